// Bandwidth-bound glue kernels of the DiT path and the pipeline (glue.cu): every kernel reads its input once
// with 128-bit accesses and writes its output once.  SURVEY.md §2b rows K2,K4,K5,K7,K8,K14-K18.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ltxv {

enum NormKind : int { NORM_RMS = 0, NORM_LAYER = 1 };

// out_bf16[r, :] = norm(x[r, :]) * (1 + scale) + shift          (ltx_transformer.rs:99-119, :72-79, :874-889)
// x f32 [rows, D]; scale / shift f32 [D] (null => no modulation).
cudaError_t launch_norm_modulate(const float* x, void* out_bf16, const float* scale, const float* shift, int rows,
                                 int D, float eps, int kind, cudaStream_t s);

// In-place on a bf16 matrix [rows, ld]: for the D columns starting at col0:
//   y = x * rsqrt(mean(x^2) + eps) * w ; if cos != null: interleaved-pair rotation with cos/sin [rows, D/2] f32
// (RmsNorm across all heads :570-571,:671-672 + apply_rotary_emb :314-339).
// groups > 1 (no RoPE): the same norm on `groups` column blocks of every row in ONE launch -- block g starts at column
// col0 + g * group_cols and uses the weight w + g * group_w (the per-layer cross-attention keys of the stacked text K/V).
cudaError_t launch_qk_norm_rope(void* x_bf16, int64_t ld, int col0, int rows, int D, const float* w, float eps,
                                const float* cos_t, const float* sin_t, cudaStream_t s, int groups = 1,
                                int group_cols = 0, int group_w = 0);

// cos/sin [S, D/2] f32 of the 3D video RoPE (LtxVideoRotaryPosEmbed::forward :436-524).
// coords != null: coords [S,3] f32 (seconds, pixel-y, pixel-x), divided by base (20, 2048, 2048).
// coords == null: token grid (f,h,w) from F,H,W, optionally multiplied by scale3 = (sf*pt/20, sh*p/2048, sw*p/2048).
cudaError_t launch_rope_table(const float* coords, int F, int H, int W, const float* scale3_host, int S, int D,
                              float theta, float* cos_t, float* sin_t, cudaStream_t s, int token0 = 0);

// q (cols [0,D)) and k (cols [D,2D)) of a fused projection buffer in ONE launch: same math as two launch_qk_norm_rope
// calls, the cos/sin row of a token is read once for both.
cudaError_t launch_qk_pair_norm_rope(void* x_bf16, int64_t ld, int rows, int D, const float* wq, const float* wk,
                                     float eps, const float* cos_t, const float* sin_t, cudaStream_t s,
                                     int table_rows = 0 /* rows of the cos/sin table (batch entries share it); 0 = rows */);

// Ulysses scatter fused with q/k RMS-norm + RoPE: local fused projections x [rows, 3D] (q | k | v) are normed /
// rotated (q, k) or copied (v) and written head-group-wise into the peers' [S_total, 3*D/nranks] buffers:
// columns of head group g go to dst[g] at row (row0 + r), column (which * D/nranks + col % (D/nranks)).
struct ScatterDst {
    void* p[8];
};
cudaError_t launch_qkv_norm_rope_scatter(const void* x_bf16, int rows, int D, int nranks, int row0, const float* wq,
                                         const float* wk, float eps, const float* cos_t, const float* sin_t,
                                         const ScatterDst& dst, cudaStream_t s);

// Deferred q/k RMS-norm of the fused QKV epilogue (EPI_QKV_ROPE, gemm.h).  ss [rows, ss_ld] holds ss_n <= 64 group sums
// of squares of q followed by ss_n of k.  k: x[row, col0 .. col0+D) *= rsqrt(sum_k / D + eps), in place, bf16;
// q: the factor itself goes to q_rscale[row] (the attention kernels fold it into the softmax scale).
cudaError_t launch_k_rms_scale(void* x_bf16, int64_t ld, int col0, int rows, int D, const float* ss, int ss_ld, int ss_n,
                               float eps, float* q_rscale, cudaStream_t s);
// out[row] = rsqrt(sum of ss[row, 0 .. ss_n) / D + eps)   (cross-attention queries: no pass over the tensor at all)
cudaError_t launch_row_rscale(const float* ss, int ss_ld, int ss_n, int rows, int D, float eps, float* out, cudaStream_t s);
// Ulysses scatter for that layout: x [rows, 3D] holds q (unscaled), k (unscaled), v.  q and v are copied, k is scaled
// like launch_k_rms_scale, all head-group-wise into dst[g] (layout as above); the q factors go to
// q_rscale_dst[every rank][row0 + r].
cudaError_t launch_qkv_scatter_scaled(const void* x_bf16, int rows, int D, int nranks, int row0, const float* ss,
                                      int ss_ld, int ss_n, float eps, const ScatterDst& dst,
                                      const ScatterDst& q_rscale_dst, cudaStream_t s);

// y[n] = act_out( sum_k act_in(x[k]) * W[n,k] + b[n] ),  W bf16 [N,K], x/y f32; batch rows handled by the caller.
enum GemvAct : int { GEMV_NONE = 0, GEMV_SILU = 1 };
cudaError_t launch_gemv(const float* x, const void* w_bf16, const float* bias, float* y, int N, int K, int act_in,
                        int act_out, cudaStream_t s);

// sinusoidal embedding [cos | sin], 256 wide.  style 0: DiT (ltx_transformer.rs:271-309, 1/10000^(i/128));
// style 1: VAE (vae.rs:172-198, exp(-ln(1e4)/128*i)).  round_bf16: replay `timestep.to_dtype(bf16)` (:1051).
// t is a device scalar; t_mul multiplies it first (VAE timestep_scale_multiplier).
cudaError_t launch_sinusoid(const float* t_dev, const float* t_mul_dev, float* out256, int style, int round_bf16,
                            cudaStream_t s);

// dst[i] = a[i] + b[i % period]  (small f32 vectors: scale_shift_table[L,6,D] + temb[6D])
cudaError_t launch_add_vec(const float* a, const float* b, float* dst, int n, int period, cudaStream_t s);

// bias[i] = (1 - mask[i]) * -10000  (ltx_transformer.rs:1059-1064)
cudaError_t launch_mask_bias(const float* mask, float* bias, int n, cudaStream_t s);

// f32 <-> bf16 converts
cudaError_t launch_f32_to_bf16(const float* src, void* dst, int64_t n, cudaStream_t s);
cudaError_t launch_bf16_to_f32(const void* src, float* dst, int64_t n, cudaStream_t s);

// x = x*(1-m) + orig*m  (skip_layer_mask blend, ltx_transformer.rs:1112-1123)
cudaError_t launch_blend(float* x, const float* orig, float m, int64_t n, cudaStream_t s);

// ---- pipeline glue (t2v_pipeline.rs) ----
// pack: [C,F,H,W] -> [S, C*pt*p*p];  unpack: inverse.  f32, bit-exact index maps (:474-550).
cudaError_t launch_pack_latents(const float* in, float* out, int C, int F, int H, int W, int p, int pt, cudaStream_t s);
cudaError_t launch_unpack_latents(const float* in, float* out, int C, int F, int H, int W, int p, int pt,
                                  cudaStream_t s);
// coords [S,3]: (clamp(8f-7,0,1000)*f32(1/fps), 32h, 32w)  (:798-847)
cudaError_t launch_video_coords(float* out, int F, int H, int W, int ts_ratio, int sp_ratio, int fps, cudaStream_t s);

// CFG / STG combine (+ optional std rescale) fused with the Euler update (:942-964, :227-243, scheduler.rs:576-582):
//   comb = u + g (c - u);  if rescale > 0: comb = r comb std(c)/std(comb) + (1-r) comb;  comb += s_stg (c - p)
//   latents += dt * comb ;  optional: noise_pred_out = comb
// uncond / perturbed may be null.  n = elements of one batch entry.  scratch: >= 4 doubles (device).
cudaError_t launch_guidance_euler(const float* cond, const float* uncond, const float* perturbed, float* latents,
                                  float* noise_pred_out, int64_t n, float guidance_scale, float guidance_rescale,
                                  float stg_scale, float dt, double* scratch, cudaStream_t s);

// Token-sharded form of the same update (multi-GPU): every rank first accumulates the partial sums of ITS shard
// (launch_guidance_stats -> acc4 = {sum c, sum c^2, sum comb, sum comb^2}, f64), then -- after a barrier -- the update
// kernel adds the partials of all ranks of the group (peer pointers) so the std is the one over the whole tensor.
struct StatParts {
    const double* p[8];
    int n;            // number of partial-sum blocks
    int64_t n_total;  // elements of the whole (unsharded) batch entry
};
cudaError_t launch_guidance_stats(const float* cond, const float* uncond, int64_t n, float guidance_scale, double* acc4,
                                  cudaStream_t s);
cudaError_t launch_guidance_euler_parts(const float* cond, const float* uncond, const float* perturbed, float* latents,
                                        float* noise_pred_out, int64_t n, float guidance_scale, float guidance_rescale,
                                        float stg_scale, float dt, const StatParts& parts, cudaStream_t s);

// Stochastic scheduler step with CALLER-SUPPLIED noise (scheduler.rs:557-575: x0 = x - sigma v;
// x <- (1 - sigma_next) x0 + sigma_next noise) and the decode-noise blend (t2v_pipeline.rs:1055-1062:
// x <- x (1 - s) + noise s).  The reference draws the noise from the device RNG (Tensor::randn -> cuRAND); the drop-in
// keeps that RNG on the Rust side and hands the tensor in, which is what makes the result reproducible.
cudaError_t launch_stochastic_step(float* latents, const float* v, const float* noise, float sigma, float sigma_next,
                                   int64_t n, cudaStream_t s);
cudaError_t launch_noise_blend(float* x, const float* noise, float scale, int64_t n, cudaStream_t s);

// x*std_c*(1/sf)+mean_c on [C, N] (per-channel), f32  (:573-594)
cudaError_t launch_denormalize(const float* in, float* out, const float* mean, const float* std, float inv_sf, int C,
                               int64_t n_per_c, cudaStream_t s);
// clamp(0.5x+0.5,0,1)*255  (:147-155)
cudaError_t launch_postprocess(const float* in, float* out, int64_t n, cudaStream_t s);
// f32 frames [B,3,F,H,W] in 0..255 -> u8 [B,F,H,W,3] (main.rs:653-667: permute, clamp, truncating cast)
cudaError_t launch_frames_to_u8(const float* in, uint8_t* out, int B, int F, int H, int W, cudaStream_t s);

uint64_t glue_launch_count();

}  // namespace ltxv
