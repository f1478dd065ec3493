// Bandwidth-bound kernels of the VAE decoder (vae_glue.cu).  Activations live in HBM as channels-last (NDHWC) bf16;
// every conv input is a zero-bordered (H, W) / replicate-padded (T) copy so that CausalConv3d becomes 27 row-shifted
// GEMM views (see gemm.h).  SURVEY.md §2b rows V2, V3, K18.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ltxv {

// z [C, F, H, W] (NCDHW, f32 or bf16) -> padded NDHWC bf16 [(F+2), (H+2), (W+2), C]; frames 0 / F+1 replicate
// frames 1 / F (non-causal decoder padding, vae.rs:388-411); the H/W border is left untouched (must be zero).
cudaError_t launch_vae_input(const void* z, int z_is_bf16, void* out_padded, int C, int F, int H, int W, cudaStream_t s);
// Slab variant (multi-GPU decode, H split across ranks): the local padded buffer covers global rows h0-1 .. h0+Hs of
// the H_full-row latent; halo rows that exist in the (replicated) latent are filled from it, rows outside stay zero.
cudaError_t launch_vae_input(const void* z, int z_is_bf16, void* out_padded, int C, int F, int H_full, int W, int h0,
                             int Hs, cudaStream_t s);

// x bf16 [T,H,W,C] (unpadded NDHWC) -> padded bf16 [(T+2),(H+2),(W+2),C] with optional
//   pixel norm (RMS over C, eps 1e-8, vae.rs:148-153), x*(1+scale)+shift (vae.rs:736-738), SiLU (vae.rs:161-163).
// scale/shift: f32 [C] or null.
// t_front: 1 = non-causal (decoder) padding as above; 2 = causal (encoder) padding [(T+2) frames, frames 0,1 = frame
// 2]; 3 = causal padding of cat(x[:1], x) [(T+3) frames] for the temporal downsamplers (vae.rs:383-387, :539-544).
// halo_up / halo_dn: the padded buffers of the ranks owning the slab above / below (peer memory) or null: the first /
// last local row is also stored into their bottom / top halo row (conv halo exchange fused into the producer).
// H_up / H_dn: slab rows of those neighbours when they differ from H (ragged slabs; 0 = same).
cudaError_t launch_vae_prep(const void* x, void* out_padded, const float* scale, const float* shift, int do_norm,
                            int do_silu, int T, int H, int W, int C, cudaStream_t s, void* halo_up = nullptr,
                            void* halo_dn = nullptr, int t_front = 1, int H_up = 0, int H_dn = 0);

// Conv3d weight [Cout, Cin, 3,3,3] (f32 or bf16, device) -> GEMM B matrix bf16 [rows_out, 27*Cin], k = tap*Cin + c.
// d2s_perm: output channel co = c'*8 + sub is stored at row sub*(Cout/8) + c' (upsampler, see EPI_CONV_D2S).
// rows_out >= Cout; extra rows are zero.
// cin_src > 0: the source tensor has only cin_src < Cin input channels (encoder conv_in: 48 padded to 64).
cudaError_t launch_conv_weight_relayout(const void* w, int w_is_bf16, void* out, int Cout, int Cin, int rows_out,
                                        int d2s_perm, cudaStream_t s, int cin_src = 0);
// bias [Cout] (f32 or bf16) -> f32 [rows_out] with the same row permutation / zero padding
cudaError_t launch_conv_bias_relayout(const void* b, int b_is_bf16, float* out, int Cout, int rows_out, int d2s_perm,
                                      cudaStream_t s);

// Tiled decode (vae.rs:2225-2434): copy a [C, bT, bH, bW] box between two NCDHW volumes (2- or 4-byte elements) ...
cudaError_t launch_copy_box(const void* src, int elem_bytes, int C, int sT, int sH, int sW, int st0, int sh0, int sw0,
                            void* dst, int dT, int dH, int dW, int dt0, int dh0, int dw0, int bT, int bH, int bW,
                            cudaStream_t s);
// ... and blend the leading `blend` slices of b along axis (1 T, 2 H, 3 W) with the trailing slices of a, in place
// (blend_t / blend_v / blend_h, vae.rs:1927-2006; blend = min(blend_extent, extent of a, extent of b)).
cudaError_t launch_blend_axis(const float* a, int aT, int aH, int aW, float* b, int bT, int bH, int bW, int C, int axis,
                              int blend_extent, cudaStream_t s);

// ---- encoder ----
// x [3, F, H, W] NCDHW (f32 or bf16) -> patchified (p = 4, vae.rs:1427-1445), causally padded NDHWC bf16
// [(F+2), (H/4+2), (W/4+2), 64]: channel c*16 + pw*4 + ph, channels 48..63 zero; frames 0,1 replicate frame 2.
cudaError_t launch_vae_patchify(const void* x, int x_is_bf16, void* out_padded, int F, int H, int W, cudaStream_t s);
// LtxVideoDownsampler3d tail (vae.rs:549-581): pixel-unshuffle of the conv output [To*st, Ho*sh, Wo*sw, Cc] plus the
// group-averaged pixel-unshuffle of x [To*st - (st-1), Ho*sh, Wo*sw, C] (first frame duplicated st-1 times) ->
// out [To, Ho, Wo, Cc*st*sh*sw]; all NDHWC bf16.
cudaError_t launch_vae_unshuffle_add(const void* conv, const void* x, void* out, int To, int Ho, int Wo, int st, int sh,
                                     int sw, int C, int Cc, cudaStream_t s);
// encoder conv_out rows [nvox, ld] f32 -> moments [2L, nvox] f32 (mean | replicated logvar channel, vae.rs:1462-1467)
cudaError_t launch_vae_moments(const float* h, float* out, int64_t nvox, int L, int ld, cudaStream_t s);
// normalize_latents (t2v_pipeline.rs:552-571) on [B, C, inner] f32; mean / std: device f32 [C]
cudaError_t launch_normalize_latents(const float* x, const float* mean, const float* std, float scaling_factor,
                                     float* out, int B, int C, int64_t inner, cudaStream_t s);

uint64_t vae_glue_launch_count();

}  // namespace ltxv
