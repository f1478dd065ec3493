// Host-side interface of the tcgen05 GEMM / implicit-GEMM conv3d kernel (gemm_tcgen05.cu).
//
// One kernel family serves every dense contraction on the LTX-Video hot path:
//   * DiT projections (SURVEY §2b K1,K3,K6,K10,K11,K12,K14): C[M,N] = A[M,K] * W[N,K]^T, bf16 in, f32 accumulate
//   * VAE CausalConv3d (V1, vae.rs:415-464) as an implicit GEMM over a zero/replicate-padded NDHWC volume:
//     the 27 taps become 27 row-shifted views of the same [voxels, Cin] matrix (see DESIGN.md "conv3d").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ltxv {

enum EpiMode : int {
    EPI_STORE_BF16 = 0,       // out_bf16[row, col] = act(acc + bias)
    EPI_STORE_F32 = 1,        // out_f32 [row, col] = acc + bias
    EPI_RESIDUAL_F32 = 2,     // res_f32[row, col] += gate[col] * (acc + bias); optional bf16 copy of the new value
    EPI_CONV_NDHWC = 3,       // out_bf16[voxel, col] = acc + bias (+ residual_bf16[voxel, col]); unpadded NDHWC
    EPI_CONV_D2S = 4,         // upsampler: depth-to-space(2,2,2) + channel-tiled residual + drop frame 0
    EPI_CONV_UNPATCHIFY = 5,  // conv_out: f32 NCDHW pixels through the p=4 unpatchify index map
    // conv + the NEXT conv's input producer in one epilogue (N <= 256 so one tile holds every channel of a voxel):
    // x = bf16(acc + bias (+ residual)) is stored to `out` if non-null, then pixel-norm / (1+scale)x+shift / SiLU of x
    // goes to the zero-bordered, T-replicated padded volume `norm_out` (what vae_prep_kernel would have produced).
    EPI_CONV_NORM_PAD = 6,
    // Linear + the weight / rotation half of the q,k RMS-norm + RoPE (ltx_transformer.rs:671-678, :314-339): for the
    // leading `qk_cols` columns (q, or q|k of the fused QKV projection, each `qk_dim` wide)
    //   v = acc + bias;  ss[row] += v^2;  y = v * w[col];  (y0, y1) <- (y0 cos - y1 sin, y1 cos + y0 sin)
    // is stored as bf16; the per-row scalar rsqrt(mean(v^2) + eps) commutes with both steps and is applied by the
    // consumer (attention: folded into the softmax scale of the row for q; one pass over k for k).  The remaining
    // columns (v of the QKV projection) are a plain bf16 store.  ss goes to qk_ss[row, 64-column group]: consumers add
    // the qk_dim / 64 entries of a row in a fixed order (independent of the tile width the launcher picked).
    EPI_QKV_ROPE = 7,
};

enum ActMode : int { ACT_NONE = 0, ACT_GELU_TANH = 1 };

struct GemmParams {
    int M, N, K;       // logical problem; K is the full reduction length (27*Cin for conv)
    int num_k_blocks;  // ceil(K / 64)
    int epi;           // EpiMode
    int act;           // ActMode (EPI_STORE_BF16 only)
    int ldo;           // leading dimension (elements) of out / residual tensors

    const float* bias;  // [N] f32 or null
    const float* gate;  // [N] f32 or null (EPI_RESIDUAL_F32: null means gate = 1)
    void* out;          // bf16 or f32 depending on epi; for EPI_RESIDUAL_F32 the optional bf16 copy (may be null)
    float* res_f32;     // EPI_RESIDUAL_F32: read-modify-write residual stream
    const void* res_bf16;  // EPI_CONV_NDHWC: optional residual (same layout as out)

    // ---- implicit-GEMM conv3d ----
    int conv;           // 0 = plain GEMM
    int cin_blocks;     // Cin / 64 (k-blocks per tap)
    int T, H, W;        // output volume (= input volume, stride 1), unpadded
    int tap_off[27];    // A-row offset of tap (kt,kh,kw) relative to the output's padded-flat index
    const void* a_ptr;  // base of the padded A volume (EPI_CONV_D2S reads its residual from here)
    int cin;            // Cin
    int post_u8_scale;  // EPI_CONV_UNPATCHIFY: 1 => also apply clamp(0.5x+0.5,0,1)*255 (t2v_pipeline.rs:147-155)
    int out_h0;         // EPI_CONV_UNPATCHIFY, H-slab decode: first pixel row of this slab in the full frame
    int out_h_full;     // ... and the full frame height in pixels (0 = 4*H, single slab)

    // ---- EPI_CONV_NORM_PAD ----
    void* norm_out;            // padded bf16 [(T + tf + (tf == 1)), H+2, W+2, N] of the SAME T, H, W (must not be `a_ptr`)
    const float* norm_scale;   // f32 [N] or null
    const float* norm_shift;   // f32 [N] or null
    int norm_do, norm_silu;    // pixel norm (RMS over N, eps 1e-8) / SiLU on or off
    int norm_tf;               // replicated front frames of norm_out: 1 non-causal (+1 behind), 2 causal, 3 causal + dup
    void* norm_halo_up;        // H-slab decode: neighbour buffers receiving this slab's first / last row, or null
    void* norm_halo_dn;
    int norm_halo_up_h, norm_halo_dn_h;  // slab rows H of those neighbours (ragged slabs); 0 = same as this slab

    int raster_g;       // tile raster group (row tiles per group, see tile_mn); set by launch_gemm_bf16

    // ---- EPI_QKV_ROPE ----
    int qk_cols, qk_dim;        // treated columns (qk_dim or 2 qk_dim) and the width of q (= of k): multiples of 64
    const float* qk_w[2];       // f32 [qk_dim] norm weights of q and of k
    const float* rope_cos;      // f32 [rope_rows, qk_dim / 2] or null (no rotation: cross-attention queries)
    const float* rope_sin;
    int rope_rows;              // table row of output row m is m % rope_rows (batched CFG: both halves share the table)
    int rope_row0;              // ... + rope_row0 (sequence-parallel shard: first token of this rank)
    float* qk_ss;               // f32 [M, qk_cols / 64] sums of squares per 64-column group (see EPI_QKV_ROPE)
};

// A: [rows_a, K_a] bf16 row-major (K contiguous). B: [N, K] bf16 row-major (nn.Linear weight layout).
// For conv, A is the padded volume viewed as [(T+2)*(H+2)*(W+2), Cin] and B is [Cout, 27*Cin] with k = tap*Cin + c.
struct GemmOperands {
    const void* a;
    int64_t a_rows, a_cols;  // a_cols = row length in elements (row stride)
    const void* b;
    int64_t b_rows, b_cols;
};

// Returns cudaSuccess or the launch / tensor-map error. block_n: 0 = auto.
cudaError_t launch_gemm_bf16(const GemmOperands& ops, const GemmParams& p, int block_n, cudaStream_t stream);

// Number of kernels launched by launch_gemm_bf16 since process start (bench.py's gpu_launches evidence).
uint64_t gemm_launch_count();

}  // namespace ltxv
