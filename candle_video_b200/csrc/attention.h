// Flash-style attention for sm_100a (attention_tcgen05.cu): S = Q K^T and O += P V on tcgen05 with the
// accumulators in TMEM, online softmax in registers, K/V streamed by TMA.  Covers the DiT's self-attention
// (ltx_transformer.rs:703-711, no mask) and its cross-attention to the text tokens (:717-741, additive
// key-padding bias) with one kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ltxv {

struct AttnParams {
    // q: [B, Sq, ldq] bf16, head h occupies columns q_col0 + h*D .. +D   (same for k, v with Skv rows)
    const void* q;
    const void* k;
    const void* v;
    int64_t ldq, ldk, ldv;        // row strides in elements
    int q_col0, k_col0, v_col0;   // column offset of head 0
    void* out;                    // [B, Sq, ldo] bf16, head h at columns h*D
    int64_t ldo;
    const float* kv_bias;         // optional [B, Skv] additive bias on the scaled scores (e.g. (1-mask)*-10000)
    int B, H, Sq, Skv, D;         // D in {64, 128}
    float scale;                  // 1/sqrt(D)  (ltx_transformer.rs:686)
    // Sequence-parallel epilogue (Ulysses gather fused into the attention kernel): when out_rows_per_peer > 0, query
    // row q belongs to rank q / out_rows_per_peer and is stored at out_peer[that rank] row (q % out_rows_per_peer),
    // column out_col0 + head*D of an [out_rows_per_peer, ldo] matrix (NVLink peer store).
    void* out_peer[8];
    int out_rows_per_peer;
    int out_col0;
    // Deferred RMS-norm of the queries (producer: GEMM epilogue EPI_QKV_ROPE, gemm.h): q holds w * q (rotated) WITHOUT the
    // per-row factor rsqrt(mean(q^2) + eps); q_rscale != null -> f32 [B * Sq] factors, the scores of query row r of batch
    // b are scaled by q_rscale[b Sq + r].  The factor commutes with q k^T, so it is folded into the row's softmax scale.
    const float* q_rscale;
};

cudaError_t launch_attention(const AttnParams& p, cudaStream_t stream);
uint64_t attention_launch_count();

}  // namespace ltxv
