// C++ host-side mirror of `LtxVideoTransformer3DModel` (ltx_transformer.rs:940-1215): owns the bf16 weights and the
// forward workspace, sequences the sm_100a kernels.  The Rust trait object behind `Box<dyn VideoTransformer3D>`
// (t2v_pipeline.rs:245-251) binds to this through the C ABI in ffi.cu.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "comm.h"
#include "model_common.h"

namespace ltxv {

struct LinearW {
    __nv_bfloat16* w = nullptr;  // [N, K]
    float* b = nullptr;          // [N]
    int N = 0, K = 0;
};

struct DitBlockW {
    LinearW qkv1;  // attn1.to_q | to_k | to_v fused: [3D, D]
    LinearW out1;  // attn1.to_out.0
    LinearW q2;    // attn2.to_q
    LinearW kv2;   // attn2.to_k | to_v fused: [2D, Xd]
    LinearW out2;  // attn2.to_out.0
    LinearW ff1;   // ff.net.0.proj [4D, D]
    LinearW ff2;   // ff.net.2      [D, 4D]
    float *norm_q1 = nullptr, *norm_k1 = nullptr, *norm_q2 = nullptr, *norm_k2 = nullptr;  // [D]
};

struct DitContext {  // hoisted, step-invariant text path for one prompt
    bool valid = false;
    int K = 0;
    DevBuf kv;         // [K, L*2D] bf16: cross-attention K (normed) | V of every layer (layer l = columns [l*2D, (l+1)*2D))
    DevBuf mask_bias;  // [K] f32 additive bias, or empty when no mask
    bool has_mask = false;
};

class LtxVideoTransformer3DModel {
public:
    LtxVideoTransformer3DModel(const ltxv_dit_config& cfg, int device);
    ~LtxVideoTransformer3DModel();

    const ltxv_dit_config& config() const { return cfg_; }
    int device() const { return device_; }
    PipeWs& pipe_ws() { return pipe_ws_; }
    std::map<std::vector<int64_t>, std::vector<size_t>>& pipe_allocs() { return pipe_allocs_; }
    int inner_dim() const { return cfg_.num_attention_heads * cfg_.attention_head_dim; }

    void load_tensor(const std::string& key, const void* data, int dtype, const int64_t* shape, int rank);
    bool has_key(const std::string& key) const { return slots_.count(key) != 0; }
    void init_random(uint64_t seed);
    void finalize();
    void set_skip_block_list(const int32_t* idx, int n);  // ltx_transformer.rs:1024
    // Ulysses sequence parallelism over ranks [first, first+count) of `comm`: forward_ctx then takes this rank's
    // contiguous token shard (S = local tokens) and exchanges heads <-> tokens around self-attention through peer memory
    void set_comm(PeerComm* comm, int first, int count);
    int sp_size() const { return sp_count_; }
    int sp_rank() const { return comm_ ? comm_->rank() - sp_first_ : 0; }

    // hoisted text path (one batch entry)
    void prepare_context(int slot, const void* enc, int enc_dtype, const float* mask, int K, cudaStream_t s);
    // forward for one batch entry against a prepared context
    void forward_ctx(int slot, const void* hidden, int hidden_dtype, const float* timestep_dev, int S, int F, int H,
                     int W, const float* rope_scale3_host, const float* video_coords, const float* skip_mask_layer_host,
                     int skip_mask_stride, void* out, int out_dtype, cudaStream_t s);
    // Two batch entries that share hidden states, timestep and coordinates and differ only in the text context
    // (the CFG pair: uncond = slot_a, cond = slot_b), run as ONE forward over 2S tokens.  The reference runs them as
    // two sequential B = 1 forwards (t2v_pipeline.rs:878-907); nothing in the block mixes batch entries, so the
    // results are the same rows, but every GEMM sees 78 instead of 39 row tiles (last-wave waste 3 % instead of 10 %),
    // the weights stream once per step and the launch count halves.  out: [2, S, out_channels].
    void prepare_pair(int slot_a, int slot_b, cudaStream_t s);
    void forward_pair(const void* hidden, int hidden_dtype, const float* timestep_dev, int S, int F, int H, int W,
                      const float* rope_scale3_host, const float* video_coords, void* out, int out_dtype, cudaStream_t s);
    // reference-shaped forward (recomputes the text path every call, like ltx_transformer.rs:1056)
    void forward(const void* hidden, int hidden_dtype, const void* enc, int enc_dtype, const float* timestep,
                 const float* mask, int B, int S, int K, int F, int H, int W, const float* rope_scale3,
                 const float* video_coords, const float* skip_layer_mask, void* out, int out_dtype, cudaStream_t s);

    static constexpr int kNumSlots = 4;
    static constexpr int kScratchSlot = 4;  // used by forward()

private:
    void add_slot(const std::string& key, void* dst, bool bf16, std::vector<int64_t> shape);
    LinearW make_linear(int N, int K);
    void ensure_workspace(int S);
    void gemm(const void* a, int64_t a_rows, const LinearW& lin, int M, int epi, int act, void* out, float* res,
              const float* gate, cudaStream_t s);
    void gemm_qk(const void* a, const LinearW& lin, int M, int qk_cols, const float* wq, const float* wk, const float* cos_t,
                 const float* sin_t, int rope_rows, int rope_row0, void* out, float* ss, cudaStream_t s);

    ltxv_dit_config cfg_;
    int device_;
    bool finalized_ = false;
    std::vector<std::unique_ptr<DevBuf>> storage_;
    std::map<std::string, ParamSlot> slots_;
    std::vector<int> skip_blocks_;

    LinearW proj_in_, te1_, te2_, te_lin_, cap1_, cap2_, proj_out_;
    float* sst_final_ = nullptr;  // [2, D]
    float* sst_blocks_ = nullptr;  // [L, 6, D] contiguous
    LinearW kv2_all_;                // attn2.to_k | to_v of every layer stacked along N: [L * 2D, Xd] (blocks_[l].kv2 points in)
    float* norm_k2_all_ = nullptr;   // [L, D] attn2.norm_k weights
    std::vector<DitBlockW> blocks_;

    DitContext ctx_[kNumSlots + 1];
    // CFG pair: [2][K, L*2D] cross-attention K|V and [2][K] key bias of the two contexts, batch-major
    DevBuf pair_kv_, pair_bias_;
    int pair_K_ = 0;
    bool pair_valid_ = false, pair_has_mask_ = false;
    void forward_impl(const DitContext* ctx, int nb, const void* hidden, int hidden_dtype, const float* timestep_dev,
                      int S, int F, int H, int W, const float* rope_scale3_host, const float* video_coords,
                      const float* skip_mask, int skip_mask_stride, void* out, int out_dtype, cudaStream_t s);

    // workspace
    int ws_S_ = 0;
    DevBuf x_, xb_, h_, qkv_, attn_, q2_, ff_, a_in_, cos_, sin_, out_f32_, orig_, qk_ss_, q2_ss_, q_rs_, q2_rs_;
    DevBuf small_;  // tp[256] | t1[D] | e[D] | temb[6D] | ada[L*6D] | fin[2D]
    DevBuf enc_bf16_, cap_mid_, enc_proj_;

    // sequence parallel state
    PeerComm* comm_ = nullptr;
    int sp_first_ = 0, sp_count_ = 1;
    int sp_S_ = 0;                       // local token capacity of the symmetric buffers
    uint64_t sp_alloc_comm_ = 0;         // communicator id / group size the symmetric buffers were carved for
    int sp_alloc_count_ = 0;
    std::map<std::vector<int64_t>, std::vector<size_t>> sp_allocs_;  // (comm id, group size, S) -> (qkv, attn, q row sums) offsets
    size_t sp_qss_off_ = 0;  // heap offset of the gathered q row sums [S_total] (fused q/k epilogue path)
    PipeWs pipe_ws_;
    std::map<std::vector<int64_t>, std::vector<size_t>> pipe_allocs_;  // pipeline_denoise_parallel exchange buffers  // (comm id, group size, S) -> (qkv, attn) offsets
    size_t sp_qkv_off_ = 0, sp_attn_off_ = 0;  // heap offsets: qkv_full [S_total, 3*D/N], attn_in [S_local, D]
};

}  // namespace ltxv
