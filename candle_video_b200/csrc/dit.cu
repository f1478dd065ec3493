// LtxVideoTransformer3DModel on B200: host-side sequencing of the sm_100a kernels (see dit.h).
//
// Data layout in HBM (one batch entry at a time; the reference's pipeline also runs B = 1 passes,
// t2v_pipeline.rs:869-907):
//   x      f32  [S, D]    residual stream (kept in f32: 28 layers of bf16 re-rounding would dominate the error)
//   xb     bf16 [S, D]    bf16 copy of x written by the attn1 out-projection epilogue (A operand of attn2.to_q)
//   h      bf16 [S, D]    AdaLN-modulated RMS-normed activations (A operand of QKV / FFN-in GEMMs)
//   qkv    bf16 [S, 3D]   fused self-attention projections; q,k are normed + rotated in place
//   attn   bf16 [S, D]    attention output (A operand of the out projections)
//   ff     bf16 [S, 4D]   GELU(FFN-in) (A operand of FFN-out)
//   ctx.kv bf16 [K, L*2D] cross-attention K (normed) | V of the text tokens for ALL layers (layer l = columns
//                         [l*2D, (l+1)*2D)): one GEMM against the stacked attn2.to_k/to_v weights + one norm launch
#include "dit.h"

#include <math.h>
#include <string.h>

#include "attention.h"
#include "gemm.h"
#include "glue.h"
#include "options.h"

namespace ltxv {

LtxVideoTransformer3DModel::LtxVideoTransformer3DModel(const ltxv_dit_config& cfg, int device)
    : cfg_(cfg), device_(device) {
    require_cuda_device(device);
    const int D = inner_dim();
    const int hd = cfg.attention_head_dim;
    // 64 / 128: the tcgen05 kernels (LTX-Video 2B / 13B); 8..32 (multiples of 8): CUDA-core fallback, there for the
    // reference's golden test geometries (2 x 16, 4 x 16 heads)
    if (hd != 64 && hd != 128 && !(hd >= 8 && hd <= 32 && hd % 8 == 0))
        fail("attention_head_dim must be 64, 128 or a multiple of 8 up to 32 (got %d)", hd);
    if (D % 32 != 0) fail("inner dim %d must be a multiple of 32", D);
    if (cfg.in_channels % 8 != 0 || cfg.out_channels % 32 != 0 || cfg.caption_channels % 8 != 0 ||
        cfg.cross_attention_dim % 8 != 0)
        fail("channel counts must be multiples of 8 (out_channels of 32)");
    if (cfg.num_layers <= 0) fail("num_layers must be positive");
    const int X = cfg.cross_attention_dim;
    if (X != D) fail("cross_attention_dim (%d) must equal the inner dim (%d): caption_projection feeds attn2", X, D);

    auto vec = [&](int64_t n) {
        storage_.emplace_back(new DevBuf());
        storage_.back()->ensure(n * sizeof(float), true);
        return storage_.back()->as<float>();
    };
    auto linear = [&](const std::string& key, LinearW& lin, int N, int K) {
        lin = make_linear(N, K);
        add_slot(key + ".weight", lin.w, true, {N, K});
        add_slot(key + ".bias", lin.b, false, {N});
    };
    linear("proj_in", proj_in_, D, cfg.in_channels);
    linear("time_embed.emb.timestep_embedder.linear_1", te1_, D, 256);
    linear("time_embed.emb.timestep_embedder.linear_2", te2_, D, D);
    linear("time_embed.linear", te_lin_, 6 * D, D);
    linear("caption_projection.linear_1", cap1_, D, cfg.caption_channels);
    linear("caption_projection.linear_2", cap2_, D, D);
    linear("proj_out", proj_out_, cfg.out_channels, D);
    sst_final_ = vec(2 * D);
    add_slot("scale_shift_table", sst_final_, false, {2, D});
    sst_blocks_ = vec(static_cast<int64_t>(cfg.num_layers) * 6 * D);

    // attn2.to_k | to_v of every layer stacked along N ([L*2D, X]) and the attn2.norm_k weights along [L*D]: the text
    // K/V of all layers come out of ONE GEMM (M = text rows) + ONE norm launch instead of 2 x L launches of 16 + 6 us
    kv2_all_ = make_linear(cfg.num_layers * 2 * D, X);
    norm_k2_all_ = vec(static_cast<int64_t>(cfg.num_layers) * D);
    blocks_.resize(cfg.num_layers);
    for (int i = 0; i < cfg.num_layers; ++i) {
        DitBlockW& b = blocks_[i];
        const std::string p = "transformer_blocks." + std::to_string(i) + ".";
        add_slot(p + "scale_shift_table", sst_blocks_ + static_cast<int64_t>(i) * 6 * D, false, {6, D});
        // fused QKV: rows [0,D) = to_q, [D,2D) = to_k, [2D,3D) = to_v
        b.qkv1 = make_linear(3 * D, D);
        const char* names[3] = {"to_q", "to_k", "to_v"};
        for (int j = 0; j < 3; ++j) {
            add_slot(p + "attn1." + names[j] + ".weight", b.qkv1.w + static_cast<int64_t>(j) * D * D, true, {D, D});
            add_slot(p + "attn1." + names[j] + ".bias", b.qkv1.b + j * D, false, {D});
        }
        linear(p + "attn1.to_out.0", b.out1, D, D);
        b.norm_q1 = vec(D);
        b.norm_k1 = vec(D);
        add_slot(p + "attn1.norm_q.weight", b.norm_q1, false, {D});
        add_slot(p + "attn1.norm_k.weight", b.norm_k1, false, {D});
        linear(p + "attn2.to_q", b.q2, D, D);
        b.kv2.N = 2 * D;
        b.kv2.K = X;
        b.kv2.w = kv2_all_.w + static_cast<int64_t>(i) * 2 * D * X;
        b.kv2.b = kv2_all_.b + static_cast<int64_t>(i) * 2 * D;
        add_slot(p + "attn2.to_k.weight", b.kv2.w, true, {D, X});
        add_slot(p + "attn2.to_k.bias", b.kv2.b, false, {D});
        add_slot(p + "attn2.to_v.weight", b.kv2.w + static_cast<int64_t>(D) * X, true, {D, X});
        add_slot(p + "attn2.to_v.bias", b.kv2.b + D, false, {D});
        linear(p + "attn2.to_out.0", b.out2, D, D);
        b.norm_q2 = vec(D);
        b.norm_k2 = norm_k2_all_ + static_cast<int64_t>(i) * D;
        add_slot(p + "attn2.norm_q.weight", b.norm_q2, false, {D});
        add_slot(p + "attn2.norm_k.weight", b.norm_k2, false, {D});
        linear(p + "ff.net.0.proj", b.ff1, 4 * D, D);
        linear(p + "ff.net.2", b.ff2, D, 4 * D);
    }
}

LtxVideoTransformer3DModel::~LtxVideoTransformer3DModel() = default;

LinearW LtxVideoTransformer3DModel::make_linear(int N, int K) {
    LinearW l;
    l.N = N;
    l.K = K;
    storage_.emplace_back(new DevBuf());
    storage_.back()->ensure(static_cast<size_t>(N) * K * 2, true);
    l.w = storage_.back()->as<__nv_bfloat16>();
    storage_.emplace_back(new DevBuf());
    // bias padded to a multiple of 256 so the GEMM epilogue's 32-wide bias loads never leave the allocation
    storage_.back()->ensure(static_cast<size_t>((N + 255) / 256 * 256) * 4, true);
    l.b = storage_.back()->as<float>();
    return l;
}

void LtxVideoTransformer3DModel::add_slot(const std::string& key, void* dst, bool bf16, std::vector<int64_t> shape) {
    ParamSlot s;
    s.key = key;
    s.dst = dst;
    s.dst_bf16 = bf16;
    s.shape = std::move(shape);
    slots_[key] = std::move(s);
}

void LtxVideoTransformer3DModel::load_tensor(const std::string& key, const void* data, int dtype, const int64_t* shape,
                                             int rank) {
    auto it = slots_.find(key);
    if (it == slots_.end()) fail("unknown transformer tensor key '%s'", key.c_str());
    ParamSlot& s = it->second;
    bool ok = rank == static_cast<int>(s.shape.size());
    for (int i = 0; ok && i < rank; ++i) ok = shape[i] == s.shape[i];
    if (!ok) {
        std::string got = "[", want = "[";
        for (int i = 0; i < rank; ++i) got += std::to_string(shape[i]) + (i + 1 < rank ? "," : "");
        for (size_t i = 0; i < s.shape.size(); ++i) want += std::to_string(s.shape[i]) + (i + 1 < s.shape.size() ? "," : "");
        fail("shape mismatch for '%s': got %s], expected %s]", key.c_str(), got.c_str(), want.c_str());
    }
    LTXV_CUDA(cudaSetDevice(device_));
    ingest_tensor(s.dst, s.dst_bf16, data, dtype, s.numel());
    s.loaded = true;
    finalized_ = false;
    for (auto& c : ctx_) c.valid = false;
}

void LtxVideoTransformer3DModel::init_random(uint64_t seed) {
    LTXV_CUDA(cudaSetDevice(device_));
    uint64_t n = 0;
    for (auto& kv : slots_) {
        ParamSlot& s = kv.second;
        const uint64_t sd = seed * 1000003ull + (++n);
        const std::string& k = s.key;
        if (k.size() >= 17 && k.compare(k.size() - 17, 17, "scale_shift_table") == 0) {
            fill_normal(s.dst, s.dst_bf16, s.numel(), 0.f, 1.0f / sqrtf(static_cast<float>(s.shape.back())), sd);
        } else if (k.find("norm_q") != std::string::npos || k.find("norm_k") != std::string::npos) {
            fill_normal(s.dst, s.dst_bf16, s.numel(), 1.0f, 0.1f, sd);
        } else if (k.size() >= 5 && k.compare(k.size() - 5, 5, ".bias") == 0) {
            fill_uniform(s.dst, s.dst_bf16, s.numel(), 0.05f, sd);
        } else {
            fill_uniform(s.dst, s.dst_bf16, s.numel(), 1.0f / sqrtf(static_cast<float>(s.shape.back())), sd);
        }
        s.loaded = true;
    }
    LTXV_CUDA(cudaDeviceSynchronize());
    for (auto& c : ctx_) c.valid = false;
    finalized_ = true;
}

void LtxVideoTransformer3DModel::finalize() {
    std::string missing;
    int n = 0;
    for (auto& kv : slots_)
        if (!kv.second.loaded) {
            if (n < 8) missing += (n ? ", " : "") + kv.first;
            ++n;
        }
    if (n) fail("%d transformer tensors were never loaded (first: %s)", n, missing.c_str());
    finalized_ = true;
}

void LtxVideoTransformer3DModel::set_skip_block_list(const int32_t* idx, int n) {
    skip_blocks_.assign(idx, idx + (n > 0 ? n : 0));
}

void LtxVideoTransformer3DModel::set_comm(PeerComm* comm, int first, int count) {
    if (comm == nullptr || count <= 1) {
        comm_ = nullptr;
        sp_first_ = 0;
        sp_count_ = 1;
        return;
    }
    if (first < 0 || first + count > comm->nranks() || comm->rank() < first || comm->rank() >= first + count)
        fail("sequence-parallel group [%d,%d) does not contain rank %d", first, first + count, comm->rank());
    if (cfg_.num_attention_heads % count != 0)
        fail("num_attention_heads (%d) must be divisible by the sequence-parallel size %d", cfg_.num_attention_heads, count);
    comm_ = comm;
    sp_first_ = first;
    sp_count_ = count;
    if (comm->id() != sp_alloc_comm_ || count != sp_alloc_count_) sp_S_ = 0;  // otherwise keep the heap buffers
    sp_alloc_comm_ = comm->id();
    sp_alloc_count_ = count;
}

void LtxVideoTransformer3DModel::ensure_workspace(int S) {
    const int D = inner_dim();
    const int L = cfg_.num_layers;
    const size_t sd = static_cast<size_t>(S) * D;
    if (S > ws_S_) {
        x_.ensure(sd * 4);
        xb_.ensure(sd * 2);
        h_.ensure(sd * 2);
        qkv_.ensure(sd * 3 * 2);
        attn_.ensure(sd * 2);
        q2_.ensure(sd * 2);
        ff_.ensure(sd * 4 * 2);
        a_in_.ensure(static_cast<size_t>(S) * cfg_.in_channels * 2);
        cos_.ensure(sd / 2 * 4);
        sin_.ensure(sd / 2 * 4);
        out_f32_.ensure(static_cast<size_t>(S) * cfg_.out_channels * 4);
        qk_ss_.ensure(static_cast<size_t>(S) * (2 * D / 64) * 4);  // sums of squares of q | k per 64-column group
        q2_ss_.ensure(static_cast<size_t>(S) * (D / 64) * 4);
        q_rs_.ensure(static_cast<size_t>(S) * 4);  // per-row factors of the deferred q norms (self / cross)
        q2_rs_.ensure(static_cast<size_t>(S) * 4);
        ws_S_ = S;
    }
    small_.ensure((256 + 2 * static_cast<size_t>(D) + 6 * D + static_cast<size_t>(L) * 6 * D + 2 * D) * 4);
    if (comm_ != nullptr && S != sp_S_) {
        // a new shard size carves fresh buffers (the symmetric heap is a bump allocator; all ranks of the group take
        // this branch in the same call, so offsets stay identical across ranks); a (communicator, group size, shard
        // size) seen before reuses its carve-out, so alternating resolutions do not exhaust the heap
        const std::vector<int64_t> key = {static_cast<int64_t>(comm_->id()), sp_count_, S};
        auto it = sp_allocs_.find(key);
        if (it == sp_allocs_.end()) {
            const size_t Dg = static_cast<size_t>(D) / sp_count_;
            const size_t q = comm_->alloc(static_cast<size_t>(S) * sp_count_ * 3 * Dg * 2);
            const size_t a = comm_->alloc(static_cast<size_t>(S) * D * 2);
            const size_t r = comm_->alloc(static_cast<size_t>(S) * sp_count_ * 4);  // q row sums of ALL tokens
            it = sp_allocs_.emplace(key, std::vector<size_t>{q, a, r}).first;
        }
        sp_qkv_off_ = it->second[0];
        sp_attn_off_ = it->second[1];
        sp_qss_off_ = it->second[2];
        sp_S_ = S;
    }
}

void LtxVideoTransformer3DModel::gemm(const void* a, int64_t a_rows, const LinearW& lin, int M, int epi, int act,
                                      void* out, float* res, const float* gate, cudaStream_t s) {
    GemmOperands ops{a, a_rows, lin.K, lin.w, lin.N, lin.K};
    GemmParams p{};
    p.M = M;
    p.N = lin.N;
    p.K = lin.K;
    p.num_k_blocks = (lin.K + 63) / 64;
    p.epi = epi;
    p.act = act;
    p.ldo = lin.N;
    p.bias = lin.b;
    p.gate = gate;
    p.out = out;
    p.res_f32 = res;
    LTXV_CUDA(launch_gemm_bf16(ops, p, 0, s));
}

// Linear whose epilogue also applies the norm weight (and RoPE) of the q / k RMS-norm and emits the rows' sums of
// squares (EPI_QKV_ROPE, gemm.h); the per-row rsqrt is applied by the consumers.
void LtxVideoTransformer3DModel::gemm_qk(const void* a, const LinearW& lin, int M, int qk_cols, const float* wq,
                                         const float* wk, const float* cos_t, const float* sin_t, int rope_rows,
                                         int rope_row0, void* out, float* ss, cudaStream_t s) {
    GemmOperands ops{a, M, lin.K, lin.w, lin.N, lin.K};
    GemmParams p{};
    p.M = M;
    p.N = lin.N;
    p.K = lin.K;
    p.num_k_blocks = (lin.K + 63) / 64;
    p.epi = EPI_QKV_ROPE;
    p.ldo = lin.N;
    p.bias = lin.b;
    p.out = out;
    p.qk_cols = qk_cols;
    p.qk_dim = inner_dim();
    p.qk_w[0] = wq;
    p.qk_w[1] = wk;
    p.rope_cos = cos_t;
    p.rope_sin = sin_t;
    p.rope_rows = rope_rows;
    p.rope_row0 = rope_row0;
    p.qk_ss = ss;
    LTXV_CUDA(launch_gemm_bf16(ops, p, 0, s));
}

// caption projection + per-layer cross K/V (ltx_transformer.rs:1056, :667-672 for attn2)
void LtxVideoTransformer3DModel::prepare_context(int slot, const void* enc, int enc_dtype, const float* mask, int K,
                                                 cudaStream_t s) {
    if (!finalized_) finalize();
    if (slot < 0 || slot > kNumSlots) fail("context slot %d out of range", slot);
    if (K <= 0) fail("text length K must be positive");
    LTXV_CUDA(cudaSetDevice(device_));
    const int D = inner_dim();
    const int L = cfg_.num_layers;
    DitContext& c = ctx_[slot];
    c.valid = false;
    c.K = K;
    const void* enc_b = enc;
    if (enc_dtype == LTXV_F32) {
        enc_bf16_.ensure(static_cast<size_t>(K) * cfg_.caption_channels * 2);
        LTXV_CUDA(launch_f32_to_bf16(static_cast<const float*>(enc), enc_bf16_.p,
                                     static_cast<int64_t>(K) * cfg_.caption_channels, s));
        enc_b = enc_bf16_.p;
    } else if (enc_dtype != LTXV_BF16) {
        fail("unsupported encoder_hidden_states dtype %d", enc_dtype);
    }
    cap_mid_.ensure(static_cast<size_t>(K) * D * 2);
    enc_proj_.ensure(static_cast<size_t>(K) * D * 2);
    c.kv.ensure(static_cast<size_t>(L) * K * 2 * D * 2);
    gemm(enc_b, K, cap1_, K, EPI_STORE_BF16, ACT_GELU_TANH, cap_mid_.p, nullptr, nullptr, s);
    gemm(cap_mid_.p, K, cap2_, K, EPI_STORE_BF16, ACT_NONE, enc_proj_.p, nullptr, nullptr, s);
    gemm(enc_proj_.p, K, kv2_all_, K, EPI_STORE_BF16, ACT_NONE, c.kv.p, nullptr, nullptr, s);  // [K, L*2D]
    LTXV_CUDA(launch_qk_norm_rope(c.kv.p, static_cast<int64_t>(L) * 2 * D, 0, K, D, norm_k2_all_, 1e-5f, nullptr, nullptr, s,
                                  L, 2 * D, D));
    c.has_mask = mask != nullptr;
    if (mask != nullptr) {
        // bias = (1 - mask) * -10000  (:1059-1064); computed with the affine kernel semantics
        c.mask_bias.ensure(static_cast<size_t>(K) * 4);
        LTXV_CUDA(launch_mask_bias(mask, c.mask_bias.as<float>(), K, s));
    }
    c.valid = true;
}

void LtxVideoTransformer3DModel::forward_ctx(int slot, const void* hidden, int hidden_dtype, const float* timestep_dev,
                                             int S, int F, int H, int W, const float* rope_scale3_host,
                                             const float* video_coords, const float* skip_mask, int skip_mask_stride,
                                             void* out, int out_dtype, cudaStream_t s) {
    if (!finalized_) finalize();
    if (slot < 0 || slot > kNumSlots || !ctx_[slot].valid) fail("context slot %d has not been prepared", slot);
    forward_impl(&ctx_[slot], 1, hidden, hidden_dtype, timestep_dev, S, F, H, W, rope_scale3_host, video_coords, skip_mask,
                 skip_mask_stride, out, out_dtype, s);
}

void LtxVideoTransformer3DModel::prepare_pair(int slot_a, int slot_b, cudaStream_t s) {
    if (slot_a < 0 || slot_a > kNumSlots || slot_b < 0 || slot_b > kNumSlots || !ctx_[slot_a].valid || !ctx_[slot_b].valid)
        fail("both context slots of a pair must have been prepared");
    const DitContext& a = ctx_[slot_a];
    const DitContext& b = ctx_[slot_b];
    if (a.K != b.K) fail("the two contexts of a pair must have the same text length (%d vs %d)", a.K, b.K);
    LTXV_CUDA(cudaSetDevice(device_));
    const int D = inner_dim(), L = cfg_.num_layers, K = a.K;
    const size_t per = static_cast<size_t>(K) * L * 2 * D * 2;  // bytes of one context's [K, L*2D] bf16 block
    pair_kv_.ensure(2 * per);                                   // [2][K, L*2D]: batch stride = K rows, as the attention kernel expects
    LTXV_CUDA(cudaMemcpyAsync(pair_kv_.p, a.kv.p, per, cudaMemcpyDeviceToDevice, s));
    LTXV_CUDA(cudaMemcpyAsync(static_cast<char*>(pair_kv_.p) + per, b.kv.p, per, cudaMemcpyDeviceToDevice, s));
    pair_has_mask_ = a.has_mask || b.has_mask;
    if (pair_has_mask_) {
        pair_bias_.ensure(static_cast<size_t>(2) * K * 4);
        const DitContext* c2[2] = {&a, &b};
        for (int i = 0; i < 2; ++i) {
            float* dst = pair_bias_.as<float>() + static_cast<size_t>(i) * K;
            if (c2[i]->has_mask) LTXV_CUDA(cudaMemcpyAsync(dst, c2[i]->mask_bias.p, K * 4, cudaMemcpyDeviceToDevice, s));
            else LTXV_CUDA(cudaMemsetAsync(dst, 0, K * 4, s));
        }
    }
    pair_K_ = K;
    pair_valid_ = true;
}

void LtxVideoTransformer3DModel::forward_pair(const void* hidden, int hidden_dtype, const float* timestep_dev, int S,
                                              int F, int H, int W, const float* rope_scale3_host,
                                              const float* video_coords, void* out, int out_dtype, cudaStream_t s) {
    if (!finalized_) finalize();
    if (!pair_valid_) fail("prepare_pair has not been called");
    if (comm_ != nullptr) fail("the batched CFG pair forward is a single-GPU path (the ranks split the branches instead)");
    forward_impl(nullptr, 2, hidden, hidden_dtype, timestep_dev, S, F, H, W, rope_scale3_host, video_coords, nullptr, 0, out,
                 out_dtype, s);
}

// nb = 1: one batch entry against `ctx1`.  nb = 2: the CFG pair (pair_kv_ / pair_bias_); the two entries share hidden
// states, timestep (hence the AdaLN modulation) and the RoPE table; all token-local kernels simply see M = 2S rows.
void LtxVideoTransformer3DModel::forward_impl(const DitContext* ctx1, int nb, const void* hidden, int hidden_dtype,
                                              const float* timestep_dev, int S, int F, int H, int W,
                                              const float* rope_scale3_host, const float* video_coords,
                                              const float* skip_mask, int skip_mask_stride, void* out, int out_dtype,
                                              cudaStream_t s) {
    if (S <= 0) fail("sequence length must be positive");
    if (video_coords == nullptr && static_cast<int64_t>(F) * H * W != static_cast<int64_t>(S) * sp_count_)
        fail("num_frames*height*width (%d*%d*%d) must equal the sequence length %d when video_coords is not given", F, H,
             W, S * sp_count_);
    if (out_dtype != LTXV_F32 && out_dtype != LTXV_BF16) fail("unsupported output dtype %d", out_dtype);
    LTXV_CUDA(cudaSetDevice(device_));
    const int D = inner_dim();
    const int L = cfg_.num_layers;
    const int heads = cfg_.num_attention_heads, hd = cfg_.attention_head_dim;
    const int M = nb * S;  // rows of every token-local kernel
    const int ctxK = nb == 1 ? ctx1->K : pair_K_;
    ensure_workspace(M);

    float* sm = small_.as<float>();
    float* tp = sm;                    // [256]
    float* t1 = tp + 256;              // [D]
    float* e = t1 + D;                 // [D]   embedded_timestep
    float* temb = e + D;               // [6D]
    float* ada = temb + 6 * D;         // [L, 6, D]
    float* fin = ada + static_cast<size_t>(L) * 6 * D;  // [2, D]

    // ---- timestep path: AdaLayerNormSingle (:262-267) ----
    LTXV_CUDA(launch_sinusoid(timestep_dev, nullptr, tp, 0, cfg_.timestep_bf16_round, s));
    LTXV_CUDA(launch_gemv(tp, te1_.w, te1_.b, t1, D, 256, GEMV_NONE, GEMV_SILU, s));
    LTXV_CUDA(launch_gemv(t1, te2_.w, te2_.b, e, D, D, GEMV_NONE, GEMV_NONE, s));
    LTXV_CUDA(launch_gemv(e, te_lin_.w, te_lin_.b, temb, 6 * D, D, GEMV_SILU, GEMV_NONE, s));
    LTXV_CUDA(launch_add_vec(sst_blocks_, temb, ada, L * 6 * D, 6 * D, s));  // :847-854
    LTXV_CUDA(launch_add_vec(sst_final_, e, fin, 2 * D, D, s));              // :1131-1147

    // ---- RoPE table (:1073-1080) ----
    float scale3[3];
    const float* sc = nullptr;
    if (video_coords == nullptr && rope_scale3_host != nullptr) {
        // :410-412  (s * patch / base) as f32
        scale3[0] = static_cast<float>(static_cast<double>(rope_scale3_host[0]) * cfg_.patch_size_t / 20.0);
        scale3[1] = static_cast<float>(static_cast<double>(rope_scale3_host[1]) * cfg_.patch_size / 2048.0);
        scale3[2] = static_cast<float>(static_cast<double>(rope_scale3_host[2]) * cfg_.patch_size / 2048.0);
        sc = scale3;
    }
    const bool sp = comm_ != nullptr;
    const int spn = sp_count_, spr = sp_rank();
    LTXV_CUDA(launch_rope_table(video_coords, F, H, W, sc, S, D, 10000.0f, cos_.as<float>(), sin_.as<float>(), s,
                                sp ? spr * S : 0));

    // ---- proj_in (:1049) ----
    const void* a_in = hidden;
    if (hidden_dtype == LTXV_F32) {
        LTXV_CUDA(launch_f32_to_bf16(static_cast<const float*>(hidden), a_in_.p,
                                     static_cast<int64_t>(S) * cfg_.in_channels, s));
        a_in = a_in_.p;
    } else if (hidden_dtype != LTXV_BF16) {
        fail("unsupported hidden_states dtype %d", hidden_dtype);
    }
    float* x = x_.as<float>();
    for (int i = 0; i < nb; ++i)  // the pair shares its hidden states: same projection into both halves of x
        gemm(a_in, S, proj_in_, S, EPI_STORE_F32, ACT_NONE, x + static_cast<size_t>(i) * S * D, nullptr, nullptr, s);

    const float attn_scale = 1.0f / sqrtf(static_cast<float>(hd));
    for (int l = 0; l < L; ++l) {
        bool skip = false;
        for (int sidx : skip_blocks_) skip = skip || (sidx == l);  // :1094-1096
        if (skip) continue;
        float m = 0.f;
        if (skip_mask != nullptr) m = skip_mask[static_cast<size_t>(l) * skip_mask_stride];
        if (m == 1.0f) continue;  // x*(1-1) + orig*1 == orig  (:1112-1123)
        const bool blend = (m != 0.0f);
        if (blend) {
            orig_.ensure(static_cast<size_t>(M) * D * 4);
            LTXV_CUDA(cudaMemcpyAsync(orig_.p, x, static_cast<size_t>(M) * D * 4, cudaMemcpyDeviceToDevice, s));
        }
        const DitBlockW& b = blocks_[l];
        const float* a6 = ada + static_cast<size_t>(l) * 6 * D;  // shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp

        // --- self-attention ---
        LTXV_CUDA(launch_norm_modulate(x, h_.p, a6 + 1 * D, a6 + 0 * D, M, D, cfg_.norm_eps, NORM_RMS, s));
        // fused path: the QKV epilogue applies norm weight + RoPE and emits sum(q^2), sum(k^2) per row; the row factors
        // are applied by one pass over k and inside the attention kernel for q (ltx_transformer.rs:671-678)
        const bool fused_qk = !options().qk_unfused && D % 64 == 0 && D / 64 <= 64;
        const int ssn = D / 64;  // 64-column groups of q (= of k)
        if (fused_qk)
            gemm_qk(h_.p, b.qkv1, M, 2 * D, b.norm_q1, b.norm_k1, cos_.as<float>(), sin_.as<float>(), S, 0, qkv_.p,
                    qk_ss_.as<float>(), s);
        else
            gemm(h_.p, M, b.qkv1, M, EPI_STORE_BF16, ACT_NONE, qkv_.p, nullptr, nullptr, s);
        const void* attn1_out = attn_.p;
        if (!sp) {
            if (fused_qk)
                LTXV_CUDA(launch_k_rms_scale(qkv_.p, 3 * D, D, M, D, qk_ss_.as<float>(), 2 * ssn, ssn, 1e-5f,
                                             q_rs_.as<float>(), s));
            else
                LTXV_CUDA(launch_qk_pair_norm_rope(qkv_.p, 3 * D, M, D, b.norm_q1, b.norm_k1, 1e-5f, cos_.as<float>(),
                                                   sin_.as<float>(), s, S));
            AttnParams ap{};
            ap.q = ap.k = ap.v = qkv_.p;
            ap.ldq = ap.ldk = ap.ldv = 3 * D;
            ap.q_col0 = 0;
            ap.k_col0 = D;
            ap.v_col0 = 2 * D;
            ap.out = attn_.p;
            ap.ldo = D;
            ap.kv_bias = nullptr;
            ap.B = nb;
            ap.H = heads;
            ap.Sq = S;
            ap.Skv = S;
            ap.D = hd;
            ap.scale = attn_scale;
            if (fused_qk) ap.q_rscale = q_rs_.as<float>();
            LTXV_CUDA(launch_attention(ap, s));
        } else {
            // Ulysses: tokens -> heads.  The norm+RoPE (or, on the fused path, the k-scale) kernel stores each head group
            // straight into the rank that owns those heads; the attention epilogue stores each query block straight into
            // the rank that owns those tokens.
            const int Dg = D / spn;
            ScatterDst dst{};
            for (int g = 0; g < spn; ++g) dst.p[g] = comm_->peer(sp_first_ + g, sp_qkv_off_);
            if (fused_qk) {
                ScatterDst qdst{};
                for (int g = 0; g < spn; ++g) qdst.p[g] = comm_->peer(sp_first_ + g, sp_qss_off_);
                LTXV_CUDA(launch_qkv_scatter_scaled(qkv_.p, S, D, spn, spr * S, qk_ss_.as<float>(), 2 * ssn, ssn, 1e-5f, dst,
                                                    qdst, s));
            } else {
                LTXV_CUDA(launch_qkv_norm_rope_scatter(qkv_.p, S, D, spn, spr * S, b.norm_q1, b.norm_k1, 1e-5f,
                                                       cos_.as<float>(), sin_.as<float>(), dst, s));
            }
            comm_->barrier(s, 1, sp_first_, spn);
            AttnParams ap{};
            ap.q = ap.k = ap.v = comm_->local(sp_qkv_off_);
            ap.ldq = ap.ldk = ap.ldv = 3 * Dg;
            ap.q_col0 = 0;
            ap.k_col0 = Dg;
            ap.v_col0 = 2 * Dg;
            ap.out = nullptr;
            ap.ldo = D;
            ap.kv_bias = nullptr;
            ap.B = 1;
            ap.H = heads / spn;
            ap.Sq = S * spn;
            ap.Skv = S * spn;
            ap.D = hd;
            ap.scale = attn_scale;
            if (fused_qk) ap.q_rscale = static_cast<const float*>(comm_->local(sp_qss_off_));
            for (int g = 0; g < spn; ++g) ap.out_peer[g] = comm_->peer(sp_first_ + g, sp_attn_off_);
            ap.out_rows_per_peer = S;
            ap.out_col0 = spr * Dg;
            LTXV_CUDA(launch_attention(ap, s));
            comm_->barrier(s, 1, sp_first_, spn);
            attn1_out = comm_->local(sp_attn_off_);
        }
        gemm(attn1_out, M, b.out1, M, EPI_RESIDUAL_F32, ACT_NONE, xb_.p, x, a6 + 2 * D, s);  // x += gate_msa * attn1

        // --- cross-attention (no norm, no gate, no RoPE; :903-909) ---
        if (fused_qk) {
            gemm_qk(xb_.p, b.q2, M, D, b.norm_q2, nullptr, nullptr, nullptr, 0, 0, q2_.p, q2_ss_.as<float>(), s);
            LTXV_CUDA(launch_row_rscale(q2_ss_.as<float>(), ssn, ssn, M, D, 1e-5f, q2_rs_.as<float>(), s));
        } else {
            gemm(xb_.p, M, b.q2, M, EPI_STORE_BF16, ACT_NONE, q2_.p, nullptr, nullptr, s);
            LTXV_CUDA(launch_qk_norm_rope(q2_.p, D, 0, M, D, b.norm_q2, 1e-5f, nullptr, nullptr, s));
        }
        {
            // layer l = columns [l*2D, (l+1)*2D) of the stacked [K, L*2D] text K/V ([2][K, L*2D] for a pair)
            const __nv_bfloat16* kv = (nb == 1 ? ctx1->kv.as<__nv_bfloat16>() : pair_kv_.as<__nv_bfloat16>()) +
                                      static_cast<size_t>(l) * 2 * D;
            AttnParams ap{};
            ap.q = q2_.p;
            ap.k = ap.v = kv;
            ap.ldq = D;
            ap.ldk = ap.ldv = static_cast<int64_t>(cfg_.num_layers) * 2 * D;
            ap.q_col0 = 0;
            ap.k_col0 = 0;
            ap.v_col0 = D;
            ap.out = attn_.p;
            ap.ldo = D;
            ap.kv_bias = nb == 1 ? (ctx1->has_mask ? ctx1->mask_bias.as<float>() : nullptr)
                                 : (pair_has_mask_ ? pair_bias_.as<float>() : nullptr);
            ap.B = nb;
            ap.H = heads;
            ap.Sq = S;
            ap.Skv = ctxK;
            ap.D = hd;
            ap.scale = attn_scale;
            if (fused_qk) ap.q_rscale = q2_rs_.as<float>();
            LTXV_CUDA(launch_attention(ap, s));
        }
        gemm(attn_.p, M, b.out2, M, EPI_RESIDUAL_F32, ACT_NONE, nullptr, x, nullptr, s);  // x += attn2

        // --- feed-forward ---
        LTXV_CUDA(launch_norm_modulate(x, h_.p, a6 + 4 * D, a6 + 3 * D, M, D, cfg_.norm_eps, NORM_RMS, s));
        gemm(h_.p, M, b.ff1, M, EPI_STORE_BF16, ACT_GELU_TANH, ff_.p, nullptr, nullptr, s);
        gemm(ff_.p, M, b.ff2, M, EPI_RESIDUAL_F32, ACT_NONE, nullptr, x, a6 + 5 * D, s);  // x += gate_mlp * ff

        if (blend) LTXV_CUDA(launch_blend(x, orig_.as<float>(), m, static_cast<int64_t>(M) * D, s));
    }

    // ---- output head (:1126-1163): LayerNorm (no affine) -> (1+scale) x + shift -> proj_out ----
    LTXV_CUDA(launch_norm_modulate(x, h_.p, fin + D, fin, M, D, 1e-6f, NORM_LAYER, s));
    if (out_dtype == LTXV_F32) {
        gemm(h_.p, M, proj_out_, M, EPI_STORE_F32, ACT_NONE, out, nullptr, nullptr, s);
    } else {
        gemm(h_.p, M, proj_out_, M, EPI_STORE_BF16, ACT_NONE, out, nullptr, nullptr, s);
    }
}

void LtxVideoTransformer3DModel::forward(const void* hidden, int hidden_dtype, const void* enc, int enc_dtype,
                                         const float* timestep, const float* mask, int B, int S, int K, int F, int H,
                                         int W, const float* rope_scale3, const float* video_coords,
                                         const float* skip_layer_mask, void* out, int out_dtype, cudaStream_t s) {
    if (B <= 0) fail("batch must be positive");
    const size_t hsz = hidden_dtype == LTXV_F32 ? 4 : 2, esz = enc_dtype == LTXV_F32 ? 4 : 2,
                 osz = out_dtype == LTXV_F32 ? 4 : 2;
    for (int b = 0; b < B; ++b) {
        const char* hb = static_cast<const char*>(hidden) + static_cast<size_t>(b) * S * cfg_.in_channels * hsz;
        const char* eb = static_cast<const char*>(enc) + static_cast<size_t>(b) * K * cfg_.caption_channels * esz;
        char* ob = static_cast<char*>(out) + static_cast<size_t>(b) * S * cfg_.out_channels * osz;
        prepare_context(kScratchSlot, eb, enc_dtype, mask ? mask + static_cast<size_t>(b) * K : nullptr, K, s);
        forward_ctx(kScratchSlot, hb, hidden_dtype, timestep + b, S, F, H, W, rope_scale3,
                    video_coords ? video_coords + static_cast<size_t>(b) * S * 3 : nullptr,
                    skip_layer_mask ? skip_layer_mask + b : nullptr, B, ob, out_dtype, s);
    }
}

}  // namespace ltxv
