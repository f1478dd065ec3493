// AutoencoderKLLtxVideo (decoder) on B200: host-side sequencing (see vae.h).
//
// Data layout in HBM: activations are channels-last bf16.  Level l of the decoder works on a volume
// (T_l, H_l, W_l, C_l) with T_{l+1} = 2 T_l - 1, H/W doubling, C = 1024, 512, 256, 128 (vae.rs:1547-1567).
//   x / x_alt : unpadded NDHWC [T,H,W,C]          residual stream of the resnets (ping-pong)
//   hb        : unpadded NDHWC                      conv1 output of a resnet
//   p_[l]     : padded [(T+2),(H+2),(W+2),C_l]      conv input: H/W border = 0 (zero padding, vae.rs:344),
//                                                   frames 0 / T+1 = copies of frames 1 / T (replicate, vae.rs:388-411)
// A conv is then one GEMM launch over the padded-flat row space with 27 row-shifted A views (gemm.h); its epilogue
// writes bias (+ residual), the depth-to-space scatter of the upsamplers, or the unpatchified f32 pixels.
#include "vae.h"

#include <algorithm>
#include <memory>

#include <math.h>

#include "gemm.h"
#include "glue.h"
#include "options.h"
#include "vae_encoder.h"
#include "vae_glue.h"

namespace ltxv {

float* AutoencoderKLLtxVideo::vec(int64_t n) {
    storage_.emplace_back(new DevBuf());
    storage_.back()->ensure(static_cast<size_t>(n) * sizeof(float), true);
    return storage_.back()->as<float>();
}

void AutoencoderKLLtxVideo::add_plain(const std::string& key, void* dst, bool bf16, std::vector<int64_t> shape) {
    VSlot v;
    v.ps.key = key;
    v.ps.dst = dst;
    v.ps.dst_bf16 = bf16;
    v.ps.shape = std::move(shape);
    slots_[key] = std::move(v);
}

void AutoencoderKLLtxVideo::add_conv(const std::string& prefix, ConvW& cw, int Cin, int Cout, bool d2s, int rows_out) {
    cw.Cin = Cin;
    cw.Cout = Cout;
    cw.rows_out = rows_out;
    cw.d2s = d2s;
    storage_.emplace_back(new DevBuf());
    storage_.back()->ensure(static_cast<size_t>(rows_out) * 27 * Cin * 2, true);
    cw.w = storage_.back()->as<__nv_bfloat16>();
    cw.b = vec((rows_out + 255) / 256 * 256);
    VSlot w;
    w.ps.key = prefix + ".conv.weight";
    w.ps.dst = cw.w;
    w.ps.shape = {Cout, Cin, 3, 3, 3};
    w.kind = CONV_W;
    w.conv = &cw;
    slots_[w.ps.key] = w;
    VSlot b;
    b.ps.key = prefix + ".conv.bias";
    b.ps.dst = cw.b;
    b.ps.dst_bf16 = false;
    b.ps.shape = {Cout};
    b.kind = CONV_B;
    b.conv = &cw;
    slots_[b.ps.key] = b;
}

void AutoencoderKLLtxVideo::add_time_embedder(const std::string& prefix, TimeEmbW& te, int dim) {
    te.dim = dim;
    auto mk = [&](LinearW& l, int N, int K, const std::string& name) {
        l.N = N;
        l.K = K;
        storage_.emplace_back(new DevBuf());
        storage_.back()->ensure(static_cast<size_t>(N) * K * 2, true);
        l.w = storage_.back()->as<__nv_bfloat16>();
        l.b = vec(N);
        add_plain(prefix + "timestep_embedder." + name + ".weight", l.w, true, {N, K});
        add_plain(prefix + "timestep_embedder." + name + ".bias", l.b, false, {N});
    };
    mk(te.l1, dim, 256, "linear_1");
    mk(te.l2, dim, dim, "linear_2");
}

AutoencoderKLLtxVideo::AutoencoderKLLtxVideo(const ltxv_vae_config& cfg, int device) : cfg_(cfg), device_(device) {
    require_cuda_device(device);
    // reversed lists, upsample_factor 2 everywhere (vae.rs:1507-1567)
    ch_[0] = cfg.decoder_block_out_channels[2];
    ch_[1] = cfg.decoder_block_out_channels[2] / 2;
    ch_[2] = cfg.decoder_block_out_channels[1] / 2;
    ch_[3] = cfg.decoder_block_out_channels[0] / 2;
    for (int l = 0; l < 4; ++l)
        if (ch_[l] != 128 && ch_[l] != 256 && ch_[l] != 512 && ch_[l] != 1024)
            fail("decoder level %d has %d channels; supported widths are 128/256/512/1024", l, ch_[l]);
    if (ch_[1] * 2 != ch_[0] || ch_[2] * 2 != ch_[1] || ch_[3] * 2 != ch_[2])
        fail("decoder_block_out_channels must halve per level (got %d,%d,%d,%d)", ch_[0], ch_[1], ch_[2], ch_[3]);
    if (cfg.latent_channels % 64 != 0) fail("latent_channels must be a multiple of 64");
    if (cfg.patch_size != 4 || cfg.out_channels != 3) fail("only patch_size 4 / 3 output channels are supported");
    const std::string P = "decoder.";
    add_conv(P + "conv_in", conv_in_, cfg.latent_channels, ch_[0], false, ch_[0]);
    const char* level_prefix[4] = {"mid_block.", "up_blocks.0.", "up_blocks.1.", "up_blocks.2."};
    for (int l = 0; l < 4; ++l) {
        const int C = ch_[l];
        const int n = cfg.decoder_layers_per_block[l];
        const std::string bp = P + level_prefix[l];
        if (l > 0) add_conv(bp + "upsamplers.0.conv", ups_[l - 1], ch_[l - 1], C * 8, true, C * 8);
        res_[l].resize(n);
        sst_[l] = vec(static_cast<int64_t>(n) * 4 * C);
        for (int i = 0; i < n; ++i) {
            const std::string rp = bp + "resnets." + std::to_string(i) + ".";
            add_conv(rp + "conv1", res_[l][i].conv1, C, C, false, C);
            add_conv(rp + "conv2", res_[l][i].conv2, C, C, false, C);
            add_plain(rp + "scale_shift_table", sst_[l] + static_cast<int64_t>(i) * 4 * C, false, {4, C});
        }
        add_time_embedder(bp + "time_embedder.", te_[l], 4 * C);
    }
    const int C3 = ch_[3];
    add_conv(P + "conv_out", conv_out_, C3, cfg.out_channels * 16, false, 64);
    add_time_embedder(P + "time_embedder.", te_final_, 2 * C3);
    sst_final_ = vec(2 * C3);
    add_plain(P + "scale_shift_table", sst_final_, false, {2, C3});
    tsm_ = vec(1);
    add_plain(P + "timestep_scale_multiplier", tsm_, false, {});
    latents_mean_ = vec(cfg.latent_channels);
    latents_std_ = vec(cfg.latent_channels);
    fill_const(latents_std_, false, cfg.latent_channels, 1.0f);  // config default (vae.rs:91-92)
    fill_const(tsm_, false, 1, 1.0f);
    add_plain("latents_mean", latents_mean_, false, {cfg.latent_channels});
    add_plain("latents_std", latents_std_, false, {cfg.latent_channels});
    LTXV_CUDA(cudaDeviceSynchronize());
}

AutoencoderKLLtxVideo::~AutoencoderKLLtxVideo() = default;

bool AutoencoderKLLtxVideo::has_key(const std::string& key) const {
    return slots_.count(key) != 0 || (enc_ && enc_->has_key(key));
}

void AutoencoderKLLtxVideo::enable_encoder(const ltxv_vae_encoder_config& cfg) {
    if (cfg.latent_channels != cfg_.latent_channels)
        fail("encoder latent_channels %d != decoder latent_channels %d", cfg.latent_channels, cfg_.latent_channels);
    enc_.reset(new LtxVideoEncoder3d(cfg, device_));
    finalized_ = false;
}

void AutoencoderKLLtxVideo::load_tensor(const std::string& key, const void* data, int dtype, const int64_t* shape,
                                        int rank) {
    auto it = slots_.find(key);
    if (it == slots_.end() && enc_ && key.rfind("encoder.", 0) == 0) {
        enc_->load_tensor(key, data, dtype, shape, rank);
        return;
    }
    if (it == slots_.end()) {
        // The reference also builds an encoder / quant convs the t2v path never runs (vae.rs:1772-1808): ignore.
        if (key.rfind("encoder.", 0) == 0 || key.rfind("quant_conv", 0) == 0 || key.rfind("post_quant_conv", 0) == 0)
            return;
        fail("unknown VAE tensor key '%s'", key.c_str());
    }
    VSlot& v = it->second;
    bool ok = rank == static_cast<int>(v.ps.shape.size());
    for (int i = 0; ok && i < rank; ++i) ok = shape[i] == v.ps.shape[i];
    if (!ok) fail("shape mismatch for VAE tensor '%s'", key.c_str());
    LTXV_CUDA(cudaSetDevice(device_));
    if (v.kind == PLAIN) {
        ingest_tensor(v.ps.dst, v.ps.dst_bf16, data, dtype, v.ps.numel());
    } else {
        DevBuf tmp;
        tmp.ensure(static_cast<size_t>(v.ps.numel()) * 4);
        ingest_tensor(tmp.p, false, data, dtype, v.ps.numel());
        const ConvW& cw = *v.conv;
        if (v.kind == CONV_W)
            LTXV_CUDA(launch_conv_weight_relayout(tmp.p, 0, cw.w, cw.Cout, cw.Cin, cw.rows_out, cw.d2s, 0));
        else
            LTXV_CUDA(launch_conv_bias_relayout(tmp.p, 0, cw.b, cw.Cout, cw.rows_out, cw.d2s, 0));
        LTXV_CUDA(cudaDeviceSynchronize());
    }
    v.ps.loaded = true;
    finalized_ = false;
}

void AutoencoderKLLtxVideo::init_random(uint64_t seed) {
    LTXV_CUDA(cudaSetDevice(device_));
    uint64_t n = 0;
    for (auto& kv : slots_) {
        VSlot& v = kv.second;
        const std::string& k = v.ps.key;
        const uint64_t sd = seed * 7919ull + (++n);
        if (k == "latents_mean" || k == "latents_std") {
            v.ps.loaded = true;
            continue;
        }
        if (k.size() >= 25 && k.compare(k.size() - 25, 25, "timestep_scale_multiplier") == 0) {
            fill_const(v.ps.dst, false, 1, 1000.0f);
        } else if (v.kind == CONV_W) {
            // generated directly in GEMM layout; the distribution is layout invariant. Padded rows stay zero.
            fill_uniform(v.ps.dst, true, static_cast<int64_t>(v.conv->Cout) * 27 * v.conv->Cin,
                         1.0f / sqrtf(27.0f * v.conv->Cin), sd);
        } else if (v.kind == CONV_B) {
            fill_uniform(v.ps.dst, false, v.conv->Cout, 0.05f, sd);
        } else if (k.size() >= 17 && k.compare(k.size() - 17, 17, "scale_shift_table") == 0) {
            fill_normal(v.ps.dst, false, v.ps.numel(), 0.f, 1.0f / sqrtf(static_cast<float>(v.ps.shape.back())), sd);
        } else if (k.size() >= 5 && k.compare(k.size() - 5, 5, ".bias") == 0) {
            fill_uniform(v.ps.dst, v.ps.dst_bf16, v.ps.numel(), 0.05f, sd);
        } else {
            fill_uniform(v.ps.dst, v.ps.dst_bf16, v.ps.numel(), 1.0f / sqrtf(static_cast<float>(v.ps.shape.back())), sd);
        }
        v.ps.loaded = true;
    }
    LTXV_CUDA(cudaDeviceSynchronize());
    if (enc_) enc_->init_random(seed);
    finalized_ = true;
}

void AutoencoderKLLtxVideo::finalize() {
    std::string missing;
    int n = 0;
    for (auto& kv : slots_) {
        const std::string& k = kv.first;
        if (kv.second.ps.loaded) continue;
        if (k == "latents_mean" || k == "latents_std") continue;  // optional (vae.rs:1827-1838)
        if (!cfg_.timestep_conditioning &&
            (k.find("time_embedder") != std::string::npos || k.find("scale_shift_table") != std::string::npos ||
             k.find("timestep_scale_multiplier") != std::string::npos))
            continue;
        // The reference tolerates missing conditioning tensors with `.ok()` (vae.rs:692,:986,:1260,:1601-1604) and
        // silently drops the conditioning; here they are required when timestep_conditioning is on.
        if (n < 8) missing += (n ? ", " : "") + k;
        ++n;
    }
    if (n) fail("%d VAE decoder tensors were never loaded (first: %s)", n, missing.c_str());
    if (enc_) enc_->finalize();
    finalized_ = true;
}

void AutoencoderKLLtxVideo::set_comm(PeerComm* comm) {
    comm_ = (comm != nullptr && comm->nranks() > 1) ? comm : nullptr;
    wsF_ = wsH_ = wsW_ = 0;  // force a workspace rebuild for the new partitioning
}

void AutoencoderKLLtxVideo::ensure_workspace(int F, int H, int W) {
    if (F == wsF_ && H == wsH_ && W == wsW_) return;
    const int N = comm_ ? comm_->nranks() : 1;
    const int r = comm_ ? comm_->rank() : 0;
    if (H < N) fail("latent height %d is smaller than the number of decode ranks %d (every slab needs a row)", H, N);
    // ragged H-slabs: H = q N + m; ranks [0, m) own q + 1 rows, the rest q (c3 / c5 latents have H = 22)
    const int q = H / N, m = H % N;
    auto rows = [&](int k) { return q + (k < m ? 1 : 0); };
    auto row0 = [&](int k) { return k * q + (k < m ? k : m); };
    const int Hmax = q + (m ? 1 : 0);
    lat_h0_ = row0(r);
    h_up_ = r > 0 ? rows(r - 1) : 0;
    h_dn_ = r < N - 1 ? rows(r + 1) : 0;
    T_[0] = F;
    Hfull_[0] = H;
    H_[0] = rows(r);  // local slab rows
    W_[0] = W;
    for (int l = 1; l < 4; ++l) {
        T_[l] = 2 * T_[l - 1] - 1;
        Hfull_[l] = 2 * Hfull_[l - 1];
        H_[l] = 2 * H_[l - 1];
        W_[l] = 2 * W_[l - 1];
    }
    size_t max_unpadded = 0;
    SlabAlloc* sa = nullptr;
    bool fresh = false;
    if (comm_ != nullptr) {
        const std::vector<int64_t> key = {static_cast<int64_t>(comm_->id()), N, F, H, W};
        auto it = slab_allocs_.find(key);
        fresh = it == slab_allocs_.end();
        sa = &slab_allocs_[key];
    }
    for (int l = 0; l < 4; ++l) {
        const size_t padded = static_cast<size_t>(T_[l] + 2) * (H_[l] + 2) * (W_[l] + 2) * ch_[l] * 2;
        if (comm_ == nullptr) {
            // geometry changed: the zero border must be re-established
            p_[l].release();
            p_[l].ensure(padded, true);
            p2_[l].release();
            p2_[l].ensure(padded, true);
            pp_[l] = 0;
        } else {
            // every rank carves the size of the LARGEST slab so that offsets stay identical across ranks
            const size_t padded_max = static_cast<size_t>(T_[l] + 2) * ((Hmax << l) + 2) * (W_[l] + 2) * ch_[l] * 2;
            for (int k = 0; k < 2; ++k) {
                if (fresh) {
                    sa->p_off[l][k] = comm_->alloc(padded_max);
                    LTXV_CUDA(cudaMemset(comm_->local(sa->p_off[l][k]), 0, padded_max));
                }
                p_off_[l][k] = sa->p_off[l][k];
            }
            pp_[l] = 0;
        }
        const size_t un = static_cast<size_t>(T_[l]) * H_[l] * W_[l] * ch_[l] * 2;
        if (un > max_unpadded) max_unpadded = un;
    }
    const size_t a0_bytes = static_cast<size_t>(F + 2) * (H_[0] + 2) * (W + 2) * cfg_.latent_channels * 2;
    if (comm_ == nullptr) {
        a0_.release();
        a0_.ensure(a0_bytes, true);
    } else {
        if (fresh) {
            const size_t a0_max = static_cast<size_t>(F + 2) * (Hmax + 2) * (W + 2) * cfg_.latent_channels * 2;
            sa->a0_off = comm_->alloc(a0_max);
            LTXV_CUDA(cudaMemset(comm_->local(sa->a0_off), 0, a0_max));
            sa->video_off = comm_->alloc(3ull * T_[3] * (4 * Hfull_[3]) * (4 * W_[3]) * 4);
        }
        a0_off_ = sa->a0_off;
        video_off_ = sa->video_off;
        // (host-synchronous on purpose: a geometry change is rare, and the neighbours' halo stores must not race the
        // zero fills above)
        LTXV_CUDA(cudaDeviceSynchronize());
        comm_->barrier(0, 0);  // every rank's halo rows are zeroed before any neighbour may write them
        LTXV_CUDA(cudaDeviceSynchronize());
    }
    xa_.ensure(max_unpadded);
    xb_.ensure(max_unpadded);
    hb_.ensure(max_unpadded);
    size_t cond = 256 + 4096 + 4096;
    for (int l = 0; l < 4; ++l) cond += static_cast<size_t>(cfg_.decoder_layers_per_block[l]) * 4 * ch_[l];
    cond += 2 * ch_[3];
    cond_.ensure(cond * 4);
    // the zero fills above ran on the legacy default stream: make them visible to whatever stream decode() is given
    LTXV_CUDA(cudaDeviceSynchronize());
    wsF_ = F;
    wsH_ = H;
    wsW_ = W;
}

bool AutoencoderKLLtxVideo::fusable(int l) const {
    const bool off = options().vae_no_fused_prep != 0;
    return !off && ch_[l] <= 256;
}

void* AutoencoderKLLtxVideo::pad_buf(int l, int k) const {
    if (comm_ != nullptr) return comm_->local(p_off_[l][k]);
    return k ? p2_[l].p : p_[l].p;
}

void* AutoencoderKLLtxVideo::prep(const void* x, int l, const float* scale, const float* shift, int do_norm,
                                  int do_silu, cudaStream_t s) {
    pp_[l] ^= 1;  // padded inputs are double-buffered: a producer never writes the volume a conv may still be reading
    if (comm_ == nullptr) {
        LTXV_CUDA(launch_vae_prep(x, pad_buf(l, pp_[l]), scale, shift, do_norm, do_silu, T_[l], H_[l], W_[l], ch_[l], s));
        return pad_buf(l, pp_[l]);
    }
    const size_t off = p_off_[l][pp_[l]];
    const int r = comm_->rank(), N = comm_->nranks();
    void* up = r > 0 ? comm_->peer(r - 1, off) : nullptr;
    void* dn = r < N - 1 ? comm_->peer(r + 1, off) : nullptr;
    LTXV_CUDA(launch_vae_prep(x, comm_->local(off), scale, shift, do_norm, do_silu, T_[l], H_[l], W_[l], ch_[l], s, up, dn,
                              1, h_up_ << l, h_dn_ << l));
    comm_->barrier(s, 0);  // halo rows from both neighbours have landed
    return comm_->local(off);
}

void AutoencoderKLLtxVideo::conv(const ConvW& cw, const void* a_padded, int T, int H, int W, int epi, void* out,
                                 const void* res, int post, cudaStream_t s, int fuse_level, const Producer* next) {
    const int Wp = W + 2, plane = (H + 2) * Wp;
    GemmOperands ops{a_padded, static_cast<int64_t>(T + 2) * plane, cw.Cin, cw.w, cw.rows_out, 27ll * cw.Cin};
    GemmParams p{};
    p.M = T * plane;
    p.N = cw.Cout;
    p.K = 27 * cw.Cin;
    p.num_k_blocks = 27 * (cw.Cin / 64);
    p.epi = epi;
    p.ldo = cw.Cout;
    p.bias = cw.b;
    p.out = out;
    p.res_bf16 = res;
    p.conv = 1;
    p.cin_blocks = cw.Cin / 64;
    p.T = T;
    p.H = H;
    p.W = W;
    p.cin = cw.Cin;
    p.a_ptr = a_padded;
    p.post_u8_scale = post;
    p.out_h0 = slab_h0_;
    p.out_h_full = slab_hfull_;
    for (int kt = 0; kt < 3; ++kt)
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) p.tap_off[(kt * 3 + kh) * 3 + kw] = kt * plane + (kh - 1) * Wp + (kw - 1);
    if (next != nullptr) {
        // the epilogue is also the producer of the next conv's input (same level, other padded buffer)
        const int l = fuse_level;
        pp_[l] ^= 1;
        p.epi = EPI_CONV_NORM_PAD;
        p.norm_out = pad_buf(l, pp_[l]);
        p.norm_scale = next->scale;
        p.norm_shift = next->shift;
        p.norm_do = next->do_norm;
        p.norm_silu = next->do_silu;
        p.norm_tf = 1;
        if (comm_ != nullptr) {
            const int r = comm_->rank(), N = comm_->nranks();
            p.norm_halo_up = r > 0 ? comm_->peer(r - 1, p_off_[l][pp_[l]]) : nullptr;
            p.norm_halo_dn = r < N - 1 ? comm_->peer(r + 1, p_off_[l][pp_[l]]) : nullptr;
            p.norm_halo_up_h = h_up_ << l;
            p.norm_halo_dn_h = h_dn_ << l;
        }
    }
    LTXV_CUDA(launch_gemm_bf16(ops, p, 0, s));
    if (next != nullptr && comm_ != nullptr) comm_->barrier(s, 0);  // halo rows from both neighbours have landed
}

// LtxVideoResnetBlock3d::forward (vae.rs:755-821), in == out
void AutoencoderKLLtxVideo::resnet(const ResnetW& rw, int l, const float* ss, __nv_bfloat16*& x, __nv_bfloat16*& x_alt,
                                   cudaStream_t s, const Producer* next, void** ready) {
    const int C = ch_[l], T = T_[l], H = H_[l], W = W_[l];
    // ss = [shift1, scale1, shift2, scale2] (vae.rs:734-735)
    if (!fusable(l)) {
        void* a1 = prep(x, l, ss ? ss + C : nullptr, ss ? ss : nullptr, 1, 1, s);
        conv(rw.conv1, a1, T, H, W, EPI_CONV_NDHWC, hb_.p, nullptr, 0, s);
        void* a2 = prep(hb_.p, l, ss ? ss + 3 * C : nullptr, ss ? ss + 2 * C : nullptr, 1, 1, s);
        conv(rw.conv2, a2, T, H, W, EPI_CONV_NDHWC, x_alt, x, 0, s);
        std::swap(x, x_alt);
        if (ready) *ready = nullptr;
        return;
    }
    // narrow level: conv1's epilogue produces conv2's input (norm2 / scale2 / shift2 / SiLU; its raw output is never
    // stored), conv2's epilogue stores the new residual stream AND produces the next consumer's input
    void* a1 = (ready && *ready) ? *ready : prep(x, l, ss ? ss + C : nullptr, ss ? ss : nullptr, 1, 1, s);
    Producer mid;
    mid.scale = ss ? ss + 3 * C : nullptr;
    mid.shift = ss ? ss + 2 * C : nullptr;
    conv(rw.conv1, a1, T, H, W, EPI_CONV_NDHWC, nullptr, nullptr, 0, s, l, &mid);
    void* a2 = pad_buf(l, pp_[l]);
    // conv2's epilogue can also produce the NEXT consumer's input (store x and the padded copy), but that doubles the
    // epilogue: at C = 128 it then outlasts the 11 us main loop of a tile (ncu: 1.34 -> 1.99 ms per conv for 0.33 ms of
    // saved prep), at C = 256 the main loop is 4x longer and it stays hidden (0.65 -> 0.67 ms for 0.08 ms saved).
    const bool no_conv2 = options().vae_no_fuse_conv2 != 0;
    const bool all_conv2 = options().vae_fuse_conv2 != 0;
    if (no_conv2 || (C < 256 && !all_conv2)) next = nullptr;
    conv(rw.conv2, a2, T, H, W, EPI_CONV_NDHWC, x_alt, x, 0, s, next ? l : -1, next);
    std::swap(x, x_alt);
    if (ready) *ready = next ? pad_buf(l, pp_[l]) : nullptr;
}

void AutoencoderKLLtxVideo::decode(const void* z, int z_dtype, const float* timestep_dev, int B, int F, int H, int W,
                                   void* out, int out_dtype, int postprocess, cudaStream_t s) {
    if (!finalized_) finalize();
    if (B <= 0 || F <= 0 || H <= 0 || W <= 0) fail("decode: invalid latent shape [%d,%d,%d,%d,%d]", B, cfg_.latent_channels, F, H, W);
    if (z_dtype != LTXV_F32 && z_dtype != LTXV_BF16) fail("unsupported latent dtype %d", z_dtype);
    if (out_dtype != LTXV_F32 && out_dtype != LTXV_BF16) fail("unsupported output dtype %d", out_dtype);
    LTXV_CUDA(cudaSetDevice(device_));
    ensure_workspace(F, H, W);
    const bool cond = cfg_.timestep_conditioning && timestep_dev != nullptr;
    const size_t zsz = z_dtype == LTXV_F32 ? 4 : 2;
    const int64_t z_elems = static_cast<int64_t>(cfg_.latent_channels) * F * H * W;
    const int To = T_[3], Ho = 4 * Hfull_[3], Wo = 4 * W_[3];
    const int prank = comm_ ? comm_->rank() : 0;
    const int64_t out_elems = 3ll * To * Ho * Wo;
    if (out_dtype == LTXV_BF16) out_f32_.ensure(static_cast<size_t>(out_elems) * 4);

    float* cb = cond_.as<float>();
    float* sin256 = cb;
    float* t1 = sin256 + 256;
    float* tproj = t1 + 4096;
    float* ss[4];
    {
        float* q = tproj + 4096;
        for (int l = 0; l < 4; ++l) {
            ss[l] = q;
            q += static_cast<size_t>(cfg_.decoder_layers_per_block[l]) * 4 * ch_[l];
        }
    }
    float* ssf = ss[3] + static_cast<size_t>(cfg_.decoder_layers_per_block[3]) * 4 * ch_[3];

    for (int b = 0; b < B; ++b) {
        const char* zb = static_cast<const char*>(z) + static_cast<size_t>(b) * z_elems * zsz;
        if (cond) {
            // temb * timestep_scale_multiplier (vae.rs:1669-1677) -> per-block embedders (vae.rs:998-1018, :1291-1298)
            LTXV_CUDA(launch_sinusoid(timestep_dev + b, tsm_, sin256, 1, 0, s));
            for (int l = 0; l < 4; ++l) {
                const int dim = te_[l].dim;
                LTXV_CUDA(launch_gemv(sin256, te_[l].l1.w, te_[l].l1.b, t1, dim, 256, GEMV_NONE, GEMV_SILU, s));
                LTXV_CUDA(launch_gemv(t1, te_[l].l2.w, te_[l].l2.b, tproj, dim, dim, GEMV_NONE, GEMV_NONE, s));
                LTXV_CUDA(launch_add_vec(sst_[l], tproj, ss[l], cfg_.decoder_layers_per_block[l] * dim, dim, s));
            }
            const int dim = te_final_.dim;  // 2 * C3: [shift | scale] (vae.rs:1702-1711)
            LTXV_CUDA(launch_gemv(sin256, te_final_.l1.w, te_final_.l1.b, t1, dim, 256, GEMV_NONE, GEMV_SILU, s));
            LTXV_CUDA(launch_gemv(t1, te_final_.l2.w, te_final_.l2.b, tproj, dim, dim, GEMV_NONE, GEMV_NONE, s));
            LTXV_CUDA(launch_add_vec(sst_final_, tproj, ssf, dim, dim, s));
        }
        // conv_in (vae.rs:1664)
        void* a0 = comm_ ? comm_->local(a0_off_) : a0_.p;
        LTXV_CUDA(launch_vae_input(zb, z_dtype == LTXV_BF16, a0, cfg_.latent_channels, F, H, W, lat_h0_, H_[0], s));
        __nv_bfloat16* x = xa_.as<__nv_bfloat16>();
        __nv_bfloat16* x_alt = xb_.as<__nv_bfloat16>();
        conv(conv_in_, a0, T_[0], H_[0], W_[0], EPI_CONV_NDHWC, x, nullptr, 0, s);
        void* ready = nullptr;  // padded buffer already filled by the previous conv's fused producer epilogue
        for (int l = 0; l < 4; ++l) {
            if (l > 0) {
                // LtxVideoUpsampler3d (vae.rs:1090-1169): conv on the raw x, depth-to-space + residual in the epilogue
                const int lp = l - 1;
                void* au = ready ? ready : prep(x, lp, nullptr, nullptr, 0, 0, s);
                ready = nullptr;
                conv(ups_[lp], au, T_[lp], H_[lp], W_[lp], EPI_CONV_D2S, x_alt, nullptr, 0, s);
                std::swap(x, x_alt);
            }
            const int C = ch_[l];
            const size_t n_res = res_[l].size();
            for (size_t i = 0; i < n_res; ++i) {
                // who consumes this resnet's output: the next resnet's norm1, the upsampler (raw x) or norm_out
                Producer next;
                if (i + 1 < n_res) {
                    const float* ssn = cond ? ss[l] + (i + 1) * 4 * C : nullptr;
                    next.scale = ssn ? ssn + C : nullptr;
                    next.shift = ssn;
                } else if (l < 3) {
                    next.do_norm = next.do_silu = 0;
                } else {
                    next.scale = cond ? ssf + C : nullptr;
                    next.shift = cond ? ssf : nullptr;
                }
                resnet(res_[l][i], l, cond ? ss[l] + i * 4 * C : nullptr, x, x_alt, s, fusable(l) ? &next : nullptr, &ready);
            }
        }
        // norm_out -> scale/shift -> SiLU -> conv_out -> unpatchify (vae.rs:1686-1725)
        const int C3 = ch_[3];
        void* af = ready ? ready : prep(x, 3, cond ? ssf + C3 : nullptr, cond ? ssf : nullptr, 1, 1, s);
        float* o32 = out_dtype == LTXV_F32 ? static_cast<float*>(out) + static_cast<size_t>(b) * out_elems
                                           : out_f32_.as<float>();
        if (comm_ == nullptr) {
            conv(conv_out_, af, T_[3], H_[3], W_[3], EPI_CONV_UNPATCHIFY, o32, nullptr, postprocess, s);
        } else {
            // every rank stores its pixel rows into rank 0's frame buffer; rank 0 hands the assembled video out
            slab_h0_ = 4 * (lat_h0_ << 3);
            slab_hfull_ = Ho;
            conv(conv_out_, af, T_[3], H_[3], W_[3], EPI_CONV_UNPATCHIFY, comm_->peer(0, video_off_), nullptr, postprocess, s);
            slab_h0_ = slab_hfull_ = 0;
            comm_->barrier(s, 0);
            if (prank == 0)
                LTXV_CUDA(cudaMemcpyAsync(o32, comm_->local(video_off_), static_cast<size_t>(out_elems) * 4,
                                          cudaMemcpyDeviceToDevice, s));
            comm_->barrier(s, 0);  // the frame buffer may be overwritten by the next decode only after the copy
            if (prank != 0) continue;
        }
        if (out_dtype == LTXV_BF16)
            LTXV_CUDA(launch_f32_to_bf16(o32, static_cast<__nv_bfloat16*>(out) + static_cast<size_t>(b) * out_elems,
                                         out_elems, s));
    }
}


// ------------------------------------------------------------------------------------------------
// tiled decode (compatibility mode)
// ------------------------------------------------------------------------------------------------
namespace {
struct TileBuf {
    DevBuf buf;
    int T = 0, H = 0, W = 0;
    float* p() const { return buf.as<float>(); }
    void shape(int t, int h, int w) {
        T = t;
        H = h;
        W = w;
        buf.ensure(static_cast<size_t>(3) * t * h * w * 4);
    }
};
}  // namespace

void AutoencoderKLLtxVideo::tiled_decode(const void* z, int z_dtype, const float* ts_b, int T, int H, int W, float* dst,
                                         const ltxv_vae_tiling& tp, cudaStream_t s) {
    const int sr = spatial_compression_ratio(), C = cfg_.latent_channels;
    const int esz = z_dtype == LTXV_F32 ? 4 : 2;
    const int Td = 8 * T - 7, Hs = H * sr, Ws = W * sr;
    const int tl_min_h = tp.tile_sample_min_height / sr, tl_min_w = tp.tile_sample_min_width / sr;
    const int tl_str_h = tp.tile_sample_stride_height / sr, tl_str_w = tp.tile_sample_stride_width / sr;
    if (tl_min_h < 1 || tl_min_w < 1 || tl_str_h < 1 || tl_str_w < 1)
        fail("tile sizes and strides must be at least %d pixels", sr);
    const int blend_h = std::max(tp.tile_sample_min_height - tp.tile_sample_stride_height, 0);
    const int blend_w = std::max(tp.tile_sample_min_width - tp.tile_sample_stride_width, 0);
    const int ncols = (W + tl_str_w - 1) / tl_str_w;
    std::unique_ptr<TileBuf[]> row_a(new TileBuf[ncols]), row_b(new TileBuf[ncols]);
    TileBuf* prev = row_a.get();
    TileBuf* cur = row_b.get();
    int ri = 0;
    for (int i = 0; i < H; i += tl_str_h, ++ri) {
        const int th = std::min(i + tl_min_h, H) - i;
        int cj = 0;
        for (int j = 0; j < W; j += tl_str_w, ++cj) {
            const int tw = std::min(j + tl_min_w, W) - j;
            tz_sp_.ensure(static_cast<size_t>(C) * T * th * tw * esz);
            LTXV_CUDA(launch_copy_box(z, esz, C, T, H, W, 0, i, j, tz_sp_.p, T, th, tw, 0, 0, 0, T, th, tw, s));
            TileBuf& tile = cur[cj];
            tile.shape(Td, th * sr, tw * sr);
            decode(tz_sp_.p, z_dtype, ts_b, 1, T, th, tw, tile.p(), LTXV_F32, 0, s);
            if (ri > 0)  // blend_v: seam with the (already blended) tile above, along H
                LTXV_CUDA(launch_blend_axis(prev[cj].p(), prev[cj].T, prev[cj].H, prev[cj].W, tile.p(), tile.T, tile.H,
                                            tile.W, 3, 2, blend_h, s));
            if (cj > 0)  // blend_h: seam with the (already blended) left neighbour, along W
                LTXV_CUDA(launch_blend_axis(cur[cj - 1].p(), cur[cj - 1].T, cur[cj - 1].H, cur[cj - 1].W, tile.p(), tile.T,
                                            tile.H, tile.W, 3, 3, blend_w, s));
            // keep the first stride rows / columns, cropped to the sample size (:2279-2289)
            const int y0 = ri * tp.tile_sample_stride_height, x0 = cj * tp.tile_sample_stride_width;
            const int hs = std::min(std::min(tp.tile_sample_stride_height, tile.H), Hs - y0);
            const int ws = std::min(std::min(tp.tile_sample_stride_width, tile.W), Ws - x0);
            LTXV_CUDA(launch_copy_box(tile.p(), 4, 3, tile.T, tile.H, tile.W, 0, 0, 0, dst, Td, Hs, Ws, 0, y0, x0, Td, hs,
                                      ws, s));
        }
        std::swap(prev, cur);
    }
    LTXV_CUDA(cudaStreamSynchronize(s));  // the tile buffers die with this frame
}

void AutoencoderKLLtxVideo::temporal_tiled_decode(const void* z, int z_dtype, const float* ts_b, int F, int H, int W,
                                                  float* dst, const ltxv_vae_tiling& tp, cudaStream_t s) {
    const int sr = spatial_compression_ratio(), tr = temporal_compression_ratio(), C = cfg_.latent_channels;
    const int esz = z_dtype == LTXV_F32 ? 4 : 2;
    const int Hs = H * sr, Ws = W * sr, n_sample = (F - 1) * tr + 1;
    const int tl_min_h = tp.tile_sample_min_height / sr, tl_min_w = tp.tile_sample_min_width / sr;
    const int tl_min_t = tp.tile_sample_min_num_frames / tr, tl_str_t = tp.tile_sample_stride_num_frames / tr;
    if (tl_str_t < 1) fail("tile_sample_stride_num_frames must be at least %d", tr);
    const int blend_t = std::max(tp.tile_sample_min_num_frames - tp.tile_sample_stride_num_frames, 0);
    const size_t plane = static_cast<size_t>(Hs) * Ws;
    int prev_T = 0, fo = 0, idx = 0;
    for (int i = 0; i < F && fo < n_sample; i += tl_str_t, ++idx) {
        const int Tl = std::min(i + tl_min_t + 1, F) - i;
        const int Tfull = 8 * Tl - 7;
        tz_tm_.ensure(static_cast<size_t>(C) * Tl * H * W * esz);
        LTXV_CUDA(launch_copy_box(z, esz, C, F, H, W, i, 0, 0, tz_tm_.p, Tl, H, W, 0, 0, 0, Tl, H, W, s));
        t_dec_.ensure(3 * static_cast<size_t>(Tfull) * plane * 4);
        if (tp.use_tiling && (H > tl_min_h || W > tl_min_w))
            tiled_decode(tz_tm_.p, z_dtype, ts_b, Tl, H, W, t_dec_.as<float>(), tp, s);
        else
            decode(tz_tm_.p, z_dtype, ts_b, 1, Tl, H, W, t_dec_.p, LTXV_F32, 0, s);
        // every tile but the first loses its last sample frame (:2398-2408); compact into t_cur_
        const int Td = (idx > 0 && Tfull > 1) ? Tfull - 1 : Tfull;
        t_cur_.ensure(3 * static_cast<size_t>(Td) * plane * 4);
        LTXV_CUDA(launch_copy_box(t_dec_.p, 4, 3, Tfull, Hs, Ws, 0, 0, 0, t_cur_.p, Td, Hs, Ws, 0, 0, 0, Td, Hs, Ws, s));
        int take;
        const float* src = t_cur_.as<float>();
        if (idx > 0) {
            // blend with the UNBLENDED previous tile (row[idx - 1], :2420), then keep the first `stride` frames
            t_work_.ensure(3 * static_cast<size_t>(Td) * plane * 4);
            LTXV_CUDA(cudaMemcpyAsync(t_work_.p, t_cur_.p, 3 * static_cast<size_t>(Td) * plane * 4,
                                      cudaMemcpyDeviceToDevice, s));
            LTXV_CUDA(launch_blend_axis(t_prev_.as<float>(), prev_T, Hs, Ws, t_work_.as<float>(), Td, Hs, Ws, 3, 1, blend_t, s));
            take = std::min(tp.tile_sample_stride_num_frames, Td);
            src = t_work_.as<float>();
        } else {
            take = std::min(tp.tile_sample_stride_num_frames + 1, Td);
        }
        take = std::min(take, n_sample - fo);  // final crop to (F-1)*8+1 frames (:2433)
        LTXV_CUDA(launch_copy_box(src, 4, 3, Td, Hs, Ws, 0, 0, 0, dst, n_sample, Hs, Ws, fo, 0, 0, take, Hs, Ws, s));
        fo += take;
        std::swap(t_prev_.p, t_cur_.p);
        std::swap(t_prev_.bytes, t_cur_.bytes);
        prev_T = Td;
    }
    if (fo != n_sample) fail("temporal tiling produced %d of %d frames (stride larger than the tile?)", fo, n_sample);
}

void AutoencoderKLLtxVideo::decode_z(const void* z, int z_dtype, const float* timestep_dev, int B, int F, int H, int W,
                                     void* out, int out_dtype, int postprocess, const ltxv_vae_tiling* tiling,
                                     cudaStream_t s) {
    const int sr = spatial_compression_ratio(), tr = temporal_compression_ratio();
    bool temporal = false, spatial = false;
    if (tiling != nullptr) {
        const ltxv_vae_tiling& tp = *tiling;
        if (tp.tile_sample_min_height < sr || tp.tile_sample_min_width < sr || tp.tile_sample_min_num_frames < tr ||
            tp.tile_sample_stride_height < sr || tp.tile_sample_stride_width < sr || tp.tile_sample_stride_num_frames < tr)
            fail("tile sizes / strides must be at least %d pixels and %d frames", sr, tr);
        temporal = tp.use_framewise_decoding && F > tp.tile_sample_min_num_frames / tr;
        spatial = !temporal && tp.use_tiling && (W > tp.tile_sample_min_width / sr || H > tp.tile_sample_min_height / sr);
    }
    if (!temporal && !spatial) {
        decode(z, z_dtype, timestep_dev, B, F, H, W, out, out_dtype, postprocess, s);
        return;
    }
    if (comm_ != nullptr) fail("tiled decode is a single-GPU compatibility mode: clear the communicator first");
    if (z_dtype != LTXV_F32 && z_dtype != LTXV_BF16) fail("unsupported latent dtype %d", z_dtype);
    if (out_dtype != LTXV_F32 && out_dtype != LTXV_BF16) fail("unsupported output dtype %d", out_dtype);
    const size_t zsz = z_dtype == LTXV_F32 ? 4 : 2;
    const int64_t z_elems = static_cast<int64_t>(cfg_.latent_channels) * F * H * W;
    const int64_t out_elems = 3ll * (8 * F - 7) * (sr * H) * (sr * W);
    DevBuf assembled;
    if (out_dtype == LTXV_BF16) assembled.ensure(static_cast<size_t>(out_elems) * 4);
    for (int b = 0; b < B; ++b) {
        const char* zb = static_cast<const char*>(z) + static_cast<size_t>(b) * z_elems * zsz;
        const float* ts_b = timestep_dev ? timestep_dev + b : nullptr;
        float* dst = out_dtype == LTXV_F32 ? static_cast<float*>(out) + static_cast<size_t>(b) * out_elems
                                           : assembled.as<float>();
        if (temporal) temporal_tiled_decode(zb, z_dtype, ts_b, F, H, W, dst, *tiling, s);
        else tiled_decode(zb, z_dtype, ts_b, F, H, W, dst, *tiling, s);
        if (postprocess) LTXV_CUDA(launch_postprocess(dst, dst, out_elems, s));
        if (out_dtype == LTXV_BF16)
            LTXV_CUDA(launch_f32_to_bf16(dst, static_cast<__nv_bfloat16*>(out) + static_cast<size_t>(b) * out_elems,
                                         out_elems, s));
        LTXV_CUDA(cudaStreamSynchronize(s));
    }
}

}  // namespace ltxv
