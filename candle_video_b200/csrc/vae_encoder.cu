// LtxVideoEncoder3d on B200: host-side sequencing (see vae_encoder.h).
//
// Data layout in HBM (same conventions as the decoder, vae.cu): level l works on a channels-last bf16 volume
// (T_l, H_l, W_l, C_l), C = 128, 256, 512, 1024, 2048.  Every conv input is a zero-bordered (H, W) copy with the
// CAUSAL temporal padding of the encoder (two copies of frame 0 in front, vae.rs:383-387), written by the pixel-norm
// producer kernel, so that CausalConv3d is one implicit-GEMM launch over 27 row-shifted views.  A downsampler is that
// same conv on the raw activations (first frame duplicated once more when it strides in T, vae.rs:539-544) followed
// by one bandwidth-bound kernel doing both pixel-unshuffles, the group mean and the add (vae.rs:549-581).
#include "vae_encoder.h"

#include <algorithm>
#include <memory>

#include <math.h>
#include <stdlib.h>

#include "gemm.h"
#include "options.h"
#include "vae_glue.h"

namespace ltxv {

void LtxVideoEncoder3d::add_conv(const std::string& prefix, ConvW& cw, int cin_src, int Cin, int Cout, int rows_out) {
    cw.Cin = Cin;
    cw.Cout = Cout;
    cw.rows_out = rows_out;
    cw.d2s = false;
    storage_.emplace_back(new DevBuf());
    storage_.back()->ensure(static_cast<size_t>(rows_out) * 27 * Cin * 2, true);
    cw.w = storage_.back()->as<__nv_bfloat16>();
    storage_.emplace_back(new DevBuf());
    storage_.back()->ensure(static_cast<size_t>((rows_out + 255) / 256 * 256) * sizeof(float), true);
    cw.b = storage_.back()->as<float>();
    Slot w;
    w.ps.key = prefix + ".conv.weight";
    w.ps.dst = cw.w;
    w.ps.shape = {Cout, cin_src, 3, 3, 3};
    w.kind = 1;
    w.conv = &cw;
    w.cin_src = cin_src;
    slots_[w.ps.key] = w;
    Slot b;
    b.ps.key = prefix + ".conv.bias";
    b.ps.dst = cw.b;
    b.ps.dst_bf16 = false;
    b.ps.shape = {Cout};
    b.kind = 2;
    b.conv = &cw;
    slots_[b.ps.key] = b;
}

LtxVideoEncoder3d::LtxVideoEncoder3d(const ltxv_vae_encoder_config& cfg, int device) : cfg_(cfg), device_(device) {
    require_cuda_device(device);
    if (cfg.in_channels != 3 || cfg.patch_size != 4) fail("encoder: only 3 input channels / patch_size 4 are supported");
    if (cfg.latent_channels <= 0 || cfg.latent_channels % 32 != 0) fail("encoder: latent_channels must be a multiple of 32");
    for (int l = 0; l < 5; ++l) {
        ch_[l] = cfg.block_out_channels[l];
        if (ch_[l] != 128 && ch_[l] != 256 && ch_[l] != 512 && ch_[l] != 1024 && ch_[l] != 2048)
            fail("encoder level %d has %d channels; supported widths are 128/256/512/1024/2048", l, ch_[l]);
        if (cfg.layers_per_block[l] < (l == 4 ? 1 : 0)) fail("encoder: invalid layers_per_block[%d]", l);
    }
    for (int l = 0; l < 4; ++l) {
        switch (cfg.downsample_types[l]) {  // DownsampleType::stride, vae.rs:484-493
            case LTXV_DOWN_SPATIAL: stride_[l][0] = 1; stride_[l][1] = 2; stride_[l][2] = 2; break;
            case LTXV_DOWN_TEMPORAL: stride_[l][0] = 2; stride_[l][1] = 1; stride_[l][2] = 1; break;
            case LTXV_DOWN_SPATIOTEMPORAL: stride_[l][0] = 2; stride_[l][1] = 2; stride_[l][2] = 2; break;
            default:
                fail("encoder: downsample type %d of block %d is not supported (the stride-2 conv layout of LTX-Video "
                     "0.9.0 is out of scope; pixel-unshuffle types only)", cfg.downsample_types[l], l);
        }
        const int S = stride_[l][0] * stride_[l][1] * stride_[l][2];
        if (ch_[l + 1] % S != 0 || (ch_[l + 1] / S) % 64 != 0 || (ch_[l] * S) % ch_[l + 1] != 0)
            fail("encoder: block %d cannot pixel-unshuffle %d -> %d channels with stride product %d", l, ch_[l], ch_[l + 1], S);
        // residual group sizes the downsampler-tail kernel is instantiated for (vae_glue.cu): S/2 and S, i.e. the
        // channel count doubles or stays the same through a block
        const int G = ch_[l] * S / ch_[l + 1];
        if (G != S / 2 && G != S)
            fail("encoder: block %d (%d -> %d channels, stride product %d) needs a residual group of %d; built: %d and %d",
                 l, ch_[l], ch_[l + 1], S, G, S / 2, S);
    }
    const std::string P = "encoder.";
    add_conv(P + "conv_in", conv_in_, cfg.in_channels * 16, 64, ch_[0], ch_[0]);
    for (int l = 0; l < 5; ++l) {
        const int C = ch_[l];
        const int n = l < 4 ? cfg.layers_per_block[l] : cfg.layers_per_block[4] - 1;  // vae.rs:1383-1386
        const std::string bp = l < 4 ? P + "down_blocks." + std::to_string(l) + "." : P + "mid_block.";
        res_[l].resize(n);
        for (int i = 0; i < n; ++i) {
            const std::string rp = bp + "resnets." + std::to_string(i) + ".";
            add_conv(rp + "conv1", res_[l][i].conv1, C, C, C, C);
            add_conv(rp + "conv2", res_[l][i].conv2, C, C, C, C);
        }
        if (l < 4) {
            const int S = stride_[l][0] * stride_[l][1] * stride_[l][2];
            add_conv(bp + "downsamplers.0.conv", down_[l], C, C, ch_[l + 1] / S, ch_[l + 1] / S);
        }
    }
    // latent + 1 output channels (vae.rs:1398-1407); computed as a 32-column multiple, extra weight rows are zero
    add_conv(P + "conv_out", conv_out_, ch_[4], ch_[4], cfg.latent_channels + 1, 256 * ((cfg.latent_channels + 32 + 255) / 256));
    LTXV_CUDA(cudaDeviceSynchronize());
}

void LtxVideoEncoder3d::load_tensor(const std::string& key, const void* data, int dtype, const int64_t* shape, int rank) {
    auto it = slots_.find(key);
    if (it == slots_.end()) fail("unknown VAE encoder tensor key '%s'", key.c_str());
    Slot& v = it->second;
    bool ok = rank == static_cast<int>(v.ps.shape.size());
    for (int i = 0; ok && i < rank; ++i) ok = shape[i] == v.ps.shape[i];
    if (!ok) fail("shape mismatch for VAE tensor '%s'", key.c_str());
    LTXV_CUDA(cudaSetDevice(device_));
    DevBuf tmp;
    tmp.ensure(static_cast<size_t>(v.ps.numel()) * 4);
    ingest_tensor(tmp.p, false, data, dtype, v.ps.numel());
    const ConvW& cw = *v.conv;
    if (v.kind == 1)
        LTXV_CUDA(launch_conv_weight_relayout(tmp.p, 0, cw.w, cw.Cout, cw.Cin, cw.rows_out, 0, 0, v.cin_src));
    else
        LTXV_CUDA(launch_conv_bias_relayout(tmp.p, 0, cw.b, cw.Cout, cw.rows_out, 0, 0));
    LTXV_CUDA(cudaDeviceSynchronize());
    v.ps.loaded = true;
    finalized_ = false;
}

void LtxVideoEncoder3d::init_random(uint64_t seed) {
    LTXV_CUDA(cudaSetDevice(device_));
    uint64_t n = 0;
    for (auto& kv : slots_) {
        Slot& v = kv.second;
        const uint64_t sd = seed * 104729ull + (++n);
        const ConvW& cw = *v.conv;
        if (v.kind == 1) {
            // generated in the reference layout and passed through the relayout, so that zero-padded input channels
            // (conv_in) and padded rows (conv_out) stay zero
            DevBuf tmp;
            tmp.ensure(static_cast<size_t>(v.ps.numel()) * 4);
            fill_uniform(tmp.p, false, v.ps.numel(), 1.0f / sqrtf(27.0f * v.cin_src), sd);
            LTXV_CUDA(launch_conv_weight_relayout(tmp.p, 0, cw.w, cw.Cout, cw.Cin, cw.rows_out, 0, 0, v.cin_src));
            LTXV_CUDA(cudaDeviceSynchronize());
        } else {
            fill_uniform(v.ps.dst, false, cw.Cout, 0.05f, sd);
        }
        v.ps.loaded = true;
    }
    LTXV_CUDA(cudaDeviceSynchronize());
    finalized_ = true;
}

void LtxVideoEncoder3d::finalize() {
    std::string missing;
    int n = 0;
    for (auto& kv : slots_) {
        if (kv.second.ps.loaded) continue;
        if (n < 8) missing += (n ? ", " : "") + kv.first;
        ++n;
    }
    if (n) fail("%d VAE encoder tensors were never loaded (first: %s)", n, missing.c_str());
    finalized_ = true;
}

void LtxVideoEncoder3d::latent_dims(int F, int H, int W, int* Fl, int* Hl, int* Wl) const {
    int t = F, h = H / 4, w = W / 4;
    if (F <= 0 || H <= 0 || W <= 0 || H % 4 != 0 || W % 4 != 0)
        fail("encode: video extent [%d,%d,%d] is not divisible by the patch size", F, H, W);
    for (int l = 0; l < 4; ++l) {
        const int st = stride_[l][0], sh = stride_[l][1], sw = stride_[l][2];
        if ((t + st - 1) % st != 0 || h % sh != 0 || w % sw != 0)
            fail("encode: video extent [%d,%d,%d] does not divide through downsampler %d (frames must be 8k+1, height "
                 "and width multiples of 32 for the default strides)", F, H, W, l);
        t = (t + st - 1) / st;
        h /= sh;
        w /= sw;
    }
    *Fl = t;
    *Hl = h;
    *Wl = w;
}

void LtxVideoEncoder3d::ensure_workspace(int F, int H, int W) {
    if (F == wsF_ && H == wsH_ && W == wsW_) return;
    int fl, hl, wl;
    latent_dims(F, H, W, &fl, &hl, &wl);
    T_[0] = F;
    H_[0] = H / 4;
    W_[0] = W / 4;
    for (int l = 0; l < 4; ++l) {
        T_[l + 1] = (T_[l] + stride_[l][0] - 1) / stride_[l][0];
        H_[l + 1] = H_[l] / stride_[l][1];
        W_[l + 1] = W_[l] / stride_[l][2];
    }
    size_t max_unpadded = 0;
    for (int l = 0; l < 5; ++l) {
        // T + 3 frames: two causal copies of frame 0, plus the one the temporal downsamplers add
        const size_t padded = static_cast<size_t>(T_[l] + 3) * (H_[l] + 2) * (W_[l] + 2) * ch_[l] * 2;
        p_[l].release();  // geometry changed: the zero border must be re-established
        p_[l].ensure(padded, true);
        q_[l].release();
        if (ch_[l] <= 256) q_[l].ensure(padded, true);  // second padded volume of the narrow levels (fused conv1 epilogue)
        const size_t un = static_cast<size_t>(T_[l] + 1) * H_[l] * W_[l] * ch_[l] * 2;
        max_unpadded = std::max(max_unpadded, un);
    }
    a_in_.release();
    a_in_.ensure(static_cast<size_t>(F + 2) * (H_[0] + 2) * (W_[0] + 2) * 64 * 2, true);
    xa_.ensure(max_unpadded);
    xb_.ensure(max_unpadded);
    hb_.ensure(max_unpadded);
    out32_.ensure(static_cast<size_t>(T_[4]) * H_[4] * W_[4] * (cfg_.latent_channels + 32) * 4);
    // the zero fills above ran on the legacy default stream: make them visible to whatever stream encode() is given
    LTXV_CUDA(cudaDeviceSynchronize());
    wsF_ = F;
    wsH_ = H;
    wsW_ = W;
}

void LtxVideoEncoder3d::conv(const ConvW& cw, const void* a_padded, int T, int H, int W, int epi, void* out,
                             const void* res, int n_cols, int ldo, cudaStream_t s, void* fused_norm_out, int fused_raw,
                             int fused_tf) {
    const int Wp = W + 2, plane = (H + 2) * Wp;
    GemmOperands ops{a_padded, static_cast<int64_t>(T + 2) * plane, cw.Cin, cw.w, cw.rows_out, 27ll * cw.Cin};
    GemmParams p{};
    p.M = T * plane;
    p.N = n_cols;
    p.K = 27 * cw.Cin;
    p.num_k_blocks = 27 * (cw.Cin / 64);
    p.epi = epi;
    p.ldo = ldo;
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
    // causal taps: output frame t reads padded frames t, t+1, t+2 = source frames t-2, t-1, t (clamped at 0)
    for (int kt = 0; kt < 3; ++kt)
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) p.tap_off[(kt * 3 + kh) * 3 + kw] = kt * plane + (kh - 1) * Wp + (kw - 1);
    if (fused_norm_out != nullptr) {
        // the epilogue is also the producer of the next conv's input: pixel norm + SiLU into the causally padded volume
        p.epi = EPI_CONV_NORM_PAD;
        p.norm_out = fused_norm_out;
        p.norm_do = p.norm_silu = fused_raw ? 0 : 1;  // raw: the downsampler's conv reads the activations as they are
        p.norm_tf = fused_tf;
    }
    LTXV_CUDA(launch_gemm_bf16(ops, p, 0, s));
}

// LtxVideoResnetBlock3d::forward (vae.rs:755-821), in == out, causal, no conditioning.
// `ready`: p_[l] already holds this resnet's conv1 input (written by the previous conv2's epilogue).  next_kind: what
// consumes the output -- 0 nothing fused, 1 another resnet (pixel norm + SiLU), 2 the downsampler (raw, `next_tf` front
// frames); on return `ready` says whether p_[l] holds that consumer's input.
void LtxVideoEncoder3d::resnet(const ResnetW& rw, int l, __nv_bfloat16*& x, __nv_bfloat16*& x_alt, cudaStream_t s,
                               bool* ready, int next_kind, int next_tf) {
    const int C = ch_[l], T = T_[l], H = H_[l], W = W_[l];
    const bool no_fuse = options().vae_no_fused_prep != 0;
    const bool no_conv2 = options().vae_no_fuse_conv2 != 0;
    if (!(ready && *ready)) LTXV_CUDA(launch_vae_prep(x, p_[l].p, nullptr, nullptr, 1, 1, T, H, W, C, s, nullptr, nullptr, 2));
    if (ready) *ready = false;
    if (C <= 256 && !no_fuse) {
        // narrow level: conv1's epilogue writes conv2's input (norm2 + SiLU) straight into the second padded volume;
        // its raw output is never stored (see EPI_CONV_NORM_PAD, gemm.h)
        conv(rw.conv1, p_[l].p, T, H, W, EPI_CONV_NDHWC, nullptr, nullptr, C, C, s, q_[l].p, 0, 2);
        // C = 256: conv2's epilogue also produces the next consumer's input into p_[l] (hidden under the longer main
        // loop there; at C = 128 it is not, see vae.cu)
        const bool fuse2 = C == 256 && !no_conv2 && next_kind != 0 && ready != nullptr;
        conv(rw.conv2, q_[l].p, T, H, W, EPI_CONV_NDHWC, x_alt, x, C, C, s, fuse2 ? p_[l].p : nullptr, next_kind == 2,
             next_kind == 2 ? next_tf : 2);
        if (fuse2) *ready = true;
    } else {
        conv(rw.conv1, p_[l].p, T, H, W, EPI_CONV_NDHWC, hb_.p, nullptr, C, C, s);
        LTXV_CUDA(launch_vae_prep(hb_.p, p_[l].p, nullptr, nullptr, 1, 1, T, H, W, C, s, nullptr, nullptr, 2));
        conv(rw.conv2, p_[l].p, T, H, W, EPI_CONV_NDHWC, x_alt, x, C, C, s);
    }
    std::swap(x, x_alt);
}

void LtxVideoEncoder3d::encode(const void* xin, int x_dtype, int B, int F, int H, int W, float* moments, cudaStream_t s) {
    if (!finalized_) finalize();
    if (B <= 0) fail("encode: invalid batch %d", B);
    if (x_dtype != LTXV_F32 && x_dtype != LTXV_BF16) fail("unsupported video dtype %d", x_dtype);
    LTXV_CUDA(cudaSetDevice(device_));
    ensure_workspace(F, H, W);
    const size_t xsz = x_dtype == LTXV_F32 ? 4 : 2;
    const int64_t x_elems = 3ll * F * H * W;
    const int L = cfg_.latent_channels;
    const int64_t nvox = static_cast<int64_t>(T_[4]) * H_[4] * W_[4];
    for (int b = 0; b < B; ++b) {
        const char* xb = static_cast<const char*>(xin) + static_cast<size_t>(b) * x_elems * xsz;
        LTXV_CUDA(launch_vae_patchify(xb, x_dtype == LTXV_BF16, a_in_.p, F, H, W, s));
        __nv_bfloat16* x = xa_.as<__nv_bfloat16>();
        __nv_bfloat16* x_alt = xb_.as<__nv_bfloat16>();
        conv(conv_in_, a_in_.p, T_[0], H_[0], W_[0], EPI_CONV_NDHWC, x, nullptr, ch_[0], ch_[0], s);
        for (int l = 0; l < 5; ++l) {
            const int st_l = l < 4 ? stride_[l][0] : 1;
            bool ready = false;
            for (size_t i = 0; i < res_[l].size(); ++i)
                resnet(res_[l][i], l, x, x_alt, s, &ready, i + 1 < res_[l].size() ? 1 : (l < 4 ? 2 : 0), 2 + (st_l - 1));
            if (l == 4) break;
            // LtxVideoDownsampler3d (vae.rs:534-582)
            const int st = stride_[l][0], sh = stride_[l][1], sw = stride_[l][2];
            const int Tp = T_[l] + st - 1;
            if (!ready)
                LTXV_CUDA(launch_vae_prep(x, p_[l].p, nullptr, nullptr, 0, 0, T_[l], H_[l], W_[l], ch_[l], s, nullptr, nullptr,
                                          2 + (st - 1)));
            conv(down_[l], p_[l].p, Tp, H_[l], W_[l], EPI_CONV_NDHWC, hb_.p, nullptr, down_[l].Cout, down_[l].Cout, s);
            LTXV_CUDA(launch_vae_unshuffle_add(hb_.p, x, x_alt, T_[l + 1], H_[l + 1], W_[l + 1], st, sh, sw, ch_[l],
                                               down_[l].Cout, s));
            std::swap(x, x_alt);
        }
        // norm_out -> SiLU -> conv_out -> last-channel replication (vae.rs:1455-1467)
        LTXV_CUDA(launch_vae_prep(x, p_[4].p, nullptr, nullptr, 1, 1, T_[4], H_[4], W_[4], ch_[4], s, nullptr, nullptr, 2));
        const int ld = L + 32;
        conv(conv_out_, p_[4].p, T_[4], H_[4], W_[4], EPI_STORE_F32, out32_.p, nullptr, ld, ld, s);
        LTXV_CUDA(launch_vae_moments(out32_.as<float>(), moments + static_cast<size_t>(b) * 2 * L * nvox, nvox, L, ld, s));
    }
}

// ------------------------------------------------------------------------------------------------
// tiled encode (compatibility mode; the reference's library default is use_tiling = true)
// ------------------------------------------------------------------------------------------------
namespace {
struct MomTile {  // one encoded tile: f32 [C, T, H, W]
    DevBuf buf;
    int T = 0, H = 0, W = 0;
    float* p() const { return buf.as<float>(); }
    void shape(int c, int t, int h, int w) {
        T = t;
        H = h;
        W = w;
        buf.ensure(static_cast<size_t>(c) * t * h * w * 4);
    }
};
}  // namespace

void LtxVideoEncoder3d::tiled_encode(const void* x, int x_dtype, int T, int H, int W, float* dst, const ltxv_vae_tiling& tp,
                                     cudaStream_t s) {
    const int sr = 32, C2 = 2 * cfg_.latent_channels;
    const int esz = x_dtype == LTXV_F32 ? 4 : 2;
    int Tl, Hl, Wl;
    latent_dims(T, H, W, &Tl, &Hl, &Wl);
    const int min_h = tp.tile_sample_min_height / sr, min_w = tp.tile_sample_min_width / sr;
    const int str_h = tp.tile_sample_stride_height / sr, str_w = tp.tile_sample_stride_width / sr;
    if (min_h < 1 || min_w < 1 || str_h < 1 || str_w < 1) fail("tile sizes and strides must be at least %d pixels", sr);
    if (tp.tile_sample_stride_height % sr || tp.tile_sample_stride_width % sr || tp.tile_sample_min_height % sr ||
        tp.tile_sample_min_width % sr)
        fail("tiled encode: tile sizes and strides must be multiples of %d pixels", sr);
    const int blend_h = std::max(min_h - str_h, 0), blend_w = std::max(min_w - str_w, 0);  // latent units (:2171-2172)
    const int ncols = (W + tp.tile_sample_stride_width - 1) / tp.tile_sample_stride_width;
    std::unique_ptr<MomTile[]> row_a(new MomTile[ncols]), row_b(new MomTile[ncols]);
    MomTile* prev = row_a.get();
    MomTile* cur = row_b.get();
    int ri = 0;
    for (int i = 0; i < H; i += tp.tile_sample_stride_height, ++ri) {
        const int th = std::min(i + tp.tile_sample_min_height, H) - i;
        int cj = 0;
        for (int j = 0; j < W; j += tp.tile_sample_stride_width, ++cj) {
            const int tw = std::min(j + tp.tile_sample_min_width, W) - j;
            tx_sp_.ensure(3ull * T * th * tw * esz);
            LTXV_CUDA(launch_copy_box(x, esz, 3, T, H, W, 0, i, j, tx_sp_.p, T, th, tw, 0, 0, 0, T, th, tw, s));
            int tl, hl, wl;
            latent_dims(T, th, tw, &tl, &hl, &wl);
            MomTile& tile = cur[cj];
            tile.shape(C2, tl, hl, wl);
            encode(tx_sp_.p, x_dtype, 1, T, th, tw, tile.p(), s);
            if (ri > 0)  // blend_v with the (already blended) tile above
                LTXV_CUDA(launch_blend_axis(prev[cj].p(), prev[cj].T, prev[cj].H, prev[cj].W, tile.p(), tile.T, tile.H,
                                            tile.W, C2, 2, blend_h, s));
            if (cj > 0)  // blend_h with the (already blended) left neighbour
                LTXV_CUDA(launch_blend_axis(cur[cj - 1].p(), cur[cj - 1].T, cur[cj - 1].H, cur[cj - 1].W, tile.p(), tile.T,
                                            tile.H, tile.W, C2, 3, blend_w, s));
            // keep the first stride rows / columns, cropped to the latent size (:2211-2222)
            const int y0 = ri * str_h, x0 = cj * str_w;
            const int hs = std::min(std::min(str_h, tile.H), Hl - y0), ws = std::min(std::min(str_w, tile.W), Wl - x0);
            if (hs > 0 && ws > 0)
                LTXV_CUDA(launch_copy_box(tile.p(), 4, C2, tile.T, tile.H, tile.W, 0, 0, 0, dst, Tl, Hl, Wl, 0, y0, x0, Tl,
                                          hs, ws, s));
        }
        std::swap(prev, cur);
    }
    LTXV_CUDA(cudaStreamSynchronize(s));  // the tile buffers die with this frame
}

void LtxVideoEncoder3d::temporal_tiled_encode(const void* x, int x_dtype, int F, int H, int W, float* dst,
                                              const ltxv_vae_tiling& tp, cudaStream_t s) {
    const int tr = 8, C2 = 2 * cfg_.latent_channels;
    const int esz = x_dtype == LTXV_F32 ? 4 : 2;
    int Fl, Hl, Wl;
    latent_dims(F, H, W, &Fl, &Hl, &Wl);
    const int min_t = tp.tile_sample_min_num_frames / tr, str_t = tp.tile_sample_stride_num_frames / tr;
    if (str_t < 1 || tp.tile_sample_stride_num_frames % tr || tp.tile_sample_min_num_frames % tr)
        fail("tile_sample_{min,stride}_num_frames must be multiples of %d", tr);
    const int blend_t = std::max(min_t - str_t, 0);
    const size_t plane = static_cast<size_t>(Hl) * Wl;
    int prev_T = 0, fo = 0, idx = 0;
    for (int i = 0; i < F && fo < Fl; i += tp.tile_sample_stride_num_frames, ++idx) {
        const int Tt = std::min(i + tp.tile_sample_min_num_frames + 1, F) - i;
        tx_tm_.ensure(3ull * Tt * H * W * esz);
        LTXV_CUDA(launch_copy_box(x, esz, 3, F, H, W, i, 0, 0, tx_tm_.p, Tt, H, W, 0, 0, 0, Tt, H, W, s));
        int tl, hl, wl;
        latent_dims(Tt, H, W, &tl, &hl, &wl);
        t_enc_.ensure(static_cast<size_t>(C2) * tl * plane * 4);
        if (tp.use_tiling && (H > tp.tile_sample_min_height || W > tp.tile_sample_min_width))
            tiled_encode(tx_tm_.p, x_dtype, Tt, H, W, t_enc_.as<float>(), tp, s);
        else
            encode(tx_tm_.p, x_dtype, 1, Tt, H, W, t_enc_.as<float>(), s);
        // the FIRST tile loses its first latent frame (:2324-2329); compact into t_cur_
        const int t0 = idx == 0 ? 1 : 0;
        const int Tc = tl - t0;
        if (Tc <= 0) fail("temporal tiled encode: the first tile has a single latent frame (clip too short for framewise encoding)");
        t_cur_.ensure(static_cast<size_t>(C2) * Tc * plane * 4);
        LTXV_CUDA(launch_copy_box(t_enc_.p, 4, C2, tl, Hl, Wl, t0, 0, 0, t_cur_.p, Tc, Hl, Wl, 0, 0, 0, Tc, Hl, Wl, s));
        int take;
        const float* src = t_cur_.as<float>();
        if (idx > 0) {
            // blend with the UNBLENDED previous tile (row[idx - 1], :2342), keep the first `stride` latent frames
            t_work_.ensure(static_cast<size_t>(C2) * Tc * plane * 4);
            LTXV_CUDA(cudaMemcpyAsync(t_work_.p, t_cur_.p, static_cast<size_t>(C2) * Tc * plane * 4, cudaMemcpyDeviceToDevice, s));
            LTXV_CUDA(launch_blend_axis(t_prev_.as<float>(), prev_T, Hl, Wl, t_work_.as<float>(), Tc, Hl, Wl, C2, 1, blend_t, s));
            src = t_work_.as<float>();
            take = std::min(str_t, Tc);
        } else {
            take = std::min(str_t + 1, Tc);
        }
        take = std::min(take, Fl - fo);
        if (take > 0)
            LTXV_CUDA(launch_copy_box(src, 4, C2, Tc, Hl, Wl, 0, 0, 0, dst, Fl, Hl, Wl, fo, 0, 0, take, Hl, Wl, s));
        fo += take;
        std::swap(t_prev_.p, t_cur_.p);
        std::swap(t_prev_.bytes, t_cur_.bytes);
        prev_T = Tc;
    }
    if (fo < Fl) fail("temporal tiled encode produced %d of %d latent frames", fo, Fl);
}

void LtxVideoEncoder3d::encode_z(const void* x, int x_dtype, int B, int F, int H, int W, const ltxv_vae_tiling* tiling,
                                 int use_framewise_encoding, float* moments, cudaStream_t s) {
    if (tiling == nullptr) return encode(x, x_dtype, B, F, H, W, moments, s);
    if (x_dtype != LTXV_F32 && x_dtype != LTXV_BF16) fail("unsupported video dtype %d", x_dtype);
    const ltxv_vae_tiling& tp = *tiling;
    const bool temporal = use_framewise_encoding && F > tp.tile_sample_min_num_frames;
    const bool spatial = tp.use_tiling && (H > tp.tile_sample_min_height || W > tp.tile_sample_min_width);
    if (!temporal && !spatial) return encode(x, x_dtype, B, F, H, W, moments, s);
    LTXV_CUDA(cudaSetDevice(device_));
    int Fl, Hl, Wl;
    latent_dims(F, H, W, &Fl, &Hl, &Wl);
    const size_t xsz = x_dtype == LTXV_F32 ? 4 : 2;
    const size_t x_elems = 3ull * F * H * W, m_elems = 2ull * cfg_.latent_channels * Fl * Hl * Wl;
    for (int b = 0; b < B; ++b) {
        const char* xb = static_cast<const char*>(x) + b * x_elems * xsz;
        float* mb = moments + b * m_elems;
        if (temporal) temporal_tiled_encode(xb, x_dtype, F, H, W, mb, tp, s);
        else tiled_encode(xb, x_dtype, F, H, W, mb, tp, s);
    }
}

}  // namespace ltxv
