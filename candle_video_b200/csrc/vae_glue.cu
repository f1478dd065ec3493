#include "common.cuh"
#include "vae_glue.h"
#define LTXV_PDL_CLASS 16
#include "launch.h"
#include "profile.h"
#include "options.h"

#include <atomic>

namespace ltxv {

namespace {
std::atomic<uint64_t> g_vae_glue_launches{0};
inline cudaError_t done() {
    g_vae_glue_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

template <typename TIn>
__global__ void vae_input_kernel(const TIn* __restrict__ z, __nv_bfloat16* __restrict__ out, int C, int F, int Hfull, int W,
                                 int h0, int Hs) {
    // enumerates the local padded rows hp in [0, Hs+2) that exist in the latent (global row h0 + hp - 1)
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t n = static_cast<int64_t>(F) * (Hs + 2) * W * C;
    if (idx >= n) return;
    const int c = static_cast<int>(idx % C);
    const int64_t vox = idx / C;
    const int w = static_cast<int>(vox % W), hp = static_cast<int>((vox / W) % (Hs + 2)),
              f = static_cast<int>(vox / (static_cast<int64_t>(W) * (Hs + 2)));
    const int h = h0 + hp - 1;
    if (h < 0 || h >= Hfull) return;  // true volume border: stays zero (vae.rs:344)
    const float v = static_cast<float>(z[((static_cast<int64_t>(c) * F + f) * Hfull + h) * W + w]);
    const __nv_bfloat16 b = __float2bfloat16(v);
    const int Hp = Hs + 2, Wp = W + 2;
    auto at = [&](int tp) { return ((static_cast<int64_t>(tp) * Hp + hp) * Wp + (w + 1)) * C + c; };
    out[at(f + 1)] = b;
    if (f == 0) out[at(0)] = b;
    if (f == F - 1) out[at(F + 1)] = b;
}

// Channels are split into 16-byte chunks of 8; chunk j of a voxel is owned by lane (j % 32) so that every warp-wide
// access is a contiguous 256/512-byte run.  C = 128 uses half a warp per voxel (2 voxels per warp), C >= 256 a full warp
// with C/256 chunks per lane.  Warps stride over the voxels (persistent grid) with scale/shift held in registers.
template <int C, int U, int MINB>
__global__ void __launch_bounds__(256, MINB)
vae_prep_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, const float* __restrict__ scale,
                const float* __restrict__ shift, int do_norm, int do_silu, int T, int H, int W, int tf,
                __nv_bfloat16* __restrict__ halo_up, __nv_bfloat16* __restrict__ halo_dn, int H_up, int H_dn) {
    // H_up / H_dn: slab rows of the neighbours owning halo_up / halo_dn (ragged H-slabs: their plane stride differs)
    // tf = replicated frames in front of the volume: 1 = the decoder's non-causal padding (one more copy of the last
    // frame behind it), 2 = the encoder's causal padding (vae.rs:383-387), 3 = causal padding of a volume whose first
    // frame is duplicated once more by the temporal downsampler (vae.rs:539-544)
    constexpr bool kMod = C <= 1024;  // C = 2048 (encoder mid block) is never modulated: keep its registers free
    constexpr int LPV = (C / 8 < 32) ? C / 8 : 32;  // lanes per voxel
    constexpr int VPW = 32 / LPV;                    // voxels per warp
    constexpr int CPL = (C / 8) / LPV;               // chunks per lane
    griddep_launch_dependents();
    griddep_wait();
    const int lane = threadIdx.x & 31;
    const int sub = lane / LPV, l = lane % LPV;
    const int Hp = H + 2, Wp = W + 2;

    float sc[kMod ? CPL : 1][8], sh[kMod ? CPL : 1][8];
    if (kMod && scale != nullptr) {
#pragma unroll
        for (int k = 0; k < (kMod ? CPL : 1); ++k) {
            const int c0 = (k * LPV + l) * 8;
            const float4 a = __ldg(reinterpret_cast<const float4*>(scale + c0));
            const float4 b = __ldg(reinterpret_cast<const float4*>(scale + c0 + 4));
            const float4 c = __ldg(reinterpret_cast<const float4*>(shift + c0));
            const float4 d = __ldg(reinterpret_cast<const float4*>(shift + c0 + 4));
            sc[k][0] = 1.f + a.x; sc[k][1] = 1.f + a.y; sc[k][2] = 1.f + a.z; sc[k][3] = 1.f + a.w;
            sc[k][4] = 1.f + b.x; sc[k][5] = 1.f + b.y; sc[k][6] = 1.f + b.z; sc[k][7] = 1.f + b.w;
            sh[k][0] = c.x; sh[k][1] = c.y; sh[k][2] = c.z; sh[k][3] = c.w;
            sh[k][4] = d.x; sh[k][5] = d.y; sh[k][6] = d.z; sh[k][7] = d.w;
        }
    }
    // One CTA pass = one (t, h) row of the volume: the frame / halo replication flags and the destination bases are
    // row-uniform, so all of the index arithmetic (two integer divisions, the 64-bit plane offsets, up to eleven
    // destinations) is paid once per row.  The previous voxel-strided loop spent 60 % of its ~205 warp instructions per
    // 256 elements there and was issue-bound at 0.43 of the copy bandwidth (profiles/r02_ncu_vae_prep.csv).
    auto load = [&](const __nv_bfloat16* src, float (&v)[CPL][8]) {
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + k * LPV * 8));
            v[k][0] = bf16_lo(u.x); v[k][1] = bf16_hi(u.x); v[k][2] = bf16_lo(u.y); v[k][3] = bf16_hi(u.y);
            v[k][4] = bf16_lo(u.z); v[k][5] = bf16_hi(u.z); v[k][6] = bf16_lo(u.w); v[k][7] = bf16_hi(u.w);
        }
    };
    auto math = [&](float (&v)[CPL][8]) {
        if (do_norm) {
            float s2 = 0.f;
#pragma unroll
            for (int k = 0; k < CPL; ++k)
#pragma unroll
                for (int i = 0; i < 8; ++i) s2 += v[k][i] * v[k][i];
#pragma unroll
            for (int o = LPV / 2; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            const float rinv = rsqrtf(s2 * (1.0f / C) + 1e-8f);
#pragma unroll
            for (int k = 0; k < CPL; ++k)
#pragma unroll
                for (int i = 0; i < 8; ++i) v[k][i] *= rinv;
        }
        if (kMod && scale != nullptr) {
#pragma unroll
            for (int k = 0; k < (kMod ? CPL : 1); ++k)
#pragma unroll
                for (int i = 0; i < 8; ++i) v[k][i] = v[k][i] * sc[k][i] + sh[k][i];
        }
        if (do_silu) {
#pragma unroll
            for (int k = 0; k < CPL; ++k)
#pragma unroll
                for (int i = 0; i < 8; ++i) v[k][i] = silu_fast_f32(v[k][i]);
        }
    };
    constexpr int VPB = 8 * VPW;              // voxels per CTA pass
    // U passes in flight: every load is issued before the first is consumed
    const int rows = T * H;
    const int64_t pl_main = static_cast<int64_t>(Hp) * Wp * C;
    const int lane_off = l * 8;
    const int wv = (threadIdx.x >> 5) * VPW + sub;  // this thread's voxel inside a pass
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int t = row / H;
        const int h = row - t * H;
        const __nv_bfloat16* src = x + static_cast<int64_t>(row) * W * C + lane_off;
        __nv_bfloat16* d_main = out + ((static_cast<int64_t>(t + tf) * Hp + h + 1) * Wp + 1) * C + lane_off;
        const bool rep_front = t == 0;                   // replicate frame 0 into the tf planes in front of it
        const bool rep_back = tf == 1 && t == T - 1;     // non-causal: one more copy of the last frame
        const bool up = halo_up != nullptr && h == 0;    // my first row = bottom halo of the slab above
        const bool dn = halo_dn != nullptr && h == H - 1;  // my last row = top halo of the slab below
        const bool extra = rep_front || rep_back || up || dn;
        for (int w0 = wv; w0 - wv < W; w0 += U * VPB) {  // (w0 - wv is CTA-uniform: the shuffles stay full-warp)
            float xv[U][CPL][8];
            int off[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int w = w0 + u * VPB;
                ok[u] = w < W;
                off[u] = (ok[u] ? w : W - 1) * C;
                load(src + off[u], xv[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                math(xv[u]);
                if (!ok[u]) continue;
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    uint4 o;
                    o.x = pack_bf16x2(xv[u][k][0], xv[u][k][1]);
                    o.y = pack_bf16x2(xv[u][k][2], xv[u][k][3]);
                    o.z = pack_bf16x2(xv[u][k][4], xv[u][k][5]);
                    o.w = pack_bf16x2(xv[u][k][6], xv[u][k][7]);
                    const int eo = off[u] + k * LPV * 8;
                    *reinterpret_cast<uint4*>(d_main + eo) = o;
                    if (extra) {
                        auto put = [&](__nv_bfloat16* base, int64_t pl) {
                            if (rep_front)
                                for (int q = 1; q <= tf; ++q) *reinterpret_cast<uint4*>(base - q * pl + eo) = o;
                            if (rep_back) *reinterpret_cast<uint4*>(base + pl + eo) = o;
                        };
                        put(d_main, pl_main);
                        if (up) {
                            const int64_t pl = static_cast<int64_t>(H_up + 2) * Wp * C;
                            __nv_bfloat16* d = halo_up + ((static_cast<int64_t>(t + tf) * (H_up + 2) + H_up + 1) * Wp + 1) * C + lane_off;
                            *reinterpret_cast<uint4*>(d + eo) = o;
                            put(d, pl);
                        }
                        if (dn) {
                            const int64_t pl = static_cast<int64_t>(H_dn + 2) * Wp * C;
                            __nv_bfloat16* d = halo_dn + (static_cast<int64_t>(t + tf) * (H_dn + 2) * Wp + 1) * C + lane_off;
                            *reinterpret_cast<uint4*>(d + eo) = o;
                            put(d, pl);
                        }
                    }
                }
            }
        }
    }
}

template <typename TIn>
__global__ void conv_weight_relayout_kernel(const TIn* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout,
                                            int Cin, int rows_out, int d2s_perm, int cin_src) {
    // out[r, tap*Cin + c]; the source has cin_src <= Cin input channels (the rest of the K row is zero)
    const int64_t K = 27ll * Cin;
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= rows_out * K) return;
    const int r = static_cast<int>(idx / K);
    const int k = static_cast<int>(idx - r * K);
    const int tap = k / Cin, c = k - tap * Cin;
    int co = r;
    if (d2s_perm) {
        const int cp = Cout >> 3;
        const int sub = r / cp, c1 = r - sub * cp;
        co = c1 * 8 + sub;
    }
    float v = 0.f;
    if (r < Cout && c < cin_src) v = static_cast<float>(w[(static_cast<int64_t>(co) * cin_src + c) * 27 + tap]);
    out[idx] = __float2bfloat16(v);
}
template <typename TIn>
__global__ void conv_bias_relayout_kernel(const TIn* __restrict__ b, float* __restrict__ out, int Cout, int rows_out,
                                          int d2s_perm) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows_out) return;
    int co = r;
    if (d2s_perm) {
        const int cp = Cout >> 3;
        const int sub = r / cp, c1 = r - sub * cp;
        co = c1 * 8 + sub;
    }
    out[r] = (r < Cout) ? static_cast<float>(b[co]) : 0.f;
}

// ---- encoder-side glue ----
// One thread per (patch voxel, colour plane slot): 4 rows x 4 pixels of one plane become 16 consecutive channels
// c*16 + pw*4 + ph (vae.rs:1427-1445); slot 3 writes the 16 zero channels that pad 48 -> 64.
template <typename TIn>
__global__ void vae_patchify_kernel(const TIn* __restrict__ x, __nv_bfloat16* __restrict__ out, int F, int H, int W) {
    const int Hq = H >> 2, Wq = W >> 2;
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t n = static_cast<int64_t>(F) * Hq * Wq * 4;
    if (idx >= n) return;
    const int c = static_cast<int>(idx & 3);
    const int64_t vox = idx >> 2;
    const int w = static_cast<int>(vox % Wq), h = static_cast<int>((vox / Wq) % Hq),
              f = static_cast<int>(vox / (static_cast<int64_t>(Wq) * Hq));
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.f;
    if (c < 3) {
        const TIn* src = x + ((static_cast<int64_t>(c) * F + f) * H + 4 * h) * W + 4 * w;
#pragma unroll
        for (int ph = 0; ph < 4; ++ph)
#pragma unroll
            for (int pw = 0; pw < 4; ++pw) v[pw * 4 + ph] = static_cast<float>(src[static_cast<int64_t>(ph) * W + pw]);
    }
    uint4 o0, o1;
    o0.x = pack_bf16x2(v[0], v[1]);   o0.y = pack_bf16x2(v[2], v[3]);   o0.z = pack_bf16x2(v[4], v[5]);   o0.w = pack_bf16x2(v[6], v[7]);
    o1.x = pack_bf16x2(v[8], v[9]);   o1.y = pack_bf16x2(v[10], v[11]); o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
    const int Hp = Hq + 2, Wp = Wq + 2;
    const int64_t plane = static_cast<int64_t>(Hp) * Wp;
    const int64_t row = (static_cast<int64_t>(f + 2) * Hp + (h + 1)) * Wp + (w + 1);  // causal: two frames in front
    for (int q = 0; q <= (f == 0 ? 2 : 0); ++q) {
        uint4* d = reinterpret_cast<uint4*>(out + (row - q * plane) * 64 + c * 16);
        d[0] = o0;
        d[1] = o1;
    }
}

// LtxVideoDownsampler3d tail (vae.rs:549-581): out[t,h,w, oc] = conv[t*st+i, h*sh+j, w*sw+k, oc / S]  (sub = oc % S)
//   + mean_{g < G} xdup[.., (oc*G + g) / S] at sub-voxel (oc*G + g) % S,   xdup[t] = x[max(t - (st-1), 0)].
// One thread per (output voxel, 8 consecutive conv channels): it owns the 8*S output channels [cc0*S, (cc0+8)*S), which
// draw on conv channels [cc0, cc0+8) and input channels [cc0*G, (cc0+8)*G) of the S source voxels -- every access is a
// 16-byte vector (the scalar version spent 24 two-byte loads per 16 bytes of output).  All operands NDHWC bf16.
template <int ST, int SH, int SW, int G>
__global__ void vae_unshuffle_add_kernel(const __nv_bfloat16* __restrict__ conv, const __nv_bfloat16* __restrict__ x,
                                         __nv_bfloat16* __restrict__ out, int To, int Ho, int Wo, int C, int Cc) {
    constexpr int S = ST * SH * SW;
    const int Cout = Cc * S;
    const int groups = Cc >> 3;
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t n = static_cast<int64_t>(To) * Ho * Wo * groups;
    if (idx >= n) return;
    const int cc0 = static_cast<int>(idx % groups) * 8;
    const int64_t vox = idx / groups;
    const int w = static_cast<int>(vox % Wo), h = static_cast<int>((vox / Wo) % Ho),
              t = static_cast<int>(vox / (static_cast<int64_t>(Wo) * Ho));
    const int Hi = Ho * SH, Wi = Wo * SW;
    float acc[8 * S];
#pragma unroll
    for (int i = 0; i < 8 * S; ++i) acc[i] = 0.f;
    auto unpack8 = [](const uint4& u, float (&f)[8]) {
        f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
        f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
    };
    // residual: input channel (local) cl = 8*g8 + e of source sub-voxel su is unshuffled channel u = cl*S + su
    // (relative to cc0*S*G) and belongs to output channel u / G
#pragma unroll
    for (int su = 0; su < S; ++su) {
        const int i = su / (SH * SW), j = (su / SW) % SH, k = su % SW;
        int ts = t * ST + i - (ST - 1);
        ts = ts < 0 ? 0 : ts;
        const int64_t xv = (static_cast<int64_t>(ts) * Hi + (h * SH + j)) * Wi + (w * SW + k);
        const uint4* xp = reinterpret_cast<const uint4*>(x + xv * C + cc0 * G);
#pragma unroll
        for (int g8 = 0; g8 < G; ++g8) {
            float f[8];
            unpack8(__ldg(xp + g8), f);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[((8 * g8 + e) * S + su) / G] += f[e];
        }
    }
    constexpr float inv_g = 1.0f / static_cast<float>(G);
#pragma unroll
    for (int sub = 0; sub < S; ++sub) {
        const int i = sub / (SH * SW), j = (sub / SW) % SH, k = sub % SW;
        const int64_t cv = (static_cast<int64_t>(t * ST + i) * Hi + (h * SH + j)) * Wi + (w * SW + k);
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(conv + cv * Cc + cc0)), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e * S + sub] = f[e] + acc[e * S + sub] * inv_g;
    }
    uint4* o4 = reinterpret_cast<uint4*>(out + vox * Cout + cc0 * S);
#pragma unroll
    for (int q = 0; q < S; ++q)
        o4[q] = make_uint4(pack_bf16x2(acc[8 * q], acc[8 * q + 1]), pack_bf16x2(acc[8 * q + 2], acc[8 * q + 3]),
                           pack_bf16x2(acc[8 * q + 4], acc[8 * q + 5]), pack_bf16x2(acc[8 * q + 6], acc[8 * q + 7]));
}

// conv_out rows [voxel, ld] f32 (columns 0..L = mean channels + one logvar channel) -> moments [2L, T*H*W] f32 with
// the last conv channel replicated over [L, 2L) (vae.rs:1462-1467)
__global__ void vae_moments_kernel(const float* __restrict__ h, float* __restrict__ out, int64_t nvox, int L, int ld) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= nvox * 2 * L) return;
    const int64_t v = idx % nvox;
    const int c = static_cast<int>(idx / nvox);
    out[idx] = h[v * ld + (c < L ? c : L)];
}

// normalize_latents (t2v_pipeline.rs:552-571): (x - mean[c]) * scaling_factor / std[c] on [B, C, inner] f32
__global__ void normalize_latents_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                         const float* __restrict__ std, float sf, float* __restrict__ out, int C,
                                         int64_t inner, int64_t n) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= n) return;
    const int c = static_cast<int>((idx / inner) % C);
    out[idx] = __fdiv_rn(__fmul_rn(__fsub_rn(x[idx], mean[c]), sf), std[c]);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// tiled-decode helpers (vae.rs:1927-2006, :2225-2290, :2358-2434): box copies and linear seam blends on NCDHW volumes
// ------------------------------------------------------------------------------------------------
struct Vol {
    int C, T, H, W;
};
template <typename E>
__global__ void copy_box_kernel(const E* __restrict__ src, Vol sv, int st0, int sh0, int sw0, E* __restrict__ dst, Vol dv,
                                int dt0, int dh0, int dw0, int bT, int bH, int bW) {
    const int64_t n = static_cast<int64_t>(sv.C) * bT * bH * bW;
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const int w = static_cast<int>(i % bW);
    const int h = static_cast<int>((i / bW) % bH);
    const int t = static_cast<int>((i / (static_cast<int64_t>(bW) * bH)) % bT);
    const int c = static_cast<int>(i / (static_cast<int64_t>(bW) * bH * bT));
    dst[((static_cast<int64_t>(c) * dv.T + dt0 + t) * dv.H + dh0 + h) * dv.W + dw0 + w] =
        src[((static_cast<int64_t>(c) * sv.T + st0 + t) * sv.H + sh0 + h) * sv.W + sw0 + w];
}
// b[.., x along axis] = a[.., La - blend + x] * (1 - x * f32(1/blend)) + b[.., x] * (x * f32(1/blend)), x < blend;
// a and b agree in every other extent.  axis: 1 = T, 2 = H, 3 = W.
__global__ void blend_axis_kernel(const float* __restrict__ a, Vol av, float* __restrict__ b, Vol bv, int axis,
                                  int blend, float inv_blend) {
    const int eT = axis == 1 ? blend : bv.T, eH = axis == 2 ? blend : bv.H, eW = axis == 3 ? blend : bv.W;
    const int64_t n = static_cast<int64_t>(bv.C) * eT * eH * eW;
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const int w = static_cast<int>(i % eW);
    const int h = static_cast<int>((i / eW) % eH);
    const int t = static_cast<int>((i / (static_cast<int64_t>(eW) * eH)) % eT);
    const int c = static_cast<int>(i / (static_cast<int64_t>(eW) * eH * eT));
    const int x = axis == 1 ? t : (axis == 2 ? h : w);
    const int at = axis == 1 ? av.T - blend + t : t, ah = axis == 2 ? av.H - blend + h : h,
              aw = axis == 3 ? av.W - blend + w : w;
    const float wgt = __fmul_rn(static_cast<float>(x), inv_blend);
    const float om = __fsub_rn(1.0f, wgt);
    const float va = a[((static_cast<int64_t>(c) * av.T + at) * av.H + ah) * av.W + aw];
    float* pb = b + ((static_cast<int64_t>(c) * bv.T + t) * bv.H + h) * bv.W + w;
    *pb = __fadd_rn(__fmul_rn(va, om), __fmul_rn(*pb, wgt));
}

uint64_t vae_glue_launch_count() { return g_vae_glue_launches.load(); }

cudaError_t launch_vae_input(const void* z, int z_is_bf16, void* out, int C, int F, int H, int W, cudaStream_t s) {
    return launch_vae_input(z, z_is_bf16, out, C, F, H, W, 0, H, s);
}

cudaError_t launch_vae_input(const void* z, int z_is_bf16, void* out, int C, int F, int H, int W, int h0, int Hs,
                             cudaStream_t s) {
    const int64_t n = static_cast<int64_t>(F) * (Hs + 2) * W * C;
    const int grid = static_cast<int>((n + 255) / 256);
    if (z_is_bf16)
        vae_input_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(z),
                                                             reinterpret_cast<__nv_bfloat16*>(out), C, F, H, W, h0, Hs);
    else
        vae_input_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(z),
                                                     reinterpret_cast<__nv_bfloat16*>(out), C, F, H, W, h0, Hs);
    return done();
}

cudaError_t launch_vae_prep(const void* x, void* out, const float* scale, const float* shift, int do_norm, int do_silu,
                            int T, int H, int W, int C, cudaStream_t s, void* halo_up, void* halo_dn, int tf, int H_up,
                            int H_dn) {
    if (H_up <= 0) H_up = H;
    if (H_dn <= 0) H_dn = H;
    if (tf < 1 || tf > 3 || (C > 1024 && scale != nullptr)) return cudaErrorInvalidValue;
    __nv_bfloat16* hu = reinterpret_cast<__nv_bfloat16*>(halo_up);
    __nv_bfloat16* hd = reinterpret_cast<__nv_bfloat16*>(halo_dn);
    const int64_t nvox = static_cast<int64_t>(T) * H * W;
    if (static_cast<int64_t>(T) * H >= (1ll << 31) - 148 * 8 || static_cast<int64_t>(W + 64) * C >= (1ll << 31))
        return cudaErrorInvalidValue;  // the kernel indexes rows and the elements of a row in 32 bits
    const int64_t rows = static_cast<int64_t>(T) * H;  // one (t, h) row per CTA pass
    const int grid = static_cast<int>(rows < 148 * 8 ? rows : 148 * 8);
    const __nv_bfloat16* xi = reinterpret_cast<const __nv_bfloat16*>(x);
    __nv_bfloat16* xo = reinterpret_cast<__nv_bfloat16*>(out);
    ProfScope prof(PROF_VAE_PREP, 4.0 * static_cast<double>(nvox) * C, s);  // bf16 in, bf16 out
    LTXV_TRACE_VARIANT("vae_prep_kernel<%d>", C);
    const int u128 = options().vae_prep_u;
    switch (C) {
        case 128:
            if (u128 == 2) launch_pdl(vae_prep_kernel<128, 2, 4>, dim3(grid), dim3(256), 0, s, xi, xo, scale, shift, do_norm, do_silu, T, H, W, tf, hu, hd, H_up, H_dn);
            else if (u128 == 3) launch_pdl(vae_prep_kernel<128, 2, 1>, dim3(grid), dim3(256), 0, s, xi, xo, scale, shift, do_norm, do_silu, T, H, W, tf, hu, hd, H_up, H_dn);
            else if (u128 == 8) launch_pdl(vae_prep_kernel<128, 8, 2>, dim3(grid), dim3(256), 0, s, xi, xo, scale, shift, do_norm, do_silu, T, H, W, tf, hu, hd, H_up, H_dn);
            else launch_pdl(vae_prep_kernel<128, 4, 3>, dim3(grid), dim3(256), 0, s, xi, xo, scale, shift, do_norm, do_silu, T, H, W, tf, hu, hd, H_up, H_dn);
            break;
        case 256: launch_pdl(vae_prep_kernel<256, 4, 3>, dim3(grid), dim3(256), 0, s, xi, xo, scale, shift, do_norm, do_silu, T, H, W, tf, hu, hd, H_up, H_dn); break;
        case 512: launch_pdl(vae_prep_kernel<512, 4, 2>, dim3(grid), dim3(256), 0, s, xi, xo, scale, shift, do_norm, do_silu, T, H, W, tf, hu, hd, H_up, H_dn); break;
        case 1024: launch_pdl(vae_prep_kernel<1024, 2, 1>, dim3(grid), dim3(256), 0, s, xi, xo, scale, shift, do_norm, do_silu, T, H, W, tf, hu, hd, H_up, H_dn); break;
        case 2048: launch_pdl(vae_prep_kernel<2048, 1, 1>, dim3(grid), dim3(256), 0, s, xi, xo, scale, shift, do_norm, do_silu, T, H, W, tf, hu, hd, H_up, H_dn); break;
        default: return cudaErrorInvalidValue;
    }
    return done();
}

cudaError_t launch_vae_patchify(const void* x, int x_is_bf16, void* out_padded, int F, int H, int W, cudaStream_t s) {
    if ((H & 3) || (W & 3)) return cudaErrorInvalidValue;
    const int64_t n = static_cast<int64_t>(F) * (H >> 2) * (W >> 2) * 4;
    const int grid = static_cast<int>((n + 255) / 256);
    if (x_is_bf16)
        vae_patchify_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                                reinterpret_cast<__nv_bfloat16*>(out_padded), F, H, W);
    else
        vae_patchify_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(x),
                                                        reinterpret_cast<__nv_bfloat16*>(out_padded), F, H, W);
    return done();
}

cudaError_t launch_vae_unshuffle_add(const void* conv, const void* x, void* out, int To, int Ho, int Wo, int st, int sh,
                                     int sw, int C, int Cc, cudaStream_t s) {
    const int S = st * sh * sw;
    const int Cout = Cc * S;
    if (Cc % 8 != 0 || (C * S) % Cout != 0) return cudaErrorInvalidValue;
    const int G = C * S / Cout;
    const int64_t n = static_cast<int64_t>(To) * Ho * Wo * (Cc >> 3);
    const int grid = static_cast<int>((n + 127) / 128);
    const __nv_bfloat16* cp = reinterpret_cast<const __nv_bfloat16*>(conv);
    const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
    __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(out);
    // the three DownsampleType strides (vae.rs:484-493) with the group sizes of the 0.9.5 widths
    if (st == 1 && sh == 2 && sw == 2 && G == 2)
        vae_unshuffle_add_kernel<1, 2, 2, 2><<<grid, 128, 0, s>>>(cp, xp, op, To, Ho, Wo, C, Cc);
    else if (st == 2 && sh == 1 && sw == 1 && G == 1)
        vae_unshuffle_add_kernel<2, 1, 1, 1><<<grid, 128, 0, s>>>(cp, xp, op, To, Ho, Wo, C, Cc);
    else if (st == 2 && sh == 2 && sw == 2 && G == 4)
        vae_unshuffle_add_kernel<2, 2, 2, 4><<<grid, 128, 0, s>>>(cp, xp, op, To, Ho, Wo, C, Cc);
    else if (st == 1 && sh == 2 && sw == 2 && G == 4)
        vae_unshuffle_add_kernel<1, 2, 2, 4><<<grid, 128, 0, s>>>(cp, xp, op, To, Ho, Wo, C, Cc);
    else if (st == 2 && sh == 1 && sw == 1 && G == 2)
        vae_unshuffle_add_kernel<2, 1, 1, 2><<<grid, 128, 0, s>>>(cp, xp, op, To, Ho, Wo, C, Cc);
    else if (st == 2 && sh == 2 && sw == 2 && G == 8)
        vae_unshuffle_add_kernel<2, 2, 2, 8><<<grid, 128, 0, s>>>(cp, xp, op, To, Ho, Wo, C, Cc);
    else
        return cudaErrorInvalidValue;  // channel ratio outside the pixel-unshuffle layouts built here
    return done();
}

cudaError_t launch_vae_moments(const float* h, float* out, int64_t nvox, int L, int ld, cudaStream_t s) {
    const int64_t n = nvox * 2 * L;
    vae_moments_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, s>>>(h, out, nvox, L, ld);
    return done();
}

cudaError_t launch_normalize_latents(const float* x, const float* mean, const float* std, float scaling_factor,
                                     float* out, int B, int C, int64_t inner, cudaStream_t s) {
    const int64_t n = static_cast<int64_t>(B) * C * inner;
    if (n <= 0) return cudaSuccess;
    normalize_latents_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, s>>>(x, mean, std, scaling_factor, out, C, inner, n);
    return done();
}

cudaError_t launch_conv_weight_relayout(const void* w, int w_is_bf16, void* out, int Cout, int Cin, int rows_out,
                                        int d2s_perm, cudaStream_t s, int cin_src) {
    if (cin_src <= 0) cin_src = Cin;
    const int64_t n = static_cast<int64_t>(rows_out) * 27 * Cin;
    const int grid = static_cast<int>((n + 255) / 256);
    if (w_is_bf16)
        conv_weight_relayout_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(
            reinterpret_cast<const __nv_bfloat16*>(w), reinterpret_cast<__nv_bfloat16*>(out), Cout, Cin, rows_out, d2s_perm,
            cin_src);
    else
        conv_weight_relayout_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(w),
                                                                reinterpret_cast<__nv_bfloat16*>(out), Cout, Cin,
                                                                rows_out, d2s_perm, cin_src);
    return done();
}
cudaError_t launch_conv_bias_relayout(const void* b, int b_is_bf16, float* out, int Cout, int rows_out, int d2s_perm,
                                      cudaStream_t s) {
    const int grid = (rows_out + 255) / 256;
    if (b_is_bf16)
        conv_bias_relayout_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(b), out, Cout,
                                                                      rows_out, d2s_perm);
    else
        conv_bias_relayout_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(b), out, Cout, rows_out,
                                                              d2s_perm);
    return done();
}


cudaError_t launch_copy_box(const void* src, int elem_bytes, int C, int sT, int sH, int sW, int st0, int sh0, int sw0,
                            void* dst, int dT, int dH, int dW, int dt0, int dh0, int dw0, int bT, int bH, int bW,
                            cudaStream_t s) {
    if (bT <= 0 || bH <= 0 || bW <= 0) return cudaSuccess;
    if (st0 < 0 || sh0 < 0 || sw0 < 0 || dt0 < 0 || dh0 < 0 || dw0 < 0 || st0 + bT > sT || sh0 + bH > sH ||
        sw0 + bW > sW || dt0 + bT > dT || dh0 + bH > dH || dw0 + bW > dW)
        return cudaErrorInvalidValue;
    const int64_t n = static_cast<int64_t>(C) * bT * bH * bW;
    const int grid = static_cast<int>((n + 255) / 256);
    const Vol sv{C, sT, sH, sW}, dv{C, dT, dH, dW};
    if (elem_bytes == 4)
        copy_box_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(src), sv, st0, sh0, sw0,
                                                    static_cast<float*>(dst), dv, dt0, dh0, dw0, bT, bH, bW);
    else if (elem_bytes == 2)
        copy_box_kernel<uint16_t><<<grid, 256, 0, s>>>(static_cast<const uint16_t*>(src), sv, st0, sh0, sw0,
                                                       static_cast<uint16_t*>(dst), dv, dt0, dh0, dw0, bT, bH, bW);
    else
        return cudaErrorInvalidValue;
    return done();
}

cudaError_t launch_blend_axis(const float* a, int aT, int aH, int aW, float* b, int bT, int bH, int bW, int C, int axis,
                              int blend_extent, cudaStream_t s) {
    if (axis < 1 || axis > 3) return cudaErrorInvalidValue;
    const int la = axis == 1 ? aT : (axis == 2 ? aH : aW), lb = axis == 1 ? bT : (axis == 2 ? bH : bW);
    int blend = blend_extent < la ? blend_extent : la;
    if (lb < blend) blend = lb;
    if (blend <= 0) return cudaSuccess;
    if ((axis != 1 && aT != bT) || (axis != 2 && aH != bH) || (axis != 3 && aW != bW)) return cudaErrorInvalidValue;
    const int eT = axis == 1 ? blend : bT, eH = axis == 2 ? blend : bH, eW = axis == 3 ? blend : bW;
    const int64_t n = static_cast<int64_t>(C) * eT * eH * eW;
    const float inv = static_cast<float>(1.0 / static_cast<double>(blend));  // affine(1/blend) of an f32 arange
    blend_axis_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, s>>>(a, Vol{C, aT, aH, aW}, b, Vol{C, bT, bH, bW}, axis,
                                                                        blend, inv);
    return done();
}

}  // namespace ltxv
