// Bandwidth-bound glue kernels (see glue.h).  Rows are processed one warp per row with 128-bit accesses and
// warp-shuffle reductions; nothing here is shaped into a GEMM.
#include "common.cuh"
#include "glue.h"
#define LTXV_PDL_CLASS 4
#include "launch.h"
#include "profile.h"

#include <atomic>

namespace ltxv {

namespace {
std::atomic<uint64_t> g_glue_launches{0};
inline cudaError_t done() {
    g_glue_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}
constexpr int kWarpsPerBlock = 8;

// ------------------------------------------------------------------------------------------------
// norm + modulate: f32 row in, bf16 row out
// ------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
norm_modulate_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, const float* __restrict__ scale,
                     const float* __restrict__ shift, int rows, int D, float eps) {
    griddep_launch_dependents();
    griddep_wait();
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<int64_t>(row) * D);
    const int nv = D >> 2;
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < nv; i += 32) {
        float4 v = xr[i];
        s1 += v.x + v.y + v.z + v.w;
        s2 += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    s2 = warp_sum(s2);
    float mean = 0.f, rinv;
    if (KIND == NORM_LAYER) {
        s1 = warp_sum(s1);
        mean = s1 / D;
        // second pass for the centred variance (matches the reference's (x-mean)^2 form, :74-77)
        float sv = 0.f;
        for (int i = lane; i < nv; i += 32) {
            float4 v = xr[i];
            float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
            sv += a * a + b * b + c * c + d * d;
        }
        sv = warp_sum(sv);
        rinv = rsqrtf(sv / D + eps);
    } else {
        rinv = rsqrtf(s2 * (1.0f / D) + eps);
    }
    uint2* orow = reinterpret_cast<uint2*>(out + static_cast<int64_t>(row) * D);
    const float4* sc4 = reinterpret_cast<const float4*>(scale);
    const float4* sh4 = reinterpret_cast<const float4*>(shift);
    for (int i = lane; i < nv; i += 32) {
        float4 v = xr[i];
        float a = (v.x - mean) * rinv, b = (v.y - mean) * rinv, c = (v.z - mean) * rinv, d = (v.w - mean) * rinv;
        if (scale != nullptr) {
            float4 sc = __ldg(sc4 + i), sh = __ldg(sh4 + i);
            a = a * (1.0f + sc.x) + sh.x;
            b = b * (1.0f + sc.y) + sh.y;
            c = c * (1.0f + sc.z) + sh.z;
            d = d * (1.0f + sc.w) + sh.w;
        }
        uint2 o;
        o.x = pack_bf16x2(a, b);
        o.y = pack_bf16x2(c, d);
        orow[i] = o;
    }
}

// RMS variant with the whole row register-resident (D = 128 * VPL): one pass over x, every load of the row in flight
// before the first is consumed (the generic kernel above reads the row twice with 4 loads in flight per lane)
template <int VPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rms_modulate_row_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, const float* __restrict__ scale,
                        const float* __restrict__ shift, int rows, float eps) {
    constexpr int D = VPL * 128;
    griddep_launch_dependents();
    griddep_wait();
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<int64_t>(row) * D);
    float4 v[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) v[i] = xr[lane + 32 * i];
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) s2 += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    s2 = warp_sum(s2);
    const float rinv = rsqrtf(s2 * (1.0f / D) + eps);
    uint2* orow = reinterpret_cast<uint2*>(out + static_cast<int64_t>(row) * D);
    const float4* sc4 = reinterpret_cast<const float4*>(scale);
    const float4* sh4 = reinterpret_cast<const float4*>(shift);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int idx = lane + 32 * i;
        float a = v[i].x * rinv, b = v[i].y * rinv, c = v[i].z * rinv, d = v[i].w * rinv;
        if (scale != nullptr) {
            const float4 sc = __ldg(sc4 + idx), sh = __ldg(sh4 + idx);
            a = a * (1.0f + sc.x) + sh.x;
            b = b * (1.0f + sc.y) + sh.y;
            c = c * (1.0f + sc.z) + sh.z;
            d = d * (1.0f + sc.w) + sh.w;
        }
        uint2 o;
        o.x = pack_bf16x2(a, b);
        o.y = pack_bf16x2(c, d);
        orow[idx] = o;
    }
}

// ------------------------------------------------------------------------------------------------
// q/k RMS norm across heads (+ RoPE), in place on bf16
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
qk_norm_rope_kernel(__nv_bfloat16* __restrict__ x, int64_t ld, int col0, int rows, int D, const float* __restrict__ w,
                    float eps, const float* __restrict__ cos_t, const float* __restrict__ sin_t, int group_cols,
                    int group_w) {
    // blockIdx.y = column group (the stacked per-layer cross-attention keys): columns + y * group_cols, weight + y * group_w
    griddep_launch_dependents();
    griddep_wait();
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    w += static_cast<int64_t>(blockIdx.y) * group_w;
    uint4* xr = reinterpret_cast<uint4*>(x + static_cast<int64_t>(row) * ld + col0 + static_cast<int64_t>(blockIdx.y) * group_cols);
    const int nv = D >> 3;  // 8 bf16 per 16 B
    float s2 = 0.f;
    for (int i = lane; i < nv; i += 32) {
        uint4 u = xr[i];
        float f[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                      bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
        for (int j = 0; j < 8; ++j) s2 += f[j] * f[j];
    }
    s2 = warp_sum(s2);
    const float rinv = rsqrtf(s2 * (1.0f / D) + eps);
    const float4* w4 = reinterpret_cast<const float4*>(w);
    const float4* c4 = cos_t ? reinterpret_cast<const float4*>(cos_t + static_cast<int64_t>(row) * (D >> 1)) : nullptr;
    const float4* s4 = sin_t ? reinterpret_cast<const float4*>(sin_t + static_cast<int64_t>(row) * (D >> 1)) : nullptr;
    for (int i = lane; i < nv; i += 32) {
        uint4 u = xr[i];
        float f[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                      bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
        float4 wa = __ldg(w4 + 2 * i), wb = __ldg(w4 + 2 * i + 1);
        const float ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = f[j] * rinv * ww[j];
        if (c4 != nullptr) {
            float4 cc = __ldg(c4 + i), ss = __ldg(s4 + i);  // 4 pairs
            const float cv[4] = {cc.x, cc.y, cc.z, cc.w}, sv[4] = {ss.x, ss.y, ss.z, ss.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float re = f[2 * j], im = f[2 * j + 1];
                f[2 * j] = re * cv[j] - im * sv[j];      // x*cos + (-x_imag)*sin
                f[2 * j + 1] = im * cv[j] + re * sv[j];  // x*cos + ( x_real)*sin
            }
        }
        uint4 o;
        o.x = pack_bf16x2(f[0], f[1]);
        o.y = pack_bf16x2(f[2], f[3]);
        o.z = pack_bf16x2(f[4], f[5]);
        o.w = pack_bf16x2(f[6], f[7]);
        xr[i] = o;
    }
}

// q AND k of the fused self-attention projection in one launch (x = [rows, >= 2D] with q at col 0, k at col D): one
// pass over the token's cos/sin row serves both (the two separate launches read the 2 x 20 MB f32 tables twice)
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
qk_pair_norm_rope_kernel(__nv_bfloat16* __restrict__ x, int64_t ld, int rows, int D, const float* __restrict__ wq,
                         const float* __restrict__ wk, float eps, const float* __restrict__ cos_t,
                         const float* __restrict__ sin_t, int table_rows) {
    griddep_launch_dependents();
    griddep_wait();
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int nv = D >> 3;
    uint4* xq = reinterpret_cast<uint4*>(x + static_cast<int64_t>(row) * ld);
    uint4* xk = reinterpret_cast<uint4*>(x + static_cast<int64_t>(row) * ld + D);
    const int trow = row % table_rows;  // batch entries share one [table_rows, D/2] table
    const float4* c4 = reinterpret_cast<const float4*>(cos_t + static_cast<int64_t>(trow) * (D >> 1));
    const float4* s4 = reinterpret_cast<const float4*>(sin_t + static_cast<int64_t>(trow) * (D >> 1));
    float sq = 0.f, sk = 0.f;
    for (int i = lane; i < nv; i += 32) {
        const uint4 a = xq[i], b = xk[i];
        const float fa[8] = {bf16_lo(a.x), bf16_hi(a.x), bf16_lo(a.y), bf16_hi(a.y),
                             bf16_lo(a.z), bf16_hi(a.z), bf16_lo(a.w), bf16_hi(a.w)};
        const float fb[8] = {bf16_lo(b.x), bf16_hi(b.x), bf16_lo(b.y), bf16_hi(b.y),
                             bf16_lo(b.z), bf16_hi(b.z), bf16_lo(b.w), bf16_hi(b.w)};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sq += fa[j] * fa[j];
            sk += fb[j] * fb[j];
        }
    }
    sq = warp_sum(sq);
    sk = warp_sum(sk);
    const float rq = rsqrtf(sq * (1.0f / D) + eps), rk = rsqrtf(sk * (1.0f / D) + eps);
    for (int i = lane; i < nv; i += 32) {
        const float4 cc = __ldg(c4 + i), ss = __ldg(s4 + i);
        const float cv[4] = {cc.x, cc.y, cc.z, cc.w}, sv[4] = {ss.x, ss.y, ss.z, ss.w};
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            uint4* xr = which == 0 ? xq : xk;
            const float* w = which == 0 ? wq : wk;
            const float rinv = which == 0 ? rq : rk;
            uint4 u = xr[i];
            float f[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                          bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
            const float4 wa = __ldg(reinterpret_cast<const float4*>(w) + 2 * i);
            const float4 wb = __ldg(reinterpret_cast<const float4*>(w) + 2 * i + 1);
            const float ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = f[j] * rinv * ww[j];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float re = f[2 * j], im = f[2 * j + 1];
                f[2 * j] = re * cv[j] - im * sv[j];
                f[2 * j + 1] = im * cv[j] + re * sv[j];
            }
            u.x = pack_bf16x2(f[0], f[1]);
            u.y = pack_bf16x2(f[2], f[3]);
            u.z = pack_bf16x2(f[4], f[5]);
            u.w = pack_bf16x2(f[6], f[7]);
            xr[i] = u;
        }
    }
}

// register-resident variant of qk_pair_norm_rope_kernel for D = 256 * VPL: q and k of the token are loaded once (all 2*VPL
// 16-byte loads in flight), normed, rotated and stored; one pass over the activations and over the cos/sin row
template <int VPL, bool SHARE>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
qk_pair_norm_rope_row_kernel(__nv_bfloat16* __restrict__ x, int64_t ld, int rows, const float* __restrict__ wq,
                             const float* __restrict__ wk, float eps, const float* __restrict__ cos_t,
                             const float* __restrict__ sin_t, int table_rows) {
    // SHARE: one warp per TABLE row: with the two CFG branches batched (rows = 2 * table_rows) the warp keeps the
    // cos/sin row in registers and applies it to token t of both branches, so the f32 tables cross HBM once instead of
    // twice.  !SHARE (D = 4096: the tables do not fit next to a row): one warp per activation row.
    constexpr int D = VPL * 256;
    griddep_launch_dependents();
    griddep_wait();
    const int w_row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w_row >= (SHARE ? table_rows : rows)) return;
    const int trow = SHARE ? w_row : w_row % table_rows;
    const int row_step = SHARE ? table_rows : rows;
    const float4* c4 = reinterpret_cast<const float4*>(cos_t + static_cast<int64_t>(trow) * (D >> 1));
    const float4* s4 = reinterpret_cast<const float4*>(sin_t + static_cast<int64_t>(trow) * (D >> 1));
    float4 cc[SHARE ? VPL : 1], ss[SHARE ? VPL : 1];
    uint4 q[VPL], k[VPL];
    {
        const uint4* xq = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(w_row) * ld);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            q[i] = xq[lane + 32 * i];
            k[i] = xq[lane + 32 * i + (D >> 3)];
        }
    }
    if (SHARE) {
#pragma unroll
        for (int i = 0; i < (SHARE ? VPL : 1); ++i) {
            cc[i] = __ldg(c4 + lane + 32 * i);
            ss[i] = __ldg(s4 + lane + 32 * i);
        }
    }
    for (int row = w_row; row < rows; row += row_step) {
        uint4* xq = reinterpret_cast<uint4*>(x + static_cast<int64_t>(row) * ld);
        uint4* xk = xq + (D >> 3);
        float sq = 0.f, sk = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const float fa[8] = {bf16_lo(q[i].x), bf16_hi(q[i].x), bf16_lo(q[i].y), bf16_hi(q[i].y),
                                 bf16_lo(q[i].z), bf16_hi(q[i].z), bf16_lo(q[i].w), bf16_hi(q[i].w)};
            const float fb[8] = {bf16_lo(k[i].x), bf16_hi(k[i].x), bf16_lo(k[i].y), bf16_hi(k[i].y),
                                 bf16_lo(k[i].z), bf16_hi(k[i].z), bf16_lo(k[i].w), bf16_hi(k[i].w)};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                sq += fa[j] * fa[j];
                sk += fb[j] * fb[j];
            }
        }
        sq = warp_sum(sq);
        sk = warp_sum(sk);
        const float rq = rsqrtf(sq * (1.0f / D) + eps), rk = rsqrtf(sk * (1.0f / D) + eps);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int idx = lane + 32 * i;
            const float4 c_i = SHARE ? cc[SHARE ? i : 0] : __ldg(c4 + idx), s_i = SHARE ? ss[SHARE ? i : 0] : __ldg(s4 + idx);
            const float cv[4] = {c_i.x, c_i.y, c_i.z, c_i.w}, sv[4] = {s_i.x, s_i.y, s_i.z, s_i.w};
#pragma unroll
            for (int which = 0; which < 2; ++which) {
                uint4 u = which == 0 ? q[i] : k[i];
                const float* w = which == 0 ? wq : wk;
                const float rinv = which == 0 ? rq : rk;
                float f[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                              bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
                const float4 wa = __ldg(reinterpret_cast<const float4*>(w) + 2 * idx);
                const float4 wb = __ldg(reinterpret_cast<const float4*>(w) + 2 * idx + 1);
                const float ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = f[j] * rinv * ww[j];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float re = f[2 * j], im = f[2 * j + 1];
                    f[2 * j] = re * cv[j] - im * sv[j];
                    f[2 * j + 1] = im * cv[j] + re * sv[j];
                }
                u.x = pack_bf16x2(f[0], f[1]);
                u.y = pack_bf16x2(f[2], f[3]);
                u.z = pack_bf16x2(f[4], f[5]);
                u.w = pack_bf16x2(f[6], f[7]);
                (which == 0 ? xq : xk)[idx] = u;
            }
        }
        if (SHARE && row + row_step < rows) {
            const uint4* nq = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(row + row_step) * ld);
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                q[i] = nq[lane + 32 * i];
                k[i] = nq[lane + 32 * i + (D >> 3)];
            }
        }
    }
}

// register-resident RMS norm x weight of ONE tensor, in place (the cross-attention queries: no RoPE), D = 256 * VPL
template <int VPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rms_weight_row_kernel(__nv_bfloat16* __restrict__ x, int64_t ld, int col0, int rows, const float* __restrict__ w, float eps,
                      int group_cols, int group_w) {
    constexpr int D = VPL * 256;
    griddep_launch_dependents();
    griddep_wait();
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    w += static_cast<int64_t>(blockIdx.y) * group_w;  // blockIdx.y = column group, see qk_norm_rope_kernel
    uint4* xr = reinterpret_cast<uint4*>(x + static_cast<int64_t>(row) * ld + col0 + static_cast<int64_t>(blockIdx.y) * group_cols);
    uint4 q[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) q[i] = xr[lane + 32 * i];
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const float f[8] = {bf16_lo(q[i].x), bf16_hi(q[i].x), bf16_lo(q[i].y), bf16_hi(q[i].y),
                            bf16_lo(q[i].z), bf16_hi(q[i].z), bf16_lo(q[i].w), bf16_hi(q[i].w)};
#pragma unroll
        for (int j = 0; j < 8; ++j) s2 += f[j] * f[j];
    }
    s2 = warp_sum(s2);
    const float rinv = rsqrtf(s2 * (1.0f / D) + eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int idx = lane + 32 * i;
        float f[8] = {bf16_lo(q[i].x), bf16_hi(q[i].x), bf16_lo(q[i].y), bf16_hi(q[i].y),
                      bf16_lo(q[i].z), bf16_hi(q[i].z), bf16_lo(q[i].w), bf16_hi(q[i].w)};
        const float4 wa = __ldg(reinterpret_cast<const float4*>(w) + 2 * idx);
        const float4 wb = __ldg(reinterpret_cast<const float4*>(w) + 2 * idx + 1);
        const float ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = f[j] * rinv * ww[j];
        uint4 o;
        o.x = pack_bf16x2(f[0], f[1]);
        o.y = pack_bf16x2(f[2], f[3]);
        o.z = pack_bf16x2(f[4], f[5]);
        o.w = pack_bf16x2(f[6], f[7]);
        xr[idx] = o;
    }
}

// Ulysses scatter fused with the q/k norm + RoPE (see glue.h)
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
qkv_norm_rope_scatter_kernel(const __nv_bfloat16* __restrict__ x, int rows, int D, int nranks, int row0,
                             const float* __restrict__ wq, const float* __restrict__ wk, float eps,
                             const float* __restrict__ cos_t, const float* __restrict__ sin_t, ScatterDst dst) {
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int nv = D >> 3;
    const int Dg = D / nranks;
    const int64_t drow = static_cast<int64_t>(row0 + row) * 3 * Dg;
    const float4* c4 = reinterpret_cast<const float4*>(cos_t + static_cast<int64_t>(row) * (D >> 1));
    const float4* s4 = reinterpret_cast<const float4*>(sin_t + static_cast<int64_t>(row) * (D >> 1));
#pragma unroll 1
    for (int which = 0; which < 3; ++which) {
        const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(row) * 3 * D + which * D);
        float rinv = 1.0f;
        const float* w = which == 0 ? wq : wk;
        if (which < 2) {
            float s2 = 0.f;
            for (int i = lane; i < nv; i += 32) {
                uint4 u = xr[i];
                float f[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                              bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
                for (int j = 0; j < 8; ++j) s2 += f[j] * f[j];
            }
            s2 = warp_sum(s2);
            rinv = rsqrtf(s2 * (1.0f / D) + eps);
        }
        for (int i = lane; i < nv; i += 32) {
            uint4 u = xr[i];
            if (which < 2) {
                float f[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                              bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
                const float4 wa = __ldg(reinterpret_cast<const float4*>(w) + 2 * i);
                const float4 wb = __ldg(reinterpret_cast<const float4*>(w) + 2 * i + 1);
                const float ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = f[j] * rinv * ww[j];
                const float4 cc = __ldg(c4 + i), ss = __ldg(s4 + i);
                const float cv[4] = {cc.x, cc.y, cc.z, cc.w}, sv[4] = {ss.x, ss.y, ss.z, ss.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float re = f[2 * j], im = f[2 * j + 1];
                    f[2 * j] = re * cv[j] - im * sv[j];
                    f[2 * j + 1] = im * cv[j] + re * sv[j];
                }
                u.x = pack_bf16x2(f[0], f[1]);
                u.y = pack_bf16x2(f[2], f[3]);
                u.z = pack_bf16x2(f[4], f[5]);
                u.w = pack_bf16x2(f[6], f[7]);
            }
            const int col = i * 8;
            const int g = col / Dg;
            __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst.p[g]) + drow + which * Dg + (col - g * Dg);
            *reinterpret_cast<uint4*>(d) = u;  // NVLink peer store (or local when g == rank)
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Deferred k RMS-norm (EPI_QKV_ROPE producers, gemm.h): k arrives as w * k (rotated) without its per-row factor; one
// in-place pass multiplies every k row by rsqrt(mean(k^2) + eps), the mean taken from the epilogue's sums of squares.
// 4 B / element (the pass it replaces moved 8 B / element of q and k plus the f32 RoPE table).
// ------------------------------------------------------------------------------------------------
// sum of the n <= 64 group sums of a row, by a whole warp: lane l takes entries l and l + 32, then a xor-shuffle tree.
// Every kernel that needs a row factor goes through this one function, so a row's factor has the same bits everywhere.
__device__ __forceinline__ float warp_row_sumsq(const float* __restrict__ ss, int n, int lane) {
    float t = lane < n ? __ldg(ss + lane) : 0.f;
    if (lane + 32 < n) t += __ldg(ss + lane + 32);
    return warp_sum(t);
}
__device__ __forceinline__ uint4 scale_bf16x8(uint4 u, float r) {
    u.x = pack_bf16x2(bf16_lo(u.x) * r, bf16_hi(u.x) * r);
    u.y = pack_bf16x2(bf16_lo(u.y) * r, bf16_hi(u.y) * r);
    u.z = pack_bf16x2(bf16_lo(u.z) * r, bf16_hi(u.z) * r);
    u.w = pack_bf16x2(bf16_lo(u.w) * r, bf16_hi(u.w) * r);
    return u;
}
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_rms_scale_kernel(__nv_bfloat16* __restrict__ x, int64_t ld, int col0, int rows, int D, const float* __restrict__ ss,
                   int ss_ld, int ss_n, float eps, float* __restrict__ q_rscale) {
    griddep_launch_dependents();
    griddep_wait();
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* srow = ss + static_cast<int64_t>(row) * ss_ld;  // ss_n groups of q, then ss_n groups of k
    const float invD = 1.0f / static_cast<float>(D);
    const float rq = rsqrtf(warp_row_sumsq(srow, ss_n, lane) * invD + eps);
    const float r = rsqrtf(warp_row_sumsq(srow + ss_n, ss_n, lane) * invD + eps);
    if (lane == 0) q_rscale[row] = rq;
    uint4* xr = reinterpret_cast<uint4*>(x + static_cast<int64_t>(row) * ld + col0);
    const int nv = D >> 3;
    // all loads of the row in flight before the first store (D = 2048: 8 x 16 B per lane)
    for (int i0 = lane; i0 < nv; i0 += 8 * 32) {
        uint4 u[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (i0 + k * 32 < nv) u[k] = xr[i0 + k * 32];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (i0 + k * 32 < nv) xr[i0 + k * 32] = scale_bf16x8(u[k], r);
    }
}

// factor of rows whose tensor needs no pass of its own (cross-attention queries): out[row] = rsqrt(sum / D + eps)
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
row_rscale_kernel(const float* __restrict__ ss, int ss_ld, int ss_n, int rows, int D, float eps, float* __restrict__ out) {
    griddep_launch_dependents();
    griddep_wait();
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float r = rsqrtf(warp_row_sumsq(ss + static_cast<int64_t>(row) * ss_ld, ss_n, lane) / static_cast<float>(D) + eps);
    if (lane == 0) out[row] = r;
}

// Ulysses scatter of the fused-epilogue layout: q (as is), k (times its row factor, same arithmetic as
// k_rms_scale_kernel) and v are copied head-group-wise into the peers' [S_total, 3 D / nranks] buffers; the row's q
// factor goes to EVERY peer's q_rscale[S_total] (the attention of each head group needs the factor of all its queries).
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
qkv_scatter_scaled_kernel(const __nv_bfloat16* __restrict__ x, int rows, int D, int nranks, int row0,
                          const float* __restrict__ ss, int ss_ld, int ss_n, float eps, ScatterDst dst, ScatterDst q_rs_dst) {
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int nv = D >> 3;
    const int Dg = D / nranks;
    const int64_t drow = static_cast<int64_t>(row0 + row) * 3 * Dg;
    const float* srow = ss + static_cast<int64_t>(row) * ss_ld;
    const float invD = 1.0f / static_cast<float>(D);
    const float rq = rsqrtf(warp_row_sumsq(srow, ss_n, lane) * invD + eps);
    const float rk = rsqrtf(warp_row_sumsq(srow + ss_n, ss_n, lane) * invD + eps);
    if (lane < nranks) reinterpret_cast<float*>(q_rs_dst.p[lane])[row0 + row] = rq;
#pragma unroll 1
    for (int which = 0; which < 3; ++which) {
        const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(row) * 3 * D + which * D);
        for (int i = lane; i < nv; i += 32) {
            uint4 u = xr[i];
            if (which == 1) u = scale_bf16x8(u, rk);
            const int col = i * 8;
            const int g = col / Dg;
            __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst.p[g]) + drow + which * Dg + (col - g * Dg);
            *reinterpret_cast<uint4*>(d) = u;  // NVLink peer store (or local when g == rank)
        }
    }
}

// ------------------------------------------------------------------------------------------------
// RoPE table
// ------------------------------------------------------------------------------------------------
__global__ void rope_table_kernel(const float* __restrict__ coords, int F, int H, int W, float m0, float m1, float m2,
                                  int S, int D, float theta_ln, float* __restrict__ cos_t, float* __restrict__ sin_t,
                                  int token0) {
    const int half = D >> 1;
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<int64_t>(S) * half) return;
    const int s = static_cast<int>(idx / half);
    const int pi = static_cast<int>(idx - static_cast<int64_t>(s) * half);
    const int steps = D / 6;
    const int rem_pairs = (D % 6) >> 1;
    float c = 1.0f, sn = 0.0f;
    if (pi >= rem_pairs) {
        const int q = pi - rem_pairs;  // = j*3 + a
        const int j = q / 3, a = q - j * 3;
        float g;
        if (coords != nullptr) {
            g = coords[static_cast<int64_t>(s) * 3 + a] * (a == 0 ? m0 : (a == 1 ? m1 : m2));
        } else {
            const int sg = s + token0;  // global token index (sequence-parallel shards start at token0)
            const int w = sg % W, h = (sg / W) % H, f = sg / (W * H);
            const float v = static_cast<float>(a == 0 ? f : (a == 1 ? h : w));
            g = v * (a == 0 ? m0 : (a == 1 ? m1 : m2));
        }
        const float lin = (steps <= 1) ? 0.0f : __fmul_rn(static_cast<float>(j), 1.0f / static_cast<float>(steps - 1));
        const float freq = __fmul_rn(expf(__fmul_rn(lin, theta_ln)), 1.5707963267948966f);
        const float gs = __fadd_rn(__fmul_rn(g, 2.0f), -1.0f);
        const float ang = __fmul_rn(gs, freq);
        sincosf(ang, &sn, &c);
    }
    cos_t[idx] = c;
    sin_t[idx] = sn;
}

// ------------------------------------------------------------------------------------------------
// GEMV (weight-read bound): one warp per output row
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
gemv_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ w, const float* __restrict__ bias,
            float* __restrict__ y, int N, int K, int act_in, int act_out) {
    const int n = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    const uint4* wr = reinterpret_cast<const uint4*>(w + static_cast<int64_t>(n) * K);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float acc = 0.f;
    for (int i = lane; i < (K >> 3); i += 32) {
        uint4 u = __ldg(wr + i);
        float4 a = x4[2 * i], b = x4[2 * i + 1];
        float xv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        if (act_in == GEMV_SILU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) xv[j] = xv[j] / (1.0f + expf(-xv[j]));
        }
        acc += xv[0] * bf16_lo(u.x) + xv[1] * bf16_hi(u.x) + xv[2] * bf16_lo(u.y) + xv[3] * bf16_hi(u.y) +
               xv[4] * bf16_lo(u.z) + xv[5] * bf16_hi(u.z) + xv[6] * bf16_lo(u.w) + xv[7] * bf16_hi(u.w);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        float v = acc + (bias ? bias[n] : 0.f);
        if (act_out == GEMV_SILU) v = v / (1.0f + expf(-v));
        y[n] = v;
    }
}

__global__ void sinusoid_kernel(const float* __restrict__ t_dev, const float* __restrict__ t_mul, float* __restrict__ out,
                                int style, int round_bf16) {
    const int i = threadIdx.x;  // 0..127
    float t = t_dev[0];
    if (t_mul != nullptr) t *= t_mul[0];
    if (round_bf16) t = __bfloat162float(__float2bfloat16(t));
    float fr;
    if (style == 0) {
        fr = 1.0f / powf(10000.0f, static_cast<float>(i) / 128.0f);
    } else {
        fr = expf(static_cast<float>(i) * (-9.210340371976184f / 128.0f));
    }
    const float a = t * fr;
    out[i] = cosf(a);
    out[128 + i] = sinf(a);
}

__global__ void add_vec_kernel(const float* a, const float* b, float* d, int n, int period) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = a[i] + b[i % period];
}

__global__ void mask_bias_kernel(const float* m, float* b, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = __fmul_rn(__fadd_rn(__fmul_rn(m[i], -1.0f), 1.0f), -10000.0f);
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
    int64_t i = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) * 4;
    if (i + 3 < n) {
        float4 v = *reinterpret_cast<const float4*>(src + i);
        uint2 o;
        o.x = pack_bf16x2(v.x, v.y);
        o.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(dst + i) = o;
    } else {
        for (; i < n; ++i) dst[i] = __float2bfloat16(src[i]);
    }
}
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int64_t n) {
    int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i < n) dst[i] = __bfloat162float(src[i]);
}
__global__ void blend_kernel(float* __restrict__ x, const float* __restrict__ orig, float m, int64_t n) {
    int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i < n) x[i] = x[i] * (1.0f - m) + orig[i] * m;
}

// ------------------------------------------------------------------------------------------------
// pack / unpack latents, coords
// ------------------------------------------------------------------------------------------------
// out[s, d], s = (f2*H2 + h2)*W2 + w2, d = ((c*pt + a)*p + i)*p + j  <-  in[c, f2*pt+a, h2*p+i, w2*p+j]
template <bool PACK>
__global__ void pack_unpack_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int F, int H, int W,
                                   int p, int pt) {
    const int F2 = F / pt, H2 = H / p, W2 = W / p;
    const int Dd = C * pt * p * p;
    const int64_t total = static_cast<int64_t>(F2) * H2 * W2 * Dd;
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= total) return;
    // idx enumerates the *packed* layout [S, Dd]
    const int d = static_cast<int>(idx % Dd);
    const int s = static_cast<int>(idx / Dd);
    const int j = d % p, i = (d / p) % p, a = (d / (p * p)) % pt, c = d / (p * p * pt);
    const int w2 = s % W2, h2 = (s / W2) % H2, f2 = s / (W2 * H2);
    const int64_t src = ((static_cast<int64_t>(c) * F + (f2 * pt + a)) * H + (h2 * p + i)) * W + (w2 * p + j);
    if (PACK) out[idx] = in[src];
    else out[src] = in[idx];
}

// p = pt = 1 fast path: [C, N] <-> [N, C] tiled transpose through shared memory (coalesced both ways)
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
    __shared__ float tile[32][33];
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 32 + threadIdx.y;
    for (int k = 0; k < 32; k += 8)
        if (x < cols && y + k < rows) tile[threadIdx.y + k][threadIdx.x] = in[static_cast<int64_t>(y + k) * cols + x];
    __syncthreads();
    x = blockIdx.y * 32 + threadIdx.x;
    y = blockIdx.x * 32 + threadIdx.y;
    for (int k = 0; k < 32; k += 8)
        if (x < rows && y + k < cols) out[static_cast<int64_t>(y + k) * rows + x] = tile[threadIdx.x][threadIdx.y + k];
}

__global__ void video_coords_kernel(float* __restrict__ out, int F, int H, int W, float ts_ratio, float sp_ratio,
                                    float inv_fps) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= F * H * W) return;
    const int w = s % W, h = (s / W) % H, f = s / (W * H);
    // affine(ts, 1-ts) -> clamp(0,1000) -> affine(1/fps)   (each an individually rounded f32 op)
    float vf = __fadd_rn(__fmul_rn(static_cast<float>(f), ts_ratio), 1.0f - ts_ratio);
    vf = fminf(fmaxf(vf, 0.0f), 1000.0f);
    vf = __fmul_rn(vf, inv_fps);
    out[s * 3 + 0] = vf;
    out[s * 3 + 1] = __fmul_rn(static_cast<float>(h), sp_ratio);
    out[s * 3 + 2] = __fmul_rn(static_cast<float>(w), sp_ratio);
}

// ------------------------------------------------------------------------------------------------
// guidance combine + Euler
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float combine_cfg(float c, float u, bool has_u, float g) {
    // uncond + (cond - uncond).affine(g, 0): three individually rounded f32 ops (t2v_pipeline.rs:947-949)
    return has_u ? __fadd_rn(u, __fmul_rn(__fsub_rn(c, u), g)) : c;
}
// pass 1 (only when guidance_rescale > 0): sums for the unbiased std of cond and of the CFG combination
__global__ void guidance_stats_kernel(const float* __restrict__ cond, const float* __restrict__ uncond, int64_t n,
                                      float g, double* __restrict__ acc) {
    double sc = 0, scc = 0, sm = 0, smm = 0;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float c = cond[i];
        const float m = combine_cfg(c, uncond[i], true, g);
        sc += c; scc += static_cast<double>(c) * c;
        sm += m; smm += static_cast<double>(m) * m;
    }
    __shared__ double red[4][32];
    double v[4] = {sc, scc, sm, smm};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0;
        for (int wi = 0; wi < (blockDim.x >> 5); ++wi) t += red[threadIdx.x][wi];
        atomicAdd(&acc[threadIdx.x], t);
    }
}
__global__ void guidance_euler_kernel(const float* __restrict__ cond, const float* __restrict__ uncond,
                                      const float* __restrict__ pert, float* __restrict__ latents,
                                      float* __restrict__ noise_out, int64_t n, float g, float r, float s_stg, float dt,
                                      StatParts parts) {
    float ratio = 1.0f;
    if (r > 0.0f && uncond != nullptr) {
        // the std runs over ALL non-batch elements (t2v_pipeline.rs:209-224): with token shards the per-rank partial
        // sums are added here in rank order, so every rank of the group derives the same ratio
        double a[4] = {0, 0, 0, 0};
        for (int k = 0; k < parts.n; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) a[j] += parts.p[k][j];
        const double nn = static_cast<double>(parts.n_total);
        const double var_c = (a[1] - a[0] * a[0] / nn) / (nn - 1.0);
        const double var_m = (a[3] - a[2] * a[2] / nn) / (nn - 1.0);
        ratio = static_cast<float>(sqrt(var_c) / sqrt(var_m));
    }
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const float c = cond[i];
    float m = combine_cfg(c, uncond ? uncond[i] : 0.f, uncond != nullptr, g);
    if (r > 0.0f && uncond != nullptr)
        m = __fadd_rn(__fmul_rn(__fmul_rn(m, ratio), r), __fmul_rn(m, 1.0f - r));
    if (pert != nullptr) m = __fadd_rn(m, __fmul_rn(__fsub_rn(c, pert[i]), s_stg));
    if (noise_out != nullptr) noise_out[i] = m;
    if (latents != nullptr) latents[i] = __fadd_rn(latents[i], __fmul_rn(m, dt));
}

// stochastic branch of FlowMatchEulerDiscreteScheduler::step (scheduler.rs:557-575), noise supplied by the caller:
//   x0 = x - sigma * v ;  x <- (1 - sigma_next) * x0 + sigma_next * noise      (every op rounded to f32 like the tensor ops)
__global__ void stochastic_step_kernel(float* __restrict__ latents, const float* __restrict__ v,
                                       const float* __restrict__ noise, float sigma, float sigma_next, int64_t n) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const float x0 = __fsub_rn(latents[i], __fmul_rn(sigma, v[i]));
    const float om = __fadd_rn(__fmul_rn(sigma_next, -1.0f), 1.0f);  // ns.affine(-1, 1)
    latents[i] = __fadd_rn(__fmul_rn(om, x0), __fmul_rn(sigma_next, noise[i]));
}
// decode-noise blend (t2v_pipeline.rs:1055-1062): x <- x * (1 - s) + noise * s
__global__ void noise_blend_kernel(float* __restrict__ x, const float* __restrict__ noise, float scale, int64_t n) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const float om = __fadd_rn(__fmul_rn(scale, -1.0f), 1.0f);
    x[i] = __fadd_rn(__fmul_rn(x[i], om), __fmul_rn(noise[i], scale));
}

__global__ void denorm_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ mean,
                              const float* __restrict__ std, float inv_sf, int C, int64_t n_per_c) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= C * n_per_c) return;
    const int c = static_cast<int>(i / n_per_c);
    out[i] = __fadd_rn(__fmul_rn(__fmul_rn(in[i], std[c]), inv_sf), mean[c]);
}
__global__ void postprocess_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n) {
    const int64_t i = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) * 4;
    if (i + 3 < n) {
        float4 v = *reinterpret_cast<const float4*>(in + i);
        v.x = fminf(fmaxf(__fmaf_rn(v.x, 0.5f, 0.5f), 0.f), 1.f) * 255.f;
        v.y = fminf(fmaxf(__fmaf_rn(v.y, 0.5f, 0.5f), 0.f), 1.f) * 255.f;
        v.z = fminf(fmaxf(__fmaf_rn(v.z, 0.5f, 0.5f), 0.f), 1.f) * 255.f;
        v.w = fminf(fmaxf(__fmaf_rn(v.w, 0.5f, 0.5f), 0.f), 1.f) * 255.f;
        *reinterpret_cast<float4*>(out + i) = v;
    } else {
        for (int64_t k = i; k < n; ++k) out[k] = fminf(fmaxf(__fmaf_rn(in[k], 0.5f, 0.5f), 0.f), 1.f) * 255.f;
    }
}

inline int blocks_for(int64_t n, int per_block) { return static_cast<int>((n + per_block - 1) / per_block); }

}  // namespace

uint64_t glue_launch_count() { return g_glue_launches.load(); }

cudaError_t launch_norm_modulate(const float* x, void* out, const float* scale, const float* shift, int rows, int D,
                                 float eps, int kind, cudaStream_t s) {
    if (D % 4 != 0 || (scale == nullptr) != (shift == nullptr)) return cudaErrorInvalidValue;
    const int grid = blocks_for(rows, kWarpsPerBlock);
    ProfScope prof(PROF_NORM_MOD, 6.0 * rows * D, s);  // f32 in, bf16 out
    LTXV_TRACE_VARIANT(kind == NORM_RMS ? (D == 2048 ? "rms_modulate_row_kernel<16>" : D == 4096 ? "rms_modulate_row_kernel<32>" : "norm_modulate_kernel<RMS>") : "norm_modulate_kernel<LAYER>");
    if (kind == NORM_RMS && D == 2048)
        launch_pdl(rms_modulate_row_kernel<16>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, s, x,
                   reinterpret_cast<__nv_bfloat16*>(out), scale, shift, rows, eps);
    else if (kind == NORM_RMS && D == 4096)
        launch_pdl(rms_modulate_row_kernel<32>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, s, x,
                   reinterpret_cast<__nv_bfloat16*>(out), scale, shift, rows, eps);
    else if (kind == NORM_LAYER)
        launch_pdl(norm_modulate_kernel<NORM_LAYER>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, s, x,
                   reinterpret_cast<__nv_bfloat16*>(out), scale, shift, rows, D, eps);
    else
        launch_pdl(norm_modulate_kernel<NORM_RMS>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, s, x,
                   reinterpret_cast<__nv_bfloat16*>(out), scale, shift, rows, D, eps);
    return done();
}

cudaError_t launch_qk_norm_rope(void* x, int64_t ld, int col0, int rows, int D, const float* w, float eps,
                                const float* cos_t, const float* sin_t, cudaStream_t s, int groups, int group_cols,
                                int group_w) {
    if (D % 8 != 0 || ld % 8 != 0 || col0 % 8 != 0 || groups < 1 || groups > 65535 || group_cols % 8 != 0 || group_w % 4 != 0 ||
        (groups > 1 && cos_t != nullptr))
        return cudaErrorInvalidValue;
    ProfScope prof(PROF_QK_ROPE, groups * (4.0 * rows * D + (cos_t ? 2.0 * rows * (D / 2) * 4 : 0.0)), s);
    LTXV_TRACE_VARIANT(cos_t == nullptr && D == 2048 ? "rms_weight_row_kernel<8>" : cos_t == nullptr && D == 4096 ? "rms_weight_row_kernel<16>" : "qk_norm_rope_kernel");
    const dim3 grid(blocks_for(rows, kWarpsPerBlock), groups);
    if (cos_t == nullptr && D == 2048)
        launch_pdl(rms_weight_row_kernel<8>, grid, dim3(kWarpsPerBlock * 32), 0, s, reinterpret_cast<__nv_bfloat16*>(x), ld,
                   col0, rows, w, eps, group_cols, group_w);
    else if (cos_t == nullptr && D == 4096)
        launch_pdl(rms_weight_row_kernel<16>, grid, dim3(kWarpsPerBlock * 32), 0, s, reinterpret_cast<__nv_bfloat16*>(x), ld,
                   col0, rows, w, eps, group_cols, group_w);
    else
        launch_pdl(qk_norm_rope_kernel, grid, dim3(kWarpsPerBlock * 32), 0, s, reinterpret_cast<__nv_bfloat16*>(x), ld, col0,
                   rows, D, w, eps, cos_t, sin_t, group_cols, group_w);
    return done();
}

cudaError_t launch_qk_pair_norm_rope(void* x, int64_t ld, int rows, int D, const float* wq, const float* wk, float eps,
                                     const float* cos_t, const float* sin_t, cudaStream_t s, int table_rows) {
    if (D % 8 != 0 || ld % 8 != 0 || ld < 2 * D || cos_t == nullptr || sin_t == nullptr) return cudaErrorInvalidValue;
    if (table_rows <= 0) table_rows = rows;
    // q,k bf16 in+out for every row; the f32 cos/sin rows once per TABLE row on the register-resident path
    const bool share = D == 2048 && rows % table_rows == 0 && rows > table_rows;
    ProfScope prof(PROF_QK_ROPE, 2.0 * rows * D * 4 + 2.0 * (share ? table_rows : rows) * (D / 2) * 4, s);
    LTXV_TRACE_VARIANT(share ? "qk_pair_norm_rope_row_kernel<8,true>" : D == 2048 ? "qk_pair_norm_rope_row_kernel<8,false>" : D == 4096 ? "qk_pair_norm_rope_row_kernel<16,false>" : "qk_pair_norm_rope_kernel");
    if (share)
        launch_pdl(qk_pair_norm_rope_row_kernel<8, true>, dim3(blocks_for(table_rows, kWarpsPerBlock)), dim3(kWarpsPerBlock * 32), 0, s,
                   reinterpret_cast<__nv_bfloat16*>(x), ld, rows, wq, wk, eps, cos_t, sin_t, table_rows);
    else if (D == 2048)
        launch_pdl(qk_pair_norm_rope_row_kernel<8, false>, dim3(blocks_for(rows, kWarpsPerBlock)), dim3(kWarpsPerBlock * 32), 0, s,
                   reinterpret_cast<__nv_bfloat16*>(x), ld, rows, wq, wk, eps, cos_t, sin_t, table_rows);
    else if (D == 4096)
        launch_pdl(qk_pair_norm_rope_row_kernel<16, false>, dim3(blocks_for(rows, kWarpsPerBlock)), dim3(kWarpsPerBlock * 32), 0, s,
                   reinterpret_cast<__nv_bfloat16*>(x), ld, rows, wq, wk, eps, cos_t, sin_t, table_rows);
    else
        launch_pdl(qk_pair_norm_rope_kernel, dim3(blocks_for(rows, kWarpsPerBlock)), dim3(kWarpsPerBlock * 32), 0, s,
                   reinterpret_cast<__nv_bfloat16*>(x), ld, rows, D, wq, wk, eps, cos_t, sin_t, table_rows);
    return done();
}

cudaError_t launch_qkv_norm_rope_scatter(const void* x, int rows, int D, int nranks, int row0, const float* wq,
                                         const float* wk, float eps, const float* cos_t, const float* sin_t,
                                         const ScatterDst& dst, cudaStream_t s) {
    if (D % (8 * nranks) != 0 || cos_t == nullptr || sin_t == nullptr) return cudaErrorInvalidValue;
    qkv_norm_rope_scatter_kernel<<<blocks_for(rows, kWarpsPerBlock), kWarpsPerBlock * 32, 0, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), rows, D, nranks, row0, wq, wk, eps, cos_t, sin_t, dst);
    return done();
}

cudaError_t launch_k_rms_scale(void* x, int64_t ld, int col0, int rows, int D, const float* ss, int ss_ld, int ss_n,
                               float eps, float* q_rscale, cudaStream_t s) {
    if (D % 8 != 0 || ld % 8 != 0 || col0 % 8 != 0 || ss == nullptr || q_rscale == nullptr || ss_n <= 0 || ss_n > 64)
        return cudaErrorInvalidValue;
    ProfScope prof(PROF_QK_ROPE, 4.0 * rows * D, s);  // bf16 in + out
    LTXV_TRACE_VARIANT("k_rms_scale_kernel");
    launch_pdl(k_rms_scale_kernel, dim3(blocks_for(rows, kWarpsPerBlock)), dim3(kWarpsPerBlock * 32), 0, s,
               reinterpret_cast<__nv_bfloat16*>(x), ld, col0, rows, D, ss, ss_ld, ss_n, eps, q_rscale);
    return done();
}

cudaError_t launch_row_rscale(const float* ss, int ss_ld, int ss_n, int rows, int D, float eps, float* out,
                              cudaStream_t s) {
    if (ss == nullptr || out == nullptr || ss_n <= 0 || ss_n > 64) return cudaErrorInvalidValue;
    ProfScope prof(PROF_QK_ROPE, 4.0 * rows * (ss_n + 1), s);
    LTXV_TRACE_VARIANT("row_rscale_kernel");
    launch_pdl(row_rscale_kernel, dim3(blocks_for(rows, kWarpsPerBlock)), dim3(kWarpsPerBlock * 32), 0, s, ss, ss_ld, ss_n,
               rows, D, eps, out);
    return done();
}

cudaError_t launch_qkv_scatter_scaled(const void* x, int rows, int D, int nranks, int row0, const float* ss, int ss_ld,
                                      int ss_n, float eps, const ScatterDst& dst, const ScatterDst& q_rs_dst,
                                      cudaStream_t s) {
    if (D % (8 * nranks) != 0 || ss == nullptr || nranks > 8 || ss_n <= 0 || ss_n > 64) return cudaErrorInvalidValue;
    LTXV_TRACE_VARIANT("qkv_scatter_scaled_kernel");
    qkv_scatter_scaled_kernel<<<blocks_for(rows, kWarpsPerBlock), kWarpsPerBlock * 32, 0, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), rows, D, nranks, row0, ss, ss_ld, ss_n, eps, dst, q_rs_dst);
    return done();
}

cudaError_t launch_rope_table(const float* coords, int F, int H, int W, const float* scale3, int S, int D, float theta,
                              float* cos_t, float* sin_t, cudaStream_t s, int token0) {
    float m0, m1, m2;
    if (coords != nullptr) {  // :454-460, multiply by f32(1/base)
        m0 = static_cast<float>(1.0 / static_cast<double>(20.0f));
        m1 = static_cast<float>(1.0 / static_cast<double>(2048.0f));
        m2 = m1;
    } else if (scale3 != nullptr) {
        m0 = scale3[0]; m1 = scale3[1]; m2 = scale3[2];
    } else {
        m0 = m1 = m2 = 1.0f;
    }
    const int64_t n = static_cast<int64_t>(S) * (D / 2);
    rope_table_kernel<<<blocks_for(n, 256), 256, 0, s>>>(coords, F, H, W, m0, m1, m2, S, D,
                                                        static_cast<float>(log(static_cast<double>(theta))), cos_t, sin_t, token0);
    return done();
}

cudaError_t launch_gemv(const float* x, const void* w, const float* bias, float* y, int N, int K, int act_in,
                        int act_out, cudaStream_t s) {
    if (K % 8 != 0) return cudaErrorInvalidValue;
    gemv_kernel<<<blocks_for(N, kWarpsPerBlock), kWarpsPerBlock * 32, 0, s>>>(
        x, reinterpret_cast<const __nv_bfloat16*>(w), bias, y, N, K, act_in, act_out);
    return done();
}

cudaError_t launch_sinusoid(const float* t_dev, const float* t_mul, float* out256, int style, int round_bf16,
                            cudaStream_t s) {
    sinusoid_kernel<<<1, 128, 0, s>>>(t_dev, t_mul, out256, style, round_bf16);
    return done();
}

cudaError_t launch_add_vec(const float* a, const float* b, float* d, int n, int period, cudaStream_t s) {
    add_vec_kernel<<<blocks_for(n, 256), 256, 0, s>>>(a, b, d, n, period);
    return done();
}

cudaError_t launch_mask_bias(const float* mask, float* bias, int n, cudaStream_t s) {
    mask_bias_kernel<<<blocks_for(n, 256), 256, 0, s>>>(mask, bias, n);
    return done();
}

cudaError_t launch_f32_to_bf16(const float* src, void* dst, int64_t n, cudaStream_t s) {
    f32_to_bf16_kernel<<<blocks_for((n + 3) / 4, 256), 256, 0, s>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n);
    return done();
}
cudaError_t launch_bf16_to_f32(const void* src, float* dst, int64_t n, cudaStream_t s) {
    bf16_to_f32_kernel<<<blocks_for(n, 256), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(src), dst, n);
    return done();
}
cudaError_t launch_blend(float* x, const float* orig, float m, int64_t n, cudaStream_t s) {
    blend_kernel<<<blocks_for(n, 256), 256, 0, s>>>(x, orig, m, n);
    return done();
}

cudaError_t launch_pack_latents(const float* in, float* out, int C, int F, int H, int W, int p, int pt, cudaStream_t s) {
    if (p <= 0 || pt <= 0 || F % pt || H % p || W % p) return cudaErrorInvalidValue;
    if (p == 1 && pt == 1) {
        const int rows = C, cols = F * H * W;  // [C, N] -> [N, C]
        transpose_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, s>>>(in, out, rows, cols);
    } else {
        const int64_t n = static_cast<int64_t>(C) * F * H * W;
        pack_unpack_kernel<true><<<blocks_for(n, 256), 256, 0, s>>>(in, out, C, F, H, W, p, pt);
    }
    return done();
}
cudaError_t launch_unpack_latents(const float* in, float* out, int C, int F, int H, int W, int p, int pt,
                                  cudaStream_t s) {
    // F,H,W here are the *unpacked* (output) dims
    if (p <= 0 || pt <= 0 || F % pt || H % p || W % p) return cudaErrorInvalidValue;
    if (p == 1 && pt == 1) {
        const int rows = F * H * W, cols = C;  // [N, C] -> [C, N]
        transpose_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, s>>>(in, out, rows, cols);
    } else {
        const int64_t n = static_cast<int64_t>(C) * F * H * W;
        pack_unpack_kernel<false><<<blocks_for(n, 256), 256, 0, s>>>(in, out, C, F, H, W, p, pt);
    }
    return done();
}
cudaError_t launch_video_coords(float* out, int F, int H, int W, int ts_ratio, int sp_ratio, int fps, cudaStream_t s) {
    video_coords_kernel<<<blocks_for(static_cast<int64_t>(F) * H * W, 256), 256, 0, s>>>(
        out, F, H, W, static_cast<float>(ts_ratio), static_cast<float>(sp_ratio),
        static_cast<float>(1.0 / static_cast<double>(fps)));
    return done();
}

cudaError_t launch_guidance_stats(const float* cond, const float* uncond, int64_t n, float g, double* acc4,
                                  cudaStream_t s) {
    if (acc4 == nullptr || uncond == nullptr) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(acc4, 0, 4 * sizeof(double), s);
    if (e != cudaSuccess) return e;
    int grid = blocks_for(n, 256 * 8);
    if (grid > 1184) grid = 1184;
    guidance_stats_kernel<<<grid, 256, 0, s>>>(cond, uncond, n, g, acc4);
    return done();
}

cudaError_t launch_guidance_euler_parts(const float* cond, const float* uncond, const float* pert, float* latents,
                                        float* noise_out, int64_t n, float g, float r, float s_stg, float dt,
                                        const StatParts& parts, cudaStream_t s) {
    if (r > 0.0f && uncond != nullptr && (parts.n < 1 || parts.n > 8 || parts.n_total < 2)) return cudaErrorInvalidValue;
    guidance_euler_kernel<<<blocks_for(n, 256), 256, 0, s>>>(cond, uncond, pert, latents, noise_out, n, g, r, s_stg, dt,
                                                            parts);
    return done();
}

cudaError_t launch_guidance_euler(const float* cond, const float* uncond, const float* pert, float* latents,
                                  float* noise_out, int64_t n, float g, float r, float s_stg, float dt, double* scratch,
                                  cudaStream_t s) {
    StatParts parts{};
    if (r > 0.0f && uncond != nullptr) {
        cudaError_t e = launch_guidance_stats(cond, uncond, n, g, scratch, s);
        if (e != cudaSuccess) return e;
        parts.p[0] = scratch;
        parts.n = 1;
        parts.n_total = n;
    }
    return launch_guidance_euler_parts(cond, uncond, pert, latents, noise_out, n, g, r, s_stg, dt, parts, s);
}

cudaError_t launch_stochastic_step(float* latents, const float* v, const float* noise, float sigma, float sigma_next,
                                   int64_t n, cudaStream_t s) {
    if (latents == nullptr || v == nullptr || noise == nullptr) return cudaErrorInvalidValue;
    stochastic_step_kernel<<<blocks_for(n, 256), 256, 0, s>>>(latents, v, noise, sigma, sigma_next, n);
    return done();
}
cudaError_t launch_noise_blend(float* x, const float* noise, float scale, int64_t n, cudaStream_t s) {
    if (x == nullptr || noise == nullptr) return cudaErrorInvalidValue;
    noise_blend_kernel<<<blocks_for(n, 256), 256, 0, s>>>(x, noise, scale, n);
    return done();
}

cudaError_t launch_denormalize(const float* in, float* out, const float* mean, const float* std, float inv_sf, int C,
                               int64_t n_per_c, cudaStream_t s) {
    denorm_kernel<<<blocks_for(C * n_per_c, 256), 256, 0, s>>>(in, out, mean, std, inv_sf, C, n_per_c);
    return done();
}
// frames f32 [B, 3, F, H, W] in 0..255 -> u8 [B, F, H, W, 3]: the per-frame permute((1,2,0)).clamp(0,255).to_dtype(U8)
// of the reference's output hand-off (examples/ltx-video/main.rs:653-667; the cast truncates toward zero).
__global__ void frames_to_u8_kernel(const float* __restrict__ in, uint8_t* __restrict__ out, int F, int64_t hw,
                                    int64_t n_px) {
    const int64_t px = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;  // over B * F * H * W
    if (px >= n_px) return;
    const int64_t fhw = static_cast<int64_t>(F) * hw;
    const int64_t b = px / fhw, r = px - b * fhw;  // r = f * hw + pixel: the offset inside one colour plane
    const float* src = in + b * 3 * fhw + r;
    uint8_t* dst = out + px * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) dst[c] = static_cast<uint8_t>(fminf(fmaxf(src[c * fhw], 0.f), 255.f));
}
cudaError_t launch_frames_to_u8(const float* in, uint8_t* out, int B, int F, int H, int W, cudaStream_t s) {
    const int64_t hw = static_cast<int64_t>(H) * W;
    const int64_t n = static_cast<int64_t>(B) * F * hw;
    if (n <= 0) return cudaSuccess;
    frames_to_u8_kernel<<<blocks_for(n, 256), 256, 0, s>>>(in, out, F, hw, n);
    return done();
}
cudaError_t launch_postprocess(const float* in, float* out, int64_t n, cudaStream_t s) {
    postprocess_kernel<<<blocks_for((n + 3) / 4, 256), 256, 0, s>>>(in, out, n);
    return done();
}

}  // namespace ltxv
