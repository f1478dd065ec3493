#include "common.cuh"
#include "vae_glue.h"

#include <atomic>

namespace ltxv {

namespace {
std::atomic<uint64_t> g_vae_glue_launches{0};
inline cudaError_t done() {
    g_vae_glue_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

template <typename TIn>
__global__ void vae_input_kernel(const TIn* __restrict__ z, __nv_bfloat16* __restrict__ out, int C, int F, int H, int W) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t n = static_cast<int64_t>(F) * H * W * C;
    if (idx >= n) return;
    const int c = static_cast<int>(idx % C);
    const int64_t vox = idx / C;
    const int w = static_cast<int>(vox % W), h = static_cast<int>((vox / W) % H), f = static_cast<int>(vox / (W * H));
    const float v = static_cast<float>(z[((static_cast<int64_t>(c) * F + f) * H + h) * W + w]);
    const __nv_bfloat16 b = __float2bfloat16(v);
    const int Hp = H + 2, Wp = W + 2;
    auto at = [&](int tp) { return ((static_cast<int64_t>(tp) * Hp + (h + 1)) * Wp + (w + 1)) * C + c; };
    out[at(f + 1)] = b;
    if (f == 0) out[at(0)] = b;
    if (f == F - 1) out[at(F + 1)] = b;
}

// one warp per voxel; each lane owns C/32 contiguous channels
template <int C>
__global__ void __launch_bounds__(256)
vae_prep_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, const float* __restrict__ scale,
                const float* __restrict__ shift, int do_norm, int do_silu, int T, int H, int W) {
    constexpr int VPL = C / 32;  // values per lane: 4, 8, 16, 32
    const int64_t vox = blockIdx.x * 8ll + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int64_t nvox = static_cast<int64_t>(T) * H * W;
    if (vox >= nvox) return;
    const __nv_bfloat16* src = x + vox * C + lane * VPL;
    float v[VPL];
    if constexpr (VPL == 4) {
        uint2 u = *reinterpret_cast<const uint2*>(src);
        v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
    } else {
#pragma unroll
        for (int q = 0; q < VPL / 8; ++q) {
            uint4 u = reinterpret_cast<const uint4*>(src)[q];
            v[8 * q + 0] = bf16_lo(u.x); v[8 * q + 1] = bf16_hi(u.x);
            v[8 * q + 2] = bf16_lo(u.y); v[8 * q + 3] = bf16_hi(u.y);
            v[8 * q + 4] = bf16_lo(u.z); v[8 * q + 5] = bf16_hi(u.z);
            v[8 * q + 6] = bf16_lo(u.w); v[8 * q + 7] = bf16_hi(u.w);
        }
    }
    if (do_norm) {
        float s2 = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) s2 += v[i] * v[i];
        s2 = warp_sum(s2);
        const float rinv = rsqrtf(s2 * (1.0f / C) + 1e-8f);
#pragma unroll
        for (int i = 0; i < VPL; ++i) v[i] *= rinv;
    }
    if (scale != nullptr) {
#pragma unroll
        for (int i = 0; i < VPL; ++i) v[i] = v[i] * (1.0f + __ldg(scale + lane * VPL + i)) + __ldg(shift + lane * VPL + i);
    }
    if (do_silu) {
#pragma unroll
        for (int i = 0; i < VPL; ++i) v[i] = silu_f32(v[i]);
    }
    const int w = static_cast<int>(vox % W), h = static_cast<int>((vox / W) % H), t = static_cast<int>(vox / (W * H));
    const int Hp = H + 2, Wp = W + 2;
    uint32_t pk[VPL / 2];
#pragma unroll
    for (int i = 0; i < VPL / 2; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    auto store = [&](int tp) {
        __nv_bfloat16* dst = out + ((static_cast<int64_t>(tp) * Hp + (h + 1)) * Wp + (w + 1)) * C + lane * VPL;
        if constexpr (VPL == 4) {
            *reinterpret_cast<uint2*>(dst) = make_uint2(pk[0], pk[1]);
        } else {
#pragma unroll
            for (int q = 0; q < VPL / 8; ++q)
                reinterpret_cast<uint4*>(dst)[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
    };
    store(t + 1);
    if (t == 0) store(0);
    if (t == T - 1) store(T + 1);
}

template <typename TIn>
__global__ void conv_weight_relayout_kernel(const TIn* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout,
                                            int Cin, int rows_out, int d2s_perm) {
    // out[r, tap*Cin + c]
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
    if (r < Cout) v = static_cast<float>(w[(static_cast<int64_t>(co) * Cin + c) * 27 + tap]);
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

}  // namespace

uint64_t vae_glue_launch_count() { return g_vae_glue_launches.load(); }

cudaError_t launch_vae_input(const void* z, int z_is_bf16, void* out, int C, int F, int H, int W, cudaStream_t s) {
    const int64_t n = static_cast<int64_t>(F) * H * W * C;
    const int grid = static_cast<int>((n + 255) / 256);
    if (z_is_bf16)
        vae_input_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(z),
                                                             reinterpret_cast<__nv_bfloat16*>(out), C, F, H, W);
    else
        vae_input_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(z),
                                                     reinterpret_cast<__nv_bfloat16*>(out), C, F, H, W);
    return done();
}

cudaError_t launch_vae_prep(const void* x, void* out, const float* scale, const float* shift, int do_norm, int do_silu,
                            int T, int H, int W, int C, cudaStream_t s) {
    const int64_t nvox = static_cast<int64_t>(T) * H * W;
    const int grid = static_cast<int>((nvox + 7) / 8);
    const __nv_bfloat16* xi = reinterpret_cast<const __nv_bfloat16*>(x);
    __nv_bfloat16* xo = reinterpret_cast<__nv_bfloat16*>(out);
    switch (C) {
        case 128: vae_prep_kernel<128><<<grid, 256, 0, s>>>(xi, xo, scale, shift, do_norm, do_silu, T, H, W); break;
        case 256: vae_prep_kernel<256><<<grid, 256, 0, s>>>(xi, xo, scale, shift, do_norm, do_silu, T, H, W); break;
        case 512: vae_prep_kernel<512><<<grid, 256, 0, s>>>(xi, xo, scale, shift, do_norm, do_silu, T, H, W); break;
        case 1024: vae_prep_kernel<1024><<<grid, 256, 0, s>>>(xi, xo, scale, shift, do_norm, do_silu, T, H, W); break;
        default: return cudaErrorInvalidValue;
    }
    return done();
}

cudaError_t launch_conv_weight_relayout(const void* w, int w_is_bf16, void* out, int Cout, int Cin, int rows_out,
                                        int d2s_perm, cudaStream_t s) {
    const int64_t n = static_cast<int64_t>(rows_out) * 27 * Cin;
    const int grid = static_cast<int>((n + 255) / 256);
    if (w_is_bf16)
        conv_weight_relayout_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(
            reinterpret_cast<const __nv_bfloat16*>(w), reinterpret_cast<__nv_bfloat16*>(out), Cout, Cin, rows_out, d2s_perm);
    else
        conv_weight_relayout_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(w),
                                                                reinterpret_cast<__nv_bfloat16*>(out), Cout, Cin,
                                                                rows_out, d2s_perm);
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

}  // namespace ltxv
