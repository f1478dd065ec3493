// Flash-style attention forward for sm_100a (non-causal; optional additive key bias).
//
// One CTA = one (batch, head, 128-query tile).  192 threads:
//   warp 0    : TMA producer  (Q once, then K_j / V_j tiles of 128 keys through 2-stage rings)
//   warp 1    : TMEM allocator + UMMA issuer:  S = Q K_j^T  (128x128xD)  ->  TMEM cols [0,128)
//                                              O += P_j V_j (128xDx128)  ->  TMEM cols [128,128+D)
//   warps 2-5 : softmax, one query row per thread (TMEM lane = row): two passes over S in TMEM
//               (row max, then exp2 / row sum), P_j written as bf16 into 128B-swizzled smem as the A operand of
//               the second MMA; O is rescaled in TMEM only when the running max grew by more than 2^8 (lazy).
// For D = 64 two CTAs are resident per SM (112 KB smem, 256 TMEM columns each) so one CTA's softmax overlaps
// the other's MMAs.
#include "attention.h"
#include "common.cuh"
#include "tensormap.h"
#include "profile.h"

#include <atomic>

namespace ltxv {

namespace {

constexpr int kTileQ = 128;
constexpr int kTileKV = 128;
constexpr int kAttnThreads = 192;
constexpr int kPBytes = kTileQ * kTileKV * 2;  // 32 KB
constexpr int kTmemColsAttn = 256;
constexpr int kOCol = 128;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int D>
struct ACfg {
    static constexpr int kQBytes = kTileQ * D * 2;
    static constexpr int kKVBytes = kTileKV * D * 2;
    static constexpr int kStages = 2;
    static constexpr int kSmemBytes = kQBytes + 2 * kStages * kKVBytes + kPBytes + 256;
    static constexpr int kAtoms = D / 64;  // 64-column (128 B) swizzle atoms per row
};

template <int D>
__global__ void __launch_bounds__(kAttnThreads, (D == 64) ? 2 : 1)
flash_attn_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                  const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ AttnParams p) {
    using C = ACfg<D>;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sq = smem;
    uint8_t* sk = sq + C::kQBytes;
    uint8_t* sv = sk + C::kStages * C::kKVBytes;
    uint8_t* sp = sv + C::kStages * C::kKVBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sp + kPBytes);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;   // [2]
    uint64_t* k_empty = bars + 3;  // [2]
    uint64_t* v_full = bars + 5;   // [2]
    uint64_t* v_empty = bars + 7;  // [2]
    uint64_t* bar_s = bars + 9;    // S_j landed in TMEM
    uint64_t* bar_p = bars + 10;   // P_j written to smem (128 arrivals)
    uint64_t* bar_pv = bars + 11;  // O += P_j V_j retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kTileQ;
    const int head = blockIdx.y;
    const int batch = blockIdx.z;
    const int n_tiles = (p.Skv + kTileKV - 1) / kTileKV;

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) {
            printf("ltxv attention: dynamic smem base not 1024B aligned\n");
            __trap();
        }
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
        }
        mbar_init(bar_s, 1);
        mbar_init(bar_p, 128);
        mbar_init(bar_pv, 1);
        fence_barrier_init();
    }
    if (warp_idx == 1) tmem_alloc<kTmemColsAttn>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp_idx == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, C::kQBytes);
#pragma unroll
            for (int a = 0; a < C::kAtoms; ++a)
                tma_load_3d(sq + a * (kTileQ * 128), &tm_q, q_full, p.q_col0 + head * D + a * 64, q0, batch);
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < n_tiles; ++j) {
                const int kv0 = j * kTileKV;
                mbar_wait(&k_empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&k_full[stage], C::kKVBytes);
#pragma unroll
                for (int a = 0; a < C::kAtoms; ++a)
                    tma_load_3d(sk + stage * C::kKVBytes + a * (kTileKV * 128), &tm_k, &k_full[stage],
                                p.k_col0 + head * D + a * 64, kv0, batch);
                mbar_wait(&v_empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&v_full[stage], C::kKVBytes);
#pragma unroll
                for (int a = 0; a < C::kAtoms; ++a)
                    tma_load_3d(sv + stage * C::kKVBytes + a * (kTileKV * 128), &tm_v, &v_full[stage],
                                p.v_col0 + head * D + a * 64, kv0, batch);
                if (++stage == C::kStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================== UMMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(kTileQ, kTileKV, false, false);
            constexpr uint32_t idesc_pv = make_idesc_bf16(kTileQ, D, false, true);  // B = V is MN-major
            const uint32_t q_addr = smem_u32(sq);
            const uint32_t p_addr = smem_u32(sp);
            const uint32_t tmem_s = tmem_base;
            const uint32_t tmem_o = tmem_base + kOCol;

            auto issue_s = [&](int stage) {
                const uint32_t k_addr = smem_u32(sk + stage * C::kKVBytes);
#pragma unroll
                for (int ks = 0; ks < D / 16; ++ks) {
                    const uint32_t off = (ks >> 2) * (kTileQ * 128) + (ks & 3) * 32;
                    umma_bf16_ss(tmem_s, make_smem_desc_sw128(q_addr + off, 1024, 0),
                                 make_smem_desc_sw128(k_addr + off, 1024, 0), idesc_s, ks != 0 ? 1u : 0u);
                }
            };

            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tcgen05_fence_after();
            issue_s(0);
            umma_commit(&k_empty[0]);
            umma_commit(bar_s);

            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < n_tiles; ++j) {
                // O += P_j V_j
                mbar_wait(bar_p, j & 1);
                mbar_wait(&v_full[stage], phase);
                tcgen05_fence_after();
                const uint32_t v_addr = smem_u32(sv + stage * C::kKVBytes);
#pragma unroll
                for (int ks = 0; ks < kTileKV / 16; ++ks) {
                    const uint64_t da = make_smem_desc_sw128(p_addr + (ks >> 2) * (kTileQ * 128) + (ks & 3) * 32, 1024, 0);
                    const uint64_t db = make_smem_desc_sw128(v_addr + ks * (16 * 128), 1024, kTileKV * 128);
                    umma_bf16_ss(tmem_o, da, db, idesc_pv, (j | ks) != 0 ? 1u : 0u);
                }
                umma_commit(&v_empty[stage]);
                umma_commit(bar_pv);
                int nstage = stage + 1;
                uint32_t nphase = phase;
                if (nstage == C::kStages) {
                    nstage = 0;
                    nphase ^= 1;
                }
                if (j + 1 < n_tiles) {
                    // S_{j+1} = Q K_{j+1}^T (softmax j has finished reading S_j: it arrived on bar_p)
                    mbar_wait(&k_full[nstage], nphase);
                    tcgen05_fence_after();
                    issue_s(nstage);
                    umma_commit(&k_empty[nstage]);
                    umma_commit(bar_s);
                }
                stage = nstage;
                phase = nphase;
            }
        }
    } else {
        // ===================== softmax warps =====================
        const int quad = warp_idx & 3;
        const int row = quad * 32 + lane;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        const uint32_t tmem_s = lane_base;
        const uint32_t tmem_o = lane_base + kOCol;
        const float c = p.scale * kLog2e;
        const float* bias = (p.kv_bias != nullptr) ? p.kv_bias + static_cast<int64_t>(batch) * p.Skv : nullptr;
        uint8_t* prow = sp + row * 128;
        const int sw = row & 7;

        float m_used = -INFINITY;
        float l = 0.f;
        for (int j = 0; j < n_tiles; ++j) {
            const int kv0 = j * kTileKV;
            const bool tail = (kv0 + kTileKV > p.Skv);
            mbar_wait(bar_s, j & 1);
            tcgen05_fence_after();
            // ---- pass 1: row max of the (scaled, biased, masked) scores, log2 domain ----
            float mx = -INFINITY;
#pragma unroll 1
            for (int cc = 0; cc < kTileKV / 32; ++cc) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_s + cc * 32, r);
                tmem_ld_wait();
                if (bias == nullptr && !tail) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int col = kv0 + cc * 32 + i;
                        float x = __uint_as_float(r[i]);
                        if (bias != nullptr && col < p.Skv) x += __ldg(bias + col) * (1.0f / p.scale);
                        if (col >= p.Skv) x = -INFINITY;
                        mx = fmaxf(mx, x);
                    }
                }
            }
            mx *= c;  // c > 0: max commutes with the positive scale
            if (j == 0) {
                m_used = mx;
            } else {
                // previous P V must have retired before O is touched or P is overwritten
                mbar_wait(bar_pv, (j - 1) & 1);
                tcgen05_fence_after();
                const float m_new = fmaxf(m_used, mx);
                const bool need = (m_new - m_used) > kRescaleThreshold;
                if (__any_sync(0xffffffffu, need)) {
                    const float alpha = ex2_approx(m_used - m_new);
#pragma unroll 1
                    for (int dc = 0; dc < D / 32; ++dc) {
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(tmem_o + dc * 32, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                        tmem_st_32x32b_x32(tmem_o + dc * 32, r);
                    }
                    tmem_st_wait();
                    l *= alpha;
                    m_used = m_new;
                }
            }
            // ---- pass 2: p = exp2(x - m_used), row sum, bf16 P into swizzled smem ----
#pragma unroll 1
            for (int cc = 0; cc < kTileKV / 32; ++cc) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_s + cc * 32, r);
                tmem_ld_wait();
                float pv[32];
                if (bias == nullptr && !tail) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) pv[i] = ex2_approx(fmaf(__uint_as_float(r[i]), c, -m_used));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int col = kv0 + cc * 32 + i;
                        float x = __uint_as_float(r[i]) * c;
                        if (bias != nullptr && col < p.Skv) x += __ldg(bias + col) * kLog2e;
                        if (col >= p.Skv) x = -INFINITY;
                        pv[i] = ex2_approx(x - m_used);
                    }
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) l += pv[i];
                // columns cc*32 .. +31 -> atom (cc>>1), 16B chunks ((cc&1)*4 + q), XOR-swizzled with the row
                uint8_t* pbase = prow + (cc >> 1) * (kTileQ * 128);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 u;
                    u.x = pack_bf16x2(pv[8 * q + 0], pv[8 * q + 1]);
                    u.y = pack_bf16x2(pv[8 * q + 2], pv[8 * q + 3]);
                    u.z = pack_bf16x2(pv[8 * q + 4], pv[8 * q + 5]);
                    u.w = pack_bf16x2(pv[8 * q + 6], pv[8 * q + 7]);
                    const int chunk = ((cc & 1) * 4 + q) ^ sw;
                    *reinterpret_cast<uint4*>(pbase + chunk * 16) = u;
                }
            }
            tcgen05_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(bar_p);
        }
        // ---- epilogue: O / l -> bf16 ----
        mbar_wait(bar_pv, (n_tiles - 1) & 1);
        tcgen05_fence_after();
        const float inv_l = 1.0f / l;
        const int qrow = q0 + row;
        __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) +
                              (static_cast<int64_t>(batch) * p.Sq + qrow) * p.ldo + head * D;
#pragma unroll 1
        for (int dc = 0; dc < D / 32; ++dc) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem_o + dc * 32, r);
            tmem_ld_wait();
            if (qrow < p.Sq) {
                uint4* d4 = reinterpret_cast<uint4*>(orow + dc * 32);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(r[8 * q + 0]) * inv_l, __uint_as_float(r[8 * q + 1]) * inv_l);
                    u.y = pack_bf16x2(__uint_as_float(r[8 * q + 2]) * inv_l, __uint_as_float(r[8 * q + 3]) * inv_l);
                    u.z = pack_bf16x2(__uint_as_float(r[8 * q + 4]) * inv_l, __uint_as_float(r[8 * q + 5]) * inv_l);
                    u.w = pack_bf16x2(__uint_as_float(r[8 * q + 6]) * inv_l, __uint_as_float(r[8 * q + 7]) * inv_l);
                    d4[q] = u;
                }
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tcgen05_fence_after();
        tmem_dealloc<kTmemColsAttn>(tmem_base);
    }
}

std::atomic<uint64_t> g_attn_launches{0};

template <int D>
cudaError_t launch_attn_impl(const AttnParams& p, cudaStream_t stream) {
    using C = ACfg<D>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e =
            cudaFuncSetAttribute(flash_attn_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    CUtensorMap tq, tk, tv;
    cudaError_t e = make_tensor_map_3d_bf16(&tq, p.q, p.B, p.Sq, p.ldq, kTileQ, 64, p.ldq, p.ldq * (int64_t)p.Sq);
    if (e != cudaSuccess) return e;
    e = make_tensor_map_3d_bf16(&tk, p.k, p.B, p.Skv, p.ldk, kTileKV, 64, p.ldk, p.ldk * (int64_t)p.Skv);
    if (e != cudaSuccess) return e;
    e = make_tensor_map_3d_bf16(&tv, p.v, p.B, p.Skv, p.ldv, kTileKV, 64, p.ldv, p.ldv * (int64_t)p.Skv);
    if (e != cudaSuccess) return e;
    dim3 grid((p.Sq + kTileQ - 1) / kTileQ, p.H, p.B);
    {
        ProfScope prof(p.kv_bias != nullptr || p.Skv != p.Sq ? PROF_ATTN_CROSS : PROF_ATTN_SELF,
                       4.0 * p.B * p.H * static_cast<double>(p.Sq) * p.Skv * D, stream);
        flash_attn_kernel<D><<<grid, kAttnThreads, C::kSmemBytes, stream>>>(tq, tk, tv, p);
    }
    g_attn_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

}  // namespace

uint64_t attention_launch_count() { return g_attn_launches.load(); }

cudaError_t launch_attention(const AttnParams& p, cudaStream_t stream) {
    if (p.B <= 0 || p.H <= 0 || p.Sq <= 0 || p.Skv <= 0) return cudaErrorInvalidValue;
    if (p.D == 64) return launch_attn_impl<64>(p, stream);
    if (p.D == 128) return launch_attn_impl<128>(p, stream);
    return cudaErrorInvalidValue;
}

}  // namespace ltxv
