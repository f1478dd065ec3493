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
#include <type_traits>
#include <stdlib.h>

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
#ifdef LTXV_ATTN_EXPERIMENT_NO_MUFU
    return fmaf(x, 1e-3f, 1.0f);  // timing experiment only: removes the MUFU work
#else
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}


__device__ __forceinline__ __nv_bfloat16* attn_out_row(const AttnParams& p, int batch, int qrow, int head, int D) {
    if (p.out_rows_per_peer > 0) {
        const int owner = qrow / p.out_rows_per_peer;
        const int lr = qrow - owner * p.out_rows_per_peer;
        return reinterpret_cast<__nv_bfloat16*>(p.out_peer[owner]) + static_cast<int64_t>(lr) * p.ldo + p.out_col0 + head * D;
    }
    return reinterpret_cast<__nv_bfloat16*>(p.out) + (static_cast<int64_t>(batch) * p.Sq + qrow) * p.ldo + head * D;
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
        if (elect_one()) {  // one elected lane: ptxas keeps the tcgen05/TMA operands on the uniform datapath
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
        if (elect_one()) {  // one elected lane: ptxas keeps the tcgen05/TMA operands on the uniform datapath
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
        __nv_bfloat16* orow = attn_out_row(p, batch, qrow < p.Sq ? qrow : 0, head, D);
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


// ================================================================================================
// v2 (head_dim 64, long key sequences): one CTA per (batch, head, 256 queries) = two 128-row query tiles that share
// every K/V tile.  384 threads, 1 CTA / SM:
//   warp 0     : TMA producer (Q0,Q1 once; K_j / V_j through 3-stage rings)
//   warp 1     : TMEM allocator + UMMA issuer: S_t = Q_t K_j^T -> TMEM cols [128t, 128t+128),
//                O_t += P_t V_j -> TMEM cols [256+64t, +64)
//   warps 4-7  : softmax of query tile 0      warps 8-11: softmax of query tile 1
// Each softmax thread pulls its whole 128-wide score row into registers with four back-to-back tcgen05.ld and
// immediately hands the S buffer back (s_free), so the tensor pipe computes S(j+1) while the exponentials of tile j
// run; the only steady-state limiter left is the MUFU pipe (128 exp2 per row per key tile, 16 exp2/clk/SM).
// ================================================================================================
#ifdef LTXV_ATTN_TIMING
__device__ long long g_attn_timing[32];
#define TMARK(idx)                                                           \
    do {                                                                     \
        if (tm_on) {                                                         \
            long long now_ = clock64();                                      \
            tm_acc[idx] += now_ - tm_last;                                   \
            tm_last = now_;                                                  \
        }                                                                    \
    } while (0)
#else
#define TMARK(idx) do { } while (0)
#endif
constexpr int kV2Threads = 384;  // warpgroup 0: control (TMA, MMA), warpgroups 1,2: softmax of query tile 0 / 1
constexpr int kV2Stages = 3;
constexpr int kV2QBytes = kTileQ * 64 * 2;   // 16 KB per query tile
constexpr int kV2KVBytes = kTileKV * 64 * 2;  // 16 KB per K or V tile
constexpr int kV2SmemBytes = 2 * kV2QBytes + 2 * kV2Stages * kV2KVBytes + 2 * kPBytes + 256;

// exp2 of 32 scores (scaled, shifted), row-sum, bf16 pack into the swizzled P tile
template <bool TAIL>
__device__ __forceinline__ void softmax_chunk(uint32_t (&r)[32], float c, float neg_m, int kv_col0, int skv,
                                              float& l0, float& l1, uint8_t* pbase, int chunk0, int sw) {
    float pv[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) pv[i] = ex2_approx(fmaf(__uint_as_float(r[i]), c, neg_m));
    if (TAIL) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (kv_col0 + i >= skv) pv[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
        l0 += pv[i];
        l1 += pv[i + 1];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint4 u;
        u.x = pack_bf16x2(pv[8 * q + 0], pv[8 * q + 1]);
        u.y = pack_bf16x2(pv[8 * q + 2], pv[8 * q + 3]);
        u.z = pack_bf16x2(pv[8 * q + 4], pv[8 * q + 5]);
        u.w = pack_bf16x2(pv[8 * q + 6], pv[8 * q + 7]);
        *reinterpret_cast<uint4*>(pbase + (((chunk0 + q) ^ sw) << 4)) = u;
    }
}
template <bool TAIL>
__device__ __forceinline__ float max32(const uint32_t (&r)[32], int kv_col0, int skv) {
    float a = -INFINITY, b = -INFINITY, c = -INFINITY, d = -INFINITY;
    if (!TAIL) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            a = fmaxf(a, __uint_as_float(r[i]));
            b = fmaxf(b, __uint_as_float(r[i + 1]));
            c = fmaxf(c, __uint_as_float(r[i + 2]));
            d = fmaxf(d, __uint_as_float(r[i + 3]));
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (kv_col0 + i < skv) a = fmaxf(a, __uint_as_float(r[i]));
    }
    return fmaxf(fmaxf(a, b), fmaxf(c, d));
}

__global__ void __launch_bounds__(kV2Threads, 1)
flash_attn2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ AttnParams p) {
    constexpr int D = 64;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sq = smem;
    uint8_t* sk = sq + 2 * kV2QBytes;
    uint8_t* sv = sk + kV2Stages * kV2KVBytes;
    uint8_t* sp = sv + kV2Stages * kV2KVBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sp + 2 * kPBytes);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;                   // [3]
    uint64_t* k_empty = bars + 1 + kV2Stages;      // [3]
    uint64_t* v_full = bars + 1 + 2 * kV2Stages;   // [3]
    uint64_t* v_empty = bars + 1 + 3 * kV2Stages;  // [3]
    uint64_t* s_full = bars + 1 + 4 * kV2Stages;   // [2]  S_t landed in TMEM
    uint64_t* s_free = s_full + 2;                 // [2]  S_t copied to registers (128 arrivals)
    uint64_t* p_full = s_free + 2;                 // [2]  P_t written to smem (128 arrivals)
    uint64_t* pv_done = p_full + 2;                // [2]  O_t += P_t V retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (2 * kTileQ);
    const int head = blockIdx.y;
    const int batch = blockIdx.z;
    const int n_tiles = (p.Skv + kTileKV - 1) / kTileKV;
    const bool two = (q0 + kTileQ) < p.Sq;  // second query tile has rows

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) {
            printf("ltxv attention v2: dynamic smem base not 1024B aligned\n");
            __trap();
        }
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
        mbar_init(q_full, 1);
        for (int i = 0; i < kV2Stages; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            mbar_init(&s_free[t], 128);
            mbar_init(&p_full[t], 128);
            mbar_init(&pv_done[t], 1);
        }
        fence_barrier_init();
    }
    if (warp_idx == 1) tmem_alloc<512>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp_idx == 0) {
        if (elect_one()) {  // one elected lane: ptxas keeps the tcgen05/TMA operands on the uniform datapath
            mbar_arrive_expect_tx(q_full, (two ? 2 : 1) * kV2QBytes);
            tma_load_3d(sq, &tm_q, q_full, p.q_col0 + head * D, q0, batch);
            if (two) tma_load_3d(sq + kV2QBytes, &tm_q, q_full, p.q_col0 + head * D, q0 + kTileQ, batch);
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < n_tiles; ++j) {
                const int kv0 = j * kTileKV;
                mbar_wait(&k_empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&k_full[stage], kV2KVBytes);
                tma_load_3d(sk + stage * kV2KVBytes, &tm_k, &k_full[stage], p.k_col0 + head * D, kv0, batch);
                mbar_wait(&v_empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&v_full[stage], kV2KVBytes);
                tma_load_3d(sv + stage * kV2KVBytes, &tm_v, &v_full[stage], p.v_col0 + head * D, kv0, batch);
                if (++stage == kV2Stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp_idx == 1) {
        if (elect_one()) {  // one elected lane: ptxas keeps the tcgen05/TMA operands on the uniform datapath
            constexpr uint32_t idesc_s = make_idesc_bf16(kTileQ, kTileKV, false, false);
            constexpr uint32_t idesc_pv = make_idesc_bf16(kTileQ, D, false, true);
            const uint32_t q_addr = smem_u32(sq);
            const uint32_t p_addr = smem_u32(sp);
            const int nt = two ? 2 : 1;

            auto issue_s = [&](int t, int stage) {
                const uint32_t k_addr = smem_u32(sk + stage * kV2KVBytes);
                const uint32_t qa = q_addr + t * kV2QBytes;
#pragma unroll
                for (int ks = 0; ks < D / 16; ++ks)
                    umma_bf16_ss(tmem_base + t * 128, make_smem_desc_sw128(qa + ks * 32, 1024, 0),
                                 make_smem_desc_sw128(k_addr + ks * 32, 1024, 0), idesc_s, ks != 0 ? 1u : 0u);
            };
            auto issue_pv = [&](int t, int stage, bool first) {
                const uint32_t v_addr = smem_u32(sv + stage * kV2KVBytes);
                const uint32_t pa = p_addr + t * kPBytes;
#pragma unroll
                for (int ks = 0; ks < kTileKV / 16; ++ks) {
                    const uint64_t da = make_smem_desc_sw128(pa + (ks >> 2) * (kTileQ * 128) + (ks & 3) * 32, 1024, 0);
                    const uint64_t db = make_smem_desc_sw128(v_addr + ks * (16 * 128), 1024, kTileKV * 128);
                    umma_bf16_ss(tmem_base + 256 + t * 64, da, db, idesc_pv, (!first || ks != 0) ? 1u : 0u);
                }
            };

            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tcgen05_fence_after();
            for (int t = 0; t < nt; ++t) {
                issue_s(t, 0);
                umma_commit(&s_full[t]);
            }
            umma_commit(&k_empty[0]);
#ifdef LTXV_ATTN_TIMING
            const bool tm_on = (blockIdx.x == 3 && blockIdx.y == 5);
            long long tm_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            long long tm_last = clock64();
#endif
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < n_tiles; ++j) {
                int nstage = stage + 1;
                uint32_t nphase = phase;
                if (nstage == kV2Stages) {
                    nstage = 0;
                    nphase ^= 1;
                }
                if (j + 1 < n_tiles) {
                    // S_t(j+1) as soon as the softmax group holds S_t(j) in registers: overlaps with its exponentials
                    mbar_wait(&k_full[nstage], nphase);
                    TMARK(0);
                    for (int t = 0; t < nt; ++t) {
                        mbar_wait(&s_free[t], j & 1);
                        TMARK(1 + t);
                        tcgen05_fence_after();
                        issue_s(t, nstage);
                        umma_commit(&s_full[t]);
                        TMARK(3);
                    }
                    umma_commit(&k_empty[nstage]);
                }
                mbar_wait(&v_full[stage], phase);
                TMARK(4);
                for (int t = 0; t < nt; ++t) {
                    mbar_wait(&p_full[t], j & 1);
                    TMARK(5 + t);
                    tcgen05_fence_after();
                    issue_pv(t, stage, j == 0);
                    umma_commit(&pv_done[t]);
                    TMARK(7);
                }
                umma_commit(&v_empty[stage]);
                stage = nstage;
                phase = nphase;
            }
#ifdef LTXV_ATTN_TIMING
            if (tm_on)
                for (int i = 0; i < 8; ++i) g_attn_timing[16 + i] = tm_acc[i];
#endif
        }
    } else if (warp_idx >= 4) {
        const int t = (warp_idx - 4) >> 2;  // query tile of this softmax group
        if (t == 0 || two) {
            const int quad = warp_idx & 3;
            const int row = quad * 32 + lane;
            const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
            const uint32_t tmem_s = lane_base + t * 128;
            const uint32_t tmem_o = lane_base + 256 + t * 64;
            const float c = p.scale * kLog2e;
            uint8_t* prow = sp + t * kPBytes + row * 128;
            const int sw = row & 7;
            float m_used = -INFINITY;
            float l0 = 0.f, l1 = 0.f;
#ifdef LTXV_ATTN_TIMING
            const bool tm_on = (blockIdx.x == 3 && blockIdx.y == 5 && lane == 0 && (warp_idx == 4 || warp_idx == 8));
            long long tm_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            long long tm_last = clock64();
#endif
            // ping-pong between the two softmax groups (named barriers 1 + t): the MUFU-heavy exponential phases of
            // the two query tiles alternate instead of colliding, so each runs at the full 16 exp2/clk/SM while the
            // other group does its TMEM load / max / waits.
            if (two && t == 1) named_bar_arrive(1, 256);  // tile 0 goes first
            auto tile = [&](int j, auto tail_tag) {
                constexpr bool TAIL = decltype(tail_tag)::value;
                const int kv0 = j * kTileKV;
                uint32_t s0[32], s1[32], s2[32], s3[32];
                mbar_wait(&s_full[t], j & 1);
                TMARK(0);
                tcgen05_fence_after();
                tmem_ld_32x32b_x32(tmem_s + 0, s0);
                tmem_ld_32x32b_x32(tmem_s + 32, s1);
                tmem_ld_32x32b_x32(tmem_s + 64, s2);
                tmem_ld_32x32b_x32(tmem_s + 96, s3);
                tmem_ld_wait();
                tcgen05_fence_before();
                mbar_arrive(&s_free[t]);  // S_t may be overwritten by the next QK^T
                TMARK(1);
                float mx = fmaxf(fmaxf(max32<TAIL>(s0, kv0, p.Skv), max32<TAIL>(s1, kv0 + 32, p.Skv)),
                                 fmaxf(max32<TAIL>(s2, kv0 + 64, p.Skv), max32<TAIL>(s3, kv0 + 96, p.Skv)));
                mx *= c;
                TMARK(2);
                if (j == 0) {
                    m_used = mx;
                } else {
                    // previous P_t V must have retired before O_t is rescaled or P_t is overwritten
                    mbar_wait(&pv_done[t], (j - 1) & 1);
                    TMARK(3);
                    tcgen05_fence_after();
                    const float m_new = fmaxf(m_used, mx);
                    if (__any_sync(0xffffffffu, (m_new - m_used) > kRescaleThreshold)) {
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
                        l0 *= alpha;
                        l1 *= alpha;
                        m_used = m_new;
                    }
                    TMARK(4);
                }
                const float neg_m = -m_used;
                if (two) named_bar_sync(1 + t, 256);  // my turn on the MUFU pipe
                softmax_chunk<TAIL>(s0, c, neg_m, kv0, p.Skv, l0, l1, prow, 0, sw);
                softmax_chunk<TAIL>(s1, c, neg_m, kv0 + 32, p.Skv, l0, l1, prow, 4, sw);
                softmax_chunk<TAIL>(s2, c, neg_m, kv0 + 64, p.Skv, l0, l1, prow + kTileQ * 128, 0, sw);
                softmax_chunk<TAIL>(s3, c, neg_m, kv0 + 96, p.Skv, l0, l1, prow + kTileQ * 128, 4, sw);
                if (two) named_bar_arrive(1 + (t ^ 1), 256);  // hand the MUFU pipe to the other tile
                TMARK(5);
                tcgen05_fence_before();
                fence_proxy_async_smem();
                mbar_arrive(&p_full[t]);
                TMARK(6);
            };
            const int n_full = p.Skv / kTileKV;  // tiles without a ragged tail
            for (int j = 0; j < n_full; ++j) tile(j, std::false_type{});
            if (n_full < n_tiles) tile(n_full, std::true_type{});
#ifdef LTXV_ATTN_TIMING
            if (tm_on)
                for (int i = 0; i < 8; ++i) g_attn_timing[(warp_idx == 4 ? 0 : 8) + i] = tm_acc[i];
#endif
            mbar_wait(&pv_done[t], (n_tiles - 1) & 1);
            tcgen05_fence_after();
            const float inv_l = 1.0f / (l0 + l1);
            const int qrow = q0 + t * kTileQ + row;
            __nv_bfloat16* orow = attn_out_row(p, batch, qrow < p.Sq ? qrow : 0, head, D);
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
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

std::atomic<uint64_t> g_attn_launches{0};

cudaError_t launch_attn2_impl(const AttnParams& p, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e =
            cudaFuncSetAttribute(flash_attn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kV2SmemBytes);
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
    dim3 grid((p.Sq + 2 * kTileQ - 1) / (2 * kTileQ), p.H, p.B);
    {
        ProfScope prof(PROF_ATTN_SELF, 4.0 * p.B * p.H * static_cast<double>(p.Sq) * p.Skv * 64, stream);
        flash_attn2_kernel<<<grid, kV2Threads, kV2SmemBytes, stream>>>(tq, tk, tv, p);
    }
    g_attn_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

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
#ifdef LTXV_ATTN_TIMING
void attention_debug_timing(long long* out32) { cudaMemcpyFromSymbol(out32, g_attn_timing, sizeof(long long) * 32); }
#endif

cudaError_t launch_attention(const AttnParams& p, cudaStream_t stream) {
    if (p.B <= 0 || p.H <= 0 || p.Sq <= 0 || p.Skv <= 0) return cudaErrorInvalidValue;
    if (p.D == 64 && p.kv_bias == nullptr && p.Skv > 2 * kTileKV && p.Sq > kTileQ && getenv("LTXV_ATTN_V1") == nullptr)
        return launch_attn2_impl(p, stream);
    if (p.D == 64) return launch_attn_impl<64>(p, stream);
    if (p.D == 128) return launch_attn_impl<128>(p, stream);
    return cudaErrorInvalidValue;
}

}  // namespace ltxv
