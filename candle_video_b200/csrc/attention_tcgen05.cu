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
#define LTXV_PDL_CLASS 8
#include "launch.h"
#include "common.cuh"
#include "tensormap.h"
#include "options.h"
#include "profile.h"

#include <atomic>
#include <map>
#include <mutex>
#include <type_traits>
#include <utility>
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


// per-row factor of the deferred q RMS-norm (AttnParams::q_rscale); 1 when the queries arrive normalised
__device__ __forceinline__ float q_row_scale(const AttnParams& p, int batch, int qrow) {
    if (p.q_rscale == nullptr) return 1.0f;
    if (qrow >= p.Sq) qrow = p.Sq - 1;  // rows of a ragged last tile: any finite value
    return __ldg(p.q_rscale + static_cast<int64_t>(batch) * p.Sq + qrow);
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
    griddep_launch_dependents();
    griddep_wait();

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
        const float c = p.scale * kLog2e * q_row_scale(p, batch, q0 + row);
        const float inv_c = kLog2e / c;  // bias is added to the UNSCALED scores below, in units of 1 / (scale of this row)
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
                        if (bias != nullptr && col < p.Skv) x += __ldg(bias + col) * inv_c;
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
constexpr int kV2QBytes = kTileQ * 64 * 2;    // 16 KB per query tile
constexpr int kV2KVBytes = kTileKV * 64 * 2;  // 16 KB per K or V tile

// ================================================================================================
// v3 (head_dim 64, long key sequences): same CTA shape as v2 (256 queries = two 128-row tiles sharing every K/V tile,
// 384 threads, 1 CTA / SM) with the per-element instruction count and the synchronisation chain cut down:
//   * P never touches shared memory: each softmax thread packs its row of probabilities to bf16x2 and stores it with
//     tcgen05.st into TMEM (cols 384+64t), and O_t += P_t V_j runs as a TMEM-A ("TS") tcgen05.mma -- no STS, no swizzle
//     address math, no proxy fence, half the smem operand traffic of the PV MMA;
//   * packed f32x2 arithmetic (FFMA2 for s*c-m, FADD2 for the row sum) and 3-input max (FMNMX3): 3.1 issue slots per
//     score instead of 4.6, so a single warp per scheduler keeps the MUFU pipe (16 exp2/clk/SM) fed;
//   * setmaxnreg: the control warpgroup drops to 48 registers, the two softmax warpgroups rise to 224 -- the 128
//     register-resident scores + 32 in-flight exponentials no longer spill;
//   * one UMMA issuer warp PER query tile (warps 1, 2): tile 0's QK^T / PV never queue behind tile 1's barriers;
//   * 4-stage K/V rings (the 64 KB that held P).
// TMEM map (512 columns): S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384) P0 [384,448) P1 [448,512).
// ================================================================================================
#ifdef LTXV_ATTN_TRACE
// absolute clock64 timestamps of the hand-off events of one CTA: [role][kv tile][event]
//   roles 0,1: softmax warp 0 of tile 0/1; 2,3: UMMA issuer of tile 0/1; 4: TMA producer
__device__ long long g_attn_trace[11][40][8];
#define TRACE(role, j, ev)                                                                     \
    do {                                                                                       \
        if (tr_on && (j) < 40) g_attn_trace[role][j][ev] = clock64();                          \
    } while (0)
#else
#define TRACE(role, j, ev) do { } while (0)
#endif
constexpr int kV3Threads = 384;
constexpr int kV3Stages = 6;
constexpr int kV3L2Ahead = 4;  // K/V tiles prefetched into L2 beyond the smem ring
constexpr int kV3SmemBytes = 2 * kV2QBytes + 2 * kV3Stages * kV2KVBytes + 512;

__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A = 128 lanes x K bf16 packed two per 32-bit column (K-major)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Warp-level mbarrier hand-offs: ONE lane polls / arrives and the warp is released / gathered with __syncwarp.
// 128 threads polling the same mbarrier word serialise in the SYNCS unit (measured: ~200 clk for an already
// satisfied wait, >1000 clk wake-up for the UMMA issuer queued behind them).
__device__ __forceinline__ void warp_mbar_wait(uint64_t* bar, uint32_t parity, int lane) {
    if (lane == 0) mbar_wait_sleep(bar, parity);
    __syncwarp();
}
__device__ __forceinline__ void warp_mbar_arrive(uint64_t* bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

template <bool TAIL>
__device__ __forceinline__ float max32_v3(uint32_t (&r)[32], int kv_col0, int skv) {
    if (TAIL) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (kv_col0 + i >= skv) r[i] = 0xff800000u;  // -inf: ignored by the max, exp2 -> 0
    }
    float a = __uint_as_float(r[0]), b = __uint_as_float(r[1]);
#pragma unroll
    for (int i = 2; i < 32; i += 4) {
        a = max3(a, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
        if (i + 3 < 32) b = max3(b, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
    }
    return fmaxf(a, b);
}
// 2^x for a pair on the FMA/ALU pipes (no MUFU): x = n + r with n = rint(x) by the 1.5*2^23 magic-number add,
// r in [-0.5, 0.5]; 2^r by a degree-3 minimax polynomial (max rel. error 7.5e-5, far below the bf16 rounding of P);
// 2^n by adding n to the exponent field.  x is clamped at -126 so the exponent arithmetic cannot wrap.
__device__ __forceinline__ void ex2_poly_pair(float x0, float x1, float& e0, float& e1) {
    const uint64_t magic2 = pack_f32x2(12582912.0f, 12582912.0f);
    const uint64_t nmagic2 = pack_f32x2(-12582912.0f, -12582912.0f);
    const uint64_t m1 = pack_f32x2(-1.0f, -1.0f);
    const uint64_t x2 = pack_f32x2(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
    const uint64_t fi2 = add_f32x2(x2, magic2);
    const uint64_t nf2 = add_f32x2(fi2, nmagic2);
    const uint64_t r2 = fma_f32x2(nf2, m1, x2);
    uint64_t p2 = fma_f32x2(r2, pack_f32x2(0.0551716649f, 0.0551716649f), pack_f32x2(0.2426111219f, 0.2426111219f));
    p2 = fma_f32x2(p2, r2, pack_f32x2(0.6932609862f, 0.6932609862f));
    p2 = fma_f32x2(p2, r2, pack_f32x2(0.9999280736f, 0.9999280736f));
    float p0, p1, f0, f1;
    unpack_f32x2(p2, p0, p1);
    unpack_f32x2(fi2, f0, f1);
    e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(f0) << 23));
    e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(f1) << 23));
}
// Early probes: the softmax warps test s_full(j+1) / pv_done(j-1) with a non-blocking mbarrier.test_wait issued in the
// middle of the exponentials and only fall back to the blocking wait when the probe failed.  A blocking wait on an
// already completed phase still costs ~230 cycles per warp (TRYWAIT round trip + warp re-convergence; clock64 trace of
// tools/attn_prof.cu), twice per key tile; P is stored in one piece at the end of the tile so that the pv_done check
// moves from the middle of the exponentials (where P V(j-1) is often still in flight) to their end.  +2.5 % at c2,
// +5 % at c3 (800 -> 820, 869 -> 911 TFLOP/s isolated).
#ifndef LTXV_ATTN_POLY
#define LTXV_ATTN_POLY 2  // pairs out of every 8 (16 scores) whose exp2 runs on the FMA pipe instead of MUFU: 0..8
#endif
// 32 scores -> exp2(s*c - m) -> row-sum (packed) -> 16 bf16x2 words
__device__ __forceinline__ void exp_chunk_v3(const uint32_t (&r)[32], uint64_t c2, uint64_t negm2, uint64_t& l2a,
                                             uint64_t& l2b, uint32_t* out16) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        float x0, x1, x2, x3;
        unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), c2, negm2), x0, x1);
        unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), c2, negm2), x2, x3);
        float e0, e1, e2, e3;
        // the MUFU pipe (16 exp2/clk/SM) is the attention bottleneck at head_dim 64: a fixed share of the
        // exponentials is computed on the otherwise idle FMA pipe
        // pair index within a group of 8 pairs: a = (i/2)&7, b = a+1; emulated pairs are spread evenly: 7,3,5,1,6,2,4,0
        constexpr int kOrder[8] = {7, 3, 5, 1, 6, 2, 4, 0};
        bool poly_a = false, poly_b = false;
#pragma unroll
        for (int q = 0; q < LTXV_ATTN_POLY; ++q) {
            poly_a = poly_a || (kOrder[q] == ((i >> 1) & 7));
            poly_b = poly_b || (kOrder[q] == (((i >> 1) + 1) & 7));
        }
        if (poly_a) {
            ex2_poly_pair(x0, x1, e0, e1);
        } else {
            e0 = ex2_approx(x0);
            e1 = ex2_approx(x1);
        }
        if (poly_b) {
            ex2_poly_pair(x2, x3, e2, e3);
        } else {
            e2 = ex2_approx(x2);
            e3 = ex2_approx(x3);
        }
        l2a = add_f32x2(l2a, pack_f32x2(e0, e1));
        l2b = add_f32x2(l2b, pack_f32x2(e2, e3));
        out16[i / 2] = pack_bf16x2(e0, e1);
        out16[i / 2 + 1] = pack_bf16x2(e2, e3);
    }
}

// Tail splitting.  The grid is one CTA per (batch, head, 256-query block) unit; with U units on 148 SMs the last wave
// holds only U mod 148 CTAs (c2: 640 units = 4.32 waves -> the 5th wave runs 48 CTAs on 148 SMs).  Those last units are
// cut along the key axis into nsplit = 148 / (U mod 148) ranges each, so the tail wave fills the machine and lasts
// 1/nsplit as long; each range CTA publishes its unnormalised (O, m, l) rows to a scratch buffer and the last CTA of a
// unit to finish (ticket counter) merges them:  O = sum_i 2^(m_i - M) O_i,  l = sum_i 2^(m_i - M) l_i,  M = max_i m_i.
struct SplitPlan {
    int n_qb;           // 256-query blocks per (batch, head)
    int n_units;        // B * H * n_qb
    int n_split_units;  // trailing units that are split (0 = none)
    int nsplit;         // key ranges per split unit (>= 1)
    float* scratch;     // [n_split_units][nsplit][kSplitCols4 float4 columns][256 rows]
    int* counters;      // [n_split_units] tickets, zero between launches
};
constexpr int kSplitCols4 = 17;  // float4 columns per row: 16 x 4 O columns + (m, l, -, -)
constexpr int kMaxSplit = 8;

__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* dst, const uint32_t (&r)[32], float inv_l) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
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

__global__ void __launch_bounds__(kV3Threads, 1)
flash_attn3_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ AttnParams p,
                   const SplitPlan sp) {
    constexpr int D = 64;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sq = smem;
    uint8_t* sk = sq + 2 * kV2QBytes;
    uint8_t* sv = sk + kV3Stages * kV2KVBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sv + kV3Stages * kV2KVBytes);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;                   // [St]
    uint64_t* k_empty = bars + 1 + kV3Stages;      // [St]  one arrival per active query tile
    uint64_t* v_full = bars + 1 + 2 * kV3Stages;   // [St]
    uint64_t* v_empty = bars + 1 + 3 * kV3Stages;  // [St]  one arrival per active query tile
    uint64_t* s_full = bars + 1 + 4 * kV3Stages;   // [2]  S_t landed in TMEM
    uint64_t* s_free = s_full + 2;                 // [2]  S_t copied to registers (128 arrivals)
    uint64_t* p_full = s_free + 2;                 // [2]  P_t stored to TMEM (128 arrivals)
    uint64_t* pv_done = p_full + 2;                // [2]  O_t += P_t V retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
    volatile uint32_t* split_flag = tmem_slot + 1;  // 1 = this CTA drew the last ticket of its split unit

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
#ifdef LTXV_ATTN_TRACE
    const long long t_entry = clock64();
    const long long t_entry_ns = static_cast<long long>(globaltimer_ns());
    if (threadIdx.x == 0 && blockIdx.x == 0) g_attn_trace[9][19][4] = t_entry_ns;
    if (threadIdx.x == 0 && blockIdx.x >= gridDim.x - 8) g_attn_trace[9][10 + (gridDim.x - 1 - blockIdx.x)][0] = t_entry_ns;
#endif
    // work item -> (unit = (batch, head, 256-query block), key-tile range).  The last `sp.n_split_units` units (the
    // partial last wave of CTAs) are cut into sp.nsplit key ranges each, see SplitPlan.
    const int item = blockIdx.x;
    const int first_split = sp.n_units - sp.n_split_units;
    int unit = item, split = 0;
    const bool is_split = item >= first_split;
    if (is_split) {
        unit = first_split + (item - first_split) / sp.nsplit;
        split = (item - first_split) % sp.nsplit;
    }
    const int qb = unit % sp.n_qb;
    const int head = (unit / sp.n_qb) % p.H;
    const int batch = unit / (sp.n_qb * p.H);
    const int q0 = qb * (2 * kTileQ);
    const int n_tiles_all = (p.Skv + kTileKV - 1) / kTileKV;
    const int jt0 = is_split ? (split * n_tiles_all) / sp.nsplit : 0;
    const int jt1 = is_split ? ((split + 1) * n_tiles_all) / sp.nsplit : n_tiles_all;
    const int n_tiles = jt1 - jt0;  // key tiles of THIS work item; j below is the local tile index
    const bool two = (q0 + kTileQ) < p.Sq;  // second query tile has rows
    const int nt = two ? 2 : 1;

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) {
            printf("ltxv attention v3: dynamic smem base not 1024B aligned\n");
            __trap();
        }
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
        mbar_init(q_full, 1);
        for (int i = 0; i < kV3Stages; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], nt);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], nt);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            mbar_init(&s_free[t], 4);  // one arrival per softmax warp
            mbar_init(&p_full[t], 4);
            mbar_init(&pv_done[t], 1);
        }
        fence_barrier_init();
    }
    if (warp_idx == 11) tmem_alloc<512>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_launch_dependents();
    griddep_wait();
#ifdef LTXV_ATTN_TRACE
    if (threadIdx.x == 0 && qb == 3 && head == 5 && split == 0) {
        g_attn_trace[10][39][0] = t_entry;
        g_attn_trace[10][39][1] = clock64();
    }
#endif

    // warps 0-7: softmax (tile 0: 0-3, tile 1: 4-7); warps 8-11: control (TMA, UMMA issuer of tile 0, of tile 1, TMEM
    // owner).  The control warps sit at the HIGH warp ids on purpose: their few instructions win the issue
    // arbitration against the always-ready softmax warps that share their scheduler.
    if (warp_idx >= 8) {
        setmaxnreg_dec<48>();
        if (warp_idx == 8) {
            // ===================== TMA producer =====================
            if (elect_one()) {
                mbar_arrive_expect_tx(q_full, nt * kV2QBytes);
                tma_load_3d(sq, &tm_q, q_full, p.q_col0 + head * D, q0, batch);
                if (two) tma_load_3d(sq + kV2QBytes, &tm_q, q_full, p.q_col0 + head * D, q0 + kTileQ, batch);
                int stage = 0;
                uint32_t phase = 0;
#ifdef LTXV_ATTN_TRACE
                const bool tr_on = (qb == 3 && head == 5 && split == 0);
#endif
                for (int j = 0; j < n_tiles; ++j) {
                    const int kv0 = (jt0 + j) * kTileKV;
                    if (j + kV3Stages + kV3L2Ahead - 1 < n_tiles) {
                        // the ring only looks kV3Stages tiles ahead and its slots free up late (after the PV MMAs
                        // retire); an L2 prefetch further out turns the ring's loads into L2 hits
                        const int kvp = kv0 + (kV3Stages + kV3L2Ahead - 1) * kTileKV;
                        tma_prefetch_3d(&tm_k, p.k_col0 + head * D, kvp, batch);
                        tma_prefetch_3d(&tm_v, p.v_col0 + head * D, kvp, batch);
                    }
                    mbar_wait_sleep(&k_empty[stage], phase ^ 1);
                    TRACE(10, j, 0);
                    mbar_arrive_expect_tx(&k_full[stage], kV2KVBytes);
                    tma_load_3d(sk + stage * kV2KVBytes, &tm_k, &k_full[stage], p.k_col0 + head * D, kv0, batch);
                    mbar_wait_sleep(&v_empty[stage], phase ^ 1);
                    TRACE(10, j, 1);
                    mbar_arrive_expect_tx(&v_full[stage], kV2KVBytes);
                    tma_load_3d(sv + stage * kV2KVBytes, &tm_v, &v_full[stage], p.v_col0 + head * D, kv0, batch);
                    if (++stage == kV3Stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        } else if (warp_idx == 9 || (warp_idx == 10 && two)) {
            // ===================== UMMA issuer of query tile t =====================
            const int t = warp_idx - 9;
            if (elect_one()) {
                constexpr uint32_t idesc_s = make_idesc_bf16(kTileQ, kTileKV, false, false);
                constexpr uint32_t idesc_pv = make_idesc_bf16(kTileQ, D, false, true);  // B = V is MN-major
                const uint32_t tmem_s = tmem_base + t * 128;
                const uint32_t tmem_o = tmem_base + 256 + t * 64;
                const uint32_t tmem_p = tmem_base + 384 + t * 64;
                // descriptors differ from tile to tile only in the 14-bit start-address field: build the constant part
                // once and add (address >> 4) per instruction (all operand addresses stay below 256 KB: no carry)
                const uint64_t desc_k0 = make_smem_desc_sw128(0, 1024, 0);
                const uint64_t desc_v0 = make_smem_desc_sw128(0, 1024, kTileKV * 128);
                const uint64_t desc_q = make_smem_desc_sw128(smem_u32(sq) + t * kV2QBytes, 1024, 0);
                const uint32_t sk_a = smem_u32(sk) >> 4, sv_a = smem_u32(sv) >> 4;
                auto issue_s = [&](int stage) {
                    const uint64_t dk = desc_k0 + (sk_a + stage * (kV2KVBytes >> 4));
#pragma unroll
                    for (int ks = 0; ks < D / 16; ++ks) {
                        umma_bf16_ss(tmem_s, desc_q + ks * 2, dk + ks * 2, idesc_s, ks != 0 ? 1u : 0u);
                    }
                };
                mbar_wait_sleep(q_full, 0);
                mbar_wait_sleep(&k_full[0], 0);
                tcgen05_fence_after();
                issue_s(0);
                umma_commit(&s_full[t]);
                umma_commit(&k_empty[0]);
#ifdef LTXV_ATTN_TIMING
                const bool tm_on = (qb == 3 && head == 5 && split == 0);
                long long tm_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                long long tm_last = clock64();
#endif
#ifdef LTXV_ATTN_TRACE
                const bool tr_on = (qb == 3 && head == 5 && split == 0);
#endif
                for (int j = 0; j < n_tiles; ++j) {
                    const int stage = j % kV3Stages, nstage = (j + 1) % kV3Stages;
                    const uint32_t phase = (j / kV3Stages) & 1, nphase = ((j + 1) / kV3Stages) & 1;
                    if (j + 1 < n_tiles) {
#ifdef LTXV_ATTN_TIMING
                        TMARK(7);
                        mbar_wait_sleep(q_full, 0);  // completed long ago: measures the fixed cost of a satisfied wait
                        TMARK(6);
#endif
                        // S_t(j+1) as soon as the softmax group holds S_t(j) in registers
                        mbar_wait_sleep(&k_full[nstage], nphase);
                        TMARK(0);
                        TRACE(8 + t, j, 0);
                        mbar_wait_sleep(&s_free[t], j & 1);
                        TMARK(1);
                        TRACE(8 + t, j, 1);
                        tcgen05_fence_after();
                        issue_s(nstage);
                        umma_commit(&s_full[t]);
                        umma_commit(&k_empty[nstage]);
                        TMARK(2);
                        TRACE(8 + t, j, 2);
                    }
                    mbar_wait_sleep(&v_full[stage], phase);
                    TMARK(3);
                    TRACE(8 + t, j, 3);
                    mbar_wait_sleep(&p_full[t], j & 1);
                    TMARK(4);
                    TRACE(8 + t, j, 4);
                    tcgen05_fence_after();
                    const uint64_t dv = desc_v0 + (sv_a + stage * (kV2KVBytes >> 4));
#pragma unroll
                    for (int ks = 0; ks < kTileKV / 16; ++ks) {
                        umma_bf16_ts(tmem_o, tmem_p + ks * 8, dv + ks * ((16 * 128) >> 4), idesc_pv, (j | ks) != 0 ? 1u : 0u);
                    }
                    umma_commit(&pv_done[t]);
                    umma_commit(&v_empty[stage]);
                    TMARK(5);
                    TRACE(8 + t, j, 5);
                }
#ifdef LTXV_ATTN_TIMING
                if (tm_on)
                    for (int i = 0; i < 8; ++i) g_attn_timing[16 + 8 * t + i] = tm_acc[i];
#endif
            }
        }
    } else {
        setmaxnreg_inc<224>();
        const int t = warp_idx >> 2;  // query tile of this softmax group
        if (t == 0 || two) {
            const int quad = warp_idx & 3;
            const int row = quad * 32 + lane;
            const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
            const uint32_t tmem_s = lane_base + t * 128;
            const uint32_t tmem_o = lane_base + 256 + t * 64;
            const uint32_t tmem_p = lane_base + 384 + t * 64;
            const float c = p.scale * kLog2e * q_row_scale(p, batch, q0 + t * kTileQ + row);
            const uint64_t c2 = pack_f32x2(c, c);
            float m_used = -INFINITY;
            uint64_t l2a = pack_f32x2(0.f, 0.f), l2b = l2a;
            uint32_t sfull_probe = 0;  // lane 0: result of the early test of s_full(j) issued during tile j-1
#ifdef LTXV_ATTN_TIMING
            const bool tm_on = (qb == 3 && head == 5 && split == 0 && lane == 0 && (warp_idx == 0 || warp_idx == 4));
            long long tm_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            long long tm_last = clock64();
#endif
#ifdef LTXV_ATTN_TRACE
            const bool tr_on = (qb == 3 && head == 5 && split == 0 && lane == 0);
#endif
#ifdef LTXV_ATTN_STAGGER
            // timing experiment: start the second query tile's softmax chain this many cycles late so that
            // the two softmax warps of a sub-partition are in opposite phases (exp vs. everything else)
            if (t == 1) {
                const long long stagger_t0 = clock64();
                while (clock64() - stagger_t0 < LTXV_ATTN_STAGGER) {
                }
            }
#endif
            auto tile = [&](int j, auto tail_tag) {
                constexpr bool TAIL = decltype(tail_tag)::value;
                const int kv0 = (jt0 + j) * kTileKV;
                uint32_t s0[32], s1[32], s2[32], s3[32];
                TRACE(warp_idx, j, 0);
                if (!__shfl_sync(0xffffffffu, sfull_probe, 0)) warp_mbar_wait(&s_full[t], j & 1, lane);
                sfull_probe = 0;
                TRACE(warp_idx, j, 1);
                TMARK(0);
                tcgen05_fence_after();
#ifdef LTXV_ATTN_EXPERIMENT_MMA_ONLY
                // timing experiment: no softmax work at all, only the barrier handshakes -> tensor-side floor
                tcgen05_fence_before();
                warp_mbar_arrive(&s_free[t], lane);
                if (j > 0) warp_mbar_wait(&pv_done[t], (j - 1) & 1, lane);
                tcgen05_fence_before();
                warp_mbar_arrive(&p_full[t], lane);
                return;
#endif
                tmem_ld_32x32b_x32(tmem_s + 0, s0);
                tmem_ld_32x32b_x32(tmem_s + 32, s1);
                tmem_ld_32x32b_x32(tmem_s + 64, s2);
                tmem_ld_32x32b_x32(tmem_s + 96, s3);
                tmem_ld_wait();
                tcgen05_fence_before();
                warp_mbar_arrive(&s_free[t], lane);  // S_t may be overwritten by the next QK^T
                TRACE(warp_idx, j, 2);
                TMARK(1);
#ifdef LTXV_ATTN_EXPERIMENT_NO_MAX
                // timing experiment only (lead a): what the key loop costs without the per-tile row maximum
                // -- the running maximum of the first tile is kept for all later tiles (wrong for peaky inputs)
                float mx = j == 0 ? fmaxf(fmaxf(max32_v3<TAIL>(s0, kv0, p.Skv), max32_v3<TAIL>(s1, kv0 + 32, p.Skv)),
                                          fmaxf(max32_v3<TAIL>(s2, kv0 + 64, p.Skv), max32_v3<TAIL>(s3, kv0 + 96, p.Skv)))
                                  : m_used / c;
#else
                float mx = fmaxf(fmaxf(max32_v3<TAIL>(s0, kv0, p.Skv), max32_v3<TAIL>(s1, kv0 + 32, p.Skv)),
                                 fmaxf(max32_v3<TAIL>(s2, kv0 + 64, p.Skv), max32_v3<TAIL>(s3, kv0 + 96, p.Skv)));
#endif
                mx *= c;
                TMARK(2);
                bool pv_waited = (j == 0);
                if (j == 0) {
                    m_used = mx;
                } else {
                    const float m_new = fmaxf(m_used, mx);
                    if (__any_sync(0xffffffffu, (m_new - m_used) > kRescaleThreshold)) {
                        // rare (first few key tiles): O_t is rescaled in TMEM, which needs the previous P_t V retired
                        warp_mbar_wait(&pv_done[t], (j - 1) & 1, lane);
                        pv_waited = true;
                        tcgen05_fence_after();
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
                        const uint64_t a2 = pack_f32x2(alpha, alpha), z2 = pack_f32x2(0.f, 0.f);
                        l2a = fma_f32x2(l2a, a2, z2);
                        l2b = fma_f32x2(l2b, a2, z2);
                        m_used = m_new;
                    }
                }
                TMARK(3);
                const uint64_t negm2 = pack_f32x2(-m_used, -m_used);
                uint32_t pk[32];
                exp_chunk_v3(s0, c2, negm2, l2a, l2b, pk);
                exp_chunk_v3(s1, c2, negm2, l2a, l2b, pk + 16);
                TMARK(4);
                TRACE(warp_idx, j, 3);
                uint32_t pk2[32];
                uint32_t pv_probe = 0;
                exp_chunk_v3(s2, c2, negm2, l2a, l2b, pk2);
                if (!pv_waited && lane == 0) pv_probe = mbar_test_wait(&pv_done[t], (j - 1) & 1);
                if (j + 1 < n_tiles && lane == 0) sfull_probe = mbar_test_wait(&s_full[t], (j + 1) & 1);
                exp_chunk_v3(s3, c2, negm2, l2a, l2b, pk2 + 16);
                TMARK(5);
                if (!pv_waited) {
                    // P_t is single-buffered: P_t V(j-1) must have read it before it is overwritten
                    if (!__shfl_sync(0xffffffffu, pv_probe, 0)) warp_mbar_wait(&pv_done[t], (j - 1) & 1, lane);
                    tcgen05_fence_after();
                }
                TRACE(warp_idx, j, 4);
                tmem_st_32x32b_x32(tmem_p, pk);          // keys  0..63  -> P columns  0..31
                tmem_st_32x32b_x32(tmem_p + 32, pk2);    // keys 64..127 -> P columns 32..63
                TMARK(6);
                tmem_st_wait();
                tcgen05_fence_before();
                warp_mbar_arrive(&p_full[t], lane);
                TRACE(warp_idx, j, 5);
                TMARK(7);
            };
#ifdef LTXV_ATTN_TRACE
            if (tr_on && warp_idx == 0) g_attn_trace[10][39][2] = clock64();
#endif
            const bool ragged = (jt1 == n_tiles_all) && (p.Skv % kTileKV != 0);  // last key tile has a tail
            const int n_full = n_tiles - (ragged ? 1 : 0);
            for (int j = 0; j < n_full; ++j) tile(j, std::false_type{});
            if (ragged) tile(n_full, std::true_type{});
#ifdef LTXV_ATTN_TIMING
            if (tm_on)
                for (int i = 0; i < 8; ++i) g_attn_timing[(warp_idx == 0 ? 0 : 8) + i] = tm_acc[i];
#endif
            warp_mbar_wait(&pv_done[t], (n_tiles - 1) & 1, lane);
            tcgen05_fence_after();
#ifdef LTXV_ATTN_TRACE
            if (tr_on && warp_idx == 0) g_attn_trace[10][39][3] = clock64();
#endif
            float la, lb, lc, ld;
            unpack_f32x2(l2a, la, lb);
            unpack_f32x2(l2b, lc, ld);
            float l_row = (la + lb) + (lc + ld);
            const int qrow = q0 + t * kTileQ + row;
            __nv_bfloat16* orow = attn_out_row(p, batch, qrow < p.Sq ? qrow : 0, head, D);
            if (!is_split) {
                const float inv_l = 1.0f / l_row;
#pragma unroll 1
                for (int dc = 0; dc < D / 32; ++dc) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(tmem_o + dc * 32, r);
                    tmem_ld_wait();
                    if (qrow < p.Sq) store_row_bf16(orow + dc * 32, r, inv_l);
                }
            } else {
#ifdef LTXV_ATTN_TRACE
                const bool ts_on = (unit == first_split && threadIdx.x == 0);
                if (ts_on) {
                    g_attn_trace[9][20 + split][0] = t_entry;
                    g_attn_trace[9][20 + split][1] = clock64();
                }
#endif
                // ---- split unit: publish (O unnormalised, m, l) of this key range; the LAST of the nsplit CTAs of the
                // unit (ticket counter) merges all partials and writes the output rows ----
                const int su = unit - first_split;
                const int r256 = t * kTileQ + row;
                // scratch is column-major per work item ([17 float4 columns][256 rows]): the 32 rows of a warp are
                // contiguous, so every store / load below is a fully coalesced 512-byte access
                float4* mine = reinterpret_cast<float4*>(sp.scratch) +
                               static_cast<int64_t>(su * sp.nsplit + split) * (kSplitCols4 * 256) + r256;
#pragma unroll 1
                for (int dc = 0; dc < D / 32; ++dc) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(tmem_o + dc * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        mine[(dc * 8 + q) * 256] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                                               __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                }
                mine[16 * 256] = make_float4(m_used, l_row, 0.f, 0.f);
                __threadfence();
                named_bar_sync(1, 128 * nt);  // all softmax threads of the CTA have published
#ifdef LTXV_ATTN_TRACE
                if (ts_on) g_attn_trace[9][20 + split][2] = clock64();
#endif
                if (threadIdx.x == 0) {
                    const int ticket = atomicAdd(sp.counters + su, 1);
                    *split_flag = (ticket == sp.nsplit - 1) ? 1 : 0;
                    if (ticket == sp.nsplit - 1) sp.counters[su] = 0;  // re-arm for the next launch on this stream
                    __threadfence();
                }
                named_bar_sync(1, 128 * nt);
#ifdef LTXV_ATTN_TRACE
                if (ts_on) g_attn_trace[9][20 + split][3] = clock64();
#endif
                if (*split_flag != 0) {
                    const float4* base = reinterpret_cast<const float4*>(sp.scratch) +
                                         static_cast<int64_t>(su * sp.nsplit) * (kSplitCols4 * 256) + r256;
                    constexpr int64_t stride = kSplitCols4 * 256;  // float4s per work item
                    // all (m, l) pairs in flight at once (one L2 round trip), then the weights 2^(m_i - M)
                    float4 ml[kMaxSplit];
#pragma unroll
                    for (int i = 0; i < kMaxSplit; ++i)
                        ml[i] = i < sp.nsplit ? __ldcg(base + i * stride + 16 * 256) : make_float4(-INFINITY, 0.f, 0.f, 0.f);
                    float m_all = -INFINITY;
#pragma unroll
                    for (int i = 0; i < kMaxSplit; ++i) m_all = fmaxf(m_all, ml[i].x);
                    float wgt[kMaxSplit];
                    float l_all = 0.f;
#pragma unroll
                    for (int i = 0; i < kMaxSplit; ++i) {
                        wgt[i] = i < sp.nsplit ? ex2_approx(ml[i].x - m_all) : 0.f;
                        l_all = fmaf(ml[i].y, wgt[i], l_all);
                    }
                    const float inv_l = 1.0f / l_all;
#pragma unroll 1
                    for (int dc = 0; dc < D / 16; ++dc) {  // 16 output columns at a time: 4 x float4 per partial in flight
                        float acc[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q) acc[q] = 0.f;
#pragma unroll
                        for (int i = 0; i < kMaxSplit; ++i) {
                            if (i < sp.nsplit) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float4 v = __ldcg(base + i * stride + (dc * 4 + q) * 256);
                                    acc[4 * q + 0] = fmaf(v.x, wgt[i], acc[4 * q + 0]);
                                    acc[4 * q + 1] = fmaf(v.y, wgt[i], acc[4 * q + 1]);
                                    acc[4 * q + 2] = fmaf(v.z, wgt[i], acc[4 * q + 2]);
                                    acc[4 * q + 3] = fmaf(v.w, wgt[i], acc[4 * q + 3]);
                                }
                            }
                        }
                        if (qrow < p.Sq) {
                            uint4* d4 = reinterpret_cast<uint4*>(orow + dc * 16);
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                uint4 u;
                                u.x = pack_bf16x2(acc[8 * q + 0] * inv_l, acc[8 * q + 1] * inv_l);
                                u.y = pack_bf16x2(acc[8 * q + 2] * inv_l, acc[8 * q + 3] * inv_l);
                                u.z = pack_bf16x2(acc[8 * q + 4] * inv_l, acc[8 * q + 5] * inv_l);
                                u.w = pack_bf16x2(acc[8 * q + 6] * inv_l, acc[8 * q + 7] * inv_l);
                                d4[q] = u;
                            }
                        }
                    }
                }
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
#ifdef LTXV_ATTN_TRACE
    if (threadIdx.x == 0 && qb == 3 && head == 5 && split == 0) g_attn_trace[10][39][4] = clock64();
    if (threadIdx.x == 0 && is_split && unit == first_split) {
        g_attn_trace[9][20 + split][4] = clock64();
        g_attn_trace[9][20 + split][5] = static_cast<long long>(globaltimer_ns());
        g_attn_trace[9][20 + split][6] = *split_flag;
    }
    if (threadIdx.x == 0 && item == 0) g_attn_trace[9][19][5] = static_cast<long long>(globaltimer_ns());
#endif
    if (warp_idx == 11) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}


// ================================================================================================
// v4 (head_dim 64, long key sequences): the v3 organisation with HALF a score row per softmax thread.  OPT-IN
// (LTXV_ATTN_V4 / option attn_v4): correct (tools/attn_test passes on it) but measured SLOWER than v3 -- 744 vs 820
// TFLOP/s at the batched c2 launch, 828 vs 911 at c3 -- so it documents a negative result rather than a fast path: four
// softmax warps per sub-partition do not raise the issue rate; the extra per-warp barrier / exchange instructions
// cost more than the added interleaving buys (fewer polynomial exponentials help it, more hurt it: issue-bound).
//
// clock64 traces of v3 (tools/attn_prof.cu) show that a softmax warp needs ~2200 cycles for the ~700 instructions of a
// key tile even when it never waits: with only two softmax warps per SM sub-partition the dependent MUFU / FMA / F2FP
// chains leave the issue port idle half of the time (ncu: issue 51 %, XU 56 %, tensor 37 %).  Here every 128-row query
// tile is served by EIGHT warps instead of four -- warps q and q + 4 of a tile share TMEM lane quadrant q and take key
// columns [0, 64) and [64, 128) of the same rows -- so each sub-partition interleaves four softmax warps.  The two halves
// of a row exchange their partial row maximum through shared memory (one 64-thread named barrier per key tile) and keep
// private partial row sums that are combined once at the end; everything else (P in TMEM, TS-form P V, lazy O rescale,
// tail splitting) is as in v3.
//   warps  0..15  softmax: tile t = w >> 3, half h = (w >> 2) & 1, quadrant = w & 3   (112 registers)
//   warp   16     TMA producer;  17, 18  UMMA issuers of tile 0 / 1;  19  TMEM owner   (32 registers)
// ================================================================================================
constexpr int kV4Threads = 640;
constexpr int kV4Stages = 5;  // 227 KB of shared memory: 5 K/V stages + the 6 KB of row-max / row-sum exchange buffers
constexpr int kV4RingBytes = 2 * kV2QBytes + 2 * kV4Stages * kV2KVBytes + 512;
constexpr int kV4SmemBytes = kV4RingBytes + 2 * 2 * 2 * 128 * 4 + 2 * 2 * 128 * 4;  // + xmax[2][2][2][128] + xsum[2][2][128]

// 32 scores -> exp2(s*c - m) -> row-sum (packed) -> 16 bf16x2 words; a polynomial share of LTXV_ATTN_POLY4 pairs per 8
#ifndef LTXV_ATTN_POLY4
#define LTXV_ATTN_POLY4 2
#endif
__device__ __forceinline__ void exp_chunk_v4(const uint32_t (&r)[32], uint64_t c2, uint64_t negm2, uint64_t& l2a,
                                             uint64_t& l2b, uint32_t* out16) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        float x0, x1, x2, x3;
        unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), c2, negm2), x0, x1);
        unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), c2, negm2), x2, x3);
        float e0, e1, e2, e3;
        constexpr int kOrder[8] = {7, 3, 5, 1, 6, 2, 4, 0};
        bool poly_a = false, poly_b = false;
#pragma unroll
        for (int q = 0; q < LTXV_ATTN_POLY4; ++q) {
            poly_a = poly_a || (kOrder[q] == ((i >> 1) & 7));
            poly_b = poly_b || (kOrder[q] == (((i >> 1) + 1) & 7));
        }
        if (poly_a) {
            ex2_poly_pair(x0, x1, e0, e1);
        } else {
            e0 = ex2_approx(x0);
            e1 = ex2_approx(x1);
        }
        if (poly_b) {
            ex2_poly_pair(x2, x3, e2, e3);
        } else {
            e2 = ex2_approx(x2);
            e3 = ex2_approx(x3);
        }
        l2a = add_f32x2(l2a, pack_f32x2(e0, e1));
        l2b = add_f32x2(l2b, pack_f32x2(e2, e3));
        out16[i / 2] = pack_bf16x2(e0, e1);
        out16[i / 2 + 1] = pack_bf16x2(e2, e3);
    }
}

__global__ void __launch_bounds__(kV4Threads, 1)
flash_attn4_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ AttnParams p,
                   const SplitPlan sp) {
    constexpr int D = 64;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sq = smem;
    uint8_t* sk = sq + 2 * kV2QBytes;
    uint8_t* sv = sk + kV4Stages * kV2KVBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sv + kV4Stages * kV2KVBytes);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;                   // [St]
    uint64_t* k_empty = bars + 1 + kV4Stages;      // [St]  one arrival per active query tile
    uint64_t* v_full = bars + 1 + 2 * kV4Stages;   // [St]
    uint64_t* v_empty = bars + 1 + 3 * kV4Stages;  // [St]  one arrival per active query tile
    uint64_t* s_full = bars + 1 + 4 * kV4Stages;   // [2]  S_t landed in TMEM
    uint64_t* s_free = s_full + 2;                 // [2]  S_t copied to registers (8 warp arrivals)
    uint64_t* p_full = s_free + 2;                 // [2]  P_t stored to TMEM (8 warp arrivals)
    uint64_t* pv_done = p_full + 2;                // [2]  O_t += P_t V retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
    volatile uint32_t* split_flag = tmem_slot + 1;  // 1 = this CTA drew the last ticket of its split unit
    float* xmax = reinterpret_cast<float*>(smem + kV4RingBytes);  // [slot 2][tile 2][half 2][row 128]
    float* xsum = xmax + 2 * 2 * 2 * 128;                         // [tile 2][half 2][row 128]

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x;
    const int first_split = sp.n_units - sp.n_split_units;
    int unit = item, split = 0;
    const bool is_split = item >= first_split;
    if (is_split) {
        unit = first_split + (item - first_split) / sp.nsplit;
        split = (item - first_split) % sp.nsplit;
    }
    const int qb = unit % sp.n_qb;
    const int head = (unit / sp.n_qb) % p.H;
    const int batch = unit / (sp.n_qb * p.H);
    const int q0 = qb * (2 * kTileQ);
    const int n_tiles_all = (p.Skv + kTileKV - 1) / kTileKV;
    const int jt0 = is_split ? (split * n_tiles_all) / sp.nsplit : 0;
    const int jt1 = is_split ? ((split + 1) * n_tiles_all) / sp.nsplit : n_tiles_all;
    const int n_tiles = jt1 - jt0;
    const bool two = (q0 + kTileQ) < p.Sq;
    const int nt = two ? 2 : 1;

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) {
            printf("ltxv attention v4: dynamic smem base not 1024B aligned\n");
            __trap();
        }
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
        mbar_init(q_full, 1);
        for (int i = 0; i < kV4Stages; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], nt);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], nt);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            mbar_init(&s_free[t], 8);  // one arrival per softmax warp of the tile
            mbar_init(&p_full[t], 8);
            mbar_init(&pv_done[t], 1);
        }
        fence_barrier_init();
    }
    if (warp_idx == 19) tmem_alloc<512>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_launch_dependents();
    griddep_wait();

    if (warp_idx >= 16) {
        // setmaxnreg.inc can only take registers that other warps of the CTA released: the launch allocates 640 x 96, the
        // control warpgroup gives back 128 x (96 - 32) = 8192 = the 512 x (112 - 96) the softmax warps ask for
        setmaxnreg_dec<32>();
        if (warp_idx == 16) {
            // ===================== TMA producer =====================
            if (elect_one()) {
                mbar_arrive_expect_tx(q_full, nt * kV2QBytes);
                tma_load_3d(sq, &tm_q, q_full, p.q_col0 + head * D, q0, batch);
                if (two) tma_load_3d(sq + kV2QBytes, &tm_q, q_full, p.q_col0 + head * D, q0 + kTileQ, batch);
                int stage = 0;
                uint32_t phase = 0;
                for (int j = 0; j < n_tiles; ++j) {
                    const int kv0 = (jt0 + j) * kTileKV;
                    if (j + kV4Stages + kV3L2Ahead - 1 < n_tiles) {
                        const int kvp = kv0 + (kV4Stages + kV3L2Ahead - 1) * kTileKV;
                        tma_prefetch_3d(&tm_k, p.k_col0 + head * D, kvp, batch);
                        tma_prefetch_3d(&tm_v, p.v_col0 + head * D, kvp, batch);
                    }
                    mbar_wait_sleep(&k_empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&k_full[stage], kV2KVBytes);
                    tma_load_3d(sk + stage * kV2KVBytes, &tm_k, &k_full[stage], p.k_col0 + head * D, kv0, batch);
                    mbar_wait_sleep(&v_empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&v_full[stage], kV2KVBytes);
                    tma_load_3d(sv + stage * kV2KVBytes, &tm_v, &v_full[stage], p.v_col0 + head * D, kv0, batch);
                    if (++stage == kV4Stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        } else if (warp_idx == 17 || (warp_idx == 18 && two)) {
            // ===================== UMMA issuer of query tile t =====================
            const int t = warp_idx - 17;
            if (elect_one()) {
                constexpr uint32_t idesc_s = make_idesc_bf16(kTileQ, kTileKV, false, false);
                constexpr uint32_t idesc_pv = make_idesc_bf16(kTileQ, D, false, true);  // B = V is MN-major
                const uint32_t tmem_s = tmem_base + t * 128;
                const uint32_t tmem_o = tmem_base + 256 + t * 64;
                const uint32_t tmem_p = tmem_base + 384 + t * 64;
                const uint64_t desc_k0 = make_smem_desc_sw128(0, 1024, 0);
                const uint64_t desc_v0 = make_smem_desc_sw128(0, 1024, kTileKV * 128);
                const uint64_t desc_q = make_smem_desc_sw128(smem_u32(sq) + t * kV2QBytes, 1024, 0);
                const uint32_t sk_a = smem_u32(sk) >> 4, sv_a = smem_u32(sv) >> 4;
                auto issue_s = [&](int stage) {
                    const uint64_t dk = desc_k0 + (sk_a + stage * (kV2KVBytes >> 4));
#pragma unroll
                    for (int ks = 0; ks < D / 16; ++ks)
                        umma_bf16_ss(tmem_s, desc_q + ks * 2, dk + ks * 2, idesc_s, ks != 0 ? 1u : 0u);
                };
                mbar_wait_sleep(q_full, 0);
                mbar_wait_sleep(&k_full[0], 0);
                tcgen05_fence_after();
                issue_s(0);
                umma_commit(&s_full[t]);
                umma_commit(&k_empty[0]);
                int stage = 0;
                uint32_t phase = 0;
                for (int j = 0; j < n_tiles; ++j) {
                    int nstage = stage + 1;
                    uint32_t nphase = phase;
                    if (nstage == kV4Stages) {
                        nstage = 0;
                        nphase ^= 1;
                    }
                    if (j + 1 < n_tiles) {
                        mbar_wait_sleep(&k_full[nstage], nphase);
                        mbar_wait_sleep(&s_free[t], j & 1);
                        tcgen05_fence_after();
                        issue_s(nstage);
                        umma_commit(&s_full[t]);
                        umma_commit(&k_empty[nstage]);
                    }
                    mbar_wait_sleep(&v_full[stage], phase);
                    mbar_wait_sleep(&p_full[t], j & 1);
                    tcgen05_fence_after();
                    const uint64_t dv = desc_v0 + (sv_a + stage * (kV2KVBytes >> 4));
#pragma unroll
                    for (int ks = 0; ks < kTileKV / 16; ++ks)
                        umma_bf16_ts(tmem_o, tmem_p + ks * 8, dv + ks * ((16 * 128) >> 4), idesc_pv, (j | ks) != 0 ? 1u : 0u);
                    umma_commit(&pv_done[t]);
                    umma_commit(&v_empty[stage]);
                    stage = nstage;
                    phase = nphase;
                }
            }
        }
    } else {
        setmaxnreg_inc<112>();
        const int t = warp_idx >> 3;         // query tile
        const int half = (warp_idx >> 2) & 1;  // key columns [64 half, 64 half + 64) of every tile
        if (t == 0 || two) {
            const int quad = warp_idx & 3;
            const int row = quad * 32 + lane;
            const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
            const uint32_t tmem_s = lane_base + t * 128 + half * 64;
            const uint32_t tmem_o = lane_base + 256 + t * 64 + half * 32;   // this thread's 32 of the 64 O columns
            const uint32_t tmem_p = lane_base + 384 + t * 64 + half * 32;   // 64 keys -> 32 packed columns
            const int pair_bar = 1 + t * 4 + quad;                          // named barrier of the two half-row warps
            float* my_max = xmax + (t * 2 + half) * 128 + row;              // + slot * 512
            const float* other_max = xmax + (t * 2 + (half ^ 1)) * 128 + row;
            const float c = p.scale * kLog2e * q_row_scale(p, batch, q0 + t * kTileQ + row);
            const uint64_t c2 = pack_f32x2(c, c);
            float m_used = -INFINITY;
            uint64_t l2a = pack_f32x2(0.f, 0.f), l2b = l2a;
            uint32_t sfull_probe = 0;
            auto tile = [&](int j, auto tail_tag) {
                constexpr bool TAIL = decltype(tail_tag)::value;
                const int kv0 = (jt0 + j) * kTileKV + half * 64;
                uint32_t s0[32], s1[32];
                if (!__shfl_sync(0xffffffffu, sfull_probe, 0)) warp_mbar_wait(&s_full[t], j & 1, lane);
                sfull_probe = 0;
                tcgen05_fence_after();
                tmem_ld_32x32b_x32(tmem_s + 0, s0);
                tmem_ld_32x32b_x32(tmem_s + 32, s1);
                tmem_ld_wait();
                tcgen05_fence_before();
                warp_mbar_arrive(&s_free[t], lane);  // S_t may be overwritten by the next QK^T
                float mx = fmaxf(max32_v3<TAIL>(s0, kv0, p.Skv), max32_v3<TAIL>(s1, kv0 + 32, p.Skv));
                // the other half of the row lives in warp (w ^ 4): exchange the partial maxima (double-buffered slot)
                my_max[(j & 1) * 512] = mx;
                named_bar_sync(pair_bar, 64);
                mx = fmaxf(mx, other_max[(j & 1) * 512]) * c;
                bool pv_waited = (j == 0);
                if (j == 0) {
                    m_used = mx;
                } else {
                    const float m_new = fmaxf(m_used, mx);
                    // both halves see the same maxima and the same m_used: the decision is identical in both warps
                    if (__any_sync(0xffffffffu, (m_new - m_used) > kRescaleThreshold)) {
                        warp_mbar_wait(&pv_done[t], (j - 1) & 1, lane);
                        pv_waited = true;
                        tcgen05_fence_after();
                        const float alpha = ex2_approx(m_used - m_new);
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(tmem_o, r);  // this half rescales its own 32 O columns
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                        tmem_st_32x32b_x32(tmem_o, r);
                        tmem_st_wait();
                        const uint64_t a2 = pack_f32x2(alpha, alpha), z2 = pack_f32x2(0.f, 0.f);
                        l2a = fma_f32x2(l2a, a2, z2);
                        l2b = fma_f32x2(l2b, a2, z2);
                        m_used = m_new;
                    }
                }
                const uint64_t negm2 = pack_f32x2(-m_used, -m_used);
                uint32_t pk[32];
                uint32_t pv_probe = 0;
                exp_chunk_v4(s0, c2, negm2, l2a, l2b, pk);
                if (!pv_waited && lane == 0) pv_probe = mbar_test_wait(&pv_done[t], (j - 1) & 1);
                if (j + 1 < n_tiles && lane == 0) sfull_probe = mbar_test_wait(&s_full[t], (j + 1) & 1);
                exp_chunk_v4(s1, c2, negm2, l2a, l2b, pk + 16);
                if (!pv_waited) {
                    // P_t is single-buffered: P_t V(j-1) must have read it before it is overwritten
                    if (!__shfl_sync(0xffffffffu, pv_probe, 0)) warp_mbar_wait(&pv_done[t], (j - 1) & 1, lane);
                    tcgen05_fence_after();
                }
                tmem_st_32x32b_x32(tmem_p, pk);
                tmem_st_wait();
                tcgen05_fence_before();
                warp_mbar_arrive(&p_full[t], lane);
            };
            const bool ragged = (jt1 == n_tiles_all) && (p.Skv % kTileKV != 0);
            const int n_full = n_tiles - (ragged ? 1 : 0);
            for (int j = 0; j < n_full; ++j) tile(j, std::false_type{});
            if (ragged) tile(n_full, std::true_type{});
            // row sum: this half's partial + the other half's
            float la, lb, lc, ld;
            unpack_f32x2(l2a, la, lb);
            unpack_f32x2(l2b, lc, ld);
            float l_row = (la + lb) + (lc + ld);
            xsum[(t * 2 + half) * 128 + row] = l_row;
            named_bar_sync(pair_bar, 64);
            {
                // add in a fixed order (half 0 first) so that both halves hold the same bits
                const float l0 = xsum[(t * 2 + 0) * 128 + row], l1 = xsum[(t * 2 + 1) * 128 + row];
                l_row = l0 + l1;
            }
            warp_mbar_wait(&pv_done[t], (n_tiles - 1) & 1, lane);
            tcgen05_fence_after();
            const int qrow = q0 + t * kTileQ + row;
            __nv_bfloat16* orow = attn_out_row(p, batch, qrow < p.Sq ? qrow : 0, head, D) + half * 32;
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem_o, r);
            tmem_ld_wait();
            if (!is_split) {
                const float inv_l = 1.0f / l_row;
                if (qrow < p.Sq) store_row_bf16(orow, r, inv_l);
            } else {
                // ---- split unit: publish (O unnormalised, m, l) of this key range; the LAST of the nsplit CTAs of the
                // unit (ticket counter) merges all partials and writes the output rows ----
                const int su = unit - first_split;
                const int r256 = t * kTileQ + row;
                float4* mine = reinterpret_cast<float4*>(sp.scratch) +
                               static_cast<int64_t>(su * sp.nsplit + split) * (kSplitCols4 * 256) + r256;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    mine[(half * 8 + q) * 256] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                                             __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                if (half == 0) mine[16 * 256] = make_float4(m_used, l_row, 0.f, 0.f);
                __threadfence();
                named_bar_sync(9, 256 * nt);  // all softmax threads of the CTA have published
                if (threadIdx.x == 0) {
                    const int ticket = atomicAdd(sp.counters + su, 1);
                    *split_flag = (ticket == sp.nsplit - 1) ? 1 : 0;
                    if (ticket == sp.nsplit - 1) sp.counters[su] = 0;  // re-arm for the next launch on this stream
                    __threadfence();
                }
                named_bar_sync(9, 256 * nt);
                if (*split_flag != 0) {
                    const float4* base = reinterpret_cast<const float4*>(sp.scratch) +
                                         static_cast<int64_t>(su * sp.nsplit) * (kSplitCols4 * 256) + r256;
                    constexpr int64_t stride = kSplitCols4 * 256;
                    float4 ml[kMaxSplit];
#pragma unroll
                    for (int i = 0; i < kMaxSplit; ++i)
                        ml[i] = i < sp.nsplit ? __ldcg(base + i * stride + 16 * 256) : make_float4(-INFINITY, 0.f, 0.f, 0.f);
                    float m_all = -INFINITY;
#pragma unroll
                    for (int i = 0; i < kMaxSplit; ++i) m_all = fmaxf(m_all, ml[i].x);
                    float wgt[kMaxSplit];
                    float l_all = 0.f;
#pragma unroll
                    for (int i = 0; i < kMaxSplit; ++i) {
                        wgt[i] = i < sp.nsplit ? ex2_approx(ml[i].x - m_all) : 0.f;
                        l_all = fmaf(ml[i].y, wgt[i], l_all);
                    }
                    const float inv_l = 1.0f / l_all;
#pragma unroll 1
                    for (int dc = 0; dc < 2; ++dc) {  // this half's 32 output columns, 16 at a time
                        float acc[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q) acc[q] = 0.f;
#pragma unroll
                        for (int i = 0; i < kMaxSplit; ++i) {
                            if (i < sp.nsplit) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float4 v = __ldcg(base + i * stride + (half * 8 + dc * 4 + q) * 256);
                                    acc[4 * q + 0] = fmaf(v.x, wgt[i], acc[4 * q + 0]);
                                    acc[4 * q + 1] = fmaf(v.y, wgt[i], acc[4 * q + 1]);
                                    acc[4 * q + 2] = fmaf(v.z, wgt[i], acc[4 * q + 2]);
                                    acc[4 * q + 3] = fmaf(v.w, wgt[i], acc[4 * q + 3]);
                                }
                            }
                        }
                        if (qrow < p.Sq) {
                            uint4* d4 = reinterpret_cast<uint4*>(orow + dc * 16);
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                uint4 u;
                                u.x = pack_bf16x2(acc[8 * q + 0] * inv_l, acc[8 * q + 1] * inv_l);
                                u.y = pack_bf16x2(acc[8 * q + 2] * inv_l, acc[8 * q + 3] * inv_l);
                                u.z = pack_bf16x2(acc[8 * q + 4] * inv_l, acc[8 * q + 5] * inv_l);
                                u.w = pack_bf16x2(acc[8 * q + 6] * inv_l, acc[8 * q + 7] * inv_l);
                                d4[q] = u;
                            }
                        }
                    }
                }
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp_idx == 19) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ================================================================================================
// Self-attention, head_dim 128 (the 13B preset: 32 heads x 128).  Same organisation as flash_attn3_kernel (two query
// tiles per CTA sharing every K/V tile, register-resident scores, P in TMEM, one UMMA issuer per tile), with the
// differences head_dim 128 forces:
//   * TMEM is full with S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512): P_t (64 packed bf16x2 columns) ALIASES the
//     first half of S_t.  A softmax thread holds its whole S row in registers before it writes P, and the issuer emits
//     QK^T(j+1) of a tile only after P_t V(j) in program order (the tensor pipe executes one thread's MMAs in order),
//     so S_t is overwritten only after P_t has been consumed; the OTHER tile's MMAs fill the pipe meanwhile.
//   * s_full(j+1) is committed after P V(j), hence also signals "O_t may be rescaled": no pv_done wait in the loop.
//   * K / V tiles are 32 KB (two 64-column swizzle atoms, 16 KB apart): 2-stage rings, Q 2 x 32 KB -> 192 KB smem.
//   * MUFU is no longer the bound (twice the FLOPs per exponential): per key tile 1024 clk of tensor work against
//     768 clk of exp2 per query tile.
// No tail splitting (c4 runs 2432 units = 16.4 waves).
// ================================================================================================
constexpr int kV5Threads = 384;
constexpr int kV5Stages = 2;
constexpr int kV5QBytes = kTileQ * 128 * 2;    // 32 KB per query tile
constexpr int kV5KVBytes = kTileKV * 128 * 2;  // 32 KB per K or V tile
constexpr int kV5SmemBytes = 2 * kV5QBytes + 2 * kV5Stages * kV5KVBytes + 512;

__global__ void __launch_bounds__(kV5Threads, 1)
flash_attn3_d128_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                        const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ AttnParams p, const int n_qb) {
    constexpr int D = 128;
    constexpr int kHalf = kTileKV * 128;  // bytes of one 64-column atom of a 128-row tile
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sq = smem;
    uint8_t* sk = sq + 2 * kV5QBytes;
    uint8_t* sv = sk + kV5Stages * kV5KVBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sv + kV5Stages * kV5KVBytes);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;
    uint64_t* k_empty = bars + 1 + kV5Stages;
    uint64_t* v_full = bars + 1 + 2 * kV5Stages;
    uint64_t* v_empty = bars + 1 + 3 * kV5Stages;
    uint64_t* s_full = bars + 1 + 4 * kV5Stages;  // [2]
    uint64_t* p_full = s_full + 2;                // [2]
    uint64_t* pv_done = p_full + 2;               // [2] last P V of the tile retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int unit = blockIdx.x;
    const int qb = unit % n_qb;
    const int head = (unit / n_qb) % p.H;
    const int batch = unit / (n_qb * p.H);
    const int q0 = qb * (2 * kTileQ);
    const int n_tiles = (p.Skv + kTileKV - 1) / kTileKV;
    const bool two = (q0 + kTileQ) < p.Sq;
    const int nt = two ? 2 : 1;

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) {
            printf("ltxv attention d128: dynamic smem base not 1024B aligned\n");
            __trap();
        }
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
        mbar_init(q_full, 1);
        for (int i = 0; i < kV5Stages; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], nt);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], nt);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            mbar_init(&p_full[t], 4);  // one arrival per softmax warp
            mbar_init(&pv_done[t], 1);
        }
        fence_barrier_init();
    }
    if (warp_idx == 11) tmem_alloc<512>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_launch_dependents();
    griddep_wait();

    if (warp_idx >= 8) {
        setmaxnreg_dec<48>();
        if (warp_idx == 8) {
            // ===================== TMA producer =====================
            if (elect_one()) {
                mbar_arrive_expect_tx(q_full, nt * kV5QBytes);
                for (int t = 0; t < nt; ++t)
                    for (int h = 0; h < 2; ++h)
                        tma_load_3d(sq + t * kV5QBytes + h * kHalf, &tm_q, q_full, p.q_col0 + head * D + h * 64,
                                    q0 + t * kTileQ, batch);
                int stage = 0;
                uint32_t phase = 0;
                for (int j = 0; j < n_tiles; ++j) {
                    const int kv0 = j * kTileKV;
                    if (j + kV5Stages + 1 < n_tiles) {
                        const int kvp = kv0 + (kV5Stages + 1) * kTileKV;
                        for (int h = 0; h < 2; ++h) {
                            tma_prefetch_3d(&tm_k, p.k_col0 + head * D + h * 64, kvp, batch);
                            tma_prefetch_3d(&tm_v, p.v_col0 + head * D + h * 64, kvp, batch);
                        }
                    }
                    mbar_wait_sleep(&k_empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&k_full[stage], kV5KVBytes);
                    for (int h = 0; h < 2; ++h)
                        tma_load_3d(sk + stage * kV5KVBytes + h * kHalf, &tm_k, &k_full[stage],
                                    p.k_col0 + head * D + h * 64, kv0, batch);
                    mbar_wait_sleep(&v_empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&v_full[stage], kV5KVBytes);
                    for (int h = 0; h < 2; ++h)
                        tma_load_3d(sv + stage * kV5KVBytes + h * kHalf, &tm_v, &v_full[stage],
                                    p.v_col0 + head * D + h * 64, kv0, batch);
                    if (++stage == kV5Stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        } else if (warp_idx == 9 || (warp_idx == 10 && two)) {
            // ===================== UMMA issuer of query tile t =====================
            const int t = warp_idx - 9;
            if (elect_one()) {
                constexpr uint32_t idesc_s = make_idesc_bf16(kTileQ, kTileKV, false, false);
                constexpr uint32_t idesc_pv = make_idesc_bf16(kTileQ, D, false, true);  // B = V is MN-major
                const uint32_t qa = smem_u32(sq) + t * kV5QBytes;
                const uint32_t tmem_s = tmem_base + t * 128;
                const uint32_t tmem_o = tmem_base + 256 + t * 128;
                const uint32_t tmem_p = tmem_s;  // aliases S_t columns [0, 64)
                auto issue_s = [&](int stage) {
                    const uint32_t k_addr = smem_u32(sk + stage * kV5KVBytes);
#pragma unroll
                    for (int ks = 0; ks < D / 16; ++ks) {
                        const uint32_t off = (ks >> 2) * kHalf + (ks & 3) * 32;
                        umma_bf16_ss(tmem_s, make_smem_desc_sw128(qa + off, 1024, 0),
                                     make_smem_desc_sw128(k_addr + off, 1024, 0), idesc_s, ks != 0 ? 1u : 0u);
                    }
                };
                mbar_wait_sleep(q_full, 0);
                mbar_wait_sleep(&k_full[0], 0);
                tcgen05_fence_after();
                issue_s(0);
                umma_commit(&s_full[t]);
                umma_commit(&k_empty[0]);
                int stage = 0;
                uint32_t phase = 0;
                for (int j = 0; j < n_tiles; ++j) {
                    int nstage = stage + 1;
                    uint32_t nphase = phase;
                    if (nstage == kV5Stages) {
                        nstage = 0;
                        nphase ^= 1;
                    }
                    mbar_wait_sleep(&v_full[stage], phase);
                    mbar_wait_sleep(&p_full[t], j & 1);
                    tcgen05_fence_after();
                    const uint32_t v_addr = smem_u32(sv + stage * kV5KVBytes);
#pragma unroll
                    for (int ks = 0; ks < kTileKV / 16; ++ks)
                        umma_bf16_ts(tmem_o, tmem_p + ks * 8,
                                     make_smem_desc_sw128(v_addr + ks * (16 * 128), 1024, kHalf), idesc_pv,
                                     (j | ks) != 0 ? 1u : 0u);
                    umma_commit(&v_empty[stage]);
                    if (j + 1 < n_tiles) {
                        // S_t(j+1) overwrites S_t / P_t: in program order behind P_t V(j)
                        mbar_wait_sleep(&k_full[nstage], nphase);
                        tcgen05_fence_after();
                        issue_s(nstage);
                        umma_commit(&s_full[t]);
                        umma_commit(&k_empty[nstage]);
                    } else {
                        umma_commit(&pv_done[t]);
                    }
                    stage = nstage;
                    phase = nphase;
                }
            }
        }
    } else {
        setmaxnreg_inc<224>();
        const int t = warp_idx >> 2;
        if (t == 0 || two) {
            const int quad = warp_idx & 3;
            const int row = quad * 32 + lane;
            const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
            const uint32_t tmem_s = lane_base + t * 128;
            const uint32_t tmem_o = lane_base + 256 + t * 128;
            const uint32_t tmem_p = tmem_s;
            const float c = p.scale * kLog2e * q_row_scale(p, batch, q0 + t * kTileQ + row);
            const uint64_t c2 = pack_f32x2(c, c);
            float m_used = -INFINITY;
            uint64_t l2a = pack_f32x2(0.f, 0.f), l2b = l2a;
            auto tile = [&](int j, auto tail_tag) {
                constexpr bool TAIL = decltype(tail_tag)::value;
                const int kv0 = j * kTileKV;
                uint32_t s0[32], s1[32], s2[32], s3[32];
                warp_mbar_wait(&s_full[t], j & 1, lane);  // also: P_t V(j-1) retired, O_t is quiescent
                tcgen05_fence_after();
                tmem_ld_32x32b_x32(tmem_s + 0, s0);
                tmem_ld_32x32b_x32(tmem_s + 32, s1);
                tmem_ld_32x32b_x32(tmem_s + 64, s2);
                tmem_ld_32x32b_x32(tmem_s + 96, s3);
                tmem_ld_wait();
                float mx = fmaxf(fmaxf(max32_v3<TAIL>(s0, kv0, p.Skv), max32_v3<TAIL>(s1, kv0 + 32, p.Skv)),
                                 fmaxf(max32_v3<TAIL>(s2, kv0 + 64, p.Skv), max32_v3<TAIL>(s3, kv0 + 96, p.Skv)));
                mx *= c;
                if (j == 0) {
                    m_used = mx;
                } else {
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
                        const uint64_t a2 = pack_f32x2(alpha, alpha), z2 = pack_f32x2(0.f, 0.f);
                        l2a = fma_f32x2(l2a, a2, z2);
                        l2b = fma_f32x2(l2b, a2, z2);
                        m_used = m_new;
                    }
                }
                const uint64_t negm2 = pack_f32x2(-m_used, -m_used);
                uint32_t pk[32];
                exp_chunk_v3(s0, c2, negm2, l2a, l2b, pk);
                exp_chunk_v3(s1, c2, negm2, l2a, l2b, pk + 16);
                tmem_st_32x32b_x32(tmem_p, pk);       // keys  0..63  -> P columns  0..31 (over S columns 0..31)
                exp_chunk_v3(s2, c2, negm2, l2a, l2b, pk);
                exp_chunk_v3(s3, c2, negm2, l2a, l2b, pk + 16);
                tmem_st_32x32b_x32(tmem_p + 32, pk);  // keys 64..127 -> P columns 32..63
                tmem_st_wait();
                tcgen05_fence_before();
                warp_mbar_arrive(&p_full[t], lane);
            };
            const bool ragged = (p.Skv % kTileKV != 0);
            const int n_full = n_tiles - (ragged ? 1 : 0);
            for (int j = 0; j < n_full; ++j) tile(j, std::false_type{});
            if (ragged) tile(n_full, std::true_type{});
            warp_mbar_wait(&pv_done[t], 0, lane);
            tcgen05_fence_after();
            float la, lb, lc, ld;
            unpack_f32x2(l2a, la, lb);
            unpack_f32x2(l2b, lc, ld);
            const float inv_l = 1.0f / ((la + lb) + (lc + ld));
            const int qrow = q0 + t * kTileQ + row;
            __nv_bfloat16* orow = attn_out_row(p, batch, qrow < p.Sq ? qrow : 0, head, D);
#pragma unroll 1
            for (int dc = 0; dc < D / 32; ++dc) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_o + dc * 32, r);
                tmem_ld_wait();
                if (qrow < p.Sq) store_row_bf16(orow + dc * 32, r, inv_l);
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp_idx == 11) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}



// ================================================================================================
// Cross-attention (head_dim 64, <= 128 keys: the T5 text tokens).  The whole key axis is ONE tile, so there is no online
// softmax and no key loop to pipeline; what has to be amortised is the per-CTA fixed cost.  The general kernel above
// spends one CTA per (head, 128 queries): 1248 CTAs of ~15 us each at c2, 62 us per launch for 5 GFLOP.  Here a CTA
// owns one (batch, head) and a contiguous range of query tiles: K and V are loaded once, Q tiles stream through a
// 2-deep TMA ring per softmax group, and the two groups (two query tiles in flight, as in v3) overlap one tile's
// exponentials with the other's TMEM load / PV MMA / output stores.  P stays in TMEM (TS-form MMA).
// TMEM map: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384) P0 [384,448) P1 [448,512).
// ================================================================================================
constexpr int kXThreads = 384;
constexpr int kXQStages = 2;
constexpr int kXSmemBytes = 2 * kV2KVBytes + 2 * kXQStages * kV2QBytes + 1024;

__global__ void __launch_bounds__(kXThreads, 1)
cross_attn_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                  const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ AttnParams p, const int ctas_per_head) {
    constexpr int D = 64;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sk = smem;
    uint8_t* sv = sk + kV2KVBytes;
    uint8_t* sq = sv + kV2KVBytes;  // [group][stage] 16 KB each
    float* sbias = reinterpret_cast<float*>(sq + 2 * kXQStages * kV2QBytes);  // [128] bias * log2e, -inf beyond Skv
    uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + 128);
    uint64_t* kv_full = bars + 0;
    uint64_t* q_full = bars + 1;                    // [2][stages]
    uint64_t* q_empty = q_full + 2 * kXQStages;     // [2][stages]
    uint64_t* s_full = q_empty + 2 * kXQStages;     // [2]
    uint64_t* s_free = s_full + 2;                  // [2] 4 warp arrivals
    uint64_t* p_full = s_free + 2;                  // [2] 4 warp arrivals
    uint64_t* pv_done = p_full + 2;                 // [2]
    uint64_t* o_free = pv_done + 2;                 // [2] 4 warp arrivals
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int bh = blockIdx.x / ctas_per_head;
    const int part = blockIdx.x % ctas_per_head;
    const int head = bh % p.H;
    const int batch = bh / p.H;
    const int nq = (p.Sq + kTileQ - 1) / kTileQ;
    const int qt0 = (part * nq) / ctas_per_head;  // this CTA's query tiles [qt0, qt1)
    const int qt1 = ((part + 1) * nq) / ctas_per_head;
    // group g handles tiles qt0 + g, qt0 + g + 2, ...
    auto n_of = [&](int g) { return (qt1 - qt0 - g + 1) / 2; };

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
        mbar_init(kv_full, 1);
        for (int i = 0; i < 2 * kXQStages; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 1);
        }
        for (int g = 0; g < 2; ++g) {
            mbar_init(&s_full[g], 1);
            mbar_init(&s_free[g], 4);
            mbar_init(&p_full[g], 4);
            mbar_init(&pv_done[g], 1);
            mbar_init(&o_free[g], 4);
        }
        fence_barrier_init();
    }
    griddep_launch_dependents();
    griddep_wait();  // the key bias / Q / K / V come from earlier kernels of the stream
    if (threadIdx.x < 128) {
        const int k = threadIdx.x;
        float b = (k < p.Skv) ? 0.f : -INFINITY;
        if (p.kv_bias != nullptr && k < p.Skv) b = __ldg(p.kv_bias + static_cast<int64_t>(batch) * p.Skv + k) * kLog2e;
        sbias[k] = b;
    }
    if (warp_idx == 11) tmem_alloc<512>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp_idx >= 8) {
        setmaxnreg_dec<48>();
        if (warp_idx == 8) {
            // ===================== TMA producer: K, V once; Q tiles of both groups =====================
            if (elect_one()) {
                mbar_arrive_expect_tx(kv_full, 2 * kV2KVBytes);
                tma_load_3d(sk, &tm_k, kv_full, p.k_col0 + head * D, 0, batch);
                tma_load_3d(sv, &tm_v, kv_full, p.v_col0 + head * D, 0, batch);
                const int n0 = n_of(0), n1 = n_of(1);
                for (int j = 0; j < n0; ++j) {
                    for (int g = 0; g < 2; ++g) {
                        if (j >= (g == 0 ? n0 : n1)) continue;
                        const int stage = j % kXQStages;
                        const uint32_t phase = (j / kXQStages) & 1;
                        uint64_t* full = &q_full[g * kXQStages + stage];
                        mbar_wait_sleep(&q_empty[g * kXQStages + stage], phase ^ 1);
                        mbar_arrive_expect_tx(full, kV2QBytes);
                        tma_load_3d(sq + (g * kXQStages + stage) * kV2QBytes, &tm_q, full, p.q_col0 + head * D,
                                    (qt0 + 2 * j + g) * kTileQ, batch);
                    }
                }
            }
        } else if (warp_idx == 9 || warp_idx == 10) {
            // ===================== UMMA issuer of group g =====================
            const int g = warp_idx - 9;
            const int n = n_of(g);
            if (n > 0 && elect_one()) {
                constexpr uint32_t idesc_s = make_idesc_bf16(kTileQ, kTileKV, false, false);
                constexpr uint32_t idesc_pv = make_idesc_bf16(kTileQ, D, false, true);
                const uint32_t k_addr = smem_u32(sk), v_addr = smem_u32(sv);
                const uint32_t tmem_s = tmem_base + g * 128;
                const uint32_t tmem_o = tmem_base + 256 + g * 64;
                const uint32_t tmem_p = tmem_base + 384 + g * 64;
                mbar_wait_sleep(kv_full, 0);
                for (int j = 0; j < n; ++j) {
                    const int stage = j % kXQStages;
                    const uint32_t phase = (j / kXQStages) & 1;
                    mbar_wait_sleep(&q_full[g * kXQStages + stage], phase);
                    if (j > 0) mbar_wait_sleep(&s_free[g], (j - 1) & 1);
                    tcgen05_fence_after();
                    const uint32_t qa = smem_u32(sq + (g * kXQStages + stage) * kV2QBytes);
#pragma unroll
                    for (int ks = 0; ks < D / 16; ++ks)
                        umma_bf16_ss(tmem_s, make_smem_desc_sw128(qa + ks * 32, 1024, 0),
                                     make_smem_desc_sw128(k_addr + ks * 32, 1024, 0), idesc_s, ks != 0 ? 1u : 0u);
                    umma_commit(&s_full[g]);
                    umma_commit(&q_empty[g * kXQStages + stage]);
                    mbar_wait_sleep(&p_full[g], j & 1);
                    if (j > 0) mbar_wait_sleep(&o_free[g], (j - 1) & 1);
                    tcgen05_fence_after();
#pragma unroll
                    for (int ks = 0; ks < kTileKV / 16; ++ks)
                        umma_bf16_ts(tmem_o, tmem_p + ks * 8,
                                     make_smem_desc_sw128(v_addr + ks * (16 * 128), 1024, kTileKV * 128), idesc_pv,
                                     ks != 0 ? 1u : 0u);
                    umma_commit(&pv_done[g]);
                }
            }
        }
    } else {
        setmaxnreg_inc<224>();
        const int g = warp_idx >> 2;
        const int n = n_of(g);
        const int quad = warp_idx & 3;
        const int row = quad * 32 + lane;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        const uint32_t tmem_s = lane_base + g * 128;
        const uint32_t tmem_o = lane_base + 256 + g * 64;
        const uint32_t tmem_p = lane_base + 384 + g * 64;
        const uint64_t one2 = pack_f32x2(1.0f, 1.0f);
        // the row factor of tile j + 1 is loaded one tile ahead: issued right before its use, the dependent global load
        // (behind a pointer the compiler keeps on the local stack) cost 17 % of the kernel's stall samples
        float qr_next = q_row_scale(p, batch, (qt0 + g) * kTileQ + row);
        for (int j = 0; j < n; ++j) {
            const float c = p.scale * kLog2e * qr_next;
            if (j + 1 < n) qr_next = q_row_scale(p, batch, (qt0 + 2 * (j + 1) + g) * kTileQ + row);
            const uint64_t c2 = pack_f32x2(c, c);
            uint32_t s0[32], s1[32], s2[32], s3[32];
            warp_mbar_wait(&s_full[g], j & 1, lane);
            tcgen05_fence_after();
            tmem_ld_32x32b_x32(tmem_s + 0, s0);
            tmem_ld_32x32b_x32(tmem_s + 32, s1);
            tmem_ld_32x32b_x32(tmem_s + 64, s2);
            tmem_ld_32x32b_x32(tmem_s + 96, s3);
            tmem_ld_wait();
            tcgen05_fence_before();
            warp_mbar_arrive(&s_free[g], lane);
            // t = s * scale*log2e + bias*log2e  (bias: additive key mask, -inf beyond the last key), in place
            auto prep = [&](uint32_t (&r)[32], int col0) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b = *reinterpret_cast<const float4*>(sbias + col0 + i);
                    float x0, x1, x2, x3;
                    unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), c2, pack_f32x2(b.x, b.y)), x0, x1);
                    unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), c2, pack_f32x2(b.z, b.w)), x2, x3);
                    r[i] = __float_as_uint(x0);
                    r[i + 1] = __float_as_uint(x1);
                    r[i + 2] = __float_as_uint(x2);
                    r[i + 3] = __float_as_uint(x3);
                }
            };
            prep(s0, 0);
            prep(s1, 32);
            prep(s2, 64);
            prep(s3, 96);
            const float mx = fmaxf(fmaxf(max32_v3<false>(s0, 0, 128), max32_v3<false>(s1, 0, 128)),
                                   fmaxf(max32_v3<false>(s2, 0, 128), max32_v3<false>(s3, 0, 128)));
            const uint64_t negm2 = pack_f32x2(-mx, -mx);
            uint64_t l2a = pack_f32x2(0.f, 0.f), l2b = l2a;
            uint32_t pk[32];
            exp_chunk_v3(s0, one2, negm2, l2a, l2b, pk);
            exp_chunk_v3(s1, one2, negm2, l2a, l2b, pk + 16);
            tmem_st_32x32b_x32(tmem_p, pk);
            exp_chunk_v3(s2, one2, negm2, l2a, l2b, pk);
            exp_chunk_v3(s3, one2, negm2, l2a, l2b, pk + 16);
            tmem_st_32x32b_x32(tmem_p + 32, pk);
            tmem_st_wait();
            tcgen05_fence_before();
            warp_mbar_arrive(&p_full[g], lane);
            float la, lb, lc, ld;
            unpack_f32x2(l2a, la, lb);
            unpack_f32x2(l2b, lc, ld);
            const float inv_l = 1.0f / ((la + lb) + (lc + ld));
            const int qrow = (qt0 + 2 * j + g) * kTileQ + row;
            __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) +
                                  (static_cast<int64_t>(batch) * p.Sq + (qrow < p.Sq ? qrow : 0)) * p.ldo + head * D;
            warp_mbar_wait(&pv_done[g], j & 1, lane);
            tcgen05_fence_after();
            uint32_t o0[32], o1[32];
            tmem_ld_32x32b_x32(tmem_o, o0);
            tmem_ld_32x32b_x32(tmem_o + 32, o1);
            tmem_ld_wait();
            tcgen05_fence_before();
            warp_mbar_arrive(&o_free[g], lane);
            if (qrow < p.Sq) {
                store_row_bf16(orow, o0, inv_l);
                store_row_bf16(orow + 32, o1, inv_l);
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp_idx == 11) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ================================================================================================
// CUDA-core fallback for SMALL head dims (8 .. 32, multiples of 8): the reference's own golden DiT geometries are
// 2 x 16 and 4 x 16 heads (scripts/gen_dit_ref.py:12-38, tests/verify_rope_parity.rs:537-567), which the tcgen05
// kernels (UMMA K = 16 per instruction over 64- / 128-wide swizzled rows) do not serve.  One thread per query row,
// 64-key tiles staged in shared memory, two passes per tile (scores + max, then exp and P V), f32 throughout.  Not a
// performance path: it exists so that fixtures generated for the reference's tests can be loaded and compared.
// ================================================================================================
constexpr int kSimtKeys = 64;
template <int DMAX>
__global__ void __launch_bounds__(128)
flash_attn_simt_kernel(const AttnParams p) {
    __shared__ __align__(16) __nv_bfloat16 sk[kSimtKeys * DMAX];
    __shared__ __align__(16) __nv_bfloat16 sv[kSimtKeys * DMAX];
    __shared__ float sbias[kSimtKeys];
    const int D = p.D;
    const int head = blockIdx.y, batch = blockIdx.z;
    const int qrow = blockIdx.x * 128 + threadIdx.x;
    const bool valid = qrow < p.Sq;
    const __nv_bfloat16* qp = reinterpret_cast<const __nv_bfloat16*>(p.q) +
                              (static_cast<int64_t>(batch) * p.Sq + (valid ? qrow : 0)) * p.ldq + p.q_col0 + head * D;
    float q[DMAX], o[DMAX];
#pragma unroll
    for (int i = 0; i < DMAX; ++i) {
        q[i] = i < D ? __bfloat162float(qp[i]) : 0.f;
        o[i] = 0.f;
    }
    const float c = p.scale * kLog2e * q_row_scale(p, batch, valid ? qrow : 0);
    float m = -INFINITY, l = 0.f;
    const int vec_per_row = D >> 3;  // 16-byte pieces per key row
    for (int kv0 = 0; kv0 < p.Skv; kv0 += kSimtKeys) {
        const int nk = min(kSimtKeys, p.Skv - kv0);
        __syncthreads();
        for (int i = threadIdx.x; i < kSimtKeys * vec_per_row; i += 128) {
            const int r = i / vec_per_row, cpos = (i - r * vec_per_row) * 8;
            uint4 ku = make_uint4(0u, 0u, 0u, 0u), vu = ku;
            if (r < nk) {
                const int64_t row = static_cast<int64_t>(batch) * p.Skv + kv0 + r;
                ku = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.k) + row * p.ldk + p.k_col0 + head * D + cpos);
                vu = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.v) + row * p.ldv + p.v_col0 + head * D + cpos);
            }
            *reinterpret_cast<uint4*>(sk + r * DMAX + cpos) = ku;
            *reinterpret_cast<uint4*>(sv + r * DMAX + cpos) = vu;
        }
        if (threadIdx.x < kSimtKeys)
            sbias[threadIdx.x] = (p.kv_bias != nullptr && threadIdx.x < nk)
                                     ? p.kv_bias[static_cast<int64_t>(batch) * p.Skv + kv0 + threadIdx.x] * kLog2e : 0.f;
        __syncthreads();
        float sc[kSimtKeys];
        float mx = -INFINITY;
#pragma unroll 4
        for (int r = 0; r < kSimtKeys; ++r) {
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < DMAX; ++i)
                if (i < D) a = fmaf(q[i], __bfloat162float(sk[r * DMAX + i]), a);
            a = r < nk ? fmaf(a, c, sbias[r]) : -INFINITY;
            sc[r] = a;
            mx = fmaxf(mx, a);
        }
        const float m_new = fmaxf(m, mx);
        const float alpha = m == -INFINITY ? 0.f : ex2_approx(m - m_new);
        l *= alpha;
#pragma unroll
        for (int i = 0; i < DMAX; ++i) o[i] *= alpha;
#pragma unroll 4
        for (int r = 0; r < kSimtKeys; ++r) {
            const float pr = ex2_approx(sc[r] - m_new);  // exp2(-inf) = 0 for the masked tail
            l += pr;
#pragma unroll
            for (int i = 0; i < DMAX; ++i)
                if (i < D) o[i] = fmaf(pr, __bfloat162float(sv[r * DMAX + i]), o[i]);
        }
        m = m_new;
    }
    if (!valid) return;
    const float inv_l = 1.0f / l;
    __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + (static_cast<int64_t>(batch) * p.Sq + qrow) * p.ldo + head * D;
    for (int i = 0; i < D; i += 2)
        *reinterpret_cast<uint32_t*>(op + i) = pack_bf16x2(o[i] * inv_l, o[i + 1] * inv_l);
}

std::atomic<uint64_t> g_attn_launches{0};

// per-device scratch for the tail split + ticket counters
struct SplitScratch {
    float* scratch = nullptr;
    int* counters = nullptr;
    int n_sm = 0;
};
// One scratch area per (device, stream): launches on the same stream are ordered, so they may share it; two streams
// (two models driven by two host threads) must not, or their partial (O, m, l) rows and tickets would interleave.
cudaError_t split_scratch(SplitScratch** out, cudaStream_t stream) {
    static std::map<std::pair<int, cudaStream_t>, SplitScratch> per_stream;
    static std::mutex mu;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    SplitScratch& s = per_stream[std::make_pair(dev, stream)];
    if (s.scratch == nullptr) {
        e = cudaDeviceGetAttribute(&s.n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        // up to (n_sm - 1) tail units x kMaxSplit key ranges x 256 rows x 17 float4 = 82 MB on a 148-SM part
        e = cudaMalloc(&s.scratch, static_cast<size_t>(s.n_sm) * kMaxSplit * 256 * kSplitCols4 * sizeof(float4));
        if (e != cudaSuccess) return e;
        e = cudaMalloc(&s.counters, static_cast<size_t>(s.n_sm) * sizeof(int));
        if (e != cudaSuccess) return e;
        e = cudaMemset(s.counters, 0, static_cast<size_t>(s.n_sm) * sizeof(int));
        if (e != cudaSuccess) return e;
    }
    *out = &s;
    return cudaSuccess;
}

cudaError_t launch_attn3_impl(const AttnParams& p, cudaStream_t stream) {
    static PerDeviceOnce configured;
    int cfg_dev = 0;
    if (configured.need(&cfg_dev)) {
        cudaError_t e =
            cudaFuncSetAttribute(flash_attn3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kV3SmemBytes);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(flash_attn4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kV4SmemBytes);
        if (e != cudaSuccess) return e;
        configured.mark(cfg_dev);
    }
    CUtensorMap tq, tk, tv;
    cudaError_t e = make_tensor_map_3d_bf16(&tq, p.q, p.B, p.Sq, p.ldq, kTileQ, 64, p.ldq, p.ldq * (int64_t)p.Sq);
    if (e != cudaSuccess) return e;
    e = make_tensor_map_3d_bf16(&tk, p.k, p.B, p.Skv, p.ldk, kTileKV, 64, p.ldk, p.ldk * (int64_t)p.Skv);
    if (e != cudaSuccess) return e;
    e = make_tensor_map_3d_bf16(&tv, p.v, p.B, p.Skv, p.ldv, kTileKV, 64, p.ldv, p.ldv * (int64_t)p.Skv);
    if (e != cudaSuccess) return e;
    SplitScratch* ss = nullptr;
    e = split_scratch(&ss, stream);
    if (e != cudaSuccess) return e;
    SplitPlan sp{};
    sp.n_qb = (p.Sq + 2 * kTileQ - 1) / (2 * kTileQ);
    sp.n_units = p.B * p.H * sp.n_qb;
    sp.nsplit = 1;
    sp.n_split_units = 0;
    sp.scratch = ss->scratch;
    sp.counters = ss->counters;
    const int n_tiles = (p.Skv + kTileKV - 1) / kTileKV;
    const int rem = sp.n_units % ss->n_sm;
    if (rem != 0 && !options().attn_nosplit) {
        // key ranges per tail unit: minimise the length of the tail, ceil(rem * ns / SMs) rounds of 1/ns unit each
        // (plus a few percent per extra range for its pipeline fill and the merge), keeping >= 4 key tiles per range
        int max_ns = n_tiles / 4;
        if (max_ns > kMaxSplit) max_ns = kMaxSplit;
        if (options().attn_nsplit_max > 0 && options().attn_nsplit_max < max_ns) max_ns = options().attn_nsplit_max;  // experiment knob
        int best_ns = 1;
        double best = 1.0;
        for (int ns = 2; ns <= max_ns; ++ns) {
            const double rounds = static_cast<double>((rem * ns + ss->n_sm - 1) / ss->n_sm);
            const double cost = rounds / ns + 0.03 * (ns - 1);
            if (cost < best - 1e-9) {
                best = cost;
                best_ns = ns;
            }
        }
        if (best_ns >= 2) {
            sp.nsplit = best_ns;
            sp.n_split_units = rem;
        }
    }
    const int n_items = sp.n_units - sp.n_split_units + sp.n_split_units * sp.nsplit;
    {
        ProfScope prof(PROF_ATTN_SELF, 4.0 * p.B * p.H * static_cast<double>(p.Sq) * p.Skv * 64, stream);
        cudaError_t le;
        if (!options().attn_v4) {
            LTXV_TRACE_VARIANT("flash_attn3_kernel nsplit=%d split_units=%d", sp.nsplit, sp.n_split_units);
            le = launch_pdl(flash_attn3_kernel, dim3(n_items), dim3(kV3Threads), kV3SmemBytes, stream, tq, tk, tv, p, sp);
        } else {
            LTXV_TRACE_VARIANT("flash_attn4_kernel nsplit=%d split_units=%d", sp.nsplit, sp.n_split_units);
            le = launch_pdl(flash_attn4_kernel, dim3(n_items), dim3(kV4Threads), kV4SmemBytes, stream, tq, tk, tv, p, sp);
        }
        if (le != cudaSuccess) return le;
    }
    g_attn_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}


cudaError_t launch_attn3_d128_impl(const AttnParams& p, cudaStream_t stream) {
    static PerDeviceOnce configured;
    int cfg_dev = 0;
    if (configured.need(&cfg_dev)) {
        cudaError_t e = cudaFuncSetAttribute(flash_attn3_d128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kV5SmemBytes);
        if (e != cudaSuccess) return e;
        configured.mark(cfg_dev);
    }
    CUtensorMap tq, tk, tv;
    cudaError_t e = make_tensor_map_3d_bf16(&tq, p.q, p.B, p.Sq, p.ldq, kTileQ, 64, p.ldq, p.ldq * (int64_t)p.Sq);
    if (e != cudaSuccess) return e;
    e = make_tensor_map_3d_bf16(&tk, p.k, p.B, p.Skv, p.ldk, kTileKV, 64, p.ldk, p.ldk * (int64_t)p.Skv);
    if (e != cudaSuccess) return e;
    e = make_tensor_map_3d_bf16(&tv, p.v, p.B, p.Skv, p.ldv, kTileKV, 64, p.ldv, p.ldv * (int64_t)p.Skv);
    if (e != cudaSuccess) return e;
    const int n_qb = (p.Sq + 2 * kTileQ - 1) / (2 * kTileQ);
    {
        ProfScope prof(PROF_ATTN_SELF, 4.0 * p.B * p.H * static_cast<double>(p.Sq) * p.Skv * 128, stream);
        LTXV_TRACE_VARIANT("flash_attn3_d128_kernel");
        cudaError_t le = launch_pdl(flash_attn3_d128_kernel, dim3(p.B * p.H * n_qb), dim3(kV5Threads), kV5SmemBytes, stream,
                                    tq, tk, tv, p, n_qb);
        if (le != cudaSuccess) return le;
    }
    g_attn_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

cudaError_t launch_cross_attn_impl(const AttnParams& p, cudaStream_t stream) {
    static PerDeviceOnce configured;
    int cfg_dev = 0;
    static int num_sms = 148;
    if (configured.need(&cfg_dev)) {
        cudaError_t e = cudaFuncSetAttribute(cross_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kXSmemBytes);
        if (e != cudaSuccess) return e;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        configured.mark(cfg_dev);
    }
    CUtensorMap tq, tk, tv;
    cudaError_t e = make_tensor_map_3d_bf16(&tq, p.q, p.B, p.Sq, p.ldq, kTileQ, 64, p.ldq, p.ldq * (int64_t)p.Sq);
    if (e != cudaSuccess) return e;
    e = make_tensor_map_3d_bf16(&tk, p.k, p.B, p.Skv, p.ldk, kTileKV, 64, p.ldk, p.ldk * (int64_t)p.Skv);
    if (e != cudaSuccess) return e;
    e = make_tensor_map_3d_bf16(&tv, p.v, p.B, p.Skv, p.ldv, kTileKV, 64, p.ldv, p.ldv * (int64_t)p.Skv);
    if (e != cudaSuccess) return e;
    const int nq = (p.Sq + kTileQ - 1) / kTileQ;
    int cph = num_sms / (p.B * p.H);  // CTAs per (batch, head): fill the SMs in one wave
    if (cph < 1) cph = 1;
    if (cph > (nq + 1) / 2) cph = (nq + 1) / 2;  // at least one tile pair per CTA
    if (cph < 1) cph = 1;
    {
        ProfScope prof(PROF_ATTN_CROSS, 4.0 * p.B * p.H * static_cast<double>(p.Sq) * p.Skv * 64, stream);
        LTXV_TRACE_VARIANT("cross_attn_kernel bias=%d", p.kv_bias != nullptr ? 1 : 0);
        cudaError_t le = launch_pdl(cross_attn_kernel, dim3(p.B * p.H * cph), dim3(kXThreads), kXSmemBytes, stream, tq, tk, tv, p, cph);
        if (le != cudaSuccess) return le;
    }
    g_attn_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

template <int D>
cudaError_t launch_attn_impl(const AttnParams& p, cudaStream_t stream) {
    using C = ACfg<D>;
    static PerDeviceOnce configured;
    int cfg_dev = 0;
    if (configured.need(&cfg_dev)) {
        cudaError_t e =
            cudaFuncSetAttribute(flash_attn_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
        if (e != cudaSuccess) return e;
        configured.mark(cfg_dev);
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
        LTXV_TRACE_VARIANT("flash_attn_kernel<%d>", D);
        cudaError_t le = launch_pdl(flash_attn_kernel<D>, grid, dim3(kAttnThreads), C::kSmemBytes, stream, tq, tk, tv, p);
        if (le != cudaSuccess) return le;
    }
    g_attn_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

}  // namespace

uint64_t attention_launch_count() { return g_attn_launches.load(); }
#ifdef LTXV_ATTN_TRACE
void attention_debug_trace(long long* out) { cudaMemcpyFromSymbol(out, g_attn_trace, sizeof(long long) * 11 * 40 * 8); }
#endif
#ifdef LTXV_ATTN_TIMING
void attention_debug_timing(long long* out32) { cudaMemcpyFromSymbol(out32, g_attn_timing, sizeof(long long) * 32); }
#endif

cudaError_t launch_attention(const AttnParams& p, cudaStream_t stream) {
    if (p.B <= 0 || p.H <= 0 || p.Sq <= 0 || p.Skv <= 0) return cudaErrorInvalidValue;
    if (p.D >= 8 && p.D <= 32 && p.D % 8 == 0) {
        // small head dims (the reference's golden test geometries): CUDA-core fallback
        if (p.out_rows_per_peer != 0 || p.ldq % 8 != 0 || p.ldk % 8 != 0 || p.ldv % 8 != 0 || p.ldo % 2 != 0 || p.q_col0 % 8 != 0 ||
            p.k_col0 % 8 != 0 || p.v_col0 % 8 != 0)
            return cudaErrorInvalidValue;
        ProfScope prof(p.kv_bias != nullptr || p.Skv != p.Sq ? PROF_ATTN_CROSS : PROF_ATTN_SELF,
                       4.0 * p.B * p.H * static_cast<double>(p.Sq) * p.Skv * p.D, stream);
        LTXV_TRACE_VARIANT("flash_attn_simt_kernel<32> D=%d", p.D);
        flash_attn_simt_kernel<32><<<dim3((p.Sq + 127) / 128, p.H, p.B), 128, 0, stream>>>(p);
        g_attn_launches.fetch_add(1, std::memory_order_relaxed);
        return cudaGetLastError();
    }
    if (p.D == 64 && p.kv_bias == nullptr && p.Skv > 2 * kTileKV && p.Sq > kTileQ && !options().attn_v1)
        return launch_attn3_impl(p, stream);
    if (p.D == 64 && p.Skv <= kTileKV && p.out_rows_per_peer == 0 && !options().attn_v1)
        return launch_cross_attn_impl(p, stream);  // one key tile: text cross-attention
    if (p.D == 128 && p.kv_bias == nullptr && p.Skv > 2 * kTileKV && p.Sq > kTileQ && !options().attn_v1)
        return launch_attn3_d128_impl(p, stream);
    if (p.D == 64) return launch_attn_impl<64>(p, stream);
    if (p.D == 128) return launch_attn_impl<128>(p, stream);
    return cudaErrorInvalidValue;
}

}  // namespace ltxv
