// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA (128B swizzle) -> smem ring -> tcgen05.mma (UMMA,
// accumulators in TMEM, double-buffered) -> tcgen05.ld epilogue with the fused element-wise tails of the
// LTX-Video DiT / VAE layers. Also runs CausalConv3d as an implicit GEMM (27 row-shifted A views).
//
// Warp roles (256 threads, 1 CTA / SM):
//   warp 0   : TMA producer (one elected lane)
//   warp 1   : UMMA issuer  (one elected lane)
//   warp 2   : TMEM allocator / deallocator
//   warp 3   : idle
//   warps 4-7: epilogue; warp w owns TMEM lanes 32*(w%4) .. +31, one accumulator row per thread
#include "common.cuh"
#include "gemm.h"
#define LTXV_PDL_CLASS 2
#include "launch.h"
#include "tensormap.h"
#include "options.h"
#include "profile.h"

#include <atomic>
#include <stdlib.h>

namespace ltxv {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kThreads = 256;
constexpr int kABytes = kBlockM * kBlockK * 2;
constexpr int kEpiRowFloats = 32;                              // unpadded rows; the float4 column is XOR-swizzled with (row & 7)
constexpr int kEpiStageBytes = 4 * 32 * kEpiRowFloats * 4;     // one 32x32 f32 transpose tile per epilogue warp

template <int BLOCK_N>
struct Cfg {
    static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (BLOCK_N >= 256) ? 4 : (BLOCK_N >= 192) ? 5 : (BLOCK_N >= 128) ? 6 : 8;
    static constexpr int kTmemStride = (BLOCK_N <= 64) ? 64 : (BLOCK_N <= 128) ? 128 : 256;
    static constexpr int kTmemCols = 2 * kTmemStride;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + kEpiStageBytes;
    static_assert(kBBytes % 1024 == 0, "B stage must keep 1024B alignment");
    static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 32 && BLOCK_N <= 256, "BLOCK_N");
};

struct Tile {
    int m0, n0;
};

// Tile raster.  Tiles are handed out in GROUPS of `g` row tiles x all column tiles (row tile fastest inside a group), so the
// ~148 tiles in flight at any moment cover a few row tiles of A against all of B: both stay L2 resident.  With the plain
// row-fastest order a wave touches EVERY row tile of A; when A does not fit in L2 (FFN-out at M = 9984: 163 MB) each of the
// 4.2 waves re-read it from HBM -- ncu: 800 MB of DRAM reads per launch against 197 MB of operands.  g = 0 keeps the old
// order.  The order changes which tile a CTA computes next, not any result bit.
__device__ __forceinline__ void tile_mn(int tile, int num_m, int num_n, int g, int& m, int& n) {
    if (g <= 0 || g >= num_m) {
        m = tile % num_m;
        n = tile / num_m;
        return;
    }
    const int per_group = g * num_n;
    const int gi = tile / per_group;
    const int m_base = gi * g;
    const int rows = min(g, num_m - m_base);
    const int local = tile - gi * per_group;
    m = m_base + local % rows;
    n = local / rows;
}
__device__ __forceinline__ Tile tile_coords(int tile, int num_m, int num_n, int g) {
    Tile t;
    int m, n;
    tile_mn(tile, num_m, num_n, g, m, n);
    t.m0 = m * kBlockM;
    t.n0 = n;
    return t;
}

// ------------------------------------------------------------------------------------------------
// Epilogue: one thread owns one accumulator row, 32 consecutive columns at a time.
// ------------------------------------------------------------------------------------------------
struct RowCtx {
    bool valid;       // row participates in stores
    int64_t out_row;  // row index in the output tensor (GEMM: m; conv: voxel index)
    int t, h, w;      // conv only
    int64_t a_row;    // conv only: row of the centre tap in the padded A volume
};

__device__ __forceinline__ RowCtx make_row_ctx(const GemmParams& p, int m) {
    RowCtx c;
    c.valid = m < p.M;
    c.out_row = m;
    c.t = c.h = c.w = 0;
    c.a_row = 0;
    if (p.conv) {
        const int wp = p.W + 2, plane = (p.H + 2) * wp;
        int t = m / plane;
        int r = m - t * plane;
        int hp = r / wp;
        int wq = r - hp * wp;
        c.t = t;
        c.h = hp - 1;
        c.w = wq - 1;
        c.valid = c.valid && hp >= 1 && hp <= p.H && wq >= 1 && wq <= p.W;
        c.out_row = (static_cast<int64_t>(t) * p.H + c.h) * p.W + c.w;
        c.a_row = static_cast<int64_t>(m) + plane;
    }
    return c;
}

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&v)[32]) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
        u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
        u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
        u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
        d4[i] = u;
    }
}

__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const RowCtx& rc, int col0, float (&v)[32]) {
    // bias (same 32 values for every thread of the warp: broadcast loads)
    if (p.bias != nullptr) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float4 b = __ldg(b4 + i);
            v[4 * i + 0] += b.x;
            v[4 * i + 1] += b.y;
            v[4 * i + 2] += b.z;
            v[4 * i + 3] += b.w;
        }
    }
    if (!rc.valid) return;

    switch (p.epi) {
        case EPI_STORE_BF16: {
            if (p.act == ACT_GELU_TANH) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = gelu_tanh_f32(v[i]);
            }
            store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.out) + rc.out_row * p.ldo + col0, v);
            break;
        }
        case EPI_STORE_F32: {
            float4* d4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + rc.out_row * p.ldo + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            break;
        }
        case EPI_RESIDUAL_F32: {
            float4* r4 = reinterpret_cast<float4*>(p.res_f32 + rc.out_row * p.ldo + col0);
            if (p.gate != nullptr) {
                const float4* g4 = reinterpret_cast<const float4*>(p.gate + col0);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4 g = __ldg(g4 + i);
                    v[4 * i + 0] *= g.x;
                    v[4 * i + 1] *= g.y;
                    v[4 * i + 2] *= g.z;
                    v[4 * i + 3] *= g.w;
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 r = r4[i];
                v[4 * i + 0] += r.x;
                v[4 * i + 1] += r.y;
                v[4 * i + 2] += r.z;
                v[4 * i + 3] += r.w;
                r4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
            if (p.out != nullptr)
                store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.out) + rc.out_row * p.ldo + col0, v);
            break;
        }
        case EPI_CONV_NDHWC: {
            if (p.res_bf16 != nullptr) {
                const uint4* r4 = reinterpret_cast<const uint4*>(
                    reinterpret_cast<const __nv_bfloat16*>(p.res_bf16) + rc.out_row * p.ldo + col0);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 u = __ldcs(r4 + i);  // streamed once: keep L2 for the A operand's (kt, kh) reuse
                    v[8 * i + 0] += bf16_lo(u.x);
                    v[8 * i + 1] += bf16_hi(u.x);
                    v[8 * i + 2] += bf16_lo(u.y);
                    v[8 * i + 3] += bf16_hi(u.y);
                    v[8 * i + 4] += bf16_lo(u.z);
                    v[8 * i + 5] += bf16_hi(u.z);
                    v[8 * i + 6] += bf16_lo(u.w);
                    v[8 * i + 7] += bf16_hi(u.w);
                }
            }
            store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.out) + rc.out_row * p.ldo + col0, v);
            break;
        }
        case EPI_CONV_D2S: {
            // Columns are stored sub-voxel major: col = sub * C' + c', sub = i*4 + j*2 + k (weights permuted at
            // load time), so a 32-column chunk lands in one output voxel. vae.rs:1142-1161.
            const int cprime = p.N >> 3;
            const int sub = col0 / cprime;
            const int c0 = col0 - sub * cprime;
            const int i = sub >> 2, j = (sub >> 1) & 1, k = sub & 1;
            const int to = 2 * rc.t + i - 1;  // drop frame 0 (vae.rs:1161)
            if (to < 0) break;
            // residual: x[(c' mod Cin/8) * 8 + sub] at the same voxel, channel-tiled (vae.rs:1101-1124)
            const int cr0 = c0 % (p.cin >> 3);
            const __nv_bfloat16* a =
                reinterpret_cast<const __nv_bfloat16*>(p.a_ptr) + rc.a_row * p.cin + cr0 * 8 + sub;
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] += __bfloat162float(a[q * 8]);
            const int Ho = 2 * p.H, Wo = 2 * p.W;
            const int64_t ov = (static_cast<int64_t>(to) * Ho + (2 * rc.h + j)) * Wo + (2 * rc.w + k);
            store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.out) + ov * cprime + c0, v);
            break;
        }
        case EPI_CONV_UNPATCHIFY: {
            // out[c3, f, 4h+j, 4w+i] = x[c3*16 + i*4 + j] (vae.rs:1626-1654); N = 48 (3 colour planes)
            float* out = reinterpret_cast<float*>(p.out);
            const int Ho = p.out_h_full > 0 ? p.out_h_full : 4 * p.H, Wo = 4 * p.W;
#pragma unroll
            for (int cl = 0; cl < 2; ++cl) {
                const int c3 = (col0 >> 4) + cl;
                if (c3 * 16 >= p.N) break;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4 o = make_float4(v[cl * 16 + 0 + j], v[cl * 16 + 4 + j], v[cl * 16 + 8 + j],
                                           v[cl * 16 + 12 + j]);
                    if (p.post_u8_scale) {
                        o.x = fminf(fmaxf(0.5f * o.x + 0.5f, 0.f), 1.f) * 255.f;
                        o.y = fminf(fmaxf(0.5f * o.y + 0.5f, 0.f), 1.f) * 255.f;
                        o.z = fminf(fmaxf(0.5f * o.z + 0.5f, 0.f), 1.f) * 255.f;
                        o.w = fminf(fmaxf(0.5f * o.w + 0.5f, 0.f), 1.f) * 255.f;
                    }
                    const int64_t idx =
                        ((static_cast<int64_t>(c3) * p.T + rc.t) * Ho + (p.out_h0 + 4 * rc.h + j)) * Wo + 4 * rc.w;
                    *reinterpret_cast<float4*>(out + idx) = o;
                }
            }
            break;
        }
        default: break;
    }
}

// ------------------------------------------------------------------------------------------------
// Coalesced epilogue for the plain (non-conv) GEMMs.  tcgen05.ld hands every thread ONE accumulator row, so direct
// stores touch 32 different rows per warp instruction (32 LSU wavefronts, 16 useful bytes each): at K = 2048 the
// f32 residual epilogue (read x, write x, write the bf16 copy) then costs more than the whole main loop
// (measured 634 TFLOP/s for the [4992,2048]x[2048,2048] projections vs 1.0-1.1 PFLOP/s for the others).  Here each
// 32x32 chunk goes through a per-warp shared-memory tile and comes back transposed: a lane holds 4 consecutive
// columns of rows (4i + lane/8), so one warp instruction covers 4 rows x 128 contiguous bytes.
// ------------------------------------------------------------------------------------------------
// residual rows of one 32x32 chunk in the transposed (coalesced) ownership; issued one chunk AHEAD of its use so the
// global round trip overlaps the TMEM load / transpose / stores of the previous chunk
__device__ __forceinline__ void epilogue_load_residual(const GemmParams& p, int m_warp0, int col0, int lane,
                                                       float4 (&res)[8]) {
    const int col = col0 + (lane & 7) * 4;
    const int rsub = lane >> 3;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m_warp0 + 4 * i + rsub;
        res[i] = (m < p.M && col < p.N) ? *reinterpret_cast<const float4*>(p.res_f32 + m * p.ldo + col)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// RoPE table rows of the 8 output rows a lane handles in epilogue_chunk_coalesced<true> (rows 4i + lane / 8 of the warp's
// 32): the table has rope_rows rows (the batched CFG pair has M = 2 x rope_rows), 32-bit arithmetic, once per tile
__device__ __forceinline__ void epilogue_rope_rows(const GemmParams& p, int m_warp0, int lane, int (&trow)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int mm = m_warp0 + 4 * i + (lane >> 3);
        if (mm >= p.M) mm = p.M - 1;
        trow[i] = static_cast<int>(static_cast<unsigned>(mm) % static_cast<unsigned>(p.rope_rows)) + p.rope_row0;
    }
}

template <bool QK>
__device__ __forceinline__ void epilogue_chunk_coalesced(const GemmParams& p, int m_warp0, int col0, int lane,
                                                         float (&v)[32], float* stage, const float4 (&res)[8],
                                                         float (&ss)[8], const int (&trow)[8]) {
    // 32x32 f32 transpose tile: row r keeps its float4 column c at slot c ^ (r & 7).  Writes (one row per lane) and reads
    // (4 rows x 8 float4 per instruction) are both 4 wavefronts per 512 B = conflict-free without padding, which leaves
    // room for one more ring stage in the two-epilogue-group pair kernels (18 -> 16 KB per group)
    float4* srow = reinterpret_cast<float4*>(stage + lane * kEpiRowFloats);
#pragma unroll
    for (int i = 0; i < 8; ++i) srow[i ^ (lane & 7)] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    __syncwarp();
    const int cq = (lane & 7) * 4;
    const int rsub = lane >> 3;
    const int col = col0 + cq;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f), g = make_float4(1.f, 1.f, 1.f, 1.f);
    if (p.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
    if (p.epi == EPI_RESIDUAL_F32 && p.gate != nullptr) g = __ldg(reinterpret_cast<const float4*>(p.gate + col));
    // all loads first (the residual rows arrive prefetched in `res`): the stores below go to rows whose addresses the
    // compiler cannot disambiguate from the next row's loads, so interleaving them would serialise 8 global round trips
    float4 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        x[i] = *reinterpret_cast<const float4*>(stage + (4 * i + rsub) * kEpiRowFloats + (((lane & 7) ^ ((4 * i + rsub) & 7)) << 2));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m_warp0 + 4 * i + rsub;
        float4 v4 = x[i];
        v4.x += b.x;
        v4.y += b.y;
        v4.z += b.z;
        v4.w += b.w;
        if (m >= p.M) continue;
        if (QK) {
            continue;  // handled below (all table loads of the chunk are issued before the first use)
        } else if (p.epi == EPI_STORE_BF16) {
            if (p.act == ACT_GELU_TANH) {
                v4.x = gelu_tanh_f32(v4.x);
                v4.y = gelu_tanh_f32(v4.y);
                v4.z = gelu_tanh_f32(v4.z);
                v4.w = gelu_tanh_f32(v4.w);
            }
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldo + col) =
                make_uint2(pack_bf16x2(v4.x, v4.y), pack_bf16x2(v4.z, v4.w));
        } else if (p.epi == EPI_STORE_F32) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + m * p.ldo + col) = v4;
        } else {  // EPI_RESIDUAL_F32: gate * y, then + x: two roundings like the reference's ops
            v4.x = __fadd_rn(__fmul_rn(v4.x, g.x), res[i].x);
            v4.y = __fadd_rn(__fmul_rn(v4.y, g.y), res[i].y);
            v4.z = __fadd_rn(__fmul_rn(v4.z, g.z), res[i].z);
            v4.w = __fadd_rn(__fmul_rn(v4.w, g.w), res[i].w);
            *reinterpret_cast<float4*>(p.res_f32 + m * p.ldo + col) = v4;
            if (p.out != nullptr)
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldo + col) =
                    make_uint2(pack_bf16x2(v4.x, v4.y), pack_bf16x2(v4.z, v4.w));
        }
    }
    if constexpr (QK) {
        // EPI_QKV_ROPE (see gemm.h).  x[] holds acc (bias not yet added); rows 4i + rsub, columns col .. col + 3.
        const bool treat = col < p.qk_cols;
        const int which = col >= p.qk_dim ? 1 : 0;
        const int cc = col - which * p.qk_dim;
        float4 w4 = make_float4(1.f, 1.f, 1.f, 1.f);
        float2 cs[8], sn[8];
        const bool rot = treat && p.rope_cos != nullptr;
        if (treat) w4 = __ldg(reinterpret_cast<const float4*>(p.qk_w[which] + cc));
        if (rot) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                // trow[i] = RoPE table row of output row 4i + rsub, computed once per tile by the caller: the
                // 64-bit `m % rope_rows` that used to sit here ran 64 times per lane and tile (8 rows x 8 chunks)
                const int64_t tr = static_cast<int64_t>(trow[i]) * (p.qk_dim >> 1) + (cc >> 1);
                cs[i] = __ldg(reinterpret_cast<const float2*>(p.rope_cos + tr));
                sn[i] = __ldg(reinterpret_cast<const float2*>(p.rope_sin + tr));
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t m = m_warp0 + 4 * i + rsub;
            float4 v4 = x[i];
            v4.x += b.x;
            v4.y += b.y;
            v4.z += b.z;
            v4.w += b.w;
            if (m >= p.M) continue;
            if (treat) {
                ss[i] += (v4.x * v4.x + v4.y * v4.y) + (v4.z * v4.z + v4.w * v4.w);
                v4.x *= w4.x;
                v4.y *= w4.y;
                v4.z *= w4.z;
                v4.w *= w4.w;
                if (rot) {
                    // interleaved pairs: rot(x)[2i] = -x[2i+1], rot(x)[2i+1] = x[2i]; f32 math (ltx_transformer.rs:314-339)
                    const float x0 = v4.x, x1 = v4.y, x2 = v4.z, x3 = v4.w;
                    v4.x = x0 * cs[i].x - x1 * sn[i].x;
                    v4.y = x1 * cs[i].x + x0 * sn[i].x;
                    v4.z = x2 * cs[i].y - x3 * sn[i].y;
                    v4.w = x3 * cs[i].y + x2 * sn[i].y;
                }
            }
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldo + col) =
                make_uint2(pack_bf16x2(v4.x, v4.y), pack_bf16x2(v4.z, v4.w));
        }
    }
    __syncwarp();  // the tile is rewritten by the next chunk
}

// EPI_QKV_ROPE: per-row sums of squares of ONE 64-column group (two 32-column chunks; a lane owns 4 columns of rows
// 4i + lane/8): add the 8 lanes of a row group and store the total at group (n0 / 64).  The grouping is by absolute
// column, not by tile, so the bits of a row's sums do not depend on the tile width the heuristic picked (single GPU and
// token shards pick different widths and must still agree bit for bit).
__device__ __forceinline__ void epilogue_store_row_ss(const GemmParams& p, int m_warp0, int n0, int lane, float (&ss)[8]) {
    const int groups_total = p.qk_cols >> 6;
    const int g0 = n0 >> 6, rsub = lane >> 3;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float t = ss[i];
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        t += __shfl_xor_sync(0xffffffffu, t, 4);
        const int64_t m = m_warp0 + 4 * i + rsub;
        if ((lane & 7) == 0 && m < p.M && n0 < p.qk_cols) p.qk_ss[m * groups_total + g0] = t;
        ss[i] = 0.f;
    }
}

// ------------------------------------------------------------------------------------------------
// EPI_CONV_NORM_PAD: the conv epilogue is also the producer of the next conv's input.  One thread owns one voxel and the
// tile holds all N <= BN channels, so the row is rounded to bf16 once (the value `x` every consumer sees), kept in
// registers as BN/2 packed words while the RMS accumulates, then normalised / modulated / SiLU'd and stored at the
// voxel's place in the padded volume (plus the replicated frames and the neighbours' halo rows).  Same arithmetic as
// vae_prep_kernel on the stored x; removes that kernel's read of x and, for a resnet's first conv, the write of x too.
// ------------------------------------------------------------------------------------------------
// The residual row of a fused-producer tile is requested into L2 while the tile's main loop still runs (the epilogue
// warps are idle then): its four dependent 64-byte loads per row otherwise pay the HBM round trip one after the other
// and, together with the second pass, outlast the main loop of the 128-channel convs (measured: 1.34 -> 1.99 ms).
__device__ __forceinline__ void prefetch_residual_row_l2(const GemmParams& p, const RowCtx& rc) {
    if (p.epi != EPI_CONV_NORM_PAD || p.res_bf16 == nullptr || !rc.valid) return;
    const char* r = reinterpret_cast<const char*>(reinterpret_cast<const __nv_bfloat16*>(p.res_bf16) + rc.out_row * p.ldo);
    for (int b = 0; b < p.N * 2; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(r + b));
}

// Two passes over the accumulator row in TMEM, 32 columns at a time (loops kept rolled: a fully unrolled version is
// ~200 KB of code and thrashes the instruction cache of the whole CTA): pass 1 accumulates the RMS of the bf16-rounded
// row, pass 2 recomputes the same rounded values, stores x and the normalised / modulated / SiLU'd padded copy.
template <int BN>
__device__ __noinline__ void epilogue_conv_norm_pad(const GemmParams& p, const RowCtx rc, uint32_t taddr,
                                                    uint64_t* tmem_empty, int arrive_on_leader, int lane) {
    const int N = p.N;
    const float* bias = p.bias;
    const __nv_bfloat16* res =
        (p.res_bf16 != nullptr && rc.valid) ? reinterpret_cast<const __nv_bfloat16*>(p.res_bf16) + rc.out_row * p.ldo : nullptr;
    // x (bf16-rounded acc + bias + residual) of one 32-column chunk as 16 packed words
    auto load_chunk = [&](int c, uint32_t (&w)[16]) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + c * 32, r);
        tmem_ld_wait();
        float v[32];
        const float4* b4 = reinterpret_cast<const float4*>(bias + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 b = __ldg(b4 + i);
            v[4 * i + 0] = __uint_as_float(r[4 * i + 0]) + b.x;
            v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + b.y;
            v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + b.z;
            v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + b.w;
        }
        if (res != nullptr) {
            const uint4* r4 = reinterpret_cast<const uint4*>(res + c * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint4 u = __ldcs(r4 + i);  // read once: streaming, evict-first
                v[8 * i + 0] += bf16_lo(u.x);
                v[8 * i + 1] += bf16_hi(u.x);
                v[8 * i + 2] += bf16_lo(u.y);
                v[8 * i + 3] += bf16_hi(u.y);
                v[8 * i + 4] += bf16_lo(u.z);
                v[8 * i + 5] += bf16_hi(u.z);
                v[8 * i + 6] += bf16_lo(u.w);
                v[8 * i + 7] += bf16_hi(u.w);
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    };
    const int n_chunks = (N < BN ? N : BN) / 32;
    __nv_bfloat16* xdst =
        (p.out != nullptr && rc.valid) ? reinterpret_cast<__nv_bfloat16*>(p.out) + rc.out_row * p.ldo : nullptr;
    // When x itself is stored (a resnet's second conv) pass 1 stores it and hands the accumulator back at once; pass 2
    // then re-reads the thread's own 64-byte pieces of x (L1/L2 hits) instead of TMEM + bias + residual.
    const bool reread = p.out != nullptr;
    auto release_tmem = [&] {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
            if (arrive_on_leader) mbar_arrive_leader(tmem_empty);
            else mbar_arrive(tmem_empty);
        }
    };
    float s2 = 0.f;
    if (p.norm_do || reread) {
#pragma unroll 1
        for (int c = 0; c < n_chunks; ++c) {
            uint32_t w[16];
            load_chunk(c, w);
            if (reread && c == n_chunks - 1) release_tmem();
            if (xdst != nullptr) {
                uint4* d4 = reinterpret_cast<uint4*>(xdst + c * 32);
#pragma unroll
                for (int q = 0; q < 4; ++q) __stcs(d4 + q, make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]));
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float a = bf16_lo(w[i]), b = bf16_hi(w[i]);
                s2 = fmaf(a, a, s2);
                s2 = fmaf(b, b, s2);
            }
        }
    }
    if (reread && !rc.valid) return;
    const float rinv = p.norm_do ? rsqrtf(s2 / static_cast<float>(N) + 1e-8f) : 1.0f;
    const int Hp = p.H + 2, Wp = p.W + 2, tf = p.norm_tf;
    const int64_t plane_elems = static_cast<int64_t>(Hp) * Wp * N;
    const bool mod = p.norm_scale != nullptr;
    const int silu = p.norm_silu;
    // destinations of this voxel: its place in the padded volume, the replicated frames when it lies in the first /
    // last frame, and the neighbours' halo rows when it lies in the first / last row of an H-slab
    const int64_t row0 = (static_cast<int64_t>(rc.t + tf) * Hp) * Wp + (rc.w + 1);
    __nv_bfloat16* d_main = reinterpret_cast<__nv_bfloat16*>(p.norm_out) + (row0 + static_cast<int64_t>(rc.h + 1) * Wp) * N;
    // the neighbours' padded volumes have their own plane stride when the slabs are ragged
    const int Hu = p.norm_halo_up_h > 0 ? p.norm_halo_up_h : p.H, Hd = p.norm_halo_dn_h > 0 ? p.norm_halo_dn_h : p.H;
    const int64_t plane_up = static_cast<int64_t>(Hu + 2) * Wp * N, plane_dn = static_cast<int64_t>(Hd + 2) * Wp * N;
    __nv_bfloat16* d_up = (p.norm_halo_up != nullptr && rc.h == 0)
                              ? reinterpret_cast<__nv_bfloat16*>(p.norm_halo_up) + (rc.t + tf) * plane_up +
                                    (static_cast<int64_t>(Hu + 1) * Wp + (rc.w + 1)) * N
                              : nullptr;
    __nv_bfloat16* d_dn = (p.norm_halo_dn != nullptr && rc.h == p.H - 1)
                              ? reinterpret_cast<__nv_bfloat16*>(p.norm_halo_dn) + (rc.t + tf) * plane_dn +
                                    static_cast<int64_t>(rc.w + 1) * N
                              : nullptr;
    const int n_front = rc.t == 0 ? tf : 0;
    const bool back = tf == 1 && rc.t == p.T - 1;
    const bool simple = rc.valid && n_front == 0 && !back && d_up == nullptr && d_dn == nullptr;
#pragma unroll 1
    for (int c = 0; c < n_chunks; ++c) {
        uint32_t w[16];
        if (reread) {
            const uint4* x4 = reinterpret_cast<const uint4*>(xdst + c * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 u = __ldcs(x4 + q);
                w[4 * q] = u.x; w[4 * q + 1] = u.y; w[4 * q + 2] = u.z; w[4 * q + 3] = u.w;
            }
        } else {
            load_chunk(c, w);
            if (c == n_chunks - 1) release_tmem();  // accumulator drained for the second time
            if (!rc.valid) continue;
        }
        uint32_t o[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                f[2 * i] = bf16_lo(w[4 * q + i]) * rinv;
                f[2 * i + 1] = bf16_hi(w[4 * q + i]) * rinv;
            }
            if (mod) {
                const float4* sc4 = reinterpret_cast<const float4*>(p.norm_scale + c * 32 + q * 8);
                const float4* sh4 = reinterpret_cast<const float4*>(p.norm_shift + c * 32 + q * 8);
                const float4 sa = __ldg(sc4), sb = __ldg(sc4 + 1), ta = __ldg(sh4), tb = __ldg(sh4 + 1);
                f[0] = f[0] * (1.f + sa.x) + ta.x; f[1] = f[1] * (1.f + sa.y) + ta.y;
                f[2] = f[2] * (1.f + sa.z) + ta.z; f[3] = f[3] * (1.f + sa.w) + ta.w;
                f[4] = f[4] * (1.f + sb.x) + tb.x; f[5] = f[5] * (1.f + sb.y) + tb.y;
                f[6] = f[6] * (1.f + sb.z) + tb.z; f[7] = f[7] * (1.f + sb.w) + tb.w;
            }
            if (silu) {
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = silu_fast_f32(f[i]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) o[4 * q + i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
        }
        auto put = [&](__nv_bfloat16* d) {
            uint4* d4 = reinterpret_cast<uint4*>(d + c * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) __stcs(d4 + q, make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));
        };
        put(d_main);
        if (!simple) {
#pragma unroll 1
            for (int which = 0; which < 3; ++which) {
                __nv_bfloat16* d = which == 0 ? d_main : (which == 1 ? d_up : d_dn);
                if (d == nullptr) continue;
                if (which != 0) put(d);
                const int64_t pl = which == 0 ? plane_elems : (which == 1 ? plane_up : plane_dn);
#pragma unroll 1
                for (int k = 1; k <= n_front; ++k) put(d - k * pl);
                if (back) put(d + pl);
            }
        }
    }
}

// conv3d k-block order: (kt, kh) outermost, then the 64-channel block, then kw -- the SAME accumulation order as the
// KW3 pair kernel (three kw taps per pipeline step), so every kernel variant produces identical bits and the result
// does not depend on which one the tile heuristic picks (single GPU vs H-slabs pick differently).
__device__ __forceinline__ void conv_kblock(const GemmParams& p, int kb, int& a_row, int& a_col, int& b_col) {
    const int per_th = 3 * p.cin_blocks;
    const int th = kb / per_th;  // kt * 3 + kh
    const int r = kb - th * per_th;
    const int cb = r / 3, kw = r - cb * 3;
    const int tap = th * 3 + kw;
    a_row += p.tap_off[tap];
    a_col = cb * kBlockK;
    b_col = (tap * p.cin_blocks + cb) * kBlockK;
}

// ------------------------------------------------------------------------------------------------
// Kernel
// ------------------------------------------------------------------------------------------------
template <int BLOCK_N, bool QK = false>  // QK: the EPI_QKV_ROPE epilogue (its own instances: the others carry none of its code)
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ GemmParams p) {
    using C = Cfg<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzle atoms need 1024 B alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + C::kStages * kABytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + C::kStages;
    uint64_t* tmem_full_bar = bars + 2 * C::kStages;
    uint64_t* tmem_empty_bar = bars + 2 * C::kStages + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 4);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int num_m = (p.M + kBlockM - 1) / kBlockM;
    const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
    const int num_tiles = num_m * num_n;
    const int num_kb = p.num_k_blocks;

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp_idx == 1 && lane == 0) {
        for (int i = 0; i < C::kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], 4);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp_idx == 2) {
        tmem_alloc<C::kTmemCols>(tmem_slot);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_launch_dependents();
    griddep_wait();  // A (and the residual / conv inputs) may come from the previous kernel of the stream

    if (warp_idx == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {  // one elected lane: ptxas keeps the tcgen05/TMA operands on the uniform datapath
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const Tile t = tile_coords(tile, num_m, num_n, p.raster_g);
                const int n0 = t.n0 * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait_sleep(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
                    int a_row = t.m0, a_col = kb * kBlockK, b_col = kb * kBlockK;
                    if (p.conv) conv_kblock(p, kb, a_row, a_col, b_col);
                    tma_load_2d(smem_a + stage * kABytes, &tmap_a, &full_bar[stage], a_col, a_row);
                    tma_load_2d(smem_b + stage * C::kBBytes, &tmap_b, &full_bar[stage], b_col, n0);
                    if (++stage == C::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================== UMMA issuer =====================
        if (elect_one()) {  // one elected lane: ptxas keeps the tcgen05/TMA operands on the uniform datapath
            constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, false, false);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait_sleep(&tmem_empty_bar[acc], acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * C::kTmemStride;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait_sleep(&full_bar[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t a_addr = smem_u32(smem_a + stage * kABytes);
                    const uint32_t b_addr = smem_u32(smem_b + stage * C::kBBytes);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                        const uint64_t da = make_smem_desc_sw128(a_addr + k * kUmmaK * 2, 1024, 0);
                        const uint64_t db = make_smem_desc_sw128(b_addr + k * kUmmaK * 2, 1024, 0);
                        umma_bf16_ss(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
                    if (kb == num_kb - 1) umma_commit(&tmem_full_bar[acc]);
                    if (++stage == C::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else if (warp_idx >= 4) {
        // ===================== epilogue =====================
        const int quad = warp_idx & 3;
        float* stage = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes + 256) + quad * (32 * kEpiRowFloats);
        const bool coalesced = !p.conv && (p.epi == EPI_STORE_BF16 || p.epi == EPI_STORE_F32 || p.epi == EPI_RESIDUAL_F32 || QK) &&
                               (p.N % 32 == 0);
        int acc = 0;
        uint32_t acc_phase = 0;
        float row_ss[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // EPI_QKV_ROPE: sums of squares of this tile's rows
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const Tile t = tile_coords(tile, num_m, num_n, p.raster_g);
            const int n0 = t.n0 * BLOCK_N;
            const int m = t.m0 + quad * 32 + lane;
            const RowCtx rc = make_row_ctx(p, m);
            float4 res_a[8], res_b[8];  // residual rows, double-buffered one chunk ahead (EPI_RESIDUAL_F32)
            const bool prefetch_res = coalesced && p.epi == EPI_RESIDUAL_F32;
            const int m_warp0 = t.m0 + quad * 32;
            int trow[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if constexpr (QK) {
                if (p.rope_cos != nullptr) epilogue_rope_rows(p, m_warp0, lane, trow);
            }
            if (prefetch_res) epilogue_load_residual(p, m_warp0, n0, lane, res_a);  // before the MMAs finish
            prefetch_residual_row_l2(p, rc);
            mbar_wait_sleep(&tmem_full_bar[acc], acc_phase);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * C::kTmemStride;
            auto chunk = [&](int c, const float4 (&res)[8]) {
                const int col0 = n0 + c * 32;
                uint32_t r[32];
                tmem_ld_32x32b_x32(taddr + c * 32, r);
                tmem_ld_wait();
                if (c == BLOCK_N / 32 - 1) {
                    // accumulator fully drained into registers: hand the TMEM stage back to the MMA warp
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
                }
                if (col0 < p.N) {
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                    if (coalesced) epilogue_chunk_coalesced<QK>(p, m_warp0, col0, lane, v, stage, res, row_ss, trow);
                    else epilogue_chunk(p, rc, col0, v);
                }
            };
            if (p.epi == EPI_CONV_NORM_PAD) {
                if constexpr (BLOCK_N == 128 || BLOCK_N == 256) epilogue_conv_norm_pad<BLOCK_N>(p, rc, taddr, &tmem_empty_bar[acc], 0, lane);
            } else {
#pragma unroll 1
                for (int c = 0; c < BLOCK_N / 32; c += 2) {
                    if (prefetch_res) epilogue_load_residual(p, m_warp0, n0 + (c + 1) * 32, lane, res_b);
                    chunk(c, res_a);
                    if (prefetch_res && c + 2 < BLOCK_N / 32) epilogue_load_residual(p, m_warp0, n0 + (c + 2) * 32, lane, res_a);
                    chunk(c + 1, res_b);
                    if constexpr (QK) epilogue_store_row_ss(p, m_warp0, n0 + c * 32, lane, row_ss);
                }
            }
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp_idx == 2) {
        tcgen05_fence_after();
        tmem_dealloc<C::kTmemCols>(tmem_base);
    }
}


// ================================================================================================
// CTA-pair variant: one 256 x 256 output tile per cluster of two CTAs (cta_group::2).
// The single-CTA kernel above moves 48 KB of operands through shared memory per 128x256x64 step twice (TMA write,
// UMMA read): ~190 B/clk against ~128 B/clk of shared-memory bandwidth, i.e. it is shared-memory bound at ~2/3 of the
// tensor peak (ncu: tensor pipe 63-69 % active).  In a pair each CTA stages 128 rows of A and HALF of the 256 B rows
// (32 KB per step) and the pair's tensor cores share the B halves: per-CTA traffic drops by a third and six stages fit.
// Roles per CTA as above; only the leader CTA's warp 1 issues MMAs; both CTAs run TMA producers and epilogues.
// ================================================================================================
//
// KW3 (conv3d only): one pipeline step covers the THREE kw taps of a (kt, kh, channel-block).  The taps read the same
// rows of the padded volume shifted by -1 / 0 / +1, so the step loads ONE box of 128 + 8 rows and the three MMAs start
// their A descriptor 0 / 1 / 2 rows (128 B each) into it; the swizzle phase follows the absolute smem address.  The A
// traffic from L2 drops 3x -- at Cin = Cout = 128 the conv re-read 1.7 MB of operands per 128x128 tile and was
// L2-bandwidth bound at ~1.0 PFLOP/s.
constexpr int kBoxRows3 = kBlockM + 8;                 // rows per KW3 A box (needs 130, keeps the 8-row swizzle atom)
constexpr int kABytes3 = 18 * 1024;                    // 136 x 128 B = 17 408, padded to a 1024-byte multiple
constexpr int kABoxBytes3 = kBoxRows3 * kBlockK * 2;   // bytes one KW3 A box transfers

// MODE 0: one 64-wide k-block per pipeline stage; MODE 1: KW3 (above); MODE 2: TWO k-blocks per stage (8 MMAs per
// full/empty barrier round trip instead of 4 -- the plain GEMMs ran at 71 % tensor-pipe activity against 99 % for KW3).
// TWO: two epilogue warpgroups like the conv variants, each with its own (swizzled, unpadded) transpose tiles; same ring depth as the one-group kernels.  Used
// where the epilogue outlasts a K = 2048 main loop with one group: EPI_QKV_ROPE (norm weight, RoPE table reads, sums of
// squares) and the f32 residual read-modify-write of the attention out-projections.
template <int BN, int MODE, bool TWO = false>  // N of the cluster tile (256 or 128); each CTA stages BN / 2 rows of B
struct PairCfg {
    static constexpr bool KW3 = MODE == 1;
    static constexpr int kKbPerStage = MODE == 2 ? 2 : 1;
    static constexpr int kBBytes = (BN / 2) * kBlockK * 2;          // one k-block's / tap's B half: 16 KB / 8 KB
    static constexpr int kAStage = KW3 ? kABytes3 : kKbPerStage * kABytes;
    static constexpr int kBStage = KW3 ? 3 * kBBytes : kKbPerStage * kBBytes;
    static constexpr int kStageBytes = kAStage + kBStage;
    static constexpr int kTxBytes = (KW3 ? kABoxBytes3 : kKbPerStage * kABytes) + kBStage;  // bytes ONE CTA credits per stage
#ifndef LTXV_PAIR_STAGE_EXP
#define LTXV_PAIR_STAGE_EXP 0  // timing experiment: ring stages removed from the plain pair kernels
#endif
    static constexpr int kStages =
        KW3 ? (BN == 256 ? 3 : 4) : ((BN == 256 ? 6 : 8) - LTXV_PAIR_STAGE_EXP) / kKbPerStage;
    static constexpr int kTmemCols = 2 * BN;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + (TWO ? 2 : 1) * kEpiStageBytes;
    static_assert(BN == 256 || BN == 128, "pair tile width");
    static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

// The conv variants (MODE 1) run TWO epilogue warpgroups, one per TMEM accumulator stage, on alternating tiles: with one
// warp per scheduler the epilogue is latency-bound (a 128x128 tile with residual, x store and the fused producer takes
// ~16 us against an 11 us main loop at C = 128), and two groups give every tile two main loops of time.
template <int MODE, bool TWO = false>
struct PairThreads {
    static constexpr int value = (MODE == 1 || TWO) ? 384 : kThreads;
};
// E2: 0 = one epilogue warpgroup; 1 = two + the EPI_QKV_ROPE epilogue code; 2 = two, plain epilogues
template <int BN, int MODE, int E2 = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PairThreads<MODE, E2 != 0>::value, 1)
gemm_pair_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ GemmParams p) {
    constexpr bool QK = E2 == 1;
    constexpr bool TWO = E2 != 0;
    using PC = PairCfg<BN, MODE, TWO>;
    constexpr bool KW3 = PC::KW3;
    constexpr int kKbPerStage = PC::kKbPerStage;
    constexpr int kPairBlockN = BN;
    constexpr int kPairBBytes = PC::kBBytes;
    constexpr int kPairStageBytes = PC::kStageBytes;
    constexpr int kPairStages = PC::kStages;
    constexpr int kAStage = PC::kAStage;
    constexpr int kBStage = PC::kBStage;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kPairStages * kAStage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPairStages * kPairStageBytes);
    uint64_t* full_bar = bars;                       // leader's copy is the live one (tx from both CTAs)
    uint64_t* empty_bar = bars + kPairStages;        // both copies live (multicast commit)
    uint64_t* tmem_full_bar = bars + 2 * kPairStages;      // [2] both copies live (multicast commit)
    uint64_t* tmem_empty_bar = bars + 2 * kPairStages + 2;  // [2] leader's copy: 8 arrivals (4 epilogue warps x 2 CTAs)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kPairStages + 4);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;

    const int num_m = (p.M + 2 * kBlockM - 1) / (2 * kBlockM);  // 256-row tiles
    const int num_n = (p.N + kPairBlockN - 1) / kPairBlockN;
    const int num_tiles = num_m * num_n;
    const int num_kb = KW3 ? p.num_k_blocks / 3 : p.num_k_blocks / kKbPerStage;  // pipeline steps per tile

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp_idx == 1 && lane == 0) {
        for (int i = 0; i < kPairStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], 8);
        }
        fence_barrier_init();
    }
    if (warp_idx == 2) tmem_alloc_2sm<PC::kTmemCols>(tmem_slot);
    tcgen05_fence_before();
    cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_launch_dependents();
    griddep_wait();

    if (warp_idx == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                int tm, tn;
                tile_mn(tile, num_m, num_n, p.raster_g, tm, tn);
                const int m0 = tm * (2 * kBlockM) + static_cast<int>(rank) * kBlockM;
                const int n0 = tn * kPairBlockN + static_cast<int>(rank) * (kPairBlockN / 2);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait_sleep(&empty_bar[stage], phase ^ 1);
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * PC::kTxBytes);  // bytes of BOTH CTAs
                    if (KW3) {
                        const int th = kb / p.cin_blocks;  // kt * 3 + kh
                        const int cb = kb - th * p.cin_blocks;
                        // rows of tap kw = 0 (shift -1); kw = 1, 2 start one / two rows further into the same box
                        tma_load_2d_2sm(smem_a + stage * kAStage, &tmap_a, &full_bar[stage], cb * kBlockK,
                                        m0 + p.tap_off[th * 3]);
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw)
                            tma_load_2d_2sm(smem_b + stage * kBStage + kw * kPairBBytes, &tmap_b, &full_bar[stage],
                                            ((th * 3 + kw) * p.cin_blocks + cb) * kBlockK, n0);
                    } else {
#pragma unroll
                        for (int j = 0; j < kKbPerStage; ++j) {
                            const int kbj = kb * kKbPerStage + j;
                            int a_row = m0, a_col = kbj * kBlockK, b_col = kbj * kBlockK;
                            if (p.conv) conv_kblock(p, kbj, a_row, a_col, b_col);
                            tma_load_2d_2sm(smem_a + stage * kAStage + j * kABytes, &tmap_a, &full_bar[stage], a_col, a_row);
                            tma_load_2d_2sm(smem_b + stage * kBStage + j * kPairBBytes, &tmap_b, &full_bar[stage], b_col, n0);
                        }
                    }
                    if (++stage == kPairStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================== UMMA issuer (leader CTA only) =====================
        if (leader && elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(2 * kBlockM, kPairBlockN, false, false);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                mbar_wait_sleep(&tmem_empty_bar[acc], acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * kPairBlockN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait_sleep(&full_bar[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t a_addr = smem_u32(smem_a + stage * kAStage);
                    const uint32_t b_addr = smem_u32(smem_b + stage * kBStage);
#pragma unroll
                    for (int kw = 0; kw < (KW3 ? 3 : kKbPerStage); ++kw) {
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                            // KW3: tap kw reads the SAME box from row kw on (one row = 128 B); otherwise sub-block kw
                            const uint32_t a_off = KW3 ? kw * 128 : kw * kABytes;
                            const uint64_t da = make_smem_desc_sw128(a_addr + a_off + k * kUmmaK * 2, 1024, 0);
                            const uint64_t db = make_smem_desc_sw128(b_addr + kw * kPairBBytes + k * kUmmaK * 2, 1024, 0);
                            umma_bf16_ss_2sm(tmem_d, da, db, idesc, (kb | kw | k) != 0 ? 1u : 0u);
                        }
                    }
                    umma_commit_2sm(&empty_bar[stage]);  // frees the stage in BOTH CTAs once these MMAs retire
                    if (kb == num_kb - 1) umma_commit_2sm(&tmem_full_bar[acc]);
                    if (++stage == kPairStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else if (warp_idx >= 4) {
        // ===================== epilogue (both CTAs: this CTA's 128 rows of the tile) =====================
        const int quad = warp_idx & 3;
        // one transpose tile per epilogue WARP (two warpgroups in the QK instances)
        float* stage_buf = reinterpret_cast<float*>(smem + kPairStages * kPairStageBytes + 256) +
                           (TWO ? warp_idx - 4 : quad) * (32 * kEpiRowFloats);
        const bool coalesced = !p.conv && (p.epi == EPI_STORE_BF16 || p.epi == EPI_STORE_F32 || p.epi == EPI_RESIDUAL_F32 || QK) &&
                               (p.N % 32 == 0);
        int acc = 0;
        uint32_t acc_phase = 0;
        constexpr int kEpiGroups = PairThreads<MODE, TWO>::value == 384 ? 2 : 1;
        const int epi_group = (warp_idx - 4) >> 2;
        int it = 0;
        float row_ss[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // EPI_QKV_ROPE: sums of squares of this tile's rows
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
            if (kEpiGroups == 2) {  // group g owns accumulator stage g: tiles g, g+2, g+4, ... of this cluster
                if ((it & 1) != epi_group) continue;
                acc = epi_group;
                acc_phase = (it >> 1) & 1;
            }
            int tm, tn;
            tile_mn(tile, num_m, num_n, p.raster_g, tm, tn);
            const int m_cta0 = tm * (2 * kBlockM) + static_cast<int>(rank) * kBlockM;
            const int n0 = tn * kPairBlockN;
            const int m_warp0 = m_cta0 + quad * 32;
            const RowCtx rc = make_row_ctx(p, m_warp0 + lane);
            float4 res_a[8], res_b[8];
            const bool prefetch_res = coalesced && p.epi == EPI_RESIDUAL_F32;
            int trow[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if constexpr (QK) {
                if (p.rope_cos != nullptr) epilogue_rope_rows(p, m_warp0, lane, trow);
            }
            if (prefetch_res) epilogue_load_residual(p, m_warp0, n0, lane, res_a);
            prefetch_residual_row_l2(p, rc);
            mbar_wait_sleep(&tmem_full_bar[acc], acc_phase);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kPairBlockN;
            auto chunk = [&](int c, const float4 (&res)[8]) {
                const int col0 = n0 + c * 32;
                uint32_t r[32];
                tmem_ld_32x32b_x32(taddr + c * 32, r);
                tmem_ld_wait();
                if (c == kPairBlockN / 32 - 1) {
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[acc]);
                }
                if (col0 < p.N) {
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                    if (coalesced) epilogue_chunk_coalesced<QK>(p, m_warp0, col0, lane, v, stage_buf, res, row_ss, trow);
                    else epilogue_chunk(p, rc, col0, v);
                }
            };
            if (p.epi == EPI_CONV_NORM_PAD) {
                if constexpr (MODE == 1) epilogue_conv_norm_pad<kPairBlockN>(p, rc, taddr, &tmem_empty_bar[acc], 1, lane);
            } else {
#pragma unroll 1
                for (int c = 0; c < kPairBlockN / 32; c += 2) {
                    if (prefetch_res) epilogue_load_residual(p, m_warp0, n0 + (c + 1) * 32, lane, res_b);
                    chunk(c, res_a);
                    if (prefetch_res && c + 2 < kPairBlockN / 32) epilogue_load_residual(p, m_warp0, n0 + (c + 2) * 32, lane, res_a);
                    chunk(c + 1, res_b);
                    if constexpr (QK) epilogue_store_row_ss(p, m_warp0, n0 + c * 32, lane, row_ss);
                }
            }
            if (kEpiGroups == 1 && ++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    tcgen05_fence_before();
    cluster_sync_all();  // neither CTA leaves while its peer may still touch its barriers / shared memory
    if (warp_idx == 2) {
        tcgen05_fence_after();
        tmem_dealloc_2sm<PC::kTmemCols>(tmem_base);
    }
}

std::atomic<uint64_t> g_launches{0};

// Row tiles per raster group (tile_mn): the tiles in flight should span all column tiles of a few row tiles.  Only for
// plain GEMMs whose A operand cannot stay L2 resident (> 64 MB); smaller ones keep the row-fastest order.
int raster_group(const GemmParams& p, int tiles_in_flight, int num_n) {
    if (p.conv || options().gemm_no_raster) return 0;
    if (static_cast<double>(p.M) * p.K * 2.0 <= 64.0 * 1024 * 1024) return 0;
    const int g = tiles_in_flight / (num_n > 0 ? num_n : 1);
    return g < 1 ? 1 : g;
}

template <int BLOCK_N, bool QK = false>
cudaError_t launch_impl(const GemmOperands& ops, const GemmParams& p, cudaStream_t stream) {
    using C = Cfg<BLOCK_N>;
    static PerDeviceOnce configured;
    int cfg_dev = 0;
    static int num_sms = 0;
    if (configured.need(&cfg_dev)) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_kernel<BLOCK_N, QK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             C::kSmemBytes);
        if (e != cudaSuccess) return e;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        configured.mark(cfg_dev);
    }
    CUtensorMap ta, tb;
    cudaError_t e = make_tensor_map_2d_bf16(&ta, ops.a, ops.a_rows, ops.a_cols, kBlockM, kBlockK);
    if (e != cudaSuccess) return e;
    e = make_tensor_map_2d_bf16(&tb, ops.b, ops.b_rows, ops.b_cols, BLOCK_N, kBlockK);
    if (e != cudaSuccess) return e;
    const int num_m = (p.M + kBlockM - 1) / kBlockM;
    const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
    const int tiles = num_m * num_n;
    const int grid = tiles < num_sms ? tiles : num_sms;
    GemmParams pr = p;
    pr.raster_g = raster_group(p, grid, num_n);
    {
        double flops = 2.0 * p.M * static_cast<double>(p.N) * p.K;
        if (p.conv) flops = 2.0 * p.T * static_cast<double>(p.H) * p.W * static_cast<double>(p.N) * p.K;  // unpadded voxels
        ProfScope prof(p.conv ? PROF_CONV : PROF_GEMM, flops, stream);
        LTXV_TRACE_VARIANT("%s<%d> epi=%d", p.conv ? "conv3d:gemm_bf16_tn_kernel" : "gemm_bf16_tn_kernel", BLOCK_N, p.epi);
        cudaError_t le = launch_pdl(gemm_bf16_tn_kernel<BLOCK_N, QK>, dim3(grid), dim3(kThreads), C::kSmemBytes, stream, ta, tb, pr);
        if (le != cudaSuccess) return le;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}


template <int BN, int MODE, int E2 = 0>
cudaError_t launch_pair_impl(const GemmOperands& ops, const GemmParams& p, cudaStream_t stream) {
    using PC = PairCfg<BN, MODE, E2 != 0>;
    constexpr bool KW3 = PC::KW3;
    static PerDeviceOnce configured;
    int cfg_dev = 0;
    static int num_sms = 0;
    if (configured.need(&cfg_dev)) {
        cudaError_t e = cudaFuncSetAttribute(gemm_pair_bf16_tn_kernel<BN, MODE, E2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             PC::kSmemBytes);
        if (e != cudaSuccess) return e;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        configured.mark(cfg_dev);
    }
    if (KW3 && (!p.conv || p.num_k_blocks != 27 * p.cin_blocks)) return cudaErrorInvalidValue;
    if (MODE == 2 && p.num_k_blocks % 2 != 0) return cudaErrorInvalidValue;
    CUtensorMap ta, tb;
    cudaError_t e = make_tensor_map_2d_bf16(&ta, ops.a, ops.a_rows, ops.a_cols, KW3 ? kBoxRows3 : kBlockM, kBlockK);
    if (e != cudaSuccess) return e;
    e = make_tensor_map_2d_bf16(&tb, ops.b, ops.b_rows, ops.b_cols, BN / 2, kBlockK);
    if (e != cudaSuccess) return e;
    const int num_m = (p.M + 2 * kBlockM - 1) / (2 * kBlockM);
    const int num_n = (p.N + BN - 1) / BN;
    const int tiles = num_m * num_n;
    const int clusters = tiles < num_sms / 2 ? tiles : num_sms / 2;
    GemmParams pr = p;
    pr.raster_g = raster_group(p, clusters, num_n);
    {
        double flops = 2.0 * p.M * static_cast<double>(p.N) * p.K;
        if (p.conv) flops = 2.0 * p.T * static_cast<double>(p.H) * p.W * static_cast<double>(p.N) * p.K;
        ProfScope prof(p.conv ? PROF_CONV : PROF_GEMM, flops, stream);
        LTXV_TRACE_VARIANT("%s<%d,%d> epi=%d", p.conv ? "conv3d:gemm_pair_bf16_tn_kernel" : "gemm_pair_bf16_tn_kernel", BN, MODE,
                           p.epi);
        cudaError_t le = launch_pdl(gemm_pair_bf16_tn_kernel<BN, MODE, E2>, dim3(2 * clusters), dim3(PairThreads<MODE, E2 != 0>::value), PC::kSmemBytes,
                                    stream, ta, tb, pr);
        if (le != cudaSuccess) return le;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

}  // namespace

uint64_t gemm_launch_count() { return g_launches.load(); }

cudaError_t launch_gemm_bf16(const GemmOperands& ops, const GemmParams& p, int block_n, cudaStream_t stream) {
    if (p.M <= 0 || p.N <= 0 || p.K <= 0) return cudaErrorInvalidValue;
    const bool norm_pad = p.epi == EPI_CONV_NORM_PAD;  // one tile must hold every channel of a voxel
    if (norm_pad && (!p.conv || p.N > 256 || p.N % 32 != 0 || p.bias == nullptr || p.norm_out == nullptr ||
                     p.norm_out == p.a_ptr || p.norm_tf < 1 || p.norm_tf > 3 ||
                     (p.norm_scale == nullptr) != (p.norm_shift == nullptr)))
        return cudaErrorInvalidValue;
    if (p.epi == EPI_QKV_ROPE) {
        if (p.conv || p.N % 64 != 0 || p.qk_dim <= 0 || p.qk_dim % 64 != 0 || (p.qk_cols != p.qk_dim && p.qk_cols != 2 * p.qk_dim) ||
            p.qk_cols > p.N || p.qk_w[0] == nullptr || (p.qk_cols > p.qk_dim && p.qk_w[1] == nullptr) || p.qk_ss == nullptr ||
            (p.rope_cos == nullptr) != (p.rope_sin == nullptr) || (p.rope_cos != nullptr && p.rope_rows <= 0))
            return cudaErrorInvalidValue;
    }
    // (sums of squares are kept per absolute 64-column group, so any tile width works for EPI_QKV_ROPE)
    if (p.epi == EPI_QKV_ROPE && block_n != 0) return cudaErrorInvalidValue;  // tile width is chosen here
    if (block_n == -2) return launch_pair_impl<256, 0>(ops, p, stream);
    if (block_n == -3) return launch_pair_impl<128, 0>(ops, p, stream);
    if (block_n == -4) return launch_pair_impl<256, 1>(ops, p, stream);   // conv3d, three kw taps per step
    if (block_n == -5) return launch_pair_impl<128, 1>(ops, p, stream);
    if (block_n == -6) return launch_pair_impl<256, 2>(ops, p, stream);   // two k-blocks per stage
    if (block_n == -7) return launch_pair_impl<256, 0, 2>(ops, p, stream);  // two epilogue warpgroups
    if (block_n == -8) return launch_pair_impl<128, 0, 2>(ops, p, stream);
    if (block_n == 0) {
        // Pick the kernel / tile width that minimises (rounds over the SMs) x (per-SM tile area) / (relative rate of
        // that tile shape).  Rates from the isolated measurements at K = 8192 (profiles/r01_gemm_*): the CTA-pair
        // kernel (256x256 per cluster = 128x256 per SM) 1.00, single-CTA 128x256 0.94, 128x192 0.88, 128x128 0.75
        // (shared-memory bound), 128x64 0.50.
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int cands[4] = {256, 192, 128, 64};
        // Token-shard launches (M <= 4096: the per-rank row count of the 4/8-GPU Ulysses split): few rounds, so the
        // narrow tiles lose nothing to the wide ones and balance better.  Re-calibrated on M = 1248 / 1672
        // (tools/bin/gemm_test 6): at K <= 2048 128x192 and 128x128 run within 5% of 128x256 per unit area, and the
        // 256x128 cluster tile is the fastest shape of all (1051 vs 931 TFLOP/s on the QKV shape at M = 1672).
        const bool shard_m = !p.conv && p.M <= 4096;
        const bool short_k = p.K <= 2048;
        const double rate_big[4] = {0.94, 0.88, 0.75, 0.50};
        const double rate_shard[4] = {0.94, 0.97, 0.95, 0.60};
        const double* rate = shard_m && short_k ? rate_shard : rate_big;
        const bool epi2 = !options().gemm_no_epi2 && (p.epi == EPI_STORE_BF16 || p.epi == EPI_QKV_ROPE);  // two epilogue warpgroups
        const double pair128_rate = !shard_m ? 0.80 : (short_k ? (epi2 ? 1.15 : 1.05) : (p.M <= 2048 ? 1.02 : 0.95));  // K = 8192 at M = 3344: 256x256 pairs 1112 vs 1027 TFLOP/s
        double best = 1e30;
        const int num_m = (p.M + kBlockM - 1) / kBlockM;
        for (int i = 0; i < 4; ++i) {
            const int c = cands[i];
            if (c > 64 && p.N <= c / 2) continue;
            if (norm_pad && (c < p.N || (c != 128 && c != 256))) continue;
            const int tiles = num_m * ((p.N + c - 1) / c);
            const int rounds = (tiles + sms - 1) / sms;
            const double cost = rounds * static_cast<double>(c) / rate[i];
            if (cost < best) {
                best = cost;
                block_n = c;
            }
        }
        // Short-K square-ish projections (the batched-CFG attention out / cross-attention q projections: N = K = 2048,
        // M = 9984): the rates above were calibrated at K = 8192; at K = 2048 the small tiles are as efficient as the
        // large ones and the 5.8 waves of 128x192 tiles beat the 4.2 rounds of 256x256 pairs (measured 1244 vs 1196
        // TFLOP/s with the bf16 store, 890 vs 866 with the f32 residual epilogue, tools/bin/gemm_test 3).
        const bool no_short_k_rule = options().gemm_no_short_k != 0;
        // ... except for the f32 residual epilogue, which runs on the pair kernel with two epilogue warpgroups (E2 = 2)
        const bool short_k_192 = !no_short_k_rule && !p.conv && p.K <= 2048 && p.N <= 2048 && p.N % 192 != 0 &&
                                 p.N > 1024 && p.M >= 8192 && !(p.epi == EPI_RESIDUAL_F32 && !options().gemm_no_epi2);
        if (short_k_192) block_n = 192;
        if (!short_k_192 && p.M > 2 * kBlockM && !options().gemm_no_pair) {
            const int num_mp = (p.M + 2 * kBlockM - 1) / (2 * kBlockM);
            int pair_bn = 0;
            if (p.N % 256 == 0) {
                const int rounds = (num_mp * (p.N / 256) + sms / 2 - 1) / (sms / 2);
                if (rounds * 256.0 / 1.00 <= best) {
                    best = rounds * 256.0;
                    pair_bn = 256;
                }
            }
            if (p.N % 128 == 0 && !(norm_pad && p.N > 128)) {  // 256x128 cluster tiles: the only way to share B when N = 128 (VAE level-3 convs)
                const int rounds = (num_mp * (p.N / 128) + sms / 2 - 1) / (sms / 2);
                if (rounds * 128.0 / pair128_rate < best) {
                    best = rounds * 128.0 / pair128_rate;
                    pair_bn = 128;
                }
            }
            const bool kw3 = p.conv && p.num_k_blocks == 27 * p.cin_blocks && !options().conv_no_kw3;
            if (norm_pad && !kw3) pair_bn = 0;  // only the KW3 pair kernels carry the fused producer epilogue
            // two k-blocks per stage measured neutral (1341 vs 1353 TFLOP/s on the QKV shape): opt-in only
            const bool k2 = !p.conv && p.num_k_blocks % 2 == 0 && options().gemm_k2 != 0;
            if (p.epi == EPI_QKV_ROPE && pair_bn == 256) return launch_pair_impl<256, 0, 1>(ops, p, stream);
            if (p.epi == EPI_QKV_ROPE && pair_bn == 128) return launch_pair_impl<128, 0, 1>(ops, p, stream);
            // f32 residual read-modify-write behind a SHORT main loop (attention out-projections, K <= 2048): two
            // epilogue warpgroups
            if (p.epi == EPI_RESIDUAL_F32 && p.K <= 2048 && pair_bn == 256 && !kw3 && !options().gemm_no_epi2)
                return launch_pair_impl<256, 0, 2>(ops, p, stream);
            // ... and the bf16-store epilogues behind a short main loop (FFN-in with GELU: 1293 -> 1400 TFLOP/s at
            // M = 9984, tools/bin/gemm_test 7; at K = 8192 the main loop hides the epilogue: one group)
            if (p.epi == EPI_STORE_BF16 && !p.conv && p.K <= 2048 && pair_bn == 256 && !options().gemm_no_epi2)
                return launch_pair_impl<256, 0, 2>(ops, p, stream);
            if (p.epi == EPI_STORE_BF16 && !p.conv && p.K <= 2048 && pair_bn == 128 && !options().gemm_no_epi2)
                return launch_pair_impl<128, 0, 2>(ops, p, stream);  // shard shapes: FFN-in at M = 1672 990 -> 1139 TFLOP/s
            if (p.epi == EPI_RESIDUAL_F32 && p.K <= 2048 && pair_bn == 128 && !kw3 && !options().gemm_no_epi2)
                return launch_pair_impl<128, 0, 2>(ops, p, stream);
            if (pair_bn == 256)
                return kw3 ? launch_pair_impl<256, 1>(ops, p, stream)
                           : (k2 ? launch_pair_impl<256, 2>(ops, p, stream) : launch_pair_impl<256, 0>(ops, p, stream));
            if (pair_bn == 128) return kw3 ? launch_pair_impl<128, 1>(ops, p, stream) : launch_pair_impl<128, 0>(ops, p, stream);
        }
    }
    if (p.epi == EPI_QKV_ROPE) {
        switch (block_n) {
            case 256: return launch_impl<256, true>(ops, p, stream);
            case 192: return launch_impl<192, true>(ops, p, stream);
            case 128: return launch_impl<128, true>(ops, p, stream);
            case 64: return launch_impl<64, true>(ops, p, stream);
            default: return cudaErrorInvalidValue;
        }
    }
    switch (block_n) {
        case 256: return launch_impl<256>(ops, p, stream);
        case 192: return launch_impl<192>(ops, p, stream);
        case 128: return launch_impl<128>(ops, p, stream);
        case 64: return launch_impl<64>(ops, p, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace ltxv
