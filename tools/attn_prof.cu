// one self-attention launch at the c2 shape for ncu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "../candle_video_b200/csrc/attention.h"
using namespace ltxv;
namespace ltxv { void attention_debug_timing(long long* out32); void attention_debug_trace(long long* out); }
__global__ void fill(__nv_bfloat16* p, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) { uint32_t x = (uint32_t)i * 2654435761u; x ^= x >> 15; p[i] = __float2bfloat16(((x >> 8) * (1.0f / 16777216.0f) - 0.5f) * 2.f); }
}
int main(int argc, char** argv) {
    // usage: attn_prof [S] [H] [D]   (defaults: the c2 shape; `attn_prof 4992 32 128` = the 13B head layout)
    const int S = argc > 1 ? atoi(argv[1]) : 4992, H = argc > 2 ? atoi(argv[2]) : 32, D = argc > 3 ? atoi(argv[3]) : 64;
    const int HD = H * D;
    __nv_bfloat16 *q, *o;
    cudaMalloc(&q, (size_t)S * 3 * HD * 2); cudaMalloc(&o, (size_t)S * HD * 2);
    fill<<<((size_t)S * 3 * HD + 255) / 256, 256>>>(q, (size_t)S * 3 * HD);
    AttnParams p{}; p.q = p.k = p.v = q; p.ldq = p.ldk = p.ldv = 3 * HD; p.k_col0 = HD; p.v_col0 = 2 * HD; p.out = o; p.ldo = HD;
    p.B = 1; p.H = H; p.Sq = S; p.Skv = S; p.D = D; p.scale = 1.0f / sqrtf((float)D);
    for (int i = 0; i < 3; ++i) launch_attention(p, 0);
    cudaDeviceSynchronize();
    printf("done %s\n", cudaGetErrorString(cudaGetLastError()));
#ifdef LTXV_ATTN_TIMING
    long long t[32];
    ltxv::attention_debug_timing(t);
    const char* names[8] = {"wait s_full", "tmem ld + s_free", "max", "rescale check", "exp first half", "wait pv_done", "sttm + exp 2nd half", "st wait+arrive"};
    for (int w = 0; w < 2; ++w) {
        long long tot = 0;
        for (int i = 0; i < 8; ++i) tot += t[w * 8 + i];
        printf("softmax warp of tile %d: total %lld clks over 39 kv tiles = %lld per tile\n", w, tot, tot / 39);
        for (int i = 0; i < 8; ++i) printf("   %-18s %8lld  (%5.1f%%)  %lld/tile\n", names[i], t[w * 8 + i], 100.0 * t[w * 8 + i] / tot, t[w * 8 + i] / 39);
    }
    const char* mn[8] = {"wait k_full", "wait s_free", "issue S", "wait v_full", "wait p_full", "issue PV", "dummy satisfied wait", "loop edge"};
    for (int w = 0; w < 2; ++w) {
        long long tot = 0;
        for (int i = 0; i < 8; ++i) tot += t[16 + 8 * w + i];
        printf("MMA thread of tile %d: total %lld clks = %lld per kv tile\n", w, tot, tot / 39);
        for (int i = 0; i < 8; ++i) printf("   %-16s %8lld (%5.1f%%) %lld/tile\n", mn[i], t[16 + 8 * w + i], 100.0 * t[16 + 8 * w + i] / tot, t[16 + 8 * w + i] / 39);
    }
#endif
#ifdef LTXV_ATTN_TRACE
    {
        static long long tr[11][40][8];
        ltxv::attention_debug_trace(&tr[0][0][0]);
        const long long t0 = tr[0][16][0];
        const char* sm[6] = {"at_wait_sfull", "sfull_seen", "sfree_arrived", "exp1_done", "pvdone_seen", "pfull_arrived"};
        const char* im[6] = {"kfull_seen", "sfree_seen", "S(j+1)_issued", "vfull_seen", "pfull_seen", "PV(j)_issued"};
        printf("CTA lifetime (clk): entry=0 setup_done=%lld softmax_loop_start=%lld pv_last_done=%lld exit_sync=%lld; first tiles: ",
               tr[10][39][1] - tr[10][39][0], tr[10][39][2] - tr[10][39][0], tr[10][39][3] - tr[10][39][0], tr[10][39][4] - tr[10][39][0]);
        for (int j = 0; j < 6; ++j) printf("sfull_seen[%d]=%lld ", j, tr[0][j][1] - tr[10][39][0]);
        printf(" ... pfull[38]=%lld\n", tr[0][38][5] - tr[10][39][0]);
        printf("grid first CTA entry (ns) = 0, its exit = %lld\n", tr[9][19][5] - tr[9][19][4]);
        for (int i = 0; i < 8; ++i) printf("  last-%d CTA entry at %lld ns\n", i, tr[9][10 + i][0] - tr[9][19][4]);
        for (int sidx = 0; sidx < 8; ++sidx)
            if (tr[9][20 + sidx][0] != 0)
                printf("split item %d of first split unit: loop_done=%lld published=%lld ticket=%lld exit=%lld (clk since entry) last=%lld exit_ns=%lld\n", sidx,
                       tr[9][20 + sidx][1] - tr[9][20 + sidx][0], tr[9][20 + sidx][2] - tr[9][20 + sidx][0],
                       tr[9][20 + sidx][3] - tr[9][20 + sidx][0], tr[9][20 + sidx][4] - tr[9][20 + sidx][0], tr[9][20 + sidx][6],
                       tr[9][20 + sidx][5] - tr[9][19][4]);
        for (int j = 15; j < 18; ++j) {
            for (int w = 0; w < 8; ++w) {
                printf("j=%d softmax warp %d (tile %d):", j, w, w >> 2);
                for (int e = 0; e < 6; ++e) printf(" %s=%lld", sm[e], tr[w][j][e] - t0);
                printf("\n");
            }
            for (int t = 0; t < 2; ++t) {
                printf("j=%d issuer  t%d:", j, t);
                for (int e = 0; e < 6; ++e) printf(" %s=%lld", im[e], tr[8 + t][j][e] - t0);
                printf("\n");
            }
            printf("j=%d producer: K(j)_issued=%lld V(j)_issued=%lld\n", j, tr[10][j][0] - t0, tr[10][j][1] - t0);
        }
    }
#endif
    return 0;
}
