// one self-attention launch at the c2 shape for ncu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include "../candle_video_b200/csrc/attention.h"
using namespace ltxv;
namespace ltxv { void attention_debug_timing(long long* out32); }
__global__ void fill(__nv_bfloat16* p, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) { uint32_t x = (uint32_t)i * 2654435761u; x ^= x >> 15; p[i] = __float2bfloat16(((x >> 8) * (1.0f / 16777216.0f) - 0.5f) * 2.f); }
}
int main() {
    const int S = 4992, H = 32, D = 64, HD = H * D;
    __nv_bfloat16 *q, *o;
    cudaMalloc(&q, (size_t)S * 3 * HD * 2); cudaMalloc(&o, (size_t)S * HD * 2);
    fill<<<((size_t)S * 3 * HD + 255) / 256, 256>>>(q, (size_t)S * 3 * HD);
    AttnParams p{}; p.q = p.k = p.v = q; p.ldq = p.ldk = p.ldv = 3 * HD; p.k_col0 = HD; p.v_col0 = 2 * HD; p.out = o; p.ldo = HD;
    p.B = 1; p.H = H; p.Sq = S; p.Skv = S; p.D = D; p.scale = 0.125f;
    for (int i = 0; i < 3; ++i) launch_attention(p, 0);
    cudaDeviceSynchronize();
    printf("done %s\n", cudaGetErrorString(cudaGetLastError()));
#ifdef LTXV_ATTN_TIMING
    long long t[32];
    ltxv::attention_debug_timing(t);
    const char* names[8] = {"wait s_full", "tmem ld + s_free", "max", "wait pv_done", "rescale", "exp+pack+store", "fence+arrive", ""};
    for (int w = 0; w < 2; ++w) {
        long long tot = 0;
        for (int i = 0; i < 7; ++i) tot += t[w * 8 + i];
        printf("softmax warp of tile %d: total %lld clks over 39 kv tiles = %lld per tile\n", w, tot, tot / 39);
        for (int i = 0; i < 7; ++i) printf("   %-18s %8lld  (%5.1f%%)  %lld/tile\n", names[i], t[w * 8 + i], 100.0 * t[w * 8 + i] / tot, t[w * 8 + i] / 39);
    }
    const char* mn[8] = {"wait k_full", "wait s_free[0]", "wait s_free[1]", "issue S", "wait v_full", "wait p_full[0]", "wait p_full[1]", "issue PV"};
    long long tot = 0;
    for (int i = 0; i < 8; ++i) tot += t[16 + i];
    printf("MMA thread: total %lld clks = %lld per kv tile\n", tot, tot / 39);
    for (int i = 0; i < 8; ++i) printf("   %-16s %8lld (%5.1f%%) %lld/tile\n", mn[i], t[16 + i], 100.0 * t[16 + i] / tot, t[16 + i] / 39);
#endif
    return 0;
}
