// Self-attention throughput at the in-step shapes (CUDA events, warm-up, L2-sized inputs): `attn_bench [B] [S] [H] [D] [iters]`
// defaults = the batched-CFG c2 launch (B = 2, S = 4992, 32 heads x 64).  Built in several -D variants by
// tools/build_attn_variants.sh to compare kernel experiments in ONE gpurun call.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "../candle_video_b200/csrc/attention.h"
using namespace ltxv;
__global__ void fill(__nv_bfloat16* p, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) { uint32_t x = (uint32_t)i * 2654435761u; x ^= x >> 15; p[i] = __float2bfloat16(((x >> 8) * (1.0f / 16777216.0f) - 0.5f) * 2.f); }
}
int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 2, S = argc > 2 ? atoi(argv[2]) : 4992, H = argc > 3 ? atoi(argv[3]) : 32,
              D = argc > 4 ? atoi(argv[4]) : 64, iters = argc > 5 ? atoi(argv[5]) : 30;
    const int HD = H * D;
    __nv_bfloat16 *q, *o;
    const size_t nq = (size_t)B * S * 3 * HD, no = (size_t)B * S * HD;
    cudaMalloc(&q, nq * 2); cudaMalloc(&o, no * 2);
    fill<<<(nq + 255) / 256, 256>>>(q, nq);
    AttnParams p{}; p.q = p.k = p.v = q; p.ldq = p.ldk = p.ldv = 3 * HD; p.k_col0 = HD; p.v_col0 = 2 * HD; p.out = o; p.ldo = HD;
    p.B = B; p.H = H; p.Sq = S; p.Skv = S; p.D = D; p.scale = 1.0f / sqrtf((float)D);
    for (int i = 0; i < 5; ++i) launch_attention(p, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) launch_attention(p, 0);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 4.0 * B * H * (double)S * S * D;
    // checksum so variants can be compared for gross errors
    __nv_bfloat16* ho = (__nv_bfloat16*)malloc(no * 2);
    cudaMemcpy(ho, o, no * 2, cudaMemcpyDeviceToHost);
    double cs = 0, ca = 0; for (size_t i = 0; i < no; i += 7) { float v = __bfloat162float(ho[i]); cs += v; ca += fabs(v); }
    printf("%s B=%d S=%d H=%d D=%d: %.1f us/launch  %.1f TFLOP/s  checksum %.6f abs %.4f  (%s)\n", argv[0], B, S, H, D,
           1000.0 * ms / iters, fl / (ms / iters * 1e-3) / 1e12, cs, ca, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
