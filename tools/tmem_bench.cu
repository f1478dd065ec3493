// Micro-benchmark: TMEM -> register read bandwidth (tcgen05.ld.32x32b.xN) per SM on sm_100a.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../candle_video_b200/csrc/common.cuh"
using namespace ltxv;

template <int X>
__device__ __forceinline__ uint32_t ld_acc(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ld_acc<32>(uint32_t taddr) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(taddr, r);
    tmem_ld_wait();
    uint32_t a = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) a ^= r[i];
    return a;
}
// 4 back-to-back x32 loads then one wait (what the attention softmax does)
template <>
__device__ __forceinline__ uint32_t ld_acc<128>(uint32_t taddr) {
    uint32_t r0[32], r1[32], r2[32], r3[32];
    tmem_ld_32x32b_x32(taddr, r0);
    tmem_ld_32x32b_x32(taddr + 32, r1);
    tmem_ld_32x32b_x32(taddr + 64, r2);
    tmem_ld_32x32b_x32(taddr + 96, r3);
    tmem_ld_wait();
    uint32_t a = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) a ^= r0[i] ^ r1[i] ^ r2[i] ^ r3[i];
    return a;
}

template <int X>
__global__ void k(uint32_t* out, long long* cycles, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(&slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) acc ^= ld_acc<X>(base + ((it * X) & 255) + (warp >> 2) * 0);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(slot);
}

template <int X>
void run(int warps) {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 4096;
    k<X><<<148, warps * 32>>>(out, cyc, 16);
    k<X><<<148, warps * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double bytes = (double)warps * iters * X * 32 * 4;
    printf("x%-3d warps/CTA=%2d: %lld clks, %.1f B/clk/SM, %.1f clk per warp-load of %d B  (%s)\n", X, warps, h[0],
           bytes / h[0], (double)h[0] / iters, X * 128, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<32>(1); run<32>(4); run<32>(8); run<32>(16);
    run<128>(1); run<128>(4); run<128>(8);
    return 0;
}
