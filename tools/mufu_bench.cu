// Micro-benchmark: throughput of ex2.approx.ftz.f32 vs ex2.approx.ftz.f16x2 vs an FMA-pipe polynomial exp2 on sm_100a.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ float ex2_f32(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned ex2_h2(unsigned x) { unsigned y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ unsigned ex2_bf2(unsigned x) { unsigned y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
// 2^x for x <= 0 (x >= -126): Cody-Waite split + degree-3 minimax on [0,1)
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -126.0f);
    float fl = floorf(x);
    float f = x - fl;
    float p = fmaf(f, 0.0555041086f, 0.2402264923f);
    p = fmaf(p, f, 0.6931471825f);
    p = fmaf(p, f, 1.0f);
    int e = (int)fl;
    return __int_as_float(__float_as_int(p) + (e << 23));
}

template <int MODE>
__global__ void k(float* out, int iters) {
    float a[8];
    unsigned h[8];
    for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0xb800b800u + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = ex2_f32(a[i]) - 1.0f;
            if (MODE == 1) h[i] = ex2_h2(h[i]) ^ 0x80008000u;
            if (MODE == 2) a[i] = ex2_poly(a[i]) - 1.0f;
            if (MODE == 3) h[i] = ex2_bf2(h[i]) ^ 0x80008000u;
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int per_op) {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    k<MODE><<<148, 1024>>>(out, 100);
    cudaEventRecord(e0);
    k<MODE><<<148, 1024>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * 1024 * iters * 8 * per_op;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s %.3f ms  %.1f Gexp/s  = %.2f exp/clk/SM at %d MHz nominal\n", name, ms, ops / ms * 1e-6,
           ops / (ms * 1e-3) / 148.0 / (clk * 1e3), clk / 1000);
    cudaFree(out);
}
int main() {
    run<0>("ex2.approx.ftz.f32", 1);
    run<1>("ex2.approx.ftz.f16x2", 2);
    run<2>("poly exp2 (FMA pipe)", 1);
    run<3>("ex2.approx.ftz.bf16x2", 2);
    return 0;
}
