// Standalone bring-up test for the tcgen05 flash-attention kernel against a naive f32 CUDA reference.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../candle_video_b200/csrc/attention.h"
#include "../candle_video_b200/csrc/tensormap.h"

using namespace ltxv;

#define CK(x)                                                                                        \
    do {                                                                                             \
        cudaError_t e_ = (x);                                                                        \
        if (e_ != cudaSuccess) {                                                                     \
            printf("CUDA error %s at %s:%d (%s)\n", cudaGetErrorString(e_), __FILE__, __LINE__,      \
                   tensor_map_last_error());                                                         \
            exit(2);                                                                                 \
        }                                                                                            \
    } while (0)

__global__ void fill_bf16(__nv_bfloat16* p, size_t n, uint32_t seed, float scale) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t x = (uint32_t)i * 2654435761u + seed;
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    float u = (x >> 8) * (1.0f / 16777216.0f) - 0.5f;
    p[i] = __float2bfloat16(u * 2.0f * scale);
}

// one thread per (b, h, q): two-pass softmax in f32
__global__ void ref_attn(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, const float* bias,
                         float* out, int B, int H, int Sq, int Skv, int D, int64_t ldq, int64_t ldk, int64_t ldv,
                         int qc, int kc, int vc, float scale) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H * Sq) return;
    int qi = idx % Sq, h = (idx / Sq) % H, b = idx / (Sq * H);
    const __nv_bfloat16* qp = q + ((int64_t)b * Sq + qi) * ldq + qc + h * D;
    float mx = -INFINITY;
    for (int j = 0; j < Skv; ++j) {
        const __nv_bfloat16* kp = k + ((int64_t)b * Skv + j) * ldk + kc + h * D;
        float s = 0;
        for (int d = 0; d < D; ++d) s += __bfloat162float(qp[d]) * __bfloat162float(kp[d]);
        s = s * scale + (bias ? bias[b * Skv + j] : 0.f);
        mx = fmaxf(mx, s);
    }
    float acc[128];
    for (int d = 0; d < D; ++d) acc[d] = 0;
    float l = 0;
    for (int j = 0; j < Skv; ++j) {
        const __nv_bfloat16* kp = k + ((int64_t)b * Skv + j) * ldk + kc + h * D;
        const __nv_bfloat16* vp = v + ((int64_t)b * Skv + j) * ldv + vc + h * D;
        float s = 0;
        for (int d = 0; d < D; ++d) s += __bfloat162float(qp[d]) * __bfloat162float(kp[d]);
        s = s * scale + (bias ? bias[b * Skv + j] : 0.f);
        float pj = expf(s - mx);
        l += pj;
        for (int d = 0; d < D; ++d) acc[d] += pj * __bfloat162float(vp[d]);
    }
    for (int d = 0; d < D; ++d) out[((int64_t)b * Sq + qi) * (H * D) + h * D + d] = acc[d] / l;
}

static int run(int B, int H, int Sq, int Skv, int D, bool packed, bool use_bias, bool timing, float qscale) {
    const int HD = H * D;
    int64_t ldq, ldk, ldv;
    int qc = 0, kc = 0, vc = 0;
    __nv_bfloat16 *q, *k, *v, *out;
    if (packed) {  // self-attention layout: one [B,S,3*HD] buffer
        ldq = ldk = ldv = 3 * HD;
        CK(cudaMalloc(&q, (size_t)B * Sq * 3 * HD * 2));
        fill_bf16<<<((size_t)B * Sq * 3 * HD + 255) / 256, 256>>>(q, (size_t)B * Sq * 3 * HD, 7, qscale);
        k = q; v = q; kc = HD; vc = 2 * HD;
    } else {  // cross-attention layout: q [B,Sq,HD], kv [B,Skv,2*HD]
        ldq = HD; ldk = ldv = 2 * HD;
        CK(cudaMalloc(&q, (size_t)B * Sq * HD * 2));
        CK(cudaMalloc(&k, (size_t)B * Skv * 2 * HD * 2));
        fill_bf16<<<((size_t)B * Sq * HD + 255) / 256, 256>>>(q, (size_t)B * Sq * HD, 8, qscale);
        fill_bf16<<<((size_t)B * Skv * 2 * HD + 255) / 256, 256>>>(k, (size_t)B * Skv * 2 * HD, 9, qscale);
        v = k; vc = HD;
    }
    CK(cudaMalloc(&out, (size_t)B * Sq * HD * 2));
    CK(cudaMemset(out, 0, (size_t)B * Sq * HD * 2));
    float* bias = nullptr;
    if (use_bias) {
        std::vector<float> hb((size_t)B * Skv);
        for (int b = 0; b < B; ++b)
            for (int j = 0; j < Skv; ++j) hb[(size_t)b * Skv + j] = (j < Skv / 3 + b) ? 0.f : -10000.f;
        CK(cudaMalloc(&bias, hb.size() * 4));
        CK(cudaMemcpy(bias, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
    }
    float* ref;
    CK(cudaMalloc(&ref, (size_t)B * Sq * HD * 4));
    const float scale = 1.0f / sqrtf((float)D);
    ref_attn<<<(B * H * Sq + 127) / 128, 128>>>(q, k, v, bias, ref, B, H, Sq, Skv, D, ldq, ldk, ldv, qc, kc, vc, scale);
    CK(cudaDeviceSynchronize());

    AttnParams p{};
    p.q = q; p.k = k; p.v = v; p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.q_col0 = qc; p.k_col0 = kc; p.v_col0 = vc;
    p.out = out; p.ldo = HD; p.kv_bias = bias; p.B = B; p.H = H; p.Sq = Sq; p.Skv = Skv; p.D = D; p.scale = scale;
    CK(launch_attention(p, 0));
    CK(cudaDeviceSynchronize());
    size_t n = (size_t)B * Sq * HD;
    std::vector<float> h_ref(n);
    std::vector<__nv_bfloat16> h_out(n);
    CK(cudaMemcpy(h_ref.data(), ref, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_out.data(), out, n * 2, cudaMemcpyDeviceToHost));
    double max_abs = 0, max_ref = 0, se = 0, sr = 0;
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) {
        double g = __bfloat162float(h_out[i]), r = h_ref[i], d = fabs(g - r);
        if (d > max_abs) max_abs = d;
        if (fabs(r) > max_ref) max_ref = fabs(r);
        se += d * d; sr += r * r;
        if (!(d <= 1e-2 + 2e-2 * fabs(r))) bad++;
    }
    printf("attn B=%d H=%d Sq=%d Skv=%d D=%d packed=%d bias=%d qscale=%.1f: max_abs=%.3e max_ref=%.3e rel_l2=%.3e bad=%zu %s\n",
           B, H, Sq, Skv, D, packed, use_bias, qscale, max_abs, max_ref, sqrt(se / (sr + 1e-30)), bad, bad ? "FAIL" : "ok");
    if (timing) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int i = 0; i < 3; ++i) launch_attention(p, 0);
        cudaEventRecord(e0);
        const int iters = 10;
        for (int i = 0; i < iters; ++i) launch_attention(p, 0);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        ms /= iters;
        printf("   time %.3f ms  %.1f TFLOP/s (4*Sq*Skv*H*D)\n", ms, 4.0 * B * H * (double)Sq * Skv * D / ms * 1e-9);
    }
    cudaFree(q); if (!packed) cudaFree(k); cudaFree(out); cudaFree(ref); if (bias) cudaFree(bias);
    return bad ? 1 : 0;
}

int main(int argc, char** argv) {
    bool big = argc > 1 && atoi(argv[1]) > 0;
    int fails = 0;
    if (argc > 1 && atoi(argv[1]) == 2) {  // timing of the Ulysses shard shapes only
        run(1, 8, 4992, 4992, 64, true, false, true, 1.0f);
        run(1, 16, 4992, 4992, 64, true, false, true, 1.0f);
        run(1, 32, 4992, 4992, 64, true, false, true, 1.0f);
        return 0;
    }
    fails += run(1, 2, 128, 128, 64, true, false, false, 1.0f);
    fails += run(1, 2, 256, 256, 64, true, false, false, 1.0f);
    fails += run(2, 4, 384, 384, 64, true, false, false, 2.0f);   // several kv tiles, batch
    fails += run(1, 3, 200, 200, 64, true, false, false, 4.0f);   // ragged q and kv tails, peaky softmax (rescale path)
    fails += run(1, 2, 1000, 1000, 64, true, false, false, 6.0f);
    fails += run(1, 2, 300, 700, 64, false, false, false, 3.0f);  // 2-tile kernel: odd query tile count, ragged kv tail
    fails += run(2, 3, 640, 1300, 64, false, false, false, 5.0f);  // 2-tile kernel: batch, peaky softmax (rescale path)
    fails += run(1, 2, 300, 2000, 64, false, false, false, 4.0f);  // tail split x4, odd query tile count, ragged kv tail
    fails += run(1, 4, 1280, 4096, 64, false, false, false, 2.0f);  // tail split: 20 units -> 7 key ranges each
    fails += run(2, 4, 384, 128, 64, false, true, false, 1.0f);   // cross-attention with key-padding bias
    fails += run(1, 4, 300, 77, 64, false, true, false, 1.0f);    // ragged text length
    fails += run(1, 2, 256, 256, 128, true, false, false, 1.0f);  // 13B head_dim
    fails += run(1, 2, 300, 128, 128, false, true, false, 1.0f);
    fails += run(1, 2, 384, 384, 128, true, false, false, 2.0f);    // head_dim-128 two-tile kernel: odd tile count
    fails += run(2, 3, 640, 1300, 128, false, false, false, 5.0f);  // batch, ragged kv tail, peaky softmax (rescale path)
    fails += run(1, 2, 1000, 1000, 128, true, false, false, 6.0f);
    if (big) {
        fails += run(1, 32, 4992, 4992, 64, true, false, true, 1.0f);
        fails += run(1, 32, 4992, 128, 64, false, true, true, 1.0f);
        fails += run(1, 8, 4992, 4992, 64, true, false, true, 1.0f);   // Ulysses shard at 8 GPUs: 160 units on 148 SMs
        fails += run(1, 16, 4992, 4992, 64, true, false, true, 1.0f);  // Ulysses shard at 4 GPUs
        fails += run(1, 32, 13376, 13376, 64, true, false, true, 1.0f);
        fails += run(1, 32, 4992, 4992, 128, true, false, true, 1.0f);
        fails += run(1, 32, 19320, 19320, 128, true, false, true, 1.0f);  // c4: 13B at 736x1280x161
    }
    printf("%s (%d failing cases)\n", fails ? "ATTN_TEST_FAIL" : "ATTN_TEST_OK", fails);
    return fails ? 1 : 0;
}
