// LtxVideoCausalConv3d::forward (vae.rs:298-465) as a standalone operator behind ltxv_causal_conv3d: the same
// implicit-GEMM tcgen05 kernel the decoder / encoder use (gemm_tcgen05.cu), wrapped with the layout conversions a caller
// holding NCDHW f32 tensors needs.  It exists so the conv can be checked against the reference's own operator-level
// test (tests/verify_conv3d_parity.rs:41-43, scripts/gen_conv3d_ref.py:28-31: [1,1024,8,16,16], 1024 -> 4096) and at
// the production volumes that select the CTA-pair / three-taps-per-step kernel variants.
#include "conv3d_op.h"

#include <atomic>

#include "gemm.h"
#include "vae_glue.h"

namespace ltxv {
namespace {

std::atomic<uint64_t> g_conv_op_launches{0};

// x [C, T, H, W] f32 (NCDHW, B = 1) -> padded NDHWC bf16 [(T+2), (H+2), (W+2), C].  tf = 1: non-causal, frames 0 and
// T+1 replicate frames 1 and T (vae.rs:388-411); tf = 2: causal, frames 0,1 replicate frame 2 (vae.rs:383-387).
// One block per (t, h) row: the W x C tile is transposed through shared memory so both sides are coalesced.
__global__ void conv_op_input_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int C, int T, int H,
                                     int W, int tf) {
    extern __shared__ float tile[];  // [cb = 32][W + 1]
    const int t = blockIdx.x / H, h = blockIdx.x % H;
    const int Hp = H + 2, Wp = W + 2;
    for (int c0 = blockIdx.y * 32; c0 < C; c0 += gridDim.y * 32) {
        for (int i = threadIdx.x; i < 32 * W; i += blockDim.x) {
            const int c = i / W, w = i % W;
            tile[c * (W + 1) + w] = (c0 + c < C) ? x[((static_cast<int64_t>(c0 + c) * T + t) * H + h) * W + w] : 0.f;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * W; i += blockDim.x) {
            const int w = i / 32, c = i % 32;
            if (c0 + c >= C) continue;
            const __nv_bfloat16 b = __float2bfloat16(tile[c * (W + 1) + w]);
            auto at = [&](int tp) { return ((static_cast<int64_t>(tp) * Hp + (h + 1)) * Wp + (w + 1)) * C + c0 + c; };
            out[at(t + tf)] = b;
            if (t == 0)
                for (int k = 0; k < tf; ++k) out[at(k)] = b;
            if (tf == 1 && t == T - 1) out[at(T + 1)] = b;
        }
        __syncthreads();
    }
}

// y [T*H*W, C] bf16 (NDHWC) -> out [C, T*H*W] f32 (NCDHW): 32 x 32 smem transpose
__global__ void conv_op_output_kernel(const __nv_bfloat16* __restrict__ y, float* __restrict__ out, int64_t nvox, int C) {
    __shared__ float tile[32][33];
    const int64_t v0 = static_cast<int64_t>(blockIdx.x) * 32;
    const int c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t v = v0 + r;
        const int c = c0 + threadIdx.x;
        tile[r][threadIdx.x] = (v < nvox && c < C) ? __bfloat162float(y[v * C + c]) : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r;
        const int64_t v = v0 + threadIdx.x;
        if (v < nvox && c < C) out[static_cast<int64_t>(c) * nvox + v] = tile[threadIdx.x][r];
    }
}

}  // namespace

uint64_t conv_op_launch_count() { return g_conv_op_launches.load(); }

void causal_conv3d(const float* x, const float* weight, const float* bias, int Cin, int Cout, int T, int H, int W,
                   bool is_causal, float* out, cudaStream_t s) {
    if (Cin <= 0 || Cout <= 0 || T <= 0 || H <= 0 || W <= 0) fail("causal_conv3d: invalid shape");
    if (Cin % 64 != 0) fail("causal_conv3d: in_channels (%d) must be a multiple of 64", Cin);
    if (Cout % 32 != 0) fail("causal_conv3d: out_channels (%d) must be a multiple of 32", Cout);
    if (W > 352) fail("causal_conv3d: width %d exceeds the staging tile of the layout kernel (352)", W);
    const int tf = is_causal ? 2 : 1;
    const int rows_out = (Cout + 63) / 64 * 64;
    const int Wp = W + 2, plane = (H + 2) * Wp;
    const int64_t nvox = static_cast<int64_t>(T) * H * W;
    DevBuf a, wb, bb, y;
    a.ensure(static_cast<size_t>(T + 2) * plane * Cin * 2, true);
    wb.ensure(static_cast<size_t>(rows_out) * 27 * Cin * 2, true);
    bb.ensure(static_cast<size_t>((rows_out + 255) / 256 * 256) * 4, true);
    y.ensure(static_cast<size_t>(nvox) * Cout * 2);
    LTXV_CUDA(cudaStreamSynchronize(0));  // the zero fills ran on the legacy default stream
    {
        const int cy = (Cin + 31) / 32 < 8 ? (Cin + 31) / 32 : 8;
        conv_op_input_kernel<<<dim3(T * H, cy), 256, 32 * (W + 1) * sizeof(float), s>>>(
            x, a.as<__nv_bfloat16>(), Cin, T, H, W, tf);
        LTXV_CUDA(cudaGetLastError());
    }
    LTXV_CUDA(launch_conv_weight_relayout(weight, 0, wb.p, Cout, Cin, rows_out, 0, s));
    if (bias != nullptr) LTXV_CUDA(launch_conv_bias_relayout(bias, 0, bb.as<float>(), Cout, rows_out, 0, s));
    GemmOperands ops{a.p, static_cast<int64_t>(T + 2) * plane, Cin, wb.p, rows_out, 27ll * Cin};
    GemmParams p{};
    p.M = T * plane;
    p.N = Cout;
    p.K = 27 * Cin;
    p.num_k_blocks = 27 * (Cin / 64);
    p.epi = EPI_CONV_NDHWC;
    p.ldo = Cout;
    p.bias = bb.as<float>();
    p.out = y.p;
    p.conv = 1;
    p.cin_blocks = Cin / 64;
    p.T = T;
    p.H = H;
    p.W = W;
    p.cin = Cin;
    p.a_ptr = a.p;
    // output frame t reads padded frames t, t+1, t+2 in both padding modes (causal: frames t-2..t of the input)
    for (int kt = 0; kt < 3; ++kt)
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) p.tap_off[(kt * 3 + kh) * 3 + kw] = kt * plane + (kh - 1) * Wp + (kw - 1);
    LTXV_CUDA(launch_gemm_bf16(ops, p, 0, s));
    conv_op_output_kernel<<<dim3(static_cast<unsigned>((nvox + 31) / 32), (Cout + 31) / 32), dim3(32, 8), 0, s>>>(
        y.as<__nv_bfloat16>(), out, nvox, Cout);
    LTXV_CUDA(cudaGetLastError());
    g_conv_op_launches.fetch_add(2, std::memory_order_relaxed);
    LTXV_CUDA(cudaStreamSynchronize(s));  // the staging buffers die with this frame
}

}  // namespace ltxv
