#include "model_common.h"

#include <atomic>
#include <mutex>

namespace ltxv {

namespace {
std::atomic<uint64_t> g_common_launches{0};

__device__ __forceinline__ uint32_t hash32(uint64_t i, uint64_t seed) {
    uint64_t z = i + seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return static_cast<uint32_t>(z >> 16);
}
__device__ __forceinline__ float u01(uint32_t h) { return (static_cast<float>(h >> 8) + 0.5f) * (1.0f / 16777216.0f); }

template <typename T>
__device__ __forceinline__ void put(T* dst, int64_t i, float v);
template <>
__device__ __forceinline__ void put<float>(float* dst, int64_t i, float v) { dst[i] = v; }
template <>
__device__ __forceinline__ void put<__nv_bfloat16>(__nv_bfloat16* dst, int64_t i, float v) {
    dst[i] = __float2bfloat16(v);
}

template <typename T>
__global__ void fill_uniform_kernel(T* dst, int64_t n, float bound, uint64_t seed) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i < n) put<T>(dst, i, (u01(hash32(i, seed)) * 2.0f - 1.0f) * bound);
}
template <typename T>
__global__ void fill_normal_kernel(T* dst, int64_t n, float mean, float std, uint64_t seed) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const float a = u01(hash32(2 * i, seed)), b = u01(hash32(2 * i + 1, seed));
    put<T>(dst, i, mean + std * sqrtf(-2.0f * logf(a)) * cosf(6.283185307179586f * b));
}
template <typename T>
__global__ void fill_const_kernel(T* dst, int64_t n, float v) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i < n) put<T>(dst, i, v);
}
template <typename TS, typename TD>
__global__ void convert_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int64_t n) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i < n) put<TD>(dst, i, static_cast<float>(src[i]));
}
inline int grid_for(int64_t n) { return static_cast<int>((n + 255) / 256); }

DevBuf g_stage[64];  // staging buffer for host -> device ingestion, one per device (a process may hold models on several GPUs)
std::mutex g_stage_mu;
}  // namespace

uint64_t common_launch_count() { return g_common_launches.load(); }

std::string& last_error_ref() {
    thread_local std::string err;
    return err;
}
void set_error(const char* fmt, ...) {
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
}
void fail(const char* fmt, ...) {
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Error(buf);
}

void require_cuda_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        fail("libltxv_b200 requires an NVIDIA B200 (sm_100a) GPU: no CUDA device available (%s); there is no CPU fallback",
             e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) fail("invalid CUDA device index %d (have %d)", device, n);
    cudaDeviceProp prop;
    LTXV_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        fail("libltxv_b200 is built for sm_100a only; device %d is sm_%d%d (%s)", device, prop.major, prop.minor, prop.name);
    LTXV_CUDA(cudaSetDevice(device));
}

void ingest_tensor(void* dst, bool dst_bf16, const void* src, int src_dtype, int64_t n) {
    if (n <= 0) return;
    if (src_dtype != LTXV_F32 && src_dtype != LTXV_BF16) fail("unsupported dtype code %d", src_dtype);
    const size_t esz = src_dtype == LTXV_F32 ? 4 : 2;
    cudaPointerAttributes attr{};
    bool on_device = false;
    if (cudaPointerGetAttributes(&attr, src) == cudaSuccess)
        on_device = (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
    else
        cudaGetLastError();
    std::lock_guard<std::mutex> lk(g_stage_mu);
    const void* dsrc = src;
    if (!on_device) {
        int dev = 0;
        LTXV_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64) fail("device index %d out of range", dev);
        DevBuf& stage = g_stage[dev];
        stage.ensure(n * esz);
        LTXV_CUDA(cudaMemcpy(stage.p, src, n * esz, cudaMemcpyHostToDevice));
        dsrc = stage.p;
    }
    if (src_dtype == LTXV_F32 && dst_bf16)
        convert_kernel<float, __nv_bfloat16><<<grid_for(n), 256>>>(static_cast<const float*>(dsrc),
                                                                  static_cast<__nv_bfloat16*>(dst), n);
    else if (src_dtype == LTXV_F32)
        convert_kernel<float, float><<<grid_for(n), 256>>>(static_cast<const float*>(dsrc), static_cast<float*>(dst), n);
    else if (dst_bf16)
        convert_kernel<__nv_bfloat16, __nv_bfloat16><<<grid_for(n), 256>>>(static_cast<const __nv_bfloat16*>(dsrc),
                                                                          static_cast<__nv_bfloat16*>(dst), n);
    else
        convert_kernel<__nv_bfloat16, float><<<grid_for(n), 256>>>(static_cast<const __nv_bfloat16*>(dsrc),
                                                                  static_cast<float*>(dst), n);
    g_common_launches++;
    LTXV_CUDA(cudaGetLastError());
    LTXV_CUDA(cudaDeviceSynchronize());
}

void fill_uniform(void* dst, bool dst_bf16, int64_t n, float bound, uint64_t seed) {
    if (dst_bf16)
        fill_uniform_kernel<__nv_bfloat16><<<grid_for(n), 256>>>(static_cast<__nv_bfloat16*>(dst), n, bound, seed);
    else
        fill_uniform_kernel<float><<<grid_for(n), 256>>>(static_cast<float*>(dst), n, bound, seed);
    g_common_launches++;
    LTXV_CUDA(cudaGetLastError());
}
void fill_normal(void* dst, bool dst_bf16, int64_t n, float mean, float std, uint64_t seed) {
    if (dst_bf16)
        fill_normal_kernel<__nv_bfloat16><<<grid_for(n), 256>>>(static_cast<__nv_bfloat16*>(dst), n, mean, std, seed);
    else
        fill_normal_kernel<float><<<grid_for(n), 256>>>(static_cast<float*>(dst), n, mean, std, seed);
    g_common_launches++;
    LTXV_CUDA(cudaGetLastError());
}
void fill_const(void* dst, bool dst_bf16, int64_t n, float v) {
    if (dst_bf16)
        fill_const_kernel<__nv_bfloat16><<<grid_for(n), 256>>>(static_cast<__nv_bfloat16*>(dst), n, v);
    else
        fill_const_kernel<float><<<grid_for(n), 256>>>(static_cast<float*>(dst), n, v);
    g_common_launches++;
    LTXV_CUDA(cudaGetLastError());
}

}  // namespace ltxv
