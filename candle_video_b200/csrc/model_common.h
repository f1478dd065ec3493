// Host-side helpers shared by the C++ model mirrors (dit.cu, vae.cu, pipeline.cu, ffi.cu): error plumbing, device
// buffers, tensor ingestion (host/device, f32/bf16) and synthetic random init.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ltxv.h"
#include "tensormap.h"

namespace ltxv {

// thread-local last error (ltxv_last_error)
std::string& last_error_ref();
void set_error(const char* fmt, ...);

struct Error : std::runtime_error {
    explicit Error(const std::string& s) : std::runtime_error(s) {}
};
[[noreturn]] void fail(const char* fmt, ...);

#define LTXV_CUDA(expr)                                                                                  \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            ::ltxv::fail("%s failed: %s (%s) at %s:%d", #expr, cudaGetErrorString(e__),                 \
                         ::ltxv::tensor_map_last_error(), __FILE__, __LINE__);                           \
    } while (0)

// Owning device buffer
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    // (re)allocate if too small; zero_fill clears the whole allocation when (re)allocated
    bool ensure(size_t n, bool zero_fill = false) {
        if (n <= bytes && p != nullptr) return false;
        release();
        LTXV_CUDA(cudaMalloc(&p, n));
        bytes = n;
        if (zero_fill) LTXV_CUDA(cudaMemset(p, 0, n));
        return true;
    }
    template <typename T>
    T* as() const {
        return reinterpret_cast<T*>(p);
    }
};

// Pinned host staging for small per-call parameters (timestep lists, the decode timestep) that are copied to the device
// asynchronously: the copy must not read a stack buffer after the call returns, and the entry points are enqueue-only
// (no stream synchronisation).  `fence` is recorded after each copy; the next writer waits for it (already complete in
// practice) before it overwrites the host side.
struct PinnedParams {
    float* h = nullptr;
    size_t cap = 0;  // floats
    cudaEvent_t fence = nullptr;
    PinnedParams() = default;
    PinnedParams(const PinnedParams&) = delete;
    PinnedParams& operator=(const PinnedParams&) = delete;
    ~PinnedParams() {
        if (h) cudaFreeHost(h);
        if (fence) cudaEventDestroy(fence);
    }
    // host buffer of at least n floats that no pending copy is still reading
    float* acquire(size_t n) {
        if (fence == nullptr) LTXV_CUDA(cudaEventCreateWithFlags(&fence, cudaEventDisableTiming));
        LTXV_CUDA(cudaEventSynchronize(fence));
        if (n > cap) {
            if (h) cudaFreeHost(h);
            h = nullptr;
            LTXV_CUDA(cudaMallocHost(&h, n * sizeof(float)));
            cap = n;
        }
        return h;
    }
    void copied(cudaStream_t s) { LTXV_CUDA(cudaEventRecord(fence, s)); }
};

// Workspace of the pipeline entry points (pipeline.cu).  It lives in the model handle it serves -- on that model's
// device, one per handle -- so two models (on one GPU or on several) never share scratch memory.
struct PipeWs {
    DevBuf cond, uncond, pert, comb, pair, coords, ts, scratch, unpacked, denorm, tdec;
    PinnedParams host;
};

// Fails loudly when there is no usable CUDA device (no CPU fallback exists).
void require_cuda_device(int device);

// dst (device, bf16 or f32) <- src (host or device, f32 or bf16), n elements; synchronous.
void ingest_tensor(void* dst, bool dst_bf16, const void* src, int src_dtype, int64_t n);

// Deterministic device-side init: uniform(-bound, bound) or normal(0, std) (hash RNG), bf16 or f32 destination.
void fill_uniform(void* dst, bool dst_bf16, int64_t n, float bound, uint64_t seed);
void fill_normal(void* dst, bool dst_bf16, int64_t n, float mean, float std, uint64_t seed);
void fill_const(void* dst, bool dst_bf16, int64_t n, float v);

uint64_t common_launch_count();

// One named parameter slot of a model: where a reference tensor lands on the device.
struct ParamSlot {
    std::string key;
    void* dst = nullptr;        // device destination (already offset for fused tensors)
    bool dst_bf16 = true;
    std::vector<int64_t> shape;  // expected reference shape
    bool loaded = false;
    int64_t numel() const {
        int64_t n = 1;
        for (auto v : shape) n *= v;
        return n;
    }
};

}  // namespace ltxv
