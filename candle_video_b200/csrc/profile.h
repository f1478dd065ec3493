// Optional per-kernel-class device timing (CUDA events on the launching stream) used by bench.py to compute the
// roofline of the dominant kernels inside a real step.  Off by default: zero overhead on the product path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ltxv {

// classes 4-7 are HBM-bound glue kernels: their `flops` field carries ALGORITHMIC BYTES instead
enum ProfClass : int {
    PROF_GEMM = 0, PROF_CONV = 1, PROF_ATTN_SELF = 2, PROF_ATTN_CROSS = 3,
    PROF_NORM_MOD = 4, PROF_QK_ROPE = 5, PROF_VAE_PREP = 6, PROF_OTHER = 7, PROF_NUM = 8
};

bool profiling_enabled();
void profiling_begin();
// returns, per class: launches, total milliseconds, total algorithmic FLOPs
void profiling_end(uint64_t* launches, double* ms, double* flops);

// Kernel-variant trace (test support): while on, every launcher records the name of the kernel variant it picked
// (template instance, tile mode, split count ...) so parity tests can assert that the production variants ran.
void trace_begin();
// "name count\n" per distinct variant, sorted by name; tracing is switched off
const char* trace_end();
bool tracing_enabled();
void trace_variant_slow(const char* fmt, ...);
#define LTXV_TRACE_VARIANT(...)                                   \
    do {                                                          \
        if (::ltxv::tracing_enabled()) ::ltxv::trace_variant_slow(__VA_ARGS__); \
    } while (0)

struct ProfScope {
    bool on;
    int cls;
    double flops;
    cudaStream_t s;
    cudaEvent_t e0, e1;
    ProfScope(int cls, double flops, cudaStream_t s);
    ~ProfScope();
};

}  // namespace ltxv
