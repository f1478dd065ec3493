// Host-side launch helper: cudaLaunchKernelEx with programmatic stream serialisation (PDL), see common.cuh.
// Only kernels that call griddep_wait() before touching global memory may be launched through it.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

#include <utility>

#include "options.h"

namespace ltxv {

// LTXV_NO_PDL is a bit mask: 1 = no PDL anywhere, 2 = not for the GEMM / conv3d kernels, 4 = not for the DiT glue,
// 8 = not for attention, 16 = not for the VAE glue.  A translation unit names its class with LTXV_PDL_CLASS before
// including this header.
#ifndef LTXV_PDL_CLASS
#define LTXV_PDL_CLASS 0
#endif
static inline bool pdl_enabled() {
    const int m = options().no_pdl;
    return !(m & 1) && !(m & LTXV_PDL_CLASS);
}

// Function attributes (cudaFuncSetAttribute: opt-in shared memory) belong to the device that was current when they
// were set.  One of these per launcher remembers, per device, whether its kernel has been configured there, so a
// process may hold models on several GPUs.
struct PerDeviceOnce {
    bool done[64] = {};
    // true if the kernel still has to be configured on the current device (returned in *dev)
    bool need(int* dev) {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) d = 0;
        *dev = d;
        return !done[d];
    }
    void mark(int dev) { done[dev] = true; }
};

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

}  // namespace ltxv
