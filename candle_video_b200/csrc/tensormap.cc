#include "tensormap.h"

#include <mutex>
#include <stdio.h>
#include <string>

namespace ltxv {

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_once;
thread_local std::string g_err;

void resolve() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = reinterpret_cast<EncodeTiledFn>(fn);
}
}  // namespace

const char* tensor_map_last_error() { return g_err.c_str(); }

cudaError_t make_tensor_map_2d_bf16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int box_rows,
                                    int box_cols, int64_t row_stride_elems) {
    std::call_once(g_once, resolve);
    if (g_encode == nullptr) {
        g_err = "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)";
        return cudaErrorNotSupported;
    }
    if (row_stride_elems < 0) row_stride_elems = cols;
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (row_stride_elems * 2) % 16 != 0) {
        g_err = "tensor map: base and row stride must be 16-byte aligned";
        return cudaErrorInvalidValue;
    }
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(row_stride_elems) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[256];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed: CUresult %d (rows=%lld cols=%lld box=%dx%d)", (int)r,
                 (long long)rows, (long long)cols, box_rows, box_cols);
        g_err = buf;
        return cudaErrorInvalidValue;
    }
    return cudaSuccess;
}

cudaError_t make_tensor_map_3d_bf16(CUtensorMap* out, const void* base, int64_t batch, int64_t rows, int64_t cols,
                                    int box_rows, int box_cols, int64_t row_stride_elems, int64_t batch_stride_elems) {
    std::call_once(g_once, resolve);
    if (g_encode == nullptr) {
        g_err = "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)";
        return cudaErrorNotSupported;
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (row_stride_elems * 2) % 16 != 0 ||
        (batch_stride_elems * 2) % 16 != 0) {
        g_err = "tensor map: base and strides must be 16-byte aligned";
        return cudaErrorInvalidValue;
    }
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(batch)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(row_stride_elems) * 2, static_cast<cuuint64_t>(batch_stride_elems) * 2};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[256];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(3d) failed: CUresult %d (batch=%lld rows=%lld cols=%lld)", (int)r,
                 (long long)batch, (long long)rows, (long long)cols);
        g_err = buf;
        return cudaErrorInvalidValue;
    }
    return cudaSuccess;
}

}  // namespace ltxv
