// Host-side TMA descriptor (CUtensorMap) construction. The driver entry point is resolved at run time through
// the CUDA runtime, so the library has no link-time dependency on libcuda (it must also build on a CPU box).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ltxv {

// 2D bf16 row-major tensor [rows, cols] (cols contiguous, row stride = row_stride_elems),
// box = [box_rows, box_cols], 128-byte swizzle (box_cols * 2 must be <= 128), zero fill out of bounds.
cudaError_t make_tensor_map_2d_bf16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int box_rows,
                                    int box_cols, int64_t row_stride_elems = -1);

// 3D bf16 tensor [batch, rows, cols] (cols contiguous), box = [1, box_rows, box_cols], 128-byte swizzle.
// Out-of-bounds rows of one batch entry are zero-filled instead of running into the next entry.
cudaError_t make_tensor_map_3d_bf16(CUtensorMap* out, const void* base, int64_t batch, int64_t rows, int64_t cols,
                                    int box_rows, int box_cols, int64_t row_stride_elems, int64_t batch_stride_elems);

const char* tensor_map_last_error();

}  // namespace ltxv
