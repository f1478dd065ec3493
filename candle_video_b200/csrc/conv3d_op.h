// Standalone LtxVideoCausalConv3d::forward (vae.rs:298-465): kernel 3x3x3, stride 1, dilation 1, on NCDHW f32 tensors.
#pragma once
#include "model_common.h"

namespace ltxv {

// x [1, Cin, T, H, W] f32, weight [Cout, Cin, 3, 3, 3] f32, bias [Cout] f32 or null (all device) -> out
// [1, Cout, T, H, W] f32.  Temporal padding: is_causal ? two copies of frame 0 in front (vae.rs:383-387)
// : one replicated frame each side (vae.rs:388-411); H/W zero padding 1 (vae.rs:344).  Operands are rounded to bf16,
// accumulation is f32 over all 27*Cin products, the result is rounded to bf16 once (like the decoder's convs).
// Synchronises `s` before returning (staging buffers are per call).
void causal_conv3d(const float* x, const float* weight, const float* bias, int Cin, int Cout, int T, int H, int W,
                   bool is_causal, float* out, cudaStream_t s);

uint64_t conv_op_launch_count();

}  // namespace ltxv
