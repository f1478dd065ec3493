// Host-side mirror of the hot part of LtxPipeline::call (t2v_pipeline.rs:627-1073); see pipeline.cu.
#pragma once
#include "dit.h"
#include "vae.h"

namespace ltxv {

float calculate_shift(int seq_len);
void scheduler_set_timesteps(int n, const float* custom_sigmas, float mu, bool has_terminal, float terminal,
                             float* sigmas_out, int64_t* timesteps_out);
void pipeline_denoise(LtxVideoTransformer3DModel& dit, const ltxv_pipeline_params& p, float* latents,
                      const void* prompt, const float* prompt_mask, const void* negative, const float* negative_mask,
                      int embeds_dtype, int K, cudaStream_t s);
void pipeline_decode(AutoencoderKLLtxVideo& vae, const ltxv_pipeline_params& p, const float* latents, float* out,
                     cudaStream_t s);

}  // namespace ltxv
