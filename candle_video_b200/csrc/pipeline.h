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
                      int embeds_dtype, int K, cudaStream_t s, const float* step_noise = nullptr);
// Multi-GPU denoise loop over all ranks of `comm` (one process per GPU).  With CFG (guidance_scale > 1) and an even
// rank count the ranks split into two groups, one per CFG branch (the reference runs the branches as independent B=1
// forwards, t2v_pipeline.rs:878-907); inside a group the tokens are sharded Ulysses-style.  Each rank keeps only its
// token shard of the latents; branch outputs are exchanged with the partner rank through peer memory and both
// partners apply the same fused combine + Euler update.  `latents` holds the full [S, C] tensor on every rank on entry
// and on exit.
struct ParallelPlan {
    int cfg_groups, sp, branch, sp_rank, s_local, token0;
};
ParallelPlan make_parallel_plan(int nranks, int rank, int S, bool do_cfg);
void pipeline_denoise_parallel(LtxVideoTransformer3DModel& dit, PeerComm& comm, const ltxv_pipeline_params& p,
                               float* latents, const void* prompt, const float* prompt_mask, const void* negative,
                               const float* negative_mask, int embeds_dtype, int K, cudaStream_t s,
                               const float* step_noise = nullptr);
// step_noise (above): [num_inference_steps, S, C] f32 or null -> stochastic sampling with caller-supplied noise.
// decode_noise: [C, F, H, W] f32 or null -> latents <- (1 - scale) latents + scale noise before the VAE.
void pipeline_decode(AutoencoderKLLtxVideo& vae, const ltxv_pipeline_params& p, const float* latents, float* out,
                     cudaStream_t s, const float* decode_noise = nullptr, float decode_noise_scale = 0.0f);

}  // namespace ltxv
