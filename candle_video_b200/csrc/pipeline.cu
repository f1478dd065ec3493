// Host-side mirror of the hot part of `LtxPipeline::call` (t2v_pipeline.rs:627-1073): schedule math (kept as in the
// reference, host f32), the denoise loop with sequential CFG / STG passes, and the decode branch.
#include "pipeline.h"

#include <math.h>

#include "glue.h"
#include "options.h"

namespace ltxv {

// t2v_pipeline.rs:159-169 with the scheduler's static SchedulerConfig {256, 4096, 0.5, 1.15} (scheduler.rs:615-640)
float calculate_shift(int seq_len) {
    const float base_shift = 0.5f, max_shift = 1.15f;
    const float m = (max_shift - base_shift) / static_cast<float>(4096 - 256);
    const float b = base_shift - m * static_cast<float>(256);
    return static_cast<float>(seq_len) * m + b;
}

// FlowMatchEulerDiscreteScheduler::set_timesteps as driven by the pipeline (scheduler.rs:274-412, :646-660)
void scheduler_set_timesteps(int n, const float* custom_sigmas, float mu, bool has_terminal, float terminal,
                             float* sigmas_out, int64_t* timesteps_out) {
    if (n <= 0) fail("num_inference_steps must be positive");
    std::vector<float> s(n);
    if (custom_sigmas != nullptr) {
        for (int i = 0; i < n; ++i) s[i] = custom_sigmas[i];
    } else {
        // t2v_pipeline.rs:752-757, linspace(1, 1/n, n) (:171-182)
        const float start = 1.0f, end = 1.0f / static_cast<float>(n);
        if (n == 1) {
            s[0] = start;
        } else {
            const float denom = static_cast<float>(n - 1);
            for (int i = 0; i < n; ++i) s[i] = start + (end - start) * static_cast<float>(i) / denom;
        }
    }
    // exponential time shift, sigma exponent 1.0 (scheduler.rs:172-179, :341-346)
    const float emu = expf(mu);
    for (int i = 0; i < n; ++i) {
        const float base = powf(1.0f / s[i] - 1.0f, 1.0f);
        s[i] = emu / (emu + base);
    }
    if (has_terminal) {  // stretch_shift_to_terminal_vec (scheduler.rs:188-207)
        const float one_minus_last = 1.0f - s[n - 1];
        const float denom = 1.0f - terminal;
        if (fabsf(denom) < 1e-12f) fail("shift_terminal too close to 1.0");
        const float scale = one_minus_last / denom;
        for (int i = 0; i < n; ++i) s[i] = 1.0f - ((1.0f - s[i]) / scale);
    }
    for (int i = 0; i < n; ++i) {
        sigmas_out[i] = s[i];
        // `x as i64` (scheduler.rs:658-659): truncation; Rust maps NaN to 0 (n = 1 with a terminal stretch is 0/0)
        const float tv = s[i] * 1000.0f;
        timesteps_out[i] = isnan(tv) ? 0 : static_cast<int64_t>(tv);
    }
    sigmas_out[n] = 0.0f;
}

namespace {
// check_inputs (t2v_pipeline.rs:323-327) + the parameter checks shared by the single- and multi-GPU loops
struct LoopGeom {
    int F, H, W, S, C;
    bool do_cfg, do_stg;
};
LoopGeom check_loop_params(const LtxVideoTransformer3DModel& dit, const ltxv_pipeline_params& p, const void* negative) {
    if (p.height % 32 != 0 || p.width % 32 != 0)
        fail("`height` and `width` must be divisible by 32, got %d and %d", p.height, p.width);
    if (p.num_frames < 1 || p.frame_rate < 1) fail("num_frames and frame_rate must be positive");
    if (p.num_inference_steps <= 0) fail("num_inference_steps must be positive");
    LoopGeom g{};
    g.F = (p.num_frames - 1) / 8 + 1;  // :743-747
    g.H = p.height / 32;
    g.W = p.width / 32;
    g.S = g.F * g.H * g.W;
    g.C = dit.config().in_channels;
    g.do_cfg = p.guidance_scale > 1.0f;  // :308-310
    g.do_stg = p.stg_scale > 0.0f;       // :304-306
    if (g.do_cfg && negative == nullptr) fail("negative prompt embeddings are required when guidance_scale > 1");
    if (p.num_skip_blocks < 0 || (p.num_skip_blocks > 0 && p.skip_block_list == nullptr))
        fail("skip_block_list is null but num_skip_blocks = %d", p.num_skip_blocks);
    return g;
}

// [num_layers] STG mask, 1 = skip (:911-923)
std::vector<float> make_stg_mask(const LtxVideoTransformer3DModel& dit, const ltxv_pipeline_params& p) {
    std::vector<float> m(dit.config().num_layers, 0.0f);
    for (int i = 0; i < p.num_skip_blocks; ++i)
        if (p.skip_block_list[i] >= 0 && p.skip_block_list[i] < dit.config().num_layers) m[p.skip_block_list[i]] = 1.0f;
    return m;
}

// integer-truncated timesteps as f32 (Tensor::full(t as f32), :874) -> device, through the handle's pinned buffer
void upload_timesteps(PipeWs& w, const std::vector<int64_t>& tsteps, cudaStream_t s) {
    const int n = static_cast<int>(tsteps.size());
    w.ts.ensure(static_cast<size_t>(n) * 4);
    float* h = w.host.acquire(n);
    for (int i = 0; i < n; ++i) h[i] = static_cast<float>(tsteps[i]);
    LTXV_CUDA(cudaMemcpyAsync(w.ts.p, h, n * 4, cudaMemcpyHostToDevice, s));
    w.host.copied(s);
}
}  // namespace

void pipeline_denoise(LtxVideoTransformer3DModel& dit, const ltxv_pipeline_params& p, float* latents,
                      const void* prompt, const float* prompt_mask, const void* negative, const float* negative_mask,
                      int embeds_dtype, int K, cudaStream_t s, const float* step_noise) {
    const LoopGeom g = check_loop_params(dit, p, negative);
    const int F = g.F, H = g.H, W = g.W, S = g.S, C = g.C;
    const bool do_cfg = g.do_cfg, do_stg = g.do_stg;
    LTXV_CUDA(cudaSetDevice(dit.device()));

    // skip-block policy (:691-697)
    if (p.skip_block_list != nullptr) {
        if (!do_stg) dit.set_skip_block_list(p.skip_block_list, p.num_skip_blocks);
        else dit.set_skip_block_list(nullptr, 0);
    }
    const int n = p.num_inference_steps;
    std::vector<float> sigmas(n + 1);
    std::vector<int64_t> tsteps(n);
    const float mu = p.custom_sigmas ? 0.0f : calculate_shift(S);  // :762-773
    scheduler_set_timesteps(n, p.custom_sigmas, mu, p.has_shift_terminal != 0, p.shift_terminal, sigmas.data(),
                            tsteps.data());

    PipeWs& w = dit.pipe_ws();
    const size_t out_bytes = static_cast<size_t>(S) * C * 4;
    w.cond.ensure(out_bytes);
    if (do_cfg) w.uncond.ensure(out_bytes);
    if (do_stg) w.pert.ensure(out_bytes);
    if (step_noise != nullptr) w.comb.ensure(out_bytes);
    w.coords.ensure(static_cast<size_t>(S) * 3 * 4);
    w.scratch.ensure(64);
    upload_timesteps(w, tsteps, s);
    LTXV_CUDA(launch_video_coords(w.coords.as<float>(), F, H, W, 8, 32, p.frame_rate, s));  // :798-847

    dit.prepare_context(0, prompt, embeds_dtype, prompt_mask, K, s);
    if (do_cfg) dit.prepare_context(1, negative, embeds_dtype, negative_mask, K, s);
    // The reference runs the CFG branches as sequential B = 1 forwards (:878-907).  They share latents, timestep and
    // coordinates, so here they run as ONE forward over 2S tokens (uncond rows first): same per-row arithmetic, better
    // tile occupancy, weights streamed once.  LTXV_NO_CFG_BATCH=1 restores the sequential order.
    const bool batch_cfg = !options().no_cfg_batch;
    const bool pair = do_cfg && batch_cfg;
    float* out_uncond = do_cfg ? w.uncond.as<float>() : nullptr;
    float* out_cond = w.cond.as<float>();
    if (pair) {
        dit.prepare_pair(1, 0, s);
        w.pair.ensure(2 * out_bytes);
        out_uncond = w.pair.as<float>();
        out_cond = w.pair.as<float>() + static_cast<size_t>(S) * C;
    }

    std::vector<float> stg_mask;
    if (do_stg) stg_mask = make_stg_mask(dit, p);
    const float* coords = w.coords.as<float>();
    for (int i = 0; i < n; ++i) {
        const float* t_dev = w.ts.as<float>() + i;
        if (pair) {
            dit.forward_pair(latents, LTXV_F32, t_dev, S, F, H, W, nullptr, coords, w.pair.p, LTXV_F32, s);
        } else {
            if (do_cfg)
                dit.forward_ctx(1, latents, LTXV_F32, t_dev, S, F, H, W, nullptr, coords, nullptr, 1, out_uncond, LTXV_F32, s);
            dit.forward_ctx(0, latents, LTXV_F32, t_dev, S, F, H, W, nullptr, coords, nullptr, 1, out_cond, LTXV_F32, s);
        }
        if (do_stg)
            dit.forward_ctx(0, latents, LTXV_F32, t_dev, S, F, H, W, nullptr, coords, stg_mask.data(), 1, w.pert.p,
                            LTXV_F32, s);
        const float dt = sigmas[i + 1] - sigmas[i];  // scheduler.rs:544-549
        if (step_noise == nullptr) {
            LTXV_CUDA(launch_guidance_euler(out_cond, out_uncond,
                                            do_stg ? w.pert.as<float>() : nullptr, latents, nullptr,
                                            static_cast<int64_t>(S) * C, p.guidance_scale, p.guidance_rescale,
                                            p.stg_scale, dt, w.scratch.as<double>(), s));
        } else {
            // stochastic_sampling = true (scheduler.rs:557-575; preset 0.9.8-distilled, configs.rs:210): the combined
            // velocity is materialised (in place of the conditional output), then x <- (1-s')(x - s v) + s' noise_i
            LTXV_CUDA(launch_guidance_euler(out_cond, out_uncond,
                                            do_stg ? w.pert.as<float>() : nullptr, nullptr, w.comb.as<float>(),
                                            static_cast<int64_t>(S) * C, p.guidance_scale, p.guidance_rescale,
                                            p.stg_scale, dt, w.scratch.as<double>(), s));
            LTXV_CUDA(launch_stochastic_step(latents, w.comb.as<float>(), step_noise + static_cast<size_t>(i) * S * C,
                                             sigmas[i], sigmas[i + 1], static_cast<int64_t>(S) * C, s));
        }
    }
}

ParallelPlan make_parallel_plan(int nranks, int rank, int S, bool do_cfg) {
    ParallelPlan pl{};
    pl.cfg_groups = (do_cfg && nranks % 2 == 0) ? 2 : 1;
    pl.sp = nranks / pl.cfg_groups;
    if (S % pl.sp != 0) fail("sequence length %d is not divisible by the sequence-parallel size %d", S, pl.sp);
    pl.branch = rank / pl.sp;   // 0 = unconditional (or the only branch), 1 = conditional
    pl.sp_rank = rank % pl.sp;
    pl.s_local = S / pl.sp;
    pl.token0 = pl.sp_rank * pl.s_local;
    return pl;
}

void pipeline_denoise_parallel(LtxVideoTransformer3DModel& dit, PeerComm& comm, const ltxv_pipeline_params& p,
                               float* latents, const void* prompt, const float* prompt_mask, const void* negative,
                               const float* negative_mask, int embeds_dtype, int K, cudaStream_t s,
                               const float* step_noise) {
    const LoopGeom g = check_loop_params(dit, p, negative);
    const int F = g.F, H = g.H, W = g.W, S = g.S, C = g.C;
    const bool do_cfg = g.do_cfg, do_stg = g.do_stg;
    LTXV_CUDA(cudaSetDevice(dit.device()));
    const int N = comm.nranks(), rank = comm.rank();
    const ParallelPlan pl = make_parallel_plan(N, rank, S, do_cfg);
    const bool split = pl.cfg_groups == 2;
    if (p.skip_block_list != nullptr) {
        if (!do_stg) dit.set_skip_block_list(p.skip_block_list, p.num_skip_blocks);
        else dit.set_skip_block_list(nullptr, 0);
    }
    dit.set_comm(&comm, pl.branch * pl.sp, pl.sp);

    const int n = p.num_inference_steps;
    std::vector<float> sigmas(n + 1);
    std::vector<int64_t> tsteps(n);
    const float mu = p.custom_sigmas ? 0.0f : calculate_shift(S);
    scheduler_set_timesteps(n, p.custom_sigmas, mu, p.has_shift_terminal != 0, p.shift_terminal, sigmas.data(),
                            tsteps.data());
    PipeWs& w = dit.pipe_ws();
    const size_t loc_elems = static_cast<size_t>(pl.s_local) * C;
    w.coords.ensure(static_cast<size_t>(S) * 3 * 4);
    w.scratch.ensure(64);
    // symmetric buffers: branch outputs [3][s_local, C] f32 (uncond, cond, perturbed), the gathered latents [S, C] and the
    // std partial sums.  Carved once per (communicator, S, shard count, C): a call with another geometry gets its own
    // buffers (same loc_elems does not imply the same S: CFG on 2 ranks vs no CFG on 2 ranks).
    const std::vector<int64_t> key = {static_cast<int64_t>(comm.id()), S, pl.sp, C, N};
    auto& cache = dit.pipe_allocs();
    auto it = cache.find(key);
    if (it == cache.end()) {
        std::vector<size_t> offs(3);
        offs[0] = comm.alloc(3 * loc_elems * 4);
        offs[1] = comm.alloc(static_cast<size_t>(S) * C * 4);
        offs[2] = comm.alloc(4 * sizeof(double));
        it = cache.emplace(key, std::move(offs)).first;
    }
    const size_t xchg_off = it->second[0], lat_off = it->second[1], stat_off = it->second[2];
    float* x_unc = static_cast<float*>(comm.local(xchg_off));
    float* x_cond = x_unc + loc_elems;
    float* x_pert = x_cond + loc_elems;
    float* lat_loc = latents + static_cast<size_t>(pl.token0) * C;  // this rank updates only its token shard

    upload_timesteps(w, tsteps, s);
    LTXV_CUDA(launch_video_coords(w.coords.as<float>(), F, H, W, 8, 32, p.frame_rate, s));
    const float* coords_loc = w.coords.as<float>() + static_cast<size_t>(pl.token0) * 3;

    // text contexts: only the branches this rank runs
    const bool run_uncond = do_cfg && (!split || pl.branch == 0);
    const bool run_cond = !split || pl.branch == 1 || !do_cfg;
    if (run_cond) dit.prepare_context(0, prompt, embeds_dtype, prompt_mask, K, s);
    if (run_uncond) dit.prepare_context(1, negative, embeds_dtype, negative_mask, K, s);
    std::vector<float> stg_mask;
    if (do_stg) stg_mask = make_stg_mask(dit, p);
    const int partner = split ? (rank + pl.sp) % N : rank;
    const bool rescale = do_cfg && p.guidance_rescale > 0.0f;
    for (int i = 0; i < n; ++i) {
        const float* t_dev = w.ts.as<float>() + i;
        if (run_uncond)
            dit.forward_ctx(1, lat_loc, LTXV_F32, t_dev, pl.s_local, F, H, W, nullptr, coords_loc, nullptr, 1, x_unc,
                            LTXV_F32, s);
        if (run_cond) {
            dit.forward_ctx(0, lat_loc, LTXV_F32, t_dev, pl.s_local, F, H, W, nullptr, coords_loc, nullptr, 1, x_cond,
                            LTXV_F32, s);
            if (do_stg)
                dit.forward_ctx(0, lat_loc, LTXV_F32, t_dev, pl.s_local, F, H, W, nullptr, coords_loc, stg_mask.data(),
                                1, x_pert, LTXV_F32, s);
        }
        if (split) {
            // hand this branch's velocity shard to the partner rank of the other CFG group (same token shard)
            float* px = static_cast<float*>(comm.peer(partner, xchg_off));
            if (pl.branch == 0) {
                LTXV_CUDA(cudaMemcpyAsync(px, x_unc, loc_elems * 4, cudaMemcpyDeviceToDevice, s));
            } else {
                LTXV_CUDA(cudaMemcpyAsync(px + loc_elems, x_cond, loc_elems * 4, cudaMemcpyDeviceToDevice, s));
                if (do_stg)
                    LTXV_CUDA(cudaMemcpyAsync(px + 2 * loc_elems, x_pert, loc_elems * 4, cudaMemcpyDeviceToDevice, s));
            }
            comm.barrier(s, 0);
        }
        const float dt = sigmas[i + 1] - sigmas[i];
        StatParts parts{};
        if (rescale) {
            // unbiased std over the WHOLE tensor (t2v_pipeline.rs:209-224): partial sums per token shard, summed by
            // every rank of the sequence-parallel group after a group barrier
            double* acc = static_cast<double*>(comm.local(stat_off));
            LTXV_CUDA(launch_guidance_stats(x_cond, x_unc, static_cast<int64_t>(loc_elems), p.guidance_scale, acc, s));
            if (pl.sp > 1) comm.barrier(s, 1, pl.branch * pl.sp, pl.sp);
            for (int k = 0; k < pl.sp; ++k)
                parts.p[k] = static_cast<const double*>(comm.peer(pl.branch * pl.sp + k, stat_off));
            parts.n = pl.sp;
            parts.n_total = static_cast<int64_t>(S) * C;
        }
        if (step_noise == nullptr) {
            LTXV_CUDA(launch_guidance_euler_parts(x_cond, do_cfg ? x_unc : nullptr, do_stg ? x_pert : nullptr, lat_loc,
                                                  nullptr, static_cast<int64_t>(loc_elems), p.guidance_scale,
                                                  p.guidance_rescale, p.stg_scale, dt, parts, s));
        } else {
            // stochastic_sampling = true (scheduler.rs:557-575): combined velocity of this token shard, then
            // x <- (1-s')(x - s v) + s' noise_i on the shard's rows of the caller's [n, S, C] noise tensor
            w.comb.ensure(loc_elems * 4);
            LTXV_CUDA(launch_guidance_euler_parts(x_cond, do_cfg ? x_unc : nullptr, do_stg ? x_pert : nullptr, nullptr,
                                                  w.comb.as<float>(), static_cast<int64_t>(loc_elems), p.guidance_scale,
                                                  p.guidance_rescale, p.stg_scale, dt, parts, s));
            LTXV_CUDA(launch_stochastic_step(lat_loc, w.comb.as<float>(),
                                             step_noise + (static_cast<size_t>(i) * S + pl.token0) * C, sigmas[i],
                                             sigmas[i + 1], static_cast<int64_t>(loc_elems), s));
        }
        // partners must not overwrite the exchange buffers / partial sums before they are consumed
        if (split) comm.barrier(s, 0);
        else if (rescale && pl.sp > 1) comm.barrier(s, 1, pl.branch * pl.sp, pl.sp);
    }
    // all-gather the latent shards: every rank stores its shard into every rank's gathered buffer
    if (N > 1) {
        for (int r = 0; r < N; ++r) {
            float* dst = static_cast<float*>(comm.peer(r, lat_off)) + static_cast<size_t>(pl.token0) * C;
            if (pl.branch == 0 || !split)
                LTXV_CUDA(cudaMemcpyAsync(dst, lat_loc, loc_elems * 4, cudaMemcpyDeviceToDevice, s));
        }
        comm.barrier(s, 0);
        LTXV_CUDA(cudaMemcpyAsync(latents, comm.local(lat_off), static_cast<size_t>(S) * C * 4, cudaMemcpyDeviceToDevice, s));
        comm.barrier(s, 0);
    }
    dit.set_comm(nullptr, 0, 1);
}

void pipeline_decode(AutoencoderKLLtxVideo& vae, const ltxv_pipeline_params& p, const float* latents, float* out,
                     cudaStream_t s, const float* decode_noise, float decode_noise_scale) {
    if (p.height % 32 != 0 || p.width % 32 != 0)
        fail("`height` and `width` must be divisible by 32, got %d and %d", p.height, p.width);
    const int F = (p.num_frames - 1) / 8 + 1, H = p.height / 32, W = p.width / 32;
    const int C = vae.config().latent_channels;
    const int64_t n = static_cast<int64_t>(C) * F * H * W;
    LTXV_CUDA(cudaSetDevice(vae.device()));
    PipeWs& w = vae.pipe_ws();
    w.unpacked.ensure(n * 4);
    w.denorm.ensure(n * 4);
    w.tdec.ensure(4);
    LTXV_CUDA(launch_unpack_latents(latents, w.unpacked.as<float>(), C, F, H, W, 1, 1, s));  // :1002-1009
    const float inv_sf = 1.0f / vae.config().scaling_factor;
    LTXV_CUDA(launch_denormalize(w.unpacked.as<float>(), w.denorm.as<float>(), vae.latents_mean(), vae.latents_std(),
                                 inv_sf, C, static_cast<int64_t>(F) * H * W, s));  // :1011-1016
    const float* t_dev = nullptr;
    if (vae.config().timestep_conditioning) {
        float* h = w.host.acquire(1);
        h[0] = p.decode_timestep;
        LTXV_CUDA(cudaMemcpyAsync(w.tdec.p, h, 4, cudaMemcpyHostToDevice, s));
        w.host.copied(s);
        t_dev = w.tdec.as<float>();
        // decode-noise blend (:1021-1065): only a timestep-conditioned VAE takes the noise branch in the reference;
        // the noise tensor is the caller's (the library has no RNG)
        if (decode_noise != nullptr && decode_noise_scale != 0.0f)
            LTXV_CUDA(launch_noise_blend(w.denorm.as<float>(), decode_noise, decode_noise_scale,
                                         static_cast<int64_t>(C) * F * H * W, s));
    }
    vae.decode(w.denorm.p, LTXV_F32, t_dev, 1, F, H, W, out, LTXV_F32, /*postprocess=*/1, s);  // :1069-1070
}

}  // namespace ltxv
