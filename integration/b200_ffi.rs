//! `src/models/ltx_video/b200_ffi.rs` -- raw bindings of libltxv_b200.so (include/ltxv.h of the B200 repository).
//!
//! Every entry point of the header is declared here, one `extern "C"` item per C declaration: first what the trait
//! implementations in `b200_models.rs` call (same order as the header), then the rest of the ABI.  Dtype codes: `LTXV_F32 = 0`, `LTXV_BF16 = 1`.
//! Link with `cargo:rustc-link-lib=dylib=ltxv_b200`.
#![allow(non_camel_case_types, dead_code)]

use std::os::raw::{c_char, c_int, c_void};

pub const LTXV_F32: c_int = 0;
pub const LTXV_BF16: c_int = 1;

#[repr(C)]
pub struct ltxv_dit {
    _private: [u8; 0],
}
#[repr(C)]
pub struct ltxv_vae {
    _private: [u8; 0],
}
#[repr(C)]
pub struct ltxv_comm {
    _private: [u8; 0],
}

/// LtxVideoTransformer3DModelConfig (ltx_transformer.rs:22-59); `timestep_bf16_round` replays
/// `timestep.to_dtype(model_dtype)` (:1051) of a bf16 run.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct ltxv_dit_config {
    pub in_channels: i32,
    pub out_channels: i32,
    pub patch_size: i32,
    pub patch_size_t: i32,
    pub num_attention_heads: i32,
    pub attention_head_dim: i32,
    pub cross_attention_dim: i32,
    pub num_layers: i32,
    pub caption_channels: i32,
    pub norm_eps: f32,
    pub timestep_bf16_round: i32,
}

/// Decoder fields of AutoencoderKLLtxVideoConfig (vae.rs:30-103).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct ltxv_vae_config {
    pub latent_channels: i32,
    pub out_channels: i32,
    pub decoder_block_out_channels: [i32; 3],
    pub decoder_layers_per_block: [i32; 4],
    pub patch_size: i32,
    pub timestep_conditioning: i32,
    pub scaling_factor: f32,
}

/// Encoder fields of AutoencoderKLLtxVideoConfig; `downsample_types`: 1 spatial, 2 temporal, 3 spatiotemporal.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct ltxv_vae_encoder_config {
    pub in_channels: i32,
    pub latent_channels: i32,
    pub block_out_channels: [i32; 5],
    pub layers_per_block: [i32; 5],
    pub downsample_types: [i32; 4],
    pub patch_size: i32,
}

/// Tiling knobs of AutoencoderKLLtxVideo (vae.rs:1848-1861), sample space.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct ltxv_vae_tiling {
    pub use_tiling: i32,
    pub use_framewise_decoding: i32,
    pub tile_sample_min_height: i32,
    pub tile_sample_min_width: i32,
    pub tile_sample_min_num_frames: i32,
    pub tile_sample_stride_height: i32,
    pub tile_sample_stride_width: i32,
    pub tile_sample_stride_num_frames: i32,
}

/// Arguments of LtxPipeline::call that the denoise loop / decode branch need (t2v_pipeline.rs:627-660).
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct ltxv_pipeline_params {
    pub height: i32,
    pub width: i32,
    pub num_frames: i32,
    pub frame_rate: i32,
    pub num_inference_steps: i32,
    pub custom_sigmas: *const f32,
    pub guidance_scale: f32,
    pub guidance_rescale: f32,
    pub stg_scale: f32,
    pub skip_block_list: *const i32,
    pub num_skip_blocks: i32,
    pub has_shift_terminal: i32,
    pub shift_terminal: f32,
    pub decode_timestep: f32,
}

extern "C" {
    pub fn ltxv_last_error() -> *const c_char;
    pub fn ltxv_version() -> *const c_char;
    pub fn ltxv_launch_count() -> u64;

    // ---- weight files ----
    pub fn ltxv_remap_official_key_raw(key: *const c_char, out: *mut c_char, out_cap: u64) -> c_int;
    pub fn ltxv_remap_official_key(key: *const c_char, out: *mut c_char, out_cap: u64, component: *mut i32) -> c_int;
    pub fn ltxv_dit_load_safetensors(m: *mut ltxv_dit, path: *const c_char, official: c_int, loaded: *mut i32,
                                     ignored: *mut i32) -> c_int;
    pub fn ltxv_vae_load_safetensors(m: *mut ltxv_vae, path: *const c_char, official: c_int, loaded: *mut i32,
                                     ignored: *mut i32) -> c_int;

    // ---- DiT ----
    pub fn ltxv_dit_config_preset(name: *const c_char, out: *mut ltxv_dit_config) -> c_int;
    pub fn ltxv_dit_create(cfg: *const ltxv_dit_config, device: c_int, out: *mut *mut ltxv_dit) -> c_int;
    pub fn ltxv_dit_destroy(m: *mut ltxv_dit);
    pub fn ltxv_dit_load_tensor(m: *mut ltxv_dit, key: *const c_char, data: *const c_void, dtype: c_int,
                                shape: *const i64, rank: c_int) -> c_int;
    pub fn ltxv_dit_finalize(m: *mut ltxv_dit) -> c_int;
    pub fn ltxv_dit_set_skip_blocks(m: *mut ltxv_dit, idx: *const i32, n: c_int) -> c_int;
    pub fn ltxv_dit_forward(m: *mut ltxv_dit, hidden: *const c_void, hidden_dtype: c_int, enc: *const c_void,
                            enc_dtype: c_int, timestep: *const f32, mask: *const f32, b: c_int, s: c_int, k: c_int,
                            f: c_int, h: c_int, w: c_int, rope_scale3: *const f32, video_coords: *const f32,
                            skip_layer_mask: *const f32, out: *mut c_void, out_dtype: c_int,
                            stream: *mut c_void) -> c_int;

    // ---- VAE ----
    pub fn ltxv_vae_config_default(out: *mut ltxv_vae_config) -> c_int;
    pub fn ltxv_vae_create(cfg: *const ltxv_vae_config, device: c_int, out: *mut *mut ltxv_vae) -> c_int;
    pub fn ltxv_vae_destroy(m: *mut ltxv_vae);
    pub fn ltxv_vae_load_tensor(m: *mut ltxv_vae, key: *const c_char, data: *const c_void, dtype: c_int,
                                shape: *const i64, rank: c_int) -> c_int;
    pub fn ltxv_vae_finalize(m: *mut ltxv_vae) -> c_int;
    pub fn ltxv_vae_latents_mean(m: *mut ltxv_vae) -> *const f32;
    pub fn ltxv_vae_latents_std(m: *mut ltxv_vae) -> *const f32;
    pub fn ltxv_vae_decode(m: *mut ltxv_vae, z: *const c_void, z_dtype: c_int, timestep: *const f32, b: c_int,
                           f: c_int, h: c_int, w: c_int, out: *mut c_void, out_dtype: c_int, postprocess: c_int,
                           stream: *mut c_void) -> c_int;
    pub fn ltxv_vae_tiling_default(out: *mut ltxv_vae_tiling) -> c_int;
    pub fn ltxv_vae_decode_tiled(m: *mut ltxv_vae, z: *const c_void, z_dtype: c_int, timestep: *const f32, b: c_int,
                                 f: c_int, h: c_int, w: c_int, tiling: *const ltxv_vae_tiling, out: *mut c_void,
                                 out_dtype: c_int, postprocess: c_int, stream: *mut c_void) -> c_int;
    pub fn ltxv_vae_encoder_config_default(out: *mut ltxv_vae_encoder_config) -> c_int;
    pub fn ltxv_vae_enable_encoder(m: *mut ltxv_vae, cfg: *const ltxv_vae_encoder_config) -> c_int;
    pub fn ltxv_vae_encode_dims(m: *mut ltxv_vae, f: c_int, h: c_int, w: c_int, fl: *mut i32, hl: *mut i32,
                                wl: *mut i32) -> c_int;
    pub fn ltxv_vae_encode(m: *mut ltxv_vae, x: *const c_void, x_dtype: c_int, b: c_int, f: c_int, h: c_int, w: c_int,
                           moments: *mut f32, stream: *mut c_void) -> c_int;
    pub fn ltxv_causal_conv3d(x: *const f32, weight: *const f32, bias: *const f32, in_channels: c_int,
                              out_channels: c_int, t: c_int, h: c_int, w: c_int, is_causal: c_int, out: *mut f32,
                              stream: *mut c_void) -> c_int;

    // ---- pipeline glue (t2v_pipeline.rs) ----
    pub fn ltxv_pack_latents(x: *const f32, out: *mut f32, b: c_int, c: c_int, f: c_int, h: c_int, w: c_int,
                             p: c_int, pt: c_int, stream: *mut c_void) -> c_int;
    pub fn ltxv_unpack_latents(x: *const f32, out: *mut f32, b: c_int, c: c_int, f: c_int, h: c_int, w: c_int,
                               p: c_int, pt: c_int, stream: *mut c_void) -> c_int;
    pub fn ltxv_video_coords(out: *mut f32, batch: c_int, f: c_int, h: c_int, w: c_int, ts_ratio: c_int,
                             sp_ratio: c_int, frame_rate: c_int, stream: *mut c_void) -> c_int;
    pub fn ltxv_guidance_euler_step(cond: *const f32, uncond: *const f32, perturbed: *const f32, latents: *mut f32,
                                    noise_pred_out: *mut f32, b: c_int, n_per_batch: i64, guidance_scale: f32,
                                    guidance_rescale: f32, stg_scale: f32, sigma: f32, sigma_next: f32,
                                    stream: *mut c_void) -> c_int;
    pub fn ltxv_denormalize_latents(x: *const f32, out: *mut f32, mean: *const f32, std: *const f32,
                                    scaling_factor: f32, b: c_int, c: c_int, n_per_channel: i64,
                                    stream: *mut c_void) -> c_int;
    pub fn ltxv_normalize_latents(x: *const f32, out: *mut f32, mean: *const f32, std: *const f32,
                                  scaling_factor: f32, b: c_int, c: c_int, n_per_channel: i64,
                                  stream: *mut c_void) -> c_int;
    pub fn ltxv_postprocess_video(x: *const f32, out: *mut f32, n: i64, stream: *mut c_void) -> c_int;
    pub fn ltxv_frames_to_u8(frames: *const f32, out: *mut u8, b: c_int, f: c_int, h: c_int, w: c_int,
                             stream: *mut c_void) -> c_int;
    pub fn ltxv_calculate_shift(seq_len: c_int, out: *mut f32) -> c_int;
    pub fn ltxv_scheduler_set_timesteps(n: c_int, custom_sigmas: *const f32, mu: f32, has_terminal: c_int,
                                        terminal: f32, sigmas_out: *mut f32, timesteps_out: *mut i64) -> c_int;

    // ---- whole-loop entry points ----
    pub fn ltxv_pipeline_denoise(dit: *mut ltxv_dit, p: *const ltxv_pipeline_params, latents: *mut f32,
                                 prompt_embeds: *const c_void, prompt_mask: *const f32, negative_embeds: *const c_void,
                                 negative_mask: *const f32, embeds_dtype: c_int, k: c_int, stream: *mut c_void) -> c_int;
    pub fn ltxv_pipeline_decode(vae: *mut ltxv_vae, p: *const ltxv_pipeline_params, latents: *const f32, out: *mut f32,
                                stream: *mut c_void) -> c_int;
    pub fn ltxv_pipeline_denoise_stochastic(dit: *mut ltxv_dit, p: *const ltxv_pipeline_params, latents: *mut f32,
                                            prompt_embeds: *const c_void, prompt_mask: *const f32,
                                            negative_embeds: *const c_void, negative_mask: *const f32,
                                            embeds_dtype: c_int, k: c_int, step_noise: *const f32,
                                            stream: *mut c_void) -> c_int;
    pub fn ltxv_pipeline_decode_noisy(vae: *mut ltxv_vae, p: *const ltxv_pipeline_params, latents: *const f32,
                                      noise: *const f32, decode_noise_scale: f32, out: *mut f32,
                                      stream: *mut c_void) -> c_int;

    // ---- multi-GPU (one process per GPU) ----
    pub fn ltxv_comm_create(nranks: c_int, rank: c_int, device: c_int, heap_bytes: u64, out: *mut *mut ltxv_comm) -> c_int;
    pub fn ltxv_comm_destroy(c: *mut ltxv_comm);
    pub fn ltxv_comm_get_handle(c: *mut ltxv_comm, handle64: *mut c_void) -> c_int;
    pub fn ltxv_comm_open(c: *mut ltxv_comm, all_handles: *const c_void) -> c_int;
    pub fn ltxv_comm_barrier(c: *mut ltxv_comm, stream: *mut c_void) -> c_int;
    pub fn ltxv_parallel_plan(nranks: c_int, rank: c_int, s: c_int, do_cfg: c_int, out6: *mut i32) -> c_int;
    pub fn ltxv_pipeline_denoise_parallel(dit: *mut ltxv_dit, c: *mut ltxv_comm, p: *const ltxv_pipeline_params,
                                          latents: *mut f32, prompt_embeds: *const c_void, prompt_mask: *const f32,
                                          negative_embeds: *const c_void, negative_mask: *const f32,
                                          embeds_dtype: c_int, k: c_int, stream: *mut c_void) -> c_int;
    pub fn ltxv_vae_set_comm(vae: *mut ltxv_vae, c: *mut ltxv_comm) -> c_int;
    pub fn ltxv_pipeline_denoise_parallel_stochastic(dit: *mut ltxv_dit, c: *mut ltxv_comm,
                                                     p: *const ltxv_pipeline_params, latents: *mut f32,
                                                     prompt_embeds: *const c_void, prompt_mask: *const f32,
                                                     negative_embeds: *const c_void, negative_mask: *const f32,
                                                     embeds_dtype: c_int, k: c_int, step_noise: *const f32,
                                                     stream: *mut c_void) -> c_int;

    // ---- the rest of include/ltxv.h: not called by b200_models.rs, declared so the binding covers the whole ABI ----
    // context slots: the text K/V of a prompt prepared once and reused by every forward of the denoise loop
    pub fn ltxv_dit_prepare_context(m: *mut ltxv_dit, slot: c_int, enc: *const c_void, enc_dtype: c_int,
                                    mask: *const f32, k: c_int, stream: *mut c_void) -> c_int;
    pub fn ltxv_dit_forward_ctx(m: *mut ltxv_dit, slot: c_int, hidden: *const c_void, hidden_dtype: c_int,
                                timestep: *const f32, s: c_int, f: c_int, h: c_int, w: c_int,
                                rope_scale3: *const f32, video_coords: *const f32, skip_layer_mask: *const f32,
                                out: *mut c_void, out_dtype: c_int, stream: *mut c_void) -> c_int;
    pub fn ltxv_dit_get_config(m: *const ltxv_dit, out: *mut ltxv_dit_config) -> c_int;
    pub fn ltxv_vae_spatial_compression_ratio(m: *const ltxv_vae) -> c_int;
    pub fn ltxv_vae_temporal_compression_ratio(m: *const ltxv_vae) -> c_int;
    pub fn ltxv_vae_encode_tiled(m: *mut ltxv_vae, x: *const c_void, x_dtype: c_int, b: c_int, f: c_int, h: c_int,
                                 w: c_int, tiling: *const ltxv_vae_tiling, use_framewise_encoding: c_int,
                                 moments: *mut f32, stream: *mut c_void) -> c_int;
    pub fn ltxv_scheduler_step_stochastic(latents: *mut f32, model_output: *const f32, noise: *const f32, n: i64,
                                          sigma: f32, sigma_next: f32, stream: *mut c_void) -> c_int;
    pub fn ltxv_decode_noise_blend(latents: *mut f32, noise: *const f32, scale: f32, n: i64,
                                   stream: *mut c_void) -> c_int;
    // host-buffer entry points (pageable or pinned host memory in, host memory out; they stage through the handle's
    // pinned buffers and synchronise before returning) -- what bench.py times as `e2e`
    pub fn ltxv_dit_forward_host(m: *mut ltxv_dit, hidden: *const c_void, hidden_dtype: c_int, enc: *const c_void,
                                 enc_dtype: c_int, timestep: *const f32, mask: *const f32, b: c_int, s: c_int,
                                 k: c_int, f: c_int, h: c_int, w: c_int, rope_scale3: *const f32,
                                 video_coords: *const f32, skip_layer_mask: *const f32, out: *mut c_void,
                                 out_dtype: c_int) -> c_int;
    pub fn ltxv_vae_decode_host(m: *mut ltxv_vae, z: *const c_void, z_dtype: c_int, timestep: *const f32, b: c_int,
                                f: c_int, h: c_int, w: c_int, out: *mut c_void, out_dtype: c_int,
                                postprocess: c_int) -> c_int;
    pub fn ltxv_vae_encode_host(m: *mut ltxv_vae, x: *const c_void, x_dtype: c_int, b: c_int, f: c_int, h: c_int,
                                w: c_int, moments: *mut f32) -> c_int;
    pub fn ltxv_pipeline_denoise_host(dit: *mut ltxv_dit, p: *const ltxv_pipeline_params, latents: *mut f32,
                                      prompt_embeds: *const c_void, prompt_mask: *const f32,
                                      negative_embeds: *const c_void, negative_mask: *const f32,
                                      embeds_dtype: c_int, k: c_int) -> c_int;
    pub fn ltxv_pipeline_decode_host(vae: *mut ltxv_vae, p: *const ltxv_pipeline_params, latents: *const f32,
                                     out: *mut f32) -> c_int;
    pub fn ltxv_pipeline_decode_host_u8(vae: *mut ltxv_vae, p: *const ltxv_pipeline_params, latents: *const f32,
                                        out: *mut u8) -> c_int;
    // weight-file listing, random initialisation (benchmarks), measurement and experiment hooks
    pub fn ltxv_safetensors_list(path: *const c_char, out: *mut c_char, out_cap: u64, n_tensors: *mut i32) -> c_int;
    pub fn ltxv_dit_init_random(m: *mut ltxv_dit, seed: u64) -> c_int;
    pub fn ltxv_vae_init_random(m: *mut ltxv_vae, seed: u64) -> c_int;
    pub fn ltxv_profile_begin() -> c_int;
    pub fn ltxv_profile_end(launches8: *mut u64, ms8: *mut f64, work8: *mut f64) -> c_int;
    pub fn ltxv_trace_begin() -> c_int;
    pub fn ltxv_trace_end(out: *mut c_char, out_cap: u64) -> c_int;
    pub fn ltxv_set_option(name: *const c_char, value: c_int) -> c_int;
    pub fn ltxv_get_option(name: *const c_char, value: *mut c_int) -> c_int;
}

/// Non-zero return code -> the message `ltxv_last_error()` holds, as the error type the reference bails with
/// (`candle_core::bail!`, e.g. ltx_transformer.rs:451-453).
pub fn check(rc: c_int) -> candle_core::Result<()> {
    if rc == 0 {
        return Ok(());
    }
    let msg = unsafe { std::ffi::CStr::from_ptr(ltxv_last_error()) }.to_string_lossy().into_owned();
    Err(candle_core::Error::Msg(msg))
}
