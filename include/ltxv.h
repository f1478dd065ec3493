/* ltxv.h -- C ABI of libltxv_b200.so: the B200-native (sm_100a) drop-in for the LTX-Video hot path of
 * FerrisMind/candle-video.
 *
 * Every entry point replaces one reference interface (Rust, cited as file:line under /root/reference):
 *
 *   ltxv_dit_*        <->  `trait VideoTransformer3D` (src/models/ltx_video/t2v_pipeline.rs:63-83) as implemented by
 *                          `LtxVideoTransformer3DModel` (ltx_transformer.rs:1029-1215); construction through
 *                          `VarBuilder` key names (ltx_transformer.rs:957-1022)
 *   ltxv_vae_*        <->  `trait VaeLtxVideo` (t2v_pipeline.rs:91-103) as implemented by `AutoencoderKLLtxVideo`
 *                          (vae.rs:2437-2463 -> :2101 decode -> :1656 LtxVideoDecoder3d::forward)
 *   ltxv_pack_latents / ltxv_unpack_latents   <->  LtxPipeline::pack_latents / unpack_latents (t2v_pipeline.rs:474-550)
 *   ltxv_video_coords                          <->  the coordinate grid built inline at t2v_pipeline.rs:798-847
 *   ltxv_guidance_euler_step                   <->  CFG/STG combine (t2v_pipeline.rs:942-964), rescale_noise_cfg (:227-243)
 *                                                   and FlowMatchEulerDiscreteScheduler::step (scheduler.rs:544-582)
 *   ltxv_denormalize_latents / ltxv_postprocess_video <-> t2v_pipeline.rs:573-594 / :146-155
 *   ltxv_scheduler_set_timesteps               <->  Scheduler::set_timesteps as driven by LtxPipeline::call
 *                                                   (scheduler.rs:274-412, :646-660; t2v_pipeline.rs:752-792) -- host math
 *   ltxv_pipeline_denoise / ltxv_pipeline_decode <-> LtxPipeline::call denoise loop (:860-994) and decode branch (:1000-1072)
 *
 * Conventions
 *   - plain C types only; opaque handles; tensors are caller-owned, contiguous, row-major.
 *   - `*_host` variants take HOST pointers and include the host<->device copies and a stream sync; all other tensor
 *     arguments are DEVICE pointers and the call only enqueues work on `stream` (a cudaStream_t; NULL = default stream)
 *     without synchronising, like a Candle CustomOp (cf. candle_flash_attn::flash_attn, ltx_transformer.rs:707).
 *   - return value 0 = success, non-zero = error; ltxv_last_error() returns the message for the calling thread
 *     (the reference returns candle_core::Error::Msg for the same conditions, e.g. ltx_transformer.rs:451-453,:834-844).
 *   - dtype codes: LTXV_F32 = 0, LTXV_BF16 = 1.  The model computes in bf16 with f32 accumulation and f32
 *     norm / softmax / GELU / RoPE math; there is no CPU fallback: without a CUDA device every
 *     compute entry point fails with an error.
 */
#ifndef LTXV_H_
#define LTXV_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LTXV_F32 0
#define LTXV_BF16 1

typedef struct ltxv_dit ltxv_dit;
typedef struct ltxv_vae ltxv_vae;

const char* ltxv_last_error(void);
const char* ltxv_version(void);
/* kernels launched by this library since process start (all streams); evidence for bench.py's gpu_launches */
uint64_t ltxv_launch_count(void);

/* Measurement hook (bench.py): between begin/end every GEMM / conv3d / attention launch is bracketed by CUDA events
 * on its stream.  end() synchronises and returns per class {0 GEMM, 1 conv3d, 2 self-attention, 3 cross-attention,
 * 4 norm+modulate, 5 q/k norm+RoPE, 6 VAE pixel-norm/modulate/SiLU, 7 reserved}: launches, summed device
 * milliseconds, and summed algorithmic FLOPs (classes 0-3) or algorithmic BYTES (classes 4-7).  All three arrays
 * have 8 entries.  Not part of the reference interface. */
int ltxv_profile_begin(void);
int ltxv_profile_end(uint64_t* launches8, double* ms8, double* work8);

/* LtxVideoCausalConv3d::forward (vae.rs:298-465; operator-level reference test tests/verify_conv3d_parity.rs:41-43):
 * 3x3x3, stride 1.  x [1,Cin,T,H,W], weight [Cout,Cin,3,3,3], bias [Cout] or NULL, out [1,Cout,T,H,W]: device f32,
 * NCDHW.  is_causal: two copies of frame 0 in front (:383-387), else one replicated frame each side (:388-411); H/W
 * zero padding.  Cin % 64 == 0, Cout % 32 == 0, W <= 352.  bf16 operands, f32 accumulation, one bf16 rounding of
 * the result (what the decoder's convs do).  Synchronises the stream before returning. */
int ltxv_causal_conv3d(const float* x, const float* weight, const float* bias, int in_channels, int out_channels,
                       int T, int H, int W, int is_causal, float* out, void* stream);

/* Experiment knobs: kernel-variant switches, each initialised once per process from its LTXV_*
 * environment variable; these calls override them at run time (names: no_cfg_batch, gemm_no_pair, conv_no_kw3,
 * gemm_no_raster, gemm_no_epi2, gemm_k2, gemm_no_short_k, attn_v1, attn_v4, attn_nosplit, attn_nsplit_max,
 * vae_prep_u, vae_no_fused_prep, vae_no_fuse_conv2, vae_fuse_conv2, no_pdl (bit mask), qk_unfused).  Unknown names
 * return an error.  Process-wide; not synchronised with in-flight calls of other threads. */
int ltxv_set_option(const char* name, int value);
int ltxv_get_option(const char* name, int* value);

/* Test hook: between begin/end every launcher records the kernel variant it selected (template instance, tile mode,
 * attention key-split count ...).  end() writes "name count\n" lines (sorted) into `out`.  The production-shape parity
 * tests use it to assert that the variants the benchmark runs are the ones they compared with the oracle.  Not part of
 * the reference interface. */
int ltxv_trace_begin(void);
int ltxv_trace_end(char* out, uint64_t out_cap);

/* --------------------------------------------------------------- weight files ------------------------------------- */
/* Official single-file checkpoints (e.g. ltx-video-2b-v0.9.5.safetensors) name tensors differently from the diffusers
 * layout the models consume.  ltxv_remap_official_key_raw = KeyRemapper::remap_key (weight_format.rs:55-143);
 * ltxv_remap_official_key additionally strips the "vae." / "model.diffusion_model." / "transformer." prefix and
 * reports the component (0 other, 1 transformer, 2 VAE) exactly as examples/ltx-video/main.rs:480-497 routes tensors. */
int ltxv_remap_official_key_raw(const char* key, char* out, uint64_t out_cap);
int ltxv_remap_official_key(const char* key, char* out, uint64_t out_cap, int32_t* component);
/* Tensor index of a safetensors file / directory, one line per tensor: "<name> <dtype> [d0,d1,...] <bytes>\n", sorted
 * by name.  `path`: a .safetensors file, a directory holding one, or a directory with `*.safetensors.index.json` and
 * its shards (loader.rs:341-371; a shard named by the index but missing on disk is an error). */
int ltxv_safetensors_list(const char* path, char* out, uint64_t out_cap, int32_t* n_tensors);

/* ------------------------------------------------------------------ DiT ------------------------------------------ */
/* LtxVideoTransformer3DModelConfig, ltx_transformer.rs:22-59 */
typedef struct ltxv_dit_config {
    int32_t in_channels;          /* 128 */
    int32_t out_channels;         /* 128 */
    int32_t patch_size;           /* 1 */
    int32_t patch_size_t;         /* 1 */
    int32_t num_attention_heads;  /* 32 */
    int32_t attention_head_dim;   /* 64 (2B) | 128 (13B) */
    int32_t cross_attention_dim;  /* 2048 | 4096 */
    int32_t num_layers;           /* 28 | 48 */
    int32_t caption_channels;     /* 4096 */
    float norm_eps;               /* 1e-6 */
    int32_t timestep_bf16_round;  /* 1: replay `timestep.to_dtype(bf16)` of a bf16 model (ltx_transformer.rs:1051) */
} ltxv_dit_config;

/* presets configs.rs:135-160: "2b", "13b" */
int ltxv_dit_config_preset(const char* name, ltxv_dit_config* out);

int ltxv_dit_create(const ltxv_dit_config* cfg, int device, ltxv_dit** out);
void ltxv_dit_destroy(ltxv_dit* m);
/* Copy one tensor into the model. `key` is the diffusers / VarBuilder name, `data` may be a
 * host or a device pointer, dtype f32 or bf16; the shape must match the reference's. */
int ltxv_dit_load_tensor(ltxv_dit* m, const char* key, const void* data, int dtype, const int64_t* shape, int rank);
/* Synthetic random-init weights of the configured architecture, generated on the device (benchmarks). */
/* Loads every tensor of `path` (see ltxv_safetensors_list) the model has a slot for: F32 / BF16 / F16 sources, cast on
 * load.  official = 1: keys go through the official->diffusers remap first and only transformer tensors are taken.
 * Still requires ltxv_dit_finalize (which reports tensors that were never loaded). */
int ltxv_dit_load_safetensors(ltxv_dit* m, const char* path, int official, int32_t* n_loaded, int32_t* n_ignored);
int ltxv_dit_init_random(ltxv_dit* m, uint64_t seed);
/* Fails (listing the missing keys) unless every tensor of the architecture has been loaded. */
int ltxv_dit_finalize(ltxv_dit* m);
/* VideoTransformer3D::set_skip_block_list, ltx_transformer.rs:1024-1026 */
int ltxv_dit_set_skip_blocks(ltxv_dit* m, const int32_t* idx, int n);
/* TransformerConfig accessors (t2v_pipeline.rs:55-61) */
int ltxv_dit_get_config(const ltxv_dit* m, ltxv_dit_config* out);

/* VideoTransformer3D::forward (t2v_pipeline.rs:63-83).  Device pointers:
 *   hidden [B,S,in_channels] (hidden_dtype), enc [B,K,caption_channels] (enc_dtype), timestep f32 [B],
 *   mask f32 [B,K] (1 keep / 0 pad; may be NULL = no mask), video_coords f32 [B,S,3] or NULL,
 *   out [B,S,out_channels] (out_dtype).
 * Host pointers (tiny control data): rope_scale3 (3 floats, or NULL), skip_layer_mask f32 [num_layers, B] or NULL. */
int ltxv_dit_forward(ltxv_dit* m, const void* hidden, int hidden_dtype, const void* enc, int enc_dtype,
                     const float* timestep, const float* mask, int B, int S, int K, int F, int H, int W,
                     const float* rope_scale3, const float* video_coords, const float* skip_layer_mask, void* out,
                     int out_dtype, void* stream);
/* Same call on HOST buffers: copies inputs to the device, runs, copies the result back, synchronises. */
int ltxv_dit_forward_host(ltxv_dit* m, const void* hidden, int hidden_dtype, const void* enc, int enc_dtype,
                          const float* timestep, const float* mask, int B, int S, int K, int F, int H, int W,
                          const float* rope_scale3, const float* video_coords, const float* skip_layer_mask, void* out,
                          int out_dtype);

/* Step-invariant text path, hoisted: caption projection (ltx_transformer.rs:186-190) and every block's cross-attention
 * K/V projection + k-norm (:667-672) computed once into context slot `slot` (0..3) for a [K,caption_channels] prompt
 * (one batch entry).  ltxv_dit_forward_ctx then runs VideoTransformer3D::forward for B = 1 against that slot. */
int ltxv_dit_prepare_context(ltxv_dit* m, int slot, const void* enc, int enc_dtype, const float* mask, int K,
                             void* stream);
int ltxv_dit_forward_ctx(ltxv_dit* m, int slot, const void* hidden, int hidden_dtype, const float* timestep, int S,
                         int F, int H, int W, const float* rope_scale3, const float* video_coords,
                         const float* skip_layer_mask, void* out, int out_dtype, void* stream);

/* ------------------------------------------------------------------ VAE ------------------------------------------ */
/* AutoencoderKLLtxVideoConfig, decoder-relevant fields (vae.rs:30-103, configs.rs:84-93) */
typedef struct ltxv_vae_config {
    int32_t latent_channels;                /* 128 */
    int32_t out_channels;                   /* 3 */
    int32_t decoder_block_out_channels[3];  /* 256, 512, 1024 */
    int32_t decoder_layers_per_block[4];    /* 5, 5, 5, 5 */
    int32_t patch_size;                     /* 4 */
    int32_t timestep_conditioning;          /* 1 */
    float scaling_factor;                   /* 1.0 */
} ltxv_vae_config;

int ltxv_vae_config_default(ltxv_vae_config* out);
int ltxv_vae_create(const ltxv_vae_config* cfg, int device, ltxv_vae** out);
void ltxv_vae_destroy(ltxv_vae* m);
/* keys carry the reference's `decoder.` prefix; `latents_mean` / `latents_std` (top level) are optional.
 * `encoder.*` keys are ignored unless ltxv_vae_enable_encoder was called (the reference builds an encoder the t2v
 * path never runs); quant_conv / post_quant_conv keys are always ignored. */
int ltxv_vae_load_tensor(ltxv_vae* m, const char* key, const void* data, int dtype, const int64_t* shape, int rank);
int ltxv_vae_load_safetensors(ltxv_vae* m, const char* path, int official, int32_t* n_loaded, int32_t* n_ignored);
int ltxv_vae_init_random(ltxv_vae* m, uint64_t seed);
int ltxv_vae_finalize(ltxv_vae* m);
/* VaeLtxVideo accessors (t2v_pipeline.rs:91-99): 128-entry f32 device vectors, ratios 32 / 8 */
const float* ltxv_vae_latents_mean(const ltxv_vae* m);
const float* ltxv_vae_latents_std(const ltxv_vae* m);
int ltxv_vae_spatial_compression_ratio(const ltxv_vae* m);
int ltxv_vae_temporal_compression_ratio(const ltxv_vae* m);

/* VaeLtxVideo::decode (t2v_pipeline.rs:102): z [B,128,F,H,W] NCDHW (z_dtype) -> out [B,3,8F-7,32H,32W] NCDHW.
 * timestep: device f32 [B] or NULL.  postprocess != 0 additionally applies LtxVideoProcessor::postprocess_video
 * (clamp(0.5x+0.5,0,1)*255, t2v_pipeline.rs:147-155) in the last kernel's epilogue. */
int ltxv_vae_decode(ltxv_vae* m, const void* z, int z_dtype, const float* timestep, int B, int F, int H, int W,
                    void* out, int out_dtype, int postprocess, void* stream);
/* Tiling knobs of AutoencoderKLLtxVideo (vae.rs:1848-1861, enable_tiling :1870-1898), all in SAMPLE space. */
typedef struct ltxv_vae_tiling {
    int32_t use_tiling;                    /* 1 */
    int32_t use_framewise_decoding;        /* 1 */
    int32_t tile_sample_min_height;        /* 512 */
    int32_t tile_sample_min_width;         /* 512 */
    int32_t tile_sample_min_num_frames;    /* 16 */
    int32_t tile_sample_stride_height;     /* 384 */
    int32_t tile_sample_stride_width;      /* 384 */
    int32_t tile_sample_stride_num_frames; /* 8 */
} ltxv_vae_tiling;
int ltxv_vae_tiling_default(ltxv_vae_tiling* out);
/* decode_z of the reference with its tiling dispatch (vae.rs:2037-2066): temporal tiling when F > min_frames/8, else
 * spatial tiling when H or W exceed the minimum tile, else the plain decoder; seams blended linearly in f32.
 * tiling == NULL is ltxv_vae_decode.  Single-GPU only (H-slab decode already removes the need for tiles). */
int ltxv_vae_decode_tiled(ltxv_vae* m, const void* z, int z_dtype, const float* timestep, int B, int F, int H, int W,
                          const ltxv_vae_tiling* tiling, void* out, int out_dtype, int postprocess, void* stream);
int ltxv_vae_decode_host(ltxv_vae* m, const void* z, int z_dtype, const float* timestep, int B, int F, int H, int W,
                         void* out, int out_dtype, int postprocess);

/* ---- encoder half: `AutoencoderKLLtxVideo::encode` (vae.rs:2070-2099) -> LtxVideoEncoder3d
 * (vae.rs:1315-1469).  The reference's t2v path never calls it; it is here for image/video conditioning callers.
 * Encoder-relevant fields of AutoencoderKLLtxVideoConfig (vae.rs:30-103); LTX-Video 0.9.5 layout only. */
enum { LTXV_DOWN_SPATIAL = 1, LTXV_DOWN_TEMPORAL = 2, LTXV_DOWN_SPATIOTEMPORAL = 3 }; /* DownsampleType, vae.rs:469-493 */
typedef struct ltxv_vae_encoder_config {
    int32_t in_channels;           /* 3 */
    int32_t latent_channels;       /* 128 */
    int32_t block_out_channels[5]; /* 128, 256, 512, 1024, 2048 */
    int32_t layers_per_block[5];   /* 4, 6, 6, 2, 2 (the last entry counts the mid block: layers - 1 resnets) */
    int32_t downsample_types[4];   /* spatial, temporal, spatiotemporal, spatiotemporal */
    int32_t patch_size;            /* 4 */
} ltxv_vae_encoder_config;
int ltxv_vae_encoder_config_default(ltxv_vae_encoder_config* out);
/* Builds the encoder inside `m`: from here on `encoder.*` keys are loaded (and required by ltxv_vae_finalize) instead
 * of ignored, and ltxv_vae_init_random also fills the encoder. */
int ltxv_vae_enable_encoder(ltxv_vae* m, const ltxv_vae_encoder_config* cfg);
/* latent extent (F', H', W') of an [F, H, W] video; fails when the extent does not divide through the downsamplers */
int ltxv_vae_encode_dims(const ltxv_vae* m, int F, int H, int W, int32_t* Fl, int32_t* Hl, int32_t* Wl);
/* x [B,3,F,H,W] NCDHW in [-1,1] (x_dtype, device) -> moments f32 [B, 2*latent, F', H', W'] (device): channels
 * [0, latent) are the posterior mean (DiagonalGaussianDistribution::mode, vae.rs:135), [latent, 2*latent) the
 * log-variance.  Untiled (encode_z with use_tiling / use_framewise_encoding off), no quant_conv. */
int ltxv_vae_encode(ltxv_vae* m, const void* x, int x_dtype, int B, int F, int H, int W, float* moments, void* stream);
/* encode_z of the reference with its tiling dispatch (vae.rs:2017-2034): temporal tiling when use_framewise_encoding and
 * F > tile_sample_min_num_frames (:2294-2356), else spatial tiling when H or W exceed the minimum tile (:2158-2223);
 * tiles are cut in sample space and blended linearly in latent space.  The reference's LIBRARY DEFAULT is
 * use_tiling = 1, use_framewise_encoding = 0 (vae.rs:1856-1858): pass ltxv_vae_tiling_default() to reproduce what its
 * encode() returns for inputs wider than 512 px.  tiling == NULL is ltxv_vae_encode.  (`use_framewise_decoding` of the
 * struct is ignored here.) */
int ltxv_vae_encode_tiled(ltxv_vae* m, const void* x, int x_dtype, int B, int F, int H, int W,
                          const ltxv_vae_tiling* tiling, int use_framewise_encoding, float* moments, void* stream);
/* same as ltxv_vae_encode with HOST buffers (copies inside) */
int ltxv_vae_encode_host(ltxv_vae* m, const void* x, int x_dtype, int B, int F, int H, int W, float* moments);

/* ------------------------------------------------------------- pipeline glue -------------------------------------- */
/* f32 device tensors.  pack: [B,C,F,H,W] -> [B,S,C*pt*p*p]; unpack is the inverse (F,H,W = unpacked dims). */
int ltxv_pack_latents(const float* in, float* out, int B, int C, int F, int H, int W, int p, int pt, void* stream);
int ltxv_unpack_latents(const float* in, float* out, int B, int C, int F, int H, int W, int p, int pt, void* stream);
/* out f32 [B,S,3]: (clamp(ts*f + 1 - ts, 0, 1000) * (float)(1/fps), sp*h, sp*w) in f,h,w token order */
int ltxv_video_coords(float* out, int B, int F, int H, int W, int ts_ratio, int sp_ratio, int fps, void* stream);
/* cond / uncond / perturbed / latents / noise_pred_out: f32 [B, n_per_batch] (uncond, perturbed, noise_pred_out,
 * latents may be NULL).  comb = u + g(c-u) [rescaled] + s(c-p); latents += (sigma_next - sigma) * comb. */
int ltxv_guidance_euler_step(const float* cond, const float* uncond, const float* perturbed, float* latents,
                             float* noise_pred_out, int B, int64_t n_per_batch, float guidance_scale,
                             float guidance_rescale, float stg_scale, float sigma, float sigma_next, void* stream);
/* in/out f32 [B,C,n_per_channel]; mean/std device f32 [C] */
int ltxv_denormalize_latents(const float* in, float* out, const float* mean, const float* std, float scaling_factor,
                             int B, int C, int64_t n_per_channel, void* stream);
/* normalize_latents (t2v_pipeline.rs:552-571): (x - mean) * scaling_factor / std, same layout */
int ltxv_normalize_latents(const float* in, float* out, const float* mean, const float* std, float scaling_factor,
                           int B, int C, int64_t n_per_channel, void* stream);
int ltxv_postprocess_video(const float* in, float* out, int64_t n, void* stream);
/* Output hand-off of the reference's example binary (examples/ltx-video/main.rs:653-667): per frame
 * permute((1,2,0)).clamp(0,255).to_dtype(U8).  frames: device f32 [B,3,F,H,W] in 0..255 (postprocess_video output);
 * out: device u8 [B,F,H,W,3] (RGB8 rows as image::save_buffer / the GIF encoder take them). */
int ltxv_frames_to_u8(const float* frames, uint8_t* out, int B, int F, int H, int W, void* stream);

/* Host-only schedule math (stays as in the reference): writes num_steps+1 sigmas (terminal 0 appended) and num_steps
 * integer-truncated timesteps.  custom_sigmas may be NULL (then linspace(1, 1/n, n) with the SD3 shift mu). */
int ltxv_calculate_shift(int seq_len, float* mu_out);
int ltxv_scheduler_set_timesteps(int num_steps, const float* custom_sigmas, float mu, int has_shift_terminal,
                                 float shift_terminal, float* sigmas_out, int64_t* timesteps_out);

/* ------------------------------------------------------------- pipeline ------------------------------------------- */
/* LtxPipeline::call with precomputed prompt embeddings (examples/ltx-video/main.rs:621-646). */
typedef struct ltxv_pipeline_params {
    int32_t height, width, num_frames, frame_rate;
    int32_t num_inference_steps;
    const float* custom_sigmas; /* host, num_inference_steps entries, or NULL */
    float guidance_scale, guidance_rescale, stg_scale;
    const int32_t* skip_block_list; /* host */
    int32_t num_skip_blocks;
    int32_t has_shift_terminal; /* preset configs.rs:113 */
    float shift_terminal;
    float decode_timestep;
} ltxv_pipeline_params;

/* Denoise loop (t2v_pipeline.rs:860-994), B = 1.  latents: device f32 [S,128] packed, updated in place.
 * prompt/negative embeds: device [K, caption_channels] (dtype), masks device f32 [K]; negative_* may be NULL when
 * guidance_scale <= 1.  Timesteps are trunc(sigma*1000) as in the reference. */
int ltxv_pipeline_denoise(ltxv_dit* dit, const ltxv_pipeline_params* p, float* latents, const void* prompt_embeds,
                          const float* prompt_mask, const void* negative_embeds, const float* negative_mask,
                          int embeds_dtype, int K, void* stream);
/* Decode branch (t2v_pipeline.rs:1000-1072) with decode_noise_scale = 0: unpack -> denormalize -> VAE decode ->
 * postprocess.  latents: device f32 [S,128]; out: device f32 [3, num_frames, height, width] in 0..255. */
int ltxv_pipeline_decode(ltxv_vae* vae, const ltxv_pipeline_params* p, const float* latents, float* out, void* stream);

/* ---- stochastic sampling and decode noise with CALLER-SUPPLIED noise -----------------------------
 * The reference draws both from the device RNG (Tensor::randn -> cuRAND, scheduler.rs:566, t2v_pipeline.rs:1055); the
 * library has no RNG of its own: the caller (the Rust pipeline, still holding Candle's generator) passes the tensors,
 * so a run is reproducible bit for bit.  All f32, every arithmetic op rounded like the reference's tensor ops. */
/* x0 = x - sigma v; x <- (1 - sigma_next) x0 + sigma_next noise   (scheduler.rs:557-575) */
int ltxv_scheduler_step_stochastic(float* latents, const float* model_output, const float* noise, int64_t n, float sigma,
                                   float sigma_next, void* stream);
/* x <- x (1 - scale) + noise scale   (t2v_pipeline.rs:1049-1062) */
int ltxv_decode_noise_blend(float* latents, const float* noise, float scale, int64_t n, void* stream);
/* ltxv_pipeline_denoise with stochastic_sampling = true (preset 0.9.8-distilled, configs.rs:210):
 * step_noise = device f32 [num_inference_steps, S, 128], slice i used by step i. */
int ltxv_pipeline_denoise_stochastic(ltxv_dit* dit, const ltxv_pipeline_params* p, float* latents,
                                     const void* prompt_embeds, const float* prompt_mask, const void* negative_embeds,
                                     const float* negative_mask, int embeds_dtype, int K, const float* step_noise,
                                     void* stream);
/* ltxv_pipeline_decode with the decode-noise blend: noise = device f32 [128, F, H, W] (NULL only with scale 0). */
int ltxv_pipeline_decode_noisy(ltxv_vae* vae, const ltxv_pipeline_params* p, const float* latents, const float* noise,
                               float decode_noise_scale, float* out, void* stream);

/* ------------------------------------------------------------- multi-GPU ------------------------------------------ */
/* One process per GPU.  The reference has no distributed code at all; these entry
 * points add the sharding BASELINE.json's north_star names: CFG cond/uncond split across two rank groups, Ulysses
 * sequence parallelism (head <-> token exchange around self-attention) inside a group, and H-slab VAE decode with
 * conv halo exchange.  All data-path traffic is NVLink peer stores issued by the producing kernels into a symmetric
 * heap (one cudaMalloc per rank whose 64-byte CUDA IPC handle the host harness exchanges once with any
 * all-gather it has); ordering is a system-scope flag barrier kernel.  No NCCL on the data path. */
typedef struct ltxv_comm ltxv_comm;
int ltxv_comm_create(int nranks, int rank, int device, uint64_t heap_bytes, ltxv_comm** out);
void ltxv_comm_destroy(ltxv_comm* c);
int ltxv_comm_get_handle(ltxv_comm* c, void* handle64);             /* out: 64 bytes */
int ltxv_comm_open(ltxv_comm* c, const void* all_handles);          /* in: nranks * 64 bytes, indexed by rank */
int ltxv_comm_barrier(ltxv_comm* c, void* stream);
/* host-only: the sharding plan of ltxv_pipeline_denoise_parallel for (nranks, rank, S tokens, CFG on/off):
 * out6 = {cfg_groups, sp_size, branch (0 uncond / 1 cond), sp_rank, local tokens, first token} */
int ltxv_parallel_plan(int nranks, int rank, int S, int do_cfg, int32_t* out6);
/* ltxv_pipeline_denoise over all ranks of `c`: `latents` is the full [S,128] f32 tensor on every rank (identical on
 * entry, identical on exit); each rank runs one CFG branch on its token shard. */
int ltxv_pipeline_denoise_parallel(ltxv_dit* dit, ltxv_comm* c, const ltxv_pipeline_params* p, float* latents,
                                   const void* prompt_embeds, const float* prompt_mask, const void* negative_embeds,
                                   const float* negative_mask, int embeds_dtype, int K, void* stream);
/* ... with stochastic_sampling = true (scheduler.rs:557-575; ltxv_pipeline_denoise_stochastic): step_noise is the same
 * full [num_inference_steps, S, 128] f32 tensor on every rank, each rank reads its token shard's rows. */
int ltxv_pipeline_denoise_parallel_stochastic(ltxv_dit* dit, ltxv_comm* c, const ltxv_pipeline_params* p, float* latents,
                                              const void* prompt_embeds, const float* prompt_mask,
                                              const void* negative_embeds, const float* negative_mask, int embeds_dtype,
                                              int K, const float* step_noise, void* stream);
/* H-slab decode over all ranks: latent heights need not be divisible by the rank count (ragged slabs: the first
 * H mod nranks ranks take one latent row more; c3 / c5 latents are 22 rows high). */
/* H-slab decode over all ranks of `c` (NULL restores single-GPU decode): ltxv_vae_decode / ltxv_pipeline_decode then
 * take the full latent on every rank and deliver the video on rank 0 (other ranks' `out` is left untouched). */
int ltxv_vae_set_comm(ltxv_vae* vae, ltxv_comm* c);

/* HOST-buffer variants of the two calls above (what a caller holding CPU tensors uses): inputs are copied to the
 * device, the loop / decode runs, the result is copied back and the stream is synchronised. */
int ltxv_pipeline_denoise_host(ltxv_dit* dit, const ltxv_pipeline_params* p, float* latents, const void* prompt_embeds,
                               const float* prompt_mask, const void* negative_embeds, const float* negative_mask,
                               int embeds_dtype, int K);
int ltxv_pipeline_decode_host(ltxv_vae* vae, const ltxv_pipeline_params* p, const float* latents, float* out);
/* same, delivering what the reference's example hands to its image / GIF writers: host u8 [frames, height, width, 3]
 * (ltxv_frames_to_u8 on the device, so a quarter of the bytes cross PCIe) */
int ltxv_pipeline_decode_host_u8(ltxv_vae* vae, const ltxv_pipeline_params* p, const float* latents, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif /* LTXV_H_ */
