//! `src/models/ltx_video/b200_models.rs` -- the two trait objects `LtxPipeline` consumes
//! (`Box<dyn VideoTransformer3D>`, `Box<dyn VaeLtxVideo>`, t2v_pipeline.rs:245-251), implemented on libltxv_b200.so.
//!
//! Written against candle-core / candle-nn 0.9.2 with cudarc's `DevicePtr::device_ptr(&stream)` (the pattern
//! candle-flash-attn 0.9.2 uses for its own `extern "C" run_mha`); with an older cudarc replace `raw_ptr` by
//! `*slice.device_ptr() as *const c_void`.  Nothing else in the crate changes: `examples/ltx-video/main.rs:554-564`
//! passes `Box::new(B200Transformer::from_tensors(..)?)` and `Box::new(B200Vae::from_tensors(..)?)` to
//! `LtxPipeline::new`.
//!
//! Ownership: the library borrows the tensors' device memory for the duration of a call and owns only its weights
//! and workspaces.  All work is enqueued on Candle's stream of the tensors' device; the device-pointer entry points
//! never synchronise.

use std::collections::HashMap;
use std::ffi::CString;
use std::os::raw::{c_int, c_void};

use candle_core::cuda_backend::cudarc::driver::DevicePtr;
use candle_core::{bail, DType, Device, Result, Storage, Tensor};
use half::bf16;

use super::b200_ffi::*;
use super::ltx_transformer::LtxVideoTransformer3DModelConfig;
use super::t2v_pipeline::{TransformerConfig, VaeConfig, VaeLtxVideo, VideoTransformer3D};
use super::vae::AutoencoderKLLtxVideoConfig;

/// dtype code of the C ABI for a tensor
fn code(t: &Tensor) -> Result<c_int> {
    match t.dtype() {
        DType::F32 => Ok(LTXV_F32),
        DType::BF16 => Ok(LTXV_BF16),
        other => bail!("libltxv_b200 takes f32 or bf16 tensors, got {other:?}"),
    }
}

/// Raw device pointer of a CONTIGUOUS f32 / bf16 CUDA tensor (start offset applied) and the raw CUstream of its
/// device.  The pointer is valid while `t` is alive; the call it is passed to is enqueued on the returned stream,
/// which is the stream every later Candle op on this device is ordered after.
fn raw_ptr(t: &Tensor) -> Result<(*const c_void, *mut c_void)> {
    if !t.is_contiguous() {
        bail!("internal: raw_ptr needs a contiguous tensor");
    }
    let (storage, layout) = t.storage_and_layout();
    let cuda = match &*storage {
        Storage::Cuda(s) => s,
        _ => bail!("libltxv_b200 has no CPU path: tensors must live on a CUDA device"),
    };
    let stream = cuda.device().cuda_stream();
    let off = layout.start_offset();
    let ptr = match t.dtype() {
        DType::F32 => {
            let s = cuda.as_cuda_slice::<f32>()?.slice(off..);
            let (p, _sync) = s.device_ptr(&stream);
            p as usize
        }
        DType::BF16 => {
            let s = cuda.as_cuda_slice::<bf16>()?.slice(off..);
            let (p, _sync) = s.device_ptr(&stream);
            p as usize
        }
        other => bail!("libltxv_b200 takes f32 or bf16 tensors, got {other:?}"),
    };
    Ok((ptr as *const c_void, stream.cu_stream() as *mut c_void))
}

fn device_index(dev: &Device) -> Result<c_int> {
    match dev {
        Device::Cuda(_) => Ok(0), // candle-video drives one GPU per process (examples/ltx-video/main.rs:210-214)
        _ => bail!("libltxv_b200 has no CPU path"),
    }
}

/// Pushes every tensor of a diffusers-named state dict (the names VarBuilder resolves: SURVEY.md appendix A) through
/// `load`; host or device tensors, f32 or bf16.
fn push_tensors(tensors: &HashMap<String, Tensor>,
                mut load: impl FnMut(*const i8, *const c_void, c_int, *const i64, c_int) -> c_int) -> Result<()> {
    for (key, t) in tensors {
        let t = match t.dtype() {
            DType::F32 | DType::BF16 => t.contiguous()?,
            _ => t.to_dtype(DType::F32)?.contiguous()?,
        };
        let shape: Vec<i64> = t.dims().iter().map(|&d| d as i64).collect();
        let ckey = CString::new(key.as_str()).map_err(|e| candle_core::Error::Msg(e.to_string()))?;
        let rc = if t.device().is_cuda() {
            let (p, _stream) = raw_ptr(&t)?;
            load(ckey.as_ptr(), p, code(&t)?, shape.as_ptr(), shape.len() as c_int)
        } else if t.dtype() == DType::F32 {
            let v = t.flatten_all()?.to_vec1::<f32>()?;
            load(ckey.as_ptr(), v.as_ptr() as *const c_void, LTXV_F32, shape.as_ptr(), shape.len() as c_int)
        } else {
            let v = t.flatten_all()?.to_vec1::<bf16>()?;
            load(ckey.as_ptr(), v.as_ptr() as *const c_void, LTXV_BF16, shape.as_ptr(), shape.len() as c_int)
        };
        check(rc)?;
    }
    Ok(())
}

// =====================================================================================================================
// VideoTransformer3D  (t2v_pipeline.rs:63-83)
// =====================================================================================================================
pub struct B200Transformer {
    handle: *mut ltxv_dit,
    cfg: TransformerConfig,
    out_channels: usize,
}

unsafe impl Send for B200Transformer {}

impl B200Transformer {
    pub fn from_tensors(cfg: &LtxVideoTransformer3DModelConfig, tensors: &HashMap<String, Tensor>, dev: &Device)
        -> Result<Self> {
        let me = Self::empty(cfg, dev)?;
        push_tensors(tensors, |k, p, dt, sh, r| unsafe { ltxv_dit_load_tensor(me.handle, k, p, dt, sh, r) })?;
        check(unsafe { ltxv_dit_finalize(me.handle) })?; // names the missing keys instead of running on zeros
        Ok(me)
    }

    /// Straight from a checkpoint: a .safetensors file, a diffusers directory or a sharded directory; `official`
    /// applies the unified-file key remap (weight_format.rs:55-143).
    pub fn from_safetensors(cfg: &LtxVideoTransformer3DModelConfig, path: &str, official: bool, dev: &Device)
        -> Result<Self> {
        let me = Self::empty(cfg, dev)?;
        let cpath = CString::new(path).map_err(|e| candle_core::Error::Msg(e.to_string()))?;
        let (mut loaded, mut ignored) = (0i32, 0i32);
        check(unsafe { ltxv_dit_load_safetensors(me.handle, cpath.as_ptr(), official as c_int, &mut loaded, &mut ignored) })?;
        check(unsafe { ltxv_dit_finalize(me.handle) })?;
        Ok(me)
    }

    /// handle with no weights yet
    fn empty(cfg: &LtxVideoTransformer3DModelConfig, dev: &Device) -> Result<Self> {
        let out_channels = if cfg.out_channels == 0 { cfg.in_channels } else { cfg.out_channels };
        let c = ltxv_dit_config {
            in_channels: cfg.in_channels as i32,
            out_channels: out_channels as i32,
            patch_size: cfg.patch_size as i32,
            patch_size_t: cfg.patch_size_t as i32,
            num_attention_heads: cfg.num_attention_heads as i32,
            attention_head_dim: cfg.attention_head_dim as i32,
            cross_attention_dim: cfg.cross_attention_dim as i32,
            num_layers: cfg.num_layers as i32,
            caption_channels: cfg.caption_channels as i32,
            norm_eps: cfg.norm_eps as f32,
            timestep_bf16_round: 1, // the bf16 run's `timestep.to_dtype(model_dtype)` (ltx_transformer.rs:1051)
        };
        let mut handle: *mut ltxv_dit = std::ptr::null_mut();
        check(unsafe { ltxv_dit_create(&c, device_index(dev)?, &mut handle) })?;
        Ok(Self {
            handle,
            cfg: TransformerConfig {
                in_channels: cfg.in_channels,
                patch_size: cfg.patch_size,
                patch_size_t: cfg.patch_size_t,
                num_layers: cfg.num_layers,
            },
            out_channels,
        })
    }

    pub fn raw(&self) -> *mut ltxv_dit {
        self.handle
    }
}

impl Drop for B200Transformer {
    fn drop(&mut self) {
        unsafe { ltxv_dit_destroy(self.handle) };
    }
}

impl VideoTransformer3D for B200Transformer {
    fn config(&self) -> &TransformerConfig {
        &self.cfg
    }

    fn set_skip_block_list(&mut self, list: Vec<usize>) {
        let v: Vec<i32> = list.iter().map(|&x| x as i32).collect();
        // the trait method cannot fail; an out-of-range index is ignored by the forward like in the reference (:1094-1096)
        let _ = unsafe { ltxv_dit_set_skip_blocks(self.handle, v.as_ptr(), v.len() as c_int) };
    }

    #[allow(clippy::too_many_arguments)]
    fn forward(&mut self, hidden_states: &Tensor, encoder_hidden_states: &Tensor, timestep: &Tensor,
               encoder_attention_mask: &Tensor, num_frames: usize, height: usize, width: usize,
               rope_interpolation_scale: Option<(f32, f32, f32)>, video_coords: Option<&Tensor>,
               skip_layer_mask: Option<&Tensor>) -> Result<Tensor> {
        let (b, s, _c) = hidden_states.dims3()?;
        let k = encoder_hidden_states.dim(1)?;
        // views are common (video_coords is a broadcast_as, t2v_pipeline.rs:843-847): make everything contiguous
        let hs = hidden_states.contiguous()?;
        let enc = encoder_hidden_states.contiguous()?;
        let ts = timestep.to_dtype(DType::F32)?.flatten_all()?.contiguous()?;
        let mask = encoder_attention_mask.to_dtype(DType::F32)?.contiguous()?;
        let coords = match video_coords {
            Some(c) => Some(c.to_dtype(DType::F32)?.contiguous()?),
            None => None,
        };
        // skip_layer_mask [num_layers, batch] is tiny and consumed on the host (which blocks to skip / blend)
        let slm: Option<Vec<f32>> = match skip_layer_mask {
            Some(m) => Some(m.to_dtype(DType::F32)?.flatten_all()?.to_vec1::<f32>()?),
            None => None,
        };
        let scale: Option<[f32; 3]> = rope_interpolation_scale.map(|(a, b, c)| [a, b, c]);
        let out = Tensor::zeros((b, s, self.out_channels), hs.dtype(), hs.device())?;

        let (p_hs, stream) = raw_ptr(&hs)?;
        let (p_enc, _) = raw_ptr(&enc)?;
        let (p_ts, _) = raw_ptr(&ts)?;
        let (p_mask, _) = raw_ptr(&mask)?;
        let p_coords = match &coords {
            Some(c) => raw_ptr(c)?.0 as *const f32,
            None => std::ptr::null(),
        };
        let (p_out, _) = raw_ptr(&out)?;
        check(unsafe {
            ltxv_dit_forward(self.handle, p_hs, code(&hs)?, p_enc, code(&enc)?, p_ts as *const f32,
                             p_mask as *const f32, b as c_int, s as c_int, k as c_int, num_frames as c_int,
                             height as c_int, width as c_int,
                             scale.as_ref().map_or(std::ptr::null(), |v| v.as_ptr()), p_coords,
                             slm.as_ref().map_or(std::ptr::null(), |v| v.as_ptr()), p_out as *mut c_void,
                             code(&out)?, stream)
        })?;
        Ok(out)
    }
}

// =====================================================================================================================
// VaeLtxVideo  (t2v_pipeline.rs:91-103)
// =====================================================================================================================
pub struct B200Vae {
    handle: *mut ltxv_vae,
    cfg: VaeConfig,
    latents_mean: Tensor,
    latents_std: Tensor,
    dtype: DType,
}

unsafe impl Send for B200Vae {}

impl B200Vae {
    pub fn from_tensors(cfg: &AutoencoderKLLtxVideoConfig, tensors: &HashMap<String, Tensor>, dtype: DType,
                        dev: &Device) -> Result<Self> {
        if cfg.decoder_block_out_channels.len() != 3 || cfg.decoder_layers_per_block.len() != 4 {
            bail!("libltxv_b200 expects 3 decoder_block_out_channels and 4 decoder_layers_per_block");
        }
        let c = ltxv_vae_config {
            latent_channels: cfg.latent_channels as i32,
            out_channels: cfg.out_channels as i32,
            decoder_block_out_channels: [cfg.decoder_block_out_channels[0] as i32,
                                         cfg.decoder_block_out_channels[1] as i32,
                                         cfg.decoder_block_out_channels[2] as i32],
            decoder_layers_per_block: [cfg.decoder_layers_per_block[0] as i32, cfg.decoder_layers_per_block[1] as i32,
                                       cfg.decoder_layers_per_block[2] as i32, cfg.decoder_layers_per_block[3] as i32],
            patch_size: cfg.patch_size as i32,
            timestep_conditioning: cfg.timestep_conditioning as i32,
            scaling_factor: cfg.scaling_factor as f32,
        };
        let mut handle: *mut ltxv_vae = std::ptr::null_mut();
        check(unsafe { ltxv_vae_create(&c, device_index(dev)?, &mut handle) })?;
        // latents_mean / latents_std: from the weights when present, else from the config (vae.rs:1827-1838)
        let mean = match tensors.get("latents_mean") {
            Some(t) => t.to_device(dev)?.to_dtype(dtype)?,
            None => Tensor::new(cfg.latents_mean.as_slice(), dev)?.to_dtype(dtype)?,
        };
        let std = match tensors.get("latents_std") {
            Some(t) => t.to_device(dev)?.to_dtype(dtype)?,
            None => Tensor::new(cfg.latents_std.as_slice(), dev)?.to_dtype(dtype)?,
        };
        let me = Self {
            handle,
            cfg: VaeConfig { scaling_factor: cfg.scaling_factor as f32, timestep_conditioning: cfg.timestep_conditioning },
            latents_mean: mean,
            latents_std: std,
            dtype,
        };
        // `encoder.*`, `quant_conv.*`, `post_quant_conv.*` are accepted and ignored (the t2v path never runs them)
        push_tensors(tensors, |k, p, dt, sh, r| unsafe { ltxv_vae_load_tensor(me.handle, k, p, dt, sh, r) })?;
        check(unsafe { ltxv_vae_finalize(me.handle) })?;
        Ok(me)
    }

    pub fn raw(&self) -> *mut ltxv_vae {
        self.handle
    }
}

impl Drop for B200Vae {
    fn drop(&mut self) {
        unsafe { ltxv_vae_destroy(self.handle) };
    }
}

impl VaeLtxVideo for B200Vae {
    fn dtype(&self) -> DType {
        self.dtype
    }
    fn spatial_compression_ratio(&self) -> usize {
        32
    }
    fn temporal_compression_ratio(&self) -> usize {
        8
    }
    fn config(&self) -> &VaeConfig {
        &self.cfg
    }
    fn latents_mean(&self) -> &Tensor {
        &self.latents_mean
    }
    fn latents_std(&self) -> &Tensor {
        &self.latents_std
    }

    /// [B, C, F, H, W] -> [B, 3, 8F-7, 32H, 32W] in the VAE's dtype (vae.rs:2437-2463 -> :2101 -> :2037)
    fn decode(&self, latents: &Tensor, timestep: Option<&Tensor>) -> Result<Tensor> {
        let (b, _c, f, h, w) = latents.dims5()?;
        let z = match latents.dtype() {
            DType::F32 | DType::BF16 => latents.contiguous()?,
            _ => latents.to_dtype(DType::F32)?.contiguous()?,
        };
        let ts = match timestep {
            Some(t) => Some(t.to_dtype(DType::F32)?.flatten_all()?.contiguous()?),
            None => None,
        };
        let out = Tensor::zeros((b, 3, 8 * f - 7, 32 * h, 32 * w), self.dtype, z.device())?;
        let (p_z, stream) = raw_ptr(&z)?;
        let p_ts = match &ts {
            Some(t) => raw_ptr(t)?.0 as *const f32,
            None => std::ptr::null(),
        };
        let (p_out, _) = raw_ptr(&out)?;
        check(unsafe {
            ltxv_vae_decode(self.handle, p_z, code(&z)?, p_ts, b as c_int, f as c_int, h as c_int, w as c_int,
                            p_out as *mut c_void, code(&out)?, 0, stream)
        })?;
        Ok(out)
    }
}
