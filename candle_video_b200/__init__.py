"""candle_video_b200 -- host-side harness over libltxv_b200.so (the B200-native LTX-Video hot path).

The product is the C-ABI shared library (`include/ltxv.h`, sources under `csrc/`): hand-written sm_100a CUDA
(tcgen05/TMEM/TMA GEMM, implicit-GEMM conv3d and flash attention + 128-bit glue kernels) sequenced by C++ mirrors of
the reference's `LtxVideoTransformer3DModel`, `AutoencoderKLLtxVideo` and the hot part of `LtxPipeline::call`.

This Python module only binds that ABI with ctypes so that tests and `bench.py` can drive it exactly the way the
reference's Rust traits would (`VideoTransformer3D::forward`, `VaeLtxVideo::decode`, t2v_pipeline.rs:63-103).  torch is
used for device memory and streams only.  There is NO CPU fallback: if the library is missing or no B200 is present,
every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, Optional, Sequence, Tuple

__all__ = [
    "lib", "LtxvError", "DitConfig", "VaeConfig", "LtxVideoTransformer3DModel", "AutoencoderKLLtxVideo",
    "pack_latents", "unpack_latents", "video_coords", "guidance_euler_step", "denormalize_latents",
    "postprocess_video", "calculate_shift", "scheduler_set_timesteps", "PipelineParams", "pipeline_denoise",
    "pipeline_decode", "launch_count", "EXPORTED_SYMBOLS", "library_path",
]

F32, BF16 = 0, 1

EXPORTED_SYMBOLS = [
    "ltxv_last_error", "ltxv_version", "ltxv_launch_count",
    "ltxv_dit_config_preset", "ltxv_dit_create", "ltxv_dit_destroy", "ltxv_dit_load_tensor", "ltxv_dit_init_random",
    "ltxv_dit_finalize", "ltxv_dit_set_skip_blocks", "ltxv_dit_get_config", "ltxv_dit_forward",
    "ltxv_dit_forward_host", "ltxv_dit_prepare_context", "ltxv_dit_forward_ctx",
    "ltxv_vae_config_default", "ltxv_vae_create", "ltxv_vae_destroy", "ltxv_vae_load_tensor", "ltxv_vae_init_random",
    "ltxv_vae_finalize", "ltxv_vae_latents_mean", "ltxv_vae_latents_std", "ltxv_vae_spatial_compression_ratio",
    "ltxv_vae_temporal_compression_ratio", "ltxv_vae_decode", "ltxv_vae_decode_host",
    "ltxv_pack_latents", "ltxv_unpack_latents", "ltxv_video_coords", "ltxv_guidance_euler_step",
    "ltxv_denormalize_latents", "ltxv_postprocess_video", "ltxv_calculate_shift", "ltxv_scheduler_set_timesteps",
    "ltxv_pipeline_denoise", "ltxv_pipeline_decode", "ltxv_pipeline_denoise_host", "ltxv_pipeline_decode_host",
    "ltxv_profile_begin", "ltxv_profile_end", "ltxv_trace_begin", "ltxv_trace_end", "ltxv_causal_conv3d", "ltxv_set_option", "ltxv_get_option",
    "ltxv_comm_create", "ltxv_comm_destroy", "ltxv_comm_get_handle", "ltxv_comm_open", "ltxv_comm_barrier",
    "ltxv_parallel_plan", "ltxv_pipeline_denoise_parallel", "ltxv_pipeline_denoise_parallel_stochastic",
    "ltxv_vae_set_comm",
    "ltxv_remap_official_key_raw", "ltxv_remap_official_key", "ltxv_safetensors_list",
    "ltxv_dit_load_safetensors", "ltxv_vae_load_safetensors",
    "ltxv_vae_tiling_default", "ltxv_vae_decode_tiled",
    "ltxv_scheduler_step_stochastic", "ltxv_decode_noise_blend", "ltxv_pipeline_denoise_stochastic",
    "ltxv_pipeline_decode_noisy",
    "ltxv_vae_encoder_config_default", "ltxv_vae_enable_encoder", "ltxv_vae_encode_dims", "ltxv_vae_encode",
    "ltxv_vae_encode_host", "ltxv_normalize_latents", "ltxv_frames_to_u8", "ltxv_pipeline_decode_host_u8",
    "ltxv_vae_encode_tiled",
]


class LtxvError(RuntimeError):
    pass


def library_path() -> Path:
    return Path(os.environ.get("LTXV_B200_LIB", Path(__file__).resolve().parent / "lib" / "libltxv_b200.so"))


class _DitConfigC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("in_channels", "out_channels", "patch_size", "patch_size_t",
                                         "num_attention_heads", "attention_head_dim", "cross_attention_dim",
                                         "num_layers", "caption_channels")] + \
               [("norm_eps", C.c_float), ("timestep_bf16_round", C.c_int32)]


class _VaeConfigC(C.Structure):
    _fields_ = [("latent_channels", C.c_int32), ("out_channels", C.c_int32),
                ("decoder_block_out_channels", C.c_int32 * 3), ("decoder_layers_per_block", C.c_int32 * 4),
                ("patch_size", C.c_int32), ("timestep_conditioning", C.c_int32), ("scaling_factor", C.c_float)]


class _VaeEncoderConfigC(C.Structure):
    _fields_ = [("in_channels", C.c_int32), ("latent_channels", C.c_int32), ("block_out_channels", C.c_int32 * 5),
                ("layers_per_block", C.c_int32 * 5), ("downsample_types", C.c_int32 * 4), ("patch_size", C.c_int32)]


class _PipelineParamsC(C.Structure):
    _fields_ = [("height", C.c_int32), ("width", C.c_int32), ("num_frames", C.c_int32), ("frame_rate", C.c_int32),
                ("num_inference_steps", C.c_int32), ("custom_sigmas", C.POINTER(C.c_float)),
                ("guidance_scale", C.c_float), ("guidance_rescale", C.c_float), ("stg_scale", C.c_float),
                ("skip_block_list", C.POINTER(C.c_int32)), ("num_skip_blocks", C.c_int32),
                ("has_shift_terminal", C.c_int32), ("shift_terminal", C.c_float), ("decode_timestep", C.c_float)]


def _load() -> C.CDLL:
    path = library_path()
    if not path.exists():
        raise LtxvError(
            f"{path} not found: build it with `python -m candle_video_b200.build` (nvcc, sm_100a). "
            "There is no CPU/PyTorch fallback for this path.")
    l = C.CDLL(str(path))
    vp, i32, i64, f32, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64
    fp = C.POINTER(C.c_float)
    l.ltxv_last_error.restype = C.c_char_p
    l.ltxv_version.restype = C.c_char_p
    l.ltxv_launch_count.restype = u64
    l.ltxv_dit_config_preset.argtypes = [C.c_char_p, C.POINTER(_DitConfigC)]
    l.ltxv_dit_create.argtypes = [C.POINTER(_DitConfigC), i32, C.POINTER(vp)]
    l.ltxv_dit_destroy.argtypes = [vp]
    l.ltxv_dit_destroy.restype = None
    l.ltxv_dit_load_tensor.argtypes = [vp, C.c_char_p, vp, i32, C.POINTER(i64), i32]
    l.ltxv_dit_init_random.argtypes = [vp, u64]
    l.ltxv_dit_finalize.argtypes = [vp]
    l.ltxv_dit_set_skip_blocks.argtypes = [vp, C.POINTER(C.c_int32), i32]
    l.ltxv_dit_get_config.argtypes = [vp, C.POINTER(_DitConfigC)]
    l.ltxv_dit_forward.argtypes = [vp, vp, i32, vp, i32, vp, vp, i32, i32, i32, i32, i32, i32, fp, vp, fp, vp, i32, vp]
    l.ltxv_dit_forward_host.argtypes = [vp, vp, i32, vp, i32, vp, vp, i32, i32, i32, i32, i32, i32, fp, vp, fp, vp, i32]
    l.ltxv_dit_prepare_context.argtypes = [vp, i32, vp, i32, vp, i32, vp]
    l.ltxv_dit_forward_ctx.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, i32, fp, vp, fp, vp, i32, vp]
    l.ltxv_vae_config_default.argtypes = [C.POINTER(_VaeConfigC)]
    l.ltxv_vae_create.argtypes = [C.POINTER(_VaeConfigC), i32, C.POINTER(vp)]
    l.ltxv_vae_destroy.argtypes = [vp]
    l.ltxv_vae_destroy.restype = None
    l.ltxv_vae_load_tensor.argtypes = [vp, C.c_char_p, vp, i32, C.POINTER(i64), i32]
    l.ltxv_vae_init_random.argtypes = [vp, u64]
    l.ltxv_vae_finalize.argtypes = [vp]
    l.ltxv_vae_latents_mean.argtypes = [vp]
    l.ltxv_vae_latents_mean.restype = vp
    l.ltxv_vae_latents_std.argtypes = [vp]
    l.ltxv_vae_latents_std.restype = vp
    l.ltxv_vae_spatial_compression_ratio.argtypes = [vp]
    l.ltxv_vae_temporal_compression_ratio.argtypes = [vp]
    l.ltxv_vae_decode.argtypes = [vp, vp, i32, vp, i32, i32, i32, i32, vp, i32, i32, vp]
    l.ltxv_vae_decode_host.argtypes = [vp, vp, i32, vp, i32, i32, i32, i32, vp, i32, i32]
    l.ltxv_pack_latents.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]
    l.ltxv_unpack_latents.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]
    l.ltxv_video_coords.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, vp]
    l.ltxv_guidance_euler_step.argtypes = [vp, vp, vp, vp, vp, i32, i64, f32, f32, f32, f32, f32, vp]
    l.ltxv_denormalize_latents.argtypes = [vp, vp, vp, vp, f32, i32, i32, i64, vp]
    l.ltxv_postprocess_video.argtypes = [vp, vp, i64, vp]
    l.ltxv_calculate_shift.argtypes = [i32, fp]
    l.ltxv_scheduler_set_timesteps.argtypes = [i32, fp, f32, i32, f32, fp, C.POINTER(i64)]
    l.ltxv_pipeline_denoise.argtypes = [vp, C.POINTER(_PipelineParamsC), vp, vp, vp, vp, vp, i32, i32, vp]
    l.ltxv_pipeline_decode.argtypes = [vp, C.POINTER(_PipelineParamsC), vp, vp, vp]
    l.ltxv_pipeline_denoise_host.argtypes = [vp, C.POINTER(_PipelineParamsC), vp, vp, vp, vp, vp, i32, i32]
    l.ltxv_pipeline_decode_host.argtypes = [vp, C.POINTER(_PipelineParamsC), vp, vp]
    l.ltxv_comm_create.argtypes = [i32, i32, i32, u64, C.POINTER(vp)]
    l.ltxv_comm_destroy.argtypes = [vp]
    l.ltxv_comm_destroy.restype = None
    l.ltxv_comm_get_handle.argtypes = [vp, vp]
    l.ltxv_comm_open.argtypes = [vp, vp]
    l.ltxv_comm_barrier.argtypes = [vp, vp]
    l.ltxv_parallel_plan.argtypes = [i32, i32, i32, i32, C.POINTER(C.c_int32)]
    l.ltxv_pipeline_denoise_parallel.argtypes = [vp, vp, C.POINTER(_PipelineParamsC), vp, vp, vp, vp, vp, i32, i32, vp]
    l.ltxv_pipeline_denoise_parallel_stochastic.argtypes = [vp, vp, C.POINTER(_PipelineParamsC), vp, vp, vp, vp, vp, i32, i32, vp, vp]
    l.ltxv_vae_set_comm.argtypes = [vp, vp]
    l.ltxv_remap_official_key_raw.argtypes = [C.c_char_p, C.c_char_p, u64]
    l.ltxv_remap_official_key.argtypes = [C.c_char_p, C.c_char_p, u64, C.POINTER(C.c_int32)]
    l.ltxv_safetensors_list.argtypes = [C.c_char_p, C.c_char_p, u64, C.POINTER(C.c_int32)]
    l.ltxv_dit_load_safetensors.argtypes = [vp, C.c_char_p, i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    l.ltxv_vae_load_safetensors.argtypes = [vp, C.c_char_p, i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    l.ltxv_scheduler_step_stochastic.argtypes = [vp, vp, vp, C.c_int64, C.c_float, C.c_float, vp]
    l.ltxv_decode_noise_blend.argtypes = [vp, vp, C.c_float, C.c_int64, vp]
    l.ltxv_pipeline_denoise_stochastic.argtypes = [vp, C.POINTER(_PipelineParamsC), vp, vp, vp, vp, vp, i32, i32, vp, vp]
    l.ltxv_pipeline_decode_noisy.argtypes = [vp, C.POINTER(_PipelineParamsC), vp, vp, C.c_float, vp, vp]
    l.ltxv_vae_tiling_default.argtypes = [C.POINTER(_VaeTilingC)]
    l.ltxv_vae_decode_tiled.argtypes = [vp, vp, i32, vp, i32, i32, i32, i32, C.POINTER(_VaeTilingC), vp, i32, i32, vp]
    l.ltxv_vae_encoder_config_default.argtypes = [C.POINTER(_VaeEncoderConfigC)]
    l.ltxv_vae_enable_encoder.argtypes = [vp, C.POINTER(_VaeEncoderConfigC)]
    l.ltxv_vae_encode_dims.argtypes = [vp, i32, i32, i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    l.ltxv_vae_encode.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp]
    l.ltxv_vae_encode_host.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
    l.ltxv_vae_encode_tiled.argtypes = [vp, vp, i32, i32, i32, i32, i32, C.POINTER(_VaeTilingC), i32, vp, vp]
    l.ltxv_normalize_latents.argtypes = [vp, vp, vp, vp, f32, i32, i32, i64, vp]
    l.ltxv_frames_to_u8.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    l.ltxv_pipeline_decode_host_u8.argtypes = [vp, C.POINTER(_PipelineParamsC), vp, vp]
    l.ltxv_profile_begin.argtypes = []
    l.ltxv_trace_begin.argtypes = []
    l.ltxv_set_option.argtypes = [C.c_char_p, i32]
    l.ltxv_get_option.argtypes = [C.c_char_p, C.POINTER(C.c_int32)]
    l.ltxv_causal_conv3d.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]
    l.ltxv_trace_end.argtypes = [C.c_char_p, u64]
    l.ltxv_profile_end.argtypes = [C.POINTER(u64), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    return l


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise LtxvError(lib().ltxv_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(lib().ltxv_launch_count())


# ------------------------------------------------------------------------------------------------------------
# torch plumbing helpers
# ------------------------------------------------------------------------------------------------------------
def _torch():
    import torch
    return torch


def _dtype_code(t) -> int:
    torch = _torch()
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise LtxvError(f"unsupported tensor dtype {t.dtype}; expected float32 or bfloat16")


def _dev(t, name: str):
    if t is None:
        return None
    if not t.is_cuda:
        raise LtxvError(f"{name} must be a CUDA tensor (no CPU path exists)")
    return t.contiguous()


def _ptr(t) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return _torch().cuda.current_stream().cuda_stream


def _f3(v: Optional[Sequence[float]]):
    if v is None:
        return None
    return (C.c_float * 3)(*[float(x) for x in v])


# ------------------------------------------------------------------------------------------------------------
# DiT
# ------------------------------------------------------------------------------------------------------------
@dataclass
class DitConfig:
    """LtxVideoTransformer3DModelConfig (ltx_transformer.rs:22-59)."""
    in_channels: int = 128
    out_channels: int = 128
    patch_size: int = 1
    patch_size_t: int = 1
    num_attention_heads: int = 32
    attention_head_dim: int = 64
    cross_attention_dim: int = 2048
    num_layers: int = 28
    caption_channels: int = 4096
    norm_eps: float = 1e-6
    timestep_bf16_round: bool = True

    def to_c(self) -> _DitConfigC:
        return _DitConfigC(self.in_channels, self.out_channels, self.patch_size, self.patch_size_t,
                           self.num_attention_heads, self.attention_head_dim, self.cross_attention_dim,
                           self.num_layers, self.caption_channels, self.norm_eps, int(self.timestep_bf16_round))

    @staticmethod
    def preset(name: str) -> "DitConfig":
        c = _DitConfigC()
        _check(lib().ltxv_dit_config_preset(name.encode(), C.byref(c)))
        return DitConfig(c.in_channels, c.out_channels, c.patch_size, c.patch_size_t, c.num_attention_heads,
                         c.attention_head_dim, c.cross_attention_dim, c.num_layers, c.caption_channels, c.norm_eps,
                         bool(c.timestep_bf16_round))


def _load_state_dict(load_fn, handle, sd: Dict[str, "object"]) -> None:
    torch = _torch()
    for key, t in sd.items():
        t = t.detach()
        if t.dtype not in (torch.float32, torch.bfloat16):
            t = t.to(torch.float32)
        t = t.contiguous()
        shape = (C.c_int64 * max(t.dim(), 1))(*t.shape) if t.dim() else (C.c_int64 * 1)(0)
        _check(load_fn(handle, key.encode(), t.data_ptr(), _dtype_code(t), shape, t.dim()))


class LtxVideoTransformer3DModel:
    """Binds `ltxv_dit_*`; mirrors LtxVideoTransformer3DModel + trait VideoTransformer3D (ltx_transformer.rs:1175-1215)."""

    def __init__(self, config: DitConfig, device: int = 0):
        self.config = config
        self._h = C.c_void_p()
        cc = config.to_c()
        _check(lib().ltxv_dit_create(C.byref(cc), device, C.byref(self._h)))
        self.device = device

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                lib().ltxv_dit_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def load_state_dict(self, sd) -> None:
        """Tensors keyed by the reference's VarBuilder names (host or CUDA, f32 or bf16)."""
        _load_state_dict(lib().ltxv_dit_load_tensor, self._h, sd)
        _check(lib().ltxv_dit_finalize(self._h))

    def load_safetensors(self, path, official: bool = False):
        """Load a .safetensors file / diffusers directory / sharded directory; official=True remaps the unified-file
        key names (weight_format.rs) and takes the transformer tensors only.  Returns (loaded, ignored)."""
        a, b = C.c_int32(), C.c_int32()
        _check(lib().ltxv_dit_load_safetensors(self._h, str(path).encode(), int(official), C.byref(a), C.byref(b)))
        _check(lib().ltxv_dit_finalize(self._h))
        return a.value, b.value

    def init_random(self, seed: int = 0) -> None:
        _check(lib().ltxv_dit_init_random(self._h, seed))

    def set_skip_block_list(self, blocks: Sequence[int]) -> None:
        arr = (C.c_int32 * max(len(blocks), 1))(*blocks)
        _check(lib().ltxv_dit_set_skip_blocks(self._h, arr, len(blocks)))

    def forward(self, hidden_states, encoder_hidden_states, timestep, encoder_attention_mask, num_frames: int,
                height: int, width: int, rope_interpolation_scale: Optional[Tuple[float, float, float]] = None,
                video_coords=None, skip_layer_mask=None, out_dtype=None):
        """VideoTransformer3D::forward (t2v_pipeline.rs:63-83) on CUDA tensors; returns [B,S,out_channels]."""
        torch = _torch()
        hs = _dev(hidden_states, "hidden_states")
        enc = _dev(encoder_hidden_states, "encoder_hidden_states")
        B, S, _ = hs.shape
        K = enc.shape[1]
        ts = _dev(timestep, "timestep").to(torch.float32).reshape(-1).contiguous()
        if ts.numel() != B:
            raise LtxvError(f"timestep must have {B} entries")
        mask = None if encoder_attention_mask is None else _dev(encoder_attention_mask, "mask").to(torch.float32).contiguous()
        coords = None if video_coords is None else _dev(video_coords, "video_coords").to(torch.float32).contiguous()
        slm = None
        if skip_layer_mask is not None:
            m = skip_layer_mask.detach().to("cpu", torch.float32).contiguous()
            slm = (C.c_float * m.numel())(*m.flatten().tolist())
        odt = hs.dtype if out_dtype is None else out_dtype
        out = torch.empty((B, S, self.config.out_channels), dtype=odt, device=hs.device)
        _check(lib().ltxv_dit_forward(self._h, _ptr(hs), _dtype_code(hs), _ptr(enc), _dtype_code(enc), _ptr(ts),
                                      _ptr(mask), B, S, K, num_frames, height, width, _f3(rope_interpolation_scale),
                                      _ptr(coords), slm, _ptr(out), _dtype_code(out), _stream()))
        return out

    def forward_host(self, hidden_states, encoder_hidden_states, timestep, encoder_attention_mask, num_frames: int,
                     height: int, width: int, rope_interpolation_scale=None, video_coords=None, skip_layer_mask=None):
        """Same call on HOST (CPU torch) tensors through ltxv_dit_forward_host: H2D + compute + D2H + sync."""
        torch = _torch()
        hs = hidden_states.contiguous()
        enc = encoder_hidden_states.contiguous()
        B, S, _ = hs.shape
        K = enc.shape[1]
        ts = timestep.to(torch.float32).reshape(-1).contiguous()
        mask = None if encoder_attention_mask is None else encoder_attention_mask.to(torch.float32).contiguous()
        coords = None if video_coords is None else video_coords.to(torch.float32).contiguous()
        slm = None
        if skip_layer_mask is not None:
            m = skip_layer_mask.to(torch.float32).contiguous()
            slm = (C.c_float * m.numel())(*m.flatten().tolist())
        out = torch.empty((B, S, self.config.out_channels), dtype=hs.dtype)
        _check(lib().ltxv_dit_forward_host(self._h, _ptr(hs), _dtype_code(hs), _ptr(enc), _dtype_code(enc), _ptr(ts),
                                           _ptr(mask), B, S, K, num_frames, height, width,
                                           _f3(rope_interpolation_scale), _ptr(coords), slm, _ptr(out),
                                           _dtype_code(out)))
        return out

    def prepare_context(self, slot: int, encoder_hidden_states, encoder_attention_mask) -> None:
        torch = _torch()
        enc = _dev(encoder_hidden_states, "encoder_hidden_states")
        if enc.dim() == 3:
            enc = enc[0]
        mask = None
        if encoder_attention_mask is not None:
            mask = _dev(encoder_attention_mask, "mask").to(torch.float32).reshape(-1).contiguous()
        _check(lib().ltxv_dit_prepare_context(self._h, slot, _ptr(enc), _dtype_code(enc), _ptr(mask), enc.shape[0],
                                              _stream()))

    def forward_ctx(self, slot: int, hidden_states, timestep, num_frames: int, height: int, width: int,
                    rope_interpolation_scale=None, video_coords=None, skip_layer_mask=None, out=None):
        torch = _torch()
        hs = _dev(hidden_states, "hidden_states")
        if hs.dim() == 3:
            hs = hs[0]
        S = hs.shape[0]
        ts = _dev(timestep, "timestep").to(torch.float32).reshape(-1).contiguous()
        coords = None if video_coords is None else _dev(video_coords, "video_coords").to(torch.float32).reshape(-1, 3).contiguous()
        slm = None
        if skip_layer_mask is not None:
            m = skip_layer_mask.detach().to("cpu", torch.float32).contiguous()
            slm = (C.c_float * m.numel())(*m.flatten().tolist())
        if out is None:
            out = torch.empty((S, self.config.out_channels), dtype=torch.float32, device=hs.device)
        _check(lib().ltxv_dit_forward_ctx(self._h, slot, _ptr(hs), _dtype_code(hs), _ptr(ts), S, num_frames, height,
                                          width, _f3(rope_interpolation_scale), _ptr(coords), slm, _ptr(out),
                                          _dtype_code(out), _stream()))
        return out


# ------------------------------------------------------------------------------------------------------------
# VAE
# ------------------------------------------------------------------------------------------------------------
@dataclass
class VaeConfig:
    """AutoencoderKLLtxVideoConfig, decoder fields (vae.rs:30-103)."""
    latent_channels: int = 128
    out_channels: int = 3
    decoder_block_out_channels: Tuple[int, int, int] = (256, 512, 1024)
    decoder_layers_per_block: Tuple[int, int, int, int] = (5, 5, 5, 5)
    patch_size: int = 4
    timestep_conditioning: bool = True
    scaling_factor: float = 1.0

    def to_c(self) -> _VaeConfigC:
        return _VaeConfigC(self.latent_channels, self.out_channels, (C.c_int32 * 3)(*self.decoder_block_out_channels),
                           (C.c_int32 * 4)(*self.decoder_layers_per_block), self.patch_size,
                           int(self.timestep_conditioning), self.scaling_factor)


DOWNSAMPLE_CODES = {"spatial": 1, "temporal": 2, "spatiotemporal": 3}  # LTXV_DOWN_*, DownsampleType (vae.rs:469-493)


@dataclass
class VaeEncoderConfig:
    """AutoencoderKLLtxVideoConfig, encoder fields (vae.rs:30-103); the LTX-Video 0.9.5 layout."""
    in_channels: int = 3
    latent_channels: int = 128
    block_out_channels: Tuple[int, int, int, int, int] = (128, 256, 512, 1024, 2048)
    layers_per_block: Tuple[int, int, int, int, int] = (4, 6, 6, 2, 2)
    downsample_types: Tuple[str, str, str, str] = ("spatial", "temporal", "spatiotemporal", "spatiotemporal")
    patch_size: int = 4

    def to_c(self) -> _VaeEncoderConfigC:
        try:
            codes = [DOWNSAMPLE_CODES[t] for t in self.downsample_types]
        except KeyError as e:
            raise LtxvError(f"unsupported downsample type {e}; pixel-unshuffle types only (spatial/temporal/spatiotemporal)")
        return _VaeEncoderConfigC(self.in_channels, self.latent_channels, (C.c_int32 * 5)(*self.block_out_channels),
                                  (C.c_int32 * 5)(*self.layers_per_block), (C.c_int32 * 4)(*codes), self.patch_size)


class _VaeTilingC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "use_tiling", "use_framewise_decoding", "tile_sample_min_height", "tile_sample_min_width",
        "tile_sample_min_num_frames", "tile_sample_stride_height", "tile_sample_stride_width",
        "tile_sample_stride_num_frames")]


@dataclass
class VaeTiling:
    """Tiling knobs of AutoencoderKLLtxVideo (vae.rs:1848-1861 defaults, enable_tiling :1870-1898), sample space."""
    use_tiling: bool = True
    use_framewise_decoding: bool = True
    tile_sample_min_height: int = 512
    tile_sample_min_width: int = 512
    tile_sample_min_num_frames: int = 16
    tile_sample_stride_height: int = 384
    tile_sample_stride_width: int = 384
    tile_sample_stride_num_frames: int = 8

    def to_c(self) -> _VaeTilingC:
        return _VaeTilingC(int(self.use_tiling), int(self.use_framewise_decoding), self.tile_sample_min_height,
                           self.tile_sample_min_width, self.tile_sample_min_num_frames,
                           self.tile_sample_stride_height, self.tile_sample_stride_width,
                           self.tile_sample_stride_num_frames)


class AutoencoderKLLtxVideo:
    """Binds `ltxv_vae_*`; mirrors AutoencoderKLLtxVideo::decode + trait VaeLtxVideo (vae.rs:2101, :2437-2463)."""

    def __init__(self, config: VaeConfig, device: int = 0):
        self.config = config
        self._h = C.c_void_p()
        cc = config.to_c()
        _check(lib().ltxv_vae_create(C.byref(cc), device, C.byref(self._h)))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                lib().ltxv_vae_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def load_state_dict(self, sd) -> None:
        _load_state_dict(lib().ltxv_vae_load_tensor, self._h, sd)
        _check(lib().ltxv_vae_finalize(self._h))

    def load_safetensors(self, path, official: bool = False):
        a, b = C.c_int32(), C.c_int32()
        _check(lib().ltxv_vae_load_safetensors(self._h, str(path).encode(), int(official), C.byref(a), C.byref(b)))
        _check(lib().ltxv_vae_finalize(self._h))
        return a.value, b.value

    def init_random(self, seed: int = 0) -> None:
        _check(lib().ltxv_vae_init_random(self._h, seed))

    @property
    def spatial_compression_ratio(self) -> int:
        return lib().ltxv_vae_spatial_compression_ratio(self._h)

    @property
    def temporal_compression_ratio(self) -> int:
        return lib().ltxv_vae_temporal_compression_ratio(self._h)

    def decode(self, latents, timestep=None, postprocess: bool = False, out_dtype=None):
        """VaeLtxVideo::decode (t2v_pipeline.rs:102): [B,128,F,H,W] -> [B,3,8F-7,32H,32W] (CUDA tensors)."""
        torch = _torch()
        z = _dev(latents, "latents")
        B, _, F, H, W = z.shape
        ts = None if timestep is None else _dev(timestep, "timestep").to(torch.float32).reshape(-1).contiguous()
        odt = torch.float32 if out_dtype is None else out_dtype
        out = torch.empty((B, 3, 8 * F - 7, 32 * H, 32 * W), dtype=odt, device=z.device)
        _check(lib().ltxv_vae_decode(self._h, _ptr(z), _dtype_code(z), _ptr(ts), B, F, H, W, _ptr(out),
                                     _dtype_code(out), int(postprocess), _stream()))
        return out

    def decode_tiled(self, latents, timestep=None, tiling: "Optional[VaeTiling]" = None, postprocess: bool = False,
                     out_dtype=None):
        """decode_z of the reference with its tiling dispatch (vae.rs:2037-2066); tiling=None is the library default
        (512/384 px tiles, 16/8 frames)."""
        torch = _torch()
        z = _dev(latents, "latents")
        B, _, F, H, W = z.shape
        ts = None if timestep is None else _dev(timestep, "timestep").to(torch.float32).reshape(-1).contiguous()
        odt = torch.float32 if out_dtype is None else out_dtype
        out = torch.empty((B, 3, 8 * F - 7, 32 * H, 32 * W), dtype=odt, device=z.device)
        tp = (tiling or VaeTiling()).to_c()
        _check(lib().ltxv_vae_decode_tiled(self._h, _ptr(z), _dtype_code(z), _ptr(ts), B, F, H, W, C.byref(tp),
                                           _ptr(out), _dtype_code(out), int(postprocess), _stream()))
        return out

    # ---- encoder half ----
    def enable_encoder(self, config: "Optional[VaeEncoderConfig]" = None) -> None:
        """Build the encoder (vae.rs:1772-1784): `encoder.*` keys are then loaded instead of ignored."""
        self.encoder_config = config or VaeEncoderConfig()
        cc = self.encoder_config.to_c()
        _check(lib().ltxv_vae_enable_encoder(self._h, C.byref(cc)))

    def encode_dims(self, num_frames: int, height: int, width: int) -> Tuple[int, int, int]:
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        _check(lib().ltxv_vae_encode_dims(self._h, num_frames, height, width, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def encode(self, video):
        """AutoencoderKLLtxVideo::encode (vae.rs:2070-2099): [B,3,F,H,W] in [-1,1] (CUDA, f32/bf16) -> moments f32
        [B, 2*latent, F', H', W']; `moments[:, :latent]` is the posterior mean (`.mode()`), the rest the logvar."""
        torch = _torch()
        x = _dev(video, "video")
        B, Cin, F, H, W = x.shape
        if Cin != 3:
            raise LtxvError("video must have 3 channels")
        fl, hl, wl = self.encode_dims(F, H, W)
        out = torch.empty((B, 2 * self.config.latent_channels, fl, hl, wl), dtype=torch.float32, device=x.device)
        _check(lib().ltxv_vae_encode(self._h, _ptr(x), _dtype_code(x), B, F, H, W, _ptr(out), _stream()))
        return out

    def encode_tiled(self, video, tiling: "Optional[VaeTiling]" = None, use_framewise_encoding: bool = False):
        """encode_z of the reference with its tiling dispatch (vae.rs:2017-2034); tiling=None is the library default
        (512/384 px tiles; framewise encoding off, vae.rs:1856-1858)."""
        torch = _torch()
        x = _dev(video, "video")
        B, Cin, F, H, W = x.shape
        if Cin != 3:
            raise LtxvError("video must have 3 channels")
        fl, hl, wl = self.encode_dims(F, H, W)
        out = torch.empty((B, 2 * self.config.latent_channels, fl, hl, wl), dtype=torch.float32, device=x.device)
        tp = (tiling or VaeTiling()).to_c()
        _check(lib().ltxv_vae_encode_tiled(self._h, _ptr(x), _dtype_code(x), B, F, H, W, C.byref(tp),
                                           int(use_framewise_encoding), _ptr(out), _stream()))
        return out

    def encode_host(self, video):
        torch = _torch()
        x = video.contiguous()
        B, Cin, F, H, W = x.shape
        if Cin != 3:
            raise LtxvError("video must have 3 channels")
        fl, hl, wl = self.encode_dims(F, H, W)
        out = torch.empty((B, 2 * self.config.latent_channels, fl, hl, wl), dtype=torch.float32)
        _check(lib().ltxv_vae_encode_host(self._h, _ptr(x), _dtype_code(x), B, F, H, W, _ptr(out)))
        return out

    def decode_host(self, latents, timestep=None, postprocess: bool = False):
        torch = _torch()
        z = latents.contiguous()
        B, _, F, H, W = z.shape
        ts = None if timestep is None else timestep.to(torch.float32).reshape(-1).contiguous()
        out = torch.empty((B, 3, 8 * F - 7, 32 * H, 32 * W), dtype=torch.float32)
        _check(lib().ltxv_vae_decode_host(self._h, _ptr(z), _dtype_code(z), _ptr(ts), B, F, H, W, _ptr(out), F32,
                                          int(postprocess)))
        return out


# ------------------------------------------------------------------------------------------------------------
# pipeline glue (t2v_pipeline.rs)
# ------------------------------------------------------------------------------------------------------------
def pack_latents(latents, patch_size: int = 1, patch_size_t: int = 1):
    """LtxPipeline::pack_latents (t2v_pipeline.rs:474-504), f32 CUDA tensor [B,C,F,H,W] -> [B,S,D]."""
    torch = _torch()
    x = _dev(latents, "latents").to(torch.float32).contiguous()
    B, Cc, F, H, W = x.shape
    p, pt = patch_size, patch_size_t
    if p <= 0 or pt <= 0 or F % pt or H % p or W % p:
        raise LtxvError("latents shape not divisible by patch sizes")
    out = torch.empty((B, (F // pt) * (H // p) * (W // p), Cc * pt * p * p), dtype=torch.float32, device=x.device)
    _check(lib().ltxv_pack_latents(_ptr(x), _ptr(out), B, Cc, F, H, W, p, pt, _stream()))
    return out


def unpack_latents(latents, num_frames: int, height: int, width: int, patch_size: int = 1, patch_size_t: int = 1):
    """LtxPipeline::unpack_latents (t2v_pipeline.rs:506-550); num_frames/height/width are the patched grid dims."""
    torch = _torch()
    x = _dev(latents, "latents").to(torch.float32).contiguous()
    B, S, D = x.shape
    p, pt = patch_size, patch_size_t
    if D % (pt * p * p):
        raise LtxvError("D is not divisible by (pt*p*p)")
    Cc = D // (pt * p * p)
    F, H, W = num_frames * pt, height * p, width * p
    out = torch.empty((B, Cc, F, H, W), dtype=torch.float32, device=x.device)
    _check(lib().ltxv_unpack_latents(_ptr(x), _ptr(out), B, Cc, F, H, W, p, pt, _stream()))
    return out


def video_coords(batch: int, f: int, h: int, w: int, frame_rate: int, device="cuda", ts_ratio: int = 8,
                 sp_ratio: int = 32):
    torch = _torch()
    out = torch.empty((batch, f * h * w, 3), dtype=torch.float32, device=device)
    _check(lib().ltxv_video_coords(_ptr(out), batch, f, h, w, ts_ratio, sp_ratio, frame_rate, _stream()))
    return out


def guidance_euler_step(cond, uncond, perturbed, latents, guidance_scale: float, guidance_rescale: float,
                        stg_scale: float, sigma: float, sigma_next: float, return_noise_pred: bool = False):
    """CFG/STG combine + Euler update; `latents` (f32 CUDA, [B,S,C]) is updated in place."""
    torch = _torch()
    c = _dev(cond, "cond")
    B = c.shape[0]
    n = c[0].numel()
    u = _dev(uncond, "uncond")
    p = _dev(perturbed, "perturbed")
    for t in (c, u, p, latents):
        if t is not None and t.dtype != torch.float32:
            raise LtxvError("guidance tensors must be float32")
    noise = torch.empty_like(c) if return_noise_pred else None
    _check(lib().ltxv_guidance_euler_step(_ptr(c), _ptr(u), _ptr(p), _ptr(latents), _ptr(noise), B, n,
                                          guidance_scale, guidance_rescale, stg_scale, sigma, sigma_next, _stream()))
    return noise


def denormalize_latents(latents, mean, std, scaling_factor: float):
    torch = _torch()
    x = _dev(latents, "latents").to(torch.float32).contiguous()
    B, Cc = x.shape[:2]
    m = _dev(mean, "mean").to(torch.float32).contiguous()
    s = _dev(std, "std").to(torch.float32).contiguous()
    out = torch.empty_like(x)
    _check(lib().ltxv_denormalize_latents(_ptr(x), _ptr(out), _ptr(m), _ptr(s), scaling_factor, B, Cc,
                                          x[0, 0].numel(), _stream()))
    return out


def normalize_latents(latents, mean, std, scaling_factor: float):
    """LtxPipeline::normalize_latents (t2v_pipeline.rs:552-571)."""
    torch = _torch()
    x = _dev(latents, "latents").to(torch.float32).contiguous()
    B, Cc = x.shape[:2]
    m = _dev(mean, "mean").to(torch.float32).contiguous()
    s = _dev(std, "std").to(torch.float32).contiguous()
    out = torch.empty_like(x)
    _check(lib().ltxv_normalize_latents(_ptr(x), _ptr(out), _ptr(m), _ptr(s), scaling_factor, B, Cc,
                                        x[0, 0].numel(), _stream()))
    return out


def postprocess_video(video):
    torch = _torch()
    x = _dev(video, "video").to(torch.float32).contiguous()
    out = torch.empty_like(x)
    _check(lib().ltxv_postprocess_video(_ptr(x), _ptr(out), x.numel(), _stream()))
    return out


def calculate_shift(seq_len: int) -> float:
    v = C.c_float()
    _check(lib().ltxv_calculate_shift(seq_len, C.byref(v)))
    return float(v.value)


def scheduler_set_timesteps(num_steps: int, mu: float, sigmas: Optional[Sequence[float]] = None,
                            shift_terminal: Optional[float] = 0.1):
    cs = None if sigmas is None else (C.c_float * num_steps)(*sigmas)
    so = (C.c_float * (num_steps + 1))()
    to = (C.c_int64 * num_steps)()
    _check(lib().ltxv_scheduler_set_timesteps(num_steps, cs, mu, int(shift_terminal is not None),
                                              float(shift_terminal or 0.0), so, to))
    return list(so), list(to)


@dataclass
class PipelineParams:
    height: int = 512
    width: int = 768
    num_frames: int = 97
    frame_rate: int = 25
    num_inference_steps: int = 40
    custom_sigmas: Optional[Sequence[float]] = None
    guidance_scale: float = 3.0
    guidance_rescale: float = 0.0
    stg_scale: float = 0.0
    skip_block_list: Optional[Sequence[int]] = None
    shift_terminal: Optional[float] = 0.1
    decode_timestep: float = 0.05

    def to_c(self):
        keep = []
        cs = None
        if self.custom_sigmas is not None:
            cs = (C.c_float * len(self.custom_sigmas))(*self.custom_sigmas)
            keep.append(cs)
        sb = None
        nsb = 0
        if self.skip_block_list is not None:
            nsb = len(self.skip_block_list)
            sb = (C.c_int32 * max(nsb, 1))(*self.skip_block_list)
            keep.append(sb)
        p = _PipelineParamsC(self.height, self.width, self.num_frames, self.frame_rate, self.num_inference_steps,
                             C.cast(cs, C.POINTER(C.c_float)) if cs is not None else None,
                             self.guidance_scale, self.guidance_rescale, self.stg_scale,
                             C.cast(sb, C.POINTER(C.c_int32)) if sb is not None else None, nsb,
                             int(self.shift_terminal is not None), float(self.shift_terminal or 0.0),
                             self.decode_timestep)
        return p, keep


def pipeline_denoise(dit: LtxVideoTransformer3DModel, params: PipelineParams, latents, prompt_embeds, prompt_mask,
                     negative_embeds=None, negative_mask=None):
    """Denoise loop of LtxPipeline::call (t2v_pipeline.rs:860-994); `latents` f32 CUDA [S,128], updated in place."""
    torch = _torch()
    p, keep = params.to_c()
    pe = _dev(prompt_embeds, "prompt_embeds")
    pe = pe[0] if pe.dim() == 3 else pe
    ne = _dev(negative_embeds, "negative_embeds")
    if ne is not None and ne.dim() == 3:
        ne = ne[0]
    pm = None if prompt_mask is None else _dev(prompt_mask, "prompt_mask").to(torch.float32).reshape(-1).contiguous()
    nm = None if negative_mask is None else _dev(negative_mask, "negative_mask").to(torch.float32).reshape(-1).contiguous()
    if latents.dtype != torch.float32 or not latents.is_cuda or not latents.is_contiguous():
        raise LtxvError("latents must be a contiguous float32 CUDA tensor")
    _check(lib().ltxv_pipeline_denoise(dit._h, C.byref(p), _ptr(latents), _ptr(pe), _ptr(pm), _ptr(ne), _ptr(nm),
                                       _dtype_code(pe), pe.shape[0], _stream()))
    return latents


def pipeline_decode(vae: AutoencoderKLLtxVideo, params: PipelineParams, latents, decode_noise=None,
                    decode_noise_scale: float = 0.0, out=None):
    """Decode branch (t2v_pipeline.rs:1000-1072).  decode_noise: f32 CUDA [128, F, H, W] drawn by the caller, blended
    as (1 - scale) latents + scale noise before the VAE (:1049-1062).  out: optional preallocated f32 CUDA
    [3, frames, height, width] (a fresh 458 MB tensor per call otherwise costs a cudaMalloc when the caching
    allocator has no free block)."""
    torch = _torch()
    p, keep = params.to_c()
    f = (params.num_frames - 1) // 8 + 1
    shape = (3, 8 * f - 7, params.height, params.width)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=latents.device)
    elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_cuda or not out.is_contiguous():
        raise LtxvError(f"out must be a contiguous float32 CUDA tensor of shape {shape}")
    if decode_noise is None and decode_noise_scale == 0.0:
        _check(lib().ltxv_pipeline_decode(vae._h, C.byref(p), _ptr(latents), _ptr(out), _stream()))
    else:
        nz = None if decode_noise is None else _dev(decode_noise, "decode_noise").to(torch.float32).contiguous()
        _check(lib().ltxv_pipeline_decode_noisy(vae._h, C.byref(p), _ptr(latents), _ptr(nz), float(decode_noise_scale),
                                                _ptr(out), _stream()))
    return out


def scheduler_step_stochastic(latents, model_output, noise, sigma: float, sigma_next: float):
    """In place: x <- (1 - sigma_next) (x - sigma v) + sigma_next noise (scheduler.rs:557-575); f32 CUDA tensors."""
    _check(lib().ltxv_scheduler_step_stochastic(_ptr(latents), _ptr(model_output), _ptr(noise), latents.numel(),
                                                float(sigma), float(sigma_next), _stream()))
    return latents


def pipeline_denoise_stochastic(dit: LtxVideoTransformer3DModel, params: PipelineParams, latents, prompt_embeds,
                                prompt_mask, step_noise, negative_embeds=None, negative_mask=None):
    """pipeline_denoise with stochastic_sampling = true; step_noise f32 CUDA [num_inference_steps, S, 128]."""
    torch = _torch()
    p, keep = params.to_c()
    pe = _dev(prompt_embeds, "prompt_embeds")
    pe = pe[0] if pe.dim() == 3 else pe
    ne = _dev(negative_embeds, "negative_embeds")
    if ne is not None and ne.dim() == 3:
        ne = ne[0]
    pm = None if prompt_mask is None else _dev(prompt_mask, "prompt_mask").to(torch.float32).reshape(-1).contiguous()
    nm = None if negative_mask is None else _dev(negative_mask, "negative_mask").to(torch.float32).reshape(-1).contiguous()
    sn = _dev(step_noise, "step_noise")
    if sn.dtype != torch.float32 or not sn.is_contiguous() or sn.numel() != params.num_inference_steps * latents.numel():
        raise LtxvError("step_noise must be a contiguous float32 CUDA tensor [num_inference_steps, S, C]")
    if latents.dtype != torch.float32 or not latents.is_cuda or not latents.is_contiguous():
        raise LtxvError("latents must be a contiguous float32 CUDA tensor")
    _check(lib().ltxv_pipeline_denoise_stochastic(dit._h, C.byref(p), _ptr(latents), _ptr(pe), _ptr(pm), _ptr(ne),
                                                  _ptr(nm), _dtype_code(pe), pe.shape[0], _ptr(sn), _stream()))
    return latents


def pipeline_denoise_host(dit: LtxVideoTransformer3DModel, params: PipelineParams, latents, prompt_embeds, prompt_mask,
                          negative_embeds=None, negative_mask=None):
    """ltxv_pipeline_denoise_host: all tensors are HOST (CPU torch) tensors; latents f32 [S,128] updated in place."""
    torch = _torch()
    p, keep = params.to_c()
    pe = prompt_embeds[0] if prompt_embeds.dim() == 3 else prompt_embeds
    pe = pe.contiguous()
    ne = None
    if negative_embeds is not None:
        ne = (negative_embeds[0] if negative_embeds.dim() == 3 else negative_embeds).contiguous()
    pm = None if prompt_mask is None else prompt_mask.to(torch.float32).reshape(-1).contiguous()
    nm = None if negative_mask is None else negative_mask.to(torch.float32).reshape(-1).contiguous()
    if latents.is_cuda or latents.dtype != torch.float32 or not latents.is_contiguous():
        raise LtxvError("latents must be a contiguous float32 CPU tensor")
    _check(lib().ltxv_pipeline_denoise_host(dit._h, C.byref(p), _ptr(latents), _ptr(pe), _ptr(pm), _ptr(ne), _ptr(nm),
                                            _dtype_code(pe), pe.shape[0]))
    return latents


def pipeline_decode_host(vae: AutoencoderKLLtxVideo, params: PipelineParams, latents, out=None):
    torch = _torch()
    p, keep = params.to_c()
    f = (params.num_frames - 1) // 8 + 1
    if out is None:
        out = torch.empty((3, 8 * f - 7, params.height, params.width), dtype=torch.float32)
    _check(lib().ltxv_pipeline_decode_host(vae._h, C.byref(p), _ptr(latents.contiguous()), _ptr(out)))
    return out


def frames_to_u8(frames):
    """Output hand-off of the reference's example (main.rs:653-667): f32 CUDA [B,3,F,H,W] in 0..255 -> u8 [B,F,H,W,3]."""
    torch = _torch()
    x = _dev(frames, "frames").to(torch.float32).contiguous()
    B, Cc, F, H, W = x.shape
    if Cc != 3:
        raise LtxvError("frames must have 3 channels")
    out = torch.empty((B, F, H, W, 3), dtype=torch.uint8, device=x.device)
    _check(lib().ltxv_frames_to_u8(_ptr(x), _ptr(out), B, F, H, W, _stream()))
    return out


def pipeline_decode_host_u8(vae: AutoencoderKLLtxVideo, params: PipelineParams, latents, out=None):
    """pipeline_decode_host delivering host u8 [frames, height, width, 3] (a quarter of the D2H bytes)."""
    torch = _torch()
    p, keep = params.to_c()
    f = (params.num_frames - 1) // 8 + 1
    if out is None:
        out = torch.empty((8 * f - 7, params.height, params.width, 3), dtype=torch.uint8)
    _check(lib().ltxv_pipeline_decode_host_u8(vae._h, C.byref(p), _ptr(latents.contiguous()), _ptr(out)))
    return out


PROFILE_CLASSES = ("gemm", "conv3d", "attn_self", "attn_cross", "norm_modulate", "qk_norm_rope", "vae_prep", "other")


def profile_begin() -> None:
    _check(lib().ltxv_profile_begin())


def profile_end() -> Dict[str, Dict[str, float]]:
    n = (C.c_uint64 * 8)()
    ms = (C.c_double * 8)()
    fl = (C.c_double * 8)()
    _check(lib().ltxv_profile_end(n, ms, fl))
    # classes 4-7 are HBM-bound glue kernels: "flops" holds their algorithmic bytes (also exposed as "bytes")
    return {PROFILE_CLASSES[i]: {"launches": int(n[i]), "ms": float(ms[i]), "flops": float(fl[i]), "bytes": float(fl[i])}
            for i in range(8)}


def causal_conv3d(x, weight, bias=None, is_causal: bool = False):
    """LtxVideoCausalConv3d::forward (vae.rs:298-465), 3x3x3 stride 1: x [1,Cin,T,H,W], weight [Cout,Cin,3,3,3] (f32
    CUDA) -> [1,Cout,T,H,W] f32."""
    torch = _torch()
    x = _dev(x, "x").to(torch.float32).contiguous()
    w = _dev(weight, "weight").to(torch.float32).contiguous()
    b = None if bias is None else _dev(bias, "bias").to(torch.float32).contiguous()
    if x.dim() != 5 or x.shape[0] != 1 or w.dim() != 5 or tuple(w.shape[2:]) != (3, 3, 3) or w.shape[1] != x.shape[1]:
        raise LtxvError("causal_conv3d expects x [1,Cin,T,H,W] and weight [Cout,Cin,3,3,3]")
    _, Cin, T, H, W = x.shape
    Cout = w.shape[0]
    out = torch.empty((1, Cout, T, H, W), dtype=torch.float32, device=x.device)
    _check(lib().ltxv_causal_conv3d(_ptr(x), _ptr(w), _ptr(b), Cin, Cout, T, H, W, int(is_causal), _ptr(out), _stream()))
    return out


def set_option(name: str, value: int) -> None:
    """Experiment knob override (ltxv_set_option)."""
    _check(lib().ltxv_set_option(name.encode(), int(value)))


def get_option(name: str) -> int:
    v = C.c_int32()
    _check(lib().ltxv_get_option(name.encode(), C.byref(v)))
    return int(v.value)


def trace_begin() -> None:
    _check(lib().ltxv_trace_begin())


def trace_end() -> Dict[str, int]:
    """{kernel variant: launches} recorded since trace_begin (ltxv_trace_*; test hook)."""
    buf = C.create_string_buffer(1 << 16)
    _check(lib().ltxv_trace_end(buf, 1 << 16))
    out = {}
    for ln in buf.value.decode().splitlines():
        name, n = ln.rsplit(" ", 1)
        out[name] = int(n)
    return out


# ------------------------------------------------------------------------------------------------------------
# weight files (host only)
# ------------------------------------------------------------------------------------------------------------
def remap_official_key_raw(key: str) -> str:
    """KeyRemapper::remap_key (weight_format.rs:55-143)."""
    buf = C.create_string_buffer(1024)
    _check(lib().ltxv_remap_official_key_raw(key.encode(), buf, 1024))
    return buf.value.decode()


def remap_official_key(key: str):
    """(model key, component) as examples/ltx-video/main.rs:480-497 routes a unified-file tensor; component in
    {"other", "transformer", "vae"}."""
    buf = C.create_string_buffer(1024)
    c = C.c_int32()
    _check(lib().ltxv_remap_official_key(key.encode(), buf, 1024, C.byref(c)))
    return buf.value.decode(), ("other", "transformer", "vae")[c.value]


def safetensors_list(path) -> list:
    """[(name, dtype, shape, nbytes)] of a safetensors file / directory / sharded directory, sorted by name."""
    cap = 1 << 24
    buf = C.create_string_buffer(cap)
    n = C.c_int32()
    _check(lib().ltxv_safetensors_list(str(path).encode(), buf, cap, C.byref(n)))
    out = []
    for ln in buf.value.decode().splitlines():
        name, dtype, shape, nbytes = ln.rsplit(" ", 3)
        dims = [int(v) for v in shape.strip("[]").split(",") if v]
        out.append((name, dtype, dims, int(nbytes)))
    assert len(out) == n.value
    return out


# ------------------------------------------------------------------------------------------------------------
# multi-GPU (one process per GPU): peer-memory communicator bootstrapped over torch.distributed
# ------------------------------------------------------------------------------------------------------------
def parallel_plan(nranks: int, rank: int, seq_len: int, do_cfg: bool) -> Dict[str, int]:
    """Host-only sharding plan of pipeline_denoise_parallel (CFG branch groups x Ulysses token shards)."""
    out = (C.c_int32 * 6)()
    _check(lib().ltxv_parallel_plan(nranks, rank, seq_len, int(do_cfg), out))
    keys = ("cfg_groups", "sp_size", "branch", "sp_rank", "local_tokens", "token0")
    return dict(zip(keys, [int(v) for v in out]))


def exchange_handles(handle: bytes, group=None):
    """All-gather the 64-byte CUDA IPC handles over torch.distributed (gloo or nccl); returns nranks*64 bytes."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine, group=group)
    return b"".join(bytes(t.cpu().tolist()) for t in allh)


class PeerComm:
    """ltxv_comm: symmetric heap + flag barrier over NVLink peer memory (see include/ltxv.h, csrc/comm.h)."""

    def __init__(self, nranks: int, rank: int, device: int, heap_bytes: int = 3 << 30, group=None):
        self.nranks, self.rank = nranks, rank
        self._h = C.c_void_p()
        _check(lib().ltxv_comm_create(nranks, rank, device, heap_bytes, C.byref(self._h)))
        if nranks > 1:
            buf = (C.c_uint8 * 64)()
            _check(lib().ltxv_comm_get_handle(self._h, buf))
            allh = exchange_handles(bytes(buf), group)
            _check(lib().ltxv_comm_open(self._h, allh))

    def barrier(self) -> None:
        _check(lib().ltxv_comm_barrier(self._h, _stream()))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                lib().ltxv_comm_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


def pipeline_denoise_parallel(dit: LtxVideoTransformer3DModel, comm: PeerComm, params: PipelineParams, latents,
                              prompt_embeds, prompt_mask, negative_embeds=None, negative_mask=None, step_noise=None):
    """Denoise loop sharded over all ranks of `comm`: CFG branch split x Ulysses sequence parallelism.  step_noise:
    optional f32 CUDA [num_inference_steps, S, 128] (stochastic sampling with the caller's noise)."""
    torch = _torch()
    p, keep = params.to_c()
    pe = _dev(prompt_embeds, "prompt_embeds")
    pe = pe[0] if pe.dim() == 3 else pe
    ne = _dev(negative_embeds, "negative_embeds")
    if ne is not None and ne.dim() == 3:
        ne = ne[0]
    pm = None if prompt_mask is None else _dev(prompt_mask, "prompt_mask").to(torch.float32).reshape(-1).contiguous()
    nm = None if negative_mask is None else _dev(negative_mask, "negative_mask").to(torch.float32).reshape(-1).contiguous()
    if latents.dtype != torch.float32 or not latents.is_cuda or not latents.is_contiguous():
        raise LtxvError("latents must be a contiguous float32 CUDA tensor")
    if step_noise is not None:
        sn = _dev(step_noise, "step_noise")
        if sn.dtype != torch.float32 or sn.numel() != params.num_inference_steps * latents.numel():
            raise LtxvError("step_noise must be a contiguous float32 CUDA tensor [num_inference_steps, S, C]")
        _check(lib().ltxv_pipeline_denoise_parallel_stochastic(dit._h, comm._h, C.byref(p), _ptr(latents), _ptr(pe),
                                                               _ptr(pm), _ptr(ne), _ptr(nm), _dtype_code(pe),
                                                               pe.shape[0], _ptr(sn), _stream()))
        return latents
    _check(lib().ltxv_pipeline_denoise_parallel(dit._h, comm._h, C.byref(p), _ptr(latents), _ptr(pe), _ptr(pm),
                                                _ptr(ne), _ptr(nm), _dtype_code(pe), pe.shape[0], _stream()))
    return latents


def vae_set_comm(vae: AutoencoderKLLtxVideo, comm: Optional[PeerComm]) -> None:
    _check(lib().ltxv_vae_set_comm(vae._h, comm._h if comm is not None else None))
