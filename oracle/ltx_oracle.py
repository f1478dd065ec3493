"""CPU f32 ORACLE for the LTX-Video hot path of FerrisMind/candle-video  --  TEST INFRASTRUCTURE ONLY.

This file is a restatement, op for op, of the reference's Rust/Candle implementation of
  * the DiT forward            (src/models/ltx_video/ltx_transformer.rs)
  * the 3D-VAE decoder         (src/models/ltx_video/vae.rs, decoder half)
  * the pipeline glue          (src/models/ltx_video/t2v_pipeline.rs, scheduler.rs)
in plain torch-CPU float32 (torch's reshape/permute/broadcast semantics are Candle's).  Every function cites the
reference file:line it follows.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it; the product (candle_video_b200 + libltxv_b200.so) never does.

PARITY PINNING STATUS: "parity unpinned" for the floating-point tensors.  The reference cannot be built here
(no cargo/rustc, Candle is not vendored) and every numeric fixture of its test-suite (`gen_*.safetensors`) is
git-ignored and absent, so there is no golden tensor to check this restatement against.  What IS pinned (see
tests/test_oracle_kat.py): the reference's self-contained known-answer tests -- AdaLN value 0.1*1.01+0.001
(tests/verify_rope_parity.rs:646-733), attention scale 1/sqrt(64) (:472-511), pack/unpack identity
(tests/verify_pipeline_parity.rs:742-771), the closed-form video-coordinate code duplicated in
tests/verify_video_coords_parity.rs:39-97, CFG formula (tests/verify_cfg_parity.rs), scheduler closed forms
(src/models/ltx_video/scheduler.rs tests), preset constants (configs.rs:289-324).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
F32 = torch.float32


# =====================================================================================================
# Configs (ltx_transformer.rs:22-59, configs.rs:135-160; vae.rs:30-103, configs.rs:84-93)
# =====================================================================================================
@dataclass
class DitConfig:
    in_channels: int = 128
    out_channels: int = 128
    patch_size: int = 1
    patch_size_t: int = 1
    num_attention_heads: int = 32
    attention_head_dim: int = 64
    cross_attention_dim: int = 2048
    num_layers: int = 28
    norm_eps: float = 1e-6
    caption_channels: int = 4096

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim


def dit_config_2b() -> DitConfig:  # configs.rs:135-149
    return DitConfig()


def dit_config_13b() -> DitConfig:  # configs.rs:151-160
    return DitConfig(num_attention_heads=32, attention_head_dim=128, cross_attention_dim=4096, num_layers=48)


@dataclass
class VaeConfig:
    latent_channels: int = 128
    out_channels: int = 3
    decoder_block_out_channels: Tuple[int, ...] = (256, 512, 1024)
    decoder_layers_per_block: Tuple[int, ...] = (5, 5, 5, 5)
    decoder_upsample_factor: Tuple[int, ...] = (2, 2, 2)
    patch_size: int = 4
    patch_size_t: int = 1
    timestep_conditioning: bool = True
    scaling_factor: float = 1.0

    def stage_channels(self) -> List[int]:
        """Channel width of mid block and of each up block's output (vae.rs:1507-1567)."""
        boc = list(reversed(self.decoder_block_out_channels))
        upf = list(reversed(self.decoder_upsample_factor))
        return [boc[0]] + [boc[i] // upf[i] for i in range(len(boc))]


# =====================================================================================================
# Small ops (ltx_transformer.rs:61-339)
# =====================================================================================================
def linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """candle_nn::Linear: x @ W^T + b, W is [out, in]."""
    return F.linear(x, w, b)


def layer_norm_no_params(x: Tensor, eps: float) -> Tensor:
    """LayerNormNoParams::forward, ltx_transformer.rs:72-79 (biased variance)."""
    d = x.shape[-1]
    mean = x.sum(-1, keepdim=True) / d
    xc = x - mean
    var = (xc * xc).sum(-1, keepdim=True) / d
    return xc / torch.sqrt(var + eps)


def rms_norm(x: Tensor, weight: Optional[Tensor], eps: float) -> Tensor:
    """RmsNorm::forward, ltx_transformer.rs:99-119."""
    d = x.shape[-1]
    ms = (x * x).sum(-1, keepdim=True) * (1.0 / d)
    y = x / torch.sqrt(ms + eps)
    if weight is not None:
        y = y * weight
    return y


def gelu_approximate(x: Tensor) -> Tensor:
    """ltx_transformer.rs:214-226."""
    inner = x + 0.044715 * (x * x * x)
    scale = torch.tensor(math.sqrt(2.0 / math.pi), dtype=F32).item()
    return 0.5 * x * (1.0 + torch.tanh(inner * scale))


def dit_timestep_embedding(t: Tensor, dim: int = 256) -> Tensor:
    """get_timestep_embedding(flip_sin_to_cos=true), ltx_transformer.rs:271-309: [cos | sin]."""
    half = dim // 2
    i = torch.arange(half, dtype=F32)
    inv_freq = 1.0 / torch.pow(torch.tensor(10000.0, dtype=F32), i / half)
    freqs = t.to(F32)[:, None] * inv_freq[None, :]
    return torch.cat([freqs.cos(), freqs.sin()], dim=-1)


def apply_rotary_emb(x: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    """ltx_transformer.rs:314-339: interleaved pairs, rot(x)[2i] = -x[2i+1], rot(x)[2i+1] = x[2i]."""
    b, s, c = x.shape
    x2 = x.reshape(b, s, c // 2, 2)
    x_rot = torch.stack([-x2[..., 1], x2[..., 0]], dim=-1).reshape(b, s, c)
    return x * cos + x_rot * sin


def rope_cos_sin(coords: Tensor, dim: int, base: Tuple[int, int, int] = (20, 2048, 2048),
                 theta: float = 10000.0) -> Tuple[Tensor, Tensor]:
    """LtxVideoRotaryPosEmbed::forward with video_coords given, ltx_transformer.rs:448-523.

    coords [B,S,3] f32 (frame-seconds, pixel-y, pixel-x)  ->  cos, sin [B,S,dim] f32.
    """
    coords = coords.to(F32)
    inv = [torch.tensor(1.0 / float(torch.tensor(float(bv), dtype=F32)), dtype=F32) for bv in base]
    grid = torch.stack([coords[..., 0] * inv[0], coords[..., 1] * inv[1], coords[..., 2] * inv[2]], dim=-1)
    steps = dim // 6
    if steps <= 1:
        lin = torch.zeros(1, dtype=F32)
    else:
        lin = torch.arange(steps, dtype=F32) * torch.tensor(1.0 / (steps - 1), dtype=F32)
    theta_ln = torch.tensor(math.log(theta), dtype=F32)
    freqs = torch.exp(lin * theta_ln) * torch.tensor(math.pi / 2.0, dtype=F32)
    grid_scaled = grid.unsqueeze(-1) * 2.0 - 1.0  # [B,S,3,1]
    fr = grid_scaled * freqs.reshape(1, 1, 1, steps)  # [B,S,3,steps]
    fr = fr.transpose(-1, -2).contiguous().flatten(2)  # [B,S,steps*3], index j*3+a
    cos = fr.cos().repeat_interleave(2, dim=-1)
    sin = fr.sin().repeat_interleave(2, dim=-1)
    rem = dim % 6
    if rem:
        b, s, _ = cos.shape
        cos = torch.cat([torch.ones(b, s, rem, dtype=F32), cos], dim=-1)
        sin = torch.cat([torch.zeros(b, s, rem, dtype=F32), sin], dim=-1)
    return cos, sin


def prepare_video_coords_grid(batch: int, f: int, h: int, w: int,
                              rope_interpolation_scale: Optional[Tuple[float, float, float]],
                              patch_size: int = 1, patch_size_t: int = 1,
                              base: Tuple[int, int, int] = (20, 2048, 2048)) -> Tensor:
    """LtxVideoRotaryPosEmbed::prepare_video_coords, ltx_transformer.rs:373-433 (used when video_coords is None).

    Returns the *normalised* grid [B,S,3] (i.e. what rope_cos_sin calls `grid`), so callers use
    rope_cos_sin_from_grid.
    """
    gf = torch.arange(f, dtype=F32).reshape(f, 1, 1).expand(f, h, w)
    gh = torch.arange(h, dtype=F32).reshape(1, h, 1).expand(f, h, w)
    gw = torch.arange(w, dtype=F32).reshape(1, 1, w).expand(f, h, w)
    grid = torch.stack([gf, gh, gw], 0)
    if rope_interpolation_scale is not None:
        sf, sh, sw = rope_interpolation_scale
        fs = torch.tensor(sf * patch_size_t / base[0], dtype=F32)
        hs = torch.tensor(sh * patch_size / base[1], dtype=F32)
        ws = torch.tensor(sw * patch_size / base[2], dtype=F32)
        grid = torch.stack([grid[0] * fs, grid[1] * hs, grid[2] * ws], 0)
    grid = grid.reshape(3, f * h * w).transpose(0, 1).contiguous()
    return grid.unsqueeze(0).expand(batch, -1, -1).contiguous()


def rope_cos_sin_from_grid(grid: Tensor, dim: int, theta: float = 10000.0) -> Tuple[Tensor, Tensor]:
    """Second half of LtxVideoRotaryPosEmbed::forward (ltx_transformer.rs:473-523) on an already normalised grid."""
    ones = (1, 1, 1)
    return rope_cos_sin(grid, dim, base=ones, theta=theta)


# =====================================================================================================
# DiT (ltx_transformer.rs:527-1173)
# =====================================================================================================
def attention(w: Dict[str, Tensor], prefix: str, heads: int, x: Tensor, enc: Optional[Tensor],
              mask_bias: Optional[Tensor], rope: Optional[Tuple[Tensor, Tensor]]) -> Tensor:
    """LtxAttention::forward, ltx_transformer.rs:648-750, manual f32 branch (:717-741) for self and cross."""
    b, q_len, _ = x.shape
    e = x if enc is None else enc
    k_len = e.shape[1]
    q = linear(x, w[prefix + "to_q.weight"], w[prefix + "to_q.bias"])
    k = linear(e, w[prefix + "to_k.weight"], w[prefix + "to_k.bias"])
    v = linear(e, w[prefix + "to_v.weight"], w[prefix + "to_v.bias"])
    q = rms_norm(q, w[prefix + "norm_q.weight"], 1e-5)  # across all heads, :570-571, :671-672
    k = rms_norm(k, w[prefix + "norm_k.weight"], 1e-5)
    if rope is not None:
        q = apply_rotary_emb(q, rope[0], rope[1])
        k = apply_rotary_emb(k, rope[0], rope[1])
    d = q.shape[-1] // heads
    q = q.reshape(b, q_len, heads, d).transpose(1, 2)
    k = k.reshape(b, k_len, heads, d).transpose(1, 2)
    v = v.reshape(b, k_len, heads, d).transpose(1, 2)
    scale = float(torch.tensor(1.0, dtype=F32) / torch.sqrt(torch.tensor(float(d), dtype=F32)))
    att = (q @ k.transpose(-1, -2)) * scale
    if mask_bias is not None:  # [B,1,K] additive bias -> broadcast over heads and queries (:627-641)
        att = att + mask_bias.unsqueeze(2)
    att = torch.softmax(att, dim=-1)
    out = att @ v
    out = out.transpose(1, 2).reshape(b, q_len, heads * d)
    return linear(out, w[prefix + "to_out.0.weight"], w[prefix + "to_out.0.bias"])


def transformer_block(w: Dict[str, Tensor], prefix: str, cfg: DitConfig, x: Tensor, enc: Tensor, temb: Tensor,
                      rope: Tuple[Tensor, Tensor], mask_bias: Optional[Tensor]) -> Tensor:
    """LtxVideoTransformerBlock::forward, ltx_transformer.rs:820-937."""
    b = x.shape[0]
    d = cfg.inner_dim
    ada = w[prefix + "scale_shift_table"].reshape(1, 1, 6, d) + temb.reshape(b, 1, 6, d)
    shift_msa, scale_msa, gate_msa = ada[:, :, 0], ada[:, :, 1], ada[:, :, 2]
    shift_mlp, scale_mlp, gate_mlp = ada[:, :, 3], ada[:, :, 4], ada[:, :, 5]
    h = rms_norm(x, None, cfg.norm_eps) * (1.0 + scale_msa) + shift_msa
    a1 = attention(w, prefix + "attn1.", cfg.num_attention_heads, h, None, None, rope)
    x = x + a1 * gate_msa
    a2 = attention(w, prefix + "attn2.", cfg.num_attention_heads, x, enc, mask_bias, None)
    x = x + a2
    h = rms_norm(x, None, cfg.norm_eps) * (1.0 + scale_mlp) + shift_mlp
    ff = linear(gelu_approximate(linear(h, w[prefix + "ff.net.0.proj.weight"], w[prefix + "ff.net.0.proj.bias"])),
                w[prefix + "ff.net.2.weight"], w[prefix + "ff.net.2.bias"])
    return x + ff * gate_mlp


def _round_bf16(t: Tensor) -> Tensor:
    return t.to(torch.bfloat16).to(F32)


def dit_forward(w: Dict[str, Tensor], cfg: DitConfig, hidden: Tensor, enc: Tensor, timestep: Tensor,
                mask: Optional[Tensor], num_frames: int, height: int, width: int,
                rope_interpolation_scale: Optional[Tuple[float, float, float]] = None,
                video_coords: Optional[Tensor] = None, skip_layer_mask: Optional[Tensor] = None,
                skip_block_list: Sequence[int] = (), timestep_to_bf16: bool = False) -> Tensor:
    """LtxVideoTransformer3DModel::forward, ltx_transformer.rs:1029-1172 (f32 path).

    timestep_to_bf16 replays the reference's `timestep.to_dtype(model_dtype)` (:1051) for a bf16 model.
    """
    hidden = hidden.to(F32)
    enc = enc.to(F32)
    x = linear(hidden, w["proj_in.weight"], w["proj_in.bias"])
    t = timestep.flatten().to(F32)
    if timestep_to_bf16:
        t = _round_bf16(t)
    # AdaLayerNormSingle :262-267
    tp = dit_timestep_embedding(t, 256)
    e = linear(F.silu(linear(tp, w["time_embed.emb.timestep_embedder.linear_1.weight"],
                             w["time_embed.emb.timestep_embedder.linear_1.bias"])),
               w["time_embed.emb.timestep_embedder.linear_2.weight"],
               w["time_embed.emb.timestep_embedder.linear_2.bias"])
    temb = linear(F.silu(e), w["time_embed.linear.weight"], w["time_embed.linear.bias"])
    # caption projection :186-190
    enc = linear(gelu_approximate(linear(enc, w["caption_projection.linear_1.weight"],
                                         w["caption_projection.linear_1.bias"])),
                 w["caption_projection.linear_2.weight"], w["caption_projection.linear_2.bias"])
    mask_bias = None
    if mask is not None:  # :1059-1071
        mask_bias = ((1.0 - mask.to(F32)) * (-10000.0)).unsqueeze(1)
    d = cfg.inner_dim
    if video_coords is not None:
        cos, sin = rope_cos_sin(video_coords, d)
    else:
        grid = prepare_video_coords_grid(x.shape[0], num_frames, height, width, rope_interpolation_scale,
                                         cfg.patch_size, cfg.patch_size_t)
        cos, sin = rope_cos_sin_from_grid(grid, d)
    for i in range(cfg.num_layers):
        if i in skip_block_list:  # :1094-1096
            continue
        orig = x
        x = transformer_block(w, f"transformer_blocks.{i}.", cfg, x, enc, temb, (cos, sin), mask_bias)
        if skip_layer_mask is not None:  # :1112-1123
            m = skip_layer_mask[i].reshape(-1, 1, 1).to(F32)
            x = x * (1.0 - m) + orig * m
    table = w["scale_shift_table"]  # [2,D]: 0 shift, 1 scale  (:1126-1147)
    ss = table.reshape(1, 1, 2, d) + e.reshape(-1, 1, 1, d)
    shift, scale = ss[:, :, 0], ss[:, :, 1]
    x = layer_norm_no_params(x, 1e-6) * (1.0 + scale) + shift
    return linear(x, w["proj_out.weight"], w["proj_out.bias"])


def dit_weight_shapes(cfg: DitConfig) -> Dict[str, Tuple[int, ...]]:
    """Diffusers key names / shapes consumed by the reference's VarBuilder."""
    d, x = cfg.inner_dim, cfg.cross_attention_dim
    s: Dict[str, Tuple[int, ...]] = {
        "proj_in.weight": (d, cfg.in_channels), "proj_in.bias": (d,),
        "scale_shift_table": (2, d),
        "time_embed.emb.timestep_embedder.linear_1.weight": (d, 256),
        "time_embed.emb.timestep_embedder.linear_1.bias": (d,),
        "time_embed.emb.timestep_embedder.linear_2.weight": (d, d),
        "time_embed.emb.timestep_embedder.linear_2.bias": (d,),
        "time_embed.linear.weight": (6 * d, d), "time_embed.linear.bias": (6 * d,),
        "caption_projection.linear_1.weight": (d, cfg.caption_channels), "caption_projection.linear_1.bias": (d,),
        "caption_projection.linear_2.weight": (d, d), "caption_projection.linear_2.bias": (d,),
        "proj_out.weight": (cfg.out_channels, d), "proj_out.bias": (cfg.out_channels,),
    }
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}."
        s[p + "scale_shift_table"] = (6, d)
        for a, kv_in in (("attn1.", d), ("attn2.", x)):
            s[p + a + "to_q.weight"] = (d, d)
            s[p + a + "to_q.bias"] = (d,)
            s[p + a + "to_k.weight"] = (d, kv_in)
            s[p + a + "to_k.bias"] = (d,)
            s[p + a + "to_v.weight"] = (d, kv_in)
            s[p + a + "to_v.bias"] = (d,)
            s[p + a + "to_out.0.weight"] = (d, d)
            s[p + a + "to_out.0.bias"] = (d,)
            s[p + a + "norm_q.weight"] = (d,)
            s[p + a + "norm_k.weight"] = (d,)
        s[p + "ff.net.0.proj.weight"] = (4 * d, d)
        s[p + "ff.net.0.proj.bias"] = (4 * d,)
        s[p + "ff.net.2.weight"] = (d, 4 * d)
        s[p + "ff.net.2.bias"] = (d,)
    return s


def _init_tensor(name: str, shape: Tuple[int, ...], gen: torch.Generator) -> Tensor:
    """Synthetic init: U(+-1/sqrt(fan_in)) for Linear/Conv weights and biases (torch default),
    N(0,1)/sqrt(C) for scale_shift tables, 1 + 0.1 N(0,1) for q/k norm weights.  Values are bf16-representable
    so the bf16 device model and the f32 oracle see identical weights."""
    if name.endswith("scale_shift_table"):
        t = torch.randn(shape, generator=gen, dtype=F32) / math.sqrt(shape[-1])
    elif "norm_q" in name or "norm_k" in name:
        t = 1.0 + 0.1 * torch.randn(shape, generator=gen, dtype=F32)
    elif name.endswith("timestep_scale_multiplier"):
        t = torch.tensor(1000.0, dtype=F32)
    elif name.endswith(".bias"):
        t = (torch.rand(shape, generator=gen, dtype=F32) * 2 - 1) * 0.05
    else:
        fan_in = 1
        for v in shape[1:]:
            fan_in *= v
        t = (torch.rand(shape, generator=gen, dtype=F32) * 2 - 1) / math.sqrt(fan_in)
    return _round_bf16(t)


def init_dit_weights(cfg: DitConfig, seed: int = 42) -> Dict[str, Tensor]:
    gen = torch.Generator().manual_seed(seed)
    return {k: _init_tensor(k, s, gen) for k, s in dit_weight_shapes(cfg).items()}


# =====================================================================================================
# VAE decoder (vae.rs:148-265, 298-465, 585-822, 951-1313, 1472-1727)
# =====================================================================================================
def vae_timestep_embedding(t: Tensor, dim: int = 256) -> Tensor:
    """get_timestep_embedding, vae.rs:172-198: exp(-ln(1e4)/half * i), [cos | sin]."""
    half = dim // 2
    coef = torch.tensor(-math.log(10000.0) / half, dtype=F32)
    emb = torch.exp(torch.arange(half, dtype=F32) * coef)
    e = t.to(F32)[:, None] * emb[None, :]
    return torch.cat([e.cos(), e.sin()], dim=1)


def vae_time_embedder(w: Dict[str, Tensor], prefix: str, t: Tensor) -> Tensor:
    """CombinedTimestepEmbedder / TimestepEmbedder, vae.rs:202-265."""
    p = prefix + "timestep_embedder."
    h = linear(vae_timestep_embedding(t, 256), w[p + "linear_1.weight"], w[p + "linear_1.bias"])
    return linear(F.silu(h), w[p + "linear_2.weight"], w[p + "linear_2.bias"])


def pixel_norm(x: Tensor, eps: float = 1e-8) -> Tensor:
    """rmsnorm_channels_first with weight = ones, vae.rs:148-153, :618-628."""
    return x / torch.sqrt((x * x).mean(dim=1, keepdim=True) + eps)


def causal_conv3d(x: Tensor, weight: Tensor, bias: Tensor, is_causal: bool = False) -> Tensor:
    """LtxVideoCausalConv3d::forward, vae.rs:374-464: replicate pad in T (causal: kt-1 left; else (kt-1)/2 both
    sides), zero pad kh/2 in H and W, stride/dilation 1.  Sum over kt of Conv2d == Conv3d."""
    kt, kh, kw = weight.shape[2:]
    if kt > 1:
        if is_causal:
            x = torch.cat([x[:, :, :1].repeat(1, 1, kt - 1, 1, 1), x], dim=2)
        else:
            p = (kt - 1) // 2
            x = torch.cat([x[:, :, :1].repeat(1, 1, p, 1, 1), x, x[:, :, -1:].repeat(1, 1, p, 1, 1)], dim=2)
    return F.conv3d(x, weight, bias, stride=1, padding=(0, kh // 2, kh // 2))


def resnet_block(w: Dict[str, Tensor], prefix: str, x: Tensor, temb: Optional[Tensor], is_causal: bool = False) -> Tensor:
    """LtxVideoResnetBlock3d::forward (in == out), vae.rs:711-821.  Decoder: non-causal convs, optional timestep
    conditioning; encoder: causal convs, no conditioning (vae.rs:863-876)."""
    b, c = x.shape[:2]
    tbl = w.get(prefix + "scale_shift_table")
    ss = None
    if tbl is not None and temb is not None:
        ss = temb.reshape(b, 4, c, 1, 1, 1) + tbl.reshape(1, 4, c, 1, 1, 1)
    h = pixel_norm(x)
    if ss is not None:
        h = h * (1.0 + ss[:, 1]) + ss[:, 0]
    h = F.silu(h)
    h = causal_conv3d(h, w[prefix + "conv1.conv.weight"], w[prefix + "conv1.conv.bias"], is_causal)
    h = pixel_norm(h)
    if ss is not None:
        h = h * (1.0 + ss[:, 3]) + ss[:, 2]
    h = F.silu(h)
    h = causal_conv3d(h, w[prefix + "conv2.conv.weight"], w[prefix + "conv2.conv.bias"], is_causal)
    return h + x


def depth_to_space(x: Tensor, st: int, sh: int, sw: int) -> Tensor:
    """vae.rs:1142-1158: out[b,c,t*st+i,h*sh+j,w*sw+k] = x[b, c*st*sh*sw + (i*sh+j)*sw + k, t,h,w]."""
    b, c, t, h, w = x.shape
    co = c // (st * sh * sw)
    x = x.reshape(b, co, st, sh, sw, t, h, w).permute(0, 1, 5, 2, 6, 3, 7, 4).contiguous()
    return x.reshape(b, co, t * st, h * sh, w * sw)


def upsampler(w: Dict[str, Tensor], prefix: str, x: Tensor, residual: bool = True) -> Tensor:
    """LtxVideoUpsampler3d::forward, stride (2,2,2), vae.rs:1090-1169."""
    wt = w[prefix + "conv.conv.weight"]
    c_in = x.shape[1]
    repeats = wt.shape[0] // c_in
    res = None
    if residual:
        res = depth_to_space(x, 2, 2, 2)
        if repeats > 1:
            res = res.repeat(1, repeats, 1, 1, 1)
        res = res[:, :, 1:]
    h = causal_conv3d(x, wt, w[prefix + "conv.conv.bias"])
    h = depth_to_space(h, 2, 2, 2)[:, :, 1:]
    return h + res if res is not None else h


def unpatchify(x: Tensor, p: int = 4, pt: int = 1) -> Tensor:
    """LtxVideoDecoder3d::unpatchify, vae.rs:1626-1654."""
    b, c, f, h, w = x.shape
    oc = c // (pt * p * p)
    x = x.reshape(b, oc, pt, p, p, f, h, w).permute(0, 1, 5, 2, 6, 4, 7, 3).contiguous()
    return x.reshape(b, oc, f * pt, h * p, w * p)


def vae_decode(w: Dict[str, Tensor], cfg: VaeConfig, z: Tensor, timestep: Optional[Tensor]) -> Tensor:
    """AutoencoderKLLtxVideo::decode -> decode_z (untiled) -> LtxVideoDecoder3d::forward,
    vae.rs:2101, :2037, :1656-1726.  Keys carry the reference's `decoder.` prefix."""
    z = z.to(F32)
    P = "decoder."
    h = causal_conv3d(z, w[P + "conv_in.conv.weight"], w[P + "conv_in.conv.bias"])
    ts = None
    if timestep is not None and cfg.timestep_conditioning:
        ts = timestep.flatten().to(F32)
        if (P + "timestep_scale_multiplier") in w:
            ts = ts * w[P + "timestep_scale_multiplier"]
    b = h.shape[0]
    # mid block :998-1033
    temb = vae_time_embedder(w, P + "mid_block.time_embedder.", ts) if ts is not None else None
    for i in range(cfg.decoder_layers_per_block[0]):
        h = resnet_block(w, P + f"mid_block.resnets.{i}.", h, temb)
    # up blocks :1274-1311
    for bi in range(len(cfg.decoder_block_out_channels)):
        bp = P + f"up_blocks.{bi}."
        temb = vae_time_embedder(w, bp + "time_embedder.", ts) if ts is not None else None
        h = upsampler(w, bp + "upsamplers.0.", h, residual=True)
        for i in range(cfg.decoder_layers_per_block[bi + 1]):
            h = resnet_block(w, bp + f"resnets.{i}.", h, temb)
    h = pixel_norm(h)
    if ts is not None:  # :1693-1721
        c = h.shape[1]
        tp = vae_time_embedder(w, P + "time_embedder.", ts).reshape(b, 2, c) + w[P + "scale_shift_table"].reshape(1, 2, c)
        shift = tp[:, 0].reshape(b, c, 1, 1, 1)
        scale = tp[:, 1].reshape(b, c, 1, 1, 1)
        h = h * (1.0 + scale) + shift
    h = F.silu(h)
    h = causal_conv3d(h, w[P + "conv_out.conv.weight"], w[P + "conv_out.conv.bias"])
    return unpatchify(h, cfg.patch_size, cfg.patch_size_t)


# ---------------------------------------------------------------------------------------------------------------
# encoder: LtxVideoEncoder3d / AutoencoderKLLtxVideo::encode, vae.rs:496-582, :841-948, :1315-1469,
# :2017-2099.  0.9.5 layout only: pixel-unshuffle downsamplers (the stride-2 "conv" type of 0.9.0 is not restated).
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class VaeEncoderConfig:  # AutoencoderKLLtxVideoConfig defaults, vae.rs:68-103
    in_channels: int = 3
    latent_channels: int = 128
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 1024, 2048)
    layers_per_block: Tuple[int, ...] = (4, 6, 6, 2, 2)
    downsample_types: Tuple[str, ...] = ("spatial", "temporal", "spatiotemporal", "spatiotemporal")
    patch_size: int = 4
    patch_size_t: int = 1


DOWNSAMPLE_STRIDE = {"spatial": (1, 2, 2), "temporal": (2, 1, 1), "spatiotemporal": (2, 2, 2)}  # vae.rs:484-493


def patchify(x: Tensor, p: int = 4, pt: int = 1) -> Tensor:
    """LtxVideoEncoder3d::patchify, vae.rs:1427-1445: channel = ((c*pt + it)*p + iw)*p + ih."""
    b, c, f, h, w = x.shape
    if f % pt or h % p or w % p:
        raise ValueError("input not divisible by patch sizes")
    x = x.reshape(b, c, f // pt, pt, h // p, p, w // p, p).permute(0, 1, 3, 7, 5, 2, 4, 6).contiguous()
    return x.reshape(b, c * pt * p * p, f // pt, h // p, w // p)


def space_to_depth(x: Tensor, st: int, sh: int, sw: int) -> Tensor:
    """vae.rs:553-556 / :574-577: out[b, ((c*st+i)*sh+j)*sw+k, t, h, w] = x[b, c, t*st+i, h*sh+j, w*sw+k]."""
    b, c, t, h, w = x.shape
    x = x.reshape(b, c, t // st, st, h // sh, sh, w // sw, sw).permute(0, 1, 3, 5, 7, 2, 4, 6).contiguous()
    return x.reshape(b, c * st * sh * sw, t // st, h // sh, w // sw)


def downsampler(w: Dict[str, Tensor], prefix: str, x: Tensor, stride: Tuple[int, int, int], out_channels: int) -> Tensor:
    """LtxVideoDownsampler3d::forward, vae.rs:534-582: duplicate the first st-1 frames, causal conv to
    out_channels/(st*sh*sw), pixel-unshuffle; residual = pixel-unshuffled input averaged in channel groups."""
    st, sh, sw = stride
    c = x.shape[1]
    group = (c * st * sh * sw) // out_channels
    if st > 1:
        x = torch.cat([x[:, :, : st - 1], x], dim=2)
    res = space_to_depth(x, st, sh, sw)
    b, cr, t, h, wd = res.shape
    res = res.reshape(b, cr // group, group, t, h, wd).mean(dim=2)
    h_ = causal_conv3d(x, w[prefix + "conv.conv.weight"], w[prefix + "conv.conv.bias"], is_causal=True)
    return space_to_depth(h_, st, sh, sw) + res


def vae_encode(w: Dict[str, Tensor], cfg: VaeEncoderConfig, x: Tensor) -> Tensor:
    """AutoencoderKLLtxVideo::encode -> encode_z (untiled, no quant_conv) -> LtxVideoEncoder3d::forward,
    vae.rs:2070, :2017-2034, :1447-1468.  x [B,3,F,H,W] in [-1,1] -> moments [B, 2*latent, F', H', W']:
    channels [0,latent) = mean, [latent, 2*latent) = logvar (one conv channel replicated, :1462-1467)."""
    x = x.to(F32)
    P = "encoder."
    h = patchify(x, cfg.patch_size, cfg.patch_size_t)
    h = causal_conv3d(h, w[P + "conv_in.conv.weight"], w[P + "conv_in.conv.bias"], is_causal=True)
    boc = cfg.block_out_channels
    for bi in range(len(boc) - 1):
        bp = P + f"down_blocks.{bi}."
        for i in range(cfg.layers_per_block[bi]):
            h = resnet_block(w, bp + f"resnets.{i}.", h, None, is_causal=True)
        h = downsampler(w, bp + "downsamplers.0.", h, DOWNSAMPLE_STRIDE[cfg.downsample_types[bi]], boc[bi + 1])
    for i in range(cfg.layers_per_block[-1] - 1):  # vae.rs:1383-1386
        h = resnet_block(w, P + f"mid_block.resnets.{i}.", h, None, is_causal=True)
    h = F.silu(pixel_norm(h))
    h = causal_conv3d(h, w[P + "conv_out.conv.weight"], w[P + "conv_out.conv.bias"], is_causal=True)
    ch = h.shape[1]
    return torch.cat([h, h[:, -1:].repeat(1, ch - 2, 1, 1, 1)], dim=1)


# ---------------------------------------------------------------------------------------------------------------
# tiled / temporal-tiled ENCODE (encode_z dispatch, vae.rs:2017-2034; library defaults use_tiling = true,
# use_framewise_encoding = false, :1856-1858).  Tiles are cut in SAMPLE space, blended in LATENT space.
# ---------------------------------------------------------------------------------------------------------------
def vae_tiled_encode(w, ecfg: "VaeEncoderConfig", x: Tensor, tp: VaeTiling, sr: int = 32) -> Tensor:
    """AutoencoderKLLtxVideo::tiled_encode, vae.rs:2158-2223."""
    H, W = x.shape[3], x.shape[4]
    lat_h, lat_w = H // sr, W // sr
    min_h, min_w = tp.tile_sample_min_height // sr, tp.tile_sample_min_width // sr
    str_h, str_w = tp.tile_sample_stride_height // sr, tp.tile_sample_stride_width // sr
    blend_h, blend_w = max(min_h - str_h, 0), max(min_w - str_w, 0)
    rows = []
    for i in range(0, H, tp.tile_sample_stride_height):
        row = []
        for j in range(0, W, tp.tile_sample_stride_width):
            tile = x[:, :, :, i:min(i + tp.tile_sample_min_height, H), j:min(j + tp.tile_sample_min_width, W)]
            row.append(vae_encode(w, ecfg, tile))
        rows.append(row)
    result_rows, prev = [], []
    for ri, row in enumerate(rows):
        res, cur = [], []
        for cj, tile in enumerate(row):
            if ri > 0:
                tile = _blend(prev[cj], tile, blend_h, 3)
            if cj > 0:
                tile = _blend(cur[cj - 1], tile, blend_w, 4)
            cur.append(tile)
            res.append(tile[:, :, :, :min(str_h, tile.shape[3]), :min(str_w, tile.shape[4])])
        result_rows.append(torch.cat(res, dim=4))
        prev = cur
    return torch.cat(result_rows, dim=3)[:, :, :, :lat_h, :lat_w]


def vae_temporal_tiled_encode(w, ecfg: "VaeEncoderConfig", x: Tensor, tp: VaeTiling, sr: int = 32, tr: int = 8) -> Tensor:
    """AutoencoderKLLtxVideo::temporal_tiled_encode, vae.rs:2294-2356 (incl. its quirks: the first tile drops its first
    latent frame, tiles are blended with the UNBLENDED previous tile, the first `stride` latent frames are kept)."""
    nf = x.shape[2]
    lat_f = (nf - 1) // tr + 1
    min_t, str_t = tp.tile_sample_min_num_frames // tr, tp.tile_sample_stride_num_frames // tr
    blend_t = max(min_t - str_t, 0)
    row = []
    for i in range(0, nf, tp.tile_sample_stride_num_frames):
        tile = x[:, :, i:min(i + tp.tile_sample_min_num_frames + 1, nf)]
        if tp.use_tiling and (tile.shape[3] > tp.tile_sample_min_height or tile.shape[4] > tp.tile_sample_min_width):
            t = vae_tiled_encode(w, ecfg, tile, tp, sr)
        else:
            t = vae_encode(w, ecfg, tile)
        if i == 0:
            t = t[:, :, 1:]
        row.append(t)
    res = []
    for idx, t in enumerate(row):
        if idx > 0:
            bl = _blend(row[idx - 1], t, blend_t, 2)
            res.append(bl[:, :, :min(str_t, bl.shape[2])])
        else:
            res.append(t[:, :, :min(str_t + 1, t.shape[2])])
    return torch.cat(res, dim=2)[:, :, :lat_f]


def vae_encode_z(w, ecfg: "VaeEncoderConfig", x: Tensor, tp: Optional[VaeTiling], use_framewise_encoding: bool = False,
                 sr: int = 32, tr: int = 8) -> Tensor:
    """encode_z dispatch, vae.rs:2017-2034 (tp = None: plain encoder)."""
    if tp is not None:
        if use_framewise_encoding and x.shape[2] > tp.tile_sample_min_num_frames:
            return vae_temporal_tiled_encode(w, ecfg, x, tp, sr, tr)
        if tp.use_tiling and (x.shape[3] > tp.tile_sample_min_height or x.shape[4] > tp.tile_sample_min_width):
            return vae_tiled_encode(w, ecfg, x, tp, sr)
    return vae_encode(w, ecfg, x)


def vae_encoder_weight_shapes(cfg: VaeEncoderConfig) -> Dict[str, Tuple[int, ...]]:
    """Appendix A (encoder half): keys as consumed by VarBuilder at vae.rs:1342-1412, :863-905, :513-524."""
    s: Dict[str, Tuple[int, ...]] = {}
    P = "encoder."

    def conv(prefix: str, cin: int, cout: int) -> None:
        s[prefix + ".conv.weight"] = (cout, cin, 3, 3, 3)
        s[prefix + ".conv.bias"] = (cout,)

    boc = cfg.block_out_channels
    conv(P + "conv_in", cfg.in_channels * cfg.patch_size_t * cfg.patch_size ** 2, boc[0])
    for bi in range(len(boc) - 1):
        c = boc[bi]
        for i in range(cfg.layers_per_block[bi]):
            conv(P + f"down_blocks.{bi}.resnets.{i}.conv1", c, c)
            conv(P + f"down_blocks.{bi}.resnets.{i}.conv2", c, c)
        st, sh, sw = DOWNSAMPLE_STRIDE[cfg.downsample_types[bi]]
        conv(P + f"down_blocks.{bi}.downsamplers.0.conv", c, boc[bi + 1] // (st * sh * sw))
    c = boc[-1]
    for i in range(cfg.layers_per_block[-1] - 1):
        conv(P + f"mid_block.resnets.{i}.conv1", c, c)
        conv(P + f"mid_block.resnets.{i}.conv2", c, c)
    conv(P + "conv_out", c, cfg.latent_channels + 1)
    return s


def init_vae_encoder_weights(cfg: VaeEncoderConfig, seed: int = 43) -> Dict[str, Tensor]:
    gen = torch.Generator().manual_seed(seed)
    return {k: _init_tensor(k, shp, gen) for k, shp in vae_encoder_weight_shapes(cfg).items()}


def normalize_latents(latents: Tensor, mean: Tensor, std: Tensor, scaling_factor: float) -> Tensor:
    """t2v_pipeline.rs:552-571: (x - mean) * scaling_factor / std per channel."""
    c = latents.shape[1]
    return (latents - mean.reshape(1, c, 1, 1, 1)) * scaling_factor / std.reshape(1, c, 1, 1, 1)


# ---------------------------------------------------------------------------------------------------------------
# tiled / temporal-tiled decode (the LIBRARY default of the reference: use_tiling = use_framewise_decoding = true,
# vae.rs:1848-1861; the example binary only enables it with --vae-tiling, main.rs:516-518)
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class VaeTiling:
    """vae.rs:1848-1861 (defaults) / enable_tiling :1870-1898; all sizes in SAMPLE space (pixels / frames)."""
    use_tiling: bool = True
    use_framewise_decoding: bool = True
    tile_sample_min_height: int = 512
    tile_sample_min_width: int = 512
    tile_sample_min_num_frames: int = 16
    tile_sample_stride_height: int = 384
    tile_sample_stride_width: int = 384
    tile_sample_stride_num_frames: int = 8


def _blend(a: Tensor, b: Tensor, blend_extent: int, dim: int) -> Tensor:
    """blend_h / blend_v / blend_t, vae.rs:1927-2006: the first `blend` slices of b along `dim` become
    a[-blend + x] * (1 - x/blend) + b[x] * (x/blend), w = x * f32(1/blend)."""
    blend = min(blend_extent, a.shape[dim], b.shape[dim])
    if blend == 0:
        return b.clone()
    w = torch.arange(blend, dtype=F32) * torch.tensor(1.0 / blend, dtype=F32)
    shape = [1] * b.dim()
    shape[dim] = blend
    w = w.reshape(shape)
    one_minus = 1.0 - w
    b_head = b.narrow(dim, 0, blend)
    b_tail = b.narrow(dim, blend, b.shape[dim] - blend)
    a_tail = a.narrow(dim, a.shape[dim] - blend, blend)
    mixed = a_tail * one_minus + b_head * w
    return torch.cat([mixed, b_tail], dim=dim)


def vae_tiled_decode(w, cfg: VaeConfig, z: Tensor, timestep, tp: VaeTiling, sr: int = 32) -> Tensor:
    """AutoencoderKLLtxVideo::tiled_decode, vae.rs:2225-2290."""
    height, width = z.shape[3], z.shape[4]
    sample_h, sample_w = height * sr, width * sr
    tl_min_h, tl_min_w = tp.tile_sample_min_height // sr, tp.tile_sample_min_width // sr
    tl_str_h, tl_str_w = tp.tile_sample_stride_height // sr, tp.tile_sample_stride_width // sr
    blend_h = max(tp.tile_sample_min_height - tp.tile_sample_stride_height, 0)
    blend_w = max(tp.tile_sample_min_width - tp.tile_sample_stride_width, 0)
    rows = []
    for i in range(0, height, tl_str_h):
        row = []
        for j in range(0, width, tl_str_w):
            tile = z[:, :, :, i:min(i + tl_min_h, height), j:min(j + tl_min_w, width)]
            row.append(vae_decode(w, cfg, tile, timestep))
        rows.append(row)
    prev_row, result_rows = [], []
    for ri, row in enumerate(rows):
        result_row, cur_row = [], []
        for cj, tile in enumerate(row):
            if ri > 0:
                tile = _blend(prev_row[cj], tile, blend_h, 3)   # blend_v acts on H (dim 3)
            if cj > 0:
                tile = _blend(cur_row[cj - 1], tile, blend_w, 4)  # blend_h acts on W (dim 4)
            cur_row.append(tile)
            hs = min(tp.tile_sample_stride_height, tile.shape[3])
            ws = min(tp.tile_sample_stride_width, tile.shape[4])
            result_row.append(tile[:, :, :, :hs, :ws])
        result_rows.append(torch.cat(result_row, dim=4))
        prev_row = cur_row
    return torch.cat(result_rows, dim=3)[:, :, :, :sample_h, :sample_w]


def vae_temporal_tiled_decode(w, cfg: VaeConfig, z: Tensor, timestep, tp: VaeTiling, sr: int = 32, tr: int = 8) -> Tensor:
    """AutoencoderKLLtxVideo::temporal_tiled_decode, vae.rs:2358-2434."""
    num_frames = z.shape[2]
    num_sample_frames = (num_frames - 1) * tr + 1
    tl_min_h, tl_min_w = tp.tile_sample_min_height // sr, tp.tile_sample_min_width // sr
    tl_min_t = tp.tile_sample_min_num_frames // tr
    tl_str_t = tp.tile_sample_stride_num_frames // tr
    blend_t = max(tp.tile_sample_min_num_frames - tp.tile_sample_stride_num_frames, 0)
    row = []
    for loop_idx, i in enumerate(range(0, num_frames, tl_str_t)):
        tile = z[:, :, i:min(i + tl_min_t + 1, num_frames)]
        if tp.use_tiling and (tile.shape[3] > tl_min_h or tile.shape[4] > tl_min_w):
            dec = vae_tiled_decode(w, cfg, tile, timestep, tp, sr)
        else:
            dec = vae_decode(w, cfg, tile, timestep)
        if loop_idx > 0 and dec.shape[2] > 1:
            dec = dec[:, :, :-1]
        row.append(dec)
    out = []
    for idx, tile in enumerate(row):
        if idx > 0:
            blended = _blend(row[idx - 1], tile, blend_t, 2)
            out.append(blended[:, :, :min(tp.tile_sample_stride_num_frames, blended.shape[2])])
        else:
            out.append(tile[:, :, :min(tp.tile_sample_stride_num_frames + 1, tile.shape[2])])
    return torch.cat(out, dim=2)[:, :, :num_sample_frames]


def vae_decode_z(w, cfg: VaeConfig, z: Tensor, timestep, tp: Optional[VaeTiling], sr: int = 32, tr: int = 8) -> Tensor:
    """decode_z dispatch, vae.rs:2037-2066: framewise first, then spatial tiling, else the plain decoder."""
    if tp is None:
        return vae_decode(w, cfg, z, timestep)
    t, h, wd = z.shape[2], z.shape[3], z.shape[4]
    if tp.use_framewise_decoding and t > tp.tile_sample_min_num_frames // tr:
        return vae_temporal_tiled_decode(w, cfg, z, timestep, tp, sr, tr)
    if tp.use_tiling and (wd > tp.tile_sample_min_width // sr or h > tp.tile_sample_min_height // sr):
        return vae_tiled_decode(w, cfg, z, timestep, tp, sr)
    return vae_decode(w, cfg, z, timestep)


def vae_weight_shapes(cfg: VaeConfig) -> Dict[str, Tuple[int, ...]]:
    """Decoder keys / shapes (vae.rs:323-335, :983-987, :1070-1079, :1257-1263, :1583-1604)."""
    ch = cfg.stage_channels()  # [1024, 512, 256, 128]
    P = "decoder."
    s: Dict[str, Tuple[int, ...]] = {
        P + "conv_in.conv.weight": (ch[0], cfg.latent_channels, 3, 3, 3), P + "conv_in.conv.bias": (ch[0],),
    }

    def time_embedder(prefix: str, dim: int) -> None:
        s[prefix + "timestep_embedder.linear_1.weight"] = (dim, 256)
        s[prefix + "timestep_embedder.linear_1.bias"] = (dim,)
        s[prefix + "timestep_embedder.linear_2.weight"] = (dim, dim)
        s[prefix + "timestep_embedder.linear_2.bias"] = (dim,)

    def resnets(prefix: str, c: int, n: int) -> None:
        for i in range(n):
            for cv in ("conv1", "conv2"):
                s[f"{prefix}resnets.{i}.{cv}.conv.weight"] = (c, c, 3, 3, 3)
                s[f"{prefix}resnets.{i}.{cv}.conv.bias"] = (c,)
            s[f"{prefix}resnets.{i}.scale_shift_table"] = (4, c)

    time_embedder(P + "mid_block.time_embedder.", 4 * ch[0])
    resnets(P + "mid_block.", ch[0], cfg.decoder_layers_per_block[0])
    for bi in range(len(cfg.decoder_block_out_channels)):
        cin, cout = ch[bi], ch[bi + 1]
        bp = P + f"up_blocks.{bi}."
        s[bp + "upsamplers.0.conv.conv.weight"] = (cout * 8, cin, 3, 3, 3)
        s[bp + "upsamplers.0.conv.conv.bias"] = (cout * 8,)
        time_embedder(bp + "time_embedder.", 4 * cout)
        resnets(bp, cout, cfg.decoder_layers_per_block[bi + 1])
    c = ch[-1]
    s[P + "conv_out.conv.weight"] = (cfg.out_channels * cfg.patch_size ** 2, c, 3, 3, 3)
    s[P + "conv_out.conv.bias"] = (cfg.out_channels * cfg.patch_size ** 2,)
    time_embedder(P + "time_embedder.", 2 * c)
    s[P + "scale_shift_table"] = (2, c)
    s[P + "timestep_scale_multiplier"] = ()
    return s


def init_vae_weights(cfg: VaeConfig, seed: int = 42) -> Dict[str, Tensor]:
    gen = torch.Generator().manual_seed(seed)
    return {k: _init_tensor(k, s, gen) for k, s in vae_weight_shapes(cfg).items()}


# =====================================================================================================
# Pipeline glue (t2v_pipeline.rs) and scheduler (scheduler.rs)
# =====================================================================================================
def pack_latents(latents: Tensor, p: int = 1, pt: int = 1) -> Tensor:
    """LtxPipeline::pack_latents, t2v_pipeline.rs:474-504."""
    b, c, f, h, w = latents.shape
    x = latents.reshape(b, c, f // pt, pt, h // p, p, w // p, p).permute(0, 2, 4, 6, 1, 3, 5, 7)
    return x.flatten(4).reshape(b, (f // pt) * (h // p) * (w // p), -1)


def unpack_latents(latents: Tensor, f: int, h: int, w: int, p: int = 1, pt: int = 1) -> Tensor:
    """LtxPipeline::unpack_latents, t2v_pipeline.rs:506-550."""
    b, _, d = latents.shape
    c = d // (pt * p * p)
    x = latents.reshape(b, f, h, w, c, pt, p, p).permute(0, 4, 1, 5, 2, 6, 3, 7).contiguous()
    return x.reshape(b, c, f * pt, h * p, w * p)


def video_coords(batch: int, f: int, h: int, w: int, frame_rate: int, ts_ratio: int = 8, sp_ratio: int = 32) -> Tensor:
    """t2v_pipeline.rs:798-847: (clamp(8f-7, 0, 1000) * f32(1/fps), 32h, 32w), token order f,h,w."""
    gf = torch.arange(f, dtype=F32).reshape(f, 1, 1).expand(f, h, w)
    gh = torch.arange(h, dtype=F32).reshape(1, h, 1).expand(f, h, w)
    gw = torch.arange(w, dtype=F32).reshape(1, 1, w).expand(f, h, w)
    c = torch.stack([gf, gh, gw], 0).flatten(1).transpose(0, 1)  # [S,3]
    tsr = torch.tensor(float(ts_ratio), dtype=F32)
    vf = (c[:, 0] * tsr + (1.0 - tsr)).clamp(0.0, 1000.0) * torch.tensor(1.0 / frame_rate, dtype=F32)
    vh = c[:, 1] * torch.tensor(float(sp_ratio), dtype=F32)
    vw = c[:, 2] * torch.tensor(float(sp_ratio), dtype=F32)
    out = torch.stack([vf, vh, vw], dim=-1)
    return out.unsqueeze(0).expand(batch, -1, -1).contiguous()


def std_over_dims_except0(x: Tensor) -> Tensor:
    """t2v_pipeline.rs:209-224: unbiased std over all non-batch elements, keepdim."""
    b = x.shape[0]
    return x.reshape(b, -1).var(dim=1, unbiased=True).sqrt().reshape([b] + [1] * (x.dim() - 1))


def rescale_noise_cfg(noise_cfg: Tensor, noise_text: Tensor, guidance_rescale: float) -> Tensor:
    """t2v_pipeline.rs:227-243."""
    ratio = std_over_dims_except0(noise_text) / std_over_dims_except0(noise_cfg)
    return (noise_cfg * ratio) * guidance_rescale + noise_cfg * (1.0 - guidance_rescale)


def guidance_combine(cond: Tensor, uncond: Optional[Tensor], perturbed: Optional[Tensor], guidance_scale: float,
                     guidance_rescale: float, stg_scale: float) -> Tensor:
    """t2v_pipeline.rs:942-964."""
    cond = cond.to(F32)
    comb = cond.clone()
    if uncond is not None:
        uncond = uncond.to(F32)
        comb = uncond + (cond - uncond) * guidance_scale
        if guidance_rescale > 0.0:
            comb = rescale_noise_cfg(comb, cond, guidance_rescale)
    if perturbed is not None:
        comb = comb + (cond - perturbed.to(F32)) * stg_scale
    return comb


def euler_step(sample: Tensor, model_output: Tensor, sigma: float, sigma_next: float) -> Tensor:
    """FlowMatchEulerDiscreteScheduler::step non-stochastic branch, scheduler.rs:544-554, :576-582."""
    dt = torch.tensor(sigma_next, dtype=F32) - torch.tensor(sigma, dtype=F32)
    return sample.to(F32) + model_output.to(F32) * dt


def stochastic_step(sample: Tensor, model_output: Tensor, noise: Tensor, sigma: float, sigma_next: float) -> Tensor:
    """FlowMatchEulerDiscreteScheduler::step, stochastic branch (scheduler.rs:557-575) with the noise passed in
    (the reference draws it with Tensor::randn on the device): x0 = x - sigma v; (1 - sigma_next) x0 + sigma_next n."""
    sample = sample.to(F32)
    cs = torch.tensor(sigma, dtype=F32)
    ns = torch.tensor(sigma_next, dtype=F32)
    x0 = sample - cs * model_output.to(F32)
    one_minus = ns * -1.0 + 1.0
    return one_minus * x0 + ns * noise.to(F32)


def decode_noise_blend(latents: Tensor, noise: Tensor, scale: float) -> Tensor:
    """t2v_pipeline.rs:1049-1062: latents * (1 - scale) + noise * scale (f32 here; model dtype in a bf16 run)."""
    sc = torch.tensor(scale, dtype=F32)
    one_minus = sc * -1.0 + 1.0
    return latents.to(F32) * one_minus + noise.to(F32) * sc


def denormalize_latents(latents: Tensor, mean: Tensor, std: Tensor, scaling_factor: float) -> Tensor:
    """t2v_pipeline.rs:573-594."""
    c = latents.shape[1]
    inv_sf = torch.tensor(1.0, dtype=F32) / torch.tensor(scaling_factor, dtype=F32)  # (1.0 / scaling_factor: f32) as f64
    return latents * std.reshape(1, c, 1, 1, 1) * inv_sf + mean.reshape(1, c, 1, 1, 1)


def postprocess_video(video: Tensor) -> Tensor:
    """LtxVideoProcessor::postprocess_video, t2v_pipeline.rs:147-155."""
    return (video * 0.5 + 0.5).clamp(0.0, 1.0) * 255.0


def frames_to_u8(video: Tensor) -> Tensor:
    """examples/ltx-video/main.rs:653-667: per frame permute((1,2,0)).clamp(0,255).to_dtype(U8) (truncating cast).
    [B,3,F,H,W] f32 in 0..255 -> [B,F,H,W,3] u8."""
    return video.permute(0, 2, 3, 4, 1).clamp(0.0, 255.0).to(torch.uint8).contiguous()


def calculate_shift(seq_len: int, base_seq_len: int = 256, max_seq_len: int = 4096, base_shift: float = 0.5,
                    max_shift: float = 1.15) -> float:
    """t2v_pipeline.rs:159-169 (all f32)."""
    f = lambda v: torch.tensor(float(v), dtype=F32)  # noqa: E731
    m = (f(max_shift) - f(base_shift)) / f(max_seq_len - base_seq_len)
    b = f(base_shift) - m * f(base_seq_len)
    return float(f(seq_len) * m + b)


def _libm_expf(x: float) -> float:
    """`f32::exp` of the reference (scheduler.rs:172-176) lowers to the platform libm's `expf`; so does the C++ host
    code of the product.  numpy's SIMD exp and a double-precision exp rounded to f32 both differ from glibc's expf by
    one ulp for some arguments (e.g. mu = 0.48341146), which flips `trunc(sigma * 1000)` at integer boundaries -- the
    oracle therefore calls the same libm function, and the integer timesteps are compared exactly."""
    import ctypes
    import ctypes.util
    global _LIBM
    if _LIBM is None:
        _LIBM = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        _LIBM.expf.restype = ctypes.c_float
        _LIBM.expf.argtypes = [ctypes.c_float]
    return float(_LIBM.expf(x))


_LIBM = None


def scheduler_set_timesteps(num_steps: int, mu: float, sigmas: Optional[Sequence[float]] = None,
                            shift_terminal: Optional[float] = 0.1) -> Tuple[List[float], List[int]]:
    """FlowMatchEulerDiscreteScheduler::set_timesteps as driven by LtxPipeline::call
    (t2v_pipeline.rs:752-792, scheduler.rs:274-412, :646-660), f32 arithmetic via numpy.

    Returns (sigmas incl. terminal 0, integer-truncated timesteps handed to the DiT)."""
    import numpy as np
    f = np.float32
    if sigmas is None:
        n = num_steps
        if n == 1:
            s = np.array([1.0], dtype=f)
        else:
            s = (f(1.0) + (f(1.0 / f(n)) - f(1.0)) * np.arange(n, dtype=f) / f(n - 1)).astype(f)
    else:
        s = np.asarray(sigmas, dtype=f)
    emu = f(_libm_expf(float(f(mu))))
    with np.errstate(divide="ignore"):
        base = (f(1.0) / s - f(1.0)).astype(f)  # sigma exponent 1.0: powf(x, 1) == x
    s = (emu / (emu + base)).astype(f)
    if shift_terminal is not None and len(s) > 0:
        one_minus_last = f(1.0) - s[-1]
        scale = one_minus_last / (f(1.0) - f(shift_terminal))
        with np.errstate(divide="ignore", invalid="ignore"):
            s = (f(1.0) - (f(1.0) - s) / scale).astype(f)
    ts = (s * f(1000.0)).astype(f)
    # Rust `f32 as i64` saturates and maps NaN to 0 (n = 1 with a terminal stretch divides 0/0 in the reference too)
    return [float(v) for v in s] + [0.0], [0 if np.isnan(v) else int(v) for v in ts]
