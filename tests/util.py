"""Shared helpers for the parity tests (error metrics, small synthetic configs)."""
import math

import torch


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_abs(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def mse(a: torch.Tensor, b: torch.Tensor) -> float:
    return float(((a.detach().double().cpu() - b.detach().double().cpu()) ** 2).mean())


def psnr_255(a: torch.Tensor, b: torch.Tensor) -> float:
    """PSNR on the 0..255 scale after the reference's postprocess (t2v_pipeline.rs:147-155)."""
    pa = (a.detach().double().cpu() * 0.5 + 0.5).clamp(0, 1) * 255
    pb = (b.detach().double().cpu() * 0.5 + 0.5).clamp(0, 1) * 255
    m = float(((pa - pb) ** 2).mean())
    return 99.0 if m == 0 else 10.0 * math.log10(255.0 ** 2 / m)


# ---- diffusers -> official (unified-file) tensor names: the inverse of weight_format.rs, for building test files ----
def to_official_dit(k: str) -> str:
    k = k.replace("proj_in", "patchify_proj").replace("time_embed", "adaln_single")
    k = k.replace("norm_q", "q_norm").replace("norm_k", "k_norm")
    return "model.diffusion_model." + k


def official_vae_keys(w):
    """{official name: tensor} for a diffusers-named decoder state dict: mid_block -> up_blocks.0,
    up_blocks.i.upsamplers.0 -> up_blocks.(2i+1), up_blocks.i -> up_blocks.(2i+2), resnets -> res_blocks, the
    decoder-level time embedder / scale-shift table get their `last_` prefix, statistics their native names."""
    import re
    out = {}
    for k, v in w.items():
        if k in ("latents_mean", "latents_std"):
            nk = "per_channel_statistics." + ("mean-of-means" if k == "latents_mean" else "std-of-means")
        elif k.startswith("decoder.mid_block"):
            nk = k.replace("decoder.mid_block", "decoder.up_blocks.0")
        elif k.startswith("decoder.up_blocks."):
            m = re.match(r"decoder\.up_blocks\.(\d+)(\.upsamplers\.0)?(.*)", k)
            i = int(m.group(1))
            nk = f"decoder.up_blocks.{2 * i + 1 if m.group(2) else 2 * i + 2}{m.group(3)}"
        elif k.startswith("decoder.time_embedder"):
            nk = k.replace("decoder.time_embedder", "decoder.last_time_embedder")
        elif k == "decoder.scale_shift_table":
            nk = "decoder.last_scale_shift_table"
        else:
            nk = k
        out["vae." + nk.replace("resnets", "res_blocks")] = v
    return out
