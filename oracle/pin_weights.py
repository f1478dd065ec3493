"""Portable deterministic tensors for the ORACLE PINNING hand-off (test infrastructure).

The reference (Rust + Candle) cannot run in this image, so nothing here is checked against it yet.  This module
defines synthetic weights / inputs that a Rust test can reproduce BIT FOR BIT without torch: every value is

    hash(seed_of_tensor, flat_index)  ->  a small integer  ->  an exactly representable bf16 number,

with splitmix64 as the hash and FNV-1a (64 bit) of the tensor's key as its seed.  `integration/b200_parity.rs` is the
Rust twin (same constants, same known-answer values): dropped into the reference crate's tests/ directory it builds
the same DiT / VAE-decoder weights, runs candle-video's own CPU f32 forward / decode and writes
`b200_parity_dump.safetensors`; tests/test_reference_pin.py compares the oracle (and, on a GPU box, the CUDA path) with
that dump whenever it is present.  Until a maintainer runs it the oracle stays "parity unpinned".
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch

from oracle import ltx_oracle as O

MASK = (1 << 64) - 1
GOLDEN = 0x9E3779B97F4A7C15


def fnv1a64(s: str) -> int:
    h = 0xCBF29CE484222325
    for b in s.encode():
        h = ((h ^ b) * 0x100000001B3) & MASK
    return h


def splitmix64(seed: int, idx: np.ndarray) -> np.ndarray:
    """z = seed + idx * GOLDEN; z = (z ^ z>>30) * C1; z = (z ^ z>>27) * C2; return z ^ z>>31   (mod 2^64)."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx.astype(np.uint64) * np.uint64(GOLDEN)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _pow2_bound(fan_in: int) -> float:
    """2^-ceil(log2(sqrt(fan_in))): the power of two at or below 1/sqrt(fan_in)."""
    k = 0
    while (1 << (2 * k)) < fan_in:
        k += 1
    return 2.0 ** -k


def pin_tensor(key: str, shape: Tuple[int, ...], base_seed: int = 0) -> torch.Tensor:
    """Value rules (all results are exact in bf16, hence identical after any f32 <-> bf16 round trip):
      * `...norm_q.weight` / `...norm_k.weight`: 1 + (h>>59 - 16) / 64          in [0.75, 1.234]
      * `...timestep_scale_multiplier`:          1000
      * everything else:  ((h>>57) * 2 - 127) / 128 * bound   (7-bit odd numerators: never zero)
            bound = 2^-5 for `.bias`, pow2_bound(last dim) for `scale_shift_table`,
                    pow2_bound(prod(shape[1:])) for weights, 2 for inputs (keys starting with "input.")
    """
    n = int(np.prod(shape)) if len(shape) else 1
    h = splitmix64(fnv1a64(key) ^ base_seed, np.arange(n, dtype=np.uint64))
    if key.endswith("timestep_scale_multiplier"):
        v = np.full(n, 1000.0, dtype=np.float32)
    elif "norm_q" in key or "norm_k" in key:
        v = (1.0 + ((h >> np.uint64(59)).astype(np.float32) - 16.0) / 64.0).astype(np.float32)
    else:
        if key.startswith("input."):
            bound = 2.0
        elif key.endswith(".bias"):
            bound = 2.0 ** -5
        elif key.endswith("scale_shift_table"):
            bound = _pow2_bound(shape[-1])
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            bound = _pow2_bound(fan_in)
        q = (h >> np.uint64(57)).astype(np.float32) * 2.0 - 127.0
        v = (q / 128.0 * bound).astype(np.float32)
    t = torch.from_numpy(v.reshape(shape if len(shape) else ()))
    assert torch.equal(t, t.bfloat16().float()), key  # exactly representable by construction
    return t


# ---- the pinned geometries (mirrored in integration/b200_parity.rs) ----
PIN_DIT = dict(num_attention_heads=4, attention_head_dim=64, cross_attention_dim=256, num_layers=2, caption_channels=256)
PIN_DIT_GRID = (2, 8, 8)   # F, H, W -> S = 128 tokens
PIN_DIT_TEXT = (16, 10)    # K text tokens, of which the first 10 are valid
PIN_DIT_TIMESTEP = 992.0   # exactly representable in bf16
PIN_VAE_LAYERS = (1, 1, 1, 1)
PIN_VAE_LATENT = (2, 4, 4)  # the reference's own VAE test shape [1,128,2,4,4] (tests/verify_vae_decode_parity.rs:41-44)
PIN_VAE_TEMB = 0.05


def pin_dit() -> Tuple[O.DitConfig, Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    cfg = O.DitConfig(**PIN_DIT)
    w = {k: pin_tensor(k, s) for k, s in O.dit_weight_shapes(cfg).items()}
    F, H, W = PIN_DIT_GRID
    K, keep = PIN_DIT_TEXT
    mask = torch.zeros(1, K)
    mask[:, :keep] = 1
    inp = {
        "hidden_states": pin_tensor("input.dit.hidden_states", (1, F * H * W, cfg.in_channels)),
        "encoder_hidden_states": pin_tensor("input.dit.encoder_hidden_states", (1, K, cfg.caption_channels)),
        "timestep": torch.tensor([PIN_DIT_TIMESTEP]),
        "encoder_attention_mask": mask,
        "video_coords": O.video_coords(1, F, H, W, 25),
    }
    return cfg, w, inp


def pin_vae() -> Tuple[O.VaeConfig, Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    cfg = O.VaeConfig(decoder_layers_per_block=PIN_VAE_LAYERS)
    w = {k: pin_tensor(k, s) for k, s in O.vae_weight_shapes(cfg).items()}
    F, H, W = PIN_VAE_LATENT
    inp = {"latents": pin_tensor("input.vae.latents", (1, 128, F, H, W)), "temb": torch.tensor([PIN_VAE_TEMB])}
    return cfg, w, inp


# known-answer values: the Rust twin asserts the same numbers before it trusts its generator
KAT = {
    "fnv1a64('proj_in.weight')": fnv1a64("proj_in.weight"),
    "splitmix64(1, 0..3)": [int(v) for v in splitmix64(1, np.arange(3, dtype=np.uint64))],
}
