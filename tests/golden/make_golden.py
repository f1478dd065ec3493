#!/usr/bin/env python
"""Regenerates tests/golden/ltx_golden_v1.safetensors.

PROVENANCE: these vectors are produced by THIS repository's CPU oracle (oracle/ltx_oracle.py, torch CPU f32), not by
candle-video: the reference cannot run in the build container (no Rust toolchain, Candle not vendored) and ships no
numeric fixtures, so float parity stays "unpinned" against the reference.  What the file pins:
  * the oracle against itself over time (tests/test_golden_cpu.py: torch upgrades / refactors must not move it), and
  * the CUDA path against COMMITTED numbers (tests/test_gpu_golden.py) on the shapes of the reference's own parity tests
    (tests/verify_vae_decode_parity.rs:41-44 latents [1,128,2,4,4], temb 0.05; a small DiT like verify_dit_parity.rs).
Weights are not stored: they are re-created from their seeds by oracle.init_*_weights (values are bf16-representable).

  python tests/golden/make_golden.py        # rewrites the fixture
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
from safetensors.torch import save_file  # noqa: E402

from oracle import ltx_oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent / "ltx_golden_v1.safetensors"

DIT_CFG = dict(num_attention_heads=4, attention_head_dim=64, cross_attention_dim=256, num_layers=2, caption_channels=256)
DIT_SHAPE = dict(F=2, H=8, W=8, K=16)
VAE_LAYERS = (1, 1, 1, 1)
ENC_CFG = dict(block_out_channels=(128, 256, 256, 512, 512), layers_per_block=(1, 1, 1, 1, 2))


def dit_case():
    cfg = O.DitConfig(**DIT_CFG)
    w = O.init_dit_weights(cfg, 42)
    g = torch.Generator().manual_seed(0)
    F, H, W, K = (DIT_SHAPE[k] for k in "FHWK")
    hidden = torch.randn(1, F * H * W, cfg.in_channels, generator=g)
    enc = torch.randn(1, K, cfg.caption_channels, generator=g)
    mask = torch.ones(1, K)
    mask[:, 11:] = 0
    coords = O.video_coords(1, F, H, W, 25)
    t = torch.tensor([993.0])
    out = O.dit_forward(w, cfg, hidden, enc, t, mask, F, H, W, None, coords, timestep_to_bf16=True)
    return {"dit.hidden": hidden, "dit.enc": enc, "dit.mask": mask, "dit.coords": coords, "dit.timestep": t, "dit.out": out}


def vae_case():
    cfg = O.VaeConfig(decoder_layers_per_block=VAE_LAYERS)
    w = O.init_vae_weights(cfg, 7)
    z = torch.randn(1, 128, 2, 4, 4, generator=torch.Generator().manual_seed(1))
    ts = torch.tensor([0.05])
    out = O.vae_decode(w, cfg, z, ts)  # [1, 3, 9, 128, 128]
    return {"vae.z": z, "vae.timestep": ts, "vae.out_sub": out[:, :, :, ::4, ::4].contiguous(),
            "vae.out_mean_std": torch.stack([out.mean(), out.std()])}


def enc_case():
    cfg = O.VaeEncoderConfig(**ENC_CFG)
    w = O.init_vae_encoder_weights(cfg, 11)
    x = torch.tanh(torch.randn(1, 3, 9, 64, 96, generator=torch.Generator().manual_seed(2)))
    m = O.vae_encode(w, cfg, x)  # [1, 256, 2, 2, 3]
    return {"enc.x": x, "enc.moments": m}


def glue_case():
    g = torch.Generator().manual_seed(3)
    c, u, p = (torch.randn(1, 96, 128, generator=g) for _ in range(3))
    lat = torch.randn(1, 96, 128, generator=g)
    comb = O.guidance_combine(c, u, p, 3.0, 0.7, 1.0)
    nxt = O.euler_step(lat, comb, 0.9, 0.8)
    sig, ts = O.scheduler_set_timesteps(40, O.calculate_shift(4992))
    return {"glue.cond": c, "glue.uncond": u, "glue.perturbed": p, "glue.latents": lat, "glue.noise_pred": comb,
            "glue.latents_next": nxt, "sched.sigmas": torch.tensor(sig, dtype=torch.float32),
            "sched.timesteps": torch.tensor(ts, dtype=torch.float32)}


def build_all():
    d = {}
    for f in (dit_case, vae_case, enc_case, glue_case):
        d.update(f())
    return {k: v.contiguous().to(torch.float32) for k, v in d.items()}


if __name__ == "__main__":
    t = build_all()
    save_file(t, str(OUT), metadata={"generator": "oracle/ltx_oracle.py (repo CPU oracle, torch f32); NOT candle-video output"})
    print(f"wrote {OUT} ({OUT.stat().st_size / 1024:.0f} KiB, {len(t)} tensors)")
