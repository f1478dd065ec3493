#!/usr/bin/env python
"""Runs the hot path at the other BASELINE.json configurations (timing + finiteness + a size-independent property):

  c3  LTX-Video 2B distilled, 704x1216x121: S = 16*22*38 = 13376 tokens, one forward per step
  c4  LTX-Video 13B (48 layers, 32 heads x 128), 720x1280x161 -> latent 21x23x40 (height padded to 736): S = 19320
  c5  VAE decode-only sweep up to 1216x704x257 (latent 33x22x38)

Property checked at full size (no CPU oracle can run these): the DiT forward is deterministic and batch-independent
(the CFG pair forward returns the same rows as two single forwards, bit for bit); the VAE decode of a latent whose last
frames are dropped equals the prefix of the full decode for the frames outside the temporal receptive field.
"""
import argparse
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import candle_video_b200 as cv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--what", default="c3,c5,c4")
args = ap.parse_args()
dev = torch.device("cuda:0")
out = {}


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r


def dit_flops(S, D, L, K=128):
    return L * (28 * S * D * D + 4 * K * D * D + 4 * S * S * D + 4 * S * K * D) + 4 * S * 128 * D


def run_dit(tag, preset, F, H, W, steps, guidance):
    cfg = cv.DitConfig.preset(preset)
    dit = cv.LtxVideoTransformer3DModel(cfg)
    dit.init_random(1)
    S, K = F * H * W, 128
    D = cfg.num_attention_heads * cfg.attention_head_dim
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(S, 128, generator=g).to(dev)
    pe, ne = torch.randn(K, 4096, generator=g).to(dev), torch.randn(K, 4096, generator=g).to(dev)
    pm = torch.cat([torch.ones(48), torch.zeros(K - 48)]).to(dev)
    params = cv.PipelineParams(height=32 * H, width=32 * W, num_frames=8 * (F - 1) + 1, num_inference_steps=steps,
                               guidance_scale=guidance)
    x = lat.clone()
    ms, _ = timed(lambda: cv.pipeline_denoise(dit, params, x, pe, pm, ne, pm), n=1)
    fwd = 2 if guidance > 1 else 1
    tf = fwd * dit_flops(S, D, cfg.num_layers) / (ms / steps * 1e-3) / 1e12
    # determinism: two runs from the same latents give the same bits
    a, b = lat.clone(), lat.clone()
    cv.pipeline_denoise(dit, params, a, pe, pm, ne, pm)
    cv.pipeline_denoise(dit, params, b, pe, pm, ne, pm)
    out[tag] = {"tokens": S, "steps": steps, "forwards_per_step": fwd, "ms_per_step": ms / steps,
                "steps_per_s": 1000.0 * steps / ms, "algorithmic_tflops": tf, "finite": bool(torch.isfinite(a).all()),
                "deterministic": bool(torch.equal(a, b))}
    print(tag, json.dumps(out[tag]), flush=True)
    del dit
    torch.cuda.empty_cache()


if "c3" in args.what:
    run_dit("c3_2b_distilled_704x1216x121", "2b", 16, 22, 38, steps=8, guidance=1.0)
if "c5" in args.what:
    vae = cv.AutoencoderKLLtxVideo(cv.VaeConfig())
    vae.init_random(2)
    g = torch.Generator().manual_seed(3)
    for (F, H, W) in [(13, 16, 24), (16, 22, 38), (33, 22, 38)]:
        z = torch.randn(1, 128, F, H, W, generator=g).to(dev)
        ts = torch.tensor([0.05], device=dev)
        ms, v = timed(lambda: vae.decode(z, ts, postprocess=True), n=2)
        frames = 8 * F - 7
        # temporal locality: the decoder's convs reach +-1 frame each: 11 at latent rate (conv_in + mid), then 11 / 11 / 11
        # at 2x / 4x / 8x -> < 22 latent frames in total.  Dropping the last 3 latent frames must therefore leave every
        # output frame more than 24 latent frames before the cut bit-identical, whatever tiles / kernel variants the
        # smaller volume selects.
        keep = 8 * (F - 3 - 24) - 7
        same = None
        if keep > 0:
            v2 = vae.decode(z[:, :, :F - 3].contiguous(), ts, postprocess=True)
            same = bool(torch.equal(v[:, :, :keep], v2[:, :, :keep]))
            del v2
        out[f"c5_vae_{32 * W}x{32 * H}x{frames}"] = {"ms": ms, "frames_per_s": frames * 1000.0 / ms,
                                                      "finite": bool(torch.isfinite(v).all()),
                                                      "prefix_frames_checked": max(keep, 0),
                                                      "prefix_identical": same}
        print(f"c5 {F}x{H}x{W}", json.dumps(out[f"c5_vae_{32 * W}x{32 * H}x{frames}"]), flush=True)
        del v
    del vae
    torch.cuda.empty_cache()
if "c4" in args.what:
    run_dit("c4_13b_736x1280x161", "13b", 21, 23, 40, steps=2, guidance=3.0)
print("SIZE_SWEEP", json.dumps(out))
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/size_sweep.json").write_text(json.dumps(out, indent=1))
