#!/usr/bin/env python
"""One profiled DiT forward (c2: S=4992, 2B) and/or one VAE decode (13x16x24 latent) for ncu.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py --what both
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_bf16 -c 4 \
      -o gpurun_out/prof_gemm python tools/profile_step.py --what dit
"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402

import candle_video_b200 as cv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--what", default="both", choices=["dit", "vae", "both"])
ap.add_argument("--layers", type=int, default=28)
args = ap.parse_args()
dev = torch.device("cuda:0")
F, H, W, K = 13, 16, 24, 128
S = F * H * W
g = torch.Generator().manual_seed(0)
if args.what in ("dit", "both"):
    cfg = cv.DitConfig.preset("2b")
    cfg.num_layers = args.layers
    dit = cv.LtxVideoTransformer3DModel(cfg)
    dit.init_random(1)
    hidden = torch.randn(S, 128, generator=g).to(dev)
    enc = torch.randn(K, 4096, generator=g).to(dev)
    mask = torch.cat([torch.ones(48), torch.zeros(K - 48)]).to(dev)
    coords = cv.video_coords(1, F, H, W, 25, device=dev)
    t = torch.tensor([993.0], device=dev)
    dit.prepare_context(0, enc, mask)
    dit.forward_ctx(0, hidden, t, F, H, W, None, coords)
if args.what in ("vae", "both"):
    vae = cv.AutoencoderKLLtxVideo(cv.VaeConfig())
    vae.init_random(2)
    z = torch.randn(1, 128, F, H, W, generator=g).to(dev)
    ts = torch.tensor([0.05], device=dev)
    vae.decode(z, ts)
torch.cuda.synchronize()
torch.cuda.profiler.start()
if args.what in ("dit", "both"):
    dit.forward_ctx(0, hidden, t, F, H, W, None, coords)
if args.what in ("vae", "both"):
    vae.decode(z, ts)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled region done")
