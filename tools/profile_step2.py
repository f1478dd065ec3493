#!/usr/bin/env python
"""The step bench.py TIMES, for ncu: one batched-CFG denoise step at c2 (ltxv_pipeline_denoise: ONE 2S = 9984-token forward
of the 28-layer 2B DiT + combine + Euler) and/or one VAE decode of the 13x16x24 latent, after a warm-up call.

  # launch list with DRAM bytes per launch (cold-cache, serialised: compare SHARES)
  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step2.py --what both
  # full set of one kernel
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:flash_attn3 -c 2 \
      -o gpurun_out/prof_attn python tools/profile_step2.py --what dit --layers 2
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
height, width, frames, K = 512, 768, 97, 128
F, H, W = 13, 16, 24
S = F * H * W
g = torch.Generator().manual_seed(0)
params = cv.PipelineParams(height=height, width=width, num_frames=frames, frame_rate=25, num_inference_steps=1,
                           custom_sigmas=[0.9], guidance_scale=3.0, shift_terminal=None, decode_timestep=0.05)
lat = torch.randn(S, 128, generator=g).to(dev)
if args.what in ("dit", "both"):
    cfg = cv.DitConfig.preset("2b")
    cfg.num_layers = args.layers
    dit = cv.LtxVideoTransformer3DModel(cfg)
    dit.init_random(1)
    pe, ne = torch.randn(K, 4096, generator=g).to(dev), torch.randn(K, 4096, generator=g).to(dev)
    pm = torch.cat([torch.ones(48), torch.zeros(K - 48)]).to(dev)
    nm = torch.cat([torch.ones(8), torch.zeros(K - 8)]).to(dev)
    cv.pipeline_denoise(dit, params, lat.clone(), pe, pm, ne, nm)
if args.what in ("vae", "both"):
    vae = cv.AutoencoderKLLtxVideo(cv.VaeConfig())
    vae.init_random(2)
    out = cv.pipeline_decode(vae, params, lat)
torch.cuda.synchronize()
torch.cuda.profiler.start()
if args.what in ("dit", "both"):
    cv.pipeline_denoise(dit, params, lat.clone(), pe, pm, ne, nm)
if args.what in ("vae", "both"):
    cv.pipeline_decode(vae, params, lat, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled region done")
