#!/usr/bin/env python
"""Multi-GPU parity check of the sharded paths, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/mgpu_check.py

Every rank builds the same small random-init DiT / VAE, runs the SINGLE-GPU path (ltxv_pipeline_denoise /
ltxv_vae_decode) as the reference and the sharded path (ltxv_pipeline_denoise_parallel: CFG branch groups x Ulysses
token shards; ltxv_vae_set_comm: H-slabs with halo rows stored into the neighbour's padded buffer) on the same inputs,
and compares.  The sharded path runs the same kernels on the same numbers, so the bar is tight: rel-L2 <= 2e-3 (only
the split-K-free GEMM tiles and the order of the f64 std partial sums may differ).  Exit code != 0 on any mismatch.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main() -> int:
    import candle_video_b200 as cv
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    comm = cv.PeerComm(world, rank, local, heap_bytes=1 << 30)
    comm.barrier()
    torch.cuda.synchronize()
    fails = []

    def report(tag, err, tol):
        ok = err <= tol
        t = torch.tensor([err], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"[mgpu N={world}] {tag}: max-over-ranks rel_l2 = {float(t):.3e} (tol {tol:.0e}) "
                  f"{'ok' if float(t) <= tol else 'FAIL'}", flush=True)
        if float(t) > tol or not ok:
            fails.append(tag)

    # ---------------- DiT denoise loop: CFG split x Ulysses ----------------
    heads = 8  # divisible by every sequence-parallel size up to 8
    dcfg = cv.DitConfig(num_attention_heads=heads, attention_head_dim=64, cross_attention_dim=heads * 64, num_layers=3,
                        caption_channels=256)
    dit = cv.LtxVideoTransformer3DModel(dcfg, device=local)
    dit.init_random(77)
    height, width, frames, K = 256, 256, 25, 24          # latent 4 x 8 x 8 = 256 tokens
    F, H, W = (frames - 1) // 8 + 1, height // 32, width // 32
    S = F * H * W
    g = torch.Generator().manual_seed(5)
    lat0 = torch.randn(S, 128, generator=g)
    pe, ne = torch.randn(K, 256, generator=g), torch.randn(K, 256, generator=g)
    pm, nm = torch.ones(K), torch.ones(K)
    pm[17:] = 0
    nm[5:] = 0
    pe, ne, pm, nm = pe.to(dev), ne.to(dev), pm.to(dev), nm.to(dev)
    cases = [("cfg", 3.0, 0.0, 0.0, None), ("cfg+rescale+stg", 3.0, 0.7, 1.0, [1]), ("no-cfg (pure Ulysses)", 1.0, 0.0, 0.0, None)]
    for tag, gs, r, stg, skip in cases:
        params = cv.PipelineParams(height=height, width=width, num_frames=frames, frame_rate=25,
                                   num_inference_steps=3, guidance_scale=gs, guidance_rescale=r, stg_scale=stg,
                                   skip_block_list=skip)
        dit.set_skip_block_list([])
        ref = lat0.to(dev).contiguous()
        cv.pipeline_denoise(dit, params, ref, pe, pm, ne, nm)
        dit.set_skip_block_list([])
        out = lat0.to(dev).contiguous()
        cv.pipeline_denoise_parallel(dit, comm, params, out, pe, pm, ne, nm)
        torch.cuda.synchronize()
        plan = cv.parallel_plan(world, rank, S, gs > 1.0)
        if rank == 0:
            print(f"[mgpu N={world}] plan for '{tag}': {plan}", flush=True)
        assert torch.isfinite(out).all()
        report(f"denoise {tag}", rel_l2(out, ref), 2e-3)

    # ---------------- VAE decode: H slabs with halo exchange ----------------
    vae = cv.AutoencoderKLLtxVideo(cv.VaeConfig(decoder_layers_per_block=(1, 1, 1, 1)), device=local)
    vae.init_random(9)
    Hl = 8  # latent rows: divisible by 2, 4, 8 ranks
    z = torch.randn(1, 128, 3, Hl, 6, generator=g).to(dev)
    ts = torch.tensor([0.05], device=dev)
    ref = vae.decode(z, ts)
    torch.cuda.synchronize()
    cv.vae_set_comm(vae, comm)
    out = vae.decode(z, ts)
    out2 = vae.decode(z, ts)  # second call reuses the carved buffers
    torch.cuda.synchronize()
    cv.vae_set_comm(vae, None)
    if rank == 0:
        e1, e2 = rel_l2(out, ref), rel_l2(out2, ref)
        print(f"[mgpu N={world}] vae slab decode: rel_l2 = {e1:.3e} / {e2:.3e} (rank 0 holds the video)", flush=True)
        if not (e1 <= 2e-3 and e2 <= 2e-3):
            fails.append("vae slabs")
    dist.barrier()
    nf = torch.tensor([len(fails)], device=dev)
    dist.all_reduce(nf, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"[mgpu N={world}] {'ALL OK' if int(nf) == 0 else 'FAILED: ' + ', '.join(fails)}", flush=True)
    del comm
    dist.destroy_process_group()
    return 1 if int(nf) else 0


if __name__ == "__main__":
    sys.exit(main())
