#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native LTX-Video hot path (BASELINE.json: DiT steps/s + VAE frames/s
at 512x768x97, LTX-Video 2B bf16, CFG flow-matching denoise + 3D-VAE decode; `configs[1]`).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (libltxv_b200.so through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

A "step" is one denoise step of LtxPipeline::call (t2v_pipeline.rs:860-994): the 2 DiT forwards of CFG (uncond + cond;
the reference runs them one after the other, here they are ONE forward over 2 x 4992 tokens that produces the same
bits, tests/test_gpu_pipeline.py::test_batched_cfg_pair_equals_sequential_forwards) + CFG combine + Euler update on a
[4992,128] latent; the VAE decode (97 frames) is timed in the same run and reported beside it.  Synthetic
latents/embeddings, random-init weights of the named architecture.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

HEIGHT, WIDTH, FRAMES, FPS, K_TEXT = 512, 768, 97, 25, 128
GUIDANCE = 3.0  # preset 0.9.5 (configs.rs:165-171); stg_scale 0 here, the STG variant is reported separately


def dit_flops(S, D=2048, L=28, K=128):
    """SURVEY.md 8(d) / BASELINE.md 3: algorithmic FLOPs of one DiT forward."""
    return (L * (28 * S * D * D + 4 * K * D * D + 4 * S * S * D + 4 * S * K * D) + 4 * S * 128 * D + 2 * K * 4096 * D
            + 2 * K * D * D + 2 * (256 * D + 7 * D * D))


def vae_flops(F, H, W):
    """Sum over the decoder's 45 convs of 2*Cin*Cout*27*T*H*W."""
    tot = 0
    T, h, w = F, H, W
    tot += 2 * 128 * 1024 * 27 * T * h * w
    ch = [1024, 512, 256, 128]
    for l in range(4):
        if l > 0:
            tot += 2 * ch[l - 1] * (8 * ch[l]) * 27 * T * h * w
            T, h, w = 2 * T - 1, 2 * h, 2 * w
        tot += 10 * 2 * ch[l] * ch[l] * 27 * T * h * w
    tot += 2 * 128 * 48 * 27 * T * h * w
    return tot


def vae_encode_flops(frames, height, width):
    """Sum over the 0.9.5 encoder's convs of 2*Cin*Cout*27*T*H*W (vae.rs:68-103 defaults; conv_in counted at its
    algorithmic 48 input channels, conv_out at 129 outputs; the temporal downsamplers run on T+1 frames)."""
    ch, layers = [128, 256, 512, 1024, 2048], [4, 6, 6, 2, 1]
    strides = [(1, 2, 2), (2, 1, 1), (2, 2, 2), (2, 2, 2)]
    T, h, w = frames, height // 4, width // 4
    tot = 2 * 48 * ch[0] * 27 * T * h * w
    for l in range(5):
        tot += layers[l] * 2 * 2 * ch[l] * ch[l] * 27 * T * h * w
        if l < 4:
            st, sh, sw = strides[l]
            tot += 2 * ch[l] * (ch[l + 1] // (st * sh * sw)) * 27 * (T + st - 1) * h * w
            T, h, w = (T + st - 1) // st, h // sh, w // sw
    tot += 2 * ch[4] * 129 * 27 * T * h * w
    return tot


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms",
                                          "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm_sorted = sorted(sm)
        # median over the busier half of the samples = "under load"
        load = sm_sorted[len(sm_sorted) // 2:]
        return {"sm_mhz": load[len(load) // 2], "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU path, restated (oracle/ltx_oracle.py).  Two measurements:
#   * cpu_block_sample: ONE transformer block of the 2B DiT at c2's S = 4992 -- the bounded sample of the headline
#     workload (a full c2 CFG step is 56 such blocks, about a minute of host time);
#   * cpu_c1_workload: BASELINE configs[0] IN FULL (28-layer DiT forward at S = 384 + full-depth VAE decode of a 4x8x12
#     latent, f32, best of 3 after one warm-up) -- nothing extrapolated.
# ----------------------------------------------------------------------------------------------------------------
def cpu_block_sample(n_iter: int, warm: int):
    """Time ONE transformer block of the 2B DiT at S=4992 (f32, all host threads).  Returns (seconds per block list,
    cores)."""
    import torch
    from oracle import ltx_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.DitConfig(num_layers=1)
    F, H, W = (FRAMES - 1) // 8 + 1, HEIGHT // 32, WIDTH // 32
    S, D = F * H * W, cfg.inner_dim
    g = torch.Generator().manual_seed(0)
    shapes = {k: v for k, v in O.dit_weight_shapes(cfg).items() if k.startswith("transformer_blocks.0.")}
    w = {k: O._init_tensor(k, s, g) for k, s in shapes.items()}
    x = torch.randn(1, S, D, generator=g)
    enc = torch.randn(1, K_TEXT, D, generator=g)
    temb = torch.randn(1, 6 * D, generator=g) * 0.1
    cos, sin = O.rope_cos_sin(O.video_coords(1, F, H, W, FPS), D)
    mask_bias = torch.zeros(1, 1, K_TEXT)
    times = []
    with torch.no_grad():
        for i in range(warm + n_iter):
            t0 = time.perf_counter()
            O.transformer_block(w, "transformer_blocks.0.", cfg, x, enc, temb, (cos, sin), mask_bias)
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
    return times, cores


C1 = {"height": 256, "width": 384, "num_frames": 25, "latent": (4, 8, 12), "text_tokens": 128}


def c1_oracle_model(seed=1):
    """Oracle weights of the full 2B DiT (28 layers) and the full-depth VAE decoder for the c1 workload.  The 28
    blocks are distinct tensors (own storage: realistic host memory traffic) cloned from two seeded blocks -- drawing
    1.9 G independent values would take longer than the measurement."""
    import torch
    from oracle import ltx_oracle as O
    cfg = O.DitConfig()
    g = torch.Generator().manual_seed(seed)
    one = O.DitConfig(num_layers=2)
    w = {k: O._init_tensor(k, s, g) for k, s in O.dit_weight_shapes(one).items()}
    for i in range(2, cfg.num_layers):
        for k in [k for k in w if k.startswith(f"transformer_blocks.{i % 2}.")]:
            w[k.replace(f"transformer_blocks.{i % 2}.", f"transformer_blocks.{i}.", 1)] = w[k].clone()
    vcfg = O.VaeConfig()
    vw = O.init_vae_weights(vcfg, seed + 1)
    return cfg, w, vcfg, vw


def c1_inputs(seed=2):
    import torch
    from oracle import ltx_oracle as O
    F, H, W = C1["latent"]
    g = torch.Generator().manual_seed(seed)
    hidden = torch.randn(1, F * H * W, 128, generator=g)
    enc = torch.randn(1, C1["text_tokens"], 4096, generator=g)
    mask = torch.ones(1, C1["text_tokens"])
    mask[:, 48:] = 0
    coords = O.video_coords(1, F, H, W, FPS)
    z = torch.randn(1, 128, F, H, W, generator=g)
    return hidden, enc, mask, coords, torch.tensor([993.0]), z, torch.tensor([0.05])


def cpu_c1_workload(repeats=3):
    """BASELINE configs[0] in full on the host cores (f32 oracle restatement), best of `repeats` after one warm-up.
    Returns (result dict, oracle outputs for the parity check, model, inputs)."""
    import torch
    from oracle import ltx_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg, w, vcfg, vw = c1_oracle_model()
    hidden, enc, mask, coords, t, z, ts = c1_inputs()
    F, H, W = C1["latent"]
    t_dit, t_vae = [], []
    vel = frames = None
    with torch.no_grad():
        for i in range(repeats + 1):
            t0 = time.perf_counter()
            vel = O.dit_forward(w, cfg, hidden, enc, t, mask, F, H, W, None, coords, timestep_to_bf16=True)
            t1 = time.perf_counter()
            frames = O.vae_decode(vw, vcfg, z, ts)
            t2 = time.perf_counter()
            if i > 0:
                t_dit.append(t1 - t0)
                t_vae.append(t2 - t1)
    S = F * H * W
    res = {"workload": "BASELINE configs[0]: 2B DiT forward (28 layers, S=384, K=128) + full VAE decode (latent 4x8x12 -> "
                       "25x256x384), f32, run in full (nothing extrapolated)",
           "dit_forward_s": min(t_dit), "vae_decode_s": min(t_vae), "repeats": repeats, "cores": cores,
           "dit_forwards_per_s": 1.0 / min(t_dit), "vae_frames_per_s": 25.0 / min(t_vae),
           "dit_gflops": dit_flops(S) / min(t_dit) / 1e9, "vae_gflops": vae_flops(F, H, W) / min(t_vae) / 1e9,
           "kind": "port"}
    return res, (vel, frames), (cfg, w, vcfg, vw), (hidden, enc, mask, coords, t, z, ts)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, cores = cpu_block_sample(args.steps, args.warmup)
    t_blk = sum(times) / len(times)
    sps = 1.0 / (t_blk * 28 * 2)
    sample = ("each bench step = 1 of the 56 transformer-block passes of a c2 CFG step (28 blocks x 2 forwards) of the 2B "
              "DiT at S=4992, f32 torch-CPU restatement of candle-video's CPU path; value = 1 / (mean block time x 56); "
              "ms_per_step is the MEASURED time of one sample step, not of a full CFG step")
    line = {
        "impl": "reference", "metric": "dit_denoise_steps_per_s", "value": sps, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * t_blk, "sample_fraction_of_step": 1.0 / 56.0,
        "extrapolated_ms_per_full_step": 1000.0 * t_blk * 56, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": sps, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference binary cannot be built here (no cargo/rustc; Candle not vendored): oracle port timed",
    }
    if not args.no_c1:
        line["c1_full"], _, _, _ = cpu_c1_workload(3)
    print(json.dumps(line))


CONFIG_NAME = "c2"


def workload_config():
    F, H, W = (FRAMES - 1) // 8 + 1, HEIGHT // 32, WIDTH // 32
    fw = 2 if GUIDANCE > 1.0 else 1
    text = {"c2": "LTX-Video 2B 0.9.5 arch, CFG denoise step (2 DiT forwards + combine + Euler) + 3D-VAE decode at "
                  "512x768x97 (BASELINE configs[1])"}.get(CONFIG_NAME, "LTX-Video " + CONFIGS[CONFIG_NAME][5])
    return {"workload": text, "name": CONFIG_NAME,
            "height": HEIGHT, "width": WIDTH, "num_frames": FRAMES, "latent": [F, H, W], "tokens": F * H * W,
            "text_tokens": K_TEXT, "guidance_scale": GUIDANCE, "forwards_per_step": fw,
            "cache": "no explicit L2 flush: every step streams the model's bf16 weights (3.8 GB for 2B, >> 126 MB L2)"}



# ----------------------------------------------------------------------------------------------------------------
# The other BASELINE configurations (configs[2..4]) and the strong-scaling suite of the sharded design
# ----------------------------------------------------------------------------------------------------------------
CONFIGS = {
    # name: (DiT preset, height, width, frames, guidance_scale, what it is)
    "c2": ("2b", 512, 768, 97, 3.0, "2B 0.9.5, CFG, 512x768x97 (configs[1])"),
    "c3": ("2b", 704, 1216, 121, 1.0, "2B 0.9.8-distilled, 1 forward/step, 704x1216x121 (configs[2])"),
    # BASELINE.json says 720x1280: 720 is not a multiple of 32 and the reference rejects it (t2v_pipeline.rs:323-327)
    "c4": ("13b", 736, 1280, 161, 3.0, "13B 0.9.8, CFG, 736x1280x161 (configs[3]; 720 is not divisible by 32)"),
    "c5": (None, 704, 1216, 257, 0.0, "VAE decode only, 128-ch latents -> 1216x704x257 (configs[4])"),
}


def config_dims(name):
    preset, h, w, fr, g, _ = CONFIGS[name]
    F, H, W = (fr - 1) // 8 + 1, h // 32, w // 32
    return preset, h, w, fr, g, F, H, W


def measure_config(cv, torch, dev, name, dit, vae, n_steps, comm=None, sync=None, maxr=None, compare=True,
                   nosplit_check=False, breakdown=False):
    """One BASELINE configuration on this rank (comm None) or sharded over all ranks of `comm`.  Returns a compact
    dict: ms/step (n_steps timed after a 1-step warm-up), ms/decode (best effort: 1 warm-up + 2 timed), and -- when
    sharded -- how the sharded result compares with the single-GPU one on the same inputs."""
    preset, h, w, fr, g, F, H, W = config_dims(name)
    S = F * H * W
    sync = sync or (lambda: torch.cuda.synchronize())
    maxr = maxr or (lambda x: x)
    gen = torch.Generator().manual_seed(1000 + S)
    lat0 = torch.randn(S, 128, generator=gen)
    out = {"S": S, "frames": 8 * F - 7}
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    if dit is not None:
        pe = torch.randn(K_TEXT, 4096, generator=gen).to(dev)
        ne = torch.randn(K_TEXT, 4096, generator=gen).to(dev)
        pm = torch.cat([torch.ones(48), torch.zeros(K_TEXT - 48)]).to(dev)
        nm = torch.cat([torch.ones(8), torch.zeros(K_TEXT - 8)]).to(dev)

        def params(n):
            return cv.PipelineParams(height=h, width=w, num_frames=fr, frame_rate=FPS, num_inference_steps=n,
                                     custom_sigmas=[0.9 - 0.1 * i for i in range(n)], guidance_scale=g,
                                     shift_terminal=None)

        def denoise(lat, n, sharded):
            if sharded:
                cv.pipeline_denoise_parallel(dit, comm, params(n), lat, pe, pm, ne if g > 1 else None, nm if g > 1 else None)
            else:
                cv.pipeline_denoise(dit, params(n), lat, pe, pm, ne if g > 1 else None, nm if g > 1 else None)

        def timed(sharded):
            lat = lat0.to(dev).contiguous()
            denoise(lat, 1, sharded)
            lat.copy_(lat0)
            sync()
            e0, e1 = ev(), ev()
            e0.record()
            denoise(lat, n_steps, sharded)
            e1.record()
            sync()
            return maxr(e0.elapsed_time(e1)) / n_steps, lat

        t1, lat1 = timed(False)
        out["single_ms_per_step"] = round(t1, 3)
        out["forwards_per_step"] = 2 if g > 1 else 1
        out["single_tflops"] = round((2 if g > 1 else 1) * dit_flops(S, *DIT_DIMS[preset]) / (t1 * 1e-3) / 1e12, 1)
        if comm is not None:
            tn, latn = timed(True)
            plan = cv.parallel_plan(comm.nranks, comm.rank, S, g > 1.0)
            out.update({"sharded_ms_per_step": round(tn, 3), "speedup": round(t1 / tn, 3),
                        "efficiency": round(t1 / tn / comm.nranks, 3),
                        "plan": f"cfg_groups {plan['cfg_groups']} x ulysses {plan['sp_size']}"})
            if compare:
                d = (latn.double() - lat1.double())
                eq = torch.tensor([int(torch.equal(latn, lat1))], device=dev)
                rl = torch.tensor([float(d.norm() / lat1.double().norm())], device=dev)
                if comm.nranks > 1:
                    import torch.distributed as dist
                    dist.all_reduce(eq, op=dist.ReduceOp.MIN)
                    dist.all_reduce(rl, op=dist.ReduceOp.MAX)
                out["sharded_vs_single_rel_l2"] = float(rl)
                out["sharded_equals_single"] = bool(int(eq))
            if breakdown:
                # where the sharded step spends its time on this rank (CUDA events per kernel class; class "other" =
                # the NVLink flag-barrier kernels: flag round trip + waiting for the slowest partner)
                lat_b = lat0.to(dev).contiguous()
                cv.profile_begin()
                denoise(lat_b, 2, True)
                prof = cv.profile_end()
                if comm.rank == 0:
                    out["sharded_breakdown_ms_per_step"] = {k: round(v["ms"] / 2, 3) for k, v in prof.items() if v["launches"]}
                    out["sharded_breakdown_launches_per_step"] = {k: v["launches"] // 2 for k, v in prof.items() if v["launches"]}
            if nosplit_check:
                # with the attention tail split off both paths run the same instruction sequence per row: bit identity
                cv.set_option("attn_nosplit", 1)
                a = lat0.to(dev).contiguous()
                b = lat0.to(dev).contiguous()
                denoise(a, 2, False)
                denoise(b, 2, True)
                sync()
                cv.set_option("attn_nosplit", 0)
                eq = torch.tensor([int(torch.equal(a, b))], device=dev)
                import torch.distributed as dist
                dist.all_reduce(eq, op=dist.ReduceOp.MIN)
                out["sharded_equals_single_tail_split_off"] = bool(int(eq))
    if vae is not None:
        frames = 8 * F - 7
        lat = lat0.to(dev).contiguous()
        p1 = cv.PipelineParams(height=h, width=w, num_frames=fr, decode_timestep=0.05)
        vid = torch.empty((3, frames, h, w), dtype=torch.float32, device=dev)

        def decode_timed():
            cv.pipeline_decode(vae, p1, lat, out=vid)
            sync()
            e0, e1 = ev(), ev()
            e0.record()
            for _ in range(2):
                cv.pipeline_decode(vae, p1, lat, out=vid)
            e1.record()
            sync()
            return maxr(e0.elapsed_time(e1) / 2)

        t1 = decode_timed()
        out["single_ms_per_decode"] = round(t1, 3)
        out["single_frames_per_s"] = round(frames * 1000.0 / t1, 1)
        out["single_vae_tflops"] = round(vae_flops(F, H, W) / (t1 * 1e-3) / 1e12, 1)
        if comm is not None:
            ref = vid.clone() if compare else None
            cv.vae_set_comm(vae, comm)
            tn = decode_timed()
            cv.vae_set_comm(vae, None)
            out.update({"sharded_ms_per_decode": round(tn, 3), "vae_speedup": round(t1 / tn, 3),
                        "vae_efficiency": round(t1 / tn / comm.nranks, 3), "vae_h_slabs": comm.nranks,
                        "sharded_frames_per_s": round(frames * 1000.0 / tn, 1)})
            if compare and comm.rank == 0:
                out["vae_sharded_equals_single"] = bool(torch.equal(vid, ref))
            del ref
        del vid
    return out


def run_vae_only(args, cv, torch, dist, dev, rank, world, local_rank, barrier, max_over_ranks):
    """--config c5: the decode-only sweep point (configs[4]); headline = decoded frames/s, sharded over all ranks."""
    vae = cv.AutoencoderKLLtxVideo(cv.VaeConfig(), device=local_rank)
    vae.init_random(4321)
    comm = cv.PeerComm(world, rank, local_rank, heap_bytes=20 << 30) if world > 1 else None
    sampler = ClockSampler(local_rank)
    sampler.start()
    r = measure_config(cv, torch, dev, args.config, None, vae, 0, comm=comm, sync=barrier, maxr=max_over_ranks)
    clocks = sampler.stop()
    ms = r.get("sharded_ms_per_decode", r["single_ms_per_decode"])
    if rank == 0:
        line = {"metric": "vae_decoded_frames_per_s", "value": r["frames"] * 1000.0 / ms, "unit": "frames/s",
                "n_gpus": world, "steps": 2, "warmup": 1, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(), "detail": r, "clocks": clocks, "gpu_launches": int(cv.launch_count())}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


DIT_DIMS = {"2b": (2048, 28, K_TEXT), "13b": (4096, 48, K_TEXT)}

# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import candle_video_b200 as cv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    global HEIGHT, WIDTH, FRAMES, GUIDANCE, CONFIG_NAME
    CONFIG_NAME = args.config
    preset, HEIGHT, WIDTH, FRAMES, GUIDANCE, F, H, W = config_dims(args.config)
    if preset is None:
        return run_vae_only(args, cv, torch, dist, dev, rank, world, local_rank, barrier, max_over_ranks)
    S = F * H * W
    frames = 8 * F - 7
    D_, L_, _ = DIT_DIMS[preset]
    dit = cv.LtxVideoTransformer3DModel(cv.DitConfig.preset(preset), device=local_rank)
    dit.init_random(1234)
    vae = cv.AutoencoderKLLtxVideo(cv.VaeConfig(), device=local_rank)
    vae.init_random(4321)

    def make_inputs(seed):
        g = torch.Generator().manual_seed(seed)
        lat = torch.randn(S, 128, generator=g).pin_memory()
        pe_h = torch.randn(K_TEXT, 4096, generator=g).pin_memory()
        ne_h = torch.randn(K_TEXT, 4096, generator=g).pin_memory()
        return lat, pe_h, ne_h

    pm_host = torch.cat([torch.ones(48), torch.zeros(K_TEXT - 48)]).pin_memory()
    nm_host = torch.cat([torch.ones(8), torch.zeros(K_TEXT - 8)]).pin_memory()
    pm, nm = pm_host.to(dev), nm_host.to(dev)

    def params(n_steps):
        if n_steps == 1:
            # single-step calls (e2e leg): an explicit sigma, because linspace(1, 1, 1) stretched to the 0.1 terminal
            # is 0/0 in the reference's schedule as well
            return cv.PipelineParams(height=HEIGHT, width=WIDTH, num_frames=FRAMES, frame_rate=FPS,
                                     num_inference_steps=1, custom_sigmas=[0.9], guidance_scale=GUIDANCE,
                                     guidance_rescale=0.0, stg_scale=0.0, shift_terminal=None, decode_timestep=0.05)
        return cv.PipelineParams(height=HEIGHT, width=WIDTH, num_frames=FRAMES, frame_rate=FPS,
                                 num_inference_steps=n_steps, guidance_scale=GUIDANCE, guidance_rescale=0.0,
                                 stg_scale=0.0, shift_terminal=0.1, decode_timestep=0.05)

    # ------------------------------------------------------------------------------------------------------------
    # parallel modes.  `videos` = videos in flight over the N GPUs.
    #   replicas: one video per GPU, no data-path exchange
    #   pairs   : one video per GPU PAIR -- the two CFG branches run on the two GPUs (the reference runs them as
    #             independent B=1 forwards), velocities exchanged over NVLink peer memory, VAE decode as 2 H-slabs
    #   sharded : ONE video over all N GPUs -- CFG branch groups x Ulysses token shards, VAE decode as N H-slabs
    # ------------------------------------------------------------------------------------------------------------
    def mode_setup(mode):
        if mode == "replicas" or world == 1:
            return None, world, 100 + rank
        if mode == "pairs":
            groups = [dist.new_group([2 * i, 2 * i + 1]) for i in range(world // 2)]
            comm = cv.PeerComm(2, rank % 2, local_rank, heap_bytes=6 << 30, group=groups[rank // 2])
            return comm, world // 2, 100 + rank // 2
        comm = cv.PeerComm(world, rank, local_rank, heap_bytes=6 << 30)
        return comm, 1, 100

    def run_mode(mode, n_steps, count_launches=False):
        """Times exactly n_steps denoise steps and the VAE decode in `mode`; returns a dict (max over ranks)."""
        comm, videos, seed = mode_setup(mode)
        lat_h, pe_h, ne_h = make_inputs(seed)
        lat = lat_h.to(dev)
        pe_d, ne_d = pe_h.to(dev), ne_h.to(dev)

        def denoise(n):
            if comm is None:
                cv.pipeline_denoise(dit, params(n), lat, pe_d, pm, ne_d, nm)
            else:
                cv.pipeline_denoise_parallel(dit, comm, params(n), lat, pe_d, pm, ne_d, nm)

        denoise(max(args.warmup, 3))  # warm-up
        lat.copy_(lat_h)
        barrier()
        l0 = cv.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        denoise(n_steps)
        e1.record()
        barrier()
        launches = cv.launch_count() - l0
        ms_step = max_over_ranks(e0.elapsed_time(e1)) / n_steps
        if comm is not None:
            cv.vae_set_comm(vae, comm)
        n_dec = max(1, min(n_steps, 3))
        # the frames land in ONE preallocated buffer: a fresh 458 MB tensor per decode made the timed loop pay a
        # cudaMalloc whenever the caching allocator had no free block (42 vs 55-59 ms per decode from run to run)
        out = cv.pipeline_decode(vae, params(1), lat)  # warm-up / workspace allocation
        barrier()
        e0.record()
        for _ in range(n_dec):
            cv.pipeline_decode(vae, params(1), lat, out=out)
        e1.record()
        barrier()
        ms_dec = max_over_ranks(e0.elapsed_time(e1) / n_dec)
        if comm is not None:
            cv.vae_set_comm(vae, None)
        finite = bool(torch.isfinite(lat).all().item())
        if comm is None or comm.rank == 0:
            finite = finite and bool(torch.isfinite(out).all().item())
        res = {"mode": mode, "videos_in_flight": videos, "steps_per_s": videos * 1000.0 / ms_step, "ms_per_step": ms_step,
               "vae_frames_per_s": videos * frames * 1000.0 / ms_dec, "vae_ms_per_decode": ms_dec,
               "launches": int(launches), "outputs_finite": finite}
        if comm is not None:
            plan = cv.parallel_plan(comm.nranks, comm.rank, S, GUIDANCE > 1.0)
            res["plan"] = {"cfg_groups": plan["cfg_groups"], "ulysses_size": plan["sp_size"], "vae_h_slabs": comm.nranks}
        del comm
        return res, lat

    if world == 1:
        headline_mode = "single"
    elif args.mode != "auto":
        headline_mode = args.mode
    else:
        # Since the two CFG branches of a step run as ONE batched forward on a GPU (78 row tiles per GEMM), one video per
        # GPU is the throughput-optimal decomposition (N=2: 38.5 vs 37.5 steps/s for pairs); pairs / sharded trade ~3 % /
        # ~40 % of the throughput for 1/2 ... 1/5 of the latency and are reported beside it under "modes".
        headline_mode = "replicas"
    if headline_mode == "pairs" and world % 2 != 0:
        raise SystemExit("bench.py: --mode pairs needs an even number of GPUs")

    sampler = ClockSampler(local_rank)
    sampler.start()
    head, latents = run_mode(headline_mode if world > 1 else "replicas", args.steps)
    clocks = sampler.stop()
    ms_per_step, vae_ms = head["ms_per_step"], head["vae_ms_per_decode"]

    line = {
        "metric": "dit_denoise_steps_per_s", "value": head["steps_per_s"], "unit": "steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if headline_mode == "sharded" else "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": workload_config(),
        "vae_frames_per_s": head["vae_frames_per_s"], "vae_ms_per_decode": vae_ms, "vae_frames": frames,
        "gpu_launches": head["launches"], "outputs_finite": head["outputs_finite"], "clocks": clocks,
    }
    if world > 1:
        desc = {"replicas": f"{world} independent replicas (one video per GPU), no data-path exchange",
                "pairs": f"{world // 2} videos in flight, one per GPU pair: the two CFG branches of a step run on the two "
                         "GPUs (velocity shards exchanged by NVLink peer stores), VAE decode as 2 H-slabs with halo rows "
                         "stored by the producer; no NCCL on the data path",
                "sharded": f"one video over all {world} GPUs: CFG branch groups x Ulysses token shards, VAE as {world} "
                           "H-slabs"}
        line["config"]["parallelism"] = desc[headline_mode]
        line["config"]["mode"] = headline_mode
        # ---- the SHARDED design on the configurations north_star assigns to it: one video over all N
        # GPUs, strong scaling against the single-GPU time measured in this same process on the same inputs ----
        names = [c for c in args.scaling_configs.split(",") if c]
        if args.config not in names:
            names.insert(0, args.config)
        if world == 8 and "c4" not in names and not args.no_c4:
            names.append("c4")
        comm = cv.PeerComm(world, rank, local_rank, heap_bytes=(20 if world == 2 else 14) << 30)
        n_sc = max(2, min(args.steps, 4))
        scal = {}
        for name in names:
            pr = config_dims(name)[0]
            d_ = dit if pr == preset else None
            own = None
            if pr is not None and d_ is None:
                own = cv.LtxVideoTransformer3DModel(cv.DitConfig.preset(pr), device=local_rank)
                own.init_random(1234)
                d_ = own
            try:
                r = measure_config(cv, torch, dev, name, d_, vae if name != "c4" else None, n_sc, comm=comm, sync=barrier,
                                   maxr=max_over_ranks, nosplit_check=(name == "c2"), breakdown=name in ("c2", "c3"))
            except cv.LtxvError as e:  # raised symmetrically on every rank (same arguments): report, keep the run
                r = {"error": str(e)[:300]}
            r["what"] = CONFIGS[name][5]
            scal[name] = r
            del own, d_
            torch.cuda.empty_cache()
        del comm
        line = dict({"strong_scaling": {"n_gpus": world, "steps_timed": n_sc, "configs": scal,
                                        "note": "sharded = ONE video over all N GPUs (CFG branch groups x Ulysses token "
                                                "shards; VAE in N H-slabs with halo rows stored by the producers); "
                                                "single = the same inputs on one GPU, timed in this process"}}, **line)
        if args.mode != "auto" or args.all_modes:
            # the other decompositions of the same N GPUs (same K steps each)
            others = {}
            for m in ("replicas", "pairs", "sharded"):
                if m == headline_mode or (m == "pairs" and (world % 2 != 0 or world == 2)):
                    continue
                r, _ = run_mode(m, args.steps)
                others[m] = r
            line["modes"] = others

    # ---- end to end through the host-buffer C ABI, every rank on its own video: per step H2D latents+embeddings,
    # D2H latents; per decode H2D latents, D2H video ----
    lat_h, pe_h, ne_h = make_inputs(100 + rank)
    lat_e2e = lat_h.clone().pin_memory()
    cv.pipeline_denoise_host(dit, params(1), lat_e2e, pe_h, pm_host, ne_h, nm_host)
    n_e2e = max(1, min(args.steps, 5))
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        cv.pipeline_denoise_host(dit, params(1), lat_e2e, pe_h, pm_host, ne_h, nm_host)
    t_e2e = max_over_ranks((time.perf_counter() - t0) / n_e2e)
    vid_host = torch.empty((3, frames, HEIGHT, WIDTH), dtype=torch.float32).pin_memory()
    cv.pipeline_decode_host(vae, params(1), lat_e2e, vid_host)
    barrier()
    t0 = time.perf_counter()
    cv.pipeline_decode_host(vae, params(1), lat_e2e, vid_host)
    t_vae_e2e = max_over_ranks(time.perf_counter() - t0)
    # the same decode delivering what the reference's example hands to its image / GIF writers (u8 HWC frames,
    # main.rs:653-667): conversion on the device, a quarter of the bytes over PCIe
    vid_u8 = torch.empty((frames, HEIGHT, WIDTH, 3), dtype=torch.uint8).pin_memory()
    cv.pipeline_decode_host_u8(vae, params(1), lat_e2e, vid_u8)
    barrier()
    t0 = time.perf_counter()
    cv.pipeline_decode_host_u8(vae, params(1), lat_e2e, vid_u8)
    t_vae_u8 = max_over_ranks(time.perf_counter() - t0)
    finite_e2e = bool(torch.isfinite(lat_e2e).all().item()) and bool(torch.isfinite(vid_host).all().item())
    line["outputs_finite"] = line["outputs_finite"] and finite_e2e
    line["e2e"] = {"value": world / t_e2e, "unit": "steps/s",
                   "h2d_bytes_per_step": int(lat_h.numel() * 4 + 2 * pe_h.numel() * 4 + 2 * K_TEXT * 4),
                   "d2h_bytes_per_step": int(lat_h.numel() * 4),
                   "api": "ltxv_pipeline_denoise_host (1 step per call: context prep + 2 forwards + Euler), one video "
                          "per GPU, wall clock, max over ranks",
                   "vae_frames_per_s": world * frames / t_vae_e2e, "vae_h2d_bytes": int(lat_h.numel() * 4),
                   "vae_d2h_bytes": int(vid_host.numel() * 4), "vae_api": "ltxv_pipeline_decode_host",
                   "vae_u8_frames_per_s": world * frames / t_vae_u8, "vae_u8_d2h_bytes": int(vid_u8.numel()),
                   "vae_u8_api": "ltxv_pipeline_decode_host_u8 (u8 [F,H,W,3] frames, the example's output hand-off)"}

    FW = 2 if GUIDANCE > 1.0 else 1
    if rank == 0:
        peaks = measured_peaks()
        # ---- instrumented pass: per-kernel-class device time inside a real step (CUDA events on the stream) ----
        lat_p = lat_h.to(dev)
        pe_d, ne_d = pe_h.to(dev), ne_h.to(dev)
        cv.profile_begin()
        cv.pipeline_denoise(dit, params(2), lat_p, pe_d, pm, ne_d, nm)
        prof_dit = cv.profile_end()
        cv.profile_begin()
        cv.pipeline_decode(vae, params(1), lat_p)
        prof_vae = cv.profile_end()
        t1 = 0.0
        if world > 1:  # single-GPU step time for the per-kernel shares
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            cv.pipeline_denoise(dit, params(2), lat_p, pe_d, pm, ne_d, nm)
            e1.record()
            torch.cuda.synchronize()
            t1 = e0.elapsed_time(e1) / 2
        step1 = ms_per_step if world == 1 else t1

        def tf(d):
            return d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 else 0.0

        def hbm(d, per):
            """HBM-bound glue class: achieved algorithmic GB/s against the measured copy bandwidth."""
            gbs = d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0.0
            return {"gb_per_s": gbs, "frac_of_hbm": gbs / peaks["hbm_gbs"], "ms": d["ms"] / per,
                    "launches": d["launches"] // per}

        gm = prof_dit["gemm"]
        traffic = ncu_traffic_bytes()
        line["roofline"] = {
            "kernel": "gemm_bf16_tn_kernel (tcgen05 GEMM, DiT projections/FFN)", "bound": "tensor",
            "achieved": tf(gm), "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
            "frac": tf(gm) / peaks["tf_sustained"], "traffic": traffic,
            "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
            "launches": gm["launches"], "avg_launch_ms": gm["ms"] / max(gm["launches"], 1),
            "share_of_step": gm["ms"] / (2 * step1),
        }
        attn_ms = prof_dit["attn_self"]["ms"] + prof_dit["attn_cross"]["ms"]
        line["roofline_all"] = {
            "dit_gemm": {"tflops": tf(gm), "ms_per_step": gm["ms"] / 2, "frac": tf(gm) / peaks["tf_sustained"]},
            "dit_attn_self": {"tflops": tf(prof_dit["attn_self"]), "ms_per_step": prof_dit["attn_self"]["ms"] / 2,
                              "frac": tf(prof_dit["attn_self"]) / peaks["tf_sustained"]},
            "dit_attn_cross": {"tflops": tf(prof_dit["attn_cross"]), "ms_per_step": prof_dit["attn_cross"]["ms"] / 2},
            "vae_conv3d": {"tflops": tf(prof_vae["conv3d"]), "ms_per_decode": prof_vae["conv3d"]["ms"],
                           "frac": tf(prof_vae["conv3d"]) / peaks["tf_sustained"]},
            "dit_norm_modulate": hbm(prof_dit["norm_modulate"], 2),
            "dit_qk_norm_rope": hbm(prof_dit["qk_norm_rope"], 2),
            "vae_prep": hbm(prof_vae["vae_prep"], 1),
            "single_gpu_ms_per_step": step1,
            "dit_step_algorithmic_tflops": FW * dit_flops(S, D_, L_) / (step1 * 1e-3) / 1e12,
            "dit_step_frac_of_peak": FW * dit_flops(S, D_, L_) / (step1 * 1e-3) / 1e12 / peaks["tf_sustained"],
            "glue_ms_per_step": step1 - (gm["ms"] + attn_ms) / 2,
            "unattributed_ms_per_step": step1 - (gm["ms"] + attn_ms + prof_dit["norm_modulate"]["ms"] +
                                                 prof_dit["qk_norm_rope"]["ms"]) / 2,
        }
        if world == 1 and not args.no_encode:
            # row f-4 (next row of the path): the VAE encoder on the clip the decoder produces, same conv3d kernel
            venc = cv.AutoencoderKLLtxVideo(cv.VaeConfig(), device=local_rank)
            venc.enable_encoder()
            venc.init_random(99)
            clip = torch.tanh(torch.randn(1, 3, FRAMES, HEIGHT, WIDTH, device=dev))
            venc.encode(clip)  # workspace + warm-up
            torch.cuda.synchronize()
            cv.profile_begin()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                venc.encode(clip)
            e1.record()
            torch.cuda.synchronize()
            prof_enc = cv.profile_end()
            enc_ms = e0.elapsed_time(e1) / 2
            line["roofline_all"]["vae_encode"] = {
                "ms_per_encode": enc_ms, "frames_per_s": FRAMES * 1000.0 / enc_ms,
                "algorithmic_tflops": vae_encode_flops(FRAMES, HEIGHT, WIDTH) / (enc_ms * 1e-3) / 1e12,
                "conv3d_tflops": tf(prof_enc["conv3d"]), "conv3d_ms": prof_enc["conv3d"]["ms"] / 2,
                "prep": hbm(prof_enc["vae_prep"], 2), "api": "ltxv_vae_encode (device buffers)"}
            del venc, clip
        if world == 1:
            line["roofline_all"]["vae_decode_algorithmic_tflops"] = vae_flops(F, H, W) / (vae_ms * 1e-3) / 1e12
            line["roofline_all"]["vae_decode_frac_of_peak"] = (vae_flops(F, H, W) / (vae_ms * 1e-3) / 1e12 /
                                                               peaks["tf_sustained"])
            line["dit_forward_ms"] = ms_per_step / FW

        if world == 1 and args.config == "c2":
            # ---- the 0.9.5 preset step (configs.rs:165-171): guidance 3, STG scale 1 on block 19, rescale 0.7 ->
            # 3 forwards per step (t2v_pipeline.rs:910-939); the headline above is the same step with stg_scale = 0 ----
            p_stg = cv.PipelineParams(height=HEIGHT, width=WIDTH, num_frames=FRAMES, frame_rate=FPS,
                                      num_inference_steps=3, guidance_scale=3.0, guidance_rescale=0.7, stg_scale=1.0,
                                      skip_block_list=[19], shift_terminal=0.1)
            lat_s = lat_h.to(dev)
            cv.pipeline_denoise(dit, p_stg, lat_s, pe_d, pm, ne_d, nm)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            cv.pipeline_denoise(dit, p_stg, lat_s, pe_d, pm, ne_d, nm)
            e1.record()
            torch.cuda.synchronize()
            ms_stg = e0.elapsed_time(e1) / 3
            dit.set_skip_block_list([])
            line["stg_preset_step"] = {
                "ms_per_step": ms_stg, "steps_per_s": 1000.0 / ms_stg, "forwards_per_step": 3,
                "algorithmic_tflops": 3 * dit_flops(S) / (ms_stg * 1e-3) / 1e12,
                "what": "0.9.5 preset: CFG 3.0 + STG 1.0 (block 19) + rescale 0.7 (configs.rs:165-171)",
                "outputs_finite": bool(torch.isfinite(lat_s).all().item())}
            # ---- the other BASELINE configurations on one GPU (configs[2..4]) ----
            if not args.no_other_configs:
                oc = {}
                for name in ("c3", "c5", "c4"):
                    if name == "c4" and args.no_c4:
                        continue
                    pr = config_dims(name)[0]
                    own = None
                    if pr == "13b":
                        own = cv.LtxVideoTransformer3DModel(cv.DitConfig.preset("13b"), device=local_rank)
                        own.init_random(1234)
                    d_ = own if own is not None else (dit if pr == "2b" else None)
                    oc[name] = measure_config(cv, torch, dev, name, d_, vae if name != "c4" else None, 2)
                    oc[name]["what"] = CONFIGS[name][5]
                    del own, d_
                    torch.cuda.empty_cache()
                line["other_configs"] = oc

        # ---- CPU baseline + parity (rank 0, N = 1 only) ----
        if world == 1 and not args.no_cpu_baseline:
            times, cores = cpu_block_sample(3, 1)
            t_blk = sum(times) / len(times)
            line["cpu_baseline"] = {
                "value": 1.0 / (t_blk * 56), "unit": "steps/s", "cores": cores, "kind": "port",
                "sample": "1 of the 56 transformer-block passes of a c2 CFG step (2B DiT, S=4992), f32 torch-CPU oracle "
                          f"restatement, mean of 3 after 1 warm-up ({t_blk:.2f} s/block), scaled x28 blocks x2 forwards"}
            line["parity"] = parity_block_c2(cv, dev)
            if not args.no_c1:
                # BASELINE configs[0] in full on the host, and the same workload on the GPU checked against it
                c1, (vel_ref, frames_ref), (cfg1, w1, vcfg1, vw1), (hidden, enc, mask, coords, t, z, ts) = cpu_c1_workload(3)
                line["cpu_baseline"]["c1_full"] = c1
                d1 = cv.LtxVideoTransformer3DModel(cv.DitConfig.preset("2b"), device=local_rank)
                d1.load_state_dict(w1)
                Fc, Hc, Wc = C1["latent"]
                d1.forward(hidden.to(dev), enc.to(dev), t.to(dev), mask.to(dev), Fc, Hc, Wc, None, coords.to(dev))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                vel = d1.forward(hidden.to(dev), enc.to(dev), t.to(dev), mask.to(dev), Fc, Hc, Wc, None, coords.to(dev))
                e1.record()
                torch.cuda.synchronize()
                c1_dit_ms = e0.elapsed_time(e1)
                v1 = cv.AutoencoderKLLtxVideo(cv.VaeConfig(), device=local_rank)
                v1.load_state_dict(vw1)
                v1.decode(z.to(dev), ts.to(dev))
                e0.record()
                fr = v1.decode(z.to(dev), ts.to(dev))
                e1.record()
                torch.cuda.synchronize()
                c1_vae_ms = e0.elapsed_time(e1)
                dv = (vel.cpu().double() - vel_ref.double())
                pa = (fr.cpu().double() * 0.5 + 0.5).clamp(0, 1) * 255
                pb = (frames_ref.double() * 0.5 + 0.5).clamp(0, 1) * 255
                m255 = float(((pa - pb) ** 2).mean())
                line["parity"].update({
                    "c1_dit_28_layers_rel_l2": float(dv.norm() / vel_ref.double().norm()),
                    "c1_dit_max_abs": float(dv.abs().max()),
                    "c1_vae_mse": float(((fr.cpu().double() - frames_ref.double()) ** 2).mean()),
                    "c1_vae_psnr_db": 99.0 if m255 == 0 else 10.0 * __import__("math").log10(255.0 ** 2 / m255),
                    "c1_gpu_dit_forward_ms": c1_dit_ms, "c1_gpu_vae_decode_ms": c1_vae_ms,
                    "tolerance": "DiT rel-L2 <= 2e-2, VAE MSE <= 1e-2 on [-1,1] and PSNR >= 35 dB on 0..255 (SURVEY 8c)"})
                del d1, v1
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def parity_block_c2(cv, dev):
    """One 2B block at c2's token count on the GPU (sequential forward, M = 4992) against the oracle on identical
    weights and inputs: the same check as tests/test_gpu_production_shapes.py, repeated inside the bench run."""
    import torch
    from oracle import ltx_oracle as O
    cfg = O.DitConfig(num_layers=1)
    w = O.init_dit_weights(cfg, 42)
    m = cv.LtxVideoTransformer3DModel(cv.DitConfig(num_layers=1), device=dev.index or 0)
    m.load_state_dict(w)
    F, H, W = (FRAMES - 1) // 8 + 1, HEIGHT // 32, WIDTH // 32
    g = torch.Generator().manual_seed(5)
    hidden = torch.randn(1, F * H * W, 128, generator=g)
    enc = torch.randn(1, K_TEXT, 4096, generator=g)
    mask = torch.ones(1, K_TEXT)
    mask[:, 48:] = 0
    coords = O.video_coords(1, F, H, W, FPS)
    t = torch.tensor([993.0])
    with torch.no_grad():
        ref = O.dit_forward(w, cfg, hidden, enc, t, mask, F, H, W, None, coords, timestep_to_bf16=True)
    cv.trace_begin()
    out = m.forward(hidden.to(dev), enc.to(dev), t.to(dev), mask.to(dev), F, H, W, None, coords.to(dev)).cpu()
    tr = cv.trace_end()
    d = out.double() - ref.double()
    return {"dit_block_c2_rel_l2": float(d.norm() / ref.double().norm()), "dit_block_c2_max_abs": float(d.abs().max()),
            "dit_block_c2_variants": sorted(tr)}


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel class (the DiT GEMMs), mean over
    the 231 GEMM launches of ONE batched-CFG step at the timed shape (M = 9984), from the committed ncu launch list of
    tools/profile_step2.py (profiles/r02e_launches_step_plus_decode.csv -> profiles/r02e_gemm_traffic.json with tools/launch_summary.py --traffic; ncu cannot
    run inside the timed bench, and it flushes the caches before every launch: cold-cache upper bound)."""
    p = ROOT / "profiles" / "r02e_gemm_traffic.json"
    try:
        return json.loads(p.read_text())["bytes_per_launch_mean"]
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c1", action="store_true", help="skip the full BASELINE configs[0] host run / GPU parity check")
    ap.add_argument("--no-encode", action="store_true", help="skip the VAE encoder measurement (roofline_all.vae_encode)")
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE configuration of the headline number")
    ap.add_argument("--scaling-configs", default="c2,c3,c5",
                    help="N > 1: configurations of the strong-scaling suite (c4 = 13B is added at N = 8)")
    ap.add_argument("--no-c4", action="store_true", help="leave the 13B configuration out of the suites")
    ap.add_argument("--no-other-configs", action="store_true", help="N = 1: skip the c3 / c4 / c5 measurements")
    ap.add_argument("--all-modes", action="store_true", help="N > 1: also time the pairs / sharded throughput modes")
    ap.add_argument("--mode", default="auto", choices=["auto", "replicas", "pairs", "sharded"],
                    help="multi-GPU decomposition of the headline number (auto = replicas, the throughput-optimal one)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
