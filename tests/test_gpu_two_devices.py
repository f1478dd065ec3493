"""Two models on two GPUs inside ONE process (the layout of a candle host that drives several devices from one binary,
INTEGRATION.md section 5): every per-call workspace lives in the model handle or is keyed by (device, stream), so calls
on cuda:0 and cuda:1 may interleave, and two host threads may drive the two models at the same time.  The same seed on
both devices must give the same bits as each device alone.  Skipped on single-GPU boxes."""
import threading

import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(dev, seed=21):
    g = torch.Generator().manual_seed(seed)
    height, width, frames, K = 256, 384, 25, 24
    F, H, W = (frames - 1) // 8 + 1, height // 32, width // 32
    lat = torch.randn(F * H * W, 128, generator=g).to(dev)
    pe, ne = torch.randn(K, 256, generator=g).to(dev), torch.randn(K, 256, generator=g).to(dev)
    pm, nm = torch.ones(K), torch.ones(K)
    pm[17:] = 0
    nm[5:] = 0
    z = torch.randn(1, 128, F, H, W, generator=g).to(dev)
    return (height, width, frames), lat, pe, pm.to(dev), ne, nm.to(dev), z


def test_models_on_two_devices_interleaved_and_threaded(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import candle_video_b200 as cv

    def small_cfg():
        return cv.DitConfig(num_attention_heads=4, attention_head_dim=64, cross_attention_dim=256, num_layers=3,
                            caption_channels=256)

    def make(d):
        with torch.cuda.device(d):
            dit = cv.LtxVideoTransformer3DModel(small_cfg(), device=d)
            dit.init_random(5)
            vae = cv.AutoencoderKLLtxVideo(cv.VaeConfig(decoder_layers_per_block=(1, 1, 1, 1)), device=d)
            vae.init_random(6)
        return dit, vae

    models = [make(0), make(1)]
    (height, width, frames), *_ = _inputs("cpu")
    params = cv.PipelineParams(height=height, width=width, num_frames=frames, frame_rate=25, num_inference_steps=3,
                               guidance_scale=3.0)

    def run(d, reps=1):
        dev = torch.device(f"cuda:{d}")
        dit, vae = models[d]
        res = None
        with torch.cuda.device(d):
            _, lat, pe, pm, ne, nm, z = _inputs(dev)
            for _ in range(reps):
                out = lat.clone()
                cv.pipeline_denoise(dit, params, out, pe, pm, ne, nm)
                video = cv.pipeline_decode(vae, params, out)
                torch.cuda.synchronize(d)
                cur = (out.cpu(), video.cpu())
                if res is not None:
                    assert torch.equal(res[0], cur[0]) and torch.equal(res[1], cur[1])
                res = cur
        return res

    # each device alone, then strictly alternating calls from one thread
    alone = [run(0), run(1)]
    assert torch.isfinite(alone[0][0]).all() and torch.isfinite(alone[0][1]).all()
    assert torch.equal(alone[0][0], alone[1][0]) and torch.equal(alone[0][1], alone[1][1])
    for _ in range(2):
        for d in (0, 1):
            got = run(d)
            assert torch.equal(got[0], alone[d][0]) and torch.equal(got[1], alone[d][1])

    # two host threads, one per device, running concurrently
    results, errors = [None, None], []

    def worker(d):
        try:
            results[d] = run(d, reps=4)
        except BaseException as e:  # noqa: BLE001 -- re-raised in the main thread
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(d,)) for d in (0, 1)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    for d in (0, 1):
        assert torch.equal(results[d][0], alone[d][0]) and torch.equal(results[d][1], alone[d][1])
