"""GPU: the CUDA path (through the C ABI) against the COMMITTED golden fixture tests/golden/ltx_golden_v1.safetensors
(provenance: tests/golden/make_golden.py -- oracle-generated, see there).  Nothing under oracle/ computes here except
the seeded weight generators, which the fixture deliberately does not store."""
from pathlib import Path

import pytest
import torch
from safetensors.torch import load_file

from oracle import ltx_oracle as O
from tests.golden import make_golden as MG
from tests.util import psnr_255, rel_l2

pytestmark = pytest.mark.gpu

GOLD = Path(__file__).resolve().parent / "golden" / "ltx_golden_v1.safetensors"


@pytest.fixture(scope="module")
def gold():
    return load_file(str(GOLD))


def test_dit_forward_vs_golden(cuda, gold):
    from tests.test_gpu_dit import build
    cfg = O.DitConfig(**MG.DIT_CFG)
    m, _ = build(cfg, 42)
    s = MG.DIT_SHAPE
    out = m.forward(gold["dit.hidden"].to(cuda), gold["dit.enc"].to(cuda), gold["dit.timestep"].to(cuda),
                    gold["dit.mask"].to(cuda), s["F"], s["H"], s["W"], None, gold["dit.coords"].to(cuda)).cpu()
    e = rel_l2(out, gold["dit.out"])
    print(f"DiT vs golden: rel_l2={e:.3e}")
    assert out.shape == gold["dit.out"].shape and e <= 2e-2


def test_vae_decode_vs_golden(cuda, gold):
    from tests.test_gpu_vae import build
    m, _, _ = build(layers=MG.VAE_LAYERS, seed=7)
    out = m.decode(gold["vae.z"].to(cuda), gold["vae.timestep"].to(cuda)).cpu()
    sub = out[:, :, :, ::4, ::4]
    ps = psnr_255(sub, gold["vae.out_sub"])
    print(f"VAE decode vs golden: psnr={ps:.1f} dB rel_l2={rel_l2(sub, gold['vae.out_sub']):.3e}")
    assert ps >= 35.0 and rel_l2(sub, gold["vae.out_sub"]) <= 3e-2
    mean_std = torch.stack([out.mean(), out.std()])
    assert torch.allclose(mean_std, gold["vae.out_mean_std"], atol=5e-3)


def test_vae_encode_vs_golden(cuda, gold):
    import candle_video_b200 as cv
    ecfg = O.VaeEncoderConfig(**MG.ENC_CFG)
    m = cv.AutoencoderKLLtxVideo(cv.VaeConfig(decoder_layers_per_block=(1, 1, 1, 1)))
    m.enable_encoder(cv.VaeEncoderConfig(block_out_channels=ecfg.block_out_channels, layers_per_block=ecfg.layers_per_block))
    sd = dict(O.init_vae_weights(O.VaeConfig(decoder_layers_per_block=(1, 1, 1, 1)), 7))
    sd.update(O.init_vae_encoder_weights(ecfg, 11))
    m.load_state_dict(sd)
    out = m.encode(gold["enc.x"].to(cuda)).cpu()
    e = rel_l2(out, gold["enc.moments"])
    print(f"VAE encode vs golden: rel_l2={e:.3e}")
    assert out.shape == gold["enc.moments"].shape and e <= 3e-2


def test_guidance_euler_vs_golden(cuda, gold):
    import candle_video_b200 as cv
    lat = gold["glue.latents"].to(cuda).clone()
    noise = cv.guidance_euler_step(gold["glue.cond"].to(cuda), gold["glue.uncond"].to(cuda), gold["glue.perturbed"].to(cuda),
                                   lat, 3.0, 0.7, 1.0, 0.9, 0.8, return_noise_pred=True).cpu()
    # the std rescale is a reduction (summation order differs from torch): f32 tolerance as in the oracle test
    assert (noise - gold["glue.noise_pred"]).abs().max() <= 1e-5 * gold["glue.noise_pred"].abs().max()
    assert (lat.cpu() - gold["glue.latents_next"]).abs().max() <= 1e-5 * gold["glue.latents_next"].abs().max()


def test_schedule_vs_golden(gold):
    import candle_video_b200 as cv
    sig, ts = cv.scheduler_set_timesteps(40, cv.calculate_shift(4992))
    assert torch.equal(torch.tensor(sig, dtype=torch.float32), gold["sched.sigmas"])
    assert torch.equal(torch.tensor(ts, dtype=torch.float32), gold["sched.timesteps"])
