"""GPU parity: ltxv_vae_decode (C ABI) vs the CPU f32 oracle (vae.rs decoder restated) on identical synthetic weights.

Shapes follow the reference's tests/verify_vae_decode_parity.rs (latents [1,128,2,4,4] and [1,128,2,8,8], temb 0.0 /
0.05); bars: MSE <= 1e-2 on [-1,1] (verify_vae_decode_parity.rs:74) and PSNR >= 35 dB on 0..255
(docs/benchmark_results.md:104).
"""
import pytest
import torch

from oracle import ltx_oracle as O
from tests.util import max_abs, mse, psnr_255, rel_l2

pytestmark = pytest.mark.gpu


def build(layers=(1, 1, 1, 1), seed=7, cond=True):
    import candle_video_b200 as cv
    cfg = O.VaeConfig(decoder_layers_per_block=layers, timestep_conditioning=cond)
    w = O.init_vae_weights(cfg, seed)
    m = cv.AutoencoderKLLtxVideo(cv.VaeConfig(decoder_layers_per_block=layers, timestep_conditioning=cond))
    m.load_state_dict(w)
    return m, w, cfg


def check(out, ref, tag):
    e, ms, ps = rel_l2(out, ref), mse(out, ref), psnr_255(out, ref)
    print(f"VAE {tag}: rel_l2={e:.3e} mse={ms:.3e} psnr={ps:.1f} dB max_abs={max_abs(out, ref):.3e} "
          f"ref_rms={ref.pow(2).mean().sqrt():.3e}")
    assert torch.isfinite(out).all()
    assert ms <= 1e-2
    assert ps >= 35.0
    assert e <= 3e-2


@pytest.mark.parametrize("F,H,W,t", [(2, 4, 4, 0.05), (2, 8, 8, 0.0), (3, 5, 7, 0.05)])
def test_vae_decode_matches_oracle(cuda, F, H, W, t):
    m, w, cfg = build()
    g = torch.Generator().manual_seed(1)
    z = torch.randn(1, 128, F, H, W, generator=g)
    ts = torch.tensor([t])
    ref = O.vae_decode(w, cfg, z, ts)
    out = m.decode(z.to(cuda), ts.to(cuda))
    assert out.shape == ref.shape == (1, 3, 8 * F - 7, 32 * H, 32 * W)
    check(out, ref, f"F{F}H{H}W{W} t={t}")


def test_vae_decode_two_resnets_batch2_bf16_latents(cuda):
    m, w, cfg = build(layers=(2, 1, 1, 2))
    g = torch.Generator().manual_seed(2)
    z = torch.randn(2, 128, 2, 4, 6, generator=g).bfloat16()
    ts = torch.tensor([0.05, 0.025])
    ref = O.vae_decode(w, cfg, z.float(), ts)
    out = m.decode(z.to(cuda), ts.to(cuda))
    check(out, ref, "batch2")


def test_vae_decode_without_timestep(cuda):
    m, w, cfg = build()
    z = torch.randn(1, 128, 2, 4, 4, generator=torch.Generator().manual_seed(3))
    ref = O.vae_decode(w, cfg, z, None)
    out = m.decode(z.to(cuda), None)
    check(out, ref, "no-temb")


def test_vae_postprocess_fused_and_host_entry(cuda):
    m, w, cfg = build()
    z = torch.randn(1, 128, 2, 4, 4, generator=torch.Generator().manual_seed(4))
    ts = torch.tensor([0.05])
    raw = m.decode(z.to(cuda), ts.to(cuda))
    post = m.decode(z.to(cuda), ts.to(cuda), postprocess=True)
    assert torch.equal(post.cpu(), O.postprocess_video(raw.cpu()))
    host = m.decode_host(z, ts)
    assert torch.equal(host, raw.cpu())


def test_vae_error_paths(cuda):
    import candle_video_b200 as cv
    m = cv.AutoencoderKLLtxVideo(cv.VaeConfig(decoder_layers_per_block=(1, 1, 1, 1)))
    with pytest.raises(cv.LtxvError, match="never loaded"):
        m.decode(torch.zeros(1, 128, 1, 2, 2, device=cuda), None)
    # encoder.* keys are accepted and ignored (the reference builds an encoder t2v never runs)
    m.__class__.load_state_dict  # noqa: B018
    import ctypes as C
    shape = (C.c_int64 * 1)(4)
    t = torch.zeros(4)
    assert cv.lib().ltxv_vae_load_tensor(m._h, b"encoder.conv_in.conv.bias", t.data_ptr(), 0, shape, 1) == 0
    assert cv.lib().ltxv_vae_load_tensor(m._h, b"decoder.bogus", t.data_ptr(), 0, shape, 1) != 0
