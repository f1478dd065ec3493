"""GPU parity of the pipeline glue through the C ABI: index maps and coordinate grids bit-exact, f32 element-wise
paths max-abs <= 1e-5 against the oracle."""
import pytest
import torch

from oracle import ltx_oracle as O
from tests.util import max_abs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,p,pt", [((1, 128, 13, 16, 24), 1, 1), ((2, 128, 4, 8, 12), 1, 1),
                                        ((1, 128, 1, 1, 1), 1, 1), ((1, 8, 4, 6, 10), 2, 2), ((2, 16, 3, 6, 9), 3, 1),
                                        ((1, 128, 7, 5, 3), 1, 1)])
def test_pack_unpack_bit_exact(cuda, shape, p, pt):
    import candle_video_b200 as cv
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(0))
    ref = O.pack_latents(x, p, pt)
    out = cv.pack_latents(x.to(cuda), p, pt)
    assert torch.equal(out.cpu(), ref)
    b, c, f, h, w = shape
    back = cv.unpack_latents(out, f // pt, h // p, w // p, p, pt)
    assert torch.equal(back.cpu(), x)  # round trip identity (tests/verify_pipeline_parity.rs:742-771)
    assert torch.equal(back.cpu(), O.unpack_latents(ref, f // pt, h // p, w // p, p, pt))


def test_pack_rejects_indivisible_shapes(cuda):
    import candle_video_b200 as cv
    with pytest.raises(cv.LtxvError, match="not divisible"):
        cv.pack_latents(torch.zeros(1, 4, 3, 4, 4, device=cuda), 2, 2)


@pytest.mark.parametrize("nf,hh,ww,fps", [(9, 256, 256, 25), (25, 512, 768, 25), (97, 512, 768, 25),
                                          (161, 512, 768, 25), (9, 512, 768, 30), (25, 768, 512, 24),
                                          (121, 704, 1216, 25), (257, 704, 1216, 25)])
def test_video_coords_bit_exact(cuda, nf, hh, ww, fps):
    import candle_video_b200 as cv
    F, H, W = (nf - 1) // 8 + 1, hh // 32, ww // 32
    ref = O.video_coords(2, F, H, W, fps)
    out = cv.video_coords(2, F, H, W, fps, device=cuda)
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("g,r,s,use_u,use_p", [(3.0, 0.0, 0.0, True, False), (3.0, 0.7, 0.0, True, False),
                                               (1.0, 0.0, 1.0, False, True), (3.0, 0.5, 1.0, True, True),
                                               (1.0, 0.0, 0.0, False, False)])
def test_guidance_euler_matches_oracle(cuda, g, r, s, use_u, use_p):
    import candle_video_b200 as cv
    gen = torch.Generator().manual_seed(5)
    B, S, C = 2, 4992, 128
    c, u, p = (torch.randn(B, S, C, generator=gen) for _ in range(3))
    lat = torch.randn(B, S, C, generator=gen)
    sigma, sigma_next = 0.9863, 0.9712
    comb = O.guidance_combine(c, u if use_u else None, p if use_p else None, g, r, s)
    ref = O.euler_step(lat, comb, sigma, sigma_next)
    lat_d = lat.to(cuda)
    noise = cv.guidance_euler_step(c.to(cuda), u.to(cuda) if use_u else None, p.to(cuda) if use_p else None, lat_d,
                                   g, r, s, sigma, sigma_next, return_noise_pred=True)
    tol = 1e-5 if r == 0.0 else 2e-5  # std ratio is a f32 reduction in the reference, f64 here
    assert max_abs(noise, comb) <= tol * max(1.0, float(comb.abs().max()))
    assert max_abs(lat_d, ref) <= 1e-5
    if r == 0.0:
        assert torch.equal(noise.cpu(), comb)  # same individually rounded f32 ops
        assert torch.equal(lat_d.cpu(), ref)


def test_denormalize_and_postprocess_bit_exact(cuda):
    import candle_video_b200 as cv
    gen = torch.Generator().manual_seed(6)
    x = torch.randn(2, 128, 3, 4, 5, generator=gen)
    mean, std = torch.randn(128, generator=gen), torch.rand(128, generator=gen) + 0.5
    for sf in (1.0, 0.13025):
        ref = O.denormalize_latents(x, mean, std, sf)
        out = cv.denormalize_latents(x.to(cuda), mean.to(cuda), std.to(cuda), sf)
        assert torch.equal(out.cpu(), ref)
    v = torch.randn(1, 3, 9, 64, 96, generator=gen) * 1.5
    assert torch.equal(cv.postprocess_video(v.to(cuda)).cpu(), O.postprocess_video(v))


def test_stochastic_step_and_decode_noise_bit_exact(cuda):
    """scheduler.rs:557-575 and t2v_pipeline.rs:1049-1062 with caller-supplied noise: f32, bit-exact vs the oracle."""
    import candle_video_b200 as cv
    g = torch.Generator().manual_seed(21)
    x, v, n = (torch.randn(4992, 128, generator=g) for _ in range(3))
    for sigma, sigma_next in [(1.0, 0.9375), (0.421, 0.25), (0.1, 0.0)]:
        ref = O.stochastic_step(x, v, n, sigma, sigma_next)
        out = cv.scheduler_step_stochastic(x.to(cuda).clone(), v.to(cuda), n.to(cuda), sigma, sigma_next)
        assert torch.equal(out.cpu(), ref)
    ref = O.decode_noise_blend(x, n, 0.025)
    out = x.to(cuda).clone()
    import ctypes as C
    cv._check(cv.lib().ltxv_decode_noise_blend(cv._ptr(out), cv._ptr(n.to(cuda)), 0.025, out.numel(), cv._stream()))
    assert torch.equal(out.cpu(), ref)


def test_frames_to_u8_bit_exact(cuda):
    """main.rs:653-667 hand-off: permute + clamp + truncating cast, bit-exact vs the oracle incl. out-of-range values."""
    import candle_video_b200 as cv
    g = torch.Generator().manual_seed(9)
    v = torch.rand(2, 3, 5, 6, 10, generator=g) * 300.0 - 20.0   # below 0 and above 255 on purpose
    v[0, 0, 0, 0, :4] = torch.tensor([0.0, 254.999, 255.0, 0.999])
    out = cv.frames_to_u8(v.to(cuda)).cpu()
    assert out.dtype == torch.uint8 and out.shape == (2, 5, 6, 10, 3)
    assert torch.equal(out, O.frames_to_u8(v))
    assert out[0, 0, 0, :4, 0].tolist() == [0, 254, 255, 0]
