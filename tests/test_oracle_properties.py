"""CPU: property tests in the style of the reference's proptests (tests/verify_vae_property_tests.rs:44-320,
tests/verify_scheduler_parity.rs:640-860), run on the oracle with `hypothesis`: the same shape ranges, the same bars."""
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import ltx_oracle as O

SHAPES = dict(batch=st.integers(1, 2), frames=st.integers(1, 4), height=st.integers(4, 15), width=st.integers(4, 15),
              seed=st.integers(0, 999))


def _case(batch, frames, height, width, seed, channels=16):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, channels, frames, height, width, generator=g)
    mean = torch.randn(channels, generator=g)
    std = torch.rand(channels, generator=g) * 1.5 + 0.25
    sf = float(torch.rand(1, generator=g)) * 1.5 + 0.5
    return x, mean, std, sf


@settings(max_examples=100, deadline=None)
@given(**SHAPES)
def test_prop_latent_normalization_roundtrip(batch, frames, height, width, seed):
    """Property 7 (verify_vae_property_tests.rs:47-115): normalize then denormalize, MSE < 1e-10."""
    x, mean, std, sf = _case(batch, frames, height, width, seed)
    back = O.denormalize_latents(O.normalize_latents(x, mean, std, sf), mean, std, sf)
    assert float(((back - x) ** 2).mean()) < 1e-10


@settings(max_examples=100, deadline=None)
@given(**SHAPES)
def test_prop_normalization_and_denormalization_formulas(batch, frames, height, width, seed):
    """Properties at :197-320: (x - mean) * sf / std and x * std / sf + mean, element by element."""
    x, mean, std, sf = _case(batch, frames, height, width, seed)
    n, d = O.normalize_latents(x, mean, std, sf), O.denormalize_latents(x, mean, std, sf)
    b, c, f, h, w = [int(torch.randint(0, s, (1,), generator=torch.Generator().manual_seed(seed))) for s in x.shape]
    assert abs(float(n[b, c, f, h, w]) - (float(x[b, c, f, h, w]) - float(mean[c])) * sf / float(std[c])) < 1e-4
    assert abs(float(d[b, c, f, h, w]) - (float(x[b, c, f, h, w]) * float(std[c]) / sf + float(mean[c]))) < 1e-4


@settings(max_examples=100, deadline=None)
@given(seed=st.integers(0, 999), i=st.integers(0, 38))
def test_prop_euler_step_is_sample_plus_dt_times_velocity(seed, i):
    """verify_scheduler_parity.rs:370-410 / :840-855: prev = sample + (sigma_next - sigma) * model_output on the
    pipeline's own 40-step schedule; sigmas strictly decrease to the appended terminal 0."""
    sig, ts = O.scheduler_set_timesteps(40, O.calculate_shift(4992))
    assert len(sig) == 41 and sig[-1] == 0.0 and all(a > b for a, b in zip(sig, sig[1:]))
    assert all(int(t) == int(s * 1000.0) or abs(t - s * 1000.0) < 1.0 for t, s in zip(ts, sig))
    g = torch.Generator().manual_seed(seed)
    x, v = torch.randn(2, 24, 128, generator=g), torch.randn(2, 24, 128, generator=g)
    out = O.euler_step(x, v, sig[i], sig[i + 1])
    assert torch.allclose(out, x + (sig[i + 1] - sig[i]) * v, atol=1e-6)


@settings(max_examples=50, deadline=None)
@given(seed=st.integers(0, 999), g_scale=st.floats(1.0, 8.0), rescale=st.floats(0.0, 1.0))
def test_prop_cfg_rescale_bounds(seed, g_scale, rescale):
    """rescale_noise_cfg (t2v_pipeline.rs:227-243): rescale = 0 is plain CFG; rescale = 1 gives the guided prediction
    the per-sample std of the conditional one."""
    g = torch.Generator().manual_seed(seed)
    c, u = torch.randn(2, 32, 128, generator=g), torch.randn(2, 32, 128, generator=g)
    plain = O.guidance_combine(c, u, None, g_scale, 0.0, 0.0)
    assert torch.allclose(plain, u + (c - u) * g_scale, atol=1e-6)
    full = O.guidance_combine(c, u, None, g_scale, 1.0, 0.0)
    assert torch.allclose(O.std_over_dims_except0(full), O.std_over_dims_except0(c), rtol=1e-4, atol=1e-5)
    mix = O.guidance_combine(c, u, None, g_scale, rescale, 0.0)
    assert torch.allclose(mix, rescale * full + (1.0 - rescale) * plain, atol=1e-4)
