"""Pin the CPU oracle against the reference's self-contained known-answer tests.

The reference's numeric fixtures (gen_*.safetensors) are git-ignored and absent, so these closed-form KATs are the
only in-tree pins: AdaLN value, attention scale, pack/unpack identity, video-coordinate closed form, CFG formula,
unbiased std, sinusoid layout, RoPE padding channels, depth-to-space / unpatchify index maps, preset constants,
Euler formula, scheduler closed forms.
"""
import math

import numpy as np
import pytest
import torch

from oracle import ltx_oracle as O


def test_adaln_known_answer():
    # tests/verify_rope_parity.rs:646-733: x[i] = 0.1 i, scale[i] = 0.01 i, shift[i] = 0.001 i; result[1] = 0.1*1.01+0.001
    x = torch.arange(8, dtype=torch.float32) * 0.1
    scale = torch.arange(8, dtype=torch.float32) * 0.01
    shift = torch.arange(8, dtype=torch.float32) * 0.001
    r = x * (1.0 + scale) + shift
    assert abs(float(r[1]) - (0.1 * 1.01 + 0.001)) < 1e-6


def test_attention_scale_known_answer():
    # tests/verify_rope_parity.rs:472-511: head_dim 64 -> scale 0.125
    cfg = O.dit_config_2b()
    assert cfg.attention_head_dim == 64
    assert abs(1.0 / math.sqrt(cfg.attention_head_dim) - 0.125) < 1e-7


def test_attention_uniform_scores_average_values():
    # with q = 0 every key gets the same weight: output = mean(v) (softmax semantic of :735)
    d, heads, S, K = 16, 2, 5, 7
    w = {}
    D = d * heads
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        w[f"a.{n}.weight"] = torch.eye(D)
        w[f"a.{n}.bias"] = torch.zeros(D)
    w["a.to_q.weight"] = torch.zeros(D, D)
    w["a.norm_q.weight"] = torch.ones(D)
    w["a.norm_k.weight"] = torch.ones(D)
    x = torch.randn(1, S, D)
    enc = torch.randn(1, K, D)
    out = O.attention(w, "a.", heads, x, enc, None, None)
    assert torch.allclose(out, enc.mean(1, keepdim=True).expand(1, S, D), atol=1e-5)


def test_mask_bias_removes_padded_keys():
    d, heads, S, K = 16, 2, 3, 6
    D = d * heads
    g = torch.Generator().manual_seed(0)
    w = {}
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        w[f"a.{n}.weight"] = torch.randn(D, D, generator=g) / math.sqrt(D)
        w[f"a.{n}.bias"] = torch.zeros(D)
    w["a.norm_q.weight"] = torch.ones(D)
    w["a.norm_k.weight"] = torch.ones(D)
    x = torch.randn(1, S, D, generator=g)
    enc = torch.randn(1, K, D, generator=g)
    mask = torch.tensor([[1., 1., 1., 1., 0., 0.]])
    bias = ((1 - mask) * -10000.0).unsqueeze(1)
    a = O.attention(w, "a.", heads, x, enc, bias, None)
    b = O.attention(w, "a.", heads, x, enc[:, :4], None, None)
    assert torch.allclose(a, b, atol=1e-5)


@pytest.mark.parametrize("shape,p,pt", [((2, 128, 3, 4, 6), 1, 1), ((1, 8, 4, 4, 6), 2, 2), ((1, 16, 2, 6, 9), 3, 1)])
def test_pack_unpack_round_trip_identity(shape, p, pt):
    # tests/verify_pipeline_parity.rs:742-771
    x = torch.randn(*shape)
    b, c, f, h, w = shape
    packed = O.pack_latents(x, p, pt)
    assert packed.shape == (b, (f // pt) * (h // p) * (w // p), c * pt * p * p)
    back = O.unpack_latents(packed, f // pt, h // p, w // p, p, pt)
    assert torch.equal(back, x)


def test_pack_index_map_p1():
    # p = pt = 1: out[b, (f*H+h)*W+w, c] = in[b,c,f,h,w]  
    x = torch.arange(2 * 3 * 2 * 3 * 4, dtype=torch.float32).reshape(2, 3, 2, 3, 4)
    p = O.pack_latents(x)
    for (b, c, f, h, w) in [(0, 0, 0, 0, 0), (1, 2, 1, 2, 3), (0, 1, 1, 0, 2)]:
        assert p[b, (f * 3 + h) * 4 + w, c] == x[b, c, f, h, w]


@pytest.mark.parametrize("nf,hh,ww,fps", [(9, 256, 256, 25), (25, 512, 768, 25), (97, 512, 768, 25),
                                          (161, 512, 768, 25), (9, 512, 768, 30), (25, 768, 512, 24)])
def test_video_coords_closed_form(nf, hh, ww, fps):
    # the configurations of tests/verify_video_coords_parity.rs:109-116 against the closed form (f32 arithmetic)
    F, H, W = (nf - 1) // 8 + 1, hh // 32, ww // 32
    c = O.video_coords(1, F, H, W, fps)
    assert c.shape == (1, F * H * W, 3)
    inv = np.float32(1.0 / fps)
    for (f, h, w) in [(0, 0, 0), (F - 1, H - 1, W - 1), (F // 2, H // 3, W // 2)]:
        s = (f * H + h) * W + w
        exp_f = np.float32(min(max(np.float32(f) * np.float32(8) + np.float32(-7), 0), 1000)) * inv
        assert c[0, s, 0].item() == float(exp_f)
        assert c[0, s, 1].item() == 32.0 * h
        assert c[0, s, 2].item() == 32.0 * w
    assert c[0, 0, 0].item() == 0.0  # causal fix: first latent frame clamps to t = 0


def test_cfg_formula_and_rescale():
    g = torch.Generator().manual_seed(1)
    u, c = torch.randn(2, 6, 8, generator=g), torch.randn(2, 6, 8, generator=g)
    out = O.guidance_combine(c, u, None, 3.0, 0.0, 0.0)
    assert torch.allclose(out, u + 3.0 * (c - u), atol=1e-6)
    # guidance_scale 1 -> cond  (tests/verify_cfg_parity.rs scale variations)
    assert torch.allclose(O.guidance_combine(c, u, None, 1.0, 0.0, 0.0), c, atol=1e-6)
    # rescale = 1 forces std(comb) == std(cond) per batch entry (unbiased std over all non-batch elements)
    r = O.guidance_combine(c, u, None, 3.0, 1.0, 0.0)
    assert torch.allclose(O.std_over_dims_except0(r), O.std_over_dims_except0(c), rtol=1e-4)
    # rescale = 0 is the identity (test_rescale_noise_cfg_zero)
    assert torch.equal(O.rescale_noise_cfg(out, c, 0.0), out * 0.0 + out)
    # STG term
    p = torch.randn(2, 6, 8, generator=g)
    assert torch.allclose(O.guidance_combine(c, None, p, 1.0, 0.0, 1.5), c + 1.5 * (c - p), atol=1e-6)


def test_std_is_unbiased():
    x = torch.tensor([[1.0, 2.0, 3.0, 4.0]])
    assert abs(float(O.std_over_dims_except0(x)) - math.sqrt(5.0 / 3.0)) < 1e-6


def test_sinusoid_layouts():
    # tests/verify_timestep_embedding.rs:55-61 values of t; layout [cos | sin], dim 256
    for t in (0.0, 0.05, 0.1, 0.5, 1.0):
        e = O.vae_timestep_embedding(torch.tensor([t * 1000.0]))
        d = O.dit_timestep_embedding(torch.tensor([t * 1000.0]))
        assert e.shape == d.shape == (1, 256)
        # frequency 0 is 1: first cos = cos(t), first sin = sin(t)
        assert abs(float(e[0, 0]) - math.cos(t * 1000.0)) < 1e-4
        assert abs(float(e[0, 128]) - math.sin(t * 1000.0)) < 1e-4
        # the two formulations (powf vs exp) agree
        assert torch.allclose(e, d, atol=2e-3)
    z = O.dit_timestep_embedding(torch.tensor([0.0]))
    assert torch.equal(z[0, :128], torch.ones(128)) and torch.equal(z[0, 128:], torch.zeros(128))


def test_rope_padding_channels_and_layout():
    # D = 2048: n = 341 frequencies, first D mod 6 = 2 channels are identity (ltx_transformer.rs:514-521)
    coords = O.video_coords(1, 2, 2, 3, 25)
    cos, sin = O.rope_cos_sin(coords, 2048)
    assert cos.shape == sin.shape == (1, 12, 2048)
    assert torch.equal(cos[..., :2], torch.ones(1, 12, 2)) and torch.equal(sin[..., :2], torch.zeros(1, 12, 2))
    # repeat_interleave(2): channel pairs share the angle
    assert torch.equal(cos[..., 2::2], cos[..., 3::2]) and torch.equal(sin[..., 2::2], sin[..., 3::2])
    # token 0: frame coord 0 -> g = 0 -> angle = freq_j * (2*0 - 1) = -freq_j; lowest frequency is pi/2 -> cos = 0, sin = -1
    assert abs(float(cos[0, 0, 2])) < 1e-6 and abs(float(sin[0, 0, 2]) + 1.0) < 1e-6
    # D = 4096 -> rem 4
    c2, s2 = O.rope_cos_sin(coords, 4096)
    assert torch.equal(c2[..., :4], torch.ones(1, 12, 4))


def test_apply_rotary_is_pairwise_rotation():
    x = torch.randn(1, 3, 8)
    ang = torch.randn(1, 3, 4)
    cos, sin = ang.cos().repeat_interleave(2, -1), ang.sin().repeat_interleave(2, -1)
    y = O.apply_rotary_emb(x, cos, sin)
    xr, xi = x[..., 0::2], x[..., 1::2]
    assert torch.allclose(y[..., 0::2], xr * ang.cos() - xi * ang.sin(), atol=1e-6)
    assert torch.allclose(y[..., 1::2], xi * ang.cos() + xr * ang.sin(), atol=1e-6)
    assert torch.allclose(y.norm(dim=-1), x.norm(dim=-1), atol=1e-5)  # rotations preserve the norm


def test_depth_to_space_and_unpatchify_index_maps():
    # R22: out[b,c,2t+i,2h+j,2w+k] = y[b, c*8+i*4+j*2+k, t,h,w]
    y = torch.arange(1 * 16 * 2 * 3 * 2, dtype=torch.float32).reshape(1, 16, 2, 3, 2)
    o = O.depth_to_space(y, 2, 2, 2)
    assert o.shape == (1, 2, 4, 6, 4)
    for (c, t, h, w, i, j, k) in [(0, 0, 0, 0, 0, 0, 0), (1, 1, 2, 1, 1, 0, 1), (1, 0, 1, 1, 0, 1, 1)]:
        assert o[0, c, 2 * t + i, 2 * h + j, 2 * w + k] == y[0, c * 8 + i * 4 + j * 2 + k, t, h, w]
    # R23: out[b,c,f,4h+j,4w+i] = x[b, c*16+i*4+j, f,h,w]  (W sub-pixel is the slower channel index)
    x = torch.arange(1 * 48 * 2 * 2 * 3, dtype=torch.float32).reshape(1, 48, 2, 2, 3)
    u = O.unpatchify(x, 4, 1)
    assert u.shape == (1, 3, 2, 8, 12)
    for (c, f, h, w, i, j) in [(0, 0, 0, 0, 0, 0), (2, 1, 1, 2, 3, 1), (1, 0, 1, 0, 2, 3)]:
        assert u[0, c, f, 4 * h + j, 4 * w + i] == x[0, c * 16 + i * 4 + j, f, h, w]


def test_causal_conv3d_padding_rules():
    # kt=3 non-causal: replicate 1 frame each side; causal: 2 copies of frame 0 on the left; H/W zero pad
    w = torch.zeros(1, 1, 3, 3, 3)
    w[0, 0, 0, 1, 1] = 1.0  # picks input frame t-1 (non-causal) / t-2 (causal)
    b = torch.zeros(1)
    x = torch.arange(4, dtype=torch.float32).reshape(1, 1, 4, 1, 1).expand(1, 1, 4, 2, 2).contiguous()
    y = O.causal_conv3d(x, w, b, is_causal=False)
    assert y[0, 0, :, 0, 0].tolist() == [0.0, 0.0, 1.0, 2.0]
    yc = O.causal_conv3d(x, w, b, is_causal=True)
    assert yc[0, 0, :, 0, 0].tolist() == [0.0, 0.0, 0.0, 1.0]
    w2 = torch.zeros(1, 1, 3, 3, 3)
    w2[0, 0, 1, 1, 0] = 1.0  # left neighbour in W: zero padded at w = 0
    y2 = O.causal_conv3d(x + 1, w2, b)
    assert (y2[0, 0, :, :, 0] == 0).all() and (y2[0, 0, :, :, 1] == x[0, 0, :, :, 0] + 1).all()


def test_preset_constants():
    # configs.rs:289-324 / :135-160
    c2, c13 = O.dit_config_2b(), O.dit_config_13b()
    assert (c2.num_layers, c2.num_attention_heads, c2.attention_head_dim, c2.cross_attention_dim) == (28, 32, 64, 2048)
    assert (c13.num_layers, c13.attention_head_dim, c13.cross_attention_dim) == (48, 128, 4096)
    v = O.VaeConfig()
    assert v.stage_channels() == [1024, 512, 256, 128]
    assert len(O.vae_weight_shapes(v)) == 2 + 4 * 4 + 3 * 2 + 20 * 5 + 2 + 4 + 2


def test_euler_and_schedule_closed_forms():
    x, v = torch.randn(1, 4, 8), torch.randn(1, 4, 8)
    assert torch.allclose(O.euler_step(x, v, 1.0, 0.75), x - 0.25 * v, atol=1e-6)
    # calculate_shift: linear through (256, 0.5) and (4096, 1.15)
    assert abs(O.calculate_shift(256) - 0.5) < 1e-6 and abs(O.calculate_shift(4096) - 1.15) < 1e-6
    assert abs(O.calculate_shift(4992) - 1.3017) < 1e-3
    sig, ts = O.scheduler_set_timesteps(40, O.calculate_shift(4992))
    assert len(sig) == 41 and len(ts) == 40 and sig[-1] == 0.0
    assert ts[:6] == [1000, 993, 986, 978, 971, 963]  # SURVEY.md Appendix C replay
    assert abs(sig[39] - 0.1) < 1e-6 and all(a > b for a, b in zip(sig, sig[1:]))
    # distilled sigmas: mu = 0 leaves them unchanged up to the terminal stretch
    s8 = [1.0, 0.9937, 0.9875, 0.9812, 0.975, 0.9094, 0.725, 0.4219]
    sig2, ts2 = O.scheduler_set_timesteps(8, 0.0, sigmas=s8, shift_terminal=None)
    assert np.allclose(sig2[:-1], s8, atol=1e-6) and ts2[0] == 1000


# ---- tiled decode restatement: blend weights and dispatch (vae.rs:1927-2066) ----
def test_blend_weights_and_extent():
    a = torch.ones(1, 1, 1, 1, 6)
    b = torch.zeros(1, 1, 1, 1, 5)
    out = O._blend(a, b, 4, 4)
    # b[x] = a[-4+x]*(1-x/4) + b[x]*(x/4): 1, .75, .5, .25 then b's own tail
    assert torch.equal(out.flatten(), torch.tensor([1.0, 0.75, 0.5, 0.25, 0.0]))
    # the blend is clipped to both extents (min(blend, a, b), :1931)
    out = O._blend(torch.full((1, 1, 2, 1, 1), 2.0), torch.zeros(1, 1, 3, 1, 1), 8, 2)
    assert torch.equal(out.flatten(), torch.tensor([2.0, 1.0, 0.0]))
    assert torch.equal(O._blend(a, b, 0, 4), b)


def test_tiling_dispatch_and_frame_count():
    cfg = O.VaeConfig(decoder_layers_per_block=(1, 1, 1, 1))
    w = O.init_vae_weights(cfg, 7)
    z = torch.randn(1, 128, 3, 2, 2, generator=torch.Generator().manual_seed(0))
    ts = torch.tensor([0.05])
    plain = O.vae_decode(w, cfg, z, ts)
    # defaults: 3 latent frames > 16/8 = 2 -> temporal tiling; result keeps (F-1)*8+1 frames (:2433)
    tiled = O.vae_decode_z(w, cfg, z, ts, O.VaeTiling())
    assert tiled.shape == plain.shape == (1, 3, 17, 64, 64)
    # both switches off, or a volume below every threshold: the plain decoder (:2055-2065)
    assert torch.equal(O.vae_decode_z(w, cfg, z, ts, O.VaeTiling(use_tiling=False, use_framewise_decoding=False)), plain)
    assert torch.equal(O.vae_decode_z(w, cfg, z[:, :, :2], ts, O.VaeTiling()), O.vae_decode(w, cfg, z[:, :, :2], ts))
    # the first temporal tile contributes its first stride+1 = 9 frames unchanged (:2426-2429)
    first = O.vae_decode(w, cfg, z[:, :, :3], ts)
    assert torch.equal(tiled[:, :, :9], first[:, :, :9])


# ---------------------------------------------------------------------------------------------------------------
# encoder restatement
# ---------------------------------------------------------------------------------------------------------------
def test_patchify_index_map_and_unpatchify_inverse():
    """vae.rs:1427-1445: out[b, (c*pt + it)*p*p + iw*p + ih, f, h, w] = x[b, c, f*pt + it, h*p + ih, w*p + iw];
    the decoder's unpatchify (vae.rs:1626-1654) is its inverse."""
    x = torch.arange(2 * 3 * 2 * 8 * 12, dtype=torch.float32).reshape(2, 3, 2, 8, 12)
    y = O.patchify(x, 4, 1)
    assert y.shape == (2, 48, 2, 2, 3)
    for (b, c, f, h, w, ih, iw) in [(0, 0, 0, 0, 0, 0, 0), (1, 2, 1, 1, 2, 3, 1), (0, 1, 1, 0, 1, 2, 3)]:
        assert y[b, c * 16 + iw * 4 + ih, f, h, w] == x[b, c, f, h * 4 + ih, w * 4 + iw]
    assert torch.equal(O.unpatchify(y, 4, 1), x)
    with pytest.raises(ValueError):
        O.patchify(torch.zeros(1, 3, 1, 6, 8), 4, 1)


def test_space_to_depth_closed_form_and_inverse():
    x = torch.randn(1, 3, 4, 6, 8, generator=torch.Generator().manual_seed(0))
    for st, sh, sw in O.DOWNSAMPLE_STRIDE.values():
        y = O.space_to_depth(x, st, sh, sw)
        assert y.shape == (1, 3 * st * sh * sw, 4 // st, 6 // sh, 8 // sw)
        c, i, j, k, t, h, w = 2, st - 1, sh - 1, 0, 1, 2, 3
        assert y[0, ((c * st + i) * sh + j) * sw + k, t, h, w] == x[0, c, t * st + i, h * sh + j, w * sw + k]
        assert torch.equal(O.depth_to_space(y, st, sh, sw), x)


def test_downsampler_constant_input():
    """With zero conv weights the downsampler is the group-mean of the unshuffled input (+ bias): a constant volume
    stays constant, and the temporal variants turn T frames into (T+1)/2 (first frame duplicated, vae.rs:539-544)."""
    for name, (st, sh, sw) in O.DOWNSAMPLE_STRIDE.items():
        cin, cout = 8, 16
        cc = cout // (st * sh * sw)
        w = {"d.conv.conv.weight": torch.zeros(cc, cin, 3, 3, 3), "d.conv.conv.bias": torch.full((cc,), 0.25)}
        x = torch.full((1, cin, 5, 4, 4), 2.0)
        y = O.downsampler(w, "d.", x, (st, sh, sw), cout)
        assert y.shape == (1, cout, (5 + st - 1) // st, 4 // sh, 4 // sw), name
        assert torch.allclose(y, torch.full_like(y, 2.25))


def test_encoder_shapes_keys_and_logvar_replication():
    cfg = O.VaeEncoderConfig(block_out_channels=(64, 64, 128, 128, 256), layers_per_block=(1, 1, 1, 1, 2))
    shapes = O.vae_encoder_weight_shapes(cfg)
    # key names as consumed by VarBuilder (vae.rs:1342-1412, :863-905, :513-524)
    assert shapes["encoder.conv_in.conv.weight"] == (64, 48, 3, 3, 3)
    assert shapes["encoder.down_blocks.0.downsamplers.0.conv.conv.weight"] == (16, 64, 3, 3, 3)   # spatial: 64/4
    assert shapes["encoder.down_blocks.1.downsamplers.0.conv.conv.weight"] == (64, 64, 3, 3, 3)   # temporal: 128/2
    assert shapes["encoder.down_blocks.3.downsamplers.0.conv.conv.weight"] == (32, 128, 3, 3, 3)  # 256/8
    assert shapes["encoder.conv_out.conv.weight"] == (129, 256, 3, 3, 3)
    assert "encoder.mid_block.resnets.0.conv1.conv.weight" in shapes
    assert "encoder.mid_block.resnets.1.conv1.conv.weight" not in shapes  # layers - 1 resnets (vae.rs:1383-1386)
    full = O.vae_encoder_weight_shapes(O.VaeEncoderConfig())
    assert sum(1 for k in full if k.endswith("conv1.conv.weight")) == 4 + 6 + 6 + 2 + 1
    w = O.init_vae_encoder_weights(cfg)
    x = torch.randn(2, 3, 9, 64, 96, generator=torch.Generator().manual_seed(1))
    m = O.vae_encode(w, cfg, x)
    assert m.shape == (2, 256, 2, 2, 3)
    assert torch.equal(m[:, 128:], m[:, 128:129].expand(-1, 128, -1, -1, -1))
    # causal: latent frame 0 only sees video frame 0
    y = x.clone()
    y[:, :, 1:] = 0
    assert torch.allclose(O.vae_encode(w, cfg, y)[:, :, 0], m[:, :, 0], atol=1e-6)


def test_normalize_denormalize_round_trip():
    g = torch.Generator().manual_seed(3)
    z = torch.randn(1, 8, 2, 3, 4, generator=g)
    mean, std = torch.randn(8, generator=g), torch.rand(8, generator=g) + 0.5
    n = O.normalize_latents(z, mean, std, 0.5)
    assert torch.allclose(n[0, 3], (z[0, 3] - mean[3]) * 0.5 / std[3])
    assert torch.allclose(O.denormalize_latents(n, mean, std, 0.5), z, atol=1e-5)


def test_frames_to_u8_kat():
    """main.rs:653-667: channel-last per frame, clamp to [0,255], truncation toward zero."""
    v = torch.zeros(1, 3, 2, 1, 2)
    v[0, :, 1, 0, 1] = torch.tensor([12.9, 300.0, -4.0])
    out = O.frames_to_u8(v)
    assert out.shape == (1, 2, 1, 2, 3) and out.dtype == torch.uint8
    assert out[0, 1, 0, 1].tolist() == [12, 255, 0]


def test_tiled_encode_dispatch_and_zero_blend_tiles():
    """encode_z (vae.rs:2017-2034): no branch -> the plain encoder; with min == stride the blends vanish (extent 0) and
    every latent tile is exactly the encoder applied to its own crop (vae.rs:2175-2189, :2211-2216)."""
    cfg = O.VaeEncoderConfig(block_out_channels=(64, 64, 128, 128, 256), layers_per_block=(1, 1, 1, 1, 2))
    w = O.init_vae_encoder_weights(cfg)
    x = torch.tanh(torch.randn(1, 3, 9, 64, 128, generator=torch.Generator().manual_seed(5)))
    assert torch.equal(O.vae_encode_z(w, cfg, x, None), O.vae_encode(w, cfg, x))
    assert torch.equal(O.vae_encode_z(w, cfg, x, O.VaeTiling()), O.vae_encode(w, cfg, x))  # 128 px <= 512: no tiling
    tp = O.VaeTiling(tile_sample_min_height=64, tile_sample_min_width=64, tile_sample_stride_height=64,
                     tile_sample_stride_width=64)
    out = O.vae_encode_z(w, cfg, x, tp)
    assert out.shape == (1, 256, 2, 2, 4)
    right = O.vae_encode(w, cfg, x[..., 64:128])
    assert torch.equal(out[..., 2:4], right)
    # framewise: 33 frames -> 5 latent frames whatever the tile bookkeeping does (first tile drops a frame, :2324-2329)
    x2 = torch.tanh(torch.randn(1, 3, 33, 32, 32, generator=torch.Generator().manual_seed(6)))
    o2 = O.vae_encode_z(w, cfg, x2, O.VaeTiling(), use_framewise_encoding=True)
    assert o2.shape == (1, 256, 5, 1, 1) and torch.isfinite(o2).all()
