"""GPU parity: ltxv_vae_encode (C ABI) vs the CPU f32 oracle restatement of LtxVideoEncoder3d /
AutoencoderKLLtxVideo::encode (vae.rs:496-582, :841-948, :1315-1469, :2017-2099) on identical synthetic weights.

The reference has no encoder golden vectors (its t2v path never encodes), so the bar is the decoder's: rel-L2 <= 3e-2 of
the bf16 device path against the f32 oracle, plus exact structural properties (logvar replication, causality)."""
import pytest
import torch

from oracle import ltx_oracle as O
from tests.util import max_abs, rel_l2

pytestmark = pytest.mark.gpu

_CACHE = {}


def build(layers=(1, 1, 1, 1, 2), seed=11):
    """Default 0.9.5 widths (128..2048) and downsample types; fewer resnets so the CPU oracle runs in seconds."""
    import candle_video_b200 as cv
    key = (layers, seed)
    if key not in _CACHE:
        ocfg = O.VaeEncoderConfig(layers_per_block=layers)
        w = O.init_vae_encoder_weights(ocfg, seed)
        m = cv.AutoencoderKLLtxVideo(cv.VaeConfig(decoder_layers_per_block=(1, 1, 1, 1)))
        m.enable_encoder(cv.VaeEncoderConfig(layers_per_block=layers))
        sd = dict(O.init_vae_weights(O.VaeConfig(decoder_layers_per_block=(1, 1, 1, 1)), 7))
        sd.update(w)
        m.load_state_dict(sd)
        _CACHE[key] = (m, w, ocfg)
    return _CACHE[key]


def video(B, F, H, W, seed=1):
    g = torch.Generator().manual_seed(seed)
    # smooth-ish content in [-1, 1] like a normalised clip
    return torch.tanh(torch.randn(B, 3, F, H, W, generator=g))


def check(out, ref, tag, L=128):
    e_mean, e_lv = rel_l2(out[:, :L], ref[:, :L]), rel_l2(out[:, L:], ref[:, L:])
    print(f"VAE encode {tag}: mean rel_l2={e_mean:.3e} logvar rel_l2={e_lv:.3e} max_abs={max_abs(out, ref):.3e} "
          f"ref_rms={ref.pow(2).mean().sqrt():.3e}")
    assert out.shape == ref.shape and torch.isfinite(out).all()
    assert e_mean <= 3e-2 and e_lv <= 3e-2
    # the log-variance half is ONE conv channel replicated (vae.rs:1462-1467): bit-identical copies
    assert torch.equal(out[:, L:], out[:, L:L + 1].expand_as(out[:, L:]))


@pytest.mark.parametrize("F,H,W", [(9, 64, 64), (17, 64, 96), (1, 96, 64)])
def test_vae_encode_matches_oracle(cuda, F, H, W):
    m, w, ocfg = build()
    x = video(1, F, H, W)
    ref = O.vae_encode(w, ocfg, x)
    out = m.encode(x.to(cuda)).cpu()
    assert out.shape == (1, 256, (F - 1) // 8 + 1, H // 32, W // 32)
    assert m.encode_dims(F, H, W) == tuple(out.shape[2:])
    check(out, ref, f"F{F} H{H} W{W}")


def test_vae_encode_batch2_bf16_input_two_resnets(cuda):
    m, w, ocfg = build(layers=(2, 1, 1, 1, 3), seed=12)
    x = video(2, 9, 64, 64, seed=2).bfloat16()
    ref = O.vae_encode(w, ocfg, x.float())
    out = m.encode(x.to(cuda)).cpu()
    check(out, ref, "batch2 bf16")
    # batch elements are independent (encode with use_slicing gives the same, vae.rs:2076-2086)
    assert torch.equal(m.encode(x[1:].to(cuda)).cpu(), out[1:])


def test_vae_encode_is_causal(cuda):
    """Every conv is causal and the downsamplers only look backwards: latent frame k depends on video frames
    <= 8k (vae.rs:383-387, :539-544).  Changing frames 9.. must leave latent frames 0 and 1 bit-identical."""
    m, _, _ = build()
    x = video(1, 17, 64, 64, seed=3)
    y = x.clone()
    y[:, :, 9:] = video(1, 8, 64, 64, seed=4)
    a, b = m.encode(x.to(cuda)).cpu(), m.encode(y.to(cuda)).cpu()
    assert torch.equal(a[:, :, :2], b[:, :, :2])
    assert not torch.equal(a[:, :, 2], b[:, :, 2])


def test_vae_encode_host_matches_device(cuda):
    m, _, _ = build()
    x = video(1, 9, 64, 64, seed=5)
    assert torch.equal(m.encode_host(x), m.encode(x.to(cuda)).cpu())


def test_vae_encode_rejects_bad_extents(cuda):
    import candle_video_b200 as cv
    m, _, _ = build()
    with pytest.raises(cv.LtxvError, match="does not divide|not divisible"):
        m.encode(torch.zeros(1, 3, 8, 64, 64, device=cuda))   # frames must be 8k+1
    with pytest.raises(cv.LtxvError, match="does not divide|not divisible"):
        m.encode(torch.zeros(1, 3, 9, 48, 64, device=cuda))   # height must be a multiple of 32
    plain = cv.AutoencoderKLLtxVideo(cv.VaeConfig(decoder_layers_per_block=(1, 1, 1, 1)))
    with pytest.raises(cv.LtxvError, match="no encoder"):
        plain.encode(torch.zeros(1, 3, 9, 64, 64, device=cuda))
    with pytest.raises(cv.LtxvError, match="unsupported downsample type"):
        cv.VaeEncoderConfig(downsample_types=("conv", "conv", "conv", "conv")).to_c()


def test_encoder_weights_are_required_once_enabled(cuda):
    import candle_video_b200 as cv
    m = cv.AutoencoderKLLtxVideo(cv.VaeConfig(decoder_layers_per_block=(1, 1, 1, 1)))
    dec = O.init_vae_weights(O.VaeConfig(decoder_layers_per_block=(1, 1, 1, 1)), 7)
    m.load_state_dict(dec)  # decoder-only deployment: fine
    m.enable_encoder(cv.VaeEncoderConfig(layers_per_block=(1, 1, 1, 1, 2)))
    with pytest.raises(cv.LtxvError, match="encoder tensors were never loaded"):
        m.load_state_dict(dec)


def test_normalize_latents_bit_exact_and_round_trip(cuda):
    import candle_video_b200 as cv
    g = torch.Generator().manual_seed(6)
    z = torch.randn(2, 128, 3, 4, 5, generator=g)
    mean, std = torch.randn(128, generator=g), torch.rand(128, generator=g) + 0.5
    for sf in (1.0, 0.7):
        out = cv.normalize_latents(z.to(cuda), mean.to(cuda), std.to(cuda), sf).cpu()
        assert torch.equal(out, O.normalize_latents(z, mean, std, sf))
        back = cv.denormalize_latents(out.to(cuda), mean.to(cuda), std.to(cuda), sf).cpu()
        assert (back - z).abs().max() <= 4e-6 * z.abs().max()


def test_encode_decode_shapes_compose(cuda):
    """encode -> mode -> normalize -> denormalize -> decode runs end to end and restores the video extent."""
    import candle_video_b200 as cv
    m, _, _ = build()
    x = video(1, 9, 64, 96, seed=8).to(cuda)
    mom = m.encode(x)
    z = mom[:, :128].contiguous()
    mean = torch.zeros(128, device=cuda)
    std = torch.ones(128, device=cuda)
    zn = cv.normalize_latents(z, mean, std, 1.0)
    assert torch.equal(zn, z)
    out = m.decode(cv.denormalize_latents(zn, mean, std, 1.0), torch.tensor([0.05], device=cuda))
    assert out.shape == x.shape and torch.isfinite(out).all()


# ---------------------------------------------------------------------------------------------------------------
# tiled / temporal-tiled encode (ltxv_vae_encode_tiled) vs the oracle restatement of encode_z (vae.rs:2017-2034,
# :2158-2223, :2294-2356), with small tiles so every seam type occurs on clips the CPU oracle encodes in seconds
# ---------------------------------------------------------------------------------------------------------------
SMALL_TILES = dict(tile_sample_min_height=64, tile_sample_min_width=64, tile_sample_stride_height=32,
                   tile_sample_stride_width=32, tile_sample_min_num_frames=16, tile_sample_stride_num_frames=8)


def test_vae_encode_spatial_tiles(cuda):
    import candle_video_b200 as cv
    m, w, ocfg = build()
    x = video(1, 9, 96, 128, seed=21)   # 3 x 4 tiles of 64 px at stride 32: vertical, horizontal and corner seams
    ref = O.vae_encode_z(w, ocfg, x, O.VaeTiling(**SMALL_TILES))
    out = m.encode_tiled(x.to(cuda), cv.VaeTiling(**SMALL_TILES)).cpu()
    check(out, ref, "spatial tiles 3x4")
    # the tiled result is NOT the untiled one: the compatibility mode is needed to reproduce the reference's default
    plain = O.vae_encode(w, ocfg, x)
    assert rel_l2(ref[:, :128], plain[:, :128]) > 10 * rel_l2(out[:, :128], ref[:, :128])


def test_vae_encode_temporal_tiles_batch2(cuda):
    import candle_video_b200 as cv
    m, w, ocfg = build()
    x = video(2, 33, 64, 64, seed=22)   # 5 temporal tiles (17,17,17,9,1 frames), no spatial tiling at 64 px
    ref = O.vae_encode_z(w, ocfg, x, O.VaeTiling(**SMALL_TILES), use_framewise_encoding=True)
    out = m.encode_tiled(x.to(cuda), cv.VaeTiling(**SMALL_TILES), use_framewise_encoding=True).cpu()
    assert out.shape == (2, 256, 5, 2, 2)
    check(out, ref, "temporal tiles, B=2")


def test_vae_encode_tiled_no_branch_is_plain(cuda):
    import candle_video_b200 as cv
    m, _, _ = build()
    x = video(1, 9, 64, 96, seed=23).to(cuda)
    # library defaults: 96 px <= 512 and framewise encoding off -> no branch (vae.rs:2019-2028)
    assert torch.equal(m.encode_tiled(x), m.encode(x))
    assert torch.equal(m.encode_tiled(x, cv.VaeTiling(use_tiling=False, **SMALL_TILES)), m.encode(x))
    with pytest.raises(cv.LtxvError, match="multiples of 32"):
        m.encode_tiled(x, cv.VaeTiling(tile_sample_min_height=64, tile_sample_min_width=64, tile_sample_stride_height=48,
                                       tile_sample_stride_width=32))
