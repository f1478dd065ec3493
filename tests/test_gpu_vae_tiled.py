"""GPU: tiled / temporal-tiled VAE decode (ltxv_vae_decode_tiled) against the oracle restatement of the reference's
decode_z dispatch, tiled_decode, temporal_tiled_decode and blend_{h,v,t} (vae.rs:1927-2066, :2225-2290, :2358-2434),
with small tiles so that every seam type occurs on volumes the CPU oracle decodes in seconds."""
import pytest
import torch

from oracle import ltx_oracle as O
from tests.util import mse, psnr_255, rel_l2

pytestmark = pytest.mark.gpu

SMALL = dict(tile_sample_min_height=128, tile_sample_min_width=128, tile_sample_stride_height=96,
             tile_sample_stride_width=96, tile_sample_min_num_frames=16, tile_sample_stride_num_frames=8)


def _both(tp_kwargs, shape, cuda, t=0.05):
    import candle_video_b200 as cv
    from tests.test_gpu_vae import build
    m, w, cfg = build()
    g = torch.Generator().manual_seed(5)
    z = torch.randn(*shape, generator=g)
    ts = torch.tensor([t] * shape[0])
    ref = O.vae_decode_z(w, cfg, z, ts, O.VaeTiling(**tp_kwargs))
    out = m.decode_tiled(z.to(cuda), ts.to(cuda), cv.VaeTiling(**tp_kwargs)).cpu()
    return m, w, cfg, z, ts, ref, out


def _check(out, ref, tag):
    e, ms, ps = rel_l2(out, ref), mse(out, ref), psnr_255(out, ref)
    print(f"tiled VAE {tag}: rel_l2={e:.3e} mse={ms:.3e} psnr={ps:.1f} dB")
    assert out.shape == ref.shape and torch.isfinite(out).all()
    assert ms <= 1e-2 and ps >= 35.0 and e <= 3e-2


def test_spatial_tiles_2x2(cuda):
    # latent 5x6 with 4/3 tiles: rows {0..4, 3..5}, cols {0..4, 3..6}: vertical, horizontal and corner seams
    _, _, _, _, _, ref, out = _both(dict(SMALL, use_framewise_decoding=False), (1, 128, 2, 5, 6), cuda)
    _check(out, ref, "spatial 2x2")


def test_temporal_and_spatial_tiles(cuda):
    # 4 latent frames, tile_latent_min_t = 2, stride 1: 4 temporal tiles (3,3,2,1 frames), each 2x2 spatial
    m, w, cfg, z, ts, ref, out = _both(SMALL, (1, 128, 4, 5, 6), cuda)
    assert ref.shape == (1, 3, 25, 160, 192)
    _check(out, ref, "temporal x spatial")
    # the tiled result is NOT the untiled one (seams are blends of differently padded decodes): compat mode is needed
    plain = O.vae_decode(w, cfg, z, ts)
    assert rel_l2(ref, plain) > 10 * rel_l2(out, ref)


def test_temporal_only_batch2(cuda):
    tp = dict(SMALL, tile_sample_min_height=512, tile_sample_min_width=512, tile_sample_stride_height=384,
              tile_sample_stride_width=384)
    _, _, _, _, _, ref, out = _both(tp, (2, 128, 3, 3, 4), cuda)
    _check(out, ref, "temporal only, B=2")


def test_no_branch_is_plain_decode(cuda):
    import candle_video_b200 as cv
    from tests.test_gpu_vae import build
    m, _, _ = build()
    z = torch.randn(1, 128, 2, 4, 4, generator=torch.Generator().manual_seed(2)).to(cuda)
    ts = torch.tensor([0.05], device=cuda)
    # library defaults: 2 latent frames <= 16/8 and 4 <= 512/32 -> no tiling branch (vae.rs:2055-2065)
    assert torch.equal(m.decode_tiled(z, ts), m.decode(z, ts))
    assert torch.equal(m.decode_tiled(z, ts, cv.VaeTiling(use_tiling=False, use_framewise_decoding=False, **SMALL)),
                       m.decode(z, ts))
    with pytest.raises(cv.LtxvError, match="at least"):
        m.decode_tiled(z, ts, cv.VaeTiling(tile_sample_stride_height=16))
