"""CPU: the algorithmic-work figures bench.py divides by agree with the survey's closed forms and with
what the oracle's weight shapes imply, so `roofline.achieved` rests on checked arithmetic."""
import importlib.util
from pathlib import Path

import pytest

from oracle import ltx_oracle as O

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_dit_flops_c2_matches_survey(bench):
    # c2 (S = 4992, K = 128, 2B) = 22.35 TFLOP per forward, 44.7 per CFG step
    assert bench.dit_flops(4992) / 1e12 == pytest.approx(22.35, abs=0.02)


def test_vae_decode_flops_c2_matches_survey_and_weight_shapes(bench):
    assert bench.vae_flops(13, 16, 24) / 1e12 == pytest.approx(48.40, abs=0.02)
    # independent count from the decoder's weight shapes and the level volumes (T' = 2T-1, H/W doubling per level)
    shapes = O.vae_weight_shapes(O.VaeConfig())
    vol = {}
    T, H, W = 13, 16, 24
    vol["conv_in"] = vol["mid_block"] = T * H * W
    for i in range(3):
        vol[f"up_blocks.{i}.upsamplers"] = T * H * W          # the upsampler conv runs on the level it leaves
        T, H, W = 2 * T - 1, 2 * H, 2 * W
        vol[f"up_blocks.{i}.resnets"] = T * H * W
    vol["conv_out"] = T * H * W
    tot = 0
    for k, shp in shapes.items():
        if not k.endswith("conv.weight") or len(shp) != 5:
            continue
        key = next(v for v in sorted(vol, key=len, reverse=True) if v in k)
        tot += 2 * shp[0] * shp[1] * 27 * vol[key]
    assert tot == bench.vae_flops(13, 16, 24)


def test_vae_encode_flops_match_weight_shapes(bench):
    cfg = O.VaeEncoderConfig()
    shapes = O.vae_encoder_weight_shapes(cfg)
    F, Hpx, Wpx = 121, 512, 768
    T, H, W = F, Hpx // 4, Wpx // 4
    tot = 2 * 48 * 128 * 27 * T * H * W
    for bi in range(4):
        c = cfg.block_out_channels[bi]
        tot += cfg.layers_per_block[bi] * 2 * 2 * c * c * 27 * T * H * W
        st, sh, sw = O.DOWNSAMPLE_STRIDE[cfg.downsample_types[bi]]
        cc = shapes[f"encoder.down_blocks.{bi}.downsamplers.0.conv.conv.weight"][0]
        tot += 2 * c * cc * 27 * (T + st - 1) * H * W
        T, H, W = (T + st - 1) // st, H // sh, W // sw
    c = cfg.block_out_channels[4]
    tot += (cfg.layers_per_block[4] - 1) * 2 * 2 * c * c * 27 * T * H * W
    tot += 2 * c * 129 * 27 * T * H * W
    assert (T, H, W) == (16, 16, 24)
    assert tot == bench.vae_encode_flops(F, Hpx, Wpx)
