"""Oracle PINNING against candle-video itself -- the loader half of the hand-off in integration/b200_parity.rs.

The reference is Rust + Candle and cannot be built in this image (no cargo/rustc, crates not vendored), so its outputs
do not exist here.  `integration/b200_parity.rs` is a cargo test for the reference crate that rebuilds the tensors of
oracle/pin_weights.py bit for bit (portable splitmix64 generator, values exact in bf16), runs candle-video's own f32
CPU forward / decode and writes `b200_parity_dump.safetensors`.  When that file is present -- committed as
tests/golden/b200_parity_dump.safetensors or named by LTXV_REFERENCE_DUMP -- the tests below compare

  * the CPU oracle with the reference's output at the reference's own f32 bars
    (DiT max-abs < 2e-3, tests/verify_dit_parity.rs:97; VAE MSE < 1e-3, docs/benchmark_results.md:103), and
  * (GPU box) the CUDA path with the reference's output at the bf16 tolerances of SURVEY.md 8(c).

Without the dump they skip: the oracle then stays "parity unpinned" (its header says so).  The generator itself is
checked here unconditionally against the known-answer values the Rust twin asserts.
"""
import os
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import ltx_oracle as O
from oracle import pin_weights as P
from tests.util import max_abs, mse, psnr_255, rel_l2

ROOT = Path(__file__).resolve().parent.parent


def _dump():
    p = Path(os.environ.get("LTXV_REFERENCE_DUMP", ROOT / "tests" / "golden" / "b200_parity_dump.safetensors"))
    if not p.exists():
        pytest.skip(f"{p} not found: run integration/b200_parity.rs inside the reference crate to produce it "
                    "(the oracle stays 'parity unpinned' until then)")
    from safetensors.torch import load_file
    return load_file(str(p))


def test_generator_known_answers():
    """The constants asserted by `generator_known_answers` in integration/b200_parity.rs."""
    assert P.fnv1a64("proj_in.weight") == 12704714928704356426
    assert [int(v) for v in P.splitmix64(1, np.arange(3, dtype=np.uint64))] == [
        6238072747940578789, 10451216379200822465, 13757245211066428519]
    t = P.pin_tensor("proj_in.weight", (256, 128))
    assert t.flatten()[:4].tolist() == [-0.02197265625, -0.05908203125, 0.00146484375, -0.06103515625]
    assert P._pow2_bound(128) == 2.0 ** -4 and P._pow2_bound(256) == 2.0 ** -4 and P._pow2_bound(257) == 2.0 ** -5
    # every generated value survives a bf16 round trip (asserted inside pin_tensor), norm weights stay near 1
    w = P.pin_tensor("transformer_blocks.0.attn1.norm_q.weight", (256,))
    assert float(w.min()) >= 0.75 and float(w.max()) <= 1.25
    assert float(P.pin_tensor("decoder.timestep_scale_multiplier", ())) == 1000.0


def test_rust_twin_lists_the_oracle_keys():
    """integration/b200_parity.rs assembles the weight keys by hand: the distinctive tail of every key the oracle needs
    (e.g. `to_q.weight`, `linear_1.bias`, `scale_shift_table`) must appear in it."""
    src = (ROOT / "integration" / "b200_parity.rs").read_text()
    _, w, _ = P.pin_dit()
    _, vw, _ = P.pin_vae()
    for k in list(w) + list(vw):
        parts = k.split(".")
        tail = ".".join(parts[-2:]) if parts[-1] in ("weight", "bias") else parts[-1]
        assert tail in src, k


def test_oracle_matches_reference_dump():
    d = _dump()
    cfg, w, inp = P.pin_dit()
    for k in ("hidden_states", "encoder_hidden_states", "timestep", "encoder_attention_mask", "video_coords"):
        assert torch.equal(d["dit." + k].float().reshape(inp[k].shape), inp[k]), f"input {k} differs from the Rust twin"
    F, H, W = P.PIN_DIT_GRID
    out = O.dit_forward(w, cfg, inp["hidden_states"], inp["encoder_hidden_states"], inp["timestep"],
                        inp["encoder_attention_mask"], F, H, W, None, inp["video_coords"])
    ref = d["dit.output"].float()
    print(f"oracle vs candle-video DiT: max_abs={max_abs(out, ref):.3e} rel_l2={rel_l2(out, ref):.3e}")
    assert max_abs(out, ref) < 2e-3  # the reference's own bar against diffusers (tests/verify_dit_parity.rs:97)
    vcfg, vw, vinp = P.pin_vae()
    assert torch.equal(d["vae.latents"].float(), vinp["latents"])
    vout = O.vae_decode(vw, vcfg, vinp["latents"], vinp["temb"])
    vref = d["vae.output"].float()
    print(f"oracle vs candle-video VAE: mse={mse(vout, vref):.3e} psnr={psnr_255(vout, vref):.1f} dB")
    assert mse(vout, vref) < 1e-3  # docs/benchmark_results.md:103


@pytest.mark.gpu
def test_cuda_path_matches_reference_dump(cuda):
    import candle_video_b200 as cv
    d = _dump()
    cfg, w, inp = P.pin_dit()
    m = cv.LtxVideoTransformer3DModel(cv.DitConfig(**P.PIN_DIT, timestep_bf16_round=True))
    m.load_state_dict(w)
    F, H, W = P.PIN_DIT_GRID
    dev = lambda t: t.to(cuda)  # noqa: E731
    out = m.forward(dev(inp["hidden_states"]), dev(inp["encoder_hidden_states"]), dev(inp["timestep"]),
                    dev(inp["encoder_attention_mask"]), F, H, W, None, dev(inp["video_coords"]))
    assert rel_l2(out, d["dit.output"].float()) <= 2e-2
    vcfg, vw, vinp = P.pin_vae()
    v = cv.AutoencoderKLLtxVideo(cv.VaeConfig(decoder_layers_per_block=P.PIN_VAE_LAYERS))
    v.load_state_dict(vw)
    vout = v.decode(dev(vinp["latents"]), dev(vinp["temb"]))
    vref = d["vae.output"].float()
    assert mse(vout, vref) <= 1e-2 and psnr_255(vout, vref) >= 35.0


@pytest.mark.gpu
def test_cuda_path_matches_oracle_on_pin_tensors(cuda):
    """Runs with or without the dump: the same portable tensors through the CUDA path vs the oracle, so that once the
    dump pins the oracle the CUDA path is pinned on exactly these numbers too."""
    import candle_video_b200 as cv
    cfg, w, inp = P.pin_dit()
    m = cv.LtxVideoTransformer3DModel(cv.DitConfig(**P.PIN_DIT, timestep_bf16_round=True))
    m.load_state_dict(w)
    F, H, W = P.PIN_DIT_GRID
    dev = lambda t: t.to(cuda)  # noqa: E731
    ref = O.dit_forward(w, cfg, inp["hidden_states"], inp["encoder_hidden_states"], inp["timestep"],
                        inp["encoder_attention_mask"], F, H, W, None, inp["video_coords"], timestep_to_bf16=True)
    out = m.forward(dev(inp["hidden_states"]), dev(inp["encoder_hidden_states"]), dev(inp["timestep"]),
                    dev(inp["encoder_attention_mask"]), F, H, W, None, dev(inp["video_coords"]))
    print(f"pin DiT: rel_l2={rel_l2(out, ref):.3e}")
    assert rel_l2(out, ref) <= 2e-2
    vcfg, vw, vinp = P.pin_vae()
    v = cv.AutoencoderKLLtxVideo(cv.VaeConfig(decoder_layers_per_block=P.PIN_VAE_LAYERS))
    v.load_state_dict(vw)
    vref = O.vae_decode(vw, vcfg, vinp["latents"], vinp["temb"])
    vout = v.decode(dev(vinp["latents"]), dev(vinp["temb"]))
    print(f"pin VAE: mse={mse(vout, vref):.3e} psnr={psnr_255(vout, vref):.1f} dB")
    assert mse(vout, vref) <= 1e-2 and psnr_255(vout, vref) >= 35.0
