"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/ltxv.h declares,
its host-only entry points (schedule math, presets) match the oracle, and compute entry points fail loudly without a GPU."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def header_symbols():
    txt = (ROOT / "include" / "ltxv.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ltxv_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import candle_video_b200 as cv
    syms = header_symbols()
    assert len(syms) >= 40
    lib = C.CDLL(str(cv.library_path()))
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(cv.EXPORTED_SYMBOLS) == syms


def test_header_cites_reference_interfaces():
    txt = (ROOT / "include" / "ltxv.h").read_text()
    for cite in ("t2v_pipeline.rs:63-83", "t2v_pipeline.rs:91-103", "t2v_pipeline.rs:474-550", "scheduler.rs:544-582"):
        assert cite in txt


def test_no_torch_types_in_signatures():
    txt = (ROOT / "include" / "ltxv.h").read_text()
    assert "torch" not in txt.lower() and "at::" not in txt and "#include <stdint.h>" in txt


def test_presets_match_reference_configs():
    import candle_video_b200 as cv
    c2, c13 = cv.DitConfig.preset("2b"), cv.DitConfig.preset("13b")
    assert (c2.num_layers, c2.num_attention_heads, c2.attention_head_dim, c2.cross_attention_dim) == (28, 32, 64, 2048)
    assert (c13.num_layers, c13.attention_head_dim, c13.cross_attention_dim, c13.caption_channels) == (48, 128, 4096, 4096)
    with pytest.raises(cv.LtxvError, match="unknown transformer preset"):
        cv.DitConfig.preset("7b")


def test_scheduler_host_math_matches_oracle():
    import candle_video_b200 as cv
    from oracle import ltx_oracle as O
    for S in (384, 4992, 13376):
        assert cv.calculate_shift(S) == pytest.approx(O.calculate_shift(S), abs=1e-6)
    for n, S in [(40, 4992), (8, 384), (2, 4992), (25, 13376)] + [(n, S) for n in (7, 20, 50) for S in range(96, 20000, 331)]:
        mu = O.calculate_shift(S)
        sig_o, ts_o = O.scheduler_set_timesteps(n, mu)
        sig_c, ts_c = cv.scheduler_set_timesteps(n, mu)
        assert len(sig_c) == n + 1 and sig_c[-1] == 0.0
        # every operation of the schedule is an IEEE f32 add / mul / div plus one libm expf (which the oracle calls
        # too): sigmas and the truncated integer timesteps are bit-exact, no +-1 allowance
        assert sig_o == sig_c
        assert ts_o == ts_c and ts_c[0] == 1000
    # degenerate n = 1 with terminal stretch: 0/0 in the reference as well -> NaN sigma, timestep 0
    sig_c, ts_c = cv.scheduler_set_timesteps(1, 1.3)
    assert sig_c[0] != sig_c[0] and ts_c == [0]
    s8 = [1.0, 0.9937, 0.9875, 0.9812, 0.975, 0.9094, 0.725, 0.4219]
    sig_c, ts_c = cv.scheduler_set_timesteps(8, 0.0, sigmas=s8, shift_terminal=None)
    sig_o, ts_o = O.scheduler_set_timesteps(8, 0.0, sigmas=s8, shift_terminal=None)
    assert sig_o == sig_c and ts_o == ts_c


def test_compute_entry_points_fail_loudly_without_gpu():
    import torch
    import candle_video_b200 as cv
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cv.LtxvError, match="no CPU fallback"):
        cv.LtxVideoTransformer3DModel(cv.DitConfig())
    with pytest.raises(cv.LtxvError, match="no CPU fallback"):
        cv.AutoencoderKLLtxVideo(cv.VaeConfig())
    with pytest.raises(cv.LtxvError, match="CUDA tensor"):
        cv.pack_latents(torch.zeros(1, 4, 1, 2, 2))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under candle_video_b200/ may import, link or execute it."""
    for p in (ROOT / "candle_video_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cc", ".h", ".cuh") and p.is_file():
            assert "oracle" not in p.read_text().replace("no CPU fallback", ""), p


def test_option_knobs_round_trip_and_reject_unknown_names():
    """ltxv_set_option / ltxv_get_option: every knob the header documents exists, holds the value it was
    given, and an unknown name is an error instead of a silent no-op.  Host-only: no GPU involved."""
    import candle_video_b200 as cv
    header = (ROOT / "include" / "ltxv.h").read_text()
    doc = header[header.index("Experiment knobs"):header.index("int ltxv_set_option")]
    names = re.findall(r"\b((?:no|gemm|conv|attn|vae|qk)_[a-z0-9_]+)\b", doc)
    assert len(names) >= 15 and "no_pdl" in names and "vae_prep_u" in names
    for n in names:
        before = cv.get_option(n)
        cv.set_option(n, 3)
        assert cv.get_option(n) == 3
        cv.set_option(n, before)
        assert cv.get_option(n) == before
    with pytest.raises(Exception):
        cv.set_option("no_such_knob", 1)
    with pytest.raises(Exception):
        cv.get_option("no_such_knob")


def test_rust_binding_declares_every_header_symbol():
    """integration/b200_ffi.rs (the file a maintainer of the reference drops into src/models/ltx_video/) declares exactly
    the entry points of include/ltxv.h -- no elisions, nothing that the library does not export."""
    header = re.sub(r"/\*.*?\*/", " ", (ROOT / "include" / "ltxv.h").read_text(), flags=re.S)  # comments out
    rust = re.sub(r"//[^\n]*", " ", (ROOT / "integration" / "b200_ffi.rs").read_text())
    declared_c = set(re.findall(r"\b(ltxv_[a-z0-9_]+)\s*\(", header))
    declared_rs = set(re.findall(r"\bfn\s+(ltxv_[a-z0-9_]+)\s*\(", rust))
    assert declared_c == declared_rs, (sorted(declared_c - declared_rs), sorted(declared_rs - declared_c))
    # argument counts agree too (a cheap guard against a signature drifting on one side only)
    def arity_c(name):
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", header, re.S)
        args = m.group(1).strip()
        return 0 if args in ("", "void") else args.count(",") + 1
    def arity_rs(name):
        m = re.search(r"\bfn\s+" + name + r"\s*\(([^;]*?)\)\s*(?:->[^;]*)?;", rust, re.S)
        args = m.group(1).strip()
        return 0 if args == "" else args.count(",") + 1
    bad = [(n, arity_c(n), arity_rs(n)) for n in sorted(declared_c) if arity_c(n) != arity_rs(n)]
    assert not bad, bad


def test_rust_model_shims_call_only_declared_entry_points():
    """integration/b200_models.rs (the trait implementations) must not call anything b200_ffi.rs does not declare."""
    ffi = set(re.findall(r"\bfn\s+(ltxv_[a-z0-9_]+)", (ROOT / "integration" / "b200_ffi.rs").read_text()))
    used = set(re.findall(r"\b(ltxv_[a-z0-9_]+)\s*\(", (ROOT / "integration" / "b200_models.rs").read_text()))
    assert len(used) >= 10 and used <= ffi, sorted(used - ffi)
