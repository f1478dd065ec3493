"""Weight-file handling on the host: the official->diffusers key remap pinned on the reference's own
unit-test vectors (weight_format.rs:171-268), tensor routing as examples/ltx-video/main.rs:480-497 does it, and the
safetensors / shard-index reader checked against files written by the `safetensors` package."""
import json

import pytest
import torch

import candle_video_b200 as cv

# (native key, diffusers key): every assert_eq! of weight_format.rs:171-268
REFERENCE_REMAP_VECTORS = [
    ("transformer.patchify_proj.weight", "transformer.proj_in.weight"),
    ("transformer.adaln_single.linear.weight", "transformer.time_embed.linear.weight"),
    ("encoder.down_blocks.0.res_blocks.0.conv1.weight", "encoder.down_blocks.0.resnets.0.conv1.weight"),
    ("encoder.down_blocks.1.conv.weight", "encoder.down_blocks.0.downsamplers.0.conv.weight"),
    ("encoder.down_blocks.2.res_blocks.0.conv1.weight", "encoder.down_blocks.1.resnets.0.conv1.weight"),
    ("encoder.down_blocks.6.res_blocks.0.weight", "encoder.down_blocks.3.resnets.0.weight"),
    ("encoder.down_blocks.8.res_blocks.0.weight", "encoder.mid_block.resnets.0.weight"),
    ("decoder.up_blocks.0.res_blocks.0.weight", "decoder.mid_block.resnets.0.weight"),
    ("decoder.up_blocks.1.conv.weight", "decoder.up_blocks.0.upsamplers.0.conv.weight"),
    ("decoder.up_blocks.2.res_blocks.0.weight", "decoder.up_blocks.0.resnets.0.weight"),
    ("decoder.up_blocks.8.res_blocks.0.weight", "decoder.up_blocks.3.resnets.0.weight"),
    ("decoder.last_time_embedder.weight", "decoder.time_embedder.weight"),
    ("per_channel_statistics.mean-of-means", "latents_mean"),
    ("per_channel_statistics.std-of-means", "latents_std"),
]


@pytest.mark.parametrize("native,diffusers", REFERENCE_REMAP_VECTORS)
def test_remap_matches_reference_vectors(native, diffusers):
    assert cv.remap_official_key_raw(native) == diffusers


def test_remap_other_rules_and_out_of_table_indices():
    # q_norm / k_norm (weight_format.rs:61-62), scale-shift table and norm3 renames (:71-73)
    assert cv.remap_official_key_raw("transformer_blocks.3.attn1.q_norm.weight") == "transformer_blocks.3.attn1.norm_q.weight"
    assert cv.remap_official_key_raw("transformer_blocks.3.attn2.k_norm.weight") == "transformer_blocks.3.attn2.norm_k.weight"
    assert cv.remap_official_key_raw("decoder.last_scale_shift_table") == "decoder.scale_shift_table"
    assert cv.remap_official_key_raw("decoder.up_blocks.2.res_blocks.1.norm3.norm.weight") == \
        "decoder.up_blocks.0.resnets.1.norm3.weight"
    # indices beyond the table keep their number (the `_ =>` arm, :96 and :137)
    assert cv.remap_official_key_raw("decoder.up_blocks.9.x") == "decoder.up_blocks.9.x"
    assert cv.remap_official_key_raw("encoder.down_blocks.12.y") == "encoder.down_blocks.12.y"
    # multi-digit indices are one number, not a prefix match of "1"
    assert cv.remap_official_key_raw("decoder.up_blocks.10.conv.weight") == "decoder.up_blocks.10.conv.weight"


def test_routing_and_prefix_strip_like_main_rs():
    assert cv.remap_official_key("model.diffusion_model.patchify_proj.weight") == ("proj_in.weight", "transformer")
    assert cv.remap_official_key("transformer.transformer_blocks.0.attn1.to_q.weight") == \
        ("transformer_blocks.0.attn1.to_q.weight", "transformer")
    assert cv.remap_official_key("vae.decoder.up_blocks.1.conv.conv.weight") == \
        ("decoder.up_blocks.0.upsamplers.0.conv.conv.weight", "vae")
    assert cv.remap_official_key("vae.per_channel_statistics.std-of-means") == ("latents_std", "vae")
    assert cv.remap_official_key("text_encoder.shared.weight")[1] == "other"
    # is_vae_key is tested first (main.rs:482): a decoder time embedder is VAE although it contains "time_embed"
    assert cv.remap_official_key("decoder.last_time_embedder.timestep_embedder.linear_1.weight") == \
        ("decoder.time_embedder.timestep_embedder.linear_1.weight", "vae")


def _tensors():
    g = torch.Generator().manual_seed(3)
    return {
        "proj_in.weight": torch.randn(8, 4, generator=g),
        "proj_in.bias": torch.randn(8, generator=g).to(torch.bfloat16),
        "blocks.0.scale_shift_table": torch.randn(6, 8, generator=g).to(torch.float16),
        "scalar": torch.tensor(2.5),
        "empty": torch.zeros(0, 3),
    }


def test_safetensors_index_matches_the_safetensors_package(tmp_path):
    from safetensors.torch import save_file
    t = _tensors()
    f = tmp_path / "model.safetensors"
    save_file(t, str(f), metadata={"format": "pt", "note": "a \"quoted\" value, {braces} and [brackets]"})
    got = cv.safetensors_list(f)
    names = {"torch.float32": "F32", "torch.bfloat16": "BF16", "torch.float16": "F16"}
    want = sorted((k, names[str(v.dtype)], list(v.shape), v.numel() * v.element_size()) for k, v in t.items())
    assert got == want
    # a directory holding the single file resolves to it (loader.rs:373-390)
    assert cv.safetensors_list(tmp_path) == want


def test_sharded_directory_through_index_json(tmp_path):
    from safetensors.torch import save_file
    t = _tensors()
    keys = sorted(t)
    shards = {"model-00001-of-00002.safetensors": keys[:2], "model-00002-of-00002.safetensors": keys[2:]}
    weight_map = {}
    for fn, ks in shards.items():
        save_file({k: t[k] for k in ks}, str(tmp_path / fn))
        weight_map.update({k: fn for k in ks})
    (tmp_path / "model.safetensors.index.json").write_text(json.dumps(
        {"metadata": {"total_size": 123}, "weight_map": weight_map}))
    got = cv.safetensors_list(tmp_path)
    assert sorted(g[0] for g in got) == keys
    # strict: a shard named by the index but missing on disk is an error (loader.rs:352-366)
    (tmp_path / "model-00002-of-00002.safetensors").unlink()
    with pytest.raises(cv.LtxvError, match="missing"):
        cv.safetensors_list(tmp_path)


def test_reader_rejects_garbage(tmp_path):
    bad = tmp_path / "bad.safetensors"
    bad.write_bytes(b"\x10\x00\x00\x00\x00\x00\x00\x00{\"a\": not json at all")
    with pytest.raises(cv.LtxvError):
        cv.safetensors_list(bad)
    short = tmp_path / "short.safetensors"
    short.write_bytes(b"\x01\x02")
    with pytest.raises(cv.LtxvError, match="too short"):
        cv.safetensors_list(short)
    huge = tmp_path / "huge.safetensors"
    huge.write_bytes((1 << 40).to_bytes(8, "little") + b"{}")
    with pytest.raises(cv.LtxvError, match="header length"):
        cv.safetensors_list(huge)
    with pytest.raises(cv.LtxvError):
        cv.safetensors_list(tmp_path / "does_not_exist")
    # offsets that do not match shape x dtype
    hdr = json.dumps({"w": {"dtype": "F32", "shape": [4], "data_offsets": [0, 8]}}).encode()
    inc = tmp_path / "inconsistent.safetensors"
    inc.write_bytes(len(hdr).to_bytes(8, "little") + hdr + b"\0" * 8)
    with pytest.raises(cv.LtxvError, match="inconsistent"):
        cv.safetensors_list(inc)


def test_official_names_round_trip_through_the_remap():
    """Every tensor name of the DiT and of the VAE decoder, renamed to the official convention, comes back to the
    diffusers name and is routed to the right component."""
    from oracle import ltx_oracle as O
    from tests.util import official_vae_keys, to_official_dit
    cfg = O.DitConfig(num_attention_heads=4, attention_head_dim=64, cross_attention_dim=256, num_layers=2,
                      caption_channels=256)
    for k in O.dit_weight_shapes(cfg):
        assert cv.remap_official_key(to_official_dit(k)) == (k, "transformer"), k
    w = {k: None for k in O.vae_weight_shapes(O.VaeConfig())} if hasattr(O, "vae_weight_shapes") else \
        O.init_vae_weights(O.VaeConfig(decoder_layers_per_block=(1, 1, 1, 1)), 7)
    off = official_vae_keys(w)
    assert len(off) == len(w)
    assert {cv.remap_official_key(k)[0] for k in off} == set(w)
    assert all(cv.remap_official_key(k)[1] == "vae" for k in off)


def _raw_safetensors(tmp_path, header: dict, data: bytes, name="bad.safetensors"):
    import json
    import struct
    h = json.dumps(header, separators=(",", ":")).encode()
    p = tmp_path / name
    p.write_bytes(struct.pack("<Q", len(h)) + h + data)
    return p


def test_safetensors_parser_rejects_malformed_headers(tmp_path):
    """Untrusted-file hygiene of the header parser (csrc/weights.cc): negative / overflowing dims, overlapping or
    non-covering data ranges and out-of-range integers fail loudly instead of wrapping around."""
    import candle_video_b200 as cv
    ok = _raw_safetensors(tmp_path, {"a": {"dtype": "F32", "shape": [2, 2], "data_offsets": [0, 16]}}, b"\0" * 16, "ok.safetensors")
    assert cv.safetensors_list(ok) == [("a", "F32", [2, 2], 16)]
    cases = {
        "negative": ({"a": {"dtype": "F32", "shape": [-2, -2], "data_offsets": [0, 16]}}, b"\0" * 16),
        "overflow": ({"a": {"dtype": "F32", "shape": [1 << 62, 8], "data_offsets": [0, 16]}}, b"\0" * 16),
        "overlap": ({"a": {"dtype": "F32", "shape": [4], "data_offsets": [0, 16]},
                     "b": {"dtype": "F32", "shape": [4], "data_offsets": [8, 24]}}, b"\0" * 24),
        "hole": ({"a": {"dtype": "F32", "shape": [2], "data_offsets": [0, 8]},
                  "b": {"dtype": "F32", "shape": [2], "data_offsets": [16, 24]}}, b"\0" * 24),
        "trailing": ({"a": {"dtype": "F32", "shape": [2], "data_offsets": [0, 8]}}, b"\0" * 12),
        "huge_int": ({"a": {"dtype": "F32", "shape": [99999999999999999999999], "data_offsets": [0, 8]}}, b"\0" * 8),
    }
    for tag, (hdr, data) in cases.items():
        p = _raw_safetensors(tmp_path, hdr, data, f"{tag}.safetensors")
        with pytest.raises(cv.LtxvError):
            cv.safetensors_list(p)
