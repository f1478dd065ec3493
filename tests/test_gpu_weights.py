"""GPU: models loaded from safetensors files (diffusers names, official unified-file names, shards) produce the same
bits as models loaded tensor by tensor through ltxv_*_load_tensor."""
import json

import pytest
import torch

from oracle import ltx_oracle as O

pytestmark = pytest.mark.gpu


from tests.util import official_vae_keys, to_official_dit  # noqa: E402


def test_dit_from_safetensors_matches_load_tensor(cuda, tmp_path):
    import candle_video_b200 as cv
    from safetensors.torch import save_file
    from tests.test_gpu_dit import build, inputs, small_cfg
    cfg = small_cfg(layers=2)
    m_ref, w = build(cfg)
    w_bf16 = {k: v.to(torch.bfloat16).contiguous() for k, v in w.items()}
    m_ref.load_state_dict(w_bf16)
    hidden, enc, mask, coords = inputs(cfg, 1, 2, 8, 8, 16, n_keep=10)
    t = torch.tensor([993.0])
    args = (hidden.to(cuda), enc.to(cuda), t.to(cuda), mask.to(cuda), 2, 8, 8, None, coords.to(cuda))
    ref = m_ref.forward(*args)

    def fresh():
        return cv.LtxVideoTransformer3DModel(cv.DitConfig(
            num_attention_heads=cfg.num_attention_heads, attention_head_dim=cfg.attention_head_dim,
            cross_attention_dim=cfg.cross_attention_dim, num_layers=cfg.num_layers,
            caption_channels=cfg.caption_channels, timestep_bf16_round=True))

    # (1) diffusers names, one bf16 file
    f1 = tmp_path / "diffusion_pytorch_model.safetensors"
    save_file(w_bf16, str(f1))
    m1 = fresh()
    loaded, ignored = m1.load_safetensors(f1)
    assert (loaded, ignored) == (len(w), 0)
    assert torch.equal(m1.forward(*args), ref)
    # (2) official unified file: transformer + VAE + unrelated tensors under native names, f16 storage
    vw = O.init_vae_weights(O.VaeConfig(decoder_layers_per_block=(1, 1, 1, 1)), 7)
    unified = {to_official_dit(k): v.to(torch.bfloat16).to(torch.float16).contiguous() for k, v in w_bf16.items()}
    exact = all(torch.equal(unified[to_official_dit(k)].to(torch.float32), v.to(torch.float32)) for k, v in w_bf16.items())
    unified.update({k: v.contiguous() for k, v in official_vae_keys(vw).items()})
    unified["text_encoder.shared.weight"] = torch.zeros(4, 4)
    f2 = tmp_path / "ltx-video-unified.safetensors"
    save_file(unified, str(f2))
    m2 = fresh()
    loaded, ignored = m2.load_safetensors(f2, official=True)
    assert loaded == len(w) and ignored == len(vw) + 1
    out2 = m2.forward(*args)
    if exact:  # bf16 -> f16 is lossless for these magnitudes
        assert torch.equal(out2, ref)
    else:
        assert float((out2 - ref).abs().max()) < 1e-2
    # (3) shards + index
    d3 = tmp_path / "sharded"
    d3.mkdir()
    keys = sorted(w_bf16)
    half = len(keys) // 2
    wm = {}
    for fn, ks in (("a.safetensors", keys[:half]), ("b.safetensors", keys[half:])):
        save_file({k: w_bf16[k] for k in ks}, str(d3 / fn))
        wm.update({k: fn for k in ks})
    (d3 / "diffusion_pytorch_model.safetensors.index.json").write_text(json.dumps({"metadata": {}, "weight_map": wm}))
    m3 = fresh()
    assert m3.load_safetensors(d3) == (len(w), 0)
    assert torch.equal(m3.forward(*args), ref)
    # a file that lacks tensors fails loudly at finalize
    save_file({k: w_bf16[k] for k in keys[:half]}, str(tmp_path / "partial.safetensors"))
    with pytest.raises(cv.LtxvError, match="never loaded"):
        fresh().load_safetensors(tmp_path / "partial.safetensors")
    # shape mismatch is reported with the key
    bad = dict(w_bf16)
    bad["proj_out.bias"] = torch.zeros(7, dtype=torch.bfloat16)
    save_file(bad, str(tmp_path / "bad.safetensors"))
    with pytest.raises(cv.LtxvError, match="proj_out.bias"):
        fresh().load_safetensors(tmp_path / "bad.safetensors")


def test_vae_from_official_file_matches_load_tensor(cuda, tmp_path):
    import candle_video_b200 as cv
    from safetensors.torch import save_file
    from tests.test_gpu_vae import build
    m_ref, w, cfg = build()
    g = torch.Generator().manual_seed(1)
    z = torch.randn(1, 128, 2, 4, 4, generator=g).to(cuda)
    ts = torch.tensor([0.05], device=cuda)
    ref = m_ref.decode(z, ts)
    off = {k: v.contiguous() for k, v in official_vae_keys(w).items()}
    off["vae.encoder.down_blocks.0.res_blocks.0.conv1.conv.weight"] = torch.zeros(2, 2)  # encoder: ignored
    off["model.diffusion_model.patchify_proj.weight"] = torch.zeros(2, 2)               # transformer: ignored
    f = tmp_path / "unified.safetensors"
    save_file(off, str(f))
    m = cv.AutoencoderKLLtxVideo(cv.VaeConfig(decoder_layers_per_block=(1, 1, 1, 1)))
    loaded, ignored = m.load_safetensors(f, official=True)
    assert loaded == len(w) and ignored == 2
    assert torch.equal(m.decode(z, ts), ref)
    # diffusers-format VAE file (what vae/diffusion_pytorch_model.safetensors holds)
    f2 = tmp_path / "vae.safetensors"
    save_file({k: v.contiguous() for k, v in w.items()}, str(f2))
    m2 = cv.AutoencoderKLLtxVideo(cv.VaeConfig(decoder_layers_per_block=(1, 1, 1, 1)))
    assert m2.load_safetensors(f2)[0] == len(w)
    assert torch.equal(m2.decode(z, ts), ref)
