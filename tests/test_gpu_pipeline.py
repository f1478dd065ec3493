"""GPU parity of the denoise loop + decode branch (ltxv_pipeline_*) against the oracle driving the same sequence of
reference operations (t2v_pipeline.rs:860-1072): sequential CFG / STG passes, combine, Euler, unpack, denormalize,
decode, postprocess.  Pipeline bar: PSNR >= 35 dB on 0..255 (docs/benchmark_results.md:104)."""
import pytest
import torch

from oracle import ltx_oracle as O
from tests.util import max_abs, psnr_255, rel_l2

pytestmark = pytest.mark.gpu


def oracle_denoise(w, cfg, latents, pe, pm, ne, nm, F, H, W, fps, n_steps, g, r, s_stg, skip, sigmas_custom=None,
                   shift_terminal=0.1):
    S = F * H * W
    mu = 0.0 if sigmas_custom is not None else O.calculate_shift(S)
    sig, ts = O.scheduler_set_timesteps(n_steps, mu, sigmas_custom, shift_terminal)
    coords = O.video_coords(1, F, H, W, fps)
    do_cfg, do_stg = g > 1.0, s_stg > 0.0
    perm_skip = list(skip) if (skip and not do_stg) else []
    lat = latents.clone()
    for i, t in enumerate(ts):
        tt = torch.tensor([float(t)])
        kw = dict(num_frames=F, height=H, width=W, video_coords=coords, skip_block_list=perm_skip, timestep_to_bf16=True)
        u = O.dit_forward(w, cfg, lat, ne, tt, nm, **kw) if do_cfg else None
        c = O.dit_forward(w, cfg, lat, pe, tt, pm, **kw)
        p = None
        if do_stg:
            slm = torch.zeros(cfg.num_layers, 1)
            for l in skip or []:
                slm[l, 0] = 1.0
            p = O.dit_forward(w, cfg, lat, pe, tt, pm, skip_layer_mask=slm, **kw)
        comb = O.guidance_combine(c, u, p, g, r, s_stg)
        lat = O.euler_step(lat, comb, sig[i], sig[i + 1])
    return lat


@pytest.mark.parametrize("g,r,s_stg,skip", [(3.0, 0.0, 0.0, None), (3.0, 0.7, 1.0, [1]), (1.0, 0.0, 0.0, [1])])
def test_denoise_loop_matches_oracle(cuda, g, r, s_stg, skip):
    import candle_video_b200 as cv
    from tests.test_gpu_dit import build, small_cfg
    cfg = small_cfg(layers=3)
    m, w = build(cfg)
    m.set_skip_block_list([])
    height, width, frames, fps, K, n_steps = 256, 256, 9, 25, 24, 4
    F, H, W = (frames - 1) // 8 + 1, height // 32, width // 32
    gen = torch.Generator().manual_seed(11)
    lat = torch.randn(1, F * H * W, 128, generator=gen)
    pe, ne = torch.randn(1, K, 256, generator=gen), torch.randn(1, K, 256, generator=gen)
    pm, nm = torch.ones(1, K), torch.ones(1, K)
    pm[:, 17:] = 0
    nm[:, 5:] = 0
    ref = oracle_denoise(w, cfg, lat, pe, pm, ne, nm, F, H, W, fps, n_steps, g, r, s_stg, skip)
    params = cv.PipelineParams(height=height, width=width, num_frames=frames, frame_rate=fps,
                               num_inference_steps=n_steps, guidance_scale=g, guidance_rescale=r, stg_scale=s_stg,
                               skip_block_list=skip)
    out = lat[0].to(cuda).contiguous()
    cv.pipeline_denoise(m, params, out, pe.to(cuda), pm.to(cuda), ne.to(cuda), nm.to(cuda))
    e = rel_l2(out, ref[0])
    print(f"denoise g={g} r={r} stg={s_stg}: rel_l2={e:.3e} max_abs={max_abs(out, ref[0]):.3e}")
    assert torch.isfinite(out).all()
    assert e <= 2e-2
    # host-buffer entry point gives the same bits
    out_h = lat[0].clone().contiguous()
    cv.pipeline_denoise_host(m, params, out_h, pe, pm, ne, nm)
    assert torch.equal(out_h, out.cpu())


def test_decode_branch_matches_oracle(cuda):
    import candle_video_b200 as cv
    from tests.test_gpu_vae import build
    m, w, cfg = build()
    height, width, frames = 128, 160, 9
    F, H, W = 2, 4, 5
    gen = torch.Generator().manual_seed(12)
    lat = torch.randn(1, F * H * W, 128, generator=gen)
    z = O.denormalize_latents(O.unpack_latents(lat, F, H, W), torch.zeros(128), torch.ones(128), 1.0)
    ref = O.postprocess_video(O.vae_decode(w, cfg, z, torch.tensor([0.05])))
    params = cv.PipelineParams(height=height, width=width, num_frames=frames, decode_timestep=0.05)
    out = cv.pipeline_decode(m, params, lat[0].to(cuda).contiguous())
    assert out.shape == (3, frames, height, width)
    mse = float(((out.cpu().double() - ref[0].double()) ** 2).mean())
    import math
    psnr = 99.0 if mse == 0 else 10 * math.log10(255.0 ** 2 / mse)
    print(f"decode branch PSNR(0..255) = {psnr:.1f} dB")
    assert psnr >= 35.0
    host = cv.pipeline_decode_host(m, params, lat[0].contiguous())
    assert torch.equal(host, out.cpu())


def test_pipeline_rejects_bad_resolution(cuda):
    import candle_video_b200 as cv
    from tests.test_gpu_dit import build, small_cfg
    m, _ = build(small_cfg(layers=1))
    params = cv.PipelineParams(height=250, width=256, num_frames=9, num_inference_steps=2, guidance_scale=1.0)
    with pytest.raises(cv.LtxvError, match="divisible by 32"):
        cv.pipeline_denoise(m, params, torch.zeros(64, 128, device=cuda), torch.zeros(4, 256, device=cuda), None)


def test_stochastic_denoise_and_noisy_decode_match_oracle(cuda):
    """stochastic_sampling = true (preset 0.9.8-distilled, configs.rs:210) and decode_noise_scale = 0.025
    (configs.rs:234) with the noise tensors injected on both sides."""
    import candle_video_b200 as cv
    from tests.test_gpu_dit import build, small_cfg
    cfg = small_cfg(layers=2)
    m, w = build(cfg)
    m.set_skip_block_list([])
    height, width, frames, fps, K, n_steps = 128, 160, 9, 25, 16, 3
    F, H, W = 2, 4, 5
    S = F * H * W
    gen = torch.Generator().manual_seed(31)
    lat = torch.randn(1, S, 128, generator=gen)
    pe = torch.randn(1, K, 256, generator=gen)
    pm = torch.ones(1, K)
    noise = torch.randn(n_steps, S, 128, generator=gen)
    sig, ts = O.scheduler_set_timesteps(n_steps, O.calculate_shift(S), None, 0.1)
    coords = O.video_coords(1, F, H, W, fps)
    ref = lat.clone()
    for i, t in enumerate(ts):  # distilled: guidance 1 -> one forward per step
        v = O.dit_forward(w, cfg, ref, pe, torch.tensor([float(t)]), pm, num_frames=F, height=H, width=W,
                          video_coords=coords, timestep_to_bf16=True)
        ref = O.stochastic_step(ref, v, noise[i][None], sig[i], sig[i + 1])
    params = cv.PipelineParams(height=height, width=width, num_frames=frames, frame_rate=fps,
                               num_inference_steps=n_steps, guidance_scale=1.0, guidance_rescale=0.0, stg_scale=0.0)
    out = lat[0].to(cuda).contiguous()
    cv.pipeline_denoise_stochastic(m, params, out, pe.to(cuda), pm.to(cuda), noise.to(cuda).contiguous())
    e = rel_l2(out, ref[0])
    print(f"stochastic denoise rel_l2={e:.3e}")
    assert e <= 2e-2
    # noisy decode
    from tests.test_gpu_vae import build as build_vae
    vm, vw, vcfg = build_vae()
    dn = torch.randn(128, F, H, W, generator=gen)
    z = O.denormalize_latents(O.unpack_latents(lat, F, H, W), torch.zeros(128), torch.ones(128), 1.0)
    z = O.decode_noise_blend(z, dn[None], 0.025)
    vref = O.postprocess_video(O.vae_decode(vw, vcfg, z, torch.tensor([0.05])))
    vout = cv.pipeline_decode(vm, cv.PipelineParams(height=height, width=width, num_frames=frames, decode_timestep=0.05),
                              lat[0].to(cuda).contiguous(), decode_noise=dn.to(cuda), decode_noise_scale=0.025)
    mse = float(((vout.cpu().double() - vref[0].double()) ** 2).mean())
    import math
    psnr = 99.0 if mse == 0 else 10 * math.log10(255.0 ** 2 / mse)
    print(f"noisy decode PSNR = {psnr:.1f} dB")
    assert psnr >= 35.0
    with pytest.raises(cv.LtxvError, match="noise tensor"):
        cv.pipeline_decode(vm, cv.PipelineParams(height=height, width=width, num_frames=frames), lat[0].to(cuda).contiguous(),
                           decode_noise_scale=0.025)


def test_batched_cfg_pair_equals_sequential_forwards(cuda):
    """The CFG pair forward (uncond + cond as one 2S-token pass) must return the rows of the two sequential B = 1
    forwards the reference runs (t2v_pipeline.rs:878-907): same bits, with and without key masks."""
    import candle_video_b200 as cv
    from tests.test_gpu_dit import build, small_cfg
    m, _ = build(small_cfg(layers=3))
    m.set_skip_block_list([])
    height, width, frames, K, n_steps = 256, 288, 17, 24, 3   # latent 3 x 8 x 9 = 216 tokens (ragged row tiles)
    S = 3 * 8 * 9
    gen = torch.Generator().manual_seed(41)
    lat = torch.randn(S, 128, generator=gen).to(cuda)
    pe, ne = torch.randn(K, 256, generator=gen).to(cuda), torch.randn(K, 256, generator=gen).to(cuda)
    pm, nm = torch.ones(K, device=cuda), torch.ones(K, device=cuda)
    pm[17:] = 0
    for masks in ((pm, nm), (pm, None), (None, None)):
        params = cv.PipelineParams(height=height, width=width, num_frames=frames, num_inference_steps=n_steps,
                                   guidance_scale=3.0, guidance_rescale=0.7)
        a = lat.clone()
        cv.pipeline_denoise(m, params, a, pe, masks[0], ne, masks[1])
        # (LTXV_NO_CFG_BATCH is latched at first use inside the library: compare against explicit sequential forwards)
        b = lat.clone()
        sig, ts = O.scheduler_set_timesteps(n_steps, O.calculate_shift(S), None, 0.1)
        coords = cv.video_coords(1, 3, 8, 9, 25, device=cuda)[0]
        m.prepare_context(0, pe, masks[0])
        m.prepare_context(1, ne, masks[1])
        for i, t in enumerate(ts):
            tt = torch.tensor([float(t)], device=cuda)
            u = m.forward_ctx(1, b, tt, 3, 8, 9, None, coords)
            c = m.forward_ctx(0, b, tt, 3, 8, 9, None, coords)
            cv.guidance_euler_step(c[None], u[None], None, b[None], 3.0, 0.7, 0.0, float(sig[i]), float(sig[i + 1]))
        assert torch.equal(a, b)


def test_kernel_variants_are_bit_identical(cuda, monkeypatch):
    """CTA-pair / KW3 conv kernels vs the single-CTA kernel on a volume large enough to select them: the k-block order
    is the same in every variant, so the decode must not change by a single bit."""
    import candle_video_b200 as cv
    from tests.test_gpu_vae import build
    m, _, _ = build()
    z = torch.randn(1, 128, 3, 8, 12, generator=torch.Generator().manual_seed(8)).to(cuda)
    ts = torch.tensor([0.05], device=cuda)
    ref = m.decode(z, ts)
    monkeypatch.setenv("LTXV_CONV_NO_KW3", "1")
    no_kw3 = m.decode(z, ts)
    monkeypatch.setenv("LTXV_GEMM_NO_PAIR", "1")
    single = m.decode(z, ts)
    assert torch.isfinite(ref).all()
    assert torch.equal(ref, no_kw3)
    assert torch.equal(ref, single)


def test_decode_host_u8_matches_f32_path(cuda):
    """ltxv_pipeline_decode_host_u8 == frames_to_u8(ltxv_pipeline_decode_host) (same kernels, u8 hand-off on device)."""
    import candle_video_b200 as cv
    from tests.test_gpu_vae import build
    vae, _, _ = build()
    params = cv.PipelineParams(height=64, width=96, num_frames=9, frame_rate=25, num_inference_steps=1,
                               custom_sigmas=[0.9], guidance_scale=1.0, guidance_rescale=0.0, stg_scale=0.0,
                               shift_terminal=None, decode_timestep=0.05)
    lat = torch.randn(2 * 2 * 3, 128, generator=torch.Generator().manual_seed(4))
    f32 = cv.pipeline_decode_host(vae, params, lat)
    u8 = cv.pipeline_decode_host_u8(vae, params, lat)
    assert u8.shape == (9, 64, 96, 3) and u8.dtype == torch.uint8
    assert torch.equal(u8, O.frames_to_u8(f32[None])[0])
