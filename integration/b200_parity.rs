//! ORACLE PINNING hand-off for the B200 drop-in (github.com/FerrisMind/candle-video, tests/ directory).
//!
//! Copy this file to `tests/b200_parity.rs` of the reference crate and run
//!
//!     cargo test --release --test b200_parity -- --nocapture
//!
//! It builds a small LTX-Video DiT and a (1,1,1,1)-layer VAE decoder from DETERMINISTIC PORTABLE weights (splitmix64 of
//! the flat index, seeded by the FNV-1a hash of the tensor's key; every value is exactly representable in bf16), runs
//! candle-video's own f32 CPU forward / decode, and writes `b200_parity_dump.safetensors`.  The B200 repository holds the
//! Python twin of the generator (`oracle/pin_weights.py`, same constants, same known-answer values) and a test
//! (`tests/test_reference_pin.py`) that compares its CPU oracle -- and, on a GPU box, the CUDA path -- with this dump:
//!
//!     LTXV_REFERENCE_DUMP=/path/to/b200_parity_dump.safetensors python -m pytest tests/test_reference_pin.py
//!
//! That comparison is what turns the oracle from "parity unpinned" into "pinned on reference output".
//!
//! Geometry (mirrors `PIN_*` in oracle/pin_weights.py):
//!   DiT : in/out 128, 4 heads x 64, 2 layers, caption 256; F,H,W = 2,8,8 (S = 128); K = 16 text tokens, 10 valid;
//!         timestep 992; video_coords given (fps 25).
//!   VAE : LtxVideoDecoder3d, decoder_block_out_channels [256,512,1024], layers [1,1,1,1], timestep conditioning on;
//!         latents [1,128,2,4,4] (the shape of tests/verify_vae_decode_parity.rs), temb 0.05.

use std::collections::HashMap;

use candle_core::{DType, Device, Tensor};
use candle_nn::VarBuilder;
use candle_video::models::ltx_video::ltx_transformer::{
    LtxVideoTransformer3DModel, LtxVideoTransformer3DModelConfig,
};
use candle_video::models::ltx_video::vae::LtxVideoDecoder3d;

const GOLDEN: u64 = 0x9E37_79B9_7F4A_7C15;

fn fnv1a64(s: &str) -> u64 {
    let mut h: u64 = 0xCBF2_9CE4_8422_2325;
    for b in s.as_bytes() {
        h = (h ^ (*b as u64)).wrapping_mul(0x0000_0100_0000_01B3);
    }
    h
}

fn splitmix64(seed: u64, idx: u64) -> u64 {
    let mut z = seed.wrapping_add(idx.wrapping_mul(GOLDEN));
    z = (z ^ (z >> 30)).wrapping_mul(0xBF58_476D_1CE4_E5B9);
    z = (z ^ (z >> 27)).wrapping_mul(0x94D0_49BB_1331_11EB);
    z ^ (z >> 31)
}

/// 2^-ceil(log2(sqrt(fan_in)))
fn pow2_bound(fan_in: usize) -> f32 {
    let mut k = 0u32;
    while (1usize << (2 * k)) < fan_in {
        k += 1;
    }
    1.0f32 / ((1u64 << k) as f32)
}

/// Same value rules as `pin_tensor` in oracle/pin_weights.py.
fn pin_tensor(key: &str, shape: &[usize], dev: &Device) -> candle_core::Result<Tensor> {
    let n: usize = shape.iter().product::<usize>().max(1);
    let seed = fnv1a64(key);
    let mut v = Vec::with_capacity(n);
    if key.ends_with("timestep_scale_multiplier") {
        v.resize(n, 1000.0f32);
    } else if key.contains("norm_q") || key.contains("norm_k") {
        for i in 0..n {
            let h = splitmix64(seed, i as u64);
            v.push(1.0f32 + ((h >> 59) as f32 - 16.0) / 64.0);
        }
    } else {
        let bound = if key.starts_with("input.") {
            2.0f32
        } else if key.ends_with(".bias") {
            1.0f32 / 32.0
        } else if key.ends_with("scale_shift_table") {
            pow2_bound(*shape.last().unwrap())
        } else {
            pow2_bound(shape[1..].iter().product::<usize>().max(1))
        };
        for i in 0..n {
            let h = splitmix64(seed, i as u64);
            let q = (h >> 57) as f32 * 2.0 - 127.0;
            v.push(q / 128.0 * bound);
        }
    }
    Tensor::from_vec(v, shape, dev)
}

fn put(map: &mut HashMap<String, Tensor>, key: &str, shape: &[usize], dev: &Device) -> candle_core::Result<()> {
    map.insert(key.to_string(), pin_tensor(key, shape, dev)?);
    Ok(())
}

/// Keys / shapes of LtxVideoTransformer3DModel (diffusers names; src/models/ltx_video/ltx_transformer.rs:957-1003).
fn dit_weights(d: usize, cross: usize, caption: usize, layers: usize, dev: &Device) -> candle_core::Result<HashMap<String, Tensor>> {
    let mut w = HashMap::new();
    put(&mut w, "proj_in.weight", &[d, 128], dev)?;
    put(&mut w, "proj_in.bias", &[d], dev)?;
    put(&mut w, "scale_shift_table", &[2, d], dev)?;
    put(&mut w, "time_embed.emb.timestep_embedder.linear_1.weight", &[d, 256], dev)?;
    put(&mut w, "time_embed.emb.timestep_embedder.linear_1.bias", &[d], dev)?;
    put(&mut w, "time_embed.emb.timestep_embedder.linear_2.weight", &[d, d], dev)?;
    put(&mut w, "time_embed.emb.timestep_embedder.linear_2.bias", &[d], dev)?;
    put(&mut w, "time_embed.linear.weight", &[6 * d, d], dev)?;
    put(&mut w, "time_embed.linear.bias", &[6 * d], dev)?;
    put(&mut w, "caption_projection.linear_1.weight", &[d, caption], dev)?;
    put(&mut w, "caption_projection.linear_1.bias", &[d], dev)?;
    put(&mut w, "caption_projection.linear_2.weight", &[d, d], dev)?;
    put(&mut w, "caption_projection.linear_2.bias", &[d], dev)?;
    put(&mut w, "proj_out.weight", &[128, d], dev)?;
    put(&mut w, "proj_out.bias", &[128], dev)?;
    for i in 0..layers {
        let p = format!("transformer_blocks.{i}.");
        put(&mut w, &format!("{p}scale_shift_table"), &[6, d], dev)?;
        for (a, kv_in) in [("attn1.", d), ("attn2.", cross)] {
            put(&mut w, &format!("{p}{a}to_q.weight"), &[d, d], dev)?;
            put(&mut w, &format!("{p}{a}to_q.bias"), &[d], dev)?;
            put(&mut w, &format!("{p}{a}to_k.weight"), &[d, kv_in], dev)?;
            put(&mut w, &format!("{p}{a}to_k.bias"), &[d], dev)?;
            put(&mut w, &format!("{p}{a}to_v.weight"), &[d, kv_in], dev)?;
            put(&mut w, &format!("{p}{a}to_v.bias"), &[d], dev)?;
            put(&mut w, &format!("{p}{a}to_out.0.weight"), &[d, d], dev)?;
            put(&mut w, &format!("{p}{a}to_out.0.bias"), &[d], dev)?;
            put(&mut w, &format!("{p}{a}norm_q.weight"), &[d], dev)?;
            put(&mut w, &format!("{p}{a}norm_k.weight"), &[d], dev)?;
        }
        put(&mut w, &format!("{p}ff.net.0.proj.weight"), &[4 * d, d], dev)?;
        put(&mut w, &format!("{p}ff.net.0.proj.bias"), &[4 * d], dev)?;
        put(&mut w, &format!("{p}ff.net.2.weight"), &[d, 4 * d], dev)?;
        put(&mut w, &format!("{p}ff.net.2.bias"), &[d], dev)?;
    }
    Ok(w)
}

fn time_embedder(w: &mut HashMap<String, Tensor>, prefix: &str, dim: usize, dev: &Device) -> candle_core::Result<()> {
    put(w, &format!("{prefix}timestep_embedder.linear_1.weight"), &[dim, 256], dev)?;
    put(w, &format!("{prefix}timestep_embedder.linear_1.bias"), &[dim], dev)?;
    put(w, &format!("{prefix}timestep_embedder.linear_2.weight"), &[dim, dim], dev)?;
    put(w, &format!("{prefix}timestep_embedder.linear_2.bias"), &[dim], dev)
}

fn resnets(w: &mut HashMap<String, Tensor>, prefix: &str, c: usize, n: usize, dev: &Device) -> candle_core::Result<()> {
    for i in 0..n {
        for cv in ["conv1", "conv2"] {
            put(w, &format!("{prefix}resnets.{i}.{cv}.conv.weight"), &[c, c, 3, 3, 3], dev)?;
            put(w, &format!("{prefix}resnets.{i}.{cv}.conv.bias"), &[c], dev)?;
        }
        put(w, &format!("{prefix}resnets.{i}.scale_shift_table"), &[4, c], dev)?;
    }
    Ok(())
}

/// Keys / shapes of LtxVideoDecoder3d with the VarBuilder rooted at "decoder" (src/models/ltx_video/vae.rs:1490-1610):
/// stage widths 1024, 512, 256, 128; one resnet per stage.
fn vae_decoder_weights(dev: &Device) -> candle_core::Result<HashMap<String, Tensor>> {
    let ch = [1024usize, 512, 256, 128];
    let mut w = HashMap::new();
    // keys carry the "decoder." prefix (that is what the generator hashes); the VarBuilder below is pushed to "decoder"
    put(&mut w, "decoder.conv_in.conv.weight", &[ch[0], 128, 3, 3, 3], dev)?;
    put(&mut w, "decoder.conv_in.conv.bias", &[ch[0]], dev)?;
    time_embedder(&mut w, "decoder.mid_block.time_embedder.", 4 * ch[0], dev)?;
    resnets(&mut w, "decoder.mid_block.", ch[0], 1, dev)?;
    for bi in 0..3 {
        let (cin, cout) = (ch[bi], ch[bi + 1]);
        let bp = format!("decoder.up_blocks.{bi}.");
        put(&mut w, &format!("{bp}upsamplers.0.conv.conv.weight"), &[cout * 8, cin, 3, 3, 3], dev)?;
        put(&mut w, &format!("{bp}upsamplers.0.conv.conv.bias"), &[cout * 8], dev)?;
        time_embedder(&mut w, &format!("{bp}time_embedder."), 4 * cout, dev)?;
        resnets(&mut w, &bp, cout, 1, dev)?;
    }
    put(&mut w, "decoder.conv_out.conv.weight", &[48, ch[3], 3, 3, 3], dev)?;
    put(&mut w, "decoder.conv_out.conv.bias", &[48], dev)?;
    time_embedder(&mut w, "decoder.time_embedder.", 2 * ch[3], dev)?;
    put(&mut w, "decoder.scale_shift_table", &[2, ch[3]], dev)?;
    put(&mut w, "decoder.timestep_scale_multiplier", &[], dev)?;
    Ok(w)
}

/// (clamp(8f-7, 0, 1000) * f32(1/fps), 32h, 32w), token order f,h,w  (src/models/ltx_video/t2v_pipeline.rs:798-847)
fn video_coords(f: usize, h: usize, w: usize, fps: f32, dev: &Device) -> candle_core::Result<Tensor> {
    let inv = 1.0f32 / fps;
    let mut v = Vec::with_capacity(f * h * w * 3);
    for fi in 0..f {
        for hi in 0..h {
            for wi in 0..w {
                let t = (8.0f32 * fi as f32 - 7.0).clamp(0.0, 1000.0) * inv;
                v.push(t);
                v.push(32.0 * hi as f32);
                v.push(32.0 * wi as f32);
            }
        }
    }
    Tensor::from_vec(v, (1, f * h * w, 3), dev)
}

#[test]
fn generator_known_answers() {
    // the same numbers are printed by `python -c "from oracle import pin_weights as P; print(P.KAT)"`
    assert_eq!(fnv1a64("proj_in.weight"), 12704714928704356426u64);
    assert_eq!(splitmix64(1, 0), 6238072747940578789u64);
    assert_eq!(splitmix64(1, 1), 10451216379200822465u64);
    assert_eq!(splitmix64(1, 2), 13757245211066428519u64);
    let dev = Device::Cpu;
    let t = pin_tensor("proj_in.weight", &[256, 128], &dev).unwrap();
    let first: Vec<f32> = t.flatten_all().unwrap().narrow(0, 0, 4).unwrap().to_vec1().unwrap();
    assert_eq!(first, vec![-0.02197265625f32, -0.05908203125, 0.00146484375, -0.06103515625]);
}

#[test]
fn dump_reference_outputs() -> anyhow::Result<()> {
    let dev = Device::Cpu; // the reference's CPU path (examples/ltx-video/main.rs:210-214, --cpu), f32
    let mut dump: HashMap<String, Tensor> = HashMap::new();

    // ---------------- DiT ----------------
    let (heads, hd, layers, caption) = (4usize, 64usize, 2usize, 256usize);
    let d = heads * hd;
    let config = LtxVideoTransformer3DModelConfig {
        in_channels: 128,
        out_channels: 128,
        patch_size: 1,
        patch_size_t: 1,
        num_attention_heads: heads,
        attention_head_dim: hd,
        cross_attention_dim: d,
        num_layers: layers,
        caption_channels: caption,
        qk_norm: "rms_norm_across_heads".to_string(),
        norm_elementwise_affine: false,
        norm_eps: 1e-6,
        attention_bias: true,
        attention_out_bias: true,
    };
    let vb = VarBuilder::from_tensors(dit_weights(d, d, caption, layers, &dev)?, DType::F32, &dev);
    let model = LtxVideoTransformer3DModel::new(&config, vb)?;
    let (f, h, w, k, keep) = (2usize, 8usize, 8usize, 16usize, 10usize);
    let hidden = pin_tensor("input.dit.hidden_states", &[1, f * h * w, 128], &dev)?;
    let enc = pin_tensor("input.dit.encoder_hidden_states", &[1, k, caption], &dev)?;
    let timestep = Tensor::from_vec(vec![992.0f32], 1, &dev)?;
    let mut m = vec![0.0f32; k];
    for v in m.iter_mut().take(keep) {
        *v = 1.0;
    }
    let mask = Tensor::from_vec(m, (1, k), &dev)?;
    let coords = video_coords(f, h, w, 25.0, &dev)?;
    let out = model.forward(&hidden, &enc, &timestep, Some(&mask), f, h, w, None, Some(&coords), None)?;
    dump.insert("dit.hidden_states".into(), hidden);
    dump.insert("dit.encoder_hidden_states".into(), enc);
    dump.insert("dit.timestep".into(), timestep);
    dump.insert("dit.encoder_attention_mask".into(), mask);
    dump.insert("dit.video_coords".into(), coords);
    dump.insert("dit.output".into(), out.to_dtype(DType::F32)?);

    // ---------------- VAE decoder ----------------
    let vb = VarBuilder::from_tensors(vae_decoder_weights(&dev)?, DType::F32, &dev);
    let decoder = LtxVideoDecoder3d::new(
        128,
        3,
        &[256, 512, 1024],
        &[true, true, true],
        &[1, 1, 1, 1],
        4,
        1,
        1e-6,
        false, // decoder_causal
        &[false, false, false, false],
        true, // timestep_conditioning
        &[true, true, true],
        &[2, 2, 2],
        vb.pp("decoder"),
    )?;
    let z = pin_tensor("input.vae.latents", &[1, 128, 2, 4, 4], &dev)?;
    let temb = Tensor::from_vec(vec![0.05f32], 1, &dev)?;
    let video = decoder.forward(&z, Some(&temb), false)?;
    dump.insert("vae.latents".into(), z);
    dump.insert("vae.temb".into(), temb);
    dump.insert("vae.output".into(), video.to_dtype(DType::F32)?);

    candle_core::safetensors::save(&dump, "b200_parity_dump.safetensors")?;
    println!("wrote b200_parity_dump.safetensors ({} tensors)", dump.len());
    Ok(())
}
