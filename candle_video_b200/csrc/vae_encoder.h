// C++ host-side mirror of `LtxVideoEncoder3d` + `AutoencoderKLLtxVideo::encode` (vae.rs:1315-1469, :2017-2099), the
// 0.9.5 layout: pixel-unshuffle downsamplers (vae.rs:496-582), causal convs, no timestep conditioning.  SURVEY.md 8f-4:
// the encoder reuses the implicit-GEMM conv3d kernel (gemm.h) and the pixel-norm producer kernel of the decoder; only
// patchify, the downsampler tail and the moments layout are new (vae_glue.h).  Owned by AutoencoderKLLtxVideo
// (vae.h: enable_encoder / encode), which routes the `encoder.*` weight keys here.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "model_common.h"
#include "vae.h"

namespace ltxv {

class LtxVideoEncoder3d {
public:
    LtxVideoEncoder3d(const ltxv_vae_encoder_config& cfg, int device);

    const ltxv_vae_encoder_config& config() const { return cfg_; }
    bool has_key(const std::string& key) const { return slots_.count(key) != 0; }
    void load_tensor(const std::string& key, const void* data, int dtype, const int64_t* shape, int rank);
    void init_random(uint64_t seed);
    void finalize();

    // latent extent of an [F, H, W] video (frames 8k+1, H and W multiples of 32 for the default strides)
    void latent_dims(int F, int H, int W, int* Fl, int* Hl, int* Wl) const;
    // x [B, 3, F, H, W] NCDHW (f32 or bf16, device) -> moments [B, 2*latent, F', H', W'] f32: mean | logvar
    void encode(const void* x, int x_dtype, int B, int F, int H, int W, float* moments, cudaStream_t s);
    // encode_z dispatch of the reference (vae.rs:2017-2034) with its tiling knobs: temporal tiling when
    // use_framewise_encoding and F > tile_sample_min_num_frames (:2294-2356), else spatial tiling when H or W exceed the
    // minimum tile (:2158-2223), each tile encoded by encode() and the seams blended linearly in LATENT space
    // (:1927-2006).  tiling == nullptr or no branch taken -> plain encode().  Compatibility mode: the library default of
    // the reference is use_tiling = true, so its encode() of anything wider than 512 px is the blended result.
    void encode_z(const void* x, int x_dtype, int B, int F, int H, int W, const ltxv_vae_tiling* tiling,
                  int use_framewise_encoding, float* moments, cudaStream_t s);

private:
    struct Slot {
        ParamSlot ps;
        int kind = 0;  // 1 conv weight, 2 conv bias
        ConvW* conv = nullptr;
        int cin_src = 0;
    };
    void add_conv(const std::string& prefix, ConvW& cw, int cin_src, int Cin, int Cout, int rows_out);
    void ensure_workspace(int F, int H, int W);
    void tiled_encode(const void* x, int x_dtype, int T, int H, int W, float* dst, const ltxv_vae_tiling& tp, cudaStream_t s);
    void temporal_tiled_encode(const void* x, int x_dtype, int F, int H, int W, float* dst, const ltxv_vae_tiling& tp,
                               cudaStream_t s);
    void conv(const ConvW& cw, const void* a_padded, int T, int H, int W, int epi, void* out, const void* res, int n_cols,
              int ldo, cudaStream_t s, void* fused_norm_out = nullptr, int fused_raw = 0, int fused_tf = 2);
    void resnet(const ResnetW& rw, int level, __nv_bfloat16*& x, __nv_bfloat16*& x_alt, cudaStream_t s, bool* ready = nullptr,
                int next_kind = 0, int next_tf = 2);

    ltxv_vae_encoder_config cfg_;
    int device_;
    bool finalized_ = false;
    std::vector<std::unique_ptr<DevBuf>> storage_;
    std::map<std::string, Slot> slots_;

    int ch_[5];             // channel width per level: 128, 256, 512, 1024, 2048
    int stride_[4][3];      // (st, sh, sw) of the downsampler leaving level l
    ConvW conv_in_, conv_out_, down_[4];
    std::vector<ResnetW> res_[5];  // down_blocks.0..3, mid_block

    int wsF_ = 0, wsH_ = 0, wsW_ = 0;
    int T_[5], H_[5], W_[5];
    DevBuf a_in_, p_[5], q_[5], xa_, xb_, hb_, out32_;
    DevBuf tx_sp_, tx_tm_, t_enc_, t_prev_, t_cur_, t_work_;  // tiled encode: sub-videos and encoded tiles
};

}  // namespace ltxv
