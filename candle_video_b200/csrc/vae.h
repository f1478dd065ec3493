// C++ host-side mirror of `AutoencoderKLLtxVideo` (decoder half; vae.rs:1472-1727, :2037-2136): owns the re-laid-out
// bf16 conv weights and the NDHWC workspace, sequences the implicit-GEMM conv3d kernel and the glue kernels.
// Bound to `Box<dyn VaeLtxVideo>` (t2v_pipeline.rs:91-103) through the C ABI in ffi.cu.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "dit.h"  // LinearW
#include "model_common.h"

namespace ltxv {

struct ConvW {
    __nv_bfloat16* w = nullptr;  // [rows_out, 27*Cin], k = tap*Cin + c
    float* b = nullptr;          // [rows_out padded]
    int Cin = 0, Cout = 0, rows_out = 0;
    bool d2s = false;  // output channels stored sub-voxel major (upsampler)
};
struct ResnetW {
    ConvW conv1, conv2;
};
struct TimeEmbW {
    LinearW l1, l2;  // 256 -> dim -> dim
    int dim = 0;
};

class LtxVideoEncoder3d;

class AutoencoderKLLtxVideo {
public:
    AutoencoderKLLtxVideo(const ltxv_vae_config& cfg, int device);
    ~AutoencoderKLLtxVideo();

    const ltxv_vae_config& config() const { return cfg_; }
    int device() const { return device_; }
    PipeWs& pipe_ws() { return pipe_ws_; }
    void load_tensor(const std::string& key, const void* data, int dtype, const int64_t* shape, int rank);
    bool has_key(const std::string& key) const;
    // encoder half (vae_encoder.h); without it `encoder.*` keys are ignored like in a decode-only deployment
    void enable_encoder(const ltxv_vae_encoder_config& cfg);
    LtxVideoEncoder3d* encoder() const { return enc_.get(); }
    void init_random(uint64_t seed);
    void finalize();
    const float* latents_mean() const { return latents_mean_; }
    const float* latents_std() const { return latents_std_; }
    int spatial_compression_ratio() const { return 32; }   // vae.rs:85
    int temporal_compression_ratio() const { return 8; }   // vae.rs:86

    // Multi-GPU decode: the volume is cut into H-slabs, one per rank of `comm`; every conv input's halo rows are
    // written by the neighbouring rank's pixel-norm kernel straight into this rank's padded buffer (peer memory), and
    // every rank's conv_out epilogue stores its pixel rows into rank 0's frame buffer.  decode() then takes the full
    // (replicated) latent on every rank and delivers the video on rank 0.
    void set_comm(PeerComm* comm);

    // z [B, C, F, H, W] -> out f32/bf16 [B, 3, 8F-7, 32H, 32W]
    void decode(const void* z, int z_dtype, const float* timestep_dev, int B, int F, int H, int W, void* out,
                int out_dtype, int postprocess, cudaStream_t s);
    // decode_z dispatch of the reference (vae.rs:2037-2066) with its tiling knobs: temporal tiling first
    // (:2358-2434), then spatial tiling (:2225-2290), each tile decoded by decode() and the seams blended linearly
    // (:1927-2006).  tiling == nullptr or no branch taken -> plain decode().  Compatibility mode for --vae-tiling
    // users: the slab decode above makes tiling unnecessary on B200, but its output differs from the blended one.
    void decode_z(const void* z, int z_dtype, const float* timestep_dev, int B, int F, int H, int W, void* out,
                  int out_dtype, int postprocess, const ltxv_vae_tiling* tiling, cudaStream_t s);

private:
    enum SlotKind { PLAIN = 0, CONV_W = 1, CONV_B = 2 };
    struct VSlot {
        ParamSlot ps;
        int kind = PLAIN;
        ConvW* conv = nullptr;
    };
    void add_plain(const std::string& key, void* dst, bool bf16, std::vector<int64_t> shape);
    void add_conv(const std::string& prefix, ConvW& cw, int Cin, int Cout, bool d2s, int rows_out);
    void add_time_embedder(const std::string& prefix, TimeEmbW& te, int dim);
    float* vec(int64_t n);
    void ensure_workspace(int F, int H, int W);
    // what the consumer of a conv's output applies before ITS conv (the job of prep()); when the level is narrow
    // enough (C <= 256) the producing conv does it in its epilogue (EPI_CONV_NORM_PAD) and prep() is skipped
    struct Producer {
        const float* scale = nullptr;
        const float* shift = nullptr;
        int do_norm = 1, do_silu = 1;
    };
    void conv(const ConvW& cw, const void* a_padded, int T, int H, int W, int epi, void* out, const void* res, int post,
              cudaStream_t s, int fuse_level = -1, const Producer* next = nullptr);
    // `next`: producer parameters of whatever consumes the resnet's output; `ready`: in/out, the padded buffer that
    // already holds the (fused) producer output for the next conv, or null
    void resnet(const ResnetW& rw, int level, const float* ss, __nv_bfloat16*& x, __nv_bfloat16*& x_alt, cudaStream_t s,
                const Producer* next = nullptr, void** ready = nullptr);
    bool fusable(int level) const;
    void* pad_buf(int level, int k) const;
    void tiled_decode(const void* z, int z_dtype, const float* ts_b, int T, int H, int W, float* dst,
                      const ltxv_vae_tiling& tp, cudaStream_t s);
    void temporal_tiled_decode(const void* z, int z_dtype, const float* ts_b, int F, int H, int W, float* dst,
                               const ltxv_vae_tiling& tp, cudaStream_t s);
    // pixel-norm / modulate / SiLU into the level's padded conv input (+ halo exchange and barrier when sharded)
    void* prep(const void* x, int level, const float* scale, const float* shift, int do_norm, int do_silu, cudaStream_t s);

    ltxv_vae_config cfg_;
    int device_;
    PipeWs pipe_ws_;
    std::unique_ptr<LtxVideoEncoder3d> enc_;
    bool finalized_ = false;
    std::vector<std::unique_ptr<DevBuf>> storage_;
    std::map<std::string, VSlot> slots_;

    int ch_[4];  // channel width per level: 1024, 512, 256, 128
    ConvW conv_in_, conv_out_, ups_[3];
    std::vector<ResnetW> res_[4];  // mid, up0, up1, up2
    float* sst_[4] = {nullptr, nullptr, nullptr, nullptr};  // [n_res, 4, C] per level
    TimeEmbW te_[4], te_final_;
    float* sst_final_ = nullptr;  // [2, C3]
    float* tsm_ = nullptr;        // timestep_scale_multiplier (device scalar)
    float *latents_mean_ = nullptr, *latents_std_ = nullptr;

    // workspace
    int wsF_ = 0, wsH_ = 0, wsW_ = 0;
    int T_[4], H_[4], W_[4];
    DevBuf a0_, p_[4], p2_[4], xa_, xb_, hb_, cond_, out_f32_;
    DevBuf tz_sp_, tz_tm_, t_dec_, t_prev_, t_cur_, t_work_;  // tiled decode: sub-latents and decoded tiles
    // sharded decode
    PeerComm* comm_ = nullptr;
    int Hfull_[4] = {0, 0, 0, 0};
    size_t p_off_[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};  // double-buffered padded inputs in the symmetric heap
    int pp_[4] = {0, 0, 0, 0};
    size_t a0_off_ = 0, video_off_ = 0;
    int slab_h0_ = 0, slab_hfull_ = 0;  // unpatchify row offset / full height of the conv_out being launched
    // Ragged H-slabs: rank r owns latent rows [slab_row0(r), slab_row0(r) + slab_rows(r)); the first H mod N ranks
    // take one row more.  h_up_ / h_dn_ = slab rows of the neighbours above / below at level 0 (their padded buffers
    // have their own plane stride), lat_h0_ = first latent row of this rank.
    int h_up_ = 0, h_dn_ = 0, lat_h0_ = 0;
    // symmetric-heap carve-outs per decode geometry (the heap is a bump allocator: a geometry seen before reuses its
    // buffers -- their zero borders are still intact -- instead of carving again)
    struct SlabAlloc {
        size_t p_off[4][2];
        size_t a0_off, video_off;
    };
    std::map<std::vector<int64_t>, SlabAlloc> slab_allocs_;
};

}  // namespace ltxv
