// extern "C" boundary (include/ltxv.h): exceptions -> error codes, handles -> C++ model mirrors.
#include <string.h>

#include "attention.h"
#include "comm.h"
#include "conv3d_op.h"
#include "dit.h"
#include "gemm.h"
#include "glue.h"
#include "options.h"
#include "pipeline.h"
#include "profile.h"
#include "vae.h"
#include "vae_encoder.h"
#include "weights.h"
#include "vae_glue.h"

using namespace ltxv;

struct ltxv_dit {
    LtxVideoTransformer3DModel model;
    DevBuf h_hidden, h_enc, h_t, h_mask, h_coords, h_out;  // device staging for the *_host entry points
    DevBuf p_lat, p_pe, p_ne, p_pm, p_nm;                   // ... and for ltxv_pipeline_denoise_host
    ltxv_dit(const ltxv_dit_config& c, int dev) : model(c, dev) {}
};
struct ltxv_comm {
    PeerComm comm;
    ltxv_comm(int n, int r, int dev, size_t heap) : comm(n, r, dev, heap) {}
};
struct ltxv_vae {
    AutoencoderKLLtxVideo model;
    DevBuf h_z, h_t, h_out;
    DevBuf p_lat, p_out, p_u8;  // staging of ltxv_pipeline_decode_host / _host_u8 (per handle: lives on the model's device)
    ltxv_vae(const ltxv_vae_config& c, int dev) : model(c, dev) {}
};

#define LTXV_TRY try {
#define LTXV_CATCH                                        \
    }                                                     \
    catch (const std::exception& e) {                     \
        set_error("%s", e.what());                        \
        return 1;                                         \
    }                                                     \
    catch (...) {                                         \
        set_error("unknown C++ exception");               \
        return 1;                                         \
    }                                                     \
    return 0;

static size_t dsize(int dtype) {
    if (dtype == LTXV_F32) return 4;
    if (dtype == LTXV_BF16) return 2;
    fail("unsupported dtype code %d", dtype);
}

extern "C" {

const char* ltxv_last_error(void) { return last_error_ref().c_str(); }
const char* ltxv_version(void) { return "ltxv_b200 0.1 (sm_100a: tcgen05 GEMM/conv3d/attention)"; }
uint64_t ltxv_launch_count(void) {
    return gemm_launch_count() + attention_launch_count() + glue_launch_count() + vae_glue_launch_count() +
           common_launch_count() + comm_launch_count() + conv_op_launch_count();
}

int ltxv_profile_begin(void) {
    profiling_begin();
    return 0;
}
int ltxv_profile_end(uint64_t* launches8, double* ms8, double* work8) {
    LTXV_TRY
    if (launches8 == nullptr || ms8 == nullptr || work8 == nullptr) fail("null argument");
    profiling_end(launches8, ms8, work8);
    LTXV_CATCH
}

int ltxv_causal_conv3d(const float* x, const float* weight, const float* bias, int in_channels, int out_channels,
                       int T, int H, int W, int is_causal, float* out, void* stream) {
    LTXV_TRY
    if (x == nullptr || weight == nullptr || out == nullptr) fail("null argument");
    causal_conv3d(x, weight, bias, in_channels, out_channels, T, H, W, is_causal != 0, out,
                  static_cast<cudaStream_t>(stream));
    LTXV_CATCH
}

int ltxv_set_option(const char* name, int value) {
    LTXV_TRY
    if (name == nullptr) fail("null argument");
    if (!set_option(name, value)) fail("unknown option '%s'", name);
    LTXV_CATCH
}
int ltxv_get_option(const char* name, int* value) {
    LTXV_TRY
    if (name == nullptr || value == nullptr) fail("null argument");
    if (!get_option(name, value)) fail("unknown option '%s'", name);
    LTXV_CATCH
}

int ltxv_trace_begin(void) {
    trace_begin();
    return 0;
}
int ltxv_trace_end(char* out, uint64_t out_cap) {
    LTXV_TRY
    if (out == nullptr) fail("null argument");
    const char* t = trace_end();
    const size_t n = strlen(t);
    if (n + 1 > out_cap) fail("output buffer too small for the variant trace (%zu bytes needed)", n + 1);
    memcpy(out, t, n + 1);
    LTXV_CATCH
}

int ltxv_remap_official_key(const char* key, char* out, uint64_t out_cap, int32_t* component) {
    LTXV_TRY
    if (key == nullptr || out == nullptr) fail("null argument");
    int c = WEIGHT_OTHER;
    const std::string r = official_key_to_model_key(key, &c);
    if (r.size() + 1 > out_cap) fail("output buffer too small for the remapped key (%zu bytes needed)", r.size() + 1);
    memcpy(out, r.c_str(), r.size() + 1);
    if (component) *component = c;
    LTXV_CATCH
}
int ltxv_remap_official_key_raw(const char* key, char* out, uint64_t out_cap) {
    LTXV_TRY
    if (key == nullptr || out == nullptr) fail("null argument");
    const std::string r = remap_official_key(key);
    if (r.size() + 1 > out_cap) fail("output buffer too small for the remapped key (%zu bytes needed)", r.size() + 1);
    memcpy(out, r.c_str(), r.size() + 1);
    LTXV_CATCH
}
int ltxv_safetensors_list(const char* path, char* out, uint64_t out_cap, int32_t* n_tensors) {
    LTXV_TRY
    if (path == nullptr || out == nullptr) fail("null argument");
    std::string text;
    int n = 0;
    for (const std::string& f : resolve_safetensors_files(path)) {
        SafeTensorsFile st(f);
        for (const SafeTensorInfo& t : st.tensors()) {
            text += t.name + " " + t.dtype + " [";
            for (size_t i = 0; i < t.shape.size(); ++i) text += (i ? "," : "") + std::to_string(t.shape[i]);
            text += "] " + std::to_string(t.bytes) + "\n";
            ++n;
        }
    }
    if (text.size() + 1 > out_cap) fail("output buffer too small for the tensor list (%zu bytes needed)", text.size() + 1);
    memcpy(out, text.c_str(), text.size() + 1);
    if (n_tensors) *n_tensors = n;
    LTXV_CATCH
}

int ltxv_dit_config_preset(const char* name, ltxv_dit_config* out) {
    LTXV_TRY
    if (out == nullptr || name == nullptr) fail("null argument");
    ltxv_dit_config c{};
    c.in_channels = 128; c.out_channels = 128; c.patch_size = 1; c.patch_size_t = 1;
    c.num_attention_heads = 32; c.attention_head_dim = 64; c.cross_attention_dim = 2048; c.num_layers = 28;
    c.caption_channels = 4096; c.norm_eps = 1e-6f; c.timestep_bf16_round = 1;
    if (strcmp(name, "2b") == 0) {
    } else if (strcmp(name, "13b") == 0) {  // configs.rs:151-160
        c.attention_head_dim = 128; c.cross_attention_dim = 4096; c.num_layers = 48;
    } else {
        fail("unknown transformer preset '%s' (expected 2b or 13b)", name);
    }
    *out = c;
    LTXV_CATCH
}

int ltxv_dit_create(const ltxv_dit_config* cfg, int device, ltxv_dit** out) {
    LTXV_TRY
    if (cfg == nullptr || out == nullptr) fail("null argument");
    *out = new ltxv_dit(*cfg, device);
    LTXV_CATCH
}
void ltxv_dit_destroy(ltxv_dit* m) { delete m; }

int ltxv_dit_load_tensor(ltxv_dit* m, const char* key, const void* data, int dtype, const int64_t* shape, int rank) {
    LTXV_TRY
    if (m == nullptr || key == nullptr || data == nullptr) fail("null argument");
    m->model.load_tensor(key, data, dtype, shape, rank);
    LTXV_CATCH
}
int ltxv_dit_load_safetensors(ltxv_dit* m, const char* path, int official, int32_t* n_loaded, int32_t* n_ignored) {
    LTXV_TRY
    if (m == nullptr || path == nullptr) fail("null argument");
    WeightSink sink;
    sink.wants = [m](const std::string& k) { return m->model.has_key(k); };
    sink.load = [m](const std::string& k, const void* d, int dt, const int64_t* sh, int r) {
        m->model.load_tensor(k, d, dt, sh, r);
    };
    int a = 0, b = 0;
    load_safetensors(path, official != 0, WEIGHT_TRANSFORMER, sink, &a, &b);
    if (n_loaded) *n_loaded = a;
    if (n_ignored) *n_ignored = b;
    LTXV_CATCH
}
int ltxv_dit_init_random(ltxv_dit* m, uint64_t seed) {
    LTXV_TRY
    if (m == nullptr) fail("null handle");
    m->model.init_random(seed);
    LTXV_CATCH
}
int ltxv_dit_finalize(ltxv_dit* m) {
    LTXV_TRY
    if (m == nullptr) fail("null handle");
    m->model.finalize();
    LTXV_CATCH
}
int ltxv_dit_set_skip_blocks(ltxv_dit* m, const int32_t* idx, int n) {
    LTXV_TRY
    if (m == nullptr) fail("null handle");
    m->model.set_skip_block_list(idx, n);
    LTXV_CATCH
}
int ltxv_dit_get_config(const ltxv_dit* m, ltxv_dit_config* out) {
    LTXV_TRY
    if (m == nullptr || out == nullptr) fail("null argument");
    *out = m->model.config();
    LTXV_CATCH
}

int ltxv_dit_forward(ltxv_dit* m, const void* hidden, int hidden_dtype, const void* enc, int enc_dtype,
                     const float* timestep, const float* mask, int B, int S, int K, int F, int H, int W,
                     const float* rope_scale3, const float* video_coords, const float* skip_layer_mask, void* out,
                     int out_dtype, void* stream) {
    LTXV_TRY
    if (m == nullptr || hidden == nullptr || enc == nullptr || timestep == nullptr || out == nullptr)
        fail("null argument");
    m->model.forward(hidden, hidden_dtype, enc, enc_dtype, timestep, mask, B, S, K, F, H, W, rope_scale3, video_coords,
                     skip_layer_mask, out, out_dtype, static_cast<cudaStream_t>(stream));
    LTXV_CATCH
}

int ltxv_dit_forward_host(ltxv_dit* m, const void* hidden, int hidden_dtype, const void* enc, int enc_dtype,
                          const float* timestep, const float* mask, int B, int S, int K, int F, int H, int W,
                          const float* rope_scale3, const float* video_coords, const float* skip_layer_mask, void* out,
                          int out_dtype) {
    LTXV_TRY
    if (m == nullptr || hidden == nullptr || enc == nullptr || timestep == nullptr || out == nullptr)
        fail("null argument");
    const ltxv_dit_config& c = m->model.config();
    const size_t nh = static_cast<size_t>(B) * S * c.in_channels * dsize(hidden_dtype);
    const size_t ne = static_cast<size_t>(B) * K * c.caption_channels * dsize(enc_dtype);
    const size_t no = static_cast<size_t>(B) * S * c.out_channels * dsize(out_dtype);
    cudaStream_t s = 0;
    LTXV_CUDA(cudaSetDevice(m->model.device()));
    m->h_hidden.ensure(nh);
    m->h_enc.ensure(ne);
    m->h_t.ensure(B * 4);
    m->h_out.ensure(no);
    LTXV_CUDA(cudaMemcpyAsync(m->h_hidden.p, hidden, nh, cudaMemcpyHostToDevice, s));
    LTXV_CUDA(cudaMemcpyAsync(m->h_enc.p, enc, ne, cudaMemcpyHostToDevice, s));
    LTXV_CUDA(cudaMemcpyAsync(m->h_t.p, timestep, B * 4, cudaMemcpyHostToDevice, s));
    const float* dmask = nullptr;
    if (mask != nullptr) {
        m->h_mask.ensure(static_cast<size_t>(B) * K * 4);
        LTXV_CUDA(cudaMemcpyAsync(m->h_mask.p, mask, static_cast<size_t>(B) * K * 4, cudaMemcpyHostToDevice, s));
        dmask = m->h_mask.as<float>();
    }
    const float* dcoords = nullptr;
    if (video_coords != nullptr) {
        m->h_coords.ensure(static_cast<size_t>(B) * S * 3 * 4);
        LTXV_CUDA(cudaMemcpyAsync(m->h_coords.p, video_coords, static_cast<size_t>(B) * S * 3 * 4, cudaMemcpyHostToDevice, s));
        dcoords = m->h_coords.as<float>();
    }
    m->model.forward(m->h_hidden.p, hidden_dtype, m->h_enc.p, enc_dtype, m->h_t.as<float>(), dmask, B, S, K, F, H, W,
                     rope_scale3, dcoords, skip_layer_mask, m->h_out.p, out_dtype, s);
    LTXV_CUDA(cudaMemcpyAsync(out, m->h_out.p, no, cudaMemcpyDeviceToHost, s));
    LTXV_CUDA(cudaStreamSynchronize(s));
    LTXV_CATCH
}

int ltxv_dit_prepare_context(ltxv_dit* m, int slot, const void* enc, int enc_dtype, const float* mask, int K,
                             void* stream) {
    LTXV_TRY
    if (m == nullptr || enc == nullptr) fail("null argument");
    if (slot < 0 || slot >= LtxVideoTransformer3DModel::kNumSlots) fail("context slot %d out of range", slot);
    m->model.prepare_context(slot, enc, enc_dtype, mask, K, static_cast<cudaStream_t>(stream));
    LTXV_CATCH
}
int ltxv_dit_forward_ctx(ltxv_dit* m, int slot, const void* hidden, int hidden_dtype, const float* timestep, int S,
                         int F, int H, int W, const float* rope_scale3, const float* video_coords,
                         const float* skip_layer_mask, void* out, int out_dtype, void* stream) {
    LTXV_TRY
    if (m == nullptr || hidden == nullptr || timestep == nullptr || out == nullptr) fail("null argument");
    if (slot < 0 || slot >= LtxVideoTransformer3DModel::kNumSlots) fail("context slot %d out of range", slot);
    m->model.forward_ctx(slot, hidden, hidden_dtype, timestep, S, F, H, W, rope_scale3, video_coords, skip_layer_mask,
                         1, out, out_dtype, static_cast<cudaStream_t>(stream));
    LTXV_CATCH
}

/* ---------------------------------------------------- VAE ---------------------------------------------------- */
int ltxv_vae_config_default(ltxv_vae_config* out) {
    LTXV_TRY
    if (out == nullptr) fail("null argument");
    ltxv_vae_config c{};
    c.latent_channels = 128; c.out_channels = 3;
    c.decoder_block_out_channels[0] = 256; c.decoder_block_out_channels[1] = 512; c.decoder_block_out_channels[2] = 1024;
    for (int i = 0; i < 4; ++i) c.decoder_layers_per_block[i] = 5;
    c.patch_size = 4; c.timestep_conditioning = 1; c.scaling_factor = 1.0f;
    *out = c;
    LTXV_CATCH
}
int ltxv_vae_create(const ltxv_vae_config* cfg, int device, ltxv_vae** out) {
    LTXV_TRY
    if (cfg == nullptr || out == nullptr) fail("null argument");
    *out = new ltxv_vae(*cfg, device);
    LTXV_CATCH
}
void ltxv_vae_destroy(ltxv_vae* m) { delete m; }
int ltxv_vae_load_tensor(ltxv_vae* m, const char* key, const void* data, int dtype, const int64_t* shape, int rank) {
    LTXV_TRY
    if (m == nullptr || key == nullptr || data == nullptr) fail("null argument");
    m->model.load_tensor(key, data, dtype, shape, rank);
    LTXV_CATCH
}
int ltxv_vae_load_safetensors(ltxv_vae* m, const char* path, int official, int32_t* n_loaded, int32_t* n_ignored) {
    LTXV_TRY
    if (m == nullptr || path == nullptr) fail("null argument");
    WeightSink sink;
    sink.wants = [m](const std::string& k) { return m->model.has_key(k); };
    sink.load = [m](const std::string& k, const void* d, int dt, const int64_t* sh, int r) {
        m->model.load_tensor(k, d, dt, sh, r);
    };
    int a = 0, b = 0;
    load_safetensors(path, official != 0, WEIGHT_VAE, sink, &a, &b);
    if (n_loaded) *n_loaded = a;
    if (n_ignored) *n_ignored = b;
    LTXV_CATCH
}
int ltxv_vae_init_random(ltxv_vae* m, uint64_t seed) {
    LTXV_TRY
    if (m == nullptr) fail("null handle");
    m->model.init_random(seed);
    LTXV_CATCH
}
int ltxv_vae_finalize(ltxv_vae* m) {
    LTXV_TRY
    if (m == nullptr) fail("null handle");
    m->model.finalize();
    LTXV_CATCH
}
const float* ltxv_vae_latents_mean(const ltxv_vae* m) { return m ? m->model.latents_mean() : nullptr; }
const float* ltxv_vae_latents_std(const ltxv_vae* m) { return m ? m->model.latents_std() : nullptr; }
int ltxv_vae_spatial_compression_ratio(const ltxv_vae* m) { return m ? m->model.spatial_compression_ratio() : 0; }
int ltxv_vae_temporal_compression_ratio(const ltxv_vae* m) { return m ? m->model.temporal_compression_ratio() : 0; }

int ltxv_vae_decode(ltxv_vae* m, const void* z, int z_dtype, const float* timestep, int B, int F, int H, int W,
                    void* out, int out_dtype, int postprocess, void* stream) {
    LTXV_TRY
    if (m == nullptr || z == nullptr || out == nullptr) fail("null argument");
    m->model.decode(z, z_dtype, timestep, B, F, H, W, out, out_dtype, postprocess, static_cast<cudaStream_t>(stream));
    LTXV_CATCH
}
int ltxv_vae_encoder_config_default(ltxv_vae_encoder_config* out) {
    LTXV_TRY
    if (out == nullptr) fail("null argument");
    // AutoencoderKLLtxVideoConfig::default (vae.rs:68-103)
    const ltxv_vae_encoder_config c = {3, 128, {128, 256, 512, 1024, 2048}, {4, 6, 6, 2, 2},
        {LTXV_DOWN_SPATIAL, LTXV_DOWN_TEMPORAL, LTXV_DOWN_SPATIOTEMPORAL, LTXV_DOWN_SPATIOTEMPORAL}, 4};
    *out = c;
    LTXV_CATCH
}
int ltxv_vae_enable_encoder(ltxv_vae* m, const ltxv_vae_encoder_config* cfg) {
    LTXV_TRY
    if (m == nullptr || cfg == nullptr) fail("null argument");
    m->model.enable_encoder(*cfg);
    LTXV_CATCH
}
static LtxVideoEncoder3d& encoder_of(const ltxv_vae* m) {
    if (m == nullptr) fail("null handle");
    if (m->model.encoder() == nullptr) fail("this VAE has no encoder: call ltxv_vae_enable_encoder first");
    return *m->model.encoder();
}
int ltxv_vae_encode_dims(const ltxv_vae* m, int F, int H, int W, int32_t* Fl, int32_t* Hl, int32_t* Wl) {
    LTXV_TRY
    if (Fl == nullptr || Hl == nullptr || Wl == nullptr) fail("null argument");
    int a, b, c;
    encoder_of(m).latent_dims(F, H, W, &a, &b, &c);
    *Fl = a;
    *Hl = b;
    *Wl = c;
    LTXV_CATCH
}
int ltxv_vae_encode(ltxv_vae* m, const void* x, int x_dtype, int B, int F, int H, int W, float* moments, void* stream) {
    LTXV_TRY
    if (x == nullptr || moments == nullptr) fail("null argument");
    encoder_of(m).encode(x, x_dtype, B, F, H, W, moments, static_cast<cudaStream_t>(stream));
    LTXV_CATCH
}
int ltxv_vae_encode_tiled(ltxv_vae* m, const void* x, int x_dtype, int B, int F, int H, int W,
                          const ltxv_vae_tiling* tiling, int use_framewise_encoding, float* moments, void* stream) {
    LTXV_TRY
    if (x == nullptr || moments == nullptr) fail("null argument");
    encoder_of(m).encode_z(x, x_dtype, B, F, H, W, tiling, use_framewise_encoding, moments, static_cast<cudaStream_t>(stream));
    LTXV_CATCH
}
int ltxv_vae_encode_host(ltxv_vae* m, const void* x, int x_dtype, int B, int F, int H, int W, float* moments) {
    LTXV_TRY
    if (x == nullptr || moments == nullptr) fail("null argument");
    LtxVideoEncoder3d& e = encoder_of(m);
    int fl, hl, wl;
    e.latent_dims(F, H, W, &fl, &hl, &wl);
    const size_t in_bytes = static_cast<size_t>(B) * 3 * F * H * W * dsize(x_dtype);
    const size_t out_bytes = static_cast<size_t>(B) * 2 * e.config().latent_channels * fl * hl * wl * 4;
    LTXV_CUDA(cudaSetDevice(m->model.device()));
    m->h_z.ensure(in_bytes);
    m->h_out.ensure(out_bytes);
    LTXV_CUDA(cudaMemcpy(m->h_z.p, x, in_bytes, cudaMemcpyHostToDevice));
    e.encode(m->h_z.p, x_dtype, B, F, H, W, m->h_out.as<float>(), 0);
    LTXV_CUDA(cudaMemcpy(moments, m->h_out.p, out_bytes, cudaMemcpyDeviceToHost));
    LTXV_CATCH
}
int ltxv_vae_tiling_default(ltxv_vae_tiling* out) {
    LTXV_TRY
    if (out == nullptr) fail("null argument");
    *out = ltxv_vae_tiling{1, 1, 512, 512, 16, 384, 384, 8};  // vae.rs:1848-1861
    LTXV_CATCH
}
int ltxv_vae_decode_tiled(ltxv_vae* m, const void* z, int z_dtype, const float* timestep, int B, int F, int H, int W,
                          const ltxv_vae_tiling* tiling, void* out, int out_dtype, int postprocess, void* stream) {
    LTXV_TRY
    if (m == nullptr || z == nullptr || out == nullptr) fail("null argument");
    m->model.decode_z(z, z_dtype, timestep, B, F, H, W, out, out_dtype, postprocess, tiling,
                      static_cast<cudaStream_t>(stream));
    LTXV_CATCH
}
int ltxv_vae_decode_host(ltxv_vae* m, const void* z, int z_dtype, const float* timestep, int B, int F, int H, int W,
                         void* out, int out_dtype, int postprocess) {
    LTXV_TRY
    if (m == nullptr || z == nullptr || out == nullptr) fail("null argument");
    const int C = m->model.config().latent_channels;
    const size_t nz = static_cast<size_t>(B) * C * F * H * W * dsize(z_dtype);
    const size_t no = static_cast<size_t>(B) * 3 * (8 * F - 7) * (32 * H) * (32 * W) * dsize(out_dtype);
    cudaStream_t s = 0;
    LTXV_CUDA(cudaSetDevice(m->model.device()));
    m->h_z.ensure(nz);
    m->h_out.ensure(no);
    LTXV_CUDA(cudaMemcpyAsync(m->h_z.p, z, nz, cudaMemcpyHostToDevice, s));
    const float* dt = nullptr;
    if (timestep != nullptr) {
        m->h_t.ensure(B * 4);
        LTXV_CUDA(cudaMemcpyAsync(m->h_t.p, timestep, B * 4, cudaMemcpyHostToDevice, s));
        dt = m->h_t.as<float>();
    }
    m->model.decode(m->h_z.p, z_dtype, dt, B, F, H, W, m->h_out.p, out_dtype, postprocess, s);
    LTXV_CUDA(cudaMemcpyAsync(out, m->h_out.p, no, cudaMemcpyDeviceToHost, s));
    LTXV_CUDA(cudaStreamSynchronize(s));
    LTXV_CATCH
}

/* ------------------------------------------------- pipeline glue ------------------------------------------------ */
int ltxv_pack_latents(const float* in, float* out, int B, int C, int F, int H, int W, int p, int pt, void* stream) {
    LTXV_TRY
    if (in == nullptr || out == nullptr) fail("null argument");
    if (p <= 0 || pt <= 0 || F % pt != 0 || H % p != 0 || W % p != 0)
        fail("latents shape not divisible by patch sizes");  // t2v_pipeline.rs:486-488
    const int64_t n = static_cast<int64_t>(C) * F * H * W;
    for (int b = 0; b < B; ++b)
        LTXV_CUDA(launch_pack_latents(in + b * n, out + b * n, C, F, H, W, p, pt, static_cast<cudaStream_t>(stream)));
    LTXV_CATCH
}
int ltxv_unpack_latents(const float* in, float* out, int B, int C, int F, int H, int W, int p, int pt, void* stream) {
    LTXV_TRY
    if (in == nullptr || out == nullptr) fail("null argument");
    if (p <= 0 || pt <= 0 || F % pt != 0 || H % p != 0 || W % p != 0) fail("latents shape not divisible by patch sizes");
    const int64_t n = static_cast<int64_t>(C) * F * H * W;
    for (int b = 0; b < B; ++b)
        LTXV_CUDA(launch_unpack_latents(in + b * n, out + b * n, C, F, H, W, p, pt, static_cast<cudaStream_t>(stream)));
    LTXV_CATCH
}
int ltxv_video_coords(float* out, int B, int F, int H, int W, int ts_ratio, int sp_ratio, int fps, void* stream) {
    LTXV_TRY
    if (out == nullptr) fail("null argument");
    if (fps <= 0) fail("frame rate must be positive");
    const int64_t n = static_cast<int64_t>(F) * H * W * 3;
    for (int b = 0; b < B; ++b)
        LTXV_CUDA(launch_video_coords(out + b * n, F, H, W, ts_ratio, sp_ratio, fps, static_cast<cudaStream_t>(stream)));
    LTXV_CATCH
}
int ltxv_guidance_euler_step(const float* cond, const float* uncond, const float* perturbed, float* latents,
                             float* noise_pred_out, int B, int64_t n, float guidance_scale, float guidance_rescale,
                             float stg_scale, float sigma, float sigma_next, void* stream) {
    LTXV_TRY
    if (cond == nullptr) fail("null argument");
    static DevBuf scratch_dev[16];  // one per device: the statistics of the std rescale
    int dev = 0;
    LTXV_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16) fail("device index %d out of range", dev);
    DevBuf& scratch = scratch_dev[dev];
    scratch.ensure(64);
    const float dt = sigma_next - sigma;
    for (int b = 0; b < B; ++b)
        LTXV_CUDA(launch_guidance_euler(cond + b * n, uncond ? uncond + b * n : nullptr,
                                        perturbed ? perturbed + b * n : nullptr, latents ? latents + b * n : nullptr,
                                        noise_pred_out ? noise_pred_out + b * n : nullptr, n, guidance_scale,
                                        guidance_rescale, stg_scale, dt, scratch.as<double>(),
                                        static_cast<cudaStream_t>(stream)));
    LTXV_CATCH
}
int ltxv_scheduler_step_stochastic(float* latents, const float* model_output, const float* noise, int64_t n, float sigma,
                                   float sigma_next, void* stream) {
    LTXV_TRY
    if (latents == nullptr || model_output == nullptr || noise == nullptr) fail("null argument");
    LTXV_CUDA(launch_stochastic_step(latents, model_output, noise, sigma, sigma_next, n, static_cast<cudaStream_t>(stream)));
    LTXV_CATCH
}
int ltxv_decode_noise_blend(float* latents, const float* noise, float scale, int64_t n, void* stream) {
    LTXV_TRY
    if (latents == nullptr || noise == nullptr) fail("null argument");
    LTXV_CUDA(launch_noise_blend(latents, noise, scale, n, static_cast<cudaStream_t>(stream)));
    LTXV_CATCH
}
int ltxv_normalize_latents(const float* in, float* out, const float* mean, const float* std, float scaling_factor,
                           int B, int C, int64_t n_per_channel, void* stream) {
    LTXV_TRY
    if (in == nullptr || out == nullptr || mean == nullptr || std == nullptr) fail("null argument");
    LTXV_CUDA(launch_normalize_latents(in, mean, std, scaling_factor, out, B, C, n_per_channel,
                                       static_cast<cudaStream_t>(stream)));
    LTXV_CATCH
}
int ltxv_denormalize_latents(const float* in, float* out, const float* mean, const float* std, float scaling_factor,
                             int B, int C, int64_t n_per_channel, void* stream) {
    LTXV_TRY
    if (in == nullptr || out == nullptr || mean == nullptr || std == nullptr) fail("null argument");
    const int64_t n = C * n_per_channel;
    for (int b = 0; b < B; ++b)
        LTXV_CUDA(launch_denormalize(in + b * n, out + b * n, mean, std, 1.0f / scaling_factor, C, n_per_channel,
                                     static_cast<cudaStream_t>(stream)));
    LTXV_CATCH
}
int ltxv_postprocess_video(const float* in, float* out, int64_t n, void* stream) {
    LTXV_TRY
    if (in == nullptr || out == nullptr) fail("null argument");
    LTXV_CUDA(launch_postprocess(in, out, n, static_cast<cudaStream_t>(stream)));
    LTXV_CATCH
}
int ltxv_calculate_shift(int seq_len, float* mu_out) {
    LTXV_TRY
    if (mu_out == nullptr) fail("null argument");
    *mu_out = calculate_shift(seq_len);
    LTXV_CATCH
}
int ltxv_scheduler_set_timesteps(int num_steps, const float* custom_sigmas, float mu, int has_shift_terminal,
                                 float shift_terminal, float* sigmas_out, int64_t* timesteps_out) {
    LTXV_TRY
    if (sigmas_out == nullptr || timesteps_out == nullptr) fail("null argument");
    scheduler_set_timesteps(num_steps, custom_sigmas, mu, has_shift_terminal != 0, shift_terminal, sigmas_out,
                            timesteps_out);
    LTXV_CATCH
}

int ltxv_pipeline_denoise(ltxv_dit* dit, const ltxv_pipeline_params* p, float* latents, const void* prompt_embeds,
                          const float* prompt_mask, const void* negative_embeds, const float* negative_mask,
                          int embeds_dtype, int K, void* stream) {
    LTXV_TRY
    if (dit == nullptr || p == nullptr || latents == nullptr || prompt_embeds == nullptr) fail("null argument");
    pipeline_denoise(dit->model, *p, latents, prompt_embeds, prompt_mask, negative_embeds, negative_mask, embeds_dtype,
                     K, static_cast<cudaStream_t>(stream));
    LTXV_CATCH
}
int ltxv_pipeline_decode(ltxv_vae* vae, const ltxv_pipeline_params* p, const float* latents, float* out, void* stream) {
    LTXV_TRY
    if (vae == nullptr || p == nullptr || latents == nullptr || out == nullptr) fail("null argument");
    pipeline_decode(vae->model, *p, latents, out, static_cast<cudaStream_t>(stream));
    LTXV_CATCH
}

int ltxv_pipeline_denoise_stochastic(ltxv_dit* dit, const ltxv_pipeline_params* p, float* latents,
                                     const void* prompt_embeds, const float* prompt_mask, const void* negative_embeds,
                                     const float* negative_mask, int embeds_dtype, int K, const float* step_noise,
                                     void* stream) {
    LTXV_TRY
    if (dit == nullptr || p == nullptr || latents == nullptr || prompt_embeds == nullptr || step_noise == nullptr)
        fail("null argument");
    pipeline_denoise(dit->model, *p, latents, prompt_embeds, prompt_mask, negative_embeds, negative_mask, embeds_dtype, K,
                     static_cast<cudaStream_t>(stream), step_noise);
    LTXV_CATCH
}
int ltxv_pipeline_decode_noisy(ltxv_vae* vae, const ltxv_pipeline_params* p, const float* latents, const float* noise,
                               float decode_noise_scale, float* out, void* stream) {
    LTXV_TRY
    if (vae == nullptr || p == nullptr || latents == nullptr || out == nullptr) fail("null argument");
    if (noise == nullptr && decode_noise_scale != 0.0f)
        fail("decode_noise_scale = %g needs a noise tensor (the library has no RNG of its own)", decode_noise_scale);
    pipeline_decode(vae->model, *p, latents, out, static_cast<cudaStream_t>(stream), noise, decode_noise_scale);
    LTXV_CATCH
}

int ltxv_comm_create(int nranks, int rank, int device, uint64_t heap_bytes, ltxv_comm** out) {
    LTXV_TRY
    if (out == nullptr) fail("null argument");
    *out = new ltxv_comm(nranks, rank, device, static_cast<size_t>(heap_bytes));
    LTXV_CATCH
}
void ltxv_comm_destroy(ltxv_comm* c) { delete c; }
int ltxv_comm_get_handle(ltxv_comm* c, void* handle64) {
    LTXV_TRY
    if (c == nullptr || handle64 == nullptr) fail("null argument");
    c->comm.get_handle(handle64);
    LTXV_CATCH
}
int ltxv_comm_open(ltxv_comm* c, const void* all_handles) {
    LTXV_TRY
    if (c == nullptr || all_handles == nullptr) fail("null argument");
    c->comm.open_peers(all_handles);
    LTXV_CATCH
}
int ltxv_comm_barrier(ltxv_comm* c, void* stream) {
    LTXV_TRY
    if (c == nullptr) fail("null argument");
    c->comm.barrier(static_cast<cudaStream_t>(stream), 0);
    LTXV_CATCH
}
int ltxv_parallel_plan(int nranks, int rank, int S, int do_cfg, int32_t* out6) {
    LTXV_TRY
    if (out6 == nullptr) fail("null argument");
    if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) fail("invalid rank %d of %d", rank, nranks);
    const ParallelPlan pl = make_parallel_plan(nranks, rank, S, do_cfg != 0);
    out6[0] = pl.cfg_groups; out6[1] = pl.sp; out6[2] = pl.branch; out6[3] = pl.sp_rank; out6[4] = pl.s_local;
    out6[5] = pl.token0;
    LTXV_CATCH
}
int ltxv_pipeline_denoise_parallel(ltxv_dit* dit, ltxv_comm* c, const ltxv_pipeline_params* p, float* latents,
                                   const void* prompt_embeds, const float* prompt_mask, const void* negative_embeds,
                                   const float* negative_mask, int embeds_dtype, int K, void* stream) {
    LTXV_TRY
    if (dit == nullptr || c == nullptr || p == nullptr || latents == nullptr || prompt_embeds == nullptr)
        fail("null argument");
    pipeline_denoise_parallel(dit->model, c->comm, *p, latents, prompt_embeds, prompt_mask, negative_embeds,
                              negative_mask, embeds_dtype, K, static_cast<cudaStream_t>(stream));
    LTXV_CATCH
}
int ltxv_pipeline_denoise_parallel_stochastic(ltxv_dit* dit, ltxv_comm* c, const ltxv_pipeline_params* p, float* latents,
                                              const void* prompt_embeds, const float* prompt_mask,
                                              const void* negative_embeds, const float* negative_mask, int embeds_dtype,
                                              int K, const float* step_noise, void* stream) {
    LTXV_TRY
    if (dit == nullptr || c == nullptr || p == nullptr || latents == nullptr || prompt_embeds == nullptr ||
        step_noise == nullptr)
        fail("null argument");
    pipeline_denoise_parallel(dit->model, c->comm, *p, latents, prompt_embeds, prompt_mask, negative_embeds,
                              negative_mask, embeds_dtype, K, static_cast<cudaStream_t>(stream), step_noise);
    LTXV_CATCH
}
int ltxv_vae_set_comm(ltxv_vae* vae, ltxv_comm* c) {
    LTXV_TRY
    if (vae == nullptr) fail("null argument");
    vae->model.set_comm(c ? &c->comm : nullptr);
    LTXV_CATCH
}

int ltxv_pipeline_denoise_host(ltxv_dit* dit, const ltxv_pipeline_params* p, float* latents, const void* prompt_embeds,
                               const float* prompt_mask, const void* negative_embeds, const float* negative_mask,
                               int embeds_dtype, int K) {
    LTXV_TRY
    if (dit == nullptr || p == nullptr || latents == nullptr || prompt_embeds == nullptr) fail("null argument");
    const ltxv_dit_config& c = dit->model.config();
    const int F = (p->num_frames - 1) / 8 + 1, H = p->height / 32, W = p->width / 32;
    const size_t nl = static_cast<size_t>(F) * H * W * c.in_channels * 4;
    const size_t ne = static_cast<size_t>(K) * c.caption_channels * dsize(embeds_dtype);
    cudaStream_t s = 0;
    LTXV_CUDA(cudaSetDevice(dit->model.device()));
    DevBuf &d_lat = dit->p_lat, &d_pe = dit->p_pe, &d_ne = dit->p_ne, &d_pm = dit->p_pm, &d_nm = dit->p_nm;
    d_lat.ensure(nl);
    d_pe.ensure(ne);
    LTXV_CUDA(cudaMemcpyAsync(d_lat.p, latents, nl, cudaMemcpyHostToDevice, s));
    LTXV_CUDA(cudaMemcpyAsync(d_pe.p, prompt_embeds, ne, cudaMemcpyHostToDevice, s));
    const float *pm = nullptr, *nm = nullptr;
    const void* nep = nullptr;
    if (prompt_mask != nullptr) {
        d_pm.ensure(K * 4);
        LTXV_CUDA(cudaMemcpyAsync(d_pm.p, prompt_mask, K * 4, cudaMemcpyHostToDevice, s));
        pm = d_pm.as<float>();
    }
    if (negative_embeds != nullptr) {
        d_ne.ensure(ne);
        LTXV_CUDA(cudaMemcpyAsync(d_ne.p, negative_embeds, ne, cudaMemcpyHostToDevice, s));
        nep = d_ne.p;
    }
    if (negative_mask != nullptr) {
        d_nm.ensure(K * 4);
        LTXV_CUDA(cudaMemcpyAsync(d_nm.p, negative_mask, K * 4, cudaMemcpyHostToDevice, s));
        nm = d_nm.as<float>();
    }
    pipeline_denoise(dit->model, *p, d_lat.as<float>(), d_pe.p, pm, nep, nm, embeds_dtype, K, s);
    LTXV_CUDA(cudaMemcpyAsync(latents, d_lat.p, nl, cudaMemcpyDeviceToHost, s));
    LTXV_CUDA(cudaStreamSynchronize(s));
    LTXV_CATCH
}
int ltxv_pipeline_decode_host(ltxv_vae* vae, const ltxv_pipeline_params* p, const float* latents, float* out) {
    LTXV_TRY
    if (vae == nullptr || p == nullptr || latents == nullptr || out == nullptr) fail("null argument");
    const int C = vae->model.config().latent_channels;
    const int F = (p->num_frames - 1) / 8 + 1, H = p->height / 32, W = p->width / 32;
    const size_t nl = static_cast<size_t>(F) * H * W * C * 4;
    const size_t no = 3ull * (8 * F - 7) * (32 * H) * (32 * W) * 4;
    cudaStream_t s = 0;
    LTXV_CUDA(cudaSetDevice(vae->model.device()));
    DevBuf &d_lat = vae->p_lat, &d_out = vae->p_out;
    d_lat.ensure(nl);
    d_out.ensure(no);
    LTXV_CUDA(cudaMemcpyAsync(d_lat.p, latents, nl, cudaMemcpyHostToDevice, s));
    pipeline_decode(vae->model, *p, d_lat.as<float>(), d_out.as<float>(), s);
    LTXV_CUDA(cudaMemcpyAsync(out, d_out.p, no, cudaMemcpyDeviceToHost, s));
    LTXV_CUDA(cudaStreamSynchronize(s));
    LTXV_CATCH
}

int ltxv_frames_to_u8(const float* frames, uint8_t* out, int B, int F, int H, int W, void* stream) {
    LTXV_TRY
    if (frames == nullptr || out == nullptr) fail("null argument");
    if (B <= 0 || F <= 0 || H <= 0 || W <= 0) fail("invalid frame extent [%d,3,%d,%d,%d]", B, F, H, W);
    LTXV_CUDA(launch_frames_to_u8(frames, out, B, F, H, W, static_cast<cudaStream_t>(stream)));
    LTXV_CATCH
}
int ltxv_pipeline_decode_host_u8(ltxv_vae* vae, const ltxv_pipeline_params* p, const float* latents, uint8_t* out) {
    LTXV_TRY
    if (vae == nullptr || p == nullptr || latents == nullptr || out == nullptr) fail("null argument");
    const int C = vae->model.config().latent_channels;
    const int F = (p->num_frames - 1) / 8 + 1, H = p->height / 32, W = p->width / 32;
    const int Fo = 8 * F - 7, Ho = 32 * H, Wo = 32 * W;
    const size_t nl = static_cast<size_t>(F) * H * W * C * 4;
    const size_t npx = static_cast<size_t>(Fo) * Ho * Wo;
    cudaStream_t s = 0;
    LTXV_CUDA(cudaSetDevice(vae->model.device()));
    DevBuf &d_lat = vae->p_lat, &d_out = vae->p_out, &d_u8 = vae->p_u8;
    d_lat.ensure(nl);
    d_out.ensure(npx * 3 * 4);
    d_u8.ensure(npx * 3);
    LTXV_CUDA(cudaMemcpyAsync(d_lat.p, latents, nl, cudaMemcpyHostToDevice, s));
    pipeline_decode(vae->model, *p, d_lat.as<float>(), d_out.as<float>(), s);
    LTXV_CUDA(launch_frames_to_u8(d_out.as<float>(), d_u8.as<uint8_t>(), 1, Fo, Ho, Wo, s));
    LTXV_CUDA(cudaMemcpyAsync(out, d_u8.p, npx * 3, cudaMemcpyDeviceToHost, s));
    LTXV_CUDA(cudaStreamSynchronize(s));
    LTXV_CATCH
}

}  // extern "C"
