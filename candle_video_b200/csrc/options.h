// Experiment knobs.  Each knob is initialised ONCE per process from its environment variable and can be
// overridden at run time through ltxv_set_option (tests toggle kernel variants inside one process that way).  Launch
// paths read plain ints: no getenv on a hot path.
#pragma once

namespace ltxv {

struct Options {
    int no_cfg_batch;        // LTXV_NO_CFG_BATCH: CFG branches as two sequential forwards (the reference's order)
    int gemm_no_pair;        // LTXV_GEMM_NO_PAIR: no CTA-pair GEMM / conv kernels
    int conv_no_kw3;         // LTXV_CONV_NO_KW3: no three-taps-per-step conv mode
    int gemm_no_raster;      // LTXV_GEMM_NO_RASTER: row-fastest tile order everywhere (no L2-friendly grouping)
    int gemm_no_epi2;        // LTXV_GEMM_NO_EPI2: no two-epilogue-warpgroup pair kernel for the short-K residual GEMMs
    int gemm_k2;             // LTXV_GEMM_K2: two k-blocks per stage (opt-in)
    int gemm_no_short_k;     // LTXV_GEMM_NO_SHORT_K_RULE: no 128x192 preference for the short-K N = K = 2048 projections
    int attn_v1;             // LTXV_ATTN_V1: general attention kernel everywhere
    int attn_v4;             // LTXV_ATTN_V4: half a score row per softmax thread (flash_attn4_kernel; measured slower)
    int attn_nosplit;        // LTXV_ATTN_NOSPLIT: no key-range tail splitting
    int attn_nsplit_max;     // LTXV_ATTN_NSPLIT=n: cap on key ranges per tail unit (0 = no cap)
    int vae_prep_u;          // LTXV_VAE_PREP_U=n: voxel passes in flight per thread in vae_prep_kernel<128> (0 = default)
    int vae_no_fused_prep;   // LTXV_VAE_NO_FUSED_PREP: no fused producer epilogue at all
    int vae_no_fuse_conv2;   // LTXV_VAE_NO_FUSE_CONV2: ... not for conv2 at C = 256
    int vae_fuse_conv2;      // LTXV_VAE_FUSE_CONV2: ... also for conv2 at C = 128 (slower)
    int no_pdl;              // LTXV_NO_PDL: plain launches instead of programmatic dependent launch
    int qk_unfused;          // LTXV_QK_UNFUSED: separate q/k norm + RoPE pass instead of the fused QKV epilogue
};

Options& options();
// returns false for an unknown name
bool set_option(const char* name, int value);
bool get_option(const char* name, int* value);

}  // namespace ltxv
