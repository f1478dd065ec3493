#include "options.h"

#include <stdlib.h>
#include <string.h>

namespace ltxv {
namespace {
struct Entry {
    const char* name;
    const char* env;
    int Options::*field;
    bool numeric;  // value = atoi(env) instead of "set / unset"
};
const Entry kEntries[] = {
    {"no_cfg_batch", "LTXV_NO_CFG_BATCH", &Options::no_cfg_batch, false},
    {"gemm_no_pair", "LTXV_GEMM_NO_PAIR", &Options::gemm_no_pair, false},
    {"conv_no_kw3", "LTXV_CONV_NO_KW3", &Options::conv_no_kw3, false},
    {"gemm_no_raster", "LTXV_GEMM_NO_RASTER", &Options::gemm_no_raster, false},
    {"gemm_no_epi2", "LTXV_GEMM_NO_EPI2", &Options::gemm_no_epi2, false},
    {"gemm_k2", "LTXV_GEMM_K2", &Options::gemm_k2, false},
    {"gemm_no_short_k", "LTXV_GEMM_NO_SHORT_K_RULE", &Options::gemm_no_short_k, false},
    {"attn_v1", "LTXV_ATTN_V1", &Options::attn_v1, false},
    {"attn_v4", "LTXV_ATTN_V4", &Options::attn_v4, false},
    {"attn_nosplit", "LTXV_ATTN_NOSPLIT", &Options::attn_nosplit, false},
    {"attn_nsplit_max", "LTXV_ATTN_NSPLIT", &Options::attn_nsplit_max, true},
    {"vae_no_fused_prep", "LTXV_VAE_NO_FUSED_PREP", &Options::vae_no_fused_prep, false},
    {"vae_no_fuse_conv2", "LTXV_VAE_NO_FUSE_CONV2", &Options::vae_no_fuse_conv2, false},
    {"vae_fuse_conv2", "LTXV_VAE_FUSE_CONV2", &Options::vae_fuse_conv2, false},
    {"vae_prep_u", "LTXV_VAE_PREP_U", &Options::vae_prep_u, true},
    {"no_pdl", "LTXV_NO_PDL", &Options::no_pdl, true},
    {"qk_unfused", "LTXV_QK_UNFUSED", &Options::qk_unfused, false},
};
Options from_env() {
    Options o{};
    for (const Entry& e : kEntries) {
        const char* v = getenv(e.env);
        if (v == nullptr) continue;
        o.*(e.field) = e.numeric ? atoi(v) : 1;
    }
    return o;
}
}  // namespace

Options& options() {
    static Options o = from_env();
    return o;
}

bool set_option(const char* name, int value) {
    for (const Entry& e : kEntries)
        if (strcmp(e.name, name) == 0) {
            options().*(e.field) = value;
            return true;
        }
    return false;
}

bool get_option(const char* name, int* value) {
    for (const Entry& e : kEntries)
        if (strcmp(e.name, name) == 0) {
            *value = options().*(e.field);
            return true;
        }
    return false;
}

}  // namespace ltxv
