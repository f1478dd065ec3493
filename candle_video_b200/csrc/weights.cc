// Weight ingestion for the drop-in boundary: safetensors files read directly (mmap + header JSON),
// diffusers-format key names as consumed by the reference's VarBuilder paths, the official single-file key remap
// (weight_format.rs:55-143, applied as main.rs:480-497 does), and sharded checkpoints through their index JSON
// (loader.rs:341-371).  Host-only code: tensors are handed to the models' load_tensor(), which does the cast-on-load
// and the one-time re-layouts (fused QKV, implicit-GEMM conv weights).
#include "weights.h"

#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <set>

#include "model_common.h"

namespace ltxv {

// ------------------------------------------------------------------------------------------------
// key remap (official unified file -> diffusers names)
// ------------------------------------------------------------------------------------------------
namespace {
void replace_all(std::string& s, const std::string& from, const std::string& to) {
    if (from.empty()) return;
    size_t pos = 0;
    while ((pos = s.find(from, pos)) != std::string::npos) {
        s.replace(pos, from.size(), to);
        pos += to.size();
    }
}

// `<prefix>.<digits>` -> table[digits] (weight_format.rs:79-141); indices outside the table keep `<prefix>.<idx>`
void remap_block_index(std::string& s, const std::string& prefix, const char* const* table, int n_table) {
    const std::string pat = prefix + ".";
    size_t pos = 0;
    std::string out;
    while (true) {
        const size_t hit = s.find(pat, pos);
        if (hit == std::string::npos) break;
        size_t d = hit + pat.size(), e = d;
        while (e < s.size() && s[e] >= '0' && s[e] <= '9') ++e;
        if (e == d) {  // no digits: not a block reference
            out.append(s, pos, hit + pat.size() - pos);
            pos = hit + pat.size();
            continue;
        }
        const long idx = strtol(s.substr(d, e - d).c_str(), nullptr, 10);
        out.append(s, pos, hit - pos);
        if (idx >= 0 && idx < n_table) out += table[idx];
        else out += prefix + "." + std::to_string(idx);
        pos = e;
    }
    out.append(s, pos, std::string::npos);
    s.swap(out);
}
}  // namespace

std::string remap_official_key(const std::string& key) {
    std::string r = key;
    // 1. transformer (weight_format.rs:58-62)
    replace_all(r, "patchify_proj", "proj_in");
    replace_all(r, "adaln_single", "time_embed");
    replace_all(r, "q_norm", "norm_q");
    replace_all(r, "k_norm", "norm_k");
    // 2. VAE (weight_format.rs:64-76)
    replace_all(r, "res_blocks", "resnets");
    static const char* const enc[9] = {"encoder.down_blocks.0", "encoder.down_blocks.0.downsamplers.0",
                                       "encoder.down_blocks.1", "encoder.down_blocks.1.downsamplers.0",
                                       "encoder.down_blocks.2", "encoder.down_blocks.2.downsamplers.0",
                                       "encoder.down_blocks.3", "encoder.down_blocks.3.downsamplers.0",
                                       "encoder.mid_block"};
    static const char* const dec[9] = {"decoder.mid_block",     "decoder.up_blocks.0.upsamplers.0",
                                       "decoder.up_blocks.0",   "decoder.up_blocks.1.upsamplers.0",
                                       "decoder.up_blocks.1",   "decoder.up_blocks.2.upsamplers.0",
                                       "decoder.up_blocks.2",   "decoder.up_blocks.3.upsamplers.0",
                                       "decoder.up_blocks.3"};
    remap_block_index(r, "encoder.down_blocks", enc, 9);
    remap_block_index(r, "decoder.up_blocks", dec, 9);
    replace_all(r, "last_time_embedder", "time_embedder");
    replace_all(r, "last_scale_shift_table", "scale_shift_table");
    replace_all(r, "norm3.norm", "norm3");
    replace_all(r, "per_channel_statistics.mean-of-means", "latents_mean");
    replace_all(r, "per_channel_statistics.std-of-means", "latents_std");
    return r;
}

static bool starts_with(const std::string& s, const char* p) { return s.rfind(p, 0) == 0; }
static bool contains(const std::string& s, const char* p) { return s.find(p) != std::string::npos; }

// weight_format.rs:145-166 (evaluated on the ORIGINAL key, VAE first, as main.rs:482-489 does)
int classify_official_key(const std::string& key) {
    if (starts_with(key, "vae.") || starts_with(key, "encoder.") || starts_with(key, "decoder.") ||
        contains(key, "per_channel_statistics") || contains(key, "latents_mean") || contains(key, "latents_std"))
        return WEIGHT_VAE;
    if (starts_with(key, "transformer.") || starts_with(key, "model.diffusion_model.") ||
        contains(key, "transformer_blocks") || contains(key, "patchify_proj") || contains(key, "proj_in") ||
        contains(key, "adaln_single") || contains(key, "time_embed"))
        return WEIGHT_TRANSFORMER;
    return WEIGHT_OTHER;
}

std::string official_key_to_model_key(const std::string& key, int* component) {
    const int c = classify_official_key(key);
    if (component) *component = c;
    std::string r = remap_official_key(key);
    if (c == WEIGHT_VAE) {
        if (starts_with(r, "vae.")) r = r.substr(4);
    } else if (c == WEIGHT_TRANSFORMER) {
        if (starts_with(r, "model.diffusion_model.")) r = r.substr(22);
        else if (starts_with(r, "transformer.")) r = r.substr(12);
    }
    return r;
}

// ------------------------------------------------------------------------------------------------
// minimal JSON reader for the two documents we meet: the safetensors header and the shard index
// ------------------------------------------------------------------------------------------------
namespace {
struct Json {
    const char* p;
    const char* end;
    void ws() {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p;
    }
    bool eat(char c) {
        ws();
        if (p < end && *p == c) {
            ++p;
            return true;
        }
        return false;
    }
    void expect(char c) {
        if (!eat(c)) fail("malformed JSON: expected '%c' (%ld bytes before the end)", c, static_cast<long>(end - p));
    }
    std::string str() {
        ws();
        if (p >= end || *p != '"') fail("malformed JSON: expected a string");
        ++p;
        std::string out;
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) {
                ++p;
                switch (*p) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {  // keys/dtypes are ASCII; keep the escape verbatim for anything else
                        out += "\\u";
                        break;
                    }
                    default: out += *p;
                }
                ++p;
            } else {
                out += *p++;
            }
        }
        if (p >= end) fail("malformed JSON: unterminated string");
        ++p;
        return out;
    }
    // bounded by [p, end): the header is a slice of an mmap'd file and is NOT NUL-terminated (strtoll would run past it)
    int64_t integer() {
        ws();
        bool neg = false;
        if (p < end && (*p == '-' || *p == '+')) neg = *p++ == '-';
        if (p >= end || *p < '0' || *p > '9') fail("malformed JSON: expected an integer");
        int64_t v = 0;
        while (p < end && *p >= '0' && *p <= '9') {
            if (__builtin_mul_overflow(v, static_cast<int64_t>(10), &v) ||
                __builtin_add_overflow(v, static_cast<int64_t>(*p - '0'), &v))
                fail("malformed JSON: integer out of range");
            ++p;
        }
        return neg ? -v : v;
    }
    void skip_value() {
        ws();
        if (p >= end) fail("malformed JSON: truncated");
        if (*p == '"') {
            str();
        } else if (*p == '{') {
            ++p;
            if (eat('}')) return;
            do {
                str();
                expect(':');
                skip_value();
            } while (eat(','));
            expect('}');
        } else if (*p == '[') {
            ++p;
            if (eat(']')) return;
            do skip_value();
            while (eat(','));
            expect(']');
        } else {
            while (p < end && *p != ',' && *p != '}' && *p != ']') ++p;
        }
    }
};

size_t dtype_size(const std::string& d) {
    if (d == "F32" || d == "I32" || d == "U32") return 4;
    if (d == "BF16" || d == "F16" || d == "I16" || d == "U16") return 2;
    if (d == "F64" || d == "I64" || d == "U64") return 8;
    if (d == "I8" || d == "U8" || d == "BOOL" || d == "F8_E4M3" || d == "F8_E5M2") return 1;
    fail("unsupported safetensors dtype '%s'", d.c_str());
}

float half_to_float(uint16_t h) {
    const uint32_t sign = (h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ffu, bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else {
            int e = -1;
            do {
                man <<= 1;
                ++e;
            } while ((man & 0x400u) == 0);
            bits = sign | ((127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

std::string read_text_file(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) fail("cannot open '%s'", path.c_str());
    std::string s;
    char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
    fclose(f);
    return s;
}
bool file_exists(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}
bool dir_exists(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
}  // namespace

// ------------------------------------------------------------------------------------------------
// SafeTensorsFile
// ------------------------------------------------------------------------------------------------
SafeTensorsFile::SafeTensorsFile(const std::string& path) : path_(path) {
    // a constructor that throws does not run the destructor: release the mapping / descriptor before re-throwing
    try {
        parse(path);
    } catch (...) {
        if (map_) munmap(const_cast<uint8_t*>(map_), size_);
        if (fd_ >= 0) close(fd_);
        map_ = nullptr;
        fd_ = -1;
        throw;
    }
}

void SafeTensorsFile::parse(const std::string& path) {
    fd_ = open(path.c_str(), O_RDONLY);
    if (fd_ < 0) fail("cannot open safetensors file '%s'", path.c_str());
    struct stat st;
    if (fstat(fd_, &st) != 0 || st.st_size < 8) fail("'%s' is too short to be a safetensors file", path.c_str());
    size_ = static_cast<size_t>(st.st_size);
    map_ = static_cast<const uint8_t*>(mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0));
    if (map_ == MAP_FAILED) {
        map_ = nullptr;
        fail("mmap of '%s' failed", path.c_str());
    }
    uint64_t hlen = 0;
    memcpy(&hlen, map_, 8);  // little-endian u64
    if (hlen > size_ - 8 || hlen > (100ull << 20)) fail("'%s': invalid safetensors header length %llu", path.c_str(),
                                                         static_cast<unsigned long long>(hlen));
    data_ = map_ + 8 + hlen;
    const size_t data_bytes = size_ - 8 - static_cast<size_t>(hlen);
    Json j{reinterpret_cast<const char*>(map_ + 8), reinterpret_cast<const char*>(map_ + 8 + hlen)};
    j.expect('{');
    if (!j.eat('}')) {
        do {
            const std::string name = j.str();
            j.expect(':');
            if (name == "__metadata__") {
                j.skip_value();
                continue;
            }
            SafeTensorInfo info;
            info.name = name;
            uint64_t begin = 0, endo = 0;
            j.expect('{');
            do {
                const std::string k = j.str();
                j.expect(':');
                if (k == "dtype") {
                    info.dtype = j.str();
                } else if (k == "shape") {
                    j.expect('[');
                    if (!j.eat(']')) {
                        do info.shape.push_back(j.integer());
                        while (j.eat(','));
                        j.expect(']');
                    }
                } else if (k == "data_offsets") {
                    j.expect('[');
                    begin = static_cast<uint64_t>(j.integer());
                    j.expect(',');
                    endo = static_cast<uint64_t>(j.integer());
                    j.expect(']');
                } else {
                    j.skip_value();
                }
            } while (j.eat(','));
            j.expect('}');
            int64_t numel = 1, nbytes = 0;
            for (auto d : info.shape)
                if (d < 0 || __builtin_mul_overflow(numel, d, &numel))
                    fail("'%s': tensor '%s' has a negative or overflowing shape", path.c_str(), name.c_str());
            if (__builtin_mul_overflow(numel, static_cast<int64_t>(dtype_size(info.dtype)), &nbytes))
                fail("'%s': tensor '%s' is too large", path.c_str(), name.c_str());
            if (endo < begin || endo > data_bytes || (endo - begin) != static_cast<uint64_t>(nbytes))
                fail("'%s': tensor '%s' has inconsistent offsets [%llu,%llu) for %lld x %s", path.c_str(), name.c_str(),
                     static_cast<unsigned long long>(begin), static_cast<unsigned long long>(endo),
                     static_cast<long long>(numel), info.dtype.c_str());
            info.data = data_ + begin;
            info.bytes = static_cast<size_t>(endo - begin);
            tensors_.push_back(std::move(info));
        } while (j.eat(','));
        j.expect('}');
    }
    // the format requires the tensors to tile the data section: no overlap, no hole (safetensors README, "Format")
    std::vector<std::pair<const uint8_t*, size_t>> spans;
    for (const SafeTensorInfo& t : tensors_) spans.emplace_back(t.data, t.bytes);
    std::sort(spans.begin(), spans.end());
    const uint8_t* cursor = data_;
    for (const auto& sp : spans) {
        if (sp.first != cursor)
            fail("'%s': tensor data ranges %s", path.c_str(), sp.first < cursor ? "overlap" : "leave a hole in the data section");
        cursor = sp.first + sp.second;
    }
    if (cursor != data_ + data_bytes) fail("'%s': %zu trailing bytes after the last tensor", path.c_str(),
                                           static_cast<size_t>(data_ + data_bytes - cursor));
    std::sort(tensors_.begin(), tensors_.end(),
              [](const SafeTensorInfo& a, const SafeTensorInfo& b) { return a.name < b.name; });
}

SafeTensorsFile::~SafeTensorsFile() {
    if (map_) munmap(const_cast<uint8_t*>(map_), size_);
    if (fd_ >= 0) close(fd_);
}

// ------------------------------------------------------------------------------------------------
// file set resolution: one file, a directory with one file, or a directory with an index + shards
// ------------------------------------------------------------------------------------------------
std::vector<std::string> resolve_safetensors_files(const std::string& path) {
    if (file_exists(path)) return {path};
    if (!dir_exists(path)) fail("'%s' is neither a safetensors file nor a directory", path.c_str());
    static const char* const index_names[] = {"model.safetensors.index.json",
                                              "diffusion_pytorch_model.safetensors.index.json"};
    for (const char* n : index_names) {
        const std::string ip = path + "/" + n;
        if (!file_exists(ip)) continue;
        const std::string text = read_text_file(ip);
        Json j{text.data(), text.data() + text.size()};
        std::set<std::string> files;
        j.expect('{');
        do {
            const std::string k = j.str();
            j.expect(':');
            if (k == "weight_map") {
                j.expect('{');
                if (!j.eat('}')) {
                    do {
                        j.str();
                        j.expect(':');
                        files.insert(j.str());
                    } while (j.eat(','));
                    j.expect('}');
                }
            } else {
                j.skip_value();
            }
        } while (j.eat(','));
        if (files.empty()) fail("'%s' has an empty weight_map", ip.c_str());
        std::vector<std::string> out;
        for (auto& f : files) {
            const std::string fp = path + "/" + f;
            if (!file_exists(fp)) fail("shard '%s' named by '%s' is missing", fp.c_str(), ip.c_str());  // strict mode
            out.push_back(fp);
        }
        return out;
    }
    static const char* const single_names[] = {"diffusion_pytorch_model.safetensors", "model.safetensors"};
    for (const char* n : single_names)
        if (file_exists(path + "/" + n)) return {path + "/" + n};
    fail("no safetensors weights found in directory '%s'", path.c_str());
}

// ------------------------------------------------------------------------------------------------
// ingestion
// ------------------------------------------------------------------------------------------------
void load_safetensors(const std::string& path, bool official, int component, const WeightSink& sink, int* n_loaded,
                      int* n_ignored) {
    int loaded = 0, ignored = 0;
    std::vector<float> scratch;
    for (const std::string& file : resolve_safetensors_files(path)) {
        SafeTensorsFile st(file);
        for (const SafeTensorInfo& t : st.tensors()) {
            std::string key = t.name;
            if (official) {
                int c = WEIGHT_OTHER;
                key = official_key_to_model_key(t.name, &c);
                if (c != component) {
                    ++ignored;
                    continue;
                }
            }
            if (!sink.wants(key)) {
                ++ignored;
                continue;
            }
            int dtype;
            const void* data = t.data;
            if (t.dtype == "F32") {
                dtype = LTXV_F32;
            } else if (t.dtype == "BF16") {
                dtype = LTXV_BF16;
            } else if (t.dtype == "F16") {  // cast on load, like VarBuilder with an explicit dtype (main.rs:507-508)
                const size_t n = t.bytes / 2;
                scratch.resize(n);
                const uint16_t* h = reinterpret_cast<const uint16_t*>(t.data);
                for (size_t i = 0; i < n; ++i) scratch[i] = half_to_float(h[i]);
                data = scratch.data();
                dtype = LTXV_F32;
            } else {
                fail("tensor '%s' in '%s' has dtype %s; F32, BF16 or F16 expected", t.name.c_str(), file.c_str(),
                     t.dtype.c_str());
            }
            sink.load(key, data, dtype, t.shape.data(), static_cast<int>(t.shape.size()));
            ++loaded;
        }
    }
    if (n_loaded) *n_loaded = loaded;
    if (n_ignored) *n_ignored = ignored;
}

}  // namespace ltxv
