// Weight ingestion: safetensors reader, official->diffusers key remap, shard index.  See weights.cc.
#pragma once
#include <stdint.h>

#include <functional>
#include <string>
#include <vector>

namespace ltxv {

enum WeightComponent : int { WEIGHT_OTHER = 0, WEIGHT_TRANSFORMER = 1, WEIGHT_VAE = 2 };

// KeyRemapper::remap_key (weight_format.rs:55-143)
std::string remap_official_key(const std::string& key);
// KeyRemapper::is_vae_key / is_transformer_key in the order main.rs:482-489 applies them
int classify_official_key(const std::string& key);
// remap + prefix strip ("vae." | "model.diffusion_model." | "transformer."), main.rs:480-497
std::string official_key_to_model_key(const std::string& key, int* component);

struct SafeTensorInfo {
    std::string name, dtype;
    std::vector<int64_t> shape;
    const uint8_t* data = nullptr;  // into the mapping
    size_t bytes = 0;
};

class SafeTensorsFile {
public:
    explicit SafeTensorsFile(const std::string& path);
    ~SafeTensorsFile();
    SafeTensorsFile(const SafeTensorsFile&) = delete;
    SafeTensorsFile& operator=(const SafeTensorsFile&) = delete;
    const std::vector<SafeTensorInfo>& tensors() const { return tensors_; }  // sorted by name

private:
    void parse(const std::string& path);
    std::string path_;
    int fd_ = -1;
    const uint8_t* map_ = nullptr;
    const uint8_t* data_ = nullptr;
    size_t size_ = 0;
    std::vector<SafeTensorInfo> tensors_;
};

// one .safetensors file | directory with a single file | directory with `*.safetensors.index.json` + shards
std::vector<std::string> resolve_safetensors_files(const std::string& path);

struct WeightSink {
    std::function<bool(const std::string&)> wants;  // does the model have a slot for this (diffusers) key?
    std::function<void(const std::string&, const void*, int, const int64_t*, int)> load;
};
void load_safetensors(const std::string& path, bool official, int component, const WeightSink& sink, int* n_loaded,
                      int* n_ignored);

}  // namespace ltxv
