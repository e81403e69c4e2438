// gguf_file.h — minimal read-only GGUF v2/v3 parser (mmap).
// Replaces the reference's use of ggml's gguf_init_from_file / gguf_get_tensor_* for the LM weights
// (reference src/loader.h:85-99, 235-271).  Key/value metadata is skipped (the reference writes and
// reads none, loader.h:227-233) but parsed correctly so files written by gguf-py also load.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace msx {

enum GgmlType : int { T_F32 = 0, T_F16 = 1, T_Q4_0 = 2, T_Q8_0 = 8, T_Q4_K = 12, T_BF16 = 30 };

// bytes of one row of k elements; -1 if unsupported type / k not a block multiple
int64_t ggml_row_size(int type, int64_t k);
const char *ggml_type_name(int type);

struct GgufTensor {
    std::string name;
    int type = 0;
    int n_dims = 0;
    int64_t ne[4] = {1, 1, 1, 1};   // ggml order: ne[0] fastest (= in-features K of a linear)
    uint64_t offset = 0;            // from start of data section
    const uint8_t *data = nullptr;  // into the mmap
    int64_t nbytes = 0;
};

class GgufFile {
public:
    GgufFile() = default;
    ~GgufFile();
    GgufFile(const GgufFile &) = delete;
    GgufFile &operator=(const GgufFile &) = delete;
    // returns false and fills err on failure
    bool open(const std::string &path, std::string &err);
    const GgufTensor *find(const std::string &name) const;
    const std::vector<GgufTensor> &tensors() const { return tensors_; }
    int version() const { return version_; }

private:
    int fd_ = -1;
    const uint8_t *map_ = nullptr;
    size_t size_ = 0;
    int version_ = 0;
    std::vector<GgufTensor> tensors_;
    std::unordered_map<std::string, size_t> index_;
};

}  // namespace msx
