// gguf_file.cpp — see gguf_file.h
#include "gguf_file.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstring>

namespace msx {

int64_t ggml_row_size(int type, int64_t k) {
    switch (type) {
        case T_F32: case 26 /* I32: PersonaPlex voice.cache */: return 4 * k;
        case T_F16: case T_BF16: return 2 * k;
        case T_Q4_0: return (k % 32) ? -1 : k / 32 * 18;
        case T_Q8_0: return (k % 32) ? -1 : k / 32 * 34;
        case T_Q4_K: return (k % 256) ? -1 : k / 256 * 144;
        default: return -1;
    }
}

const char *ggml_type_name(int type) {
    switch (type) {
        case T_F32: return "f32"; case T_F16: return "f16"; case T_BF16: return "bf16";
        case T_Q4_0: return "q4_0"; case T_Q8_0: return "q8_0"; case T_Q4_K: return "q4_k";
        default: return "unsupported";
    }
}

GgufFile::~GgufFile() {
    if (map_) munmap(const_cast<uint8_t *>(map_), size_);
    if (fd_ >= 0) close(fd_);
}

namespace {
struct Cursor {
    const uint8_t *p, *end;
    bool ok = true;
    template <typename T> T get() {
        T v{};
        if (p + sizeof(T) > end) { ok = false; return v; }
        memcpy(&v, p, sizeof(T)); p += sizeof(T); return v;
    }
    std::string str() {
        uint64_t n = get<uint64_t>();
        if (!ok || n > (uint64_t)(end - p)) { ok = false; return {}; }
        std::string s((const char *)p, (size_t)n); p += n; return s;
    }
    void skip(uint64_t n) { if (n > (uint64_t)(end - p)) ok = false; else p += n; }
};

// GGUF metadata value types
enum { V_U8, V_I8, V_U16, V_I16, V_U32, V_I32, V_F32, V_BOOL, V_STR, V_ARR, V_U64, V_I64, V_F64 };
size_t scalar_size(uint32_t t) {
    switch (t) {
        case V_U8: case V_I8: case V_BOOL: return 1;
        case V_U16: case V_I16: return 2;
        case V_U32: case V_I32: case V_F32: return 4;
        case V_U64: case V_I64: case V_F64: return 8;
        default: return 0;
    }
}
bool skip_value(Cursor &c, uint32_t t, uint32_t *align_out, bool is_align) {
    if (t == V_STR) { c.str(); return c.ok; }
    if (t == V_ARR) {
        uint32_t et = c.get<uint32_t>(); uint64_t n = c.get<uint64_t>();
        if (!c.ok) return false;
        if (et == V_STR) { for (uint64_t i = 0; i < n && c.ok; i++) c.str(); return c.ok; }
        size_t es = scalar_size(et); if (!es) return false;
        if (n > (uint64_t)(c.end - c.p) / es) { c.ok = false; return false; }      // n * es cannot wrap
        c.skip(n * es); return c.ok;
    }
    size_t s = scalar_size(t); if (!s) return false;
    if (is_align && t == V_U32) { *align_out = c.get<uint32_t>(); return c.ok; }
    c.skip(s); return c.ok;
}
}  // namespace

bool GgufFile::open(const std::string &path, std::string &err) {
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) { err = "cannot open " + path; return false; }
    struct stat st;
    if (fstat(fd_, &st) != 0 || st.st_size < 24) { err = "cannot stat / file too small: " + path; return false; }
    size_ = (size_t)st.st_size;
    void *m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
    if (m == MAP_FAILED) { err = "mmap failed: " + path; map_ = nullptr; return false; }
    map_ = (const uint8_t *)m;
    madvise(m, size_, MADV_SEQUENTIAL);

    Cursor c{map_, map_ + size_};
    if (c.get<uint32_t>() != 0x46554747u) { err = "not a GGUF file (bad magic): " + path; return false; }
    version_ = (int)c.get<uint32_t>();
    if (version_ != 2 && version_ != 3) { err = "unsupported GGUF version " + std::to_string(version_); return false; }
    uint64_t n_tensors = c.get<uint64_t>(), n_kv = c.get<uint64_t>();
    // header counts come from the (downloaded) file: every key / tensor record takes at least 12 / 24 bytes of it
    if (!c.ok || n_kv > size_ / 12 || n_tensors > size_ / 24) { err = "corrupt GGUF header (tensor / key counts exceed the file size)"; return false; }
    uint32_t alignment = 32;
    for (uint64_t i = 0; i < n_kv; i++) {
        std::string key = c.str();
        uint32_t t = c.get<uint32_t>();
        if (!c.ok || !skip_value(c, t, &alignment, key == "general.alignment")) { err = "corrupt GGUF metadata"; return false; }
    }
    if (alignment == 0 || (alignment & (alignment - 1))) { err = "bad general.alignment"; return false; }
    tensors_.resize(n_tensors);
    for (uint64_t i = 0; i < n_tensors; i++) {
        GgufTensor &t = tensors_[i];
        t.name = c.str();
        t.n_dims = (int)c.get<uint32_t>();
        if (!c.ok || t.n_dims < 1 || t.n_dims > 4) { err = "corrupt GGUF tensor info"; return false; }
        for (int d = 0; d < t.n_dims; d++) {
            const uint64_t v = c.get<uint64_t>();
            if (!c.ok || v == 0 || v > (uint64_t)1 << 40) { err = "corrupt GGUF tensor info (dimension of " + t.name + ")"; return false; }
            t.ne[d] = (int64_t)v;
        }
        t.type = (int)c.get<uint32_t>();
        t.offset = c.get<uint64_t>();
        if (!c.ok) { err = "corrupt GGUF tensor info"; return false; }
    }
    size_t hdr = (size_t)(c.p - map_);
    size_t data_off = (hdr + alignment - 1) / alignment * alignment;
    if (data_off > size_) { err = "corrupt GGUF header (no tensor data)"; return false; }
    const uint64_t data_bytes = size_ - data_off;
    for (auto &t : tensors_) {
        int64_t rs = ggml_row_size(t.type, t.ne[0]);
        if (rs < 0) { t.data = nullptr; t.nbytes = -1; }   // unsupported type: only an error if the LM needs it
        else {
            // checked product (dimensions are <= 2^40 each, the file is far smaller than 2^63 bytes)
            unsigned __int128 nb = (unsigned __int128)(uint64_t)rs;
            for (int d = 1; d < 4; d++) { nb *= (uint64_t)t.ne[d]; if (nb > data_bytes) break; }
            if (nb > data_bytes || t.offset > data_bytes - (uint64_t)nb) { err = "GGUF tensor " + t.name + " exceeds file size"; return false; }
            t.nbytes = (int64_t)nb;
            t.data = map_ + data_off + t.offset;
        }
        index_[t.name] = (size_t)(&t - tensors_.data());
    }
    return true;
}

const GgufTensor *GgufFile::find(const std::string &name) const {
    auto it = index_.find(name);
    return it == index_.end() ? nullptr : &tensors_[it->second];
}

}  // namespace msx
