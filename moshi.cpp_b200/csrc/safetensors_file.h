// safetensors_file.h — minimal read-only safetensors parser (mmap): 8-byte little-endian header length, a JSON object
// {name: {"dtype": "BF16", "shape": [..], "data_offsets": [begin, end]}, "__metadata__": {...}}, then the data.
// Replaces the reference's SafeTensorFile (src/context.h:78-226) for the file-to-GGUF quantiser.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace msx {

struct SafeTensor {
    std::string name, dtype;
    std::vector<int64_t> shape;     // torch order: last dimension fastest
    const uint8_t *data = nullptr;
    uint64_t nbytes = 0;
};

class SafeTensorsFile {
public:
    SafeTensorsFile() = default;
    ~SafeTensorsFile() {
        if (map_) munmap(const_cast<uint8_t *>(map_), size_);
        if (fd_ >= 0) close(fd_);
    }
    SafeTensorsFile(const SafeTensorsFile &) = delete;
    SafeTensorsFile &operator=(const SafeTensorsFile &) = delete;

    bool open(const std::string &path, std::string &err) {
        fd_ = ::open(path.c_str(), O_RDONLY);
        if (fd_ < 0) { err = "cannot open " + path; return false; }
        struct stat st;
        if (fstat(fd_, &st) != 0) { err = "cannot stat " + path; return false; }
        size_ = (size_t)st.st_size;
        if (size_ < 10) { err = "not a safetensors file (too short)"; return false; }
        void *p = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (p == MAP_FAILED) { err = "cannot mmap " + path; return false; }
        map_ = (const uint8_t *)p;
        uint64_t hlen;
        memcpy(&hlen, map_, 8);
        if (hlen < 2 || hlen > size_ - 8) { err = "corrupt safetensors header length"; return false; }
        s_ = (const char *)map_ + 8; n_ = (size_t)hlen; i_ = 0;
        const uint8_t *data = map_ + 8 + hlen;
        const uint64_t data_size = size_ - 8 - hlen;
        if (!parse(data, data_size)) { err = "corrupt safetensors header near byte " + std::to_string(i_); return false; }
        return true;
    }
    const std::vector<SafeTensor> &tensors() const { return tensors_; }

private:
    void ws() { while (i_ < n_ && (s_[i_] == ' ' || s_[i_] == '\n' || s_[i_] == '\t' || s_[i_] == '\r')) i_++; }
    bool eat(char c) { ws(); if (i_ < n_ && s_[i_] == c) { i_++; return true; } return false; }
    bool str(std::string &out) {
        ws();
        if (i_ >= n_ || s_[i_] != '"') return false;
        out.clear();
        for (i_++; i_ < n_ && s_[i_] != '"'; i_++) {
            if (s_[i_] == '\\' && i_ + 1 < n_) { i_++; out.push_back(s_[i_] == 'n' ? '\n' : s_[i_] == 't' ? '\t' : s_[i_]); }
            else out.push_back(s_[i_]);
        }
        return i_ < n_ && s_[i_++] == '"';
    }
    bool integer(int64_t &v) {
        ws();
        size_t b = i_;
        v = 0;
        while (i_ < n_ && s_[i_] >= '0' && s_[i_] <= '9') v = v * 10 + (s_[i_++] - '0');
        return i_ > b;
    }
    bool int_array(std::vector<int64_t> &out) {
        out.clear();
        if (!eat('[')) return false;
        if (eat(']')) return true;
        do { int64_t v; if (!integer(v)) return false; out.push_back(v); } while (eat(','));
        return eat(']');
    }
    bool skip_value() {            // strings, numbers, literals, nested objects / arrays (only __metadata__ gets here)
        ws();
        if (i_ >= n_) return false;
        if (s_[i_] == '"') { std::string t; return str(t); }
        if (s_[i_] == '{' || s_[i_] == '[') {
            const char close = s_[i_] == '{' ? '}' : ']';
            i_++;
            if (eat(close)) return true;
            do {
                if (close == '}') { std::string k; if (!str(k) || !eat(':')) return false; }
                if (!skip_value()) return false;
            } while (eat(','));
            return eat(close);
        }
        size_t b = i_;
        while (i_ < n_ && s_[i_] != ',' && s_[i_] != '}' && s_[i_] != ']') i_++;
        return i_ > b;
    }
    bool parse(const uint8_t *data, uint64_t data_size) {
        if (!eat('{')) return false;
        if (eat('}')) return true;
        do {
            std::string name;
            if (!str(name) || !eat(':')) return false;
            if (name == "__metadata__") { if (!skip_value()) return false; continue; }
            SafeTensor t;
            t.name = name;
            std::vector<int64_t> off;
            if (!eat('{')) return false;
            do {
                std::string key;
                if (!str(key) || !eat(':')) return false;
                if (key == "dtype") { if (!str(t.dtype)) return false; }
                else if (key == "shape") { if (!int_array(t.shape)) return false; }
                else if (key == "data_offsets") { if (!int_array(off)) return false; }
                else if (!skip_value()) return false;
            } while (eat(','));
            if (!eat('}')) return false;
            if (off.size() != 2 || off[0] > off[1] || (uint64_t)off[1] > data_size) return false;
            t.data = data + off[0];
            t.nbytes = (uint64_t)(off[1] - off[0]);
            tensors_.push_back(std::move(t));
        } while (eat(','));
        return eat('}');
    }

    int fd_ = -1;
    const uint8_t *map_ = nullptr;
    size_t size_ = 0;
    const char *s_ = nullptr;
    size_t n_ = 0, i_ = 0;
    std::vector<SafeTensor> tensors_;
};

}  // namespace msx
