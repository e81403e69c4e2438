// converters.inl — file-to-file converters: GGUF -> quantised GGUF, safetensors -> GGUF (moshi_lm_quantize + moshi_lm_save_gguf,
// src/moshi.cpp:654-695, src/loader.h:149-233).  Included by engine.cu.

// ---- GGUF -> GGUF quantiser (reference: `moshi-sts -q q4_k -g out.gguf`: moshi_lm_quantize + moshi_lm_save_gguf,
// src/moshi.cpp:654-695, WeightLoader::save_gguf src/loader.h:227-233) --------------------------------------------------
namespace {
// the tensors moshi_scaled_embedding_t fetches (lm_utils.h:131-147): lm.text_emb, lm.emb.{c}, lm.depformer_emb.{k},
// lm.depformer_text_emb — a q4_k model stores them as Q4_0
bool is_embedding_table_name(const std::string &name) {
    const std::string tail = ".weight";
    if (name.size() <= tail.size() || name.compare(name.size() - tail.size(), tail.size(), tail)) return false;
    std::string stem = name.substr(0, name.size() - tail.size());
    size_t e = stem.size();
    while (e > 0 && stem[e - 1] >= '0' && stem[e - 1] <= '9') e--;
    if (e < stem.size() && e > 0 && stem[e - 1] == '.') stem.resize(e - 1);
    return stem.size() >= 3 && stem.compare(stem.size() - 3, 3, "emb") == 0;
}
// loader.h:161-172: Q4_K needs K % 256 == 0 else Q4_0, Q4_0 / Q8_0 need K % 32 == 0 else the tensor stays as it is
int on_load_type(int quantize, const std::string &name, int type, int n_dims, int64_t K) {
    if (!quantize || !is_float_type(type) || n_dims != 2 || name.rfind("lm.", 0) != 0) return type;
    if (name.find("condition_provider") != std::string::npos) return type;      // fetched without a destination type (tts.h:16-35)
    int dst = quantize;
    if (dst == T_Q4_K && is_embedding_table_name(name)) dst = T_Q4_0;
    if (dst == T_Q4_K && K % 256) dst = T_Q4_0;
    if ((dst == T_Q4_0 || dst == T_Q8_0) && K % 32) dst = type;
    return dst;
}
struct FileCloser { void operator()(FILE *f) const { if (f) fclose(f); } };
template <typename T> bool put(FILE *f, T v) { return fwrite(&v, sizeof(T), 1, f) == 1; }

// one tensor of the output GGUF: `src` holds rows x K values of src_type; dst_type is a block format (quantised on the
// GPU), F32 (host cast, the norm vectors) or src_type (copied)
struct OutTensor {
    std::string name;
    int n_dims = 0;
    int64_t ne[4] = {1, 1, 1, 1};
    int src_type = 0, dst_type = 0;
    const uint8_t *src = nullptr;
    uint64_t src_bytes = 0;
};
uint64_t out_tensor_bytes(const OutTensor &t) {
    if (t.dst_type == t.src_type) return t.src_bytes;
    int64_t rows = 1;
    for (int d = 1; d < t.n_dims; d++) rows *= t.ne[d];
    return (uint64_t)ggml_row_size(t.dst_type, t.ne[0]) * (uint64_t)rows;
}
// GGUF v3, no key/value pairs (the reference writes none, loader.h:227-233), 32-byte alignment
int write_gguf_impl(msx_model *m, const std::vector<OutTensor> &ts, const char *out_path) {
    std::vector<uint64_t> offset(ts.size()), nbytes(ts.size());
    uint64_t data_bytes = 0;
    for (size_t i = 0; i < ts.size(); i++) {
        nbytes[i] = out_tensor_bytes(ts[i]);
        offset[i] = data_bytes;
        data_bytes = (data_bytes + nbytes[i] + 31) / 32 * 32;
    }
    std::unique_ptr<FILE, FileCloser> out(fopen(out_path, "wb"));
    if (!out) return fail(MSX_ERR_IO, std::string("cannot open ") + out_path + " for writing");
    FILE *o = out.get();
    bool ok = fwrite("GGUF", 1, 4, o) == 4 && put<uint32_t>(o, 3) && put<uint64_t>(o, ts.size()) && put<uint64_t>(o, 0);
    for (size_t i = 0; ok && i < ts.size(); i++) {
        ok = put<uint64_t>(o, ts[i].name.size()) && fwrite(ts[i].name.data(), 1, ts[i].name.size(), o) == ts[i].name.size() &&
             put<uint32_t>(o, (uint32_t)ts[i].n_dims);
        for (int d = 0; ok && d < ts[i].n_dims; d++) ok = put<uint64_t>(o, (uint64_t)ts[i].ne[d]);
        ok = ok && put<uint32_t>(o, (uint32_t)ts[i].dst_type) && put<uint64_t>(o, offset[i]);
    }
    static const uint8_t zeros[32] = {0};
    auto pad32 = [&](uint64_t pos) { const size_t n = (size_t)((32 - pos % 32) % 32); return n == 0 || fwrite(zeros, 1, n, o) == n; };
    ok = ok && pad32((uint64_t)ftell(o));
    std::vector<uint8_t> host;
    for (size_t i = 0; ok && i < ts.size(); i++) {
        const OutTensor &t = ts[i];
        const uint8_t *src = t.src;
        if (t.dst_type == T_F32 && t.src_type != T_F32) {           // bf16 / f16 -> f32 (exact)
            const size_t n = (size_t)(t.src_bytes / 2);
            host.resize(n * 4);
            float *dst = reinterpret_cast<float *>(host.data());
            const uint16_t *h = reinterpret_cast<const uint16_t *>(t.src);
            for (size_t e = 0; e < n; e++) {
                if (t.src_type == T_BF16) { const uint32_t u = (uint32_t)h[e] << 16; memcpy(dst + e, &u, 4); }
                else { __half v; memcpy(&v, h + e, 2); dst[e] = __half2float(v); }
            }
            src = host.data();
        } else if (t.dst_type != t.src_type) {
            if (int e = ensure_staging(m, (size_t)t.src_bytes)) return e;
            CU(cudaMemcpy(m->staging, t.src, (size_t)t.src_bytes, cudaMemcpyHostToDevice));
            const uint8_t *blocks = nullptr;
            if (int e = quantize_staging(m, t.src_type, t.dst_type, t.ne[0], t.ne[1], &blocks)) return e;
            host.resize((size_t)nbytes[i]);
            CU(cudaMemcpy(host.data(), blocks, (size_t)nbytes[i], cudaMemcpyDeviceToHost));
            src = host.data();
        }
        ok = fwrite(src, 1, (size_t)nbytes[i], o) == (size_t)nbytes[i] && pad32(nbytes[i]);
    }
    if (!ok || fflush(o) != 0) return fail(MSX_ERR_IO, std::string("write failed: ") + out_path);
    return 0;
}
// a failed conversion never leaves a truncated output file behind
int write_gguf(msx_model *m, const std::vector<OutTensor> &ts, const char *out_path) {
    const int rc = write_gguf_impl(m, ts, out_path);
    if (rc != 0) unlink(out_path);
    return rc;
}
bool same_file(const char *a, const char *b) {
    struct stat sa, sb;
    return stat(a, &sa) == 0 && stat(b, &sb) == 0 && sa.st_dev == sb.st_dev && sa.st_ino == sb.st_ino;
}
bool ends_with(const std::string &s, const char *tail) {
    const size_t n = strlen(tail);
    return s.size() >= n && s.compare(s.size() - n, n, tail) == 0;
}
}  // namespace

extern "C" int msx_gguf_quantize(const char *in_path, const char *out_path, int quantize, int device) {
    if (!in_path || !out_path) return fail(MSX_ERR_ARG, "null argument");
    if (quantize != 0 && quantize != T_Q8_0 && quantize != T_Q4_K) return fail(MSX_ERR_ARG, "quantize takes 0 (copy), 8 (q8_0) or 12 (q4_k)");
    if (same_file(in_path, out_path)) return fail(MSX_ERR_ARG, "output path is the input file (it is memory-mapped while the output is written)");
    GgufFile f;
    std::string err;
    if (!f.open(in_path, err)) {
        const bool io = err.rfind("cannot open", 0) == 0 || err.rfind("cannot stat", 0) == 0;
        return fail(io ? MSX_ERR_IO : MSX_ERR_FORMAT, err);
    }
    std::unique_ptr<msx_model> m;
    if (int e = device_setup(device, m)) return e;
    std::vector<OutTensor> ts;
    for (const GgufTensor &g : f.tensors()) {
        if (!g.data) return fail(MSX_ERR_FORMAT, "tensor " + g.name + " has unsupported type " + std::to_string(g.type));
        OutTensor t;
        t.name = g.name; t.n_dims = g.n_dims;
        for (int d = 0; d < 4; d++) t.ne[d] = g.ne[d];
        t.src_type = g.type; t.dst_type = on_load_type(quantize, g.name, g.type, g.n_dims, g.ne[0]);
        t.src = g.data; t.src_bytes = (uint64_t)g.nbytes;
        ts.push_back(std::move(t));
    }
    return write_gguf(m.get(), ts, out_path);
}

// The reference's own starting point: model.safetensors (bf16 / f16 / f32, torch names) -> GGUF with the names and
// splits its loader produces (WeightLoader::from_safetensor + save_gguf, src/loader.h:77-83, 149-233):
//   name -> "lm." + name (loader.h:101-105 strips that prefix again when it looks a tensor up);
//   *.in_proj_weight [n*3*d, d] -> n tensors *.in_projs.{i}.weight, *.out_proj.weight [n*d, d] -> *.out_projs.{i}.weight
//   (per-step depformer weights, src/moshi/modules/transformer.h:764-849);
//   vectors (norm alpha [1,1,d], biases) -> F32 (loader.h:204-209); 2-D float tensors -> the quantisation rules above.
extern "C" int msx_safetensors_to_gguf(const char *in_path, const char *out_path, int quantize, int device) {
    if (!in_path || !out_path) return fail(MSX_ERR_ARG, "null argument");
    if (quantize != 0 && quantize != T_Q8_0 && quantize != T_Q4_K) return fail(MSX_ERR_ARG, "quantize takes 0 (copy), 8 (q8_0) or 12 (q4_k)");
    if (same_file(in_path, out_path)) return fail(MSX_ERR_ARG, "output path is the input file (it is memory-mapped while the output is written)");
    SafeTensorsFile f;
    std::string err;
    if (!f.open(in_path, err)) {
        const bool io = err.rfind("cannot open", 0) == 0 || err.rfind("cannot stat", 0) == 0 || err.rfind("cannot mmap", 0) == 0;
        return fail(io ? MSX_ERR_IO : MSX_ERR_FORMAT, err);
    }
    std::vector<OutTensor> ts;                  // the tensor list is validated on the host before the device is touched
    for (const SafeTensor &st : f.tensors()) {
        const int type = st.dtype == "F32" ? T_F32 : st.dtype == "F16" ? T_F16 : st.dtype == "BF16" ? T_BF16 : -1;
        if (type < 0) continue;      // integer / bool bookkeeping tensors: the reference's loader never fetches them, save_gguf never writes them
        if (st.shape.empty() || st.shape.size() > 4) return fail(MSX_ERR_FORMAT, "tensor " + st.name + " has unsupported rank");
        int64_t count = 1;
        for (int64_t v : st.shape) {
            if (v <= 0 || v > ((int64_t)1 << 40) || count > ((int64_t)1 << 46) / v) return fail(MSX_ERR_FORMAT, "tensor " + st.name + ": bad shape");
            count *= v;
        }
        if ((uint64_t)count * (type == T_F32 ? 4 : 2) != st.nbytes) return fail(MSX_ERR_FORMAT, "tensor " + st.name + ": shape does not match its bytes");
        const std::string name = "lm." + st.name;
        const int64_t K = st.shape.back();
        int64_t parts = 1;
        std::string stem;
        if (st.shape.size() == 2 && ends_with(name, "in_proj_weight")) { parts = st.shape[0] / (3 * K); stem = name.substr(0, name.size() - strlen("in_proj_weight")) + "in_projs."; }
        else if (st.shape.size() == 2 && ends_with(name, ".out_proj.weight")) { parts = st.shape[0] / K; stem = name.substr(0, name.size() - strlen(".out_proj.weight")) + ".out_projs."; }
        if (parts < 1 || (parts > 1 && st.shape[0] % parts)) return fail(MSX_ERR_FORMAT, "tensor " + st.name + ": cannot split into per-step weights");
        for (int64_t p = 0; p < parts; p++) {
            OutTensor t;
            t.name = stem.empty() ? name : stem + std::to_string(p) + ".weight";
            t.n_dims = (int)st.shape.size();
            for (int d = 0; d < t.n_dims; d++) t.ne[d] = st.shape[st.shape.size() - 1 - d];      // dimensions are inverted
            if (!stem.empty()) t.ne[1] = st.shape[0] / parts;
            t.src_type = type;
            t.src_bytes = st.nbytes / (uint64_t)parts;
            t.src = st.data + (uint64_t)p * t.src_bytes;
            t.dst_type = count == K ? T_F32 : on_load_type(quantize, t.name, type, t.n_dims, K);
            ts.push_back(std::move(t));
        }
    }
    std::unique_ptr<msx_model> m;
    if (int e = device_setup(device, m)) return e;
    return write_gguf(m.get(), ts, out_path);
}
