// engine.cu — model loading/repack, per-stream state, CUDA-graph step orchestration and the C ABI
// declared in include/moshi_b200.h.  See DESIGN.md for the mapping to the reference.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/moshi_b200.h"
#include "attention.cuh"
#include "common.cuh"
#include "gemv.cuh"
#include "gguf_file.h"
#include "local_attn.cuh"
#include "misc_kernels.cuh"
#include "mma_gemm.cuh"
#include "safetensors_file.h"
#include "sample.cuh"
#include "step_kernel.cuh"
#include "tts_kernels.cuh"

using namespace msx;

// -------------------------------------------------------------------------------------------------
// errors
// -------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CU(expr)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(MSX_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));               \
    } while (0)

extern "C" const char *msx_last_error(void) { return g_err.c_str(); }

// ---- NCCL, bound at run time (tensor-parallel streams only) -------------------------------------------
// dlopen by soname: inside a torch process this resolves to the NCCL torch has already loaded, so one
// process never holds two NCCL copies; a process that never asks for tensor parallelism never loads it.
namespace {
struct Nccl {
    typedef struct { char internal[128]; } UniqueId;
    typedef void *Comm;
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(Comm *, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;
};
Nccl &nccl() {
    static Nccl n;
    static bool tried = false;
    if (tried) return n;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { n.why = std::string("cannot load libnccl: ") + dlerror(); return n; }
    n.GetUniqueId = (int (*)(Nccl::UniqueId *))dlsym(h, "ncclGetUniqueId");
    n.CommInitRank = (int (*)(Nccl::Comm *, int, Nccl::UniqueId, int))dlsym(h, "ncclCommInitRank");
    n.CommDestroy = (int (*)(Nccl::Comm))dlsym(h, "ncclCommDestroy");
    n.AllReduce = (int (*)(const void *, void *, size_t, int, int, Nccl::Comm, cudaStream_t))dlsym(h, "ncclAllReduce");
    n.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllReduce && n.GetErrorString;
    if (!n.ok) n.why = "libnccl lacks an expected symbol";
    return n;
}
constexpr int kNcclFloat64 = 8, kNcclSum = 0;       // ncclDataType_t / ncclRedOp_t values (nccl.h)
}  // namespace
extern "C" const char *msx_version(void) { return "moshi_b200 0.1 (sm_100a)"; }
extern "C" int msx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// -------------------------------------------------------------------------------------------------
// model
// -------------------------------------------------------------------------------------------------
struct LayerW {
    const float *norm1 = nullptr, *norm2 = nullptr;
    std::vector<QLinear> in_proj, out_proj, lin_in, lin_out;
    // cross-attention layers (TTS): LayerNorm weight / bias, in_proj [dim -> 3 dim] (q | k | v rows), out_proj
    const float *norm_cross_w = nullptr, *norm_cross_b = nullptr;
    QLinear cross_in, cross_out;
};

struct msx_model {
    msx_config cfg{};
    int device = 0;
    int num_sms = 148;
    int hidden = 0, dep_hidden = 0, dep_cap = 0, dep_nw = 0;
    // tensor parallelism over the temporal transformer: this rank owns heads [h0, h1) and the
    // hidden slice [f0, f1); everything else (embeddings, text head, depformer) is replicated
    int tp_rank = 0, tp_world = 1;
    int heads_local = 0, adim = 0, h0 = 0, f0 = 0, hidden_local = 0;
    int quantize = 0;                 // T_Q8_0 / T_Q4_K: f32 / f16 / bf16 tensors of the file are quantised while loading
    uint8_t *qstaging = nullptr; size_t qstaging_bytes = 0;
    std::vector<EmbTable> emb;        // [n_q+1]: text, audio 0..n_q-1
    EmbTable *d_emb = nullptr;        // device copy of `emb`
    EmbTable dep_text_emb;
    std::vector<EmbTable> dep_emb;    // [dep_q-1]
    std::vector<LayerW> layers, dep_layers;
    const float *out_norm = nullptr;
    const float *rope_freq = nullptr, *dep_rope_freq = nullptr;   // [Dh/2] RoPE frequencies (host-computed)
    QLinear text_linear;
    std::vector<QLinear> dep_in, linears, extra_heads;
    // TTS family: demuxed text embedding projections (temporal: repacked linears; depformer: GGUF-format rows),
    // low-rank projections of the depformer embeddings
    QLinear text_out1, text_out2;
    EmbTable dep_text_out1, dep_text_out2, dep_text_lr;
    std::vector<EmbTable> dep_emb_lr;
    bool dep_small = false;           // depformer embeddings go through small_linear_kernel
    // TTS voice conditioners (tts.h:5-35), present when the GGUF carries lm.condition_provider.conditioners.*
    FloatTensor cfg_embed, cfg_proj, control_embed, control_proj, spk_pad, spk_proj;
    const float *cond_freq = nullptr; // [dim/2] timestep-embedding frequencies
    bool has_conditioners = false;
    std::vector<void *> allocs;
    std::unordered_map<const void *, QTiles> tiles;   // MMA unit layout of a linear, keyed by its qs plane (batch.inl)
    // stream layout of a linear for the persistent step kernel (step_kernel.cuh), keyed by its qs plane
    struct StreamW { const uint8_t *p = nullptr; int gran = 1; };
    std::unordered_map<const void *, StreamW> wstream;
    bool stream_ok = true;            // every linear of the decode step has a stream-layout copy of one weight type
    int stream_type = 0;
    QLinear dep_in_all;               // depformer_in[w_k] of all dep_q steps as ONE matrix [dep_q * dep_dim][dim] (stream layout only)
    int64_t weight_bytes_per_frame = 0;
    int64_t device_bytes = 0;
    uint8_t *staging = nullptr;
    size_t staging_bytes = 0;

    ~msx_model() {
        cudaSetDevice(device);
        for (void *p : allocs) cudaFree(p);
        if (staging) cudaFree(staging);
        if (qstaging) cudaFree(qstaging);
    }
};

namespace {

int dev_alloc(msx_model *m, void **p, size_t bytes) {
    CU(cudaMalloc(p, std::max<size_t>(bytes, 16)));
    m->allocs.push_back(*p);
    m->device_bytes += (int64_t)bytes;
    return 0;
}

int ensure_staging(msx_model *m, size_t bytes) {
    if (bytes <= m->staging_bytes) return 0;
    if (m->staging) cudaFree(m->staging);
    m->staging = nullptr; m->staging_bytes = 0;
    CU(cudaMalloc((void **)&m->staging, bytes));
    m->staging_bytes = bytes;
    return 0;
}

// quantise-on-load (moshi_lm_quantize on an unquantised file): the float rows already sit in m->staging; on return
// *blocks points at GGUF-format rows of dst_type (Q8_0, Q4_0 or Q4_K) on the device
int quantize_staging(msx_model *m, int src_type, int dst_type, int64_t K, int64_t rows, const uint8_t **blocks) {
    const int64_t rs = ggml_row_size(dst_type, K);
    if (rs < 0) return fail(MSX_ERR_FORMAT, std::string("K is not a multiple of the block size of ") + ggml_type_name(dst_type));
    const size_t need = (size_t)rs * rows;
    if (need > m->qstaging_bytes) {
        if (m->qstaging) cudaFree(m->qstaging);
        m->qstaging = nullptr; m->qstaging_bytes = 0;
        CU(cudaMalloc((void **)&m->qstaging, need));
        m->qstaging_bytes = need;
    }
    if (dst_type == T_Q4_K) {
        const long long nblk = (long long)(K / 256) * rows;
        quantize_rows_q4_K_kernel<<<(unsigned)((nblk * 8 + kQ4kQuantThreads - 1) / kQ4kQuantThreads), kQ4kQuantThreads>>>(
            m->staging, src_type, nblk, m->qstaging);
    } else {
        const long long nblk = (long long)(K / 32) * rows;
        const unsigned grid = (unsigned)((nblk * 32 + 255) / 256);
        if (dst_type == T_Q8_0) quantize_rows_q8_0_kernel<<<grid, 256>>>(m->staging, src_type, nblk, m->qstaging);
        else if (dst_type == T_Q4_0) quantize_rows_q4_0_kernel<<<grid, 256>>>(m->staging, src_type, nblk, m->qstaging);
        else return fail(MSX_ERR_ARG, "quantise-on-load: unsupported target type");
    }
    CU(cudaGetLastError());
    *blocks = m->qstaging;
    return 0;
}
bool is_float_type(int t) { return t == T_F32 || t == T_F16 || t == T_BF16; }

// Upload a GGUF tensor [rows][K] and repack it into device tiles. perm_half: see repack kernels.
// Raw GGUF blocks (device) -> stream layout of the persistent step kernel
int upload_stream(msx_model *m, const uint8_t *d_blocks, int type, int64_t K, int64_t rows, int perm_half, const void *key) {
    const size_t qbytes = (size_t)ggml_row_size(type, K) * rows;
    void *ws = nullptr;
    if (int e = dev_alloc(m, &ws, qbytes)) return e;
    const long long n = (long long)rows * (K / 256);
    const int gran = perm_half > 0 ? 2 : 1;
    sk::repack_stream_kernel<<<(unsigned)((n + 255) / 256), 256>>>(d_blocks, (uint8_t *)ws, type, (int)rows, (int)K, gran, m->num_sms, perm_half);
    CU(cudaGetLastError());
    m->wstream[key] = msx_model::StreamW{(const uint8_t *)ws, gran};
    m->stream_type = type;
    return 0;
}

// blocks_copy: optional device buffer that receives the (quantised) GGUF blocks of this matrix as they are
int upload_linear(msx_model *m, const void *host, int type, int64_t K, int64_t rows, int perm_half, QLinear *out, uint8_t *blocks_copy = nullptr) {
    const bool on_load = m->quantize && is_float_type(type);
    if (type != T_Q4_K && type != T_Q8_0 && !on_load)
        return fail(MSX_ERR_FORMAT, std::string("linear weights must be q4_k or q8_0, got ") + ggml_type_name(type));
    // loader.h:161-172 would fall back to Q4_0 rows for K % 256 != 0; the GEMV paths take Q4_K / Q8_0 only
    if (on_load && m->quantize == T_Q4_K && K % 256)
        return fail(MSX_ERR_FORMAT, "quantise-on-load q4_k: a linear with K % 256 != 0 would become q4_0, which the linears do not take");
    const int64_t rs = ggml_row_size(type, K);
    if (rs < 0) return fail(MSX_ERR_FORMAT, "K is not a multiple of the block size");
    const size_t raw = (size_t)rs * rows;
    if (int e = ensure_staging(m, raw)) return e;
    CU(cudaMemcpy(m->staging, host, raw, cudaMemcpyHostToDevice));
    const uint8_t *src_blocks = m->staging;
    if (on_load) { if (int e = quantize_staging(m, type, m->quantize, K, rows, &src_blocks)) return e; type = m->quantize; }
    QLinear w;
    w.type = type; w.K = (int)K; w.rows = (int)rows; w.gs = K >= 4096 ? 32 : 16; w.gate = perm_half > 0;
    void *qs = nullptr, *sc = nullptr, *dd = nullptr;
    if (type == T_Q4_K) {
        if (int e = dev_alloc(m, &qs, (size_t)rows * K / 2)) return e;
        if (int e = dev_alloc(m, &sc, (size_t)rows * (K / 64) * 4)) return e;
        if (int e = dev_alloc(m, &dd, (size_t)rows * (K / 256) * 4)) return e;
        const long long n = (long long)rows * (K / 64);
        repack_q4k_kernel<<<(unsigned)((n + 255) / 256), 256>>>(src_blocks, (uint8_t *)qs, (uint32_t *)sc, (uint32_t *)dd,
                                                                (int)rows, (int)K, w.gs, perm_half);
    } else {
        if (int e = dev_alloc(m, &qs, (size_t)rows * K)) return e;
        if (int e = dev_alloc(m, &dd, (size_t)rows * (K / 32) * 2)) return e;
        const long long n = (long long)rows * (K / 32);
        repack_q8_0_kernel<<<(unsigned)((n + 255) / 256), 256>>>(src_blocks, (uint8_t *)qs, (uint16_t *)dd, (int)rows, (int)K, w.gs, perm_half);
    }
    CU(cudaGetLastError());
    w.qs = (const uint8_t *)qs; w.sc = (const uint32_t *)sc; w.dd = dd;
    // second copy in the stream layout of the persistent step kernel: CTA spans of 32-row x super-block units, GGUF bytes exactly
    if (m->tp_world == 1 && K % 256 == 0 && rows % (perm_half > 0 ? 2 : 1) == 0 && (m->stream_type == 0 || m->stream_type == type)) {
        if (int e = upload_stream(m, src_blocks, type, K, rows, perm_half, w.qs)) return e;
    } else m->stream_ok = false;
    if (blocks_copy) CU(cudaMemcpyAsync(blocks_copy, src_blocks, (size_t)ggml_row_size(type, K) * rows, cudaMemcpyDeviceToDevice, 0));
    CU(cudaDeviceSynchronize());
    *out = w;
    return 0;
}

int upload_table(msx_model *m, const void *host, int type, int64_t K, int64_t rows, EmbTable *out) {
    int64_t rs = ggml_row_size(type, K);
    if (rs < 0 || type == T_Q4_K)
        return fail(MSX_ERR_FORMAT, std::string("embedding table type not supported: ") + ggml_type_name(type));
    void *d = nullptr;
    if (m->quantize && is_float_type(type) && K % 32 == 0) {
        // the reference quantises embedding tables with the model (lm_utils.h:131-147): float rows -> Q8_0 rows for a
        // q8_0 model, Q4_0 rows for a q4_k model
        const int dst_type = m->quantize == T_Q4_K ? T_Q4_0 : T_Q8_0;
        if (int e = ensure_staging(m, (size_t)rs * rows)) return e;
        CU(cudaMemcpy(m->staging, host, (size_t)rs * rows, cudaMemcpyHostToDevice));
        const uint8_t *blocks = nullptr;
        if (int e = quantize_staging(m, type, dst_type, K, rows, &blocks)) return e;
        type = dst_type; rs = ggml_row_size(type, K);
        if (int e = dev_alloc(m, &d, (size_t)rs * rows)) return e;
        CU(cudaMemcpy(d, blocks, (size_t)rs * rows, cudaMemcpyDeviceToDevice));
        out->data = (const uint8_t *)d; out->type = type; out->K = (int)K; out->rows = (int)rows; out->row_bytes = (int)rs;
        return 0;
    }
    if (int e = dev_alloc(m, &d, (size_t)rs * rows)) return e;
    CU(cudaMemcpy(d, host, (size_t)rs * rows, cudaMemcpyHostToDevice));
    out->data = (const uint8_t *)d; out->type = type; out->K = (int)K; out->rows = (int)rows; out->row_bytes = (int)rs;
    return 0;
}

struct Loader {
    msx_model *m;
    GgufFile &f;
    int64_t linear_bytes = 0;   // GGUF bytes of the last linear loaded

    const GgufTensor *need(const std::string &name) {
        const GgufTensor *t = f.find(name);
        if (!t) { fail(MSX_ERR_FORMAT, "tensor missing in GGUF: " + name); return nullptr; }
        if (!t->data) { fail(MSX_ERR_FORMAT, "tensor " + name + " has unsupported type " + std::to_string(t->type)); return nullptr; }
        return t;
    }
    int linear(const std::string &name, int64_t K, int64_t rows, QLinear *out, int perm_half = 0, uint8_t *blocks_copy = nullptr) {
        const GgufTensor *t = need(name);
        if (!t) return MSX_ERR_FORMAT;
        if ((K > 0 && t->ne[0] != K) || (rows > 0 && t->ne[1] != rows))
            return fail(MSX_ERR_FORMAT, "shape mismatch for " + name + ": got [" + std::to_string(t->ne[0]) + "," +
                                            std::to_string(t->ne[1]) + "], want [" + std::to_string(K) + "," + std::to_string(rows) + "]");
        linear_bytes = (m->quantize && is_float_type(t->type)) ? t->ne[1] * ggml_row_size(m->quantize, t->ne[0]) : t->nbytes;
        return upload_linear(m, t->data, t->type, t->ne[0], t->ne[1], perm_half, out, blocks_copy);
    }
    // tensor-parallel shard of a linear: the listed row ranges (concatenated) x the K-slice [k0, k1) of every row
    int linear_slice(const std::string &name, int64_t K, int64_t rows, const std::vector<std::pair<int64_t, int64_t>> &ranges,
                     int64_t k0, int64_t k1, int perm_half, QLinear *out) {
        const GgufTensor *t = need(name);
        if (!t) return MSX_ERR_FORMAT;
        if (t->ne[0] != K || t->ne[1] != rows) return fail(MSX_ERR_FORMAT, "shape mismatch for " + name);
        if (t->type != T_Q4_K && t->type != T_Q8_0) return fail(MSX_ERR_FORMAT, name + ": tensor-parallel shards need q4_k or q8_0 weights");
        const int64_t bw = t->type == T_Q4_K ? 256 : 32, bb = t->type == T_Q4_K ? 144 : 34;
        if (k0 % bw || k1 % bw || k1 <= k0 || k1 > K) return fail(MSX_ERR_ARG, name + ": K-slice is not block aligned");
        const int64_t rs = ggml_row_size(t->type, K), srs = (k1 - k0) / bw * bb;
        int64_t n = 0;
        for (auto &r : ranges) n += r.second - r.first;
        std::vector<uint8_t> buf((size_t)n * srs);
        int64_t i = 0;
        for (auto &rg : ranges)
            for (int64_t r = rg.first; r < rg.second; r++, i++)
                memcpy(buf.data() + (size_t)i * srs, (const uint8_t *)t->data + (size_t)r * rs + (size_t)(k0 / bw) * bb, (size_t)srs);
        linear_bytes = n * srs;
        return upload_linear(m, buf.data(), t->type, k1 - k0, n, perm_half, out);
    }
    int table(const std::string &name, int64_t K, int64_t rows, EmbTable *out) {
        const GgufTensor *t = need(name);
        if (!t) return MSX_ERR_FORMAT;
        if (t->ne[0] != K || t->ne[1] != rows)
            return fail(MSX_ERR_FORMAT, "shape mismatch for " + name);
        return upload_table(m, t->data, t->type, K, rows, out);
    }
    // an unquantised tensor kept in its file type (the conditioners: loader.h fetch() without a destination type)
    int float_tensor(const std::string &name, int64_t ne0, FloatTensor *out) {
        const GgufTensor *t = need(name);
        if (!t) return MSX_ERR_FORMAT;
        if (!is_float_type(t->type)) return fail(MSX_ERR_FORMAT, name + " must be f32 / f16 / bf16");
        if (ne0 > 0 && t->ne[0] != ne0) return fail(MSX_ERR_FORMAT, "shape mismatch for " + name);
        void *d = nullptr;
        if (int e = dev_alloc(m, &d, (size_t)t->nbytes)) return e;
        CU(cudaMemcpy(d, t->data, (size_t)t->nbytes, cudaMemcpyHostToDevice));
        out->data = (const uint8_t *)d; out->type = t->type; out->ne0 = (int32_t)t->ne[0];
        out->ne1 = (int32_t)(t->ne[1] * t->ne[2] * t->ne[3]);
        return 0;
    }
    int vec_f32(const std::string &name, int64_t n, const float **out) {
        const GgufTensor *t = need(name);
        if (!t) return MSX_ERR_FORMAT;
        if (t->type != T_F32 || t->ne[0] != n) return fail(MSX_ERR_FORMAT, "norm tensor " + name + " must be f32[" + std::to_string(n) + "]");
        void *d = nullptr;
        if (int e = dev_alloc(m, &d, (size_t)n * 4)) return e;
        CU(cudaMemcpy(d, t->data, (size_t)n * 4, cudaMemcpyHostToDevice));
        *out = (const float *)d;
        return 0;
    }
};

int check_config(const msx_config *c) {
    if (!c) return fail(MSX_ERR_ARG, "config is null");
    if (c->dim <= 0 || c->num_heads <= 0 || c->num_layers <= 0 || c->context <= 0) return fail(MSX_ERR_ARG, "bad temporal dims");
    if (c->dim % c->num_heads) return fail(MSX_ERR_ARG, "dim % num_heads != 0");
    const int dh = c->dim / c->num_heads;
    if (dh != 64 && dh != 128) return fail(MSX_ERR_ARG, "head dim must be 64 or 128");
    if (c->n_q < 0 || c->n_q + 1 > MSX_MAX_CODEBOOKS || c->dep_q < 0 || c->dep_q > MSX_MAX_STEPS) return fail(MSX_ERR_ARG, "bad codebook counts");
    if (c->n_delays < c->n_q + 1) return fail(MSX_ERR_ARG, "delays shorter than n_q + 1");
    if (c->dep_q > 0) {
        if (c->dep_dim <= 0 || c->dep_heads <= 0 || c->dep_layers <= 0) return fail(MSX_ERR_ARG, "bad depformer dims");
        const int ddh = c->dep_dim / c->dep_heads;
        if (c->dep_dim % c->dep_heads || (ddh != 64 && ddh != 128)) return fail(MSX_ERR_ARG, "depformer head dim must be 64 or 128");
        if (c->dep_context <= 0 && c->schedule_len <= 0) return fail(MSX_ERR_ARG, "depformer needs a context or a schedule");
        if (c->schedule_len && c->schedule_len < c->dep_q) return fail(MSX_ERR_ARG, "schedule shorter than dep_q");
    }
    return 0;
}

}  // namespace

extern "C" int msx_model_load_gguf(const char *path, const msx_config *cfg, int device, msx_model **out) {
    return msx_model_load_gguf_tp(path, cfg, device, 0, 1, out);
}

extern "C" int msx_model_load_gguf_tp(const char *path, const msx_config *cfg, int device, int tp_rank, int tp_world, msx_model **out) {
    return msx_model_load_gguf_ex(path, cfg, device, tp_rank, tp_world, 0, out);
}

extern "C" int msx_model_load_gguf_ex(const char *path, const msx_config *cfg, int device, int tp_rank, int tp_world, int quantize,
                                      msx_model **out) {
    if (!path || !out) return fail(MSX_ERR_ARG, "null argument");
    if (quantize != 0 && quantize != T_Q8_0 && quantize != T_Q4_K) return fail(MSX_ERR_ARG, "quantise-on-load takes 0 (as is), 8 (q8_0) or 12 (q4_k)");
    if (quantize && tp_world > 1) return fail(MSX_ERR_ARG, "quantise-on-load is not combined with tensor-parallel shards");
    *out = nullptr;
    if (int e = check_config(cfg)) return e;
    if (tp_world < 1 || tp_rank < 0 || tp_rank >= tp_world) return fail(MSX_ERR_ARG, "bad tensor-parallel rank / world");
    if (tp_world > 1) {
        if (cfg->num_heads % tp_world) return fail(MSX_ERR_ARG, "num_heads must be divisible by the tensor-parallel world size");
        if (cfg->cross_attention) return fail(MSX_ERR_ARG, "tensor parallelism does not cover cross-attention layers");
    }
    GgufFile f;
    std::string err;
    if (!f.open(path, err)) {
        const bool io = err.rfind("cannot open", 0) == 0 || err.rfind("cannot stat", 0) == 0;
        return fail(io ? MSX_ERR_IO : MSX_ERR_FORMAT, err);
    }
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(MSX_ERR_CUDA, "no such CUDA device " + std::to_string(device));
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(MSX_ERR_CUDA, std::string("moshi_b200 is built for sm_100a only; device is ") + prop.name);

    std::unique_ptr<msx_model> m(new msx_model);
    m->cfg = *cfg; m->device = device; m->num_sms = prop.multiProcessorCount;
    const msx_config &c = m->cfg;
    Loader L{m.get(), f};
    const int d = c.dim;
    const int Dh = d / c.num_heads;
    m->tp_rank = tp_rank; m->tp_world = tp_world; m->quantize = quantize;
    m->heads_local = c.num_heads / tp_world; m->h0 = tp_rank * m->heads_local; m->adim = m->heads_local * Dh;

    // embeddings (lm.h:386-391)
    m->emb.resize(c.n_q + 1);
    if (int e = L.table("lm.text_emb.weight", d, c.text_card + 1, &m->emb[0])) return e;
    if (c.demux_second_stream) {      // lm_utils.h:14-40
        if (int e = L.linear("lm.text_emb.out1.weight", d, d, &m->text_out1)) return e;
        if (int e = L.linear("lm.text_emb.out2.weight", d, d, &m->text_out2)) return e;
    }
    for (int q = 0; q < c.n_q; q++)
        if (int e = L.table("lm.emb." + std::to_string(q) + ".weight", d, c.card + 1, &m->emb[q + 1])) return e;
    {
        void *p = nullptr;
        if (int e = dev_alloc(m.get(), &p, sizeof(EmbTable) * m->emb.size())) return e;
        CU(cudaMemcpy(p, m->emb.data(), sizeof(EmbTable) * m->emb.size(), cudaMemcpyHostToDevice));
        m->d_emb = (EmbTable *)p;
    }
    // temporal transformer (transformer.h:1042-1080)
    m->layers.resize(c.num_layers);
    int64_t wb = 0;
    for (int i = 0; i < c.num_layers; i++) {
        LayerW &l = m->layers[i];
        const std::string p = "lm.transformer.layers." + std::to_string(i) + ".";
        l.in_proj.resize(1); l.out_proj.resize(1); l.lin_in.resize(1); l.lin_out.resize(1);
        if (int e = L.vec_f32(p + "norm1.alpha", d, &l.norm1)) return e;
        if (int e = L.vec_f32(p + "norm2.alpha", d, &l.norm2)) return e;
        if (tp_world == 1) {
            if (int e = L.linear(p + "self_attn.in_projs.0.weight", d, 3 * d, &l.in_proj[0])) return e;
            wb += L.linear_bytes;
            if (int e = L.linear(p + "self_attn.out_projs.0.weight", d, d, &l.out_proj[0])) return e;
            wb += L.linear_bytes;
        } else {
            // q | k | v rows of this rank's heads; out_proj columns of the same heads (partial sums, all-reduced)
            const int64_t r0 = (int64_t)m->h0 * Dh, r1 = r0 + m->adim;
            if (int e = L.linear_slice(p + "self_attn.in_projs.0.weight", d, 3 * d, {{r0, r1}, {d + r0, d + r1}, {2 * d + r0, 2 * d + r1}}, 0, d, 0, &l.in_proj[0])) return e;
            wb += L.linear_bytes;
            if (int e = L.linear_slice(p + "self_attn.out_projs.0.weight", d, d, {{0, d}}, r0, r1, 0, &l.out_proj[0])) return e;
            wb += L.linear_bytes;
        }
        if (c.cross_attention) {      // transformer.h:1053-1056; bias is optional (torch.h:62-68)
            if (int e = L.vec_f32(p + "norm_cross.weight", d, &l.norm_cross_w)) return e;
            if (f.find(p + "norm_cross.bias")) if (int e = L.vec_f32(p + "norm_cross.bias", d, &l.norm_cross_b)) return e;
            if (int e = L.linear(p + "cross_attention.in_projs.0.weight", d, 3 * d, &l.cross_in)) return e;
            wb += L.linear_bytes / 3;     // per frame only the q rows are read; k / v rows once per conditioning
            if (int e = L.linear(p + "cross_attention.out_projs.0.weight", d, d, &l.cross_out)) return e;
            wb += L.linear_bytes;
        }
        const GgufTensor *t = L.need(p + "gating.linear_in.weight");
        if (!t) return MSX_ERR_FORMAT;
        const int F = (int)(t->ne[1] / 2);
        if (i == 0) {
            m->hidden = F;
            // hidden slice of this rank, on super-block (256) boundaries when the width allows it, else on 32
            const int unit = F % 256 == 0 ? 256 : 32, nu = F / unit;
            m->f0 = (int)((long long)tp_rank * nu / tp_world) * unit;
            m->hidden_local = (int)((long long)(tp_rank + 1) * nu / tp_world) * unit - m->f0;
            if (m->hidden_local <= 0) return fail(MSX_ERR_ARG, "hidden size too small for this tensor-parallel world size");
        }
        if (F != m->hidden || t->ne[1] != 2 * F) return fail(MSX_ERR_FORMAT, "inconsistent gating hidden size");
        if (tp_world == 1) {
            if (int e = L.linear(p + "gating.linear_in.weight", d, 2 * F, &l.lin_in[0], /*perm_half=*/F)) return e;
            wb += L.linear_bytes;
            if (int e = L.linear(p + "gating.linear_out.weight", F, d, &l.lin_out[0])) return e;
            wb += L.linear_bytes;
        } else {
            const int64_t a0 = m->f0, a1 = m->f0 + m->hidden_local;
            if (int e = L.linear_slice(p + "gating.linear_in.weight", d, 2 * F, {{a0, a1}, {F + a0, F + a1}}, 0, d, m->hidden_local, &l.lin_in[0])) return e;
            wb += L.linear_bytes;
            if (int e = L.linear_slice(p + "gating.linear_out.weight", F, d, {{0, d}}, a0, a1, 0, &l.lin_out[0])) return e;
            wb += L.linear_bytes;
        }
    }
    if (int e = L.vec_f32("lm.out_norm.alpha", d, &m->out_norm)) return e;
    if (int e = L.linear("lm.text_linear.weight", d, c.text_card, &m->text_linear)) return e;
    wb += L.linear_bytes;

    // depformer (lm.h:371-385; lm_default.h:72-83 for the number of per-step weights)
    if (c.dep_q > 0) {
        const int dd = c.dep_dim;
        int nw = c.dep_q;
        if (c.schedule_len) { nw = 0; for (int i = 0; i < c.schedule_len; i++) nw = std::max(nw, c.schedule[i] + 1); }
        m->dep_nw = nw;
        m->dep_cap = c.dep_context ? c.dep_context : c.schedule_len;
        m->dep_in.resize(nw);
        std::vector<int64_t> dep_in_bytes(nw), layer_bytes(nw, 0);
        // the GGUF blocks of every depformer_in are kept on the device until the per-step concatenation below
        const GgufTensor *dep_in0 = f.find("lm.depformer_in.0.weight");
        const int dep_in_type = !dep_in0 ? 0 : (m->quantize && is_float_type(dep_in0->type)) ? m->quantize : dep_in0->type;
        const int64_t dep_in_rs = ggml_row_size(dep_in_type, d);
        uint8_t *cat_w = nullptr;
        if (dep_in_rs > 0 && tp_world == 1) CU(cudaMalloc((void **)&cat_w, (size_t)nw * dd * dep_in_rs));
        struct CatFree { uint8_t *p; ~CatFree() { if (p) cudaFree(p); } } cat_free{cat_w};
        for (int k = 0; k < nw; k++) {
            if (int e = L.linear("lm.depformer_in." + std::to_string(k) + ".weight", d, dd, &m->dep_in[k], 0, cat_w ? cat_w + (size_t)k * dd * dep_in_rs : nullptr)) return e;
            dep_in_bytes[k] = L.linear_bytes;
            if (m->dep_in[k].type != dep_in_type) m->stream_ok = false;
        }
        if (cat_w && m->stream_ok && d % 256 == 0) {
            // depformer_in[w_k] . t_out of ALL codebook steps does not depend on the chain: one matrix, one phase of the step kernel
            uint8_t *cat_k = nullptr;
            CU(cudaMalloc((void **)&cat_k, (size_t)c.dep_q * dd * dep_in_rs));
            CatFree cat_k_free{cat_k};
            for (int k = 0; k < c.dep_q; k++) {
                const int wsel = c.schedule_len ? c.schedule[k] : k, w = nw == 1 ? 0 : wsel;
                CU(cudaMemcpy(cat_k + (size_t)k * dd * dep_in_rs, cat_w + (size_t)w * dd * dep_in_rs, (size_t)dd * dep_in_rs, cudaMemcpyDeviceToDevice));
            }
            m->dep_in_all = m->dep_in[0];
            m->dep_in_all.rows = c.dep_q * dd;
            m->dep_in_all.qs = reinterpret_cast<const uint8_t *>(&m->dep_in_all);      // key only: this matrix exists in stream layout alone
            if (int e = upload_stream(m.get(), cat_k, dep_in_type, d, (int64_t)c.dep_q * dd, 0, m->dep_in_all.qs)) return e;
            CU(cudaDeviceSynchronize());
        } else m->stream_ok = false;
        // low-rank / demux depformer embeddings: table rows are [lr] wide and go through a small projection
        // (lm_utils.h:126-217; lm_default.h:196-214)
        const int de = c.dep_low_rank ? c.dep_low_rank : dd;
        m->dep_small = c.dep_low_rank || c.demux_second_stream;
        if (m->dep_small && (de > kSmallMaxK || de % 32)) return fail(MSX_ERR_FORMAT, "low-rank embedding width must be a multiple of 32, <= 2048");
        auto small = [&](const std::string &name, EmbTable *out) -> int {
            const GgufTensor *t = L.need(name);
            if (!t) return MSX_ERR_FORMAT;
            if (t->type != T_Q4_0 && t->type != T_Q8_0 && !(m->quantize && is_float_type(t->type)))
                return fail(MSX_ERR_FORMAT, name + ": small projections must be q4_0 or q8_0");
            if (m->quantize == T_Q4_K && is_float_type(t->type) && de % 256 == 0)
                return fail(MSX_ERR_FORMAT, name + ": the reference would make this projection q4_k (K % 256 == 0); small projections take q4_0 / q8_0");
            return L.table(name, de, dd, out);
        };
        if (int e = L.table("lm.depformer_text_emb.weight", de, c.text_card + 1, &m->dep_text_emb)) return e;
        if (c.demux_second_stream) {
            if (int e = small("lm.depformer_text_emb.out1.weight", &m->dep_text_out1)) return e;
            if (int e = small("lm.depformer_text_emb.out2.weight", &m->dep_text_out2)) return e;
        } else if (c.dep_low_rank) {
            if (int e = small("lm.depformer_text_emb.low_rank.weight", &m->dep_text_lr)) return e;
        }
        m->dep_emb.resize(c.dep_q - 1);
        m->dep_emb_lr.resize(c.dep_low_rank ? c.dep_q - 1 : 0);
        for (int k = 0; k < c.dep_q - 1; k++) {
            if (int e = L.table("lm.depformer_emb." + std::to_string(k) + ".weight", de, c.card + 1, &m->dep_emb[k])) return e;
            if (c.dep_low_rank) if (int e = small("lm.depformer_emb." + std::to_string(k) + ".low_rank.weight", &m->dep_emb_lr[k])) return e;
        }
        m->dep_layers.resize(c.dep_layers);
        for (int i = 0; i < c.dep_layers; i++) {
            LayerW &l = m->dep_layers[i];
            const std::string p = "lm.depformer.layers." + std::to_string(i) + ".";
            if (int e = L.vec_f32(p + "norm1.alpha", dd, &l.norm1)) return e;
            if (int e = L.vec_f32(p + "norm2.alpha", dd, &l.norm2)) return e;
            l.in_proj.resize(nw); l.out_proj.resize(nw); l.lin_in.resize(nw); l.lin_out.resize(nw);
            for (int k = 0; k < nw; k++) {
                const std::string ks = std::to_string(k);
                if (int e = L.linear(p + "self_attn.in_projs." + ks + ".weight", dd, 3 * dd, &l.in_proj[k])) return e;
                layer_bytes[k] += L.linear_bytes;
                if (int e = L.linear(p + "self_attn.out_projs." + ks + ".weight", dd, dd, &l.out_proj[k])) return e;
                layer_bytes[k] += L.linear_bytes;
                // per-step gating names: "gating.{k}.linear_in" (transformer.h:1057-1063); single-weight: "gating.linear_in"
                std::string gname = p + "gating." + ks + ".linear_in.weight", oname = p + "gating." + ks + ".linear_out.weight";
                if (nw == 1 && !f.find(gname)) { gname = p + "gating.linear_in.weight"; oname = p + "gating.linear_out.weight"; }
                const GgufTensor *t = L.need(gname);
                if (!t) return MSX_ERR_FORMAT;
                const int Fd = (int)(t->ne[1] / 2);
                if (i == 0 && k == 0) m->dep_hidden = Fd;
                if (Fd != m->dep_hidden) return fail(MSX_ERR_FORMAT, "inconsistent depformer hidden size");
                if (int e = L.linear(gname, dd, 2 * Fd, &l.lin_in[k], Fd)) return e;
                layer_bytes[k] += L.linear_bytes;
                if (int e = L.linear(oname, Fd, dd, &l.lin_out[k])) return e;
                layer_bytes[k] += L.linear_bytes;
            }
        }
        m->linears.resize(c.dep_q);
        for (int k = 0; k < c.dep_q; k++) {
            if (int e = L.linear("lm.linears." + std::to_string(k) + ".weight", dd, c.card, &m->linears[k])) return e;
            const int w = nw == 1 ? 0 : (c.schedule_len ? c.schedule[k] : k);
            wb += L.linear_bytes + dep_in_bytes[w] + layer_bytes[w];
        }
    }
    m->extra_heads.resize(c.extra_heads);
    for (int j = 0; j < c.extra_heads; j++)
        if (int e = L.linear("lm.extra_heads." + std::to_string(j) + ".weight", d, 0, &m->extra_heads[j])) return e;
    m->weight_bytes_per_frame = wb;
    // RoPE frequencies exactly as ggml_timestep_embedding computes them on the host CPU:
    // freq_j = expf(-logf(max_period) * j / half)   (rope.h:8-20)
    auto make_freq = [&](int dh, int max_period, const float **out) -> int {
        if (!max_period) return 0;
        const int half = dh / 2;
        std::vector<float> fr(half);
        for (int j = 0; j < half; j++) fr[j] = (float)expf(-logf((float)max_period) * j / half);
        void *p = nullptr;
        if (int e = dev_alloc(m.get(), &p, half * 4)) return e;
        CU(cudaMemcpy(p, fr.data(), half * 4, cudaMemcpyHostToDevice));
        *out = (const float *)p;
        return 0;
    };
    if (int e = make_freq(c.dim / c.num_heads, c.max_period, &m->rope_freq)) return e;
    if (c.dep_q > 0)
        if (int e = make_freq(c.dep_dim / c.dep_heads, c.dep_max_period, &m->dep_rope_freq)) return e;
    // voice conditioners (tts.h:16-35): optional — files made for externally computed conditioning do not carry them
    const std::string cp = "lm.condition_provider.conditioners.";
    if (c.cross_attention && f.find(cp + "cfg.embed.weight")) {
        if (int e = L.float_tensor(cp + "cfg.embed.weight", 0, &m->cfg_embed)) return e;
        if (int e = L.float_tensor(cp + "cfg.output_proj.weight", m->cfg_embed.ne0, &m->cfg_proj)) return e;
        if (int e = L.float_tensor(cp + "control.embed.weight", 0, &m->control_embed)) return e;
        if (int e = L.float_tensor(cp + "control.output_proj.weight", m->control_embed.ne0, &m->control_proj)) return e;
        if (int e = L.float_tensor(cp + "speaker_wavs.learnt_padding", d, &m->spk_pad)) return e;
        if (int e = L.float_tensor(cp + "speaker_wavs.output_proj.weight", 0, &m->spk_proj)) return e;
        if (m->cfg_proj.ne1 != d || m->control_proj.ne1 != d || m->spk_proj.ne1 != d || m->cfg_embed.ne1 < 3)
            return fail(MSX_ERR_FORMAT, "conditioner projections must map to dim; cfg.embed needs >= 3 rows");
        if (int e = make_freq(d, 10000, &m->cond_freq)) return e;          // ggml_timestep_embedding(positions, dim, 10000)
        m->has_conditioners = true;
    }
    if (m->staging) { cudaFree(m->staging); m->staging = nullptr; m->staging_bytes = 0; }
    if (m->qstaging) { cudaFree(m->qstaging); m->qstaging = nullptr; m->qstaging_bytes = 0; }
    *out = m.release();
    return 0;
}

extern "C" void msx_model_free(msx_model *m) { delete m; }
extern "C" int msx_model_config(const msx_model *m, msx_config *out) {
    if (!m || !out) return fail(MSX_ERR_ARG, "null argument");
    *out = m->cfg; return 0;
}
extern "C" int64_t msx_model_weight_bytes_per_frame(const msx_model *m) { return m ? m->weight_bytes_per_frame : 0; }
extern "C" int64_t msx_model_device_bytes(const msx_model *m) { return m ? m->device_bytes : 0; }
extern "C" int msx_model_device(const msx_model *m) { return m ? m->device : -1; }

// -------------------------------------------------------------------------------------------------
// kernel launch helpers
// -------------------------------------------------------------------------------------------------
namespace {

// kernel families, for msx_profile_frame()
enum Family : int {
    FAM_EMBED = 0, FAM_IN_PROJ, FAM_ATTN, FAM_OUT_PROJ, FAM_LIN_IN, FAM_LIN_OUT, FAM_TEXT_HEAD, FAM_FINALIZE,
    FAM_DEP_IN, FAM_DEP_IN_PROJ, FAM_DEP_ATTN, FAM_DEP_OUT_PROJ, FAM_DEP_LIN_IN, FAM_DEP_LIN_OUT, FAM_DEP_HEAD, FAM_DEP_FINALIZE,
    FAM_STEP_TEMPORAL, FAM_STEP_DEPFORMER,
    FAM_COUNT
};
const char *kFamilyNames[FAM_COUNT] = {
    "embed", "in_proj", "attn", "out_proj", "linear_in", "linear_out", "text_head", "finalize",
    "dep_in", "dep_in_proj", "dep_attn", "dep_out_proj", "dep_linear_in", "dep_linear_out", "dep_head", "dep_finalize",
    "step_temporal", "step_depformer"};

int tiles_of(msx_model *m, const QLinear &w, QTiles *out);     // batch.inl
int ensure_all_tiles(msx_model *m);

// tiles of a small matrix one CTA takes at least: the grid shrinks below one CTA per SM for tiny matrices (measured on B200,
// moshi 7B frame: 4 -> 2.153 ms, 8 -> 2.152, 16 -> 2.231, 32 -> 2.649, 64 -> 3.616)
constexpr int kTilesPerCta = 8;

struct Launcher {
    cudaStream_t st;
    int num_sms;
    int count = 0;
    msx_model *model = nullptr;
    cudaError_t err = cudaSuccess;
    // optional per-launch timing (eager mode only): events[i], events[i+1] bracket launch i
    std::vector<cudaEvent_t> *events = nullptr;
    std::vector<int> *families = nullptr;
    int fam = 0;
    void begin() {
        if (events && events->empty()) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); events->push_back(e); }
    }
    void check() {
        if (err == cudaSuccess) err = cudaGetLastError();
        count++;
        if (events) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); events->push_back(e); families->push_back(fam); }
    }

    // cudaLaunchKernelEx with the programmatic-stream-serialization attribute (captured into the graph as a
    // programmatic dependency edge): the kernel may begin before its predecessor has drained
    template <typename K, typename... Args>
    void launch_pdl(K kernel, dim3 grid, dim3 block, size_t smem, Args... args) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
        if (err == cudaSuccess) err = e;
    }
    bool pdl = true;

    void gemv(const GemvArgs &a, int pro, int epi, int family = 0) {
        fam = family; begin();
        const int tr = tile_rows(a.w.gs);
        const int n_tiles = (a.w.rows + tr - 1) / tr;
        const int grid = std::max(1, std::min(num_sms, (n_tiles + kTilesPerCta - 1) / kTilesPerCta));
        const int smem = gemv_smem_bytes(a.w.type, a.w.K);
        if (a.tp) {          // tensor-parallel partial sums pushed to the peers: separate instantiations
            if (a.w.type == T_Q4_K) {
                if (a.w.gs == 32) launch_pdl(gemv_kernel<12, 32, true>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
                else launch_pdl(gemv_kernel<12, 16, true>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
            } else {
                if (a.w.gs == 32) launch_pdl(gemv_kernel<8, 32, true>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
                else launch_pdl(gemv_kernel<8, 16, true>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
            }
        } else if ((epi == EPI_STORE || epi == EPI_RESID || epi == EPI_GATE) && !a.xparts && !a.norm_out) {
            // the lean kernels: store / residual / gate epilogues only (4 of every 5 launches of a frame)
            if (a.w.type == T_Q4_K) {
                if (a.w.gs == 32) launch_pdl(gemv_kernel<12, 32, false, true>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
                else launch_pdl(gemv_kernel<12, 16, false, true>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
            } else {
                if (a.w.gs == 32) launch_pdl(gemv_kernel<8, 32, false, true>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
                else launch_pdl(gemv_kernel<8, 16, false, true>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
            }
        } else if (a.w.type == T_Q4_K) {
            if (a.w.gs == 32) launch_pdl(gemv_kernel<12, 32>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
            else launch_pdl(gemv_kernel<12, 16>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
        } else {
            if (a.w.gs == 32) launch_pdl(gemv_kernel<8, 32>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
            else launch_pdl(gemv_kernel<8, 16>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
        }
        check();
    }

    // n GEMVs of one shape over the same input in one launch (lean store epilogue)
    void gemv_multi(const GemvArgs &a, const GemvMulti &mm, int n, int family) {
        fam = family; begin();
        const int smem = gemv_smem_bytes(a.w.type, a.w.K);
        const dim3 grid(n * mm.per), block(kGemvThreads);
        if (a.w.type == T_Q4_K) {
            if (a.w.gs == 32) launch_pdl(gemv_multi_kernel<12, 32>, grid, block, smem, a, mm, (int)PRO_PLAIN, (int)EPI_STORE);
            else launch_pdl(gemv_multi_kernel<12, 16>, grid, block, smem, a, mm, (int)PRO_PLAIN, (int)EPI_STORE);
        } else {
            if (a.w.gs == 32) launch_pdl(gemv_multi_kernel<8, 32>, grid, block, smem, a, mm, (int)PRO_PLAIN, (int)EPI_STORE);
            else launch_pdl(gemv_multi_kernel<8, 16>, grid, block, smem, a, mm, (int)PRO_PLAIN, (int)EPI_STORE);
        }
        check();
    }
    int gemv_ctas(const QLinear &w) const {
        const int tr = tile_rows(w.gs);
        return std::max(1, std::min(num_sms, ((w.rows + tr - 1) / tr + kTilesPerCta - 1) / kTilesPerCta));
    }

    // fused local attention + out_proj (tiny rings)
    void gemv_local_attn(const GemvArgs &g, const AttnArgs &a, int heads, int dh, int pro, int epi, int family = 0) {
        fam = family; begin();
        const int tr = tile_rows(g.w.gs);
        const int n_tiles = (g.w.rows + tr - 1) / tr;
        const int grid = std::max(1, std::min(num_sms, (n_tiles + kTilesPerCta - 1) / kTilesPerCta));
        const int region = (gemv_smem_bytes(g.w.type, g.w.K) + 15) / 16 * 16;
        const int smem = local_attn_smem_bytes(region, a.dim, dh);
#define MSX_LA(WT, LN, DH) launch_pdl(gemv_local_attn_kernel<WT, LN, DH>, dim3(grid), dim3(kGemvThreads), smem, g, a, heads, pro, epi, region)
        if (g.w.type == T_Q4_K) {
            if (g.w.gs == 32) { if (dh == 64) MSX_LA(12, 32, 64); else MSX_LA(12, 32, 128); }
            else { if (dh == 64) MSX_LA(12, 16, 64); else MSX_LA(12, 16, 128); }
        } else {
            if (g.w.gs == 32) { if (dh == 64) MSX_LA(8, 32, 64); else MSX_LA(8, 32, 128); }
            else { if (dh == 64) MSX_LA(8, 16, 64); else MSX_LA(8, 16, 128); }
        }
#undef MSX_LA
        check();
    }

    void layer_norm(const float *x, const float *w, const float *b, float *y, int n, float eps, int family) {
        LayerNormArgs a; a.x = x; a.w = w; a.b = b; a.y = y; a.n = n; a.eps = eps;
        fam = family; begin();
        launch_pdl(layer_norm_kernel, dim3(1), dim3(kLnThreads), 0, a);
        check();
    }
    void cross_attn(const CrossAttnArgs &a, int heads, int dh, int family) {
        fam = family; begin();
        if (dh == 128) launch_pdl(cross_attn_kernel<128>, dim3(heads), dim3(kCrossThreads), (size_t)cross_attn_smem<128>(a.tc), a);
        else launch_pdl(cross_attn_kernel<64>, dim3(heads), dim3(kCrossThreads), (size_t)cross_attn_smem<64>(a.tc), a);
        check();
    }
    void small_linear(const SmallLinearArgs &a, int family) {
        fam = family; begin();
        launch_pdl(small_linear_kernel, dim3((a.w.rows + kSmallThreads - 1) / kSmallThreads), dim3(kSmallThreads), 0, a);
        check();
    }

    void attn(const AttnArgs &a, int heads, int dh, int split, int family = 0, int n_streams = 1) {
        fam = family; begin();
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(split, heads, n_streams);
        cfg.blockDim = dim3(kThreads, 1, 1);
        cfg.stream = st;
        cudaLaunchAttribute at[2];
        int na = 0;
        if (pdl) { at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[na].val.programmaticStreamSerializationAllowed = 1; na++; }
        if (split > 1) { at[na].id = cudaLaunchAttributeClusterDimension; at[na].val.clusterDim.x = split; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1; na++; }
        cfg.attrs = at; cfg.numAttrs = na;
        cudaError_t e;
        if (dh == 128) {
            cfg.dynamicSmemBytes = attn_smem_bytes<128>(a.cap, split);
            e = split > 1 ? cudaLaunchKernelEx(&cfg, attn_kernel<128, true>, a) : cudaLaunchKernelEx(&cfg, attn_kernel<128, false>, a);
        } else {
            cfg.dynamicSmemBytes = attn_smem_bytes<64>(a.cap, split);
            e = split > 1 ? cudaLaunchKernelEx(&cfg, attn_kernel<64, true>, a) : cudaLaunchKernelEx(&cfg, attn_kernel<64, false>, a);
        }
        if (err == cudaSuccess) err = e;
        check();
    }
};

// rows [row0, row0 + rows) of a repacked linear (torch_nn_linear_view, torch.h:103-118)
QLinear linear_rows(const QLinear &w, int row0, int rows) {
    QLinear v = w;
    v.rows = rows;
    if (w.type == T_Q4_K) {
        v.qs = w.qs + (size_t)row0 * (w.K >> 1);
        v.sc = w.sc + (size_t)row0 * (w.K >> 6);
        v.dd = reinterpret_cast<const uint32_t *>(w.dd) + (size_t)row0 * (w.K >> 8);
    } else {
        v.qs = w.qs + (size_t)row0 * w.K;
        v.dd = reinterpret_cast<const uint16_t *>(w.dd) + (size_t)row0 * (w.K >> 5);
    }
    return v;
}

int attn_split_for(int heads, int cap, int num_sms) {
    // short rings: at most one CTA per SM; long rings (KV streaming dominates): up to two CTAs per SM
    const int budget = cap > 1024 ? 2 * num_sms : num_sms;
    int s = 1;
    while (s * 2 <= kAttnMaxSplit && heads * s * 2 <= budget && cap / (s * 2) >= 1) s *= 2;
    return s;
}

}  // namespace

// -------------------------------------------------------------------------------------------------
// stream
// -------------------------------------------------------------------------------------------------
static void free_prefill(struct msx_batch *b);
struct StepBuffers {      // LL vectors of the persistent step kernel (step_kernel.cuh), one set per stream
    sk::LL *xA = nullptr, *xB = nullptr, *qkv = nullptr, *ctx = nullptr, *gate = nullptr, *tkeys = nullptr;
    sk::LL *scores = nullptr;
    sk::LL *dep_d = nullptr, *dxA = nullptr, *dxB = nullptr, *dqkv = nullptr, *dctx = nullptr, *dgate = nullptr, *dkeys = nullptr;
};
struct msx_stream {
    msx_model *m = nullptr;
    int cap = 0;
    int attn_split = 1;
    cudaStream_t st = nullptr;
    Ctrl *ctrl = nullptr;            // device
    int32_t *h_in = nullptr;         // pinned: text_override, tokens[40], force[40], pad
    int32_t *h_out = nullptr;        // pinned: out_tokens[41]
    int32_t *h_err = nullptr;        // pinned: Ctrl::error
    uint16_t *kc = nullptr, *vc = nullptr, *dkc = nullptr, *dvc = nullptr;
    float *x = nullptr, *qkv = nullptr, *ctx = nullptr, *gate = nullptr, *tout = nullptr, *text_logits = nullptr;
    float *dx = nullptr, *dqkv = nullptr, *dctx = nullptr, *dgate = nullptr, *audio_logits = nullptr, *vad_logits = nullptr;
    float *rope_cs = nullptr;        // [Dh] cos | sin of the current temporal position
    // TTS family
    float *cond_sum = nullptr;       // [dim] or null
    float *kv_cross = nullptr;       // [L][tc][2*dim] f32 cross-attention memory
    int tc = 0;
    float *cnx = nullptr, *cq = nullptr, *cctx = nullptr;     // layer-norm output, cross q, cross context [dim]
    float *demux_l = nullptr, *demux_r = nullptr, *demux_y1 = nullptr, *demux_y2 = nullptr;   // [dim]
    float *embed_in = nullptr;       // [dim] voice-embedding prompt row (msx_step_temporal_embedding)
    float *dep_e = nullptr;          // [dep_dim] embedding of the previous token after its low-rank / demux projection
    float *dep_d = nullptr;          // [dep_q][dep_dim] depformer_in[k] . t_out of every step, computed by one launch up front
    // tensor parallelism: double partial sums of out_proj / linear_out, all-reduced with NCCL inside the graph
    double *tp_partial = nullptr;    // [dim]
    void *nccl_comm = nullptr;
    // peer-memory all-reduce (msx_stream_tp_export / _connect): arena mapped by the peers through CUDA IPC
    uint8_t *tp_arena = nullptr;     // [inbox 2 x world x dim x {lo, seq, hi, seq} | epoch]
    TpCtx *d_tp = nullptr;           // device copy of the context
    uint32_t *tp_frame_ctr = nullptr;
    std::vector<void *> tp_peer_maps;
    bool tp_p2p = false;
    struct msx_batch *prefill = nullptr;   // batched-T prompt prefill context (batch.inl), created on first use
    bool embed_override_next = false;
    int32_t *d_feed = nullptr;       // msx_run_resident_async
    cudaGraphExec_t g_temporal = nullptr, g_depformer = nullptr;
    int launches_temporal = 0, launches_depformer = 0;
    // sampling (sampling.h:46-64): temperature <= 0 = greedy; noise = Exp(1) draws supplied by the host per frame
    float temp_text = 0.f, temp_audio = 0.f;
    int top_k_text = 25, top_k_audio = 250;
    float *d_noise = nullptr, *h_noise = nullptr, *d_probs = nullptr;
    int noise_floats = 0;
    bool noise_fresh = false;
    // persistent step kernel (step_kernel.cuh): phase programs of the two stacks, LL vectors, launch counter
    StepBuffers step_buf;
    sk::StepPhase *d_prog_t = nullptr, *d_prog_d = nullptr;
    int n_prog_t = 0, n_prog_d = 0;
    std::vector<int> prog_fam_t, prog_fam_d;   // kernel family of every phase (timeline)
    uint32_t *d_epoch = nullptr;
    bool step_kernel = false;        // the graphs hold one cooperative step_kernel launch each
    int flags = 0;
    int host_offset = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<void *> allocs;

    ~msx_stream() {
        if (m) cudaSetDevice(m->device);
        if (prefill) free_prefill(prefill);
        if (g_temporal) cudaGraphExecDestroy(g_temporal);
        if (g_depformer) cudaGraphExecDestroy(g_depformer);
        if (nccl_comm) nccl().CommDestroy(nccl_comm);
        for (void *p : tp_peer_maps) cudaIpcCloseMemHandle(p);
        for (void *p : allocs) cudaFree(p);
        if (h_in) cudaFreeHost(h_in);
        if (h_out) cudaFreeHost(h_out);
        if (h_err) cudaFreeHost(h_err);
        if (h_noise) cudaFreeHost(h_noise);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (st) cudaStreamDestroy(st);
    }
};

namespace {

int salloc(msx_stream *s, void **p, size_t bytes) {
    CU(cudaMalloc(p, std::max<size_t>(bytes, 16)));
    CU(cudaMemset(*p, 0, std::max<size_t>(bytes, 16)));
    s->allocs.push_back(*p);
    return 0;
}

size_t kv_elems(const msx_stream *s) { return (size_t)s->m->cfg.num_layers * s->cap * s->m->adim; }
size_t dkv_elems(const msx_stream *s) { return (size_t)s->m->cfg.dep_layers * s->m->dep_cap * s->m->cfg.dep_dim; }

// one transformer layer (transformer.h:910-1039) as 5 launches
void enqueue_layer(Launcher &L, const msx_stream *s, const LayerW &lw, int w, bool temporal, int layer, int pos_const) {
    const msx_model *m = s->m; const msx_config &c = m->cfg;
    const int dim = temporal ? c.dim : c.dep_dim;
    // tensor parallelism shards the temporal layers only: this rank's heads / hidden slice (adim == dim for one rank)
    const bool tp = temporal && m->tp_world > 1;
    const int heads = temporal ? m->heads_local : c.dep_heads;
    const int adim = temporal ? m->adim : c.dep_dim;
    const int cap = temporal ? s->cap : m->dep_cap;
    float *x = temporal ? s->x : s->dx, *qkv = temporal ? s->qkv : s->dqkv, *ctx = temporal ? s->ctx : s->dctx, *gate = temporal ? s->gate : s->dgate;
    // out[dim] += W_shard . in : partial sums in double -> NCCL all-reduce (sum) -> rounded once into the residual stream
    auto reduce_into_x = [&](GemvArgs &gg, int family) {
        if (s->tp_p2p) {
            // GEMV pushes its partial sums into every rank's inbox over NVLink; the consumer waits for the flags
            const int idx = 2 * layer + (family == FAM_LIN_OUT ? 1 : 0);
            gg.out_f64 = nullptr; gg.out = nullptr; gg.tp = s->d_tp; gg.tp_idx = idx;
            L.gemv(gg, PRO_PLAIN, EPI_STORE_F64, family);
            gg.tp = nullptr;
            L.fam = family; L.begin();
            L.launch_pdl(tp_apply_p2p_kernel, dim3(1), dim3(1024), 0, x, (const TpCtx *)s->d_tp, idx);
            L.check();
            return;
        }
        gg.out_f64 = s->tp_partial; gg.out = nullptr;
        L.gemv(gg, PRO_PLAIN, EPI_STORE_F64, family);
        L.fam = family; L.begin();
        const int rc = nccl().AllReduce(s->tp_partial, s->tp_partial, (size_t)dim, kNcclFloat64, kNcclSum, s->nccl_comm, L.st);
        if (rc != 0 && L.err == cudaSuccess) L.err = cudaErrorUnknown;
        L.check();
        L.fam = family; L.begin();
        L.launch_pdl(tp_apply_kernel, dim3((dim + 255) / 256), dim3(256), 0, x, (const double *)s->tp_partial, dim);
        L.check();
    };
    GemvArgs g;
    g.ctrl = s->ctrl; g.eps = 1e-8f;
    // x -> rms_norm1 -> in_proj -> qkv
    g.w = lw.in_proj[w]; g.x = x; g.alpha = lw.norm1; g.out = qkv;
    L.gemv(g, PRO_RMS, EPI_STORE, temporal ? FAM_IN_PROJ : FAM_DEP_IN_PROJ);
    // rope + kv insert + attention
    AttnArgs a;
    a.qkv = qkv; a.ctx = ctx; a.ctrl = s->ctrl; a.pos_const = pos_const; a.cap = cap; a.dim = adim;
    a.max_period = temporal ? c.max_period : c.dep_max_period;
    a.rope_freq = temporal ? m->rope_freq : m->dep_rope_freq;
    a.rope_cs = (temporal && c.max_period) ? s->rope_cs : nullptr;
    a.small_ctx = 32;        // up to 32 valid slots one CTA per head handles the ring alone (no cluster barriers)
    const size_t lstride = (size_t)cap * adim;
    a.kc = (temporal ? s->kc : s->dkc) + (size_t)layer * lstride;
    a.vc = (temporal ? s->vc : s->dvc) + (size_t)layer * lstride;
    if (!temporal && cap <= 64) {
        // tiny ring: every CTA recomputes the attention of all heads in its prologue -> one launch
        g.w = lw.out_proj[w]; g.x = nullptr; g.alpha = nullptr; g.out = x;
        L.gemv_local_attn(g, a, heads, adim / heads, PRO_PLAIN, EPI_RESID, FAM_DEP_OUT_PROJ);
    } else {
        L.attn(a, heads, adim / heads, temporal ? s->attn_split : 1, temporal ? FAM_ATTN : FAM_DEP_ATTN);
        // out_proj + residual
        g.w = lw.out_proj[w]; g.x = ctx; g.alpha = nullptr; g.out = x;
        if (tp) reduce_into_x(g, FAM_OUT_PROJ);
        else L.gemv(g, PRO_PLAIN, EPI_RESID, temporal ? FAM_OUT_PROJ : FAM_DEP_OUT_PROJ);
    }
    if (temporal && lw.cross_in.qs && s->kv_cross && s->tc > 0) {
        // x += cross_attention(layer_norm(x)) over the conditioning memory (transformer.h:936-943, 714-762)
        L.layer_norm(x, lw.norm_cross_w, lw.norm_cross_b, s->cnx, dim, 0.0f, FAM_ATTN);
        g.w = linear_rows(lw.cross_in, 0, dim); g.x = s->cnx; g.alpha = nullptr; g.out = s->cq;
        L.gemv(g, PRO_PLAIN, EPI_STORE, FAM_IN_PROJ);
        CrossAttnArgs ca;
        ca.q = s->cq; ca.kv = s->kv_cross + (size_t)layer * s->tc * 2 * dim; ca.ctx = s->cctx; ca.tc = s->tc; ca.dim = dim;
        L.cross_attn(ca, heads, dim / heads, FAM_ATTN);
        g.w = lw.cross_out; g.x = s->cctx; g.alpha = nullptr; g.out = x;
        L.gemv(g, PRO_PLAIN, EPI_RESID, FAM_OUT_PROJ);
    }
    // rms_norm2 -> linear_in -> silu gate
    g.w = lw.lin_in[w]; g.x = x; g.alpha = lw.norm2; g.out = gate;
    L.gemv(g, PRO_RMS, EPI_GATE, temporal ? FAM_LIN_IN : FAM_DEP_LIN_IN);
    // linear_out + residual
    g.w = lw.lin_out[w]; g.x = gate; g.alpha = nullptr; g.out = x;
    if (tp) reduce_into_x(g, FAM_LIN_OUT);
    else L.gemv(g, PRO_PLAIN, EPI_RESID, temporal ? FAM_LIN_OUT : FAM_DEP_LIN_OUT);
}

void enqueue_temporal(Launcher &L, const msx_stream *s) {
    const msx_model *m = s->m; const msx_config &c = m->cfg;
    EmbedArgs e;
    e.tables = m->d_emb; e.n_tables = c.n_q + 1; e.dim = c.dim; e.ctrl = s->ctrl; e.x = s->x;
    if (c.max_period) { e.rope_cs = s->rope_cs; e.rope_freq = m->rope_freq; e.dh = c.dim / c.num_heads; }
    if (c.demux_second_stream) {
        // text embedding = out1(row[left]) + out2(row[right]) * right_scale (lm_utils.h:42-86)
        DemuxRowsArgs dr;
        dr.table = m->emb[0]; dr.ctrl = s->ctrl; dr.num_embeddings = c.text_card + 1; dr.left = s->demux_l; dr.right = s->demux_r;
        L.fam = FAM_EMBED; L.begin();
        L.launch_pdl(demux_rows_kernel, dim3((c.dim + 255) / 256), dim3(256), 0, dr);
        L.check();
        GemvArgs g1;
        g1.ctrl = s->ctrl; g1.w = m->text_out1; g1.x = s->demux_l; g1.out = s->demux_y1;
        L.gemv(g1, PRO_PLAIN, EPI_STORE, FAM_EMBED);
        g1.w = m->text_out2; g1.x = s->demux_r; g1.out = s->demux_y2;
        L.gemv(g1, PRO_PLAIN, EPI_STORE, FAM_EMBED);
        e.text_pre1 = s->demux_y1; e.text_pre2 = s->demux_y2; e.num_embeddings = c.text_card + 1;
    }
    e.cond_sum = s->cond_sum;
    e.embed_in = s->embed_in;
    L.fam = FAM_EMBED; L.begin();
    L.launch_pdl(embed_kernel, dim3((c.dim + kThreads - 1) / kThreads), dim3(kThreads), 0, e);
    L.check();
    for (int l = 0; l < c.num_layers; l++) enqueue_layer(L, s, m->layers[l], 0, true, l, -1);
    // out_norm -> transformer_out (kept for depformer / VAD) -> text_linear -> greedy token (lm.h:671-674, 864-868)
    GemvArgs g;
    g.ctrl = s->ctrl; g.eps = 1e-8f;
    g.w = m->text_linear; g.x = s->x; g.alpha = m->out_norm; g.norm_out = s->tout; g.out = s->text_logits;
    g.key = &s->ctrl->text_key;
    L.gemv(g, PRO_RMS, EPI_ARGMAX, FAM_TEXT_HEAD);
    if (s->temp_text > 0.f) {      // moshi_sample_token: softmax(l / temp) -> top-k -> p / Exp(1) -> argmax
        SampleArgs sa;
        sa.logits = s->text_logits; sa.n = c.text_card; sa.k = std::min(std::min(s->top_k_text, c.text_card), kSampleMaxK);
        sa.inv_temp = 1.f / s->temp_text; sa.noise = s->d_noise; sa.key = &s->ctrl->text_key; sa.probs = s->d_probs;
        L.fam = FAM_TEXT_HEAD; L.begin();
        L.launch_pdl(sample_kernel, dim3(1), dim3(kSampleThreads), 0, sa);
        L.check();
    }
    L.fam = FAM_FINALIZE;
    L.launch_pdl(finalize_temporal_kernel, dim3(1), dim3(32), 0, s->ctrl, c.dep_q > 0 ? 1 : 0, s->tp_p2p ? s->tp_frame_ctr : (uint32_t *)nullptr);
    L.check();
}

void enqueue_depformer(Launcher &L, const msx_stream *s) {
    const msx_model *m = s->m; const msx_config &c = m->cfg;
    auto weights_of = [&](int k) { const int wsel = c.schedule_len ? c.schedule[k] : k; return m->dep_nw == 1 ? 0 : wsel; };   // lm.h:457-462, transformer.h:74-83
    // depformer_in[w_k](transformer_out) does not depend on the codebook chain: all dep_q of them in ONE launch up front
    // (same kernel body, same arithmetic); the steps then only add the previous token's embedding (lm.h:464-467, 494-516)
    const bool hoist = c.dep_q <= kGemvMultiMax;
    if (hoist) {
        GemvMulti mm;
        for (int k = 0; k < c.dep_q; k++) {
            const QLinear &w = m->dep_in[weights_of(k)];
            mm.qs[k] = w.qs; mm.sc[k] = w.sc; mm.dd[k] = w.dd; mm.out[k] = s->dep_d + (size_t)k * c.dep_dim;
        }
        GemvArgs g;
        g.ctrl = s->ctrl; g.eps = 1e-8f; g.w = m->dep_in[weights_of(0)]; g.x = s->tout; g.out = s->dep_d;
        mm.per = L.gemv_ctas(g.w);
        L.gemv_multi(g, mm, c.dep_q, FAM_DEP_IN);
    }
    for (int k = 0; k < c.dep_q; k++) {
        const int w = weights_of(k);
        const float *dk = s->dep_d + (size_t)k * c.dep_dim;
        GemvArgs g;
        g.ctrl = s->ctrl; g.eps = 1e-8f;
        g.w = m->dep_in[w]; g.x = s->tout; g.out = s->dx;
        if (m->dep_small) {
            // previous token's embedding through its low-rank / demux projection first (lm_utils.h:42-66, 155-168, 209-217)
            SmallLinearArgs sl;
            sl.ctrl = s->ctrl; sl.step = k; sl.out = s->dep_e; sl.num_embeddings = c.text_card + 1;
            if (k == 0 && c.demux_second_stream) {
                sl.table = m->dep_text_emb; sl.w = m->dep_text_out1; sl.mode = 2;
                L.small_linear(sl, FAM_DEP_IN);
                sl.w = m->dep_text_out2; sl.mode = 3;
                if (hoist) { sl.addvec = dk; sl.dst = s->dx; }
                L.small_linear(sl, FAM_DEP_IN);
            } else {
                sl.table = k == 0 ? m->dep_text_emb : m->dep_emb[k - 1];
                sl.w = k == 0 ? m->dep_text_lr : m->dep_emb_lr[k - 1];
                sl.mode = k == 0 ? 0 : 1;
                if (hoist) { sl.addvec = dk; sl.dst = s->dx; }
                L.small_linear(sl, FAM_DEP_IN);
            }
            if (!hoist) {
                g.addvec = s->dep_e;
                L.gemv(g, PRO_PLAIN, EPI_ADD_VEC, FAM_DEP_IN);
            }
        } else if (hoist) {
            L.fam = FAM_DEP_IN; L.begin();
            L.launch_pdl(dep_embed_add_kernel, dim3((c.dep_dim + 255) / 256), dim3(256), 0, (const Ctrl *)s->ctrl, dk,
                         k == 0 ? m->dep_text_emb : m->dep_emb[k - 1], k, s->dx, (int)c.dep_dim);
            L.check();
        } else {
            g.emb = k == 0 ? m->dep_text_emb : m->dep_emb[k - 1];
            g.emb_step = k;
            L.gemv(g, PRO_PLAIN, EPI_ADD_EMB, FAM_DEP_IN);
        }
        for (int l = 0; l < c.dep_layers; l++) enqueue_layer(L, s, m->dep_layers[l], w, false, l, k);
        // linears[k] -> logits -> greedy token (no final norm, lm.h:472)
        GemvArgs h;
        h.ctrl = s->ctrl;
        h.w = m->linears[k]; h.x = s->dx; h.out = s->audio_logits + (size_t)k * c.card; h.key = &s->ctrl->audio_key[k];
        L.gemv(h, PRO_PLAIN, EPI_ARGMAX, FAM_DEP_HEAD);
        if (s->temp_audio > 0.f) {
            const int kk = std::min(std::min(s->top_k_audio, c.card), kSampleMaxK);
            SampleArgs sa;
            sa.logits = h.out; sa.n = c.card; sa.k = kk; sa.inv_temp = 1.f / s->temp_audio;
            sa.noise = s->d_noise + kSampleMaxK + (size_t)k * kSampleMaxK; sa.key = &s->ctrl->audio_key[k]; sa.probs = s->d_probs;
            L.fam = FAM_DEP_HEAD; L.begin();
            L.launch_pdl(sample_kernel, dim3(1), dim3(kSampleThreads), 0, sa);
            L.check();
        }
    }
    L.fam = FAM_DEP_FINALIZE;
    L.launch_pdl(finalize_depformer_kernel, dim3(1), dim3(64), 0, s->ctrl, (int)c.dep_q);
    L.check();
}

template <typename F>
int capture(msx_stream *s, F &&body, cudaGraphExec_t *exec, int *launches) {
    Launcher L{s->st, s->m->num_sms};
    L.model = s->m;
    CU(cudaStreamBeginCapture(s->st, cudaStreamCaptureModeThreadLocal));
    body(L);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(s->st, &graph);
    if (L.err != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return fail(MSX_ERR_CUDA, std::string("kernel launch failed during capture: ") + cudaGetErrorString(L.err)); }
    if (e != cudaSuccess) return fail(MSX_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(MSX_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    *launches = L.count;
    return 0;
}

int set_smem_attrs() {
    const int big = 220 * 1024;   // dynamic part; the kernels also have a little static shared memory
    CU(cudaFuncSetAttribute(gemv_kernel<12, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_multi_kernel<12, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_multi_kernel<12, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_multi_kernel<8, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_multi_kernel<8, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_kernel<12, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_kernel<8, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_kernel<8, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_kernel<12, 32, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_kernel<12, 16, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_kernel<8, 32, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_kernel<8, 16, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_kernel<12, 32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_kernel<12, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_kernel<8, 32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_kernel<8, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_local_attn_kernel<12, 32, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_local_attn_kernel<12, 32, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_local_attn_kernel<12, 16, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_local_attn_kernel<12, 16, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_local_attn_kernel<8, 32, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_local_attn_kernel<8, 32, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_local_attn_kernel<8, 16, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(gemv_local_attn_kernel<8, 16, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(sk::step_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, sk::kSmemBytes));
    CU(cudaFuncSetAttribute(sk::step_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, sk::kSmemBytes));
    // all kernels stay below the 48 KB default except long-context attention with split 1
    CU(cudaFuncSetAttribute(attn_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CU(cudaFuncSetAttribute(attn_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CU(cudaFuncSetAttribute(attn_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CU(cudaFuncSetAttribute(attn_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    return 0;
}

}  // namespace

#include "step_program.inl"

static int build_graphs(msx_stream *sp);

extern "C" int msx_stream_create(msx_model *m, int context_override, msx_stream **out) {
    return msx_stream_create_ex(m, context_override, 0, out);
}

extern "C" int msx_tp_unique_id(uint8_t *out128) {
    if (!out128) return fail(MSX_ERR_ARG, "null argument");
    Nccl &n = nccl();
    if (!n.ok) return fail(MSX_ERR_STATE, n.why);
    Nccl::UniqueId id;
    const int rc = n.GetUniqueId(&id);
    if (rc != 0) return fail(MSX_ERR_CUDA, std::string("ncclGetUniqueId: ") + n.GetErrorString(rc));
    memcpy(out128, id.internal, 128);
    return 0;
}

static int stream_create_impl(msx_model *m, int context_override, int flags, const uint8_t *nccl_id, msx_stream **out);

extern "C" int msx_stream_create_ex(msx_model *m, int context_override, int flags, msx_stream **out) {
    return stream_create_impl(m, context_override, flags, nullptr, out);
}
// tensor-parallel stream: collective over the ranks of the model's tensor-parallel group (every rank calls it with the
// same 128-byte id obtained from msx_tp_unique_id on one rank)
extern "C" int msx_stream_create_tp(msx_model *m, int context_override, const uint8_t *nccl_id, msx_stream **out) {
    return stream_create_impl(m, context_override, 0, nccl_id, out);
}

static int stream_create_impl(msx_model *m, int context_override, int flags, const uint8_t *nccl_id, msx_stream **out) {
    if (!m || !out) return fail(MSX_ERR_ARG, "null argument");
    *out = nullptr;
    if (m->tp_world > 1 && !nccl_id) return fail(MSX_ERR_STATE, "tensor-parallel model: create its streams with msx_stream_create_tp");
    CU(cudaSetDevice(m->device));
    if (int e = set_smem_attrs()) return e;
    std::unique_ptr<msx_stream> s(new msx_stream);
    s->m = m; s->flags = flags;
    const msx_config &c = m->cfg;
    s->cap = context_override > 0 ? std::min(context_override, c.context) : c.context;
    s->attn_split = attn_split_for(m->heads_local, s->cap, m->num_sms);
    CU(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
    CU(cudaEventCreate(&s->ev0)); CU(cudaEventCreate(&s->ev1));
    CU(cudaMallocHost((void **)&s->h_in, kCtrlInBytes));
    CU(cudaMallocHost((void **)&s->h_out, kCtrlOutBytes));
    CU(cudaMallocHost((void **)&s->h_err, 4));
    *s->h_err = 0;
    if (int e = salloc(s.get(), (void **)&s->ctrl, sizeof(Ctrl))) return e;
    if (int e = salloc(s.get(), (void **)&s->kc, kv_elems(s.get()) * 2)) return e;
    if (int e = salloc(s.get(), (void **)&s->vc, kv_elems(s.get()) * 2)) return e;
    if (int e = salloc(s.get(), (void **)&s->x, (size_t)c.dim * 4)) return e;
    if (int e = salloc(s.get(), (void **)&s->qkv, (size_t)m->adim * 3 * 4)) return e;
    if (int e = salloc(s.get(), (void **)&s->ctx, (size_t)m->adim * 4)) return e;
    if (int e = salloc(s.get(), (void **)&s->gate, (size_t)(m->tp_world > 1 ? m->hidden_local : m->hidden) * 4)) return e;
    if (m->tp_world > 1) {
        Nccl &n = nccl();
        if (!n.ok) return fail(MSX_ERR_STATE, n.why);
        if (int e = salloc(s.get(), (void **)&s->tp_partial, (size_t)c.dim * 8)) return e;
        Nccl::UniqueId id;
        memcpy(id.internal, nccl_id, 128);
        const int rc = n.CommInitRank(&s->nccl_comm, m->tp_world, id, m->tp_rank);
        if (rc != 0) { s->nccl_comm = nullptr; return fail(MSX_ERR_CUDA, std::string("ncclCommInitRank: ") + n.GetErrorString(rc)); }
    }
    if (int e = salloc(s.get(), (void **)&s->tout, (size_t)c.dim * 4)) return e;
    if (int e = salloc(s.get(), (void **)&s->text_logits, (size_t)c.text_card * 4)) return e;
    if (int e = salloc(s.get(), (void **)&s->rope_cs, (size_t)(c.dim / c.num_heads) * 4)) return e;
    if (int e = salloc(s.get(), (void **)&s->embed_in, (size_t)c.dim * 4)) return e;
    if (c.dep_q > 0) {
        if (int e = salloc(s.get(), (void **)&s->dkc, dkv_elems(s.get()) * 2)) return e;
        if (int e = salloc(s.get(), (void **)&s->dvc, dkv_elems(s.get()) * 2)) return e;
        if (int e = salloc(s.get(), (void **)&s->dx, (size_t)c.dep_dim * 4)) return e;
        if (int e = salloc(s.get(), (void **)&s->dqkv, (size_t)c.dep_dim * 3 * 4)) return e;
        if (int e = salloc(s.get(), (void **)&s->dctx, (size_t)c.dep_dim * 4)) return e;
        if (int e = salloc(s.get(), (void **)&s->dgate, (size_t)m->dep_hidden * 4)) return e;
        if (int e = salloc(s.get(), (void **)&s->audio_logits, (size_t)c.dep_q * c.card * 4)) return e;
    }
    if (c.extra_heads > 0)
        if (int e = salloc(s.get(), (void **)&s->vad_logits, 64 * 4)) return e;
    if (c.cross_attention) {
        if (int e = salloc(s.get(), (void **)&s->cnx, (size_t)c.dim * 4)) return e;
        if (int e = salloc(s.get(), (void **)&s->cq, (size_t)c.dim * 4)) return e;
        if (int e = salloc(s.get(), (void **)&s->cctx, (size_t)c.dim * 4)) return e;
    }
    if (c.demux_second_stream) {
        if (int e = salloc(s.get(), (void **)&s->demux_l, (size_t)c.dim * 4)) return e;
        if (int e = salloc(s.get(), (void **)&s->demux_r, (size_t)c.dim * 4)) return e;
        if (int e = salloc(s.get(), (void **)&s->demux_y1, (size_t)c.dim * 4)) return e;
        if (int e = salloc(s.get(), (void **)&s->demux_y2, (size_t)c.dim * 4)) return e;
    }
    if (m->dep_small) {
        if (int e = salloc(s.get(), (void **)&s->dep_e, (size_t)c.dep_dim * 4)) return e;
    }
    // depformer_in[k] . t_out of every step (hoisted launch): needed with and without the small-embedding path
    if (int e = salloc(s.get(), (void **)&s->dep_d, (size_t)c.dep_q * c.dep_dim * 4)) return e;
    s->noise_floats = kSampleMaxK * (1 + MSX_MAX_STEPS);
    if (int e = salloc(s.get(), (void **)&s->d_noise, (size_t)s->noise_floats * 4)) return e;
    if (int e = salloc(s.get(), (void **)&s->d_probs, (size_t)std::max(c.text_card, c.card) * 4)) return e;
    CU(cudaMallocHost((void **)&s->h_noise, (size_t)s->noise_floats * 4));
    // ctrl: n_in, no overrides
    Ctrl hc;
    memset(&hc, 0, sizeof(hc));
    hc.n_in = c.n_q + 1;
    hc.text_override = INT32_MIN;
    for (int i = 0; i < 40; i++) hc.force[i] = INT32_MIN;
    CU(cudaMemcpy(s->ctrl, &hc, sizeof(hc), cudaMemcpyHostToDevice));

    if (int e = build_graphs(s.get())) return e;
    CU(cudaStreamSynchronize(s->st));
    *out = s.release();
    return 0;
}

static int build_graphs(msx_stream *sp) {
    struct Holder { msx_stream *p; msx_stream *get() const { return p; } msx_stream *operator->() const { return p; } } s{sp};
    msx_model *m = sp->m;
    const msx_config &c = m->cfg;
    if (sp->g_temporal) { cudaGraphExecDestroy(sp->g_temporal); sp->g_temporal = nullptr; }
    if (sp->g_depformer) { cudaGraphExecDestroy(sp->g_depformer); sp->g_depformer = nullptr; }
    // MSX_STREAM_STEP_KERNEL: each stack of the frame is ONE persistent kernel (step_kernel.cuh); models it does not take, and
    // every stream without the flag, run as PDL-chained launches (measured faster on B200: profiles/r2_step_kernel.md)
    sp->step_kernel = false;
    if (step_kernel_eligible(sp)) {
        int per_sm = 0;
        const void *fn = m->stream_type == T_Q4_K ? (const void *)sk::step_kernel<12> : (const void *)sk::step_kernel<8>;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, sk::kThreads, sk::kSmemBytes));
        if (per_sm >= 1) {
            if (!sp->d_prog_t) if (int e = build_step_programs(sp)) return e;
            if (int e = capture(s.get(), [&](Launcher &L) { enqueue_step_kernel(L, s.get(), true); }, &s->g_temporal, &s->launches_temporal)) return e;
            if (c.dep_q > 0)
                if (int e = capture(s.get(), [&](Launcher &L) { enqueue_step_kernel(L, s.get(), false); }, &s->g_depformer, &s->launches_depformer)) return e;
            sp->step_kernel = true;
            return 0;
        }
    }
    if (int e = capture(s.get(), [&](Launcher &L) { enqueue_temporal(L, s.get()); }, &s->g_temporal, &s->launches_temporal)) return e;
    if (c.dep_q > 0) {
        if (int e = capture(s.get(), [&](Launcher &L) { enqueue_depformer(L, s.get()); }, &s->g_depformer, &s->launches_depformer)) return e;
    }
    return 0;
}

extern "C" int msx_stream_set_sampling(msx_stream *s, float temp_text, float temp_audio, int top_k_text, int top_k_audio) {
    if (!s) return fail(MSX_ERR_ARG, "null stream");
    if (top_k_text < 1 || top_k_audio < 1) return fail(MSX_ERR_ARG, "top_k must be >= 1");
    if (std::min(top_k_text, s->m->cfg.text_card) > kSampleMaxK || std::min(top_k_audio, s->m->cfg.card) > kSampleMaxK)
        return fail(MSX_ERR_ARG, "top_k > 256 is not supported");
    CU(cudaSetDevice(s->m->device));
    CU(cudaStreamSynchronize(s->st));
    s->temp_text = temp_text; s->temp_audio = temp_audio; s->top_k_text = top_k_text; s->top_k_audio = top_k_audio;
    if (int e = build_graphs(s)) return e;
    CU(cudaStreamSynchronize(s->st));
    return 0;
}

// noise_text[top_k_text], noise_audio[dep_q][top_k_audio]: Exp(1) draws in candidate order (descending probability)
extern "C" int msx_stream_set_noise(msx_stream *s, const float *noise_text, const float *noise_audio) {
    if (!s) return fail(MSX_ERR_ARG, "null stream");
    const msx_config &c = s->m->cfg;
    CU(cudaSetDevice(s->m->device));
    CU(cudaStreamSynchronize(s->st));     // the pinned staging buffer may still be in flight
    const int kt = std::min(std::min(s->top_k_text, c.text_card), kSampleMaxK), ka = std::min(std::min(s->top_k_audio, c.card), kSampleMaxK);
    for (int i = 0; i < s->noise_floats; i++) s->h_noise[i] = 1.f;
    if (noise_text) memcpy(s->h_noise, noise_text, (size_t)kt * 4);
    if (noise_audio) for (int k = 0; k < c.dep_q; k++) memcpy(s->h_noise + kSampleMaxK + (size_t)k * kSampleMaxK, noise_audio + (size_t)k * ka, (size_t)ka * 4);
    CU(cudaMemcpyAsync(s->d_noise, s->h_noise, (size_t)(kSampleMaxK * (1 + c.dep_q)) * 4, cudaMemcpyHostToDevice, s->st));
    s->noise_fresh = true;
    return 0;
}

extern "C" void msx_stream_free(msx_stream *s) { delete s; }

extern "C" int msx_stream_reset(msx_stream *s) {
    if (!s) return fail(MSX_ERR_ARG, "null stream");
    CU(cudaSetDevice(s->m->device));
    CU(cudaStreamSynchronize(s->st));
    CU(cudaMemsetAsync(s->kc, 0, kv_elems(s) * 2, s->st));
    CU(cudaMemsetAsync(s->vc, 0, kv_elems(s) * 2, s->st));
    if (s->dkc) { CU(cudaMemsetAsync(s->dkc, 0, dkv_elems(s) * 2, s->st)); CU(cudaMemsetAsync(s->dvc, 0, dkv_elems(s) * 2, s->st)); }
    CU(cudaMemsetAsync(s->tout, 0, (size_t)s->m->cfg.dim * 4, s->st));
    CU(cudaMemsetAsync(&s->ctrl->offset, 0, 4, s->st));
    CU(cudaStreamSynchronize(s->st));
    s->host_offset = 0;
    return 0;
}

extern "C" int msx_stream_offset(const msx_stream *s) { return s ? s->host_offset : -1; }
extern "C" int64_t msx_stream_kv_bytes_next(const msx_stream *s) {
    if (!s) return 0;
    const int n_valid = std::min(s->host_offset + 1, s->cap);
    return (int64_t)n_valid * 2 * s->m->adim * 2 * s->m->cfg.num_layers;      // this rank's share under tensor parallelism
}
extern "C" int msx_stream_launches_per_frame(const msx_stream *s) { return s ? s->launches_temporal + s->launches_depformer : 0; }

namespace {

// Token ids index embedding tables on the device: reject anything outside the tables here, on the host (a stray id would
// read device memory out of bounds and the resulting fault would take every stream of the process with it).
//   input tokens (lm.h:555-584): -1 = zero embedding, other negatives = row 0 ("ungenerated", lm_utils.h:172-182), else a row of
//   the table (text: text_card + 1 rows, or the two-stream demux range; audio: card + 1 rows)
//   text_override / force (depformer feed-forward, lm.h:494-516): INT32_MIN = none, else a row of the step's table
int check_tokens(const msx_model *m, const int32_t *tokens, int n_rows, int32_t text_override, const int32_t *force) {
    const msx_config &c = m->cfg;
    const long long n_text = (long long)c.text_card + 1;
    const long long text_max = c.demux_second_stream ? n_text * (n_text + 1) - 1 : n_text - 1;
    if (tokens)
        for (int r = 0; r < n_rows; r++) {
            const int32_t *t = tokens + (size_t)r * (c.n_q + 1);
            if (t[0] < -2 || t[0] > text_max) return fail(MSX_ERR_ARG, "text token " + std::to_string(t[0]) + " is outside the embedding table");
            for (int i = 1; i <= c.n_q; i++)
                if (t[i] < -2 || t[i] > c.card) {
                    std::string rowtxt;
                    for (int j = 0; j <= c.n_q; j++) rowtxt += " " + std::to_string(t[j]);
                    return fail(MSX_ERR_ARG, "audio token " + std::to_string(t[i]) + " (codebook " + std::to_string(i - 1) + ") is outside the embedding table; row:" + rowtxt);
                }
        }
    if (text_override != INT32_MIN && (text_override < -2 || text_override > text_max))
        return fail(MSX_ERR_ARG, "text token " + std::to_string(text_override) + " is outside the depformer text embedding table");
    if (force)
        for (int k = 0; k < c.dep_q; k++)
            if (force[k] != INT32_MIN && (force[k] < 0 || force[k] > c.card))
                return fail(MSX_ERR_ARG, "forced audio token " + std::to_string(force[k]) + " (step " + std::to_string(k) + ") is outside the embedding table");
    return 0;
}

int push_inputs(msx_stream *s, const int32_t *tokens, int32_t text_override, const int32_t *force) {
    const msx_config &c = s->m->cfg;
    if (int e = check_tokens(s->m, tokens, 1, text_override, force)) return e;
    int32_t *h = s->h_in;
    h[0] = text_override;
    if (tokens) for (int i = 0; i < c.n_q + 1; i++) h[1 + i] = tokens[i];
    for (int i = 0; i < 40; i++) h[41 + i] = (force && i < c.dep_q) ? force[i] : INT32_MIN;
    h[81] = s->embed_override_next ? 1 : 0;
    s->embed_override_next = false;
    CU(cudaMemcpyAsync(reinterpret_cast<uint8_t *>(s->ctrl) + kCtrlInOffset, h, kCtrlInBytes, cudaMemcpyHostToDevice, s->st));
    return 0;
}

int pull_outputs(msx_stream *s) {
    CU(cudaMemcpyAsync(s->h_out, reinterpret_cast<uint8_t *>(s->ctrl) + kCtrlOutOffset, kCtrlOutBytes, cudaMemcpyDeviceToHost, s->st));
    CU(cudaMemcpyAsync(s->h_err, &s->ctrl->error, 4, cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    if (*s->h_err) return fail(MSX_ERR_CUDA, "persistent kernel: grid barrier watchdog fired (device-side timeout)");
    return 0;
}

}  // namespace

extern "C" int msx_step_temporal(msx_stream *s, const int32_t *tokens, int32_t *text_token, float *text_logits, float *transformer_out) {
    if (!s || !tokens) return fail(MSX_ERR_ARG, "null argument");
    const msx_config &c = s->m->cfg;
    CU(cudaSetDevice(s->m->device));
    if (int e = push_inputs(s, tokens, INT32_MIN, nullptr)) return e;
    CU(cudaGraphLaunch(s->g_temporal, s->st));
    s->host_offset++;
    if (int e = pull_outputs(s)) return e;
    if (text_token) *text_token = s->h_out[0];
    if (text_logits) CU(cudaMemcpy(text_logits, s->text_logits, (size_t)c.text_card * 4, cudaMemcpyDeviceToHost));
    if (transformer_out) CU(cudaMemcpy(transformer_out, s->tout, (size_t)c.dim * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// PersonaPlex voice-embedding prompt (lm.h:694-709, 1005-1036): the temporal step on a given f32 embedding row
extern "C" int msx_step_temporal_embedding(msx_stream *s, const float *x, int32_t *text_token, float *text_logits, float *transformer_out) {
    if (!s || !x) return fail(MSX_ERR_ARG, "null argument");
    const msx_config &c = s->m->cfg;
    CU(cudaSetDevice(s->m->device));
    CU(cudaMemcpyAsync(s->embed_in, x, (size_t)c.dim * 4, cudaMemcpyHostToDevice, s->st));
    s->embed_override_next = true;
    if (int e = push_inputs(s, nullptr, INT32_MIN, nullptr)) return e;
    CU(cudaGraphLaunch(s->g_temporal, s->st));
    s->host_offset++;
    if (int e = pull_outputs(s)) return e;
    if (text_token) *text_token = s->h_out[0];
    if (text_logits) CU(cudaMemcpy(text_logits, s->text_logits, (size_t)c.text_card * 4, cudaMemcpyDeviceToHost));
    if (transformer_out) CU(cudaMemcpy(transformer_out, s->tout, (size_t)c.dim * 4, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int msx_step_depformer(msx_stream *s, int32_t text_token, const int32_t *force, int32_t *audio_tokens, float *audio_logits) {
    if (!s) return fail(MSX_ERR_ARG, "null stream");
    const msx_config &c = s->m->cfg;
    if (c.dep_q <= 0) return fail(MSX_ERR_STATE, "model has no depformer");
    CU(cudaSetDevice(s->m->device));
    // tokens[] of the input block are left as they are in h_in (already consumed by the temporal step)
    if (int e = push_inputs(s, nullptr, text_token, force)) return e;
    CU(cudaGraphLaunch(s->g_depformer, s->st));
    if (int e = pull_outputs(s)) return e;
    if (audio_tokens) for (int k = 0; k < c.dep_q; k++) audio_tokens[k] = s->h_out[1 + k];
    if (audio_logits) CU(cudaMemcpy(audio_logits, s->audio_logits, (size_t)c.dep_q * c.card * 4, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int msx_step(msx_stream *s, const int32_t *tokens, int32_t *out_tokens) {
    if (!s || !tokens) return fail(MSX_ERR_ARG, "null argument");
    const msx_config &c = s->m->cfg;
    CU(cudaSetDevice(s->m->device));
    if (int e = push_inputs(s, tokens, INT32_MIN, nullptr)) return e;
    CU(cudaGraphLaunch(s->g_temporal, s->st));
    s->host_offset++;
    if (c.dep_q > 0) CU(cudaGraphLaunch(s->g_depformer, s->st));
    if (int e = pull_outputs(s)) return e;
    if (out_tokens) for (int k = 0; k < 1 + c.dep_q; k++) out_tokens[k] = s->h_out[k];
    return 0;
}

// ---- peer-memory all-reduce for tensor-parallel streams ---------------------------------------------------------------
static size_t tp_arena_bytes(const msx_stream *s) { return (size_t)2 * s->m->tp_world * s->m->cfg.dim * 16 + 64; }

extern "C" int msx_stream_tp_export(msx_stream *s, uint8_t *handle64) {
    if (!s || !handle64) return fail(MSX_ERR_ARG, "null argument");
    if (s->m->tp_world < 2) return fail(MSX_ERR_STATE, "not a tensor-parallel stream");
    if (s->m->tp_world > 8) return fail(MSX_ERR_ARG, "peer-memory all-reduce supports up to 8 ranks");
    CU(cudaSetDevice(s->m->device));
    if (!s->tp_arena) {
        CU(cudaMalloc((void **)&s->tp_arena, tp_arena_bytes(s)));        // plain cudaMalloc: IPC-exportable
        s->allocs.push_back(s->tp_arena);
        CU(cudaMemset(s->tp_arena, 0, tp_arena_bytes(s)));
    }
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, s->tp_arena));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t");
    memcpy(handle64, &h, 64);
    return 0;
}

// handles[world][64] in rank order (own entry ignored).  Re-captures the graphs with the fused GEMV -> peer push path.
extern "C" int msx_stream_tp_connect(msx_stream *s, const uint8_t *handles) {
    if (!s || !handles) return fail(MSX_ERR_ARG, "null argument");
    msx_model *m = s->m;
    if (m->tp_world < 2 || !s->tp_arena) return fail(MSX_ERR_STATE, "call msx_stream_tp_export on every rank first");
    CU(cudaSetDevice(m->device));
    CU(cudaStreamSynchronize(s->st));
    TpCtx h;
    h.rank = m->tp_rank; h.world = m->tp_world; h.dim = m->cfg.dim;
    if (m->cfg.dim > 1024 * kTpApplyPer) return fail(MSX_ERR_ARG, "peer-memory all-reduce: dim too large for the apply kernel");
    const size_t inbox_bytes = (size_t)2 * m->tp_world * m->cfg.dim * 16;
    for (int r = 0; r < m->tp_world; r++) {
        uint8_t *base = s->tp_arena;
        if (r != m->tp_rank) {
            cudaIpcMemHandle_t ih;
            memcpy(&ih, handles + (size_t)r * 64, 64);
            void *p = nullptr;
            CU(cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
            s->tp_peer_maps.push_back(p);
            base = (uint8_t *)p;
        }
        h.inbox[r] = reinterpret_cast<uint4 *>(base);
    }
    h.frame_ctr = reinterpret_cast<uint32_t *>(s->tp_arena + inbox_bytes);
    h.reduces_per_frame = 2 * m->cfg.num_layers;
    s->tp_frame_ctr = h.frame_ctr;
    h.error = &s->ctrl->error;
    if (!s->d_tp) if (int e = salloc(s, (void **)&s->d_tp, sizeof(TpCtx))) return e;
    CU(cudaMemcpy(s->d_tp, &h, sizeof(h), cudaMemcpyHostToDevice));
    s->tp_p2p = true;
    if (int e = build_graphs(s)) return e;
    CU(cudaStreamSynchronize(s->st));
    return 0;
}

extern "C" int msx_stream_set_condition(msx_stream *s, const float *cond_sum, const float *cond_cross, int tc) {
    if (!s) return fail(MSX_ERR_ARG, "null stream");
    msx_model *m = s->m; const msx_config &c = m->cfg;
    if (cond_cross && !c.cross_attention) return fail(MSX_ERR_STATE, "model has no cross-attention layers (moshi_lm_set_voice_condition returns -1 likewise, moshi.cpp:729-731)");
    if (cond_cross && tc <= 0) return fail(MSX_ERR_ARG, "tc must be positive");
    CU(cudaSetDevice(m->device));
    CU(cudaStreamSynchronize(s->st));
    const int dim = c.dim;
    if (cond_sum) {
        if (!s->cond_sum) if (int e = salloc(s, (void **)&s->cond_sum, (size_t)dim * 4)) return e;
        CU(cudaMemcpy(s->cond_sum, cond_sum, (size_t)dim * 4, cudaMemcpyHostToDevice));
    } else s->cond_sum = nullptr;       // (allocation stays in the stream's arena)
    if (cond_cross) {
        // init(): k | v = in_proj rows [dim, 3 dim) applied to every condition column, kept in f32 (transformer.h:343-396)
        float *d_cross = nullptr;
        CU(cudaMalloc((void **)&d_cross, (size_t)tc * dim * 4));
        CU(cudaMemcpy(d_cross, cond_cross, (size_t)tc * dim * 4, cudaMemcpyHostToDevice));
        if (tc != s->tc || !s->kv_cross) if (int e = salloc(s, (void **)&s->kv_cross, (size_t)c.num_layers * tc * 2 * dim * 4)) { cudaFree(d_cross); return e; }
        s->tc = tc;
        Launcher L{s->st, m->num_sms};
        L.pdl = false;
        for (int l = 0; l < c.num_layers; l++) {
            const QLinear kvw = linear_rows(m->layers[l].cross_in, dim, 2 * dim);
            for (int i = 0; i < tc; i++) {
                GemvArgs g;
                g.ctrl = s->ctrl; g.w = kvw; g.x = d_cross + (size_t)i * dim; g.out = s->kv_cross + ((size_t)l * tc + i) * 2 * dim;
                L.gemv(g, PRO_PLAIN, EPI_STORE);
            }
        }
        cudaError_t e = cudaStreamSynchronize(s->st);
        cudaFree(d_cross);
        if (L.err != cudaSuccess || e != cudaSuccess) return fail(MSX_ERR_CUDA, std::string("cross-attention memory: ") + cudaGetErrorString(L.err != cudaSuccess ? L.err : e));
    } else { s->tc = 0; }
    if (int e = build_graphs(s)) return e;     // the graphs bake the conditioning pointers and tc
    CU(cudaStreamSynchronize(s->st));
    return 0;
}

// voice_condition() (src/moshi.cpp:296-366): condition_sum = cfg_proj . cfg_embed[2] + control_proj . control_embed[0]
// ("cfg 2.0", "control ok"); condition_cross [5T][dim] = the projected speaker embedding in the first T rows, the learnt
// padding in the other 4T, plus the sinusoidal position embedding; then the cross-attention K / V memory as in
// msx_stream_set_condition.  speaker_wavs is the voice file's tensor as stored: [channels][frames], frames fastest.
extern "C" int msx_stream_set_voice(msx_stream *s, const float *speaker_wavs, int channels, int frames, float *sum_out, float *cross_out) {
    if (!s || !speaker_wavs) return fail(MSX_ERR_ARG, "null argument");
    msx_model *m = s->m;
    if (!m->cfg.cross_attention) return fail(MSX_ERR_STATE, "model has no cross-attention layers (moshi_lm_load_voice_condition returns -1, moshi.cpp:740-742)");
    if (!m->has_conditioners) return fail(MSX_ERR_STATE, "the GGUF carries no lm.condition_provider.conditioners.* tensors (moshi_lm_load_voice_condition returns -2, moshi.cpp:744-745)");
    if (frames <= 0 || channels != m->spk_proj.ne0) return fail(MSX_ERR_ARG, "speaker_wavs must be [" + std::to_string(m->spk_proj.ne0) + "][frames]");
    CU(cudaSetDevice(m->device));
    CU(cudaStreamSynchronize(s->st));
    const int dim = m->cfg.dim, tc = 5 * frames;
    float *buf = nullptr;      // wavs [C][T] | speaker [T][dim] | cfg [dim] | control [dim] | sum [dim] | cross [5T][dim]
    const size_t n_w = (size_t)channels * frames, n_s = (size_t)frames * dim, n_c = (size_t)tc * dim;
    CU(cudaMalloc((void **)&buf, (n_w + n_s + 3 * (size_t)dim + n_c) * 4));
    float *d_w = buf, *d_s = d_w + n_w, *d_cfg = d_s + n_s, *d_ctl = d_cfg + dim, *d_sum = d_ctl + dim, *d_cross = d_sum + dim;
    std::vector<float> sum(dim), cross(n_c);
    cudaError_t e = cudaMemcpy(d_w, speaker_wavs, n_w * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        const dim3 rows((dim + 7) / 8, 1);
        CondLinearArgs a;
        a.w = m->cfg_proj; a.table = m->cfg_embed; a.row = 2; a.y = d_cfg;
        cond_linear_kernel<<<rows, 256, 0, s->st>>>(a);
        a.w = m->control_proj; a.table = m->control_embed; a.row = 0; a.y = d_ctl;
        cond_linear_kernel<<<rows, 256, 0, s->st>>>(a);
        cond_add_kernel<<<(dim + 255) / 256, 256, 0, s->st>>>(d_cfg, d_ctl, d_sum, dim);
        CondLinearArgs b;                                              // column t of the transposed wavs: x[k] = wavs[k][t]
        b.w = m->spk_proj; b.row = -1; b.x = d_w; b.xstride = frames; b.xcol = 1; b.y = d_s;
        cond_linear_kernel<<<dim3((dim + 7) / 8, frames), 256, 0, s->st>>>(b);
        cond_cross_kernel<<<tc, 256, 0, s->st>>>(d_s, m->spk_pad, m->cond_freq, d_cross, frames, dim);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(sum.data(), d_sum, (size_t)dim * 4, cudaMemcpyDeviceToHost, s->st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(cross.data(), d_cross, n_c * 4, cudaMemcpyDeviceToHost, s->st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->st);
    }
    cudaFree(buf);
    if (e != cudaSuccess) return fail(MSX_ERR_CUDA, std::string("voice conditioners: ") + cudaGetErrorString(e));
    if (sum_out) memcpy(sum_out, sum.data(), (size_t)dim * 4);
    if (cross_out) memcpy(cross_out, cross.data(), n_c * 4);
    return msx_stream_set_condition(s, sum.data(), cross.data(), tc);
}

// moshi_lm_set_voice_condition + moshi_lm_load_voice_condition (moshi.cpp:729-760): the voice file is a safetensors whose
// "speaker_wavs" tensor ([1,] channels, frames; f32 / f16 / bf16) feeds the conditioners
extern "C" int msx_model_has_conditioners(const msx_model *m) { return m && m->has_conditioners ? 1 : 0; }

extern "C" int msx_stream_load_voice(msx_stream *s, const char *path) {
    if (!s || !path) return fail(MSX_ERR_ARG, "null argument");
    SafeTensorsFile f;
    std::string err;
    if (!f.open(path, err)) {
        const bool io = err.rfind("cannot", 0) == 0;
        return fail(io ? MSX_ERR_IO : MSX_ERR_FORMAT, err);
    }
    for (const SafeTensor &t : f.tensors()) {
        if (t.name != "speaker_wavs") continue;
        std::vector<int64_t> shape = t.shape;
        while (shape.size() > 2 && shape.front() == 1) shape.erase(shape.begin());
        const int esz = t.dtype == "F32" ? 4 : (t.dtype == "F16" || t.dtype == "BF16") ? 2 : 0;
        if (shape.size() != 2 || !esz || (uint64_t)(shape[0] * shape[1] * esz) != t.nbytes)
            return fail(MSX_ERR_FORMAT, "speaker_wavs must be a [channels, frames] float tensor");
        std::vector<float> w((size_t)(shape[0] * shape[1]));
        for (size_t i = 0; i < w.size(); i++) {
            if (esz == 4) memcpy(&w[i], t.data + i * 4, 4);
            else {
                uint16_t h; memcpy(&h, t.data + i * 2, 2);
                if (t.dtype == "BF16") { const uint32_t u = (uint32_t)h << 16; memcpy(&w[i], &u, 4); }
                else { __half v; memcpy(&v, &h, 2); w[i] = __half2float(v); }
            }
        }
        return msx_stream_set_voice(s, w.data(), (int)shape[0], (int)shape[1], nullptr, nullptr);
    }
    return fail(MSX_ERR_FORMAT, std::string(path) + " has no speaker_wavs tensor");
}

extern "C" int msx_vad(msx_stream *s, float *vad) {
    if (!s || !vad) return fail(MSX_ERR_ARG, "null argument");
    const msx_model *m = s->m;
    if (m->cfg.extra_heads <= 2) { *vad = 0.f; return 0; }      // lm.h:973-975
    CU(cudaSetDevice(m->device));
    const QLinear &w = m->extra_heads[2];
    if (w.rows > 64) return fail(MSX_ERR_ARG, "extra head wider than 64");
    Launcher L{s->st, m->num_sms};
    GemvArgs g;
    g.ctrl = s->ctrl; g.w = w; g.x = s->tout; g.out = s->vad_logits;
    L.gemv(g, PRO_PLAIN, EPI_STORE);
    if (L.err != cudaSuccess) return fail(MSX_ERR_CUDA, cudaGetErrorString(L.err));
    float h[64];
    CU(cudaMemcpyAsync(h, s->vad_logits, w.rows * 4, cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    // ggml_soft_max over the head's outputs, element 0 (lm.h:968-971)
    float mx = h[0];
    for (int i = 1; i < w.rows; i++) mx = std::max(mx, h[i]);
    double sum = 0;
    for (int i = 0; i < w.rows; i++) { h[i] = expf(h[i] - mx); sum += h[i]; }
    *vad = h[0] * (float)(1.0 / sum);
    return 0;
}

extern "C" int msx_run_resident(msx_stream *s, const int32_t *frames, int n_frames, int n_steps, int32_t *out_tokens, float *elapsed_ms) {
    if (!s || !frames || n_frames <= 0 || n_steps <= 0) return fail(MSX_ERR_ARG, "bad argument");
    const msx_config &c = s->m->cfg;
    CU(cudaSetDevice(s->m->device));
    const int n_in = c.n_q + 1, n_out = 1 + c.dep_q;
    if (int e = check_tokens(s->m, frames, n_frames, INT32_MIN, nullptr)) return e;
    int32_t *d_feed = nullptr, *d_trace = nullptr;
    CU(cudaMalloc((void **)&d_feed, (size_t)n_frames * n_in * 4));
    if (out_tokens) CU(cudaMalloc((void **)&d_trace, (size_t)n_steps * n_out * 4));
    CU(cudaMemcpy(d_feed, frames, (size_t)n_frames * n_in * 4, cudaMemcpyHostToDevice));
    if (int e = push_inputs(s, nullptr, INT32_MIN, nullptr)) return e;
    Ctrl hdr;                       // first 32 bytes: offset, frame, feed_n, n_in, feed, trace
    memset(&hdr, 0, sizeof(hdr));
    hdr.offset = s->host_offset; hdr.frame = 0; hdr.feed_n = n_frames; hdr.n_in = n_in; hdr.feed = d_feed; hdr.trace = d_trace;
    CU(cudaMemcpyAsync(s->ctrl, &hdr, kCtrlInOffset, cudaMemcpyHostToDevice, s->st));
    CU(cudaStreamSynchronize(s->st));
    CU(cudaEventRecord(s->ev0, s->st));
    for (int i = 0; i < n_steps; i++) {
        CU(cudaGraphLaunch(s->g_temporal, s->st));
        if (c.dep_q > 0) CU(cudaGraphLaunch(s->g_depformer, s->st));
    }
    CU(cudaEventRecord(s->ev1, s->st));
    CU(cudaStreamSynchronize(s->st));
    s->host_offset += n_steps;
    if (elapsed_ms) CU(cudaEventElapsedTime(elapsed_ms, s->ev0, s->ev1));
    if (out_tokens) CU(cudaMemcpy(out_tokens, d_trace, (size_t)n_steps * n_out * 4, cudaMemcpyDeviceToHost));
    int32_t zero2[2] = {0, 0};
    CU(cudaMemcpy(&s->ctrl->frame, zero2, 8, cudaMemcpyHostToDevice));   // frame = 0, feed_n = 0 -> host mode
    cudaFree(d_feed);
    if (d_trace) cudaFree(d_trace);
    return 0;
}

// Eager (non-graph) run of one fused frame with a CUDA event after every launch: per-family kernel time.
extern "C" int msx_profile_frame(msx_stream *s, const int32_t *tokens, int32_t *out_tokens,
                                 float *family_ms, int32_t *family_launches, int max_families) {
    if (!s || !tokens || !family_ms || !family_launches) return fail(MSX_ERR_ARG, "null argument");
    const msx_config &c = s->m->cfg;
    CU(cudaSetDevice(s->m->device));
    for (int i = 0; i < max_families; i++) { family_ms[i] = 0.f; family_launches[i] = 0; }
    if (int e = push_inputs(s, tokens, INT32_MIN, nullptr)) return e;
    std::vector<cudaEvent_t> ev;
    std::vector<int> fam;
    Launcher L{s->st, s->m->num_sms};
    L.model = s->m;
    L.events = &ev; L.families = &fam;
    if (s->step_kernel) { enqueue_step_kernel(L, s, true); s->host_offset++; if (c.dep_q > 0) enqueue_step_kernel(L, s, false); }
    else {
        enqueue_temporal(L, s);
        s->host_offset++;
        if (c.dep_q > 0) enqueue_depformer(L, s);
    }
    if (L.err != cudaSuccess) return fail(MSX_ERR_CUDA, std::string("launch failed: ") + cudaGetErrorString(L.err));
    if (int e = pull_outputs(s)) return e;
    if (out_tokens) for (int k = 0; k < 1 + c.dep_q; k++) out_tokens[k] = s->h_out[k];
    for (size_t i = 0; i + 1 < ev.size(); i++) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
        const int f = fam[i];
        if (f < max_families) { family_ms[f] += ms; family_launches[f] += 1; }
    }
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    return 0;
}
// One frame through the step kernels with the in-kernel timeline enabled.  rows: [n_phases][n_cta][9] int64 =
// {family, start, end, after prologue, after main loop, input loaded, rms scale known, epilogue stored, -} (globaltimer ns; 0 where
// a phase has no such stage); temporal
// phases first.  Returns the number of phases in *n_phases and the CTA count in *n_cta; rows must hold max_rows * 9 values.
extern "C" int msx_step_timeline(msx_stream *s, const int32_t *tokens, long long *rows, int max_rows, int *n_phases, int *n_cta) {
    if (!s || !tokens || !rows || !n_phases || !n_cta) return fail(MSX_ERR_ARG, "null argument");
    if (!s->step_kernel) return fail(MSX_ERR_STATE, "stream does not run the persistent step kernel");
    CU(cudaSetDevice(s->m->device));
    const int nt = s->n_prog_t, nd = s->n_prog_d, nc = s->m->num_sms;
    if ((size_t)(nt + nd) * nc > (size_t)max_rows) return fail(MSX_ERR_ARG, "timeline buffer too small");
    long long *d = nullptr;
    const size_t words = (size_t)(nt + nd) * nc * 8;
    CU(cudaMalloc((void **)&d, words * 8));
    CU(cudaMemset(d, 0, words * 8));
    if (int e = push_inputs(s, tokens, INT32_MIN, nullptr)) { cudaFree(d); return e; }
    Launcher L{s->st, s->m->num_sms};
    enqueue_step_kernel(L, s, true, d);
    s->host_offset++;
    if (nd) enqueue_step_kernel(L, s, false, d + (size_t)nt * nc * 8);
    if (L.err != cudaSuccess) { cudaFree(d); return fail(MSX_ERR_CUDA, std::string("launch failed: ") + cudaGetErrorString(L.err)); }
    if (int e = pull_outputs(s)) { cudaFree(d); return e; }
    std::vector<long long> h(words);
    CU(cudaMemcpy(h.data(), d, words * 8, cudaMemcpyDeviceToHost));
    cudaFree(d);
    for (int i = 0; i < nt + nd; i++)
        for (int c = 0; c < nc; c++) {
            long long *r = rows + ((size_t)i * nc + c) * 9;
            r[0] = i < nt ? s->prog_fam_t[i] : s->prog_fam_d[i - nt];
            for (int j = 0; j < 8; j++) r[1 + j] = h[((size_t)i * nc + c) * 8 + j];
        }
    *n_phases = nt + nd; *n_cta = nc;
    return 0;
}

extern "C" int msx_family_count(void) { return FAM_COUNT; }
extern "C" const char *msx_family_name(int i) { return (i >= 0 && i < FAM_COUNT) ? kFamilyNames[i] : ""; }

// CUDA-event stopwatch on the stream the step kernels and copies run on (bench.py e2e timing)
extern "C" int msx_timer_start(msx_stream *s) {
    if (!s) return fail(MSX_ERR_ARG, "null stream");
    CU(cudaSetDevice(s->m->device));
    CU(cudaStreamSynchronize(s->st));
    CU(cudaEventRecord(s->ev0, s->st));
    return 0;
}
extern "C" int msx_timer_stop(msx_stream *s, float *elapsed_ms) {
    if (!s || !elapsed_ms) return fail(MSX_ERR_ARG, "null argument");
    CU(cudaSetDevice(s->m->device));
    CU(cudaEventRecord(s->ev1, s->st));
    CU(cudaEventSynchronize(s->ev1));
    CU(cudaEventElapsedTime(elapsed_ms, s->ev0, s->ev1));
    return 0;
}

// Asynchronous variant of msx_run_resident for several streams on one GPU: enqueue n_steps fused frames on the
// stream's own CUDA stream and return; msx_stream_wait() joins and yields the device time.  feed buffers are owned
// by the stream until the wait.
extern "C" int msx_run_resident_async(msx_stream *s, const int32_t *frames, int n_frames, int n_steps) {
    if (!s || !frames || n_frames <= 0 || n_steps <= 0) return fail(MSX_ERR_ARG, "bad argument");
    const msx_config &c = s->m->cfg;
    CU(cudaSetDevice(s->m->device));
    const int n_in = c.n_q + 1;
    if (int e = check_tokens(s->m, frames, n_frames, INT32_MIN, nullptr)) return e;
    if (s->d_feed) { cudaFree(s->d_feed); s->d_feed = nullptr; }
    CU(cudaMalloc((void **)&s->d_feed, (size_t)n_frames * n_in * 4));
    CU(cudaMemcpy(s->d_feed, frames, (size_t)n_frames * n_in * 4, cudaMemcpyHostToDevice));
    if (int e = push_inputs(s, nullptr, INT32_MIN, nullptr)) return e;
    Ctrl hdr;
    memset(&hdr, 0, sizeof(hdr));
    hdr.offset = s->host_offset; hdr.frame = 0; hdr.feed_n = n_frames; hdr.n_in = n_in; hdr.feed = s->d_feed; hdr.trace = nullptr;
    CU(cudaMemcpyAsync(s->ctrl, &hdr, kCtrlInOffset, cudaMemcpyHostToDevice, s->st));
    CU(cudaStreamSynchronize(s->st));
    CU(cudaEventRecord(s->ev0, s->st));
    for (int i = 0; i < n_steps; i++) {
        CU(cudaGraphLaunch(s->g_temporal, s->st));
        if (c.dep_q > 0) CU(cudaGraphLaunch(s->g_depformer, s->st));
    }
    CU(cudaEventRecord(s->ev1, s->st));
    s->host_offset += n_steps;
    return 0;
}
extern "C" int msx_stream_wait(msx_stream *s, float *elapsed_ms) {
    if (!s) return fail(MSX_ERR_ARG, "null stream");
    CU(cudaSetDevice(s->m->device));
    CU(cudaStreamSynchronize(s->st));
    if (elapsed_ms) CU(cudaEventElapsedTime(elapsed_ms, s->ev0, s->ev1));
    if (s->d_feed) {
        int32_t zero2[2] = {0, 0};
        CU(cudaMemcpy(&s->ctrl->frame, zero2, 8, cudaMemcpyHostToDevice));   // back to host mode
        cudaFree(s->d_feed); s->d_feed = nullptr;
    }
    return 0;
}

extern "C" int msx_stream_get_kv(msx_stream *s, int layer, int head, int slot, uint16_t *k, uint16_t *v) {
    if (!s || !k || !v) return fail(MSX_ERR_ARG, "null argument");
    const msx_config &c = s->m->cfg;
    if (layer < 0 || layer >= c.num_layers || head < 0 || head >= c.num_heads || slot < 0 || slot >= s->cap) return fail(MSX_ERR_ARG, "index out of range");
    CU(cudaSetDevice(s->m->device));
    const int dh = c.dim / c.num_heads;
    const size_t o = (((size_t)layer * c.num_heads + head) * s->cap + slot) * dh;
    CU(cudaMemcpy(k, s->kc + o, dh * 2, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(v, s->vc + o, dh * 2, cudaMemcpyDeviceToHost));
    return 0;
}

// -------------------------------------------------------------------------------------------------
// LMGen host logic (reference lm.h:715-743 state, lm.h:778-979 step; greedy, no state machine)
// -------------------------------------------------------------------------------------------------
// Per-generator libc random state.  The reference draws its Exp(1) noise with rand() (context.h:464-480), i.e. glibc's
// process-global TYPE_3 additive-feedback generator; random_r on a private 128-byte state produces the same sequence for the
// same seed (rand() before any srand() behaves like seed 1), so a single generator reproduces the reference's draws while
// several generators / threads no longer perturb each other.
struct GenRng {
    struct random_data rd;
    char state[128];
    GenRng() { seed(1u); }
    void seed(unsigned v) { memset(&rd, 0, sizeof(rd)); memset(state, 0, sizeof(state)); initstate_r(v, state, sizeof(state), &rd); }
    int next() { int32_t r = 0; random_r(&rd, &r); return (int)r; }
    float exp1() { return -logf(next() / (float)RAND_MAX); }
};

struct msx_gen {
    GenRng rng;
    msx_stream *s = nullptr;
    msx_config cfg{};
    msx_step_fn fn = nullptr;         // host-logic tests: the model step is a caller-supplied callback
    void *user = nullptr;
    int offset = 0, CT = 0, ncb = 0, max_delay = 0, delay_steps = 0;
    std::vector<int32_t> cache;       // [CT][ncb], init -2 = lm_ungenerated_token_id
    std::vector<int32_t> initial;     // {text_card, card, card, ...}
    // TTS hooks (lm.h:877-899 on_text_hook, 915-931 on_audio_hook) and the frames to swallow after an audio prefix
    msx_text_hook text_hook = nullptr; void *text_user = nullptr;
    msx_audio_hook audio_hook = nullptr; void *audio_user = nullptr;
    int skip = 0;
};

static void gen_init_impl(msx_gen *g, int delay_steps);
static void gen_init(msx_gen *g, int delay_steps) { gen_init_impl(g, delay_steps); }

extern "C" int msx_gen_create(msx_stream *s, int delay_steps, msx_gen **out) {
    if (!s || !out) return fail(MSX_ERR_ARG, "null argument");
    const msx_config &c = s->m->cfg;
    auto *g = new msx_gen;
    g->s = s; g->cfg = c;
    gen_init(g, delay_steps);
    *out = g;
    return 0;
}

extern "C" int msx_gen_create_with_callback(const msx_config *cfg, int delay_steps, msx_step_fn fn, void *user, msx_gen **out) {
    if (!cfg || !fn || !out) return fail(MSX_ERR_ARG, "null argument");
    if (cfg->n_q < 0 || cfg->n_q + 1 > MSX_MAX_CODEBOOKS || cfg->dep_q < 0 || cfg->dep_q > MSX_MAX_STEPS || cfg->n_delays < cfg->n_q + 1)
        return fail(MSX_ERR_ARG, "bad codebook counts / delays");
    auto *g = new msx_gen;
    g->cfg = *cfg; g->fn = fn; g->user = user;
    gen_init(g, delay_steps);
    *out = g;
    return 0;
}

static void gen_init_impl(msx_gen *g, int delay_steps) {
    const msx_config &c = g->cfg;
    g->ncb = c.n_q + 1; g->delay_steps = delay_steps;
    int md = c.delays[0];
    for (int i = 0; i < c.n_delays; i++) md = std::max(md, c.delays[i]);     // lm_default.h:177-183
    g->max_delay = md;
    g->CT = md + 2 + (c.personaplex ? 1 : 0);                                // lm.h:727-729
    g->cache.assign((size_t)g->CT * g->ncb, -2);
    g->initial.assign(g->ncb, c.card);
    g->initial[0] = c.text_card;
}
extern "C" void msx_gen_free(msx_gen *g) { delete g; }
extern "C" void msx_gen_seed(msx_gen *g, unsigned seed) { if (g) g->rng.seed(seed); }
extern "C" int msx_gen_offset(const msx_gen *g) { return g ? g->offset : -1; }
extern "C" int msx_gen_max_delay(const msx_gen *g) { return g ? g->max_delay : -1; }

// moshi_lmgen_step_voice_prompt, embedding variant (lm.h:1005-1050): one prompt frame = temporal step on the row,
// text token forced to 3, depformer step (its tokens are dropped), offset++
extern "C" int msx_gen_prompt_embedding(msx_gen *g, const float *row) {
    if (!g || !row || !g->s) return fail(MSX_ERR_ARG, "null argument / callback generator");
    int32_t text = 0, audio[MSX_MAX_STEPS];
    msx_stream *s = g->s; const msx_config &c = g->cfg;
    if (s->temp_text > 0.f || s->temp_audio > 0.f) {
        // the sampled graphs read this frame's Exp(1) noise: draw it exactly like msx_gen_step (and like the reference's
        // moshi_lmgen_step_voice_prompt, whose two graphs consume kt + dep_q * ka draws per prompt frame)
        const int kt = std::min(std::min(s->top_k_text, c.text_card), kSampleMaxK), ka = std::min(std::min(s->top_k_audio, c.card), kSampleMaxK);
        std::vector<float> nt(kt), na((size_t)std::max(1, c.dep_q) * ka);
        if (s->temp_text > 0.f) for (int i = 0; i < kt; i++) nt[i] = g->rng.exp1();
        if (s->temp_audio > 0.f) for (size_t i = 0; i < (size_t)c.dep_q * ka; i++) na[i] = g->rng.exp1();
        if (int e = msx_stream_set_noise(s, nt.data(), na.data())) return e;
    }
    if (int e = msx_step_temporal_embedding(g->s, row, &text, nullptr, nullptr)) return e;
    if (g->cfg.dep_q > 0) if (int e = msx_step_depformer(g->s, 3, nullptr, audio, nullptr)) return e;
    g->offset++;
    return 0;
}
// the token delay ring as stored with the voice ("voice.cache"): cache[CT][n_q+1] row-major here
extern "C" int msx_gen_set_cache(msx_gen *g, const int32_t *cache) {
    if (!g || !cache) return fail(MSX_ERR_ARG, "null argument");
    for (size_t i = 0; i < g->cache.size(); i++) g->cache[i] = cache[i];
    return 0;
}
extern "C" int msx_gen_cache_rows(const msx_gen *g) { return g ? g->CT : -1; }

// T prompt frames with ALL n_q+1 tokens given, as one batched-T prefill: the host side of moshi_lmgen_step's "provided"
// branch (ring writes lm.h:812-818, input gather 826-833, no output write-back 933-943, offset++) for every frame, then
// msx_stream_prefill on the gathered inputs.  The libc rand() draws the per-frame sampler would have consumed are
// consumed here too, so a sampled conversation continues with the reference's random sequence.
extern "C" int msx_stream_prefill(msx_stream *s, const int32_t *tokens, int T);
extern "C" int msx_gen_prefill(msx_gen *g, const int32_t *rows, int T) {
    if (!g || !rows || T <= 0 || !g->s) return fail(MSX_ERR_ARG, "bad argument / callback generator");
    const msx_config &c = g->cfg;
    const int CT = g->CT, ncb = g->ncb;
    std::vector<int32_t> inputs((size_t)T * ncb), cache = g->cache;
    int offset = g->offset;
    for (int f = 0; f < T; f++) {
        for (int i = 0; i < ncb; i++) cache[(size_t)((offset + c.delays[i]) % CT) * ncb + i] = rows[(size_t)f * ncb + i];
        const int pos = offset % CT;
        for (int i = 0; i < ncb; i++) inputs[(size_t)f * ncb + i] = (offset <= c.delays[i]) ? g->initial[i] : cache[(size_t)pos * ncb + i];
        offset++;
    }
    if (int e = msx_stream_prefill(g->s, inputs.data(), T)) return e;
    g->cache.swap(cache); g->offset = offset;
    if (g->s->temp_text > 0.f || g->s->temp_audio > 0.f) {
        const int kt = std::min(std::min(g->s->top_k_text, c.text_card), kSampleMaxK), ka = std::min(std::min(g->s->top_k_audio, c.card), kSampleMaxK);
        const long draws = (long)T * ((g->s->temp_text > 0.f ? kt : 0) + (g->s->temp_audio > 0.f ? (long)c.dep_q * ka : 0));
        for (long i = 0; i < draws; i++) (void)g->rng.next();
    }
    return 0;
}

extern "C" int msx_gen_set_text_hook(msx_gen *g, msx_text_hook fn, void *user) {
    if (!g) return fail(MSX_ERR_ARG, "null generator");
    g->text_hook = fn; g->text_user = user;
    return 0;
}
extern "C" int msx_gen_set_audio_hook(msx_gen *g, msx_audio_hook fn, void *user) {
    if (!g) return fail(MSX_ERR_ARG, "null generator");
    g->audio_hook = fn; g->audio_user = user;
    return 0;
}

// moshi_lmgen_step (lm.h:778-979) around the model call: gen_prepare = ring write of the incoming tokens + input gather,
// gen_finish = audio hooks, ring write of the generated tokens, delayed emit.  Shared by msx_gen_step (one stream) and
// msx_bgen_step (a batch of streams, batch.inl).
struct GenPrep { bool provided = false; int32_t input[MSX_MAX_CODEBOOKS]; };
static int gen_prepare(msx_gen *g, const int32_t *in_tokens, int n_in, GenPrep *p) {
    const msx_config &c = g->cfg;
    const int CT = g->CT, ncb = g->ncb;
    int dep_q = c.dep_q;
    if (c.personaplex) dep_q = 8;                                             // lm.h:802-805
    const int dep_q_1 = dep_q + 1;
    const int needed = ncb - dep_q - 1;
    p->provided = false;
    if (needed > 0) {
        if (!in_tokens || n_in < needed) return fail(MSX_ERR_ARG, "not enough input tokens");   // reference: assert (lm.h:810)
        if (n_in == ncb) {
            for (int i = 0; i < ncb; i++) g->cache[(size_t)((g->offset + c.delays[i]) % CT) * ncb + i] = in_tokens[i];
            p->provided = true;
        } else {
            for (int i = 0; i < needed; i++)
                g->cache[(size_t)((g->offset + c.delays[dep_q_1 + i]) % CT) * ncb + dep_q_1 + i] = in_tokens[i];
        }
    }
    const int pos = g->offset % CT;
    for (int i = 0; i < ncb; i++) p->input[i] = (g->offset <= c.delays[i]) ? g->initial[i] : g->cache[(size_t)pos * ncb + i];
    return 0;
}
static int gen_finish(msx_gen *g, const GenPrep &p, int32_t *out, int depformer_replace_tokens, int32_t *out_text, int32_t *out_audio) {
    const msx_config &c = g->cfg;
    const int CT = g->CT, ncb = g->ncb;
    int dep_q = c.dep_q;
    if (c.personaplex) dep_q = 8;
    const int dep_q_1 = dep_q + 1;
    const int text_token = out[0];
    int32_t *audio = out + 1;
    if (c.dep_q > 0 && g->delay_steps)
        for (int q = 0; q < c.dep_q; q++)
            if (g->offset < c.delays[q + 1] + g->delay_steps) audio[q] = -1;  // lm.h:915-921
    if (c.dep_q > 0 && g->audio_hook) {                                       // audio prefix (lm.h:922-931)
        const int sk = g->audio_hook(g->audio_user, g->offset, audio, c.dep_q);
        if (sk >= 0) g->skip = sk;
    }
    g->offset++;
    if (!p.provided) {
        const int pp = g->offset % CT;
        g->cache[(size_t)pp * ncb + 0] = text_token;
        for (int q = 0; q < c.dep_q; q++) g->cache[(size_t)pp * ncb + q + 1] = audio[q];
    }
    for (int q = 0; q < c.dep_q; q++) out_audio[q] = audio[q];
    if (g->skip > 0) { --g->skip; return 0; }                                 // lm.h:944-947
    if (g->offset <= g->max_delay || depformer_replace_tokens) return 0;
    *out_text = g->cache[(size_t)((g->offset - g->max_delay + c.delays[0]) % CT) * ncb + 0];
    for (int i = 1; i < dep_q_1; i++)
        out_audio[i - 1] = g->cache[(size_t)((g->offset - g->max_delay + c.delays[i]) % CT) * ncb + i];
    for (int q = 0; q < c.dep_q; q++)
        if (out_audio[q] == -1) return 0;
    return 1;
}

extern "C" int msx_gen_step(msx_gen *g, const int32_t *in_tokens, int n_in, int depformer_replace_tokens, int32_t *out_text, int32_t *out_audio) {
    if (!g || !out_text || !out_audio) return fail(MSX_ERR_ARG, "null argument");
    msx_stream *s = g->s;
    const msx_config &c = g->cfg;
    GenPrep p;
    if (int e = gen_prepare(g, in_tokens, n_in, &p)) return e;
    const int32_t *input = p.input;

    int32_t out[1 + MSX_MAX_STEPS];
    for (int i = 0; i < 1 + MSX_MAX_STEPS; i++) out[i] = -1;
    if (!g->fn && (s->temp_text > 0.f || s->temp_audio > 0.f)) {
        // Exp(1) draws exactly like GraphContext::_exponential_compute (context.h:464-480): one libc rand() per
        // top-k candidate, text graph first, then the depformer codebooks in order
        const int kt = std::min(std::min(s->top_k_text, c.text_card), kSampleMaxK), ka = std::min(std::min(s->top_k_audio, c.card), kSampleMaxK);
        std::vector<float> nt(kt), na((size_t)std::max(1, c.dep_q) * ka);
        if (s->temp_text > 0.f) for (int i = 0; i < kt; i++) nt[i] = g->rng.exp1();
        if (s->temp_audio > 0.f && !depformer_replace_tokens) for (size_t i = 0; i < (size_t)c.dep_q * ka; i++) na[i] = g->rng.exp1();
        if (int e = msx_stream_set_noise(s, nt.data(), na.data())) return e;
    }
    if (g->fn) {
        if (int e = g->fn(g->user, input, depformer_replace_tokens, out)) return fail(MSX_ERR_STATE, "step callback failed: " + std::to_string(e));
        if (g->text_hook) out[0] = g->text_hook(g->text_user, g->offset, out[0]);               // (callback mode: applied after the fact)
        if (depformer_replace_tokens) for (int q = 0; q < c.dep_q; q++) out[1 + q] = -1;      // lm.h:909-913
    } else if (g->text_hook) {
        // the text token is rewritten between the two graphs (state machine / text prefix, lm.h:877-899)
        if (int e = msx_step_temporal(s, input, &out[0], nullptr, nullptr)) return e;
        out[0] = g->text_hook(g->text_user, g->offset, out[0]);
        if (c.dep_q > 0 && !depformer_replace_tokens)
            if (int e = msx_step_depformer(s, out[0], nullptr, out + 1, nullptr)) return e;
    } else if (c.dep_q > 0 && !depformer_replace_tokens) {
        if (int e = msx_step(s, input, out)) return e;                        // temporal + depformer, one sync
    } else {
        if (int e = msx_step_temporal(s, input, &out[0], nullptr, nullptr)) return e;
    }
    return gen_finish(g, p, out, depformer_replace_tokens, out_text, out_audio);
}

// -------------------------------------------------------------------------------------------------
// unit-level test entry points
// -------------------------------------------------------------------------------------------------
namespace {
int device_setup(int device, std::unique_ptr<msx_model> &m) {
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(MSX_ERR_CUDA, "no such CUDA device");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(MSX_ERR_CUDA, "sm_100a device required");
    m.reset(new msx_model);
    m->device = device; m->num_sms = prop.multiProcessorCount;
    return set_smem_attrs();
}
}  // namespace

extern "C" int msx_test_gemv(int device, int type, const void *w, int64_t k, int64_t rows, const float *x, const float *alpha, int prologue, float *y) {
    if (!w || !x || !y) return fail(MSX_ERR_ARG, "null argument");
    std::unique_ptr<msx_model> m;
    if (int e = device_setup(device, m)) return e;
    QLinear ql;
    if (int e = upload_linear(m.get(), w, type, k, rows, 0, &ql)) return e;
    float *dx = nullptr, *dy = nullptr, *da = nullptr;
    if (int e = dev_alloc(m.get(), (void **)&dx, (size_t)k * 4)) return e;
    if (int e = dev_alloc(m.get(), (void **)&dy, (size_t)rows * 4)) return e;
    CU(cudaMemcpy(dx, x, (size_t)k * 4, cudaMemcpyHostToDevice));
    if (prologue == PRO_RMS) {
        if (!alpha) return fail(MSX_ERR_ARG, "alpha required for the rms prologue");
        if (int e = dev_alloc(m.get(), (void **)&da, (size_t)k * 4)) return e;
        CU(cudaMemcpy(da, alpha, (size_t)k * 4, cudaMemcpyHostToDevice));
    }
    Launcher L{nullptr, m->num_sms};
    GemvArgs g;
    g.w = ql; g.x = dx; g.alpha = da; g.eps = 1e-8f; g.out = dy;
    L.gemv(g, prologue == PRO_RMS ? PRO_RMS : PRO_PLAIN, EPI_STORE);
    if (L.err != cudaSuccess) return fail(MSX_ERR_CUDA, std::string("gemv launch: ") + cudaGetErrorString(L.err));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(y, dy, (size_t)rows * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// Micro-benchmark of the fused GEMV kernel: n_mats copies of one random [rows][k] matrix (rotated so every
// launch streams cold weights from HBM when n_mats * bytes > L2), iters launches timed with CUDA events.
extern "C" int msx_bench_gemv(int device, int type, const void *w, int64_t k, int64_t rows, int n_mats, int iters,
                              int prologue, int epilogue, float *avg_us) {
    if (!w || !avg_us || n_mats < 1 || iters < 1) return fail(MSX_ERR_ARG, "bad argument");
    std::unique_ptr<msx_model> m;
    if (int e = device_setup(device, m)) return e;
    std::vector<QLinear> mats(n_mats);
    for (int i = 0; i < n_mats; i++)
        if (int e = upload_linear(m.get(), w, type, k, rows, epilogue == EPI_GATE ? (int)(rows / 2) : 0, &mats[i])) return e;
    float *dx = nullptr, *dy = nullptr, *da = nullptr;
    if (int e = dev_alloc(m.get(), (void **)&dx, (size_t)k * 4)) return e;
    if (int e = dev_alloc(m.get(), (void **)&dy, (size_t)std::max<int64_t>(rows, k) * 4)) return e;
    if (int e = dev_alloc(m.get(), (void **)&da, (size_t)k * 4)) return e;
    std::vector<float> hx(k), ha(k, 1.0f);
    for (int64_t i = 0; i < k; i++) hx[i] = (float)((i * 2654435761u) % 2001) / 1000.f - 1.f;
    CU(cudaMemcpy(dx, hx.data(), (size_t)k * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(da, ha.data(), (size_t)k * 4, cudaMemcpyHostToDevice));
    CU(cudaMemset(dy, 0, (size_t)std::max<int64_t>(rows, k) * 4));
    cudaStream_t st;
    CU(cudaStreamCreate(&st));
    Launcher L{st, m->num_sms};
    L.model = m.get();
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    unsigned long long *dkey = nullptr;
    if (int e = dev_alloc(m.get(), (void **)&dkey, 8)) return e;
    CU(cudaMemset(dkey, 0, 8));
    // like the real step: the launches are captured into a CUDA graph and replayed
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < iters; i++) {
        GemvArgs g;
        g.w = mats[i % n_mats]; g.x = dx; g.alpha = da; g.eps = 1e-8f; g.out = dy; g.key = dkey;
        L.gemv(g, prologue, epilogue);
    }
    CU(cudaStreamEndCapture(st, &graph));
    if (L.err != cudaSuccess) return fail(MSX_ERR_CUDA, std::string("gemv launch: ") + cudaGetErrorString(L.err));
    CU(cudaGraphInstantiate(&exec, graph, 0));
    CU(cudaGraphLaunch(exec, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaEventRecord(e0, st));
    CU(cudaGraphLaunch(exec, st));
    CU(cudaEventRecord(e1, st));
    CU(cudaStreamSynchronize(st));
    cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    *avg_us = ms * 1000.f / iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(st);
    return 0;
}

extern "C" int msx_test_dequant_rows(int device, int type, const void *table, int64_t k, int64_t table_rows, const int32_t *row_ids, int n_rows, float *out) {
    if (!table || !row_ids || !out) return fail(MSX_ERR_ARG, "null argument");
    std::unique_ptr<msx_model> m;
    if (int e = device_setup(device, m)) return e;
    EmbTable t;
    if (int e = upload_table(m.get(), table, type, k, table_rows, &t)) return e;
    int32_t *ids = nullptr; float *o = nullptr;
    if (int e = dev_alloc(m.get(), (void **)&ids, (size_t)n_rows * 4)) return e;
    if (int e = dev_alloc(m.get(), (void **)&o, (size_t)n_rows * k * 4)) return e;
    CU(cudaMemcpy(ids, row_ids, (size_t)n_rows * 4, cudaMemcpyHostToDevice));
    const long long n = (long long)n_rows * k;
    dequant_rows_kernel<<<(unsigned)((n + 255) / 256), 256>>>(t, ids, n_rows, o);
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, o, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int msx_test_dequant_repacked(int device, int type, const void *w, int64_t k, int64_t rows, float *out) {
    if (!w || !out) return fail(MSX_ERR_ARG, "null argument");
    std::unique_ptr<msx_model> m;
    if (int e = device_setup(device, m)) return e;
    QLinear ql;
    if (int e = upload_linear(m.get(), w, type, k, rows, 0, &ql)) return e;
    float *o = nullptr;
    const long long n = (long long)rows * k;
    if (int e = dev_alloc(m.get(), (void **)&o, (size_t)n * 4)) return e;
    dequant_repacked_kernel<<<(unsigned)((n + 255) / 256), 256>>>(ql, o);
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, o, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// GGUF blocks of the on-load quantisers (dst_type 8 = Q8_0, 2 = Q4_0, 12 = Q4_K) for `rows` rows of k f32 / f16 / bf16 values
extern "C" int msx_test_quantize_rows(int device, int src_type, int dst_type, const void *x, int64_t k, int64_t rows, void *out) {
    if (!x || !out) return fail(MSX_ERR_ARG, "null argument");
    if (!is_float_type(src_type)) return fail(MSX_ERR_ARG, "source must be f32 / f16 / bf16");
    std::unique_ptr<msx_model> m;
    if (int e = device_setup(device, m)) return e;
    const size_t raw = (size_t)ggml_row_size(src_type, k) * rows;
    if (int e = ensure_staging(m.get(), raw)) return e;
    CU(cudaMemcpy(m->staging, x, raw, cudaMemcpyHostToDevice));
    const uint8_t *blocks = nullptr;
    if (int e = quantize_staging(m.get(), src_type, dst_type, k, rows, &blocks)) return e;
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(out, blocks, (size_t)ggml_row_size(dst_type, k) * rows, cudaMemcpyDeviceToHost));
    return 0;
}

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

#include "batch.inl"

static void free_prefill(struct msx_batch *b) { delete b; }
