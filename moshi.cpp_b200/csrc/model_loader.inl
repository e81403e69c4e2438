// model_loader.inl — msx_model: GGUF tensors -> device arenas (repack / on-load quantisation / stream layout), tensor-parallel
// shards, msx_model_* entry points.  Reference: WeightLoader src/loader.h:85-271, tensor names lm.h:370-395.  Included by engine.cu.

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
    // tc layout of a Q4_K linear for the tcgen05 prefill GEMM (tc_gemm.cuh), keyed by its qs plane
    std::unordered_map<const void *, const uint8_t *> wtc;
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
    // third copy for the tensor-core prompt prefill: [128-row tile][super-block][row][GGUF block]
    if (m->tp_world == 1 && type == T_Q4_K && K % 256 == 0 && rows % tc::kM == 0) {
        void *wt = nullptr;
        if (int e = dev_alloc(m, &wt, (size_t)rows * (K / 256) * 144)) return e;
        const long long n = (long long)rows * (K / 256) * 9;
        tc::tc_layout_kernel<<<(unsigned)((n + 255) / 256), 256>>>(src_blocks, (uint8_t *)wt, (int)rows, (int)(K / 256), perm_half);
        CU(cudaGetLastError());
        m->wtc[w.qs] = (const uint8_t *)wt;
    }
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
