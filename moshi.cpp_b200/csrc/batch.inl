// batch.inl — lock-step batch of independent conversation streams on one GPU (included by engine.cu).
//
// SURVEY.md §8e / BASELINE.json config 5: streams are independent (private KV rings, delay state, positions);
// stepping n of them together lets every weight matrix be read from HBM once per frame for all of them
// (mma_gemm.cuh).  The step is the single-stream step (enqueue_temporal / enqueue_depformer above) with
//   * one activation-quantisation launch + one tensor-core GEMM launch per linear layer,
//   * attention / embedding / bookkeeping kernels launched with one grid slice per stream.
// Streams may sit at different positions (msx_batch_reset_stream): everything per-frame lives in the
// stream's own Ctrl block.

struct msx_batch {
    msx_model *m = nullptr;
    int n = 0, cap = 0, attn_split = 1;
    cudaStream_t st = nullptr;
    Ctrl *ctrl = nullptr;                 // device [n]
    int32_t *h_in = nullptr, *h_out = nullptr;   // pinned [n][84], [n][44]
    uint16_t *kc = nullptr, *vc = nullptr, *dkc = nullptr, *dvc = nullptr;
    float *x = nullptr, *qkv = nullptr, *ctx = nullptr, *gate = nullptr, *tout = nullptr, *text_logits = nullptr, *rope_cs = nullptr;
    float *dx = nullptr, *dqkv = nullptr, *dctx = nullptr, *dgate = nullptr, *audio_logits = nullptr;
    uint8_t *img = nullptr;               // activation image of the GEMM being fed
    uint8_t *img_tout = nullptr;          // image of transformer_out (depformer_in input, reused by every codebook step)
    cudaGraphExec_t g_temporal = nullptr, g_depformer = nullptr;
    int launches_temporal = 0, launches_depformer = 0;
    std::vector<int> host_offset;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<void *> allocs;
    // batched-T prefill of ONE stream: the columns are consecutive positions of `prefill_of`, they
    // share its KV rings and its CUDA stream; n_active <= n columns are live in the launches being enqueued
    msx_stream *prefill_of = nullptr;
    int n_active = 0;
    bool tc = false;                      // prefill with 64 columns per pass on the tcgen05 GEMM (tc_gemm.cuh) instead of 8 on mma.sync
    uint16_t *victim_k = nullptr, *victim_v = nullptr;             // prefill: old ring rows of the slots a pass overwrites [H][64][Dh]
    double *tc_partial = nullptr; int tc_max_tiles = 0;            // stream-K partial sums [cta][2][64][128]
    unsigned int *tc_tickets = nullptr;                            // [tiles] arrival counters (zero between launches)
    // sampling (sampling.h:46-64) per stream: temperature <= 0 = greedy; Exp(1) noise supplied by the host per frame
    float temp_text = 0.f, temp_audio = 0.f;
    int top_k_text = 25, top_k_audio = 250;
    float *d_noise = nullptr, *h_noise = nullptr, *d_probs = nullptr;      // noise [n][1 + MSX_MAX_STEPS][kSampleMaxK]
    uint8_t *h_hdr = nullptr;             // pinned [n][32]: Ctrl headers (position per column)

    ~msx_batch() {
        if (m) cudaSetDevice(m->device);
        if (g_temporal) cudaGraphExecDestroy(g_temporal);
        if (g_depformer) cudaGraphExecDestroy(g_depformer);
        for (void *p : allocs) cudaFree(p);
        if (h_in) cudaFreeHost(h_in);
        if (h_out) cudaFreeHost(h_out);
        if (h_hdr) cudaFreeHost(h_hdr);
        if (h_noise) cudaFreeHost(h_noise);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (st && !prefill_of) cudaStreamDestroy(st);
    }
};

namespace {

int balloc(msx_batch *b, void **p, size_t bytes) {
    CU(cudaMalloc(p, std::max<size_t>(bytes, 16)));
    CU(cudaMemset(*p, 0, std::max<size_t>(bytes, 16)));
    b->allocs.push_back(*p);
    return 0;
}

// units layout of one linear (created on first use, kept for the life of the model)
int tiles_of(msx_model *m, const QLinear &w, QTiles *out) {
    auto it = m->tiles.find(w.qs);
    if (it != m->tiles.end()) { *out = it->second; return 0; }
    if (w.type != T_Q4_K && w.type != T_Q8_0) return fail(MSX_ERR_ARG, "batched streams need q4_k or q8_0 linear weights");
    if (w.type == T_Q8_0 && (w.K % 128)) return fail(MSX_ERR_ARG, "batched q8_0 streams need K % 128 == 0");
    QTiles t;
    t.type = w.type; t.K = w.K; t.rows = w.rows; t.nsb = w.K / unit_weights_of(w.type); t.n_tiles = (w.rows + 15) / 16;
    void *u = nullptr;
    if (int e = dev_alloc(m, &u, (size_t)t.n_tiles * t.nsb * unit_bytes_of(w.type))) return e;
    if (w.type == T_Q4_K) {
        const long long n = (long long)t.n_tiles * t.nsb * 148;
        tile_q4k_kernel<<<(unsigned)((n + 255) / 256), 256>>>(w, w.gate, (uint8_t *)u, t.n_tiles);
    } else {
        const long long n = (long long)t.n_tiles * t.nsb * 136;
        tile_q8_0_kernel<<<(unsigned)((n + 255) / 256), 256>>>(w, w.gate, (uint8_t *)u, t.n_tiles);
    }
    CU(cudaGetLastError());
    t.units = (const uint8_t *)u;
    m->tiles[w.qs] = t;
    *out = t;
    return 0;
}

int ensure_all_tiles(msx_model *m) {
    QTiles t;
    auto all = [&](const std::vector<QLinear> &v) -> int { for (const QLinear &w : v) if (int e = tiles_of(m, w, &t)) return e; return 0; };
    for (const LayerW &l : m->layers) { if (int e = all(l.in_proj)) return e; if (int e = all(l.out_proj)) return e; if (int e = all(l.lin_in)) return e; if (int e = all(l.lin_out)) return e; }
    for (const LayerW &l : m->dep_layers) { if (int e = all(l.in_proj)) return e; if (int e = all(l.out_proj)) return e; if (int e = all(l.lin_in)) return e; if (int e = all(l.lin_out)) return e; }
    if (int e = tiles_of(m, m->text_linear, &t)) return e;
    if (int e = all(m->dep_in)) return e;
    if (int e = all(m->linears)) return e;
    CU(cudaDeviceSynchronize());
    return 0;
}

struct BatchLauncher {
    Launcher &L;
    msx_batch *b;
    int err = 0;

    void quant(const float *x, int ld, const float *alpha, float *norm_out, int norm_ld, int K, int family, uint8_t *img = nullptr) {
        const int wt = b->m->text_linear.type;          // a model is q4_k or q8_0 throughout
        QuantArgs q;
        q.x = x; q.ld = ld; q.alpha = alpha; q.eps = 1e-8f; q.norm_out = norm_out; q.norm_ld = norm_ld; q.img = img ? img : b->img; q.K = K;
        q.plain = b->tc ? 1 : 0;
        L.fam = family; L.begin();
        if (wt == T_Q4_K) L.launch_pdl(quant_q8k_kernel, dim3(b->n_active, quant_parts_for(K)), dim3(kGemmThreads), 0, q);
        else L.launch_pdl(quant_q8_0_kernel, dim3(b->n_active, quant_parts_for(K)), dim3(kGemmThreads), 0, q);
        L.check();
    }
    // y = W (norm?)(x): one launch with the fused prologue when the inner dimension is short, else quantise + GEMM
    void linear(const float *x, int ld, const float *alpha, const QLinear &w, float *out, int out_ld, int epi, int family, int key_index = -1) {
        if (b->tc) {                     // more than 8 columns: quantise (plain image) + tcgen05 GEMM
            quant(x, ld, alpha, nullptr, 0, w.K, family);
            gemm(w, out, out_ld, epi, family, key_index);
            return;
        }
        if (fuse_small && w.type == T_Q4_K && gemm_can_fuse_quant(w.K, b->n_active)) {
            gemm(w, out, out_ld, epi, family, key_index, nullptr, 0, nullptr, x, ld, alpha);
        } else {
            quant(x, ld, alpha, nullptr, 0, w.K, family);
            gemm(w, out, out_ld, epi, family, key_index);
        }
    }
    bool fuse_small = true;
    // tcgen05 GEMM over the quantised image (tc_gemm.cuh); arg-max and embedding-add epilogues run as a small kernel behind it
    void tc_gemm(const QLinear &w, float *out, int ld, int epi, int family, int key_index, const EmbTable *emb, int emb_step, const uint8_t *img) {
        auto it = b->m->wtc.find(w.qs);
        if (it == b->m->wtc.end()) { err = fail(MSX_ERR_STATE, "linear without a tensor-core layout in a wide batch"); return; }
        tc::TcMatmulArgs g;
        g.w = it->second; g.K = w.K; g.rows = w.rows; g.img = img ? img : b->img; g.out = out; g.ld = ld; g.nb = b->n_active;
        g.epi = (epi == EPI_ARGMAX || epi == EPI_ADD_EMB) ? (int)EPI_STORE : epi;
        g.partial = b->tc_partial; g.tickets = b->tc_tickets;
        if (w.rows / tc::kM > b->tc_max_tiles) { err = fail(MSX_ERR_STATE, "tc gemm: ticket array too small"); return; }
        const dim3 grid(tc::grid_for(w.rows / tc::kM, w.K >> 8, L.num_sms)), block(tc::kThreads);
        L.fam = family; L.begin();
        const int nc = tc::columns_for(b->n_active);
        if (nc == 16) L.launch_pdl(tc::tc_matmul_q4k_kernel<16>, grid, block, (size_t)tc::kSmemBytes, g);
        else if (nc == 32) L.launch_pdl(tc::tc_matmul_q4k_kernel<32>, grid, block, (size_t)tc::kSmemBytes, g);
        else L.launch_pdl(tc::tc_matmul_q4k_kernel<64>, grid, block, (size_t)tc::kSmemBytes, g);
        L.check();
        if (epi == EPI_ARGMAX) {
            L.begin();
            L.launch_pdl(tc::argmax_rows_kernel, dim3(b->n_active), dim3(256), 0, (const float *)out, ld, w.rows, b->ctrl, key_index);
            L.check();
        } else if (epi == EPI_ADD_EMB) {
            L.begin();
            L.launch_pdl(tc::dep_embed_add_cols_kernel, dim3((w.rows + 255) / 256, b->n_active), dim3(256), 0, (const Ctrl *)b->ctrl, *emb, emb_step, out, ld, w.rows);
            L.check();
        }
    }
    void gemm(const QLinear &w, float *out, int ld, int epi, int family, int key_index = -1, const EmbTable *emb = nullptr, int emb_step = 0,
              const uint8_t *img = nullptr, const float *xsrc = nullptr, int xld = 0, const float *alpha = nullptr) {
        if (b->tc) { tc_gemm(w, out, ld, epi, family, key_index, emb, emb_step, img); return; }
        MatmulArgs g;
        if (int e = tiles_of(b->m, w, &g.w)) { err = e; return; }
        g.xsrc = xsrc; g.xld = xld; g.alpha = alpha; g.eps = 1e-8f;
        g.img = img ? img : b->img; g.out = out; g.ld = ld; g.nb = b->n_active; g.epi = epi; g.ctrl = b->ctrl; g.key_index = key_index; g.emb_step = emb_step;
        if (emb) g.emb = *emb;
        g.stages = gemm_stages_for(w.K, w.type);
        L.fam = family; L.begin();
        const dim3 grid(gemm_grid_for(g.w.n_tiles, L.num_sms)), block(kGemmThreads);
        const size_t smem = (size_t)gemm_smem_bytes(w.K, g.stages, w.type);
        const bool lean = epi == EPI_STORE || epi == EPI_RESID || epi == EPI_GATE;
        if (w.type == T_Q4_K) {
            if (lean && xsrc) L.launch_pdl(dq_matmul_mma_kernel<12, 2>, grid, block, smem, g);
            else if (lean) L.launch_pdl(dq_matmul_mma_kernel<12, 1>, grid, block, smem, g);
            else L.launch_pdl(dq_matmul_mma_kernel<12, 0>, grid, block, smem, g);
        } else {
            if (lean) L.launch_pdl(dq_matmul_mma_kernel<8, 1>, grid, block, smem, g);
            else L.launch_pdl(dq_matmul_mma_kernel<8, 0>, grid, block, smem, g);
        }
        L.check();
    }
};

void enqueue_layer_b(BatchLauncher &B, const LayerW &lw, int w, bool temporal, int layer, int pos_const) {
    msx_batch *b = B.b; const msx_model *m = b->m; const msx_config &c = m->cfg;
    const int dim = temporal ? c.dim : c.dep_dim, heads = temporal ? c.num_heads : c.dep_heads;
    const int cap = temporal ? b->cap : m->dep_cap;
    const int hidden = temporal ? m->hidden : m->dep_hidden;
    float *x = temporal ? b->x : b->dx, *qkv = temporal ? b->qkv : b->dqkv, *ctx = temporal ? b->ctx : b->dctx, *gate = temporal ? b->gate : b->dgate;
    B.linear(x, dim, lw.norm1, lw.in_proj[w], qkv, 3 * dim, EPI_STORE, temporal ? FAM_IN_PROJ : FAM_DEP_IN_PROJ);
    AttnArgs a;
    a.qkv = qkv; a.ctx = ctx; a.ctrl = b->ctrl; a.pos_const = pos_const; a.cap = cap; a.dim = dim;
    a.max_period = temporal ? c.max_period : c.dep_max_period;
    a.rope_freq = temporal ? m->rope_freq : m->dep_rope_freq;
    a.rope_cs = (temporal && c.max_period) ? b->rope_cs : nullptr;
    a.small_ctx = 32;
    const size_t lstride = (size_t)cap * dim;
    const int n_layers = temporal ? c.num_layers : c.dep_layers;
    a.kc = (temporal ? b->kc : b->dkc) + (size_t)layer * lstride;
    a.vc = (temporal ? b->vc : b->dvc) + (size_t)layer * lstride;
    a.kv_bstride = (int64_t)n_layers * lstride; a.qkv_bstride = 3 * dim; a.ctx_bstride = dim;
    if (b->prefill_of) {
        // columns = consecutive positions of one stream: insert all their K / V rows into ITS ring first, then attend
        a.kc = b->prefill_of->kc + (size_t)layer * lstride; a.vc = b->prefill_of->vc + (size_t)layer * lstride;
        a.kv_bstride = 0; a.skip_insert = 1;
        a.victim_k = b->victim_k; a.victim_v = b->victim_v; a.victim_n = b->n_active;
        B.L.fam = FAM_ATTN; B.L.begin();
        if (dim / heads == 128) B.L.launch_pdl(kv_insert_kernel<128>, dim3(heads, b->n_active), dim3(64), 0, a);
        else B.L.launch_pdl(kv_insert_kernel<64>, dim3(heads, b->n_active), dim3(64), 0, a);
        B.L.check();
    }
    B.L.attn(a, heads, dim / heads, temporal ? b->attn_split : 1, temporal ? FAM_ATTN : FAM_DEP_ATTN, b->n_active);
    B.linear(ctx, dim, nullptr, lw.out_proj[w], x, dim, EPI_RESID, temporal ? FAM_OUT_PROJ : FAM_DEP_OUT_PROJ);
    B.linear(x, dim, lw.norm2, lw.lin_in[w], gate, hidden, EPI_GATE, temporal ? FAM_LIN_IN : FAM_DEP_LIN_IN);
    B.linear(gate, hidden, nullptr, lw.lin_out[w], x, dim, EPI_RESID, temporal ? FAM_LIN_OUT : FAM_DEP_LIN_OUT);
}

void enqueue_temporal_b(BatchLauncher &B) {
    msx_batch *b = B.b; const msx_model *m = b->m; const msx_config &c = m->cfg;
    Launcher &L = B.L;
    EmbedArgs e;
    e.tables = m->d_emb; e.n_tables = c.n_q + 1; e.dim = c.dim; e.ctrl = b->ctrl; e.x = b->x;
    if (c.max_period) { e.rope_cs = b->rope_cs; e.rope_freq = m->rope_freq; e.dh = c.dim / c.num_heads; }
    L.fam = FAM_EMBED; L.begin();
    L.launch_pdl(embed_kernel, dim3((c.dim + kThreads - 1) / kThreads, b->n_active), dim3(kThreads), 0, e);
    L.check();
    for (int l = 0; l < c.num_layers; l++) enqueue_layer_b(B, m->layers[l], 0, true, l, -1);
    if (b->prefill_of) return;           // prompt frames only populate the KV rings: no head, no sampling, no depformer
    B.quant(b->x, c.dim, m->out_norm, b->tout, c.dim, c.dim, FAM_TEXT_HEAD);
    B.gemm(m->text_linear, b->text_logits, c.text_card, EPI_ARGMAX, FAM_TEXT_HEAD, -1);
    if (b->temp_text > 0.f) {          // moshi_sample_token per stream: overwrites the arg-max key with the sampled token
        SampleArgs sa;
        sa.logits = b->text_logits; sa.logits_stride = c.text_card; sa.n = c.text_card;
        sa.k = std::min(std::min(b->top_k_text, c.text_card), kSampleMaxK); sa.inv_temp = 1.f / b->temp_text;
        sa.noise = b->d_noise; sa.noise_stride = (1 + MSX_MAX_STEPS) * kSampleMaxK;
        sa.probs = b->d_probs; sa.probs_stride = std::max(c.text_card, c.card); sa.ctrl = b->ctrl; sa.key_index = -1;
        L.fam = FAM_TEXT_HEAD; L.begin();
        L.launch_pdl(sample_kernel, dim3(b->n_active), dim3(kSampleThreads), 0, sa);
        L.check();
    }
    L.fam = FAM_FINALIZE; L.begin();
    L.launch_pdl(finalize_temporal_kernel, dim3(b->n_active), dim3(32), 0, b->ctrl, c.dep_q > 0 ? 1 : 0, (uint32_t *)nullptr);
    L.check();
}

void enqueue_depformer_b(BatchLauncher &B) {
    msx_batch *b = B.b; const msx_model *m = b->m; const msx_config &c = m->cfg;
    Launcher &L = B.L;
    for (int k = 0; k < c.dep_q; k++) {
        const int wsel = c.schedule_len ? c.schedule[k] : k;
        const int w = m->dep_nw == 1 ? 0 : wsel;
        // transformer_out is quantised once per frame into its own image and reused by all dep_q depformer_in GEMMs
        if (k == 0) B.quant(b->tout, c.dim, nullptr, nullptr, 0, c.dim, FAM_DEP_IN, b->img_tout);
        B.gemm(m->dep_in[w], b->dx, c.dep_dim, EPI_ADD_EMB, FAM_DEP_IN, -1, k == 0 ? &m->dep_text_emb : &m->dep_emb[k - 1], k, b->img_tout);
        for (int l = 0; l < c.dep_layers; l++) enqueue_layer_b(B, m->dep_layers[l], w, false, l, k);
        B.linear(b->dx, c.dep_dim, nullptr, m->linears[k], b->audio_logits + (size_t)k * c.card, c.dep_q * c.card, EPI_ARGMAX, FAM_DEP_HEAD, k);
        if (b->temp_audio > 0.f) {
            SampleArgs sa;
            sa.logits = b->audio_logits + (size_t)k * c.card; sa.logits_stride = c.dep_q * c.card; sa.n = c.card;
            sa.k = std::min(std::min(b->top_k_audio, c.card), kSampleMaxK); sa.inv_temp = 1.f / b->temp_audio;
            sa.noise = b->d_noise + (size_t)(1 + k) * kSampleMaxK; sa.noise_stride = (1 + MSX_MAX_STEPS) * kSampleMaxK;
            sa.probs = b->d_probs; sa.probs_stride = std::max(c.text_card, c.card); sa.ctrl = b->ctrl; sa.key_index = k;
            L.fam = FAM_DEP_HEAD; L.begin();
            L.launch_pdl(sample_kernel, dim3(b->n_active), dim3(kSampleThreads), 0, sa);
            L.check();
        }
    }
    L.fam = FAM_DEP_FINALIZE; L.begin();
    L.launch_pdl(finalize_depformer_kernel, dim3(b->n_active), dim3(64), 0, b->ctrl, (int)c.dep_q);
    L.check();
}

template <typename F>
int capture_b(msx_batch *b, F &&body, cudaGraphExec_t *exec, int *launches) {
    Launcher L{b->st, b->m->num_sms};
    BatchLauncher B{L, b};
    CU(cudaStreamBeginCapture(b->st, cudaStreamCaptureModeThreadLocal));
    body(B);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(b->st, &graph);
    if (B.err) { if (graph) cudaGraphDestroy(graph); return B.err; }
    if (L.err != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return fail(MSX_ERR_CUDA, std::string("kernel launch failed during capture: ") + cudaGetErrorString(L.err)); }
    if (e != cudaSuccess) return fail(MSX_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(MSX_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    *launches = L.count;
    return 0;
}

int push_inputs_b(msx_batch *b, const int32_t *tokens) {
    const msx_config &c = b->m->cfg;
    if (int e = check_tokens(b->m, tokens, b->n, INT32_MIN, nullptr)) return e;
    const int w = (int)(kCtrlInBytes / 4);
    for (int s = 0; s < b->n; s++) {
        int32_t *h = b->h_in + (size_t)s * w;
        h[0] = INT32_MIN;
        if (tokens) for (int i = 0; i < c.n_q + 1; i++) h[1 + i] = tokens[(size_t)s * (c.n_q + 1) + i];
        for (int i = 0; i < 40; i++) h[41 + i] = INT32_MIN;
        h[81] = h[82] = h[83] = 0;
    }
    CU(cudaMemcpy2DAsync(reinterpret_cast<uint8_t *>(b->ctrl) + kCtrlInOffset, sizeof(Ctrl), b->h_in, kCtrlInBytes, kCtrlInBytes, b->n,
                         cudaMemcpyHostToDevice, b->st));
    return 0;
}

int pull_outputs_b(msx_batch *b) {
    CU(cudaMemcpy2DAsync(b->h_out, kCtrlOutBytes, reinterpret_cast<uint8_t *>(b->ctrl) + kCtrlOutOffset, sizeof(Ctrl), kCtrlOutBytes, b->n,
                         cudaMemcpyDeviceToHost, b->st));
    CU(cudaStreamSynchronize(b->st));
    return 0;
}

}  // namespace

static int batch_create_impl(msx_model *m, int n_streams, int context_override, msx_stream *prefill_of, msx_batch **out);
// every linear of the temporal stack has the tc layout (Q4_K, rows % 128 == 0, K % 256 == 0, one GPU)
static bool tc_prefill_ok(const msx_model *m) {
    if (m->layers.empty()) return false;
    for (const LayerW &l : m->layers)
        for (const QLinear *w : {&l.in_proj[0], &l.out_proj[0], &l.lin_in[0], &l.lin_out[0]})
            if (!m->wtc.count(w->qs)) return false;
    return true;
}
// every linear of the whole decode step has the tc layout: batches of more than 8 streams run on the tcgen05 GEMM
static bool tc_batch_ok(const msx_model *m) {
    if (!tc_prefill_ok(m) || !m->wtc.count(m->text_linear.qs)) return false;
    for (const QLinear &w : m->dep_in) if (!m->wtc.count(w.qs)) return false;
    for (const QLinear &w : m->linears) if (!m->wtc.count(w.qs)) return false;
    for (const LayerW &l : m->dep_layers)
        for (const std::vector<QLinear> *v : {&l.in_proj, &l.out_proj, &l.lin_in, &l.lin_out})
            for (const QLinear &w : *v) if (!m->wtc.count(w.qs)) return false;
    return true;
}
extern "C" int msx_batch_create(msx_model *m, int n_streams, int context_override, msx_batch **out) {
    return batch_create_impl(m, n_streams, context_override, nullptr, out);
}
static int batch_create_impl(msx_model *m, int n_streams, int context_override, msx_stream *prefill_of, msx_batch **out) {
    if (!m || !out) return fail(MSX_ERR_ARG, "null argument");
    *out = nullptr;
    // a prefill context of 64 columns runs its linears on the tcgen05 GEMM (Q4_K models whose temporal linears have the tc layout)
    // ... and so does a batch of 9..64 streams when every linear of the step has it
    const bool tc_mode = prefill_of ? (n_streams == tc::kN && tc_prefill_ok(m)) : (n_streams > kMmaCols && n_streams <= tc::kN && tc_batch_ok(m));
    if (n_streams < 1 || (n_streams > kMmaCols && !tc_mode))
        return fail(MSX_ERR_ARG, "a batch holds 1..8 streams (up to 64 for Q4_K models whose linears all have 128-row tiles and 256-multiple inner dimensions)");
    if (m->cfg.cross_attention || m->cfg.demux_second_stream || m->cfg.dep_low_rank)
        return fail(MSX_ERR_ARG, "batched streams do not cover the TTS-family layers (cross-attention, demux / low-rank embeddings)");
    CU(cudaSetDevice(m->device));
    if (int e = set_smem_attrs()) return e;
    CU(cudaFuncSetAttribute(dq_matmul_mma_kernel<12, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemMax));
    CU(cudaFuncSetAttribute(dq_matmul_mma_kernel<12, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemMax));
    CU(cudaFuncSetAttribute(dq_matmul_mma_kernel<12, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemMax));
    CU(cudaFuncSetAttribute(dq_matmul_mma_kernel<8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemMax));
    CU(cudaFuncSetAttribute(dq_matmul_mma_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemMax));
    if (tc_mode) {
        CU(cudaFuncSetAttribute(tc::tc_matmul_q4k_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes));
        CU(cudaFuncSetAttribute(tc::tc_matmul_q4k_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes));
        CU(cudaFuncSetAttribute(tc::tc_matmul_q4k_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes));
    }
    else if (int e = ensure_all_tiles(m)) return e;
    std::unique_ptr<msx_batch> b(new msx_batch);
    b->m = m; b->n = n_streams; b->n_active = n_streams; b->prefill_of = prefill_of; b->tc = tc_mode;
    const msx_config &c = m->cfg;
    const size_t n = (size_t)n_streams;
    b->cap = context_override > 0 ? std::min(context_override, c.context) : c.context;
    b->attn_split = attn_split_for(c.num_heads * n_streams, b->cap, m->num_sms);
    b->host_offset.assign(n_streams, 0);
    if (prefill_of) b->st = prefill_of->st;
    else CU(cudaStreamCreateWithFlags(&b->st, cudaStreamNonBlocking));
    CU(cudaEventCreate(&b->ev0)); CU(cudaEventCreate(&b->ev1));
    CU(cudaMallocHost((void **)&b->h_hdr, kCtrlInOffset * n));
    CU(cudaMallocHost((void **)&b->h_in, kCtrlInBytes * n));
    CU(cudaMallocHost((void **)&b->h_out, kCtrlOutBytes * n));
    if (int e = balloc(b.get(), (void **)&b->ctrl, sizeof(Ctrl) * n)) return e;
    const size_t kv = (size_t)c.num_layers * b->cap * c.dim;
    if (!prefill_of) {
        if (int e = balloc(b.get(), (void **)&b->kc, kv * 2 * n)) return e;
        if (int e = balloc(b.get(), (void **)&b->vc, kv * 2 * n)) return e;
    }
    if (int e = balloc(b.get(), (void **)&b->x, n * c.dim * 4)) return e;
    if (int e = balloc(b.get(), (void **)&b->qkv, n * c.dim * 3 * 4)) return e;
    if (int e = balloc(b.get(), (void **)&b->ctx, n * c.dim * 4)) return e;
    if (int e = balloc(b.get(), (void **)&b->gate, n * m->hidden * 4)) return e;
    if (int e = balloc(b.get(), (void **)&b->tout, n * c.dim * 4)) return e;
    if (int e = balloc(b.get(), (void **)&b->text_logits, n * c.text_card * 4)) return e;
    if (int e = balloc(b.get(), (void **)&b->rope_cs, n * (c.dim / c.num_heads) * 4)) return e;
    int maxK = std::max(c.dim, m->hidden);
    if (prefill_of) {
        const size_t vb = (size_t)c.num_heads * kVictimRows * (c.dim / c.num_heads) * 2;
        if (int e = balloc(b.get(), (void **)&b->victim_k, vb)) return e;
        if (int e = balloc(b.get(), (void **)&b->victim_v, vb)) return e;
    }
    if (c.dep_q > 0 && !prefill_of) {
        const size_t dkv = (size_t)c.dep_layers * m->dep_cap * c.dep_dim;
        if (int e = balloc(b.get(), (void **)&b->dkc, dkv * 2 * n)) return e;
        if (int e = balloc(b.get(), (void **)&b->dvc, dkv * 2 * n)) return e;
        if (int e = balloc(b.get(), (void **)&b->dx, n * c.dep_dim * 4)) return e;
        if (int e = balloc(b.get(), (void **)&b->dqkv, n * c.dep_dim * 3 * 4)) return e;
        if (int e = balloc(b.get(), (void **)&b->dctx, n * c.dep_dim * 4)) return e;
        if (int e = balloc(b.get(), (void **)&b->dgate, n * m->dep_hidden * 4)) return e;
        if (int e = balloc(b.get(), (void **)&b->audio_logits, n * c.dep_q * c.card * 4)) return e;
        maxK = std::max(maxK, std::max(c.dep_dim, m->dep_hidden));
    }
    const int wt = m->text_linear.type;
    if (!tc_mode && gemm_stages_for(maxK, wt) < 2) return fail(MSX_ERR_ARG, "inner dimension too large for the batched GEMM's shared-memory image");
    if (int e = balloc(b.get(), (void **)&b->img, tc_mode ? tc::image_bytes(maxK) : (size_t)act_image_bytes(maxK, wt))) return e;
    if (tc_mode) {
        int max_tiles = 0;
        for (const QLinear *w : {&m->layers[0].in_proj[0], &m->layers[0].out_proj[0], &m->layers[0].lin_in[0], &m->layers[0].lin_out[0], &m->text_linear})
            max_tiles = std::max(max_tiles, w->rows / tc::kM);
        for (const LayerW &l : m->dep_layers)
            for (const std::vector<QLinear> *v : {&l.in_proj, &l.out_proj, &l.lin_in, &l.lin_out})
                for (const QLinear &w : *v) max_tiles = std::max(max_tiles, w.rows / tc::kM);
        for (const QLinear &w : m->dep_in) max_tiles = std::max(max_tiles, w.rows / tc::kM);
        for (const QLinear &w : m->linears) max_tiles = std::max(max_tiles, w.rows / tc::kM);
        b->tc_max_tiles = max_tiles;
        if (int e = balloc(b.get(), (void **)&b->tc_partial, tc::partial_bytes(m->num_sms))) return e;
        if (int e = balloc(b.get(), (void **)&b->tc_tickets, (size_t)max_tiles * 4)) return e;
    }
    if (int e = balloc(b.get(), (void **)&b->img_tout, tc_mode ? tc::image_bytes(c.dim) : (size_t)act_image_bytes(c.dim, wt))) return e;
    if (!prefill_of) {
        const size_t nf = n * (1 + MSX_MAX_STEPS) * kSampleMaxK;
        if (int e = balloc(b.get(), (void **)&b->d_noise, nf * 4)) return e;
        if (int e = balloc(b.get(), (void **)&b->d_probs, n * std::max(c.text_card, c.card) * 4)) return e;
        CU(cudaMallocHost((void **)&b->h_noise, nf * 4));
        for (size_t i = 0; i < nf; i++) b->h_noise[i] = 1.f;
    }
    std::vector<Ctrl> hc(n_streams);
    memset(hc.data(), 0, sizeof(Ctrl) * n);
    for (Ctrl &h : hc) { h.n_in = c.n_q + 1; h.text_override = INT32_MIN; for (int i = 0; i < 40; i++) h.force[i] = INT32_MIN; }
    CU(cudaMemcpy(b->ctrl, hc.data(), sizeof(Ctrl) * n, cudaMemcpyHostToDevice));
    if (int e = capture_b(b.get(), [&](BatchLauncher &B) { enqueue_temporal_b(B); }, &b->g_temporal, &b->launches_temporal)) return e;
    if (c.dep_q > 0 && !prefill_of)
        if (int e = capture_b(b.get(), [&](BatchLauncher &B) { enqueue_depformer_b(B); }, &b->g_depformer, &b->launches_depformer)) return e;
    CU(cudaStreamSynchronize(b->st));
    *out = b.release();
    return 0;
}

// ---- batched-T prompt prefill -------------------------------------------------------------------------------------
// Prompt frames (PersonaPlex voice / system prompt rows, lm.h:983-1134: all n_q+1 tokens given) only populate the
// temporal KV rings — their logits, sampled tokens and depformer results are discarded by moshi_lmgen_step when the
// tokens are "provided" (lm.h:933-943).  Eight consecutive positions are therefore run as the eight columns of the
// tensor-core GEMM: every weight matrix is read once per 8 prompt frames instead of once per frame.
extern "C" int msx_stream_prefill(msx_stream *s, const int32_t *tokens, int T) {
    if (!s || !tokens || T <= 0) return fail(MSX_ERR_ARG, "bad argument");
    msx_model *m = s->m; const msx_config &c = m->cfg;
    if (m->tp_world > 1 || c.cross_attention || c.demux_second_stream) return fail(MSX_ERR_STATE, "prefill covers plain single-GPU models");
    if (s->host_offset < 0 || T > (1 << 24)) return fail(MSX_ERR_ARG, "bad prompt length");
    CU(cudaSetDevice(m->device));
    if (!s->prefill) {
        msx_batch *pb = nullptr;
        if (int e = batch_create_impl(m, tc_prefill_ok(m) ? tc::kN : kMmaCols, s->cap, s, &pb)) return e;
        s->prefill = pb;
    }
    msx_batch *b = s->prefill;
    const int n_in = c.n_q + 1;
    for (int c0 = 0, nb = 0; c0 < T; c0 += nb) {
        // all columns of a pass are inserted before any of them attends; the old rows of the slots they overwrite are kept aside
        // (AttnArgs::victim_*) for the columns that still see them, so a pass is b->n (64 or 8) positions anywhere on the ring
        nb = std::min(std::min(b->n, T - c0), s->cap);
        CU(cudaStreamSynchronize(b->st));            // the pinned staging buffers are reused per chunk
        for (int j = 0; j < b->n; j++) {
            Ctrl hdr; memset(&hdr, 0, sizeof(hdr));
            hdr.offset = s->host_offset + c0 + std::min(j, nb - 1); hdr.n_in = n_in;
            memcpy(b->h_hdr + (size_t)j * kCtrlInOffset, &hdr, kCtrlInOffset);
        }
        CU(cudaMemcpy2DAsync(b->ctrl, sizeof(Ctrl), b->h_hdr, kCtrlInOffset, kCtrlInOffset, b->n, cudaMemcpyHostToDevice, b->st));
        std::vector<int32_t> tk((size_t)b->n * n_in, 0);
        memcpy(tk.data(), tokens + (size_t)c0 * n_in, (size_t)nb * n_in * 4);
        if (int e = push_inputs_b(b, tk.data())) return e;
        if (nb == b->n) {
            CU(cudaGraphLaunch(b->g_temporal, b->st));
        } else {                                       // tail: exactly nb live columns, launched eagerly
            Launcher L{b->st, m->num_sms};
            BatchLauncher B{L, b};
            b->n_active = nb;
            enqueue_temporal_b(B);
            b->n_active = b->n;
            if (B.err) return B.err;
            if (L.err != cudaSuccess) return fail(MSX_ERR_CUDA, std::string("prefill launch: ") + cudaGetErrorString(L.err));
        }
    }
    s->host_offset += T;
    CU(cudaMemcpyAsync(&s->ctrl->offset, &s->host_offset, 4, cudaMemcpyHostToDevice, s->st));
    CU(cudaStreamSynchronize(s->st));
    return 0;
}

// per-family kernel time of ONE eagerly launched full prefill pass (CUDA event after every launch; the positions are inserted like
// msx_stream_prefill would, so the stream advances by the pass).  Tool for profiles/, not on the product path.
extern "C" int msx_stream_prefill_profile(msx_stream *s, const int32_t *tokens, float *family_ms, int32_t *family_launches, int max_families) {
    if (!s || !tokens || !family_ms || !family_launches) return fail(MSX_ERR_ARG, "null argument");
    if (int e = msx_stream_prefill(s, tokens, 1)) return e;              // creates the prefill context
    msx_batch *b = s->prefill; msx_model *m = s->m; const msx_config &c = m->cfg;
    if (b->n > s->cap) return fail(MSX_ERR_STATE, "ring shorter than a pass");
    const int n_in = c.n_q + 1;
    for (int i = 0; i < max_families; i++) { family_ms[i] = 0.f; family_launches[i] = 0; }
    CU(cudaStreamSynchronize(b->st));
    for (int j = 0; j < b->n; j++) {
        Ctrl hdr; memset(&hdr, 0, sizeof(hdr));
        hdr.offset = s->host_offset + j; hdr.n_in = n_in;
        memcpy(b->h_hdr + (size_t)j * kCtrlInOffset, &hdr, kCtrlInOffset);
    }
    CU(cudaMemcpy2DAsync(b->ctrl, sizeof(Ctrl), b->h_hdr, kCtrlInOffset, kCtrlInOffset, b->n, cudaMemcpyHostToDevice, b->st));
    if (int e = push_inputs_b(b, tokens + n_in)) return e;
    std::vector<cudaEvent_t> ev;
    std::vector<int> fam;
    Launcher L{b->st, m->num_sms};
    L.events = &ev; L.families = &fam;
    BatchLauncher B{L, b};
    enqueue_temporal_b(B);
    if (B.err) return B.err;
    if (L.err != cudaSuccess) return fail(MSX_ERR_CUDA, std::string("prefill launch: ") + cudaGetErrorString(L.err));
    CU(cudaStreamSynchronize(b->st));
    for (size_t i = 0; i + 1 < ev.size(); i++) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
        const int f = fam[i];
        if (f < max_families) { family_ms[f] += ms; family_launches[f] += 1; }
    }
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    s->host_offset += b->n;
    CU(cudaMemcpyAsync(&s->ctrl->offset, &s->host_offset, 4, cudaMemcpyHostToDevice, s->st));
    CU(cudaStreamSynchronize(s->st));
    return 0;
}

static int batch_build_graphs(msx_batch *b) {
    const msx_config &c = b->m->cfg;
    if (b->g_temporal) { cudaGraphExecDestroy(b->g_temporal); b->g_temporal = nullptr; }
    if (b->g_depformer) { cudaGraphExecDestroy(b->g_depformer); b->g_depformer = nullptr; }
    if (int e = capture_b(b, [&](BatchLauncher &B) { enqueue_temporal_b(B); }, &b->g_temporal, &b->launches_temporal)) return e;
    if (c.dep_q > 0 && !b->prefill_of)
        if (int e = capture_b(b, [&](BatchLauncher &B) { enqueue_depformer_b(B); }, &b->g_depformer, &b->launches_depformer)) return e;
    return 0;
}

// temperatures / top-k of every stream of the batch (moshi_lm_start: 250 audio / 25 text); temperature <= 0 = greedy
extern "C" int msx_batch_set_sampling(msx_batch *b, float temp_text, float temp_audio, int top_k_text, int top_k_audio) {
    if (!b || b->prefill_of) return fail(MSX_ERR_ARG, "bad batch");
    if (top_k_text < 1 || top_k_audio < 1) return fail(MSX_ERR_ARG, "top_k must be >= 1");
    if (std::min(top_k_text, b->m->cfg.text_card) > kSampleMaxK || std::min(top_k_audio, b->m->cfg.card) > kSampleMaxK)
        return fail(MSX_ERR_ARG, "top_k > 256 is not supported");
    CU(cudaSetDevice(b->m->device));
    CU(cudaStreamSynchronize(b->st));
    b->temp_text = temp_text; b->temp_audio = temp_audio; b->top_k_text = top_k_text; b->top_k_audio = top_k_audio;
    if (int e = batch_build_graphs(b)) return e;
    CU(cudaStreamSynchronize(b->st));
    return 0;
}

// Exp(1) draws of the next frame: noise_text [n][kt], noise_audio [n][dep_q][ka] (kt / ka = min(top_k, cardinality, 256))
extern "C" int msx_batch_set_noise(msx_batch *b, const float *noise_text, const float *noise_audio) {
    if (!b || b->prefill_of) return fail(MSX_ERR_ARG, "bad batch");
    const msx_config &c = b->m->cfg;
    CU(cudaSetDevice(b->m->device));
    CU(cudaStreamSynchronize(b->st));
    const int kt = std::min(std::min(b->top_k_text, c.text_card), kSampleMaxK), ka = std::min(std::min(b->top_k_audio, c.card), kSampleMaxK);
    const size_t per = (size_t)(1 + MSX_MAX_STEPS) * kSampleMaxK;
    for (int s = 0; s < b->n; s++) {
        float *h = b->h_noise + per * s;
        if (noise_text) memcpy(h, noise_text + (size_t)s * kt, (size_t)kt * 4);
        if (noise_audio) for (int k = 0; k < c.dep_q; k++) memcpy(h + (size_t)(1 + k) * kSampleMaxK, noise_audio + ((size_t)s * c.dep_q + k) * ka, (size_t)ka * 4);
    }
    CU(cudaMemcpyAsync(b->d_noise, b->h_noise, per * b->n * 4, cudaMemcpyHostToDevice, b->st));
    return 0;
}

extern "C" void msx_batch_free(msx_batch *b) { delete b; }
extern "C" int msx_batch_size(const msx_batch *b) { return b ? b->n : 0; }
extern "C" int msx_batch_launches_per_frame(const msx_batch *b) { return b ? b->launches_temporal + b->launches_depformer : 0; }
extern "C" int msx_batch_offset(const msx_batch *b, int stream) { return (b && stream >= 0 && stream < b->n) ? b->host_offset[stream] : -1; }
extern "C" int64_t msx_batch_kv_bytes_next(const msx_batch *b) {
    if (!b) return 0;
    int64_t t = 0;
    for (int s = 0; s < b->n; s++) t += (int64_t)std::min(b->host_offset[s] + 1, b->cap) * 2 * b->m->cfg.dim * 2 * b->m->cfg.num_layers;
    return t;
}

// stream < 0: all streams.  A stream can be restarted while the others keep their context.
extern "C" int msx_batch_reset_stream(msx_batch *b, int stream) {
    if (!b || stream >= b->n) return fail(MSX_ERR_ARG, "bad argument");
    const msx_config &c = b->m->cfg;
    CU(cudaSetDevice(b->m->device));
    CU(cudaStreamSynchronize(b->st));
    const size_t kv = (size_t)c.num_layers * b->cap * c.dim * 2;
    const size_t dkv = c.dep_q > 0 ? (size_t)c.dep_layers * b->m->dep_cap * c.dep_dim * 2 : 0;
    for (int s = (stream < 0 ? 0 : stream); s < (stream < 0 ? b->n : stream + 1); s++) {
        CU(cudaMemsetAsync(reinterpret_cast<uint8_t *>(b->kc) + kv * s, 0, kv, b->st));
        CU(cudaMemsetAsync(reinterpret_cast<uint8_t *>(b->vc) + kv * s, 0, kv, b->st));
        if (dkv) {
            CU(cudaMemsetAsync(reinterpret_cast<uint8_t *>(b->dkc) + dkv * s, 0, dkv, b->st));
            CU(cudaMemsetAsync(reinterpret_cast<uint8_t *>(b->dvc) + dkv * s, 0, dkv, b->st));
        }
        CU(cudaMemsetAsync(b->tout + (size_t)s * c.dim, 0, (size_t)c.dim * 4, b->st));
        CU(cudaMemsetAsync(&b->ctrl[s].offset, 0, 4, b->st));
        b->host_offset[s] = 0;
    }
    CU(cudaStreamSynchronize(b->st));
    return 0;
}

// tokens [n][n_q+1] -> out_tokens [n][1+dep_q]: one fused frame for every stream of the batch
extern "C" int msx_batch_step(msx_batch *b, const int32_t *tokens, int32_t *out_tokens) {
    if (!b || !tokens) return fail(MSX_ERR_ARG, "null argument");
    const msx_config &c = b->m->cfg;
    CU(cudaSetDevice(b->m->device));
    if (int e = push_inputs_b(b, tokens)) return e;
    CU(cudaGraphLaunch(b->g_temporal, b->st));
    if (c.dep_q > 0) CU(cudaGraphLaunch(b->g_depformer, b->st));
    for (int &o : b->host_offset) o++;
    if (int e = pull_outputs_b(b)) return e;
    const int w = (int)(kCtrlOutBytes / 4);
    if (out_tokens)
        for (int s = 0; s < b->n; s++)
            for (int k = 0; k < 1 + c.dep_q; k++) out_tokens[(size_t)s * (1 + c.dep_q) + k] = b->h_out[(size_t)s * w + k];
    return 0;
}

// logits of the last step of one stream (parity checks)
extern "C" int msx_batch_get_logits(msx_batch *b, int stream, float *text_logits, float *audio_logits) {
    if (!b || stream < 0 || stream >= b->n) return fail(MSX_ERR_ARG, "bad argument");
    const msx_config &c = b->m->cfg;
    CU(cudaSetDevice(b->m->device));
    CU(cudaStreamSynchronize(b->st));
    if (text_logits) CU(cudaMemcpy(text_logits, b->text_logits + (size_t)stream * c.text_card, (size_t)c.text_card * 4, cudaMemcpyDeviceToHost));
    if (audio_logits && c.dep_q > 0)
        CU(cudaMemcpy(audio_logits, b->audio_logits + (size_t)stream * c.dep_q * c.card, (size_t)c.dep_q * c.card * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// frames [n][n_frames][n_q+1] resident on the device, n_steps frames replayed back to back for every stream
extern "C" int msx_batch_run_resident(msx_batch *b, const int32_t *frames, int n_frames, int n_steps, int32_t *out_tokens, float *elapsed_ms) {
    if (!b || !frames || n_frames <= 0 || n_steps <= 0) return fail(MSX_ERR_ARG, "bad argument");
    if (int e = check_tokens(b->m, frames, b->n * n_frames, INT32_MIN, nullptr)) return e;
    const msx_config &c = b->m->cfg;
    CU(cudaSetDevice(b->m->device));
    const int n_in = c.n_q + 1, n_out = 1 + c.dep_q;
    int32_t *d_feed = nullptr, *d_trace = nullptr;
    CU(cudaMalloc((void **)&d_feed, (size_t)b->n * n_frames * n_in * 4));
    if (out_tokens) CU(cudaMalloc((void **)&d_trace, (size_t)b->n * n_steps * n_out * 4));
    CU(cudaMemcpy(d_feed, frames, (size_t)b->n * n_frames * n_in * 4, cudaMemcpyHostToDevice));
    if (int e = push_inputs_b(b, nullptr)) return e;
    for (int s = 0; s < b->n; s++) {
        Ctrl hdr;
        memset(&hdr, 0, sizeof(hdr));
        hdr.offset = b->host_offset[s]; hdr.frame = 0; hdr.feed_n = n_frames; hdr.n_in = n_in;
        hdr.feed = d_feed + (size_t)s * n_frames * n_in;
        hdr.trace = d_trace ? d_trace + (size_t)s * n_steps * n_out : nullptr;
        CU(cudaMemcpyAsync(&b->ctrl[s], &hdr, kCtrlInOffset, cudaMemcpyHostToDevice, b->st));
        CU(cudaStreamSynchronize(b->st));      // hdr is a stack object
    }
    CU(cudaEventRecord(b->ev0, b->st));
    for (int i = 0; i < n_steps; i++) {
        CU(cudaGraphLaunch(b->g_temporal, b->st));
        if (c.dep_q > 0) CU(cudaGraphLaunch(b->g_depformer, b->st));
    }
    CU(cudaEventRecord(b->ev1, b->st));
    CU(cudaStreamSynchronize(b->st));
    for (int &o : b->host_offset) o += n_steps;
    if (elapsed_ms) CU(cudaEventElapsedTime(elapsed_ms, b->ev0, b->ev1));
    if (out_tokens) CU(cudaMemcpy(out_tokens, d_trace, (size_t)b->n * n_steps * n_out * 4, cudaMemcpyDeviceToHost));
    int32_t zero2[2] = {0, 0};
    for (int s = 0; s < b->n; s++) CU(cudaMemcpy(&b->ctrl[s].frame, zero2, 8, cudaMemcpyHostToDevice));
    cudaFree(d_feed);
    if (d_trace) cudaFree(d_trace);
    return 0;
}

// per-family kernel time of one eager batched frame (CUDA event after every launch)
extern "C" int msx_batch_profile_frame(msx_batch *b, const int32_t *tokens, float *family_ms, int32_t *family_launches, int max_families) {
    if (!b || !tokens || !family_ms || !family_launches) return fail(MSX_ERR_ARG, "null argument");
    const msx_config &c = b->m->cfg;
    CU(cudaSetDevice(b->m->device));
    for (int i = 0; i < max_families; i++) { family_ms[i] = 0.f; family_launches[i] = 0; }
    if (int e = push_inputs_b(b, tokens)) return e;
    std::vector<cudaEvent_t> ev;
    std::vector<int> fam;
    Launcher L{b->st, b->m->num_sms};
    L.events = &ev; L.families = &fam;
    BatchLauncher B{L, b};
    enqueue_temporal_b(B);
    if (c.dep_q > 0) enqueue_depformer_b(B);
    for (int &o : b->host_offset) o++;
    if (L.err != cudaSuccess) return fail(MSX_ERR_CUDA, std::string("launch failed: ") + cudaGetErrorString(L.err));
    if (int e = pull_outputs_b(b)) return e;
    for (size_t i = 0; i + 1 < ev.size(); i++) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
        const int f = fam[i];
        if (f < max_families) { family_ms[f] += ms; family_launches[f] += 1; }
    }
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    return 0;
}

// test hook: y[nb][rows] = W x[nb][k] through quant_q8k_kernel + dq_matmul_q4k_kernel (EPI_STORE)
extern "C" int msx_test_gemm_batch(int device, int type, const void *w, int64_t k, int64_t rows, const float *x, int nb, const float *alpha,
                                   float *y) {
    if (!w || !x || !y || nb < 1 || nb > kMmaCols) return fail(MSX_ERR_ARG, "bad argument");
    std::unique_ptr<msx_model> m;
    if (int e = device_setup(device, m)) return e;
    CU(cudaFuncSetAttribute(dq_matmul_q4k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemMax));
    QLinear ql;
    if (int e = upload_linear(m.get(), w, type, k, rows, 0, &ql)) return e;
    QTiles qt;
    if (int e = tiles_of(m.get(), ql, &qt)) return e;
    CU(cudaFuncSetAttribute(dq_matmul_q8_0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemMax));
    if (gemm_stages_for((int)k, type) < 2) return fail(MSX_ERR_ARG, "k too large");
    float *dx = nullptr, *dy = nullptr, *da = nullptr; uint8_t *img = nullptr;
    if (int e = dev_alloc(m.get(), (void **)&dx, (size_t)nb * k * 4)) return e;
    if (int e = dev_alloc(m.get(), (void **)&dy, (size_t)nb * rows * 4)) return e;
    if (int e = dev_alloc(m.get(), (void **)&img, (size_t)act_image_bytes((int)k, type))) return e;
    CU(cudaMemset(img, 0xff, (size_t)act_image_bytes((int)k, type)));      // dead columns hold garbage in production too
    CU(cudaMemcpy(dx, x, (size_t)nb * k * 4, cudaMemcpyHostToDevice));
    if (alpha) {
        if (int e = dev_alloc(m.get(), (void **)&da, (size_t)k * 4)) return e;
        CU(cudaMemcpy(da, alpha, (size_t)k * 4, cudaMemcpyHostToDevice));
    }
    QuantArgs q;
    q.x = dx; q.ld = (int)k; q.alpha = da; q.eps = 1e-8f; q.img = img; q.K = (int)k;
    if (type == T_Q4_K) quant_q8k_kernel<<<dim3(nb, quant_parts_for((int)k)), kGemmThreads>>>(q);
    else quant_q8_0_kernel<<<dim3(nb, quant_parts_for((int)k)), kGemmThreads>>>(q);
    MatmulArgs g;
    g.w = qt; g.img = img; g.out = dy; g.ld = (int)rows; g.nb = nb; g.epi = EPI_STORE;
    g.stages = gemm_stages_for((int)k, type);
    if (type == T_Q4_K) dq_matmul_q4k_kernel<<<gemm_grid_for(qt.n_tiles, m->num_sms), kGemmThreads, gemm_smem_bytes((int)k, g.stages, 12)>>>(g);
    else dq_matmul_q8_0_kernel<<<gemm_grid_for(qt.n_tiles, m->num_sms), kGemmThreads, gemm_smem_bytes((int)k, g.stages, 8)>>>(g);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(y, dy, (size_t)nb * rows * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// micro-benchmark: quant + GEMM pairs over n_mats rotating copies of one matrix, replayed from a CUDA graph
extern "C" int msx_bench_gemm_batch_ex(int device, const void *w, int64_t k, int64_t rows, int nb, int n_mats, int iters, int epilogue,
                                       int with_quant, float *avg_us, long long *stamps_out /* [iters][148][8] or null */);
extern "C" int msx_bench_gemm_batch(int device, const void *w, int64_t k, int64_t rows, int nb, int n_mats, int iters, int epilogue,
                                    int with_quant, float *avg_us) {
    return msx_bench_gemm_batch_ex(device, w, k, rows, nb, n_mats, iters, epilogue, with_quant, avg_us, nullptr);
}
extern "C" int msx_bench_gemm_batch_ex(int device, const void *w, int64_t k, int64_t rows, int nb, int n_mats, int iters, int epilogue,
                                       int with_quant, float *avg_us, long long *stamps_out) {
    if (!w || !avg_us || n_mats < 1 || iters < 1 || nb < 1 || nb > kMmaCols) return fail(MSX_ERR_ARG, "bad argument");
    std::unique_ptr<msx_model> m;
    if (int e = device_setup(device, m)) return e;
    CU(cudaFuncSetAttribute(dq_matmul_q4k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemMax));
    std::vector<QTiles> mats(n_mats);
    for (int i = 0; i < n_mats; i++) {
        QLinear ql;
        if (int e = upload_linear(m.get(), w, T_Q4_K, k, rows, epilogue == EPI_GATE ? (int)(rows / 2) : 0, &ql)) return e;
        if (int e = tiles_of(m.get(), ql, &mats[i])) return e;
    }
    float *dx = nullptr, *dy = nullptr; uint8_t *img = nullptr; Ctrl *ctrl = nullptr;
    const size_t outn = (size_t)nb * std::max<int64_t>(rows, k);
    if (int e = dev_alloc(m.get(), (void **)&dx, (size_t)nb * k * 4)) return e;
    if (int e = dev_alloc(m.get(), (void **)&dy, outn * 4)) return e;
    if (int e = dev_alloc(m.get(), (void **)&img, (size_t)act_image_bytes((int)k))) return e;
    if (int e = dev_alloc(m.get(), (void **)&ctrl, sizeof(Ctrl) * nb)) return e;
    std::vector<float> hx((size_t)nb * k);
    for (size_t i = 0; i < hx.size(); i++) hx[i] = (float)((i * 2654435761u) % 2001) / 1000.f - 1.f;
    CU(cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
    CU(cudaMemset(dy, 0, outn * 4));
    CU(cudaMemset(ctrl, 0, sizeof(Ctrl) * nb));
    CU(cudaMemset(img, 0, (size_t)act_image_bytes((int)k)));
    long long *d_stamps = nullptr;
    if (stamps_out) { if (int e = dev_alloc(m.get(), (void **)&d_stamps, (size_t)iters * 148 * 8 * 8)) return e; CU(cudaMemset(d_stamps, 0, (size_t)iters * 148 * 8 * 8)); }
    cudaStream_t st;
    CU(cudaStreamCreate(&st));
    Launcher L{st, m->num_sms};
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < iters; i++) {
        QuantArgs q;
        q.x = dx; q.ld = (int)k; q.eps = 1e-8f; q.img = img; q.K = (int)k;
        if (with_quant || i == 0) L.launch_pdl(quant_q8k_kernel, dim3(nb, quant_parts_for((int)k)), dim3(kGemmThreads), 0, q);
        MatmulArgs g;
        g.w = mats[i % n_mats]; g.img = img; g.out = dy; g.ld = (int)(epilogue == EPI_GATE ? rows / 2 : rows); g.nb = nb; g.epi = epilogue; g.ctrl = ctrl;
        if (d_stamps) g.stamps = d_stamps + (size_t)i * 148 * 8;
        g.stages = gemm_stages_for((int)k);
        L.launch_pdl(dq_matmul_q4k_kernel, dim3(gemm_grid_for(g.w.n_tiles, m->num_sms)), dim3(kGemmThreads), (size_t)gemm_smem_bytes((int)k, g.stages), g);
    }
    CU(cudaStreamEndCapture(st, &graph));
    if (L.err != cudaSuccess) return fail(MSX_ERR_CUDA, std::string("gemm launch: ") + cudaGetErrorString(L.err));
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
    if (stamps_out) CU(cudaMemcpy(stamps_out, d_stamps, (size_t)iters * 148 * 8 * 8, cudaMemcpyDeviceToHost));
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(st);
    return 0;
}


// ---- batched generator: one delay ring (moshi_lmgen_state_t) per stream of a batch -------------------------------------
// The per-stream host logic is msx_gen's (gen_prepare / gen_finish, lm.h:778-979); the model step is ONE msx_batch_step.
struct msx_bgen {
    msx_batch *b = nullptr;
    std::vector<std::unique_ptr<msx_gen>> gens;
    std::vector<int32_t> inputs, outs;
};

extern "C" int msx_bgen_create(msx_batch *b, int delay_steps, msx_bgen **out) {
    if (!b || !out || b->prefill_of) return fail(MSX_ERR_ARG, "bad argument");
    std::unique_ptr<msx_bgen> g(new msx_bgen);
    g->b = b;
    for (int i = 0; i < b->n; i++) {
        std::unique_ptr<msx_gen> gi(new msx_gen);
        gi->cfg = b->m->cfg;
        gen_init(gi.get(), delay_steps);
        g->gens.push_back(std::move(gi));
    }
    g->inputs.resize((size_t)b->n * (b->m->cfg.n_q + 1));
    g->outs.resize((size_t)b->n * (1 + b->m->cfg.dep_q));
    *out = g.release();
    return 0;
}
extern "C" void msx_bgen_free(msx_bgen *g) { delete g; }
extern "C" int msx_bgen_offset(const msx_bgen *g, int stream) { return (g && stream >= 0 && stream < g->b->n) ? g->gens[stream]->offset : -1; }

// a new conversation takes over slot `stream`: fresh delay ring, KV rings cleared, position 0; the other slots go on
extern "C" int msx_bgen_reset_stream(msx_bgen *g, int stream) {
    if (!g || stream < 0 || stream >= g->b->n) return fail(MSX_ERR_ARG, "bad argument");
    const int ds = g->gens[stream]->delay_steps;
    std::unique_ptr<msx_gen> gi(new msx_gen);
    gi->cfg = g->b->m->cfg;
    gen_init(gi.get(), ds);
    g->gens[stream] = std::move(gi);
    return msx_batch_reset_stream(g->b, stream);
}

// in_tokens [n][n_in] (n_in = the user codes of a frame, or n_q+1 when everything is provided) -> out_text [n],
// out_audio [n][dep_q], valid [n] (1 = this stream emitted a frame, 0 = still inside its delay window)
extern "C" int msx_bgen_step(msx_bgen *g, const int32_t *in_tokens, int n_in, int32_t *out_text, int32_t *out_audio, int32_t *valid) {
    if (!g || !out_text || !out_audio || !valid) return fail(MSX_ERR_ARG, "null argument");
    msx_batch *b = g->b; const msx_config &c = b->m->cfg;
    const int n = b->n, ncb = c.n_q + 1;
    std::vector<GenPrep> prep(n);
    for (int i = 0; i < n; i++) {
        if (int e = gen_prepare(g->gens[i].get(), in_tokens ? in_tokens + (size_t)i * n_in : nullptr, n_in, &prep[i])) return e;
        memcpy(g->inputs.data() + (size_t)i * ncb, prep[i].input, (size_t)ncb * 4);
    }
    if (b->temp_text > 0.f || b->temp_audio > 0.f) {
        // Exp(1) draws like context.h:464-480 (text candidates, then the codebooks), every conversation from its own generator state
        const int kt = std::min(std::min(b->top_k_text, c.text_card), kSampleMaxK), ka = std::min(std::min(b->top_k_audio, c.card), kSampleMaxK);
        std::vector<float> nt((size_t)n * kt, 1.f), na((size_t)n * std::max(1, c.dep_q) * ka, 1.f);
        for (int i = 0; i < n; i++) {
            if (b->temp_text > 0.f) for (int j = 0; j < kt; j++) nt[(size_t)i * kt + j] = g->gens[i]->rng.exp1();
            if (b->temp_audio > 0.f) for (size_t j = 0; j < (size_t)c.dep_q * ka; j++) na[(size_t)i * c.dep_q * ka + j] = g->gens[i]->rng.exp1();
        }
        if (int e = msx_batch_set_noise(b, nt.data(), na.data())) return e;
    }
    if (int e = msx_batch_step(b, g->inputs.data(), g->outs.data())) return e;
    for (int i = 0; i < n; i++) {
        int32_t out[1 + MSX_MAX_STEPS];
        for (int k = 0; k < 1 + MSX_MAX_STEPS; k++) out[k] = -1;
        memcpy(out, g->outs.data() + (size_t)i * (1 + c.dep_q), (size_t)(1 + c.dep_q) * 4);
        const int rc = gen_finish(g->gens[i].get(), prep[i], out, 0, &out_text[i], out_audio + (size_t)i * c.dep_q);
        if (rc < 0) return rc;
        valid[i] = rc;
    }
    return 0;
}
