// stream.inl — msx_stream: per-conversation state (KV rings, control block, scratch), enqueue of the temporal / depformer
// launch chains, CUDA-graph capture, msx_stream_* and msx_step_* entry points.  Reference: StateContext / ScratchContext
// src/context.h:227-780, graph build lm.h:446-553, 659-690.  Included by engine.cu.

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
    auto reduce_into_x = [&](MatvecArgs &gg, int family) {
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
    MatvecArgs g;
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
    a.small_ctx = 64;        // up to 64 valid slots one CTA per head handles the ring alone (no cluster barriers; profiles/r2_attention.md)
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
        MatvecArgs g1;
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
    MatvecArgs g;
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
        MatvecMulti mm;
        for (int k = 0; k < c.dep_q; k++) {
            const QLinear &w = m->dep_in[weights_of(k)];
            mm.qs[k] = w.qs; mm.sc[k] = w.sc; mm.dd[k] = w.dd; mm.out[k] = s->dep_d + (size_t)k * c.dep_dim;
        }
        MatvecArgs g;
        g.ctrl = s->ctrl; g.eps = 1e-8f; g.w = m->dep_in[weights_of(0)]; g.x = s->tout; g.out = s->dep_d;
        mm.per = L.gemv_ctas(g.w);
        L.gemv_multi(g, mm, c.dep_q, FAM_DEP_IN);
    }
    for (int k = 0; k < c.dep_q; k++) {
        const int w = weights_of(k);
        const float *dk = s->dep_d + (size_t)k * c.dep_dim;
        MatvecArgs g;
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
        MatvecArgs h;
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
    CU(cudaFuncSetAttribute(dq_matvec_kernel<12, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_multi_kernel<12, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_multi_kernel<12, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_multi_kernel<8, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_multi_kernel<8, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_kernel<12, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_kernel<8, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_kernel<8, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_kernel<12, 32, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_kernel<12, 16, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_kernel<8, 32, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_kernel<8, 16, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_kernel<12, 32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_kernel<12, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_kernel<8, 32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_kernel<8, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_local_attn_kernel<12, 32, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_local_attn_kernel<12, 32, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_local_attn_kernel<12, 16, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_local_attn_kernel<12, 16, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_local_attn_kernel<8, 32, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_local_attn_kernel<8, 32, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_local_attn_kernel<8, 16, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU(cudaFuncSetAttribute(dq_matvec_local_attn_kernel<8, 16, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
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
