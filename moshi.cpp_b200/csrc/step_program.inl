// step_program.inl — host side of the persistent step kernel (step_kernel.cuh): phase programs of the two stacks of a
// frame and their launches.  Included by engine.cu (needs msx_model / msx_stream / Launcher).
//
// temporal program   (lm.h:659-690, transformer.h:910-1039):  embed | L x (in_proj | attention | out_proj | linear_in |
//                     linear_out) | out_norm + text_linear + arg-max | finalize
// depformer program  (lm.h:446-553):  depformer_in of all steps | dep_q x (embedding add | Ld x (in_proj | attention |
//                     out_proj | linear_in | linear_out) | linears[k] + arg-max) | finalize

namespace {

// Models the step kernel takes: one weight type, every linear in stream layout, self-attention only, plain embeddings,
// one GPU, greedy decoding.  Opt-in (MSX_STREAM_STEP_KERNEL); everything else runs on the PDL-chained launches.
bool step_kernel_eligible(const msx_stream *s) {
    const msx_model *m = s->m; const msx_config &c = m->cfg;
    if (!(s->flags & MSX_STREAM_STEP_KERNEL)) return false;
    if (!m->stream_ok || m->tp_world != 1 || c.cross_attention || c.demux_second_stream || m->dep_small) return false;
    if (m->stream_type != T_Q4_K && m->stream_type != T_Q8_0) return false;
    if (s->temp_text > 0.f || s->temp_audio > 0.f) return false;
    if (c.dep_q > 0 && c.dep_max_period) return false;      // the kernel's RoPE table is the temporal position's
    if (m->num_sms > sk::kMaxCta) return false;
    const int dh = c.dim / c.num_heads;
    auto fits = [&](const QLinear &w, int gran) {
        if (w.K > sk::kMaxK || w.K % 256) return false;
        for (int cta = 0; cta < m->num_sms; cta++)
            if (sk::row_begin(w.rows, gran, m->num_sms, cta + 1) - sk::row_begin(w.rows, gran, m->num_sms, cta) > sk::kMaxRowsCta) return false;
        return m->wstream.count(w.qs) != 0;
    };
    for (const LayerW &l : m->layers)
        if (!fits(l.in_proj[0], 1) || !fits(l.out_proj[0], 1) || !fits(l.lin_in[0], 2) || !fits(l.lin_out[0], 1)) return false;
    if (!fits(m->text_linear, 1)) return false;
    if (c.num_heads > m->num_sms) return false;
    // attention scratch: 32 KB of partial contexts | q | k, v | scores / probabilities of the whole ring
    if (32768 + dh * 8 + (s->cap + 2) * 4 > sk::kAttnScratch) return false;
    if (c.dep_q > 0) {
        if (!fits(m->dep_in_all, 1)) return false;
        for (const LayerW &l : m->dep_layers)
            for (size_t w = 0; w < l.in_proj.size(); w++)
                if (!fits(l.in_proj[w], 1) || !fits(l.out_proj[w], 1) || !fits(l.lin_in[w], 2) || !fits(l.lin_out[w], 1)) return false;
        for (const QLinear &w : m->linears) if (!fits(w, 1)) return false;
        if (c.dep_heads > m->num_sms) return false;
        if (32768 + (c.dep_dim / c.dep_heads) * 8 + (m->dep_cap + 2) * 4 > sk::kAttnScratch) return false;
    }
    return true;
}

int build_step_programs(msx_stream *s) {
    msx_model *m = s->m; const msx_config &c = m->cfg;
    const int n_cta = m->num_sms;
    StepBuffers &b = s->step_buf;
    auto ll = [&](sk::LL **p, size_t n) { return salloc(s, (void **)p, (n + 8) * sizeof(sk::LL)); };
    const int dh = c.dim / c.num_heads;
    int S = 1;                                                 // split factor of the temporal attention: power of two, <= 8, heads * S CTAs
    while (S * 2 <= sk::kMaxSplit && c.num_heads * S * 2 <= n_cta && dh / (S * 2) >= 8) S *= 2;
    if (int e = ll(&b.xA, c.dim)) return e;
    if (int e = ll(&b.xB, c.dim)) return e;
    if (int e = ll(&b.qkv, (size_t)3 * c.dim)) return e;
    if (int e = ll(&b.ctx, c.dim)) return e;
    if (int e = ll(&b.gate, m->hidden)) return e;
    if (int e = ll(&b.tkeys, (size_t)2 * n_cta)) return e;
    if (int e = ll(&b.scores, (size_t)c.num_heads * s->cap)) return e;
    if (int e = salloc(s, (void **)&s->d_epoch, 64)) return e;
    const uint32_t one = 1u;
    CU(cudaMemcpy(s->d_epoch, &one, 4, cudaMemcpyHostToDevice));

    auto stream_of = [&](const QLinear &w) { return m->wstream.at(w.qs); };
    auto gemv = [&](std::vector<sk::StepPhase> &prog, const QLinear &w, int pro, int epi, int fam) -> sk::StepPhase & {
        sk::StepPhase ph;
        ph.type = sk::PH_GEMV; ph.pro = pro; ph.epi = epi; ph.fam = fam;
        const auto sw = stream_of(w);
        ph.w = sw.p; ph.gran = sw.gran; ph.K = w.K; ph.rows = w.rows; ph.eps = 1e-8f;
        prog.push_back(ph);
        return prog.back();
    };
    // one transformer layer; returns with the residual stream back in xa
    auto layer = [&](std::vector<sk::StepPhase> &prog, const LayerW &lw, int w, bool temporal, int li, int step, sk::LL *xa, sk::LL *xb,
                     sk::LL *qkv, sk::LL *ctx, sk::LL *gate, int &x_src) {
        const int heads = temporal ? c.num_heads : c.dep_heads, dim = temporal ? c.dim : c.dep_dim, cap = temporal ? s->cap : m->dep_cap;
        {   sk::StepPhase &g = gemv(prog, lw.in_proj[w], PRO_RMS, EPI_STORE, temporal ? FAM_IN_PROJ : FAM_DEP_IN_PROJ);
            g.x_ll = xa; g.x_src = x_src; g.alpha = lw.norm1; g.out = qkv; }
        const int p_qkv = (int)prog.size() - 1;
        {   sk::StepPhase a;
            a.type = sk::PH_ATTN; a.fam = temporal ? FAM_ATTN : FAM_DEP_ATTN; a.x_ll = qkv; a.x_src = p_qkv; a.out = ctx;
            a.heads = heads; a.dh = dim / heads; a.cap = cap; a.split = temporal ? S : 1;
            a.pos_const = temporal ? -1 : step; a.step = step;
            a.max_period = temporal ? c.max_period : c.dep_max_period;
            const size_t lstride = (size_t)cap * dim;
            a.kc = (temporal ? s->kc : s->dkc) + (size_t)li * lstride;
            a.vc = (temporal ? s->vc : s->dvc) + (size_t)li * lstride;
            a.scores = b.scores;
            prog.push_back(a); }
        const int p_ctx = (int)prog.size() - 1;
        {   sk::StepPhase &g = gemv(prog, lw.out_proj[w], PRO_PLAIN, EPI_RESID, temporal ? FAM_OUT_PROJ : FAM_DEP_OUT_PROJ);
            g.x_ll = ctx; g.x_src = p_ctx; g.resid = xa; g.resid_src = x_src; g.out = xb; }
        const int p_xb = (int)prog.size() - 1;
        {   sk::StepPhase &g = gemv(prog, lw.lin_in[w], PRO_RMS, EPI_GATE, temporal ? FAM_LIN_IN : FAM_DEP_LIN_IN);
            g.x_ll = xb; g.x_src = p_xb; g.alpha = lw.norm2; g.out = gate; }
        const int p_gate = (int)prog.size() - 1;
        {   sk::StepPhase &g = gemv(prog, lw.lin_out[w], PRO_PLAIN, EPI_RESID, temporal ? FAM_LIN_OUT : FAM_DEP_LIN_OUT);
            g.x_ll = gate; g.x_src = p_gate; g.resid = xb; g.resid_src = p_xb; g.out = xa; }
        x_src = (int)prog.size() - 1;
    };

    // ---- temporal ----
    std::vector<sk::StepPhase> pt;
    {   sk::StepPhase e;
        e.type = sk::PH_EMBED; e.fam = FAM_EMBED; e.tables = m->d_emb; e.n_tables = c.n_q + 1; e.dim = c.dim; e.out = b.xA; e.embed_in = s->embed_in;
        pt.push_back(e); }
    int x_src = 0;
    for (int l = 0; l < c.num_layers; l++) layer(pt, m->layers[l], 0, true, l, 0, b.xA, b.xB, b.qkv, b.ctx, b.gate, x_src);
    {   sk::StepPhase &g = gemv(pt, m->text_linear, PRO_RMS, EPI_ARGMAX, FAM_TEXT_HEAD);
        g.x_ll = b.xA; g.x_src = x_src; g.alpha = m->out_norm; g.norm_out = s->tout; g.out_plain = s->text_logits; g.keys = b.tkeys; }
    {   sk::StepPhase f;
        f.type = sk::PH_FINALIZE_T; f.fam = FAM_FINALIZE; f.prev_keys = b.tkeys; f.prev_src = (int)pt.size() - 1; f.has_depformer = c.dep_q > 0 ? 1 : 0;
        pt.push_back(f); }
    if (pt.size() > 4000) return fail(MSX_ERR_ARG, "step program too long");
    if (int e = salloc(s, (void **)&s->d_prog_t, pt.size() * sizeof(sk::StepPhase))) return e;
    CU(cudaMemcpy(s->d_prog_t, pt.data(), pt.size() * sizeof(sk::StepPhase), cudaMemcpyHostToDevice));
    s->n_prog_t = (int)pt.size();
    for (const sk::StepPhase &ph : pt) s->prog_fam_t.push_back(ph.fam);

    // ---- depformer ----
    if (c.dep_q > 0) {
        const int dd = c.dep_dim;
        if (int e = ll(&b.dep_d, (size_t)c.dep_q * dd)) return e;
        if (int e = ll(&b.dxA, dd)) return e;
        if (int e = ll(&b.dxB, dd)) return e;
        if (int e = ll(&b.dqkv, (size_t)3 * dd)) return e;
        if (int e = ll(&b.dctx, dd)) return e;
        if (int e = ll(&b.dgate, m->dep_hidden)) return e;
        if (int e = ll(&b.dkeys, (size_t)c.dep_q * 2 * n_cta)) return e;
        std::vector<sk::StepPhase> pd;
        {   sk::StepPhase &g = gemv(pd, m->dep_in_all, PRO_PLAIN, EPI_STORE, FAM_DEP_IN);
            g.x_plain = s->tout; g.out = b.dep_d; }
        sk::StepPhase fin;
        fin.type = sk::PH_FINALIZE_D; fin.fam = FAM_DEP_FINALIZE; fin.dep_q = c.dep_q; fin.prev_keys = b.dkeys; fin.keys_stride = 2 * n_cta;
        int prev_head = -1;
        for (int k = 0; k < c.dep_q; k++) {
            const int wsel = c.schedule_len ? c.schedule[k] : k, w = m->dep_nw == 1 ? 0 : wsel;     // lm.h:457-462, transformer.h:74-83
            {   sk::StepPhase e;
                e.type = sk::PH_DEP_EMBED; e.fam = FAM_DEP_IN; e.step = k; e.dim = dd; e.x_ll = b.dep_d + (size_t)k * dd; e.x_src = 0;
                e.emb = k == 0 ? m->dep_text_emb : m->dep_emb[k - 1];
                e.prev_keys = k > 0 ? b.dkeys + (size_t)(k - 1) * 2 * n_cta : nullptr; e.prev_src = prev_head;
                e.out = b.dxA;
                pd.push_back(e); }
            int dx_src = (int)pd.size() - 1;
            for (int l = 0; l < c.dep_layers; l++) layer(pd, m->dep_layers[l], w, false, l, k, b.dxA, b.dxB, b.dqkv, b.dctx, b.dgate, dx_src);
            {   sk::StepPhase &g = gemv(pd, m->linears[k], PRO_PLAIN, EPI_ARGMAX, FAM_DEP_HEAD);       // no final norm (lm.h:472)
                g.x_ll = b.dxA; g.x_src = dx_src; g.out_plain = s->audio_logits + (size_t)k * c.card; g.keys = b.dkeys + (size_t)k * 2 * n_cta; }
            prev_head = (int)pd.size() - 1;
            fin.key_src[k] = prev_head;
        }
        pd.push_back(fin);
        if (pd.size() > 4000) return fail(MSX_ERR_ARG, "step program too long");
        if (int e = salloc(s, (void **)&s->d_prog_d, pd.size() * sizeof(sk::StepPhase))) return e;
        CU(cudaMemcpy(s->d_prog_d, pd.data(), pd.size() * sizeof(sk::StepPhase), cudaMemcpyHostToDevice));
        s->n_prog_d = (int)pd.size();
        for (const sk::StepPhase &ph : pd) s->prog_fam_d.push_back(ph.fam);
    }
    return 0;
}

void enqueue_step_kernel(Launcher &L, const msx_stream *s, bool temporal, long long *dbg = nullptr) {
    const msx_model *m = s->m; const msx_config &c = m->cfg;
    sk::StepArgs a;
    a.phases = temporal ? s->d_prog_t : s->d_prog_d;
    a.n_phases = temporal ? s->n_prog_t : s->n_prog_d;
    a.ctrl = s->ctrl; a.epoch = s->d_epoch; a.dbg = dbg;
    if (temporal && c.max_period) { a.rope_dh = c.dim / c.num_heads; a.rope_freq = m->rope_freq; }
    if (!temporal && c.dep_max_period) { a.rope_dh = c.dep_dim / c.dep_heads; a.rope_freq = m->dep_rope_freq; }
    void *args[] = {(void *)&a};
    L.fam = temporal ? FAM_STEP_TEMPORAL : FAM_STEP_DEPFORMER; L.begin();
    const void *fn = m->stream_type == T_Q4_K ? (const void *)sk::step_kernel<12> : (const void *)sk::step_kernel<8>;
    cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(m->num_sms), dim3(sk::kThreads), args, (size_t)sk::kSmemBytes, L.st);
    if (L.err == cudaSuccess) L.err = e;
    L.check();
}

}  // namespace
