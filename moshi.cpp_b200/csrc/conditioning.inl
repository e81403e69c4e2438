// conditioning.inl — TTS conditioning: condition_sum / cross-attention memory, voice conditioners, VAD head.
// Reference: moshi.cpp:296-366, 729-760, 851-883; tts.h:5-35; lm.h:973-975.  Included by engine.cu.

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
                MatvecArgs g;
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
    MatvecArgs g;
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
