// resident.inl — device-resident frame loops (bench / tests), per-family profiling, the step-kernel timeline, stream timers
// and the KV-row test hook.  Included by engine.cu.

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

// The resident loop with one event between the two graphs of a frame: device time of the temporal stack and of the depformer
// inside the REAL pipelined run (graphs, PDL) — unlike msx_profile_frame, which serialises every launch.
extern "C" int msx_run_resident_split(msx_stream *s, const int32_t *frames, int n_frames, int n_steps, float *temporal_ms, float *depformer_ms) {
    if (!s || !frames || n_frames <= 0 || n_steps <= 0 || n_steps > 4096) return fail(MSX_ERR_ARG, "bad argument");
    const msx_config &c = s->m->cfg;
    CU(cudaSetDevice(s->m->device));
    const int n_in = c.n_q + 1;
    if (int e = check_tokens(s->m, frames, n_frames, INT32_MIN, nullptr)) return e;
    int32_t *d_feed = nullptr;
    CU(cudaMalloc((void **)&d_feed, (size_t)n_frames * n_in * 4));
    CU(cudaMemcpy(d_feed, frames, (size_t)n_frames * n_in * 4, cudaMemcpyHostToDevice));
    if (int e = push_inputs(s, nullptr, INT32_MIN, nullptr)) return e;
    Ctrl hdr;
    memset(&hdr, 0, sizeof(hdr));
    hdr.offset = s->host_offset; hdr.frame = 0; hdr.feed_n = n_frames; hdr.n_in = n_in; hdr.feed = d_feed; hdr.trace = nullptr;
    CU(cudaMemcpyAsync(s->ctrl, &hdr, kCtrlInOffset, cudaMemcpyHostToDevice, s->st));
    CU(cudaStreamSynchronize(s->st));
    std::vector<cudaEvent_t> ev((size_t)2 * n_steps + 1);
    for (cudaEvent_t &e : ev) CU(cudaEventCreate(&e));
    CU(cudaEventRecord(ev[0], s->st));
    for (int i = 0; i < n_steps; i++) {
        CU(cudaGraphLaunch(s->g_temporal, s->st));
        CU(cudaEventRecord(ev[2 * i + 1], s->st));
        if (c.dep_q > 0) CU(cudaGraphLaunch(s->g_depformer, s->st));
        CU(cudaEventRecord(ev[2 * i + 2], s->st));
    }
    CU(cudaStreamSynchronize(s->st));
    s->host_offset += n_steps;
    double t = 0.0, d = 0.0;
    for (int i = 0; i < n_steps; i++) {
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, ev[2 * i], ev[2 * i + 1]); cudaEventElapsedTime(&b, ev[2 * i + 1], ev[2 * i + 2]);
        t += a; d += b;
    }
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    if (temporal_ms) *temporal_ms = (float)t;
    if (depformer_ms) *depformer_ms = (float)d;
    int32_t zero2[2] = {0, 0};
    CU(cudaMemcpy(&s->ctrl->frame, zero2, 8, cudaMemcpyHostToDevice));
    cudaFree(d_feed);
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
