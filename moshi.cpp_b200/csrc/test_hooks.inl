// test_hooks.inl — unit-level entry points used by tests/ and bench.py only (single GEMV, dequantisers, row quantisers,
// graph-replay GEMV bench).  Included by engine.cu.

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
    MatvecArgs g;
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
        MatvecArgs g;
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
