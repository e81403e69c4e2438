// mimi_rvq.inl — C ABI of the Mimi split residual vector quantiser (mimi_rvq.cuh).  Included by engine.cu.
struct msx_rvq {
    int device = 0, n_sem = 0, n_rest = 0, bins = 0, D = 0, dim = 0;
    float *cb[2] = {nullptr, nullptr}, *cb_t[2] = {nullptr, nullptr};          // first / rest: [n][bins][D] and transposed [n][D][bins]
    uint16_t *in_w[2] = {nullptr, nullptr}, *out_w[2] = {nullptr, nullptr};    // F16 [D][dim] / [dim][D]
    // work buffers, grown on demand
    int cap_T = 0;
    float *x = nullptr, *p = nullptr, *y = nullptr, *y2 = nullptr;
    int32_t *codes = nullptr;
    unsigned long long *keys = nullptr;
    cudaStream_t st = nullptr;
    ~msx_rvq() {
        for (int i = 0; i < 2; i++) { cudaFree(cb[i]); cudaFree(cb_t[i]); cudaFree(in_w[i]); cudaFree(out_w[i]); }
        cudaFree(x); cudaFree(p); cudaFree(y); cudaFree(y2); cudaFree(codes); cudaFree(keys);
        if (st) cudaStreamDestroy(st);
    }
};
namespace {
int rvq_reserve(msx_rvq *q, int T) {
    if (T <= q->cap_T) return 0;
    cudaFree(q->x); cudaFree(q->p); cudaFree(q->y); cudaFree(q->y2); cudaFree(q->codes); cudaFree(q->keys);
    q->x = q->p = q->y = q->y2 = nullptr; q->codes = nullptr; q->keys = nullptr; q->cap_T = 0;
    const int n_q = q->n_sem + q->n_rest;
    CU(cudaMalloc((void **)&q->x, (size_t)T * q->dim * 4));
    CU(cudaMalloc((void **)&q->p, (size_t)T * q->D * 4));
    CU(cudaMalloc((void **)&q->y, (size_t)T * q->dim * 4));
    CU(cudaMalloc((void **)&q->y2, (size_t)T * q->dim * 4));
    CU(cudaMalloc((void **)&q->codes, (size_t)T * n_q * 4));
    CU(cudaMalloc((void **)&q->keys, (size_t)T * 8));
    CU(cudaMemset(q->keys, 0, (size_t)T * 8));
    q->cap_T = T;
    return 0;
}
// y[T][n_out] = conv1d (kernel 1) of x[T][n_in] with the F16 weight w [n_out][n_in]
void rvq_conv(msx_rvq *q, const uint16_t *w, int n_in, int n_out, const float *x, int T, float *y) {
    CondLinearArgs a;
    a.w.data = reinterpret_cast<const uint8_t *>(w); a.w.type = 1; a.w.ne0 = n_in; a.w.ne1 = n_out;
    a.x = x; a.xstride = 1; a.xcol = n_in; a.y = y;
    cond_linear_kernel<<<dim3((n_out + 7) / 8, T), 256, 0, q->st>>>(a);
}
}  // namespace

extern "C" int msx_rvq_create(int device, int n_sem, int n_rest, int bins, int D, int dim, const float *cb_first, const float *cb_rest,
                              const uint16_t *in_first, const uint16_t *in_rest, const uint16_t *out_first, const uint16_t *out_rest, msx_rvq **out) {
    if (!out) return fail(MSX_ERR_ARG, "null argument");
    *out = nullptr;
    if (n_sem < 1 || n_rest < 0 || bins < 1 || D < 1 || dim < 1 || D > 4096 || !cb_first || !in_first || !out_first || (n_rest > 0 && (!cb_rest || !in_rest || !out_rest)))
        return fail(MSX_ERR_ARG, "bad quantiser shape");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(MSX_ERR_CUDA, "no such CUDA device");
    CU(cudaSetDevice(device));
    std::unique_ptr<msx_rvq> q(new msx_rvq);
    q->device = device; q->n_sem = n_sem; q->n_rest = n_rest; q->bins = bins; q->D = D; q->dim = dim;
    CU(cudaStreamCreateWithFlags(&q->st, cudaStreamNonBlocking));
    const float *cbs[2] = {cb_first, cb_rest};
    const uint16_t *ins[2] = {in_first, in_rest}, *outs[2] = {out_first, out_rest};
    const int ns[2] = {n_sem, n_rest};
    for (int i = 0; i < 2; i++) {
        if (ns[i] == 0) continue;
        const size_t n = (size_t)ns[i] * bins * D;
        CU(cudaMalloc((void **)&q->cb[i], n * 4)); CU(cudaMalloc((void **)&q->cb_t[i], n * 4));
        CU(cudaMemcpy(q->cb[i], cbs[i], n * 4, cudaMemcpyHostToDevice));
        rvq::transpose_kernel<<<(unsigned)((n + 255) / 256), 256, 0, q->st>>>(q->cb[i], q->cb_t[i], bins, D, (long long)n);
        CU(cudaMalloc((void **)&q->in_w[i], (size_t)D * dim * 2)); CU(cudaMalloc((void **)&q->out_w[i], (size_t)D * dim * 2));
        CU(cudaMemcpy(q->in_w[i], ins[i], (size_t)D * dim * 2, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(q->out_w[i], outs[i], (size_t)D * dim * 2, cudaMemcpyHostToDevice));
    }
    CU(cudaStreamSynchronize(q->st));
    *out = q.release();
    return 0;
}
extern "C" void msx_rvq_free(msx_rvq *q) { delete q; }

// mimi_quantizer_encode: x [T][dim] host -> codes [n_q][T] host
extern "C" int msx_rvq_encode(msx_rvq *q, const float *x, int T, int n_q, int32_t *codes) {
    if (!q || !x || !codes || T < 1 || T > 65535 || n_q < 1 || n_q > q->n_sem + q->n_rest) return fail(MSX_ERR_ARG, "bad argument (1 <= T <= 65535 frames per call)");
    CU(cudaSetDevice(q->device));
    if (int e = rvq_reserve(q, T)) return e;
    CU(cudaMemcpyAsync(q->x, x, (size_t)T * q->dim * 4, cudaMemcpyHostToDevice, q->st));
    for (int part = 0; part < 2; part++) {
        const int l0 = part == 0 ? 0 : q->n_sem, l1 = std::min(n_q, part == 0 ? q->n_sem : q->n_sem + q->n_rest);
        if (l1 <= l0) continue;
        rvq_conv(q, q->in_w[part], q->dim, q->D, q->x, T, q->p);
        for (int l = l0; l < l1; l++) {
            const size_t off = (size_t)(l - l0) * q->bins * q->D;
            rvq::nearest_kernel<<<dim3((q->bins + rvq::kNearestThreads - 1) / rvq::kNearestThreads, T), rvq::kNearestThreads, (size_t)q->D * 4, q->st>>>(
                q->cb_t[part] + off, q->bins, q->D, q->p, q->keys);
            rvq::apply_kernel<<<T, 256, 0, q->st>>>(q->cb[part] + off, q->D, q->p, q->keys, q->codes + (size_t)l * T);
        }
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(codes, q->codes, (size_t)n_q * T * 4, cudaMemcpyDeviceToHost, q->st));
    CU(cudaStreamSynchronize(q->st));
    return 0;
}
// mimi_decode_latent: codes [K][T] host -> latent [T][dim] host
extern "C" int msx_rvq_decode(msx_rvq *q, const int32_t *codes, int K, int T, float *y) {
    if (!q || !codes || !y || T < 1 || T > 65535 || K < 1 || K > q->n_sem + q->n_rest) return fail(MSX_ERR_ARG, "bad argument (1 <= T <= 65535 frames per call)");
    for (size_t i = 0; i < (size_t)K * T; i++)
        if (codes[i] < 0 || codes[i] >= q->bins) return fail(MSX_ERR_ARG, "code out of range");
    CU(cudaSetDevice(q->device));
    if (int e = rvq_reserve(q, T)) return e;
    CU(cudaMemcpyAsync(q->codes, codes, (size_t)K * T * 4, cudaMemcpyHostToDevice, q->st));
    const int k1 = std::min(K, q->n_sem);
    const unsigned blocks = (unsigned)(((size_t)T * q->D + 255) / 256);
    rvq::decode_kernel<<<blocks, 256, 0, q->st>>>(q->cb[0], k1, q->bins, q->D, q->codes, T, q->p);
    rvq_conv(q, q->out_w[0], q->D, q->dim, q->p, T, q->y);
    if (K > q->n_sem) {
        rvq::decode_kernel<<<blocks, 256, 0, q->st>>>(q->cb[1], K - q->n_sem, q->bins, q->D, q->codes + (size_t)q->n_sem * T, T, q->p);
        rvq_conv(q, q->out_w[1], q->D, q->dim, q->p, T, q->y2);
        const int n = T * q->dim;
        cond_add_kernel<<<(n + 255) / 256, 256, 0, q->st>>>(q->y, q->y2, q->y, n);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(y, q->y, (size_t)T * q->dim * 4, cudaMemcpyDeviceToHost, q->st));
    CU(cudaStreamSynchronize(q->st));
    return 0;
}
// device-timed repeat of encode / decode on resident buffers (bench hook): ms per call
extern "C" int msx_rvq_bench(msx_rvq *q, int T, int n_q, int reps, float *encode_ms, float *decode_ms) {
    if (!q || T < 1 || T > 65535 || n_q < 1 || n_q > q->n_sem + q->n_rest || reps < 1) return fail(MSX_ERR_ARG, "bad argument");
    CU(cudaSetDevice(q->device));
    if (int e = rvq_reserve(q, T)) return e;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    CU(cudaMemsetAsync(q->x, 0, (size_t)T * q->dim * 4, q->st));
    for (int phase = 0; phase < 2; phase++) {
        CU(cudaEventRecord(e0, q->st));
        for (int r = 0; r < reps; r++) {
            if (phase == 0) {
                for (int part = 0; part < 2; part++) {
                    const int l0 = part == 0 ? 0 : q->n_sem, l1 = std::min(n_q, part == 0 ? q->n_sem : q->n_sem + q->n_rest);
                    if (l1 <= l0) continue;
                    rvq_conv(q, q->in_w[part], q->dim, q->D, q->x, T, q->p);
                    for (int l = l0; l < l1; l++) {
                        const size_t off = (size_t)(l - l0) * q->bins * q->D;
                        rvq::nearest_kernel<<<dim3((q->bins + rvq::kNearestThreads - 1) / rvq::kNearestThreads, T), rvq::kNearestThreads, (size_t)q->D * 4, q->st>>>(
                            q->cb_t[part] + off, q->bins, q->D, q->p, q->keys);
                        rvq::apply_kernel<<<T, 256, 0, q->st>>>(q->cb[part] + off, q->D, q->p, q->keys, q->codes + (size_t)l * T);
                    }
                }
            } else {
                const int k1 = std::min(n_q, q->n_sem);
                const unsigned blocks = (unsigned)(((size_t)T * q->D + 255) / 256);
                rvq::decode_kernel<<<blocks, 256, 0, q->st>>>(q->cb[0], k1, q->bins, q->D, q->codes, T, q->p);
                rvq_conv(q, q->out_w[0], q->D, q->dim, q->p, T, q->y);
                if (n_q > q->n_sem) {
                    rvq::decode_kernel<<<blocks, 256, 0, q->st>>>(q->cb[1], n_q - q->n_sem, q->bins, q->D, q->codes + (size_t)q->n_sem * T, T, q->p);
                    rvq_conv(q, q->out_w[1], q->D, q->dim, q->p, T, q->y2);
                    const int n = T * q->dim;
                    cond_add_kernel<<<(n + 255) / 256, 256, 0, q->st>>>(q->y, q->y2, q->y, n);
                }
            }
        }
        CU(cudaEventRecord(e1, q->st));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (phase == 0 && encode_ms) *encode_ms = ms / reps;
        if (phase == 1 && decode_ms) *decode_ms = ms / reps;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CU(cudaGetLastError());
    return 0;
}
