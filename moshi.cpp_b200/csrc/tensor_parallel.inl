// tensor_parallel.inl — fused GEMV -> all-reduce over peer memory for tensor-parallel streams (BASELINE.json config 4):
// CUDA-IPC export / connect of the per-stream inbox.  No reference counterpart.  Included by engine.cu.

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
