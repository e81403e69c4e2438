// launcher.inl — kernel families and the Launcher (PDL launches, CUDA-graph capture bookkeeping, per-launch events for
// msx_profile_frame).  Included by engine.cu.

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

    void gemv(const MatvecArgs &a, int pro, int epi, int family = 0) {
        fam = family; begin();
        const int tr = tile_rows(a.w.gs);
        const int n_tiles = (a.w.rows + tr - 1) / tr;
        const int grid = std::max(1, std::min(num_sms, (n_tiles + kTilesPerCta - 1) / kTilesPerCta));
        const int smem = gemv_smem_bytes(a.w.type, a.w.K);
        if (a.tp) {          // tensor-parallel partial sums pushed to the peers: separate instantiations
            if (a.w.type == T_Q4_K) {
                if (a.w.gs == 32) launch_pdl(dq_matvec_kernel<12, 32, true>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
                else launch_pdl(dq_matvec_kernel<12, 16, true>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
            } else {
                if (a.w.gs == 32) launch_pdl(dq_matvec_kernel<8, 32, true>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
                else launch_pdl(dq_matvec_kernel<8, 16, true>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
            }
        } else if ((epi == EPI_STORE || epi == EPI_RESID || epi == EPI_GATE) && !a.xparts && !a.norm_out) {
            // the lean kernels: store / residual / gate epilogues only (4 of every 5 launches of a frame)
            if (a.w.type == T_Q4_K) {
                if (a.w.gs == 32) launch_pdl(dq_matvec_kernel<12, 32, false, true>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
                else launch_pdl(dq_matvec_kernel<12, 16, false, true>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
            } else {
                if (a.w.gs == 32) launch_pdl(dq_matvec_kernel<8, 32, false, true>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
                else launch_pdl(dq_matvec_kernel<8, 16, false, true>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
            }
        } else if (a.w.type == T_Q4_K) {
            if (a.w.gs == 32) launch_pdl(dq_matvec_kernel<12, 32>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
            else launch_pdl(dq_matvec_kernel<12, 16>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
        } else {
            if (a.w.gs == 32) launch_pdl(dq_matvec_kernel<8, 32>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
            else launch_pdl(dq_matvec_kernel<8, 16>, dim3(grid), dim3(kGemvThreads), smem, a, pro, epi);
        }
        check();
    }

    // n GEMVs of one shape over the same input in one launch (lean store epilogue)
    void gemv_multi(const MatvecArgs &a, const MatvecMulti &mm, int n, int family) {
        fam = family; begin();
        const int smem = gemv_smem_bytes(a.w.type, a.w.K);
        const dim3 grid(n * mm.per), block(kGemvThreads);
        if (a.w.type == T_Q4_K) {
            if (a.w.gs == 32) launch_pdl(dq_matvec_multi_kernel<12, 32>, grid, block, smem, a, mm, (int)PRO_PLAIN, (int)EPI_STORE);
            else launch_pdl(dq_matvec_multi_kernel<12, 16>, grid, block, smem, a, mm, (int)PRO_PLAIN, (int)EPI_STORE);
        } else {
            if (a.w.gs == 32) launch_pdl(dq_matvec_multi_kernel<8, 32>, grid, block, smem, a, mm, (int)PRO_PLAIN, (int)EPI_STORE);
            else launch_pdl(dq_matvec_multi_kernel<8, 16>, grid, block, smem, a, mm, (int)PRO_PLAIN, (int)EPI_STORE);
        }
        check();
    }
    int gemv_ctas(const QLinear &w) const {
        const int tr = tile_rows(w.gs);
        return std::max(1, std::min(num_sms, ((w.rows + tr - 1) / tr + kTilesPerCta - 1) / kTilesPerCta));
    }

    // fused local attention + out_proj (tiny rings)
    void gemv_local_attn(const MatvecArgs &g, const AttnArgs &a, int heads, int dh, int pro, int epi, int family = 0) {
        fam = family; begin();
        const int tr = tile_rows(g.w.gs);
        const int n_tiles = (g.w.rows + tr - 1) / tr;
        const int grid = std::max(1, std::min(num_sms, (n_tiles + kTilesPerCta - 1) / kTilesPerCta));
        const int region = (gemv_smem_bytes(g.w.type, g.w.K) + 15) / 16 * 16;
        const int smem = local_attn_smem_bytes(region, a.dim, dh);
#define MSX_LA(WT, LN, DH) launch_pdl(dq_matvec_local_attn_kernel<WT, LN, DH>, dim3(grid), dim3(kGemvThreads), smem, g, a, heads, pro, epi, region)
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
