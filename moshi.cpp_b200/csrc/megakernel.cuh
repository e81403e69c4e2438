// megakernel.cuh — persistent "phase program" kernel (sm_100a).
//
// The reference runs one frame as ~3000 ggml graph nodes (HOT LOOP A / B); the first
// CUDA path here ran it as 420 kernel launches, and ncu showed every launch paying ~5 us of fixed
// latency (launch, dependent prologue loads, tail) — the depformer alone was 257 launches for 58 us
// worth of HBM traffic.  This kernel is launched ONCE per graph (cooperatively, one CTA per SM) and
// walks a device-resident list of phases; phases are separated by a grid-wide barrier instead of a
// kernel boundary.  The arithmetic of every phase is the same device code the standalone kernels use
// (gemv_body, attention), so results are bit-identical to the multi-kernel path.
#pragma once
#include "attention.cuh"
#include "common.cuh"
#include "gemv.cuh"
#include "misc_kernels.cuh"

namespace msx {

enum PhaseType : int {
    PH_GEMV = 0,            // fused dequant-GEMV (any prologue / epilogue)
    PH_GEMV_LOCAL_ATTN = 1, // every CTA first recomputes the (tiny) ring attention of ALL heads into shared
                            // memory, then runs the out_proj GEMV on it (depformer: cap <= 64 slots)
    PH_EMBED = 2,
    PH_FINALIZE_TEMPORAL = 3,
    PH_FINALIZE_DEPFORMER = 4,
};

struct Phase {
    int32_t type = 0, pro = 0, epi = 0, pad_ = 0;
    GemvArgs g;
    AttnArgs a;
    EmbedArgs e;
    int32_t heads = 0, dh = 0, dep_q = 0, has_depformer = 0;
};

struct MegaArgs {
    const Phase *phases = nullptr;
    int32_t n_phases = 0;
    Ctrl *ctrl = nullptr;
    long long *dbg = nullptr;       // optional timeline: 5 globaltimer stamps per phase from CTA 0
};

// ---- geometry ---------------------------------------------------------------------------------------------
constexpr int kMegaThreads = 512;                    // 15 consumer warps + 1 prefetch warp, one CTA per SM
constexpr int kMegaConsumers = kMegaThreads - 32;
constexpr int kMegaConsumerWarps = kMegaConsumers / 32;
constexpr unsigned long long kPrefetchWindow = 160 * 1024;   // bytes of weights kept in flight / in L2 ahead of the consumers, per CTA

// ---- grid barrier ------------------------------------------------------------------------------------
// Monotonic 64-bit arrival counter in the control block; barrier number b of this launch completes when
// counter >= base + b * gridDim.x, where base was read at kernel start and is advanced by CTA 0 after the
// last barrier.  All CTAs are co-resident (cooperative launch), so spinning is safe; a clock64() watchdog
// turns a would-be hang into an error flag.  Only the consumer threads take part (named barrier 1).
// Polling load: relaxed at gpu scope (served by L2).  An acquire here would make the SM invalidate its L1
// (CCTL.IVALL) at every barrier; everything that is written by other CTAs during a launch is read with
// ld.global.cg (L1 bypass) instead, so L1-resident constants / descriptors stay hot across barriers.
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_volatile_s32(const int *p) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long atom_release_add_u64(unsigned long long *p, unsigned long long v) {
    unsigned long long old;
    asm volatile("atom.release.gpu.global.add.u64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(v) : "memory");
    return old;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Arrivals are counted on bar_counter; the last arriver publishes the target on bar_flag, which is the only
// word the waiters poll (read-only line: the polls do not queue behind the arrival atomics).
__device__ __forceinline__ void grid_barrier(Ctrl *c, unsigned long long target, const BlockGeom &bg) {
    block_sync(bg);                                        // all consumer writes of this CTA are issued
    if (threadIdx.x == 0) {
        const unsigned long long old = atom_release_add_u64(&c->bar_counter, 1ull);   // release, cumulative over the CTA
        if (ld_volatile_s32(&c->bar_mode) != 2) {          // default: poll the arrival counter itself (1.2 us vs 1.9 us for the flag variant on B200)
            const long long t0 = clock64();
            while (ld_acquire_u64(&c->bar_counter) < target) { if (clock64() - t0 > 4000000000ll) { c->error = 1; break; } }
        } else if (old + 1 == target) st_release_u64(&c->bar_flag, target);
        else {
            const long long t0 = clock64();
            while (ld_acquire_u64(&c->bar_flag) < target) {
                if (clock64() - t0 > 4000000000ll) { c->error = 1; break; }     // ~2 s: never hang the device
                if (ld_volatile_s32(&c->error)) break;
            }
        }
    }
    block_sync(bg);                                        // activations are re-read with ld.global.cg (L1 bypass)
}

// L2 prefetch of a byte span; the span is widened to 16-byte boundaries (allocations are 256-byte granular)
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, unsigned long long bytes) {
    const unsigned long long a0 = reinterpret_cast<unsigned long long>(p) & ~15ull;
    const unsigned long long a1 = (reinterpret_cast<unsigned long long>(p) + bytes + 15ull) & ~15ull;
    if (a1 > a0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"((unsigned)(a1 - a0)) : "memory");
}

// ---- ring attention of all heads, recomputed by every CTA (depformer) ---------------------------------
// Same arithmetic as attn_kernel (bf16-rounded q and p, products summed in double, exact softmax), one
// warp per head, results to ctx_s[dim] in shared memory.  CTA `writer` also performs the ring insert.
template <int DH>
__device__ __forceinline__ void attn_local(const AttnArgs &a, int heads, float *ctx_s, float *scratch, bool writer, int nwarps) {
    constexpr int LPS = DH / 8;              // lanes per slot
    constexpr int SPI = 32 / LPS;            // slots per warp iteration
    const int lane = threadIdx.x & 31, warp = uniform_warp_id();
    const int cap = a.cap;
    const int pos = a.pos_const >= 0 ? a.pos_const : a.ctrl->offset;
    const int slot = pos % cap;
    const int n_valid = (pos >= cap - 1) ? cap : pos + 1;
    // per-warp scratch: q[DH] f32 | knew[DH] bf16 | vnew[DH] bf16 | scores[64] f32
    float *q_s = scratch + warp * (DH + DH + 64);
    uint16_t *knew = reinterpret_cast<uint16_t *>(q_s + DH);
    uint16_t *vnew = knew + DH;
    float *sc_s = q_s + DH + DH;
    const float scale = 1.f / sqrtf((float)DH);
    for (int h = warp; h < heads; h += nwarps) {
        const float *q = a.qkv + h * DH, *k = a.qkv + a.dim + h * DH, *v = a.qkv + 2 * a.dim + h * DH;
        for (int j = lane; j < DH / 2; j += 32) {
            const float2 qq = __ldcg(reinterpret_cast<const float2 *>(q + 2 * j));
            const float2 kk = __ldcg(reinterpret_cast<const float2 *>(k + 2 * j));
            const float2 vv = __ldcg(reinterpret_cast<const float2 *>(v + 2 * j));
            if (a.max_period) {
                const float arg = (float)pos * a.rope_freq[j];
                const float cs = (float)cos((double)arg), sn = (float)sin((double)arg);
                q_s[j] = bf16_round(__fsub_rn(__fmul_rn(qq.x, cs), __fmul_rn(qq.y, sn)));
                q_s[DH / 2 + j] = bf16_round(__fadd_rn(__fmul_rn(qq.x, sn), __fmul_rn(qq.y, cs)));
                knew[j] = f32_to_bf16_bits(__fsub_rn(__fmul_rn(kk.x, cs), __fmul_rn(kk.y, sn)));
                knew[DH / 2 + j] = f32_to_bf16_bits(__fadd_rn(__fmul_rn(kk.x, sn), __fmul_rn(kk.y, cs)));
            } else {
                q_s[2 * j] = bf16_round(qq.x); q_s[2 * j + 1] = bf16_round(qq.y);
                knew[2 * j] = f32_to_bf16_bits(kk.x); knew[2 * j + 1] = f32_to_bf16_bits(kk.y);
            }
            vnew[2 * j] = f32_to_bf16_bits(vv.x); vnew[2 * j + 1] = f32_to_bf16_bits(vv.y);
        }
        __syncwarp();
        if (writer) {
            const size_t o = ((size_t)h * cap + slot) * DH;
            for (int t = lane; t < DH / 4; t += 32) {
                reinterpret_cast<uint2 *>(a.kc + o)[t] = reinterpret_cast<const uint2 *>(knew)[t];
                reinterpret_cast<uint2 *>(a.vc + o)[t] = reinterpret_cast<const uint2 *>(vnew)[t];
            }
        }
        const int g = lane / LPS, sl = lane % LPS;
        float qv[8];
#pragma unroll
        for (int i = 0; i < 8; i++) qv[i] = q_s[sl * 8 + i];
        float lmax = -INFINITY;
        for (int i0 = 0; i0 < n_valid; i0 += SPI) {
            const int i = i0 + g;
            const bool valid = i < n_valid;
            uint4 kk = make_uint4(0, 0, 0, 0);
            if (valid) {
                if (i == slot) kk = reinterpret_cast<const uint4 *>(knew)[sl];
                else kk = __ldcg(reinterpret_cast<const uint4 *>(a.kc + ((size_t)h * cap + i) * DH + sl * 8));
            }
            double d = 0.0;
            d += (double)(bf16_bits_to_f32(kk.x & 0xffff) * qv[0]); d += (double)(bf16_bits_to_f32(kk.x >> 16) * qv[1]);
            d += (double)(bf16_bits_to_f32(kk.y & 0xffff) * qv[2]); d += (double)(bf16_bits_to_f32(kk.y >> 16) * qv[3]);
            d += (double)(bf16_bits_to_f32(kk.z & 0xffff) * qv[4]); d += (double)(bf16_bits_to_f32(kk.z >> 16) * qv[5]);
            d += (double)(bf16_bits_to_f32(kk.w & 0xffff) * qv[6]); d += (double)(bf16_bits_to_f32(kk.w >> 16) * qv[7]);
#pragma unroll
            for (int o = LPS / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
            const float s = (float)d * scale + 0.0f;
            if (valid) { if (sl == 0) sc_s[i] = s; lmax = fmaxf(lmax, s); }
        }
        lmax = warp_max(lmax);
        __syncwarp();
        double lsum = 0.0;
        for (int i = lane; i < n_valid; i += 32) { const float e = (float)exp((double)(sc_s[i] - lmax)); sc_s[i] = e; lsum += (double)e; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
        const float inv = (float)(1.0 / lsum);
        __syncwarp();
        // context: lane owns dims {2*lane, 2*lane+1} (+64 for DH = 128)
        for (int d0 = 2 * lane; d0 < DH; d0 += 64) {
            double acc0 = 0.0, acc1 = 0.0;
            for (int i = 0; i < n_valid; i++) {
                const float p = bf16_round(sc_s[i] * inv);
                uint32_t vv;
                if (i == slot) vv = *reinterpret_cast<const uint32_t *>(vnew + d0);
                else vv = __ldcg(reinterpret_cast<const uint32_t *>(a.vc + ((size_t)h * cap + i) * DH + d0));
                acc0 += (double)(bf16_bits_to_f32(vv & 0xffff) * p);
                acc1 += (double)(bf16_bits_to_f32(vv >> 16) * p);
            }
            ctx_s[h * DH + d0] = (float)acc0; ctx_s[h * DH + d0 + 1] = (float)acc1;
        }
        __syncwarp();
    }
}

// ---- standalone fused kernel: local ring attention (all heads, every CTA) + out_proj GEMV -----------------
// Used by the PDL-chained path for transformers whose ring is tiny (depformer: <= 64 slots): one launch
// instead of attention + out_proj.  smem: [gemv region][ctx_s dim floats][attn scratch]
template <int WT, int LANES, int DH>
__global__ void __launch_bounds__(kGemvThreads, 1) gemv_local_attn_kernel(const GemvArgs g, const AttnArgs a, const int heads, const int pro,
                                                                          const int epi, const int gemv_region) {
    extern __shared__ __align__(16) uint8_t smem[];
    griddep_launch();
    const BlockGeom bg{kGemvThreads, kGemvThreads / 32};
    float *ctx_s = reinterpret_cast<float *>(smem + gemv_region);
    float *scratch = ctx_s + a.dim;
    // the out_proj weights do not depend on qkv: gemv_body requests its first steps, THEN waits for the previous kernel and
    // runs the attention (the whole 0.6 MB matrix of a depformer layer is in flight under it)
    auto mid = [&]() {
        griddep_wait();                              // qkv comes from the previous kernel
        attn_local<DH>(a, heads, ctx_s, scratch, blockIdx.x == 0, bg.nwarps);
        block_sync(bg);
    };
    gemv_body<WT, LANES, false, false, true, decltype(mid)>(g, ctx_s, false, pro, epi, smem, blockIdx.x, gridDim.x, bg, nullptr, mid);   // lean body: residual epilogue only
}
__host__ __device__ inline int local_attn_smem_bytes(int gemv_bytes, int dim, int dh) {
    return (gemv_bytes + 15) / 16 * 16 + dim * 4 + (kGemvThreads / 32) * (2 * dh + 64) * 4 + 64;
}

// ---- phase dispatch -------------------------------------------------------------------------------------
__device__ __noinline__ void gemv_phase(const GemvArgs &g, const float *x_over, bool norm_out_cta, int pro, int epi, uint8_t *smem,
                                        int cta, int n_cta, const BlockGeom bg, unsigned long long *progress) {
    if (g.w.type == 12) {
        if (g.w.gs == 32) gemv_body<12, 32>(g, x_over, norm_out_cta, pro, epi, smem, cta, n_cta, bg, progress);
        else gemv_body<12, 16>(g, x_over, norm_out_cta, pro, epi, smem, cta, n_cta, bg, progress);
    } else {
        if (g.w.gs == 32) gemv_body<8, 32>(g, x_over, norm_out_cta, pro, epi, smem, cta, n_cta, bg, progress);
        else gemv_body<8, 16>(g, x_over, norm_out_cta, pro, epi, smem, cta, n_cta, bg, progress);
    }
}

// shared memory: [gemv region: max gemv_smem_bytes][ctx_s: max_dim floats][attn_local scratch: consumer warps]
__host__ __device__ inline int mega_smem_bytes(int max_gemv_bytes, int max_local_dim, int max_dh) {
    return (max_gemv_bytes + 15) / 16 * 16 + max_local_dim * 4 + kMegaConsumerWarps * (2 * max_dh + 64) * 4 + 64;
}

// bytes of repacked weights per row (all planes)
__device__ __forceinline__ unsigned long long row_bytes_all(const QLinear &w) {
    return w.type == 12 ? (unsigned long long)(w.K >> 1) + (w.K >> 6) * 4 + (w.K >> 8) * 4
                        : (unsigned long long)w.K + (w.K >> 5) * 2;
}

// The prefetch warp walks the whole program ahead of the consumers: for every GEMV phase it issues L2 bulk
// prefetches for exactly the weight spans this CTA will read (weights do not depend on activations, so it
// runs through grid barriers), throttled to kPrefetchWindow bytes ahead of what the consumers have retired.
__device__ __forceinline__ void prefetch_warp_loop(const MegaArgs &m, int cta, int n_cta, volatile unsigned long long *consumed) {
    const int lane = threadIdx.x & 31;
    unsigned long long issued = 0;
    for (int p = 0; p < m.n_phases; p++) {
        const Phase *ph = m.phases + p;
        const int type = __ldg(&ph->type);
        if (type != PH_GEMV && type != PH_GEMV_LOCAL_ATTN) continue;
        QLinear w;
        {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&ph->g.w);
            uint32_t *dst = reinterpret_cast<uint32_t *>(&w);
#pragma unroll
            for (int i = 0; i < (int)(sizeof(QLinear) / 4); i++) dst[i] = __ldg(src + i);
        }
        if (lane == 0 && __ldg(&ph->pro) == PRO_RMS) {
            const float *alpha = *reinterpret_cast<const float *const *>(&ph->g.alpha);
            prefetch_l2_bulk(alpha, (unsigned long long)w.K * 4);
        }
        const int tr = tile_rows(w.gs);
        const int n_tiles = (w.rows + tr - 1) / tr;
        const int t_begin = (int)((long long)cta * n_tiles / n_cta), t_end = (int)((long long)(cta + 1) * n_tiles / n_cta);
        int r0 = t_begin * tr, r1 = min(t_end * tr, w.rows);
        const unsigned long long rb = row_bytes_all(w);
        // chunks of whole tiles, <= 16 KB of qs each; lanes 0..2 take one plane each
        const int rows_per_chunk = max(4, (int)(16384 / max(1ull, rb)) / 4 * 4);
        for (int r = r0; r < r1; r += rows_per_chunk) {
            const int nr = min(rows_per_chunk, r1 - r);
            const unsigned long long bytes = rb * nr;
            // throttle
            if (lane == 0) {
                const long long t0 = clock64();
                while (issued + bytes > *consumed + kPrefetchWindow) {
                    if (clock64() - t0 > 4000000000ll) break;
                    __nanosleep(64);
                }
            }
            __syncwarp();
            if (w.type == 12) {
                if (lane == 0) prefetch_l2_bulk(w.qs + (size_t)r * (w.K >> 1), (unsigned long long)nr * (w.K >> 1));
                else if (lane == 1) prefetch_l2_bulk(w.sc + (size_t)r * (w.K >> 6), (unsigned long long)nr * (w.K >> 6) * 4);
                else if (lane == 2) prefetch_l2_bulk(reinterpret_cast<const uint32_t *>(w.dd) + (size_t)r * (w.K >> 8), (unsigned long long)nr * (w.K >> 8) * 4);
            } else {
                if (lane == 0) prefetch_l2_bulk(w.qs + (size_t)r * w.K, (unsigned long long)nr * w.K);
                else if (lane == 1) prefetch_l2_bulk(reinterpret_cast<const uint8_t *>(w.dd) + (size_t)r * (w.K >> 5) * 2, (size_t)nr * (w.K >> 5) * 2);
            }
            issued += bytes;
        }
    }
}

__global__ void __launch_bounds__(kMegaThreads, 1) mega_kernel(const MegaArgs m, const int gemv_region, const int local_dim) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ unsigned long long s_base;
    __shared__ unsigned long long s_consumed;
    __shared__ __align__(16) unsigned char s_phase_raw[2][sizeof(Phase)];
    Ctrl *c = m.ctrl;
    const int cta = blockIdx.x, n_cta = gridDim.x;
    if (threadIdx.x == 0) { s_base = c->bar_base; s_consumed = 0ull; }
    {   // descriptor of phase 0
        const int n4 = (int)(sizeof(Phase) / 4);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(m.phases);
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_phase_raw[0]);
        for (int i = threadIdx.x; i < n4; i += kMegaThreads) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    if (threadIdx.x >= kMegaConsumers) {                   // ---- prefetch warp ----
        prefetch_warp_loop(m, cta, n_cta, &s_consumed);
        return;
    }
    // ---- consumer warps ----
    BlockGeom bg{kMegaConsumers, kMegaConsumerWarps};
    const unsigned long long base = s_base;
    float *ctx_s = reinterpret_cast<float *>(smem + gemv_region);
    float *scratch = ctx_s + local_dim;
    constexpr int n4 = (int)(sizeof(Phase) / 4);

    for (int p = 0; p < m.n_phases; p++) {
        const Phase &ph = *reinterpret_cast<const Phase *>(s_phase_raw[p & 1]);
        // fetch the next descriptor while this phase runs (stored to the other buffer before the barrier)
        uint32_t next_word = 0;
        const bool fetch_next = (p + 1 < m.n_phases) && threadIdx.x < n4;
        if (fetch_next) next_word = __ldg(reinterpret_cast<const uint32_t *>(m.phases + p + 1) + threadIdx.x);

        if (m.dbg && cta == 0) { bg.stamp = m.dbg + (size_t)p * 8; if (threadIdx.x == 0) bg.stamp[0] = global_ns(); }
        if (ph.type == PH_GEMV || ph.type == PH_GEMV_LOCAL_ATTN) {
            const float *x_over = nullptr;
            if (ph.type == PH_GEMV_LOCAL_ATTN) {
                if (ph.dh == 64) attn_local<64>(ph.a, ph.heads, ctx_s, scratch, cta == 0, bg.nwarps);
                else attn_local<128>(ph.a, ph.heads, ctx_s, scratch, cta == 0, bg.nwarps);
                block_sync(bg);
                x_over = ctx_s;
            }
            gemv_phase(ph.g, x_over, cta == 0, ph.pro, ph.epi, smem, cta, n_cta, bg, &s_consumed);
        } else if (ph.type == PH_EMBED) {
            embed_body(ph.e, cta, n_cta, bg.nthr);
        } else if (ph.type == PH_FINALIZE_TEMPORAL) {
            if (cta == 0) finalize_temporal_body(c, ph.has_depformer);
        } else if (ph.type == PH_FINALIZE_DEPFORMER) {
            if (cta == 0) finalize_depformer_body(c, ph.dep_q);
        }
        if (fetch_next) reinterpret_cast<uint32_t *>(s_phase_raw[(p + 1) & 1])[threadIdx.x] = next_word;
        if (bg.stamp && threadIdx.x == 0) bg.stamp[3] = global_ns();
        if (p + 1 < m.n_phases) grid_barrier(c, base + (unsigned long long)(p + 1) * n_cta, bg);
        if (bg.stamp && threadIdx.x == 0) bg.stamp[4] = global_ns();
    }
    // all CTAs have arrived at every barrier of this launch before CTA 0 gets here
    if (cta == 0 && threadIdx.x == 0) c->bar_base = base + (unsigned long long)(m.n_phases - 1) * n_cta;
}

}  // namespace msx
