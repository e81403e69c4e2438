// attention.cuh — RoPE + ring-KV insert + single-query attention over the bf16 ring cache (sm_100a).
//
// One launch replaces, per layer, the reference's graph nodes
//     q/k/v views, cont, reshape, permute                       transformer.h:498-549
//     moshi_apply_rope (+ timestep embedding)                   src/moshi/modules/rope.h:8-128
//     moshi_kv_cache_insert_kv (ggml_set_rows f32->bf16)         transformer.h:238-249
//     bias-pattern window copy                                   torch.h:205-223, transformer.h:1273-1277
//     SDPA = mul_mat(K,q) -> soft_max_ext -> cont(transpose V) -> mul_mat   src/torch.h:225-237
// For T = 1 the bias LUT reduces to "slot i visible iff i <= pos or pos >= cap-1" (SURVEY.md §3.4),
// so only the n_valid = min(pos+1, cap) valid slots are read (the reference reads all `cap` slots and
// re-copies V every layer).
//
// Numerics mirror ggml's CPU path with a bf16 cache: q and the normalised probabilities are rounded
// to bf16 before the two contractions, K/V are rounded to bf16 on insert, softmax normalises by the
// full-row sum BEFORE the bf16 rounding — hence the cluster-wide max/sum exchange below instead of
// an online-softmax merge.
//
// Parallelisation: grid (S, H).  The S CTAs of one head form a thread-block cluster and split the
// valid slots; row max, row sum and the partial context vectors are exchanged through distributed
// shared memory.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

namespace msx {
namespace cg = cooperative_groups;

struct AttnArgs {
    const float *qkv = nullptr;     // [3*dim] = q | k | v, each (h d)
    uint16_t *kc = nullptr;         // this layer's K ring [H][cap][DH] bf16
    uint16_t *vc = nullptr;
    float *ctx = nullptr;           // [dim] attention output (h d)
    const Ctrl *ctrl = nullptr;
    int32_t pos_const = -1;         // >= 0: baked position (depformer step k), else ctrl->offset
    int32_t cap = 0;
    int32_t dim = 0;
    int32_t max_period = 0;         // 0 = no RoPE
    int32_t small_ctx = 0;          // n_valid <= small_ctx: rank 0 of the cluster handles the head alone
    const float *rope_freq = nullptr;  // [DH/2] expf(-logf(max_period)*j/half), computed on the host at load
    const float *rope_cs = nullptr;    // optional [DH]: cos | sin of this step's position, written once per frame by the embed kernel
    // batched steps: grid.z = stream; per-stream strides in elements (0 for a single stream)
    int64_t kv_bstride = 0;
    int32_t qkv_bstride = 0, ctx_bstride = 0;
    int32_t skip_insert = 0;        // the ring rows of this step were written by kv_insert_kernel (batched-T prefill)
    // batched-T prefill past the ring's first lap: the pass's columns overwrite the `victim_n` oldest slots before anybody attends;
    // kv_insert_kernel saves the old rows here ([H][kVictimRows][DH] bf16) and column c reads them in place of the rows of the
    // columns after it (which hold positions it must not see) — the T > 1 window of torch.h:170-223 with serial-step semantics
    uint16_t *victim_k = nullptr, *victim_v = nullptr;
    int32_t victim_n = 0;
    long long *dbg = nullptr;       // optional timeline (scripts/attn_probe.cu): [CTA][8] globaltimer ns
};

constexpr int kAttnMaxSplit = 8;
constexpr int kVictimRows = 64;       // columns of a prefill pass
// K / V rows of the valid slots stream through a shared-memory ring of kAttnRing chunks of kAttnChunkRows rows, filled with
// cp.async.bulk (TMA) + mbarrier completion: the rows of a (head, split) are contiguous in the ring cache [H][cap][DH], so a
// chunk is ONE bulk copy.  The K chunks are followed by the V chunks of the same rows; the first kAttnRing chunks are
// requested at kernel entry, BEFORE the PDL wait (ring rows of earlier frames do not depend on the predecessor), and the V
// chunks land while the row maximum / row sum travel across the cluster, so the HBM stream does not stop at the softmax.
// (Round 1 fetched 8 rows per lane group per round trip through registers and reached 28 % of the measured HBM peak with a
// full 3000-slot ring; see profiles/r2_attention.md.)
constexpr int kAttnRing = 6;          // 8 and 10 slots measured: no change (the passes are bound by the F2F conversions, profiles/r2_attention.md)
constexpr int kAttnChunkRows = 32;

// shared-memory layout (bytes): ring[kAttnRing][32][DH] bf16 | full[R], empty[R] mbarriers | x_sum[8] f64 | red[8] f64 |
//                               x_ctx[8][DH] f64 | part[NG][DH] f64 | q[DH] f32 | knew,vnew [2*DH] bf16 | x_max[8] f32 | scores[per] f32
template <int DH>
__host__ __device__ inline int attn_ring_bytes() { return kAttnRing * kAttnChunkRows * DH * 2 + (kAttnRing * 20 + 127) / 128 * 128; }   // + full[R], empty[R], cnt[R]
template <int DH>
__host__ __device__ inline int attn_smem_bytes(int cap, int S) {
    const int per = (cap + S - 1) / S + 1;
    const int ng = kThreads / (DH / 8);
    return attn_ring_bytes<DH>() + (64 + 64 + kAttnMaxSplit * DH * 8 + ng * DH * 8 + DH * 4 + DH * 4 + 32 + per * 4 + 15) / 16 * 16;
}

// q' (bf16-rounded, RoPE'd), and the bf16 K / V rows of this step for head h; threads tid < DH/2 work
template <int DH>
__device__ __forceinline__ void rope_rows(const AttnArgs &a, int h, int pos, int tid, float *q_s, uint16_t *knew, uint16_t *vnew) {
    const float *q = a.qkv + h * DH, *k = a.qkv + a.dim + h * DH, *v = a.qkv + 2 * a.dim + h * DH;
    if (tid < DH / 2) {
        const int j = tid;
        float cs = 1.f, sn = 0.f;
        if (a.max_period) {
            // ggml_timestep_embedding: freq = expf(-logf(max_period) * j / half); arg = pos * freq.
            // cos/sin through double: correctly rounded fp32 irrespective of the libm (order-independent parity)
            if (a.rope_cs) { cs = __ldcg(a.rope_cs + j); sn = __ldcg(a.rope_cs + DH / 2 + j); }
            else {
                const float arg = (float)pos * a.rope_freq[j];
                cs = (float)cos((double)arg); sn = (float)sin((double)arg);
            }
        }
        const float qr = q[2 * j], qi = q[2 * j + 1], kr = k[2 * j], ki = k[2 * j + 1];
        float qo_r, qo_i, ko_r, ko_i;
        if (a.max_period) {
            qo_r = __fsub_rn(__fmul_rn(qr, cs), __fmul_rn(qi, sn)); qo_i = __fadd_rn(__fmul_rn(qr, sn), __fmul_rn(qi, cs));
            ko_r = __fsub_rn(__fmul_rn(kr, cs), __fmul_rn(ki, sn)); ko_i = __fadd_rn(__fmul_rn(kr, sn), __fmul_rn(ki, cs));
            q_s[j] = bf16_round(qo_r); q_s[DH / 2 + j] = bf16_round(qo_i);
            knew[j] = f32_to_bf16_bits(ko_r); knew[DH / 2 + j] = f32_to_bf16_bits(ko_i);
        } else {
            q_s[2 * j] = bf16_round(qr); q_s[2 * j + 1] = bf16_round(qi);
            knew[2 * j] = f32_to_bf16_bits(kr); knew[2 * j + 1] = f32_to_bf16_bits(ki);
        }
        vnew[2 * j] = f32_to_bf16_bits(v[2 * j]); vnew[2 * j + 1] = f32_to_bf16_bits(v[2 * j + 1]);
    }
}

template <int DH, bool CLUSTER>
__global__ void __launch_bounds__(kThreads) attn_kernel(const AttnArgs a0) {
    extern __shared__ __align__(16) uint8_t smem[];
    AttnArgs a = a0;
    {
        const int b = blockIdx.z;
        a.qkv += (size_t)b * a.qkv_bstride; a.ctx += (size_t)b * a.ctx_bstride;
        a.kc += (size_t)b * a.kv_bstride; a.vc += (size_t)b * a.kv_bstride;
        a.ctrl += b; if (a.rope_cs) a.rope_cs += (size_t)b * DH;
    }
    constexpr int LPS = DH / 8;             // lanes per slot (8 dims = 16 B of bf16 each)
    constexpr int NG = kThreads / LPS;      // slots per pass over a chunk
    constexpr int CH = kAttnChunkRows, CHB = CH * DH * 2;
    static_assert(CH % NG == 0 || NG % CH == 0, "chunk / group geometry");
    griddep_launch();
    int S = gridDim.x, c = blockIdx.x;
    const int h = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_id();
    const int cap = a.cap;
    // the position is stable for the whole graph (advanced by the finalize kernel of the previous frame), so it may be read
    // before the PDL wait
    const int pos = a.pos_const >= 0 ? a.pos_const : a.ctrl->offset;
    const int slot = pos % cap;
    const int n_valid = (pos >= cap - 1) ? cap : pos + 1;
    // prefill pass: slot i was overwritten by column jv of this pass (slots of a pass are consecutive modulo cap); columns after
    // mine hold future positions -> read the saved old row instead
    const int vic_n = a.victim_n, my_col = (int)blockIdx.z;
    const int vic_s0 = vic_n ? (((pos - my_col) % cap) + cap) % cap : 0;
    auto victim_of = [&](int i) { const int jv = i - vic_s0 + (i < vic_s0 ? cap : 0); return (jv > my_col && jv < vic_n) ? jv : -1; };
    const int per = (cap + S - 1) / S + 1;
    // short context: the split is not worth three cluster barriers — rank 0 handles the head alone
    // (uniform decision across the cluster, so nobody waits at a barrier)
    const bool use_cluster = CLUSTER && n_valid > min(a.small_ctx, per - 1);
    if (CLUSTER && !use_cluster) { if (c != 0) return; S = 1; c = 0; }
    // "this CTA has started" — arrive now, wait right before the first remote shared-memory write (after the score pass): by then
    // every CTA of the cluster has long arrived, so the start-up handshake costs nothing
    if (use_cluster) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    const int lo = (int)((long long)n_valid * c / S), hi = (int)((long long)n_valid * (c + 1) / S);
    const int n_ck = (hi - lo + CH - 1) / CH, n_chunks = 2 * n_ck;      // K chunks, then V chunks of the same rows

    uint8_t *ring = smem;
    const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t bars = ring_u32 + kAttnRing * CHB;                    // full[R] then empty[R]
    int *cnt = reinterpret_cast<int *>(ring + kAttnRing * CHB + 2 * kAttnRing * 8);   // [R] warps through the chunk in a slot
    uint8_t *rest = smem + attn_ring_bytes<DH>();
    double *x_sum = reinterpret_cast<double *>(rest);                  // [kAttnMaxSplit] cluster exchange: row sums
    double *dred = x_sum + kAttnMaxSplit;                              // [8] block reduce scratch
    float *red = reinterpret_cast<float *>(dred);                      // aliases dred (used at different times)
    double *x_ctx = dred + 8;                                          // [kAttnMaxSplit][DH] cluster exchange: partial contexts
    double *part = x_ctx + kAttnMaxSplit * DH;                         // [NG][DH]
    float *q_s = reinterpret_cast<float *>(part + NG * DH);            // [DH] bf16-rounded q'
    uint16_t *knew = reinterpret_cast<uint16_t *>(q_s + DH);           // [DH]
    uint16_t *vnew = knew + DH;                                        // [DH]
    float *x_max = reinterpret_cast<float *>(vnew + DH);               // [kAttnMaxSplit] cluster exchange: row maxima
    float *sc_s = x_max + kAttnMaxSplit;                               // [per]
    (void)per;

    // chunk j of the stream: rows [lo + (j % n_ck) * CH, ...) of K (j < n_ck) or V
    auto issue = [&](int j) {
        const int jj = j < n_ck ? j : j - n_ck;
        const int r0 = lo + jj * CH, nr = min(CH, hi - r0);
        const uint16_t *src = (j < n_ck ? a.kc : a.vc) + ((size_t)h * cap + r0) * DH;
        const uint32_t s = (uint32_t)(j % kAttnRing);
        mbar_expect_tx(bars + s * 8, (uint32_t)nr * DH * 2);
        bulk_g2s(ring_u32 + s * CHB, src, (uint32_t)nr * DH * 2, bars + s * 8);
    };
    if (tid == 0) {
        for (int s = 0; s < kAttnRing; s++) { mbar_init(bars + s * 8, 1); mbar_init(bars + (kAttnRing + s) * 8, kWarps); cnt[s] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // (batched-T prefill: the rows of this pass's other columns are written by the kv_insert kernel right before us)
        if (!a.skip_insert) for (int j = 0; j < min(kAttnRing, n_chunks); j++) issue(j);
    }
    long long *stamp = a.dbg && tid == 0 ? a.dbg + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr;
    if (stamp) stamp[0] = global_ns();
    __syncthreads();       // barriers initialised before anybody waits on them
    griddep_wait();        // qkv comes from the previous kernel (PDL)
    if (tid == 0 && a.skip_insert) for (int j = 0; j < min(kAttnRing, n_chunks); j++) issue(j);

    // every warp arrives on empty[slot] when it is done with chunk j; the LAST warp through refills the slot with chunk
    // j + kAttnRing (nobody waits for a slower warp: a fixed refiller made warp 0 the pace of the whole CTA)
    auto release = [&](int j) {
        __syncwarp();
        if (lane == 0) {
            const uint32_t s = (uint32_t)(j % kAttnRing);
            mbar_arrive(bars + (kAttnRing + s) * 8);
            if (j + kAttnRing < n_chunks && atomicAdd(&cnt[s], 1) == kWarps - 1) {
                cnt[s] = 0;
                mbar_wait(bars + (kAttnRing + s) * 8, (uint32_t)((j / kAttnRing) & 1));
                issue(j + kAttnRing);
            }
        }
    };

    // ---- 1. RoPE (interleaved pairs -> [re half | im half]) and the new K/V row -------------------
    rope_rows<DH>(a, h, pos, tid, q_s, knew, vnew);
    __syncthreads();
    // ring insert (moshi_kv_cache_insert_kv): DH bf16 = DH/4 8-byte pieces per row
    if (a.skip_insert) {
    } else if (c == 0 && tid < DH / 4) {
        const size_t o = ((size_t)h * cap + slot) * DH;
        reinterpret_cast<uint2 *>(a.kc + o)[tid] = reinterpret_cast<const uint2 *>(knew)[tid];
    } else if (c == 0 && tid >= 64 && tid < 64 + DH / 4) {
        const size_t o = ((size_t)h * cap + slot) * DH;
        reinterpret_cast<uint2 *>(a.vc + o)[tid - 64] = reinterpret_cast<const uint2 *>(vnew)[tid - 64];
    }

    // ---- 2. scores over this CTA's share of the valid slots ----------------------------------------
    // Work item = 16 rows of a chunk, taken by ONE warp: lanes [0,16) own dims [0, DH/2) of row (lane & 15), lanes [16,32) the other
    // half, so a row is reduced inside its lane pair (one shuffle) and every lane runs four independent accumulation chains over
    // DH/2 elements.  Lane r reads its row's 16-byte pieces in the rotated order (u + r) mod (DH/16): rows are DH * 2 bytes apart,
    // i.e. all in the same banks, and the rotation spreads a quarter-warp over eight different pieces (conflict-free for DH = 128).
    // bf16 x bf16 products are exact in fp32; they are summed in double (order-independent).
    const int g = tid / LPS, sl = tid % LPS;       // context pass: row group, 8-dim slice
    const float scale = 1.f / sqrtf((float)DH);
    float lmax = -INFINITY;
    {
        constexpr int HP = DH / 16;              // 16-byte pieces per half row
        const int hl = lane & 15, hs = lane >> 4;
        for (int j = 0; j < n_ck; j++) {
            mbar_wait(bars + (j % kAttnRing) * 8, (uint32_t)((j / kAttnRing) & 1));     // every warp: keeps all warps inside the ring window
            const uint16_t *ck = reinterpret_cast<const uint16_t *>(ring + (j % kAttnRing) * CHB);
#pragma unroll
            for (int half = 0; half < 2; half++) {
                if (((2 * j + half) & (kWarps - 1)) != warp) continue;                  // item 2j + half belongs to warp (2j + half) mod 8
                const int r = half * 16 + hl, i = lo + j * CH + r;
                double d = 0.0;
                if (i < hi) {
                    const uint16_t *rowp = (i == slot && !a.skip_insert) ? knew : ck + r * DH;
                    if (vic_n) { const int jv = victim_of(i); if (jv >= 0) rowp = a.victim_k + ((size_t)h * kVictimRows + jv) * DH; }
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                    for (int u = 0; u < HP; u++) {
                        const int pc = hs * HP + ((u + hl) & (HP - 1));
                        const uint4 kk = *reinterpret_cast<const uint4 *>(rowp + pc * 8);
                        const float4 q0 = *reinterpret_cast<const float4 *>(q_s + pc * 8), q1 = *reinterpret_cast<const float4 *>(q_s + pc * 8 + 4);
                        a0 += (double)(bf16_bits_to_f32(kk.x & 0xffff) * q0.x); a1 += (double)(bf16_bits_to_f32(kk.x >> 16) * q0.y);
                        a2 += (double)(bf16_bits_to_f32(kk.y & 0xffff) * q0.z); a3 += (double)(bf16_bits_to_f32(kk.y >> 16) * q0.w);
                        a0 += (double)(bf16_bits_to_f32(kk.z & 0xffff) * q1.x); a1 += (double)(bf16_bits_to_f32(kk.z >> 16) * q1.y);
                        a2 += (double)(bf16_bits_to_f32(kk.w & 0xffff) * q1.z); a3 += (double)(bf16_bits_to_f32(kk.w >> 16) * q1.w);
                    }
                    d = (a0 + a1) + (a2 + a3);
                }
                d += __shfl_xor_sync(0xffffffffu, d, 16);
                if (i < hi) {
                    const float sv = (float)d * scale + 0.0f;
                    if (hs == 0) sc_s[i - lo] = sv;
                    lmax = fmaxf(lmax, sv);
                }
            }
            release(j);
        }
    }
    if (stamp) stamp[1] = global_ns();      // score pass done (warp 0)
    lmax = warp_max(lmax);
    if (lane == 0) red[warp] = lmax;
    __syncthreads();
    float cmax = red[0];
#pragma unroll
    for (int w = 1; w < kWarps; w++) cmax = fmaxf(cmax, red[w]);
    __syncthreads();

    float gmax = cmax;
    if (use_cluster) {
        cg::cluster_group cl = cg::this_cluster();
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // every CTA of the cluster has started (its shared memory exists)
        if (tid < S) cl.map_shared_rank(x_max, tid)[c] = cmax;     // scatter my max to every CTA of the cluster
        cl.sync();
        gmax = x_max[0];
        for (int r = 1; r < S; r++) gmax = fmaxf(gmax, x_max[r]);
    }

    if (stamp) stamp[2] = global_ns();      // row maximum known
    // ---- 3. exp and row sum (ggml soft_max: expf(x - max), sum in double, scale by 1/sum) -----------
    double lsum = 0.0;
    for (int i = lo + tid; i < hi; i += kThreads) { const float e = (float)exp((double)(sc_s[i - lo] - gmax)); sc_s[i - lo] = e; lsum += (double)e; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if (lane == 0) dred[warp] = lsum;
    __syncthreads();
    double csum = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) csum += dred[w];
    double gsum = csum;
    if (use_cluster) {
        cg::cluster_group cl = cg::this_cluster();
        if (tid < S) cl.map_shared_rank(x_sum, tid)[c] = csum;
        cl.sync();
        gsum = 0.0;
        for (int r = 0; r < S; r++) gsum += x_sum[r];
    }
    const float inv = (float)(1.0 / gsum);
    for (int i = lo + tid; i < hi; i += kThreads) sc_s[i - lo] = bf16_round(sc_s[i - lo] * inv);    // p_i, rounded to bf16 once per row
    __syncthreads();
    if (stamp) stamp[3] = global_ns();      // probabilities known

    // ---- 4. context = sum_i bf16(p_i) * V_i over this CTA's slots ------------------------------------
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = 0.0;
    for (int j = n_ck; j < n_chunks; j++) {
        mbar_wait(bars + (j % kAttnRing) * 8, (uint32_t)((j / kAttnRing) & 1));
        const uint16_t *ck = reinterpret_cast<const uint16_t *>(ring + (j % kAttnRing) * CHB);
#pragma unroll
        for (int r = g; r < CH; r += NG) {
            const int i = lo + (j - n_ck) * CH + r;
            if (i < hi) {
                uint4 vv = (i == slot && !a.skip_insert) ? reinterpret_cast<const uint4 *>(vnew)[sl] : reinterpret_cast<const uint4 *>(ck + r * DH)[sl];
                if (vic_n) { const int jv = victim_of(i); if (jv >= 0) vv = reinterpret_cast<const uint4 *>(a.victim_v + ((size_t)h * kVictimRows + jv) * DH)[sl]; }
                const float p = sc_s[i - lo];
                acc[0] += (double)(bf16_bits_to_f32(vv.x & 0xffff) * p); acc[1] += (double)(bf16_bits_to_f32(vv.x >> 16) * p);
                acc[2] += (double)(bf16_bits_to_f32(vv.y & 0xffff) * p); acc[3] += (double)(bf16_bits_to_f32(vv.y >> 16) * p);
                acc[4] += (double)(bf16_bits_to_f32(vv.z & 0xffff) * p); acc[5] += (double)(bf16_bits_to_f32(vv.z >> 16) * p);
                acc[6] += (double)(bf16_bits_to_f32(vv.w & 0xffff) * p); acc[7] += (double)(bf16_bits_to_f32(vv.w >> 16) * p);
            }
        }
        release(j);
    }
    if (stamp) stamp[4] = global_ns();      // context pass done (warp 0)
#pragma unroll
    for (int i = 0; i < 8; i++) part[g * DH + sl * 8 + i] = acc[i];
    __syncthreads();
    double tot = 0.0;
    if (tid < DH) {
        for (int gg = 0; gg < NG; gg++) tot += part[gg * DH + tid];
    }
    if (use_cluster) {
        cg::cluster_group cl = cg::this_cluster();
        if (tid < DH) cl.map_shared_rank(x_ctx, 0)[c * DH + tid] = tot;    // gather partials in rank 0
        cl.sync();
        if (c == 0 && tid < DH) {
            double t = 0.0;
            for (int r = 0; r < S; r++) t += x_ctx[r * DH + tid];
            a.ctx[h * DH + tid] = (float)t;
        }
    } else {
        if (tid < DH) a.ctx[h * DH + tid] = (float)tot;
    }
    if (stamp) stamp[5] = global_ns();
}

// Batched-T prefill: the K / V rows of ALL columns of a chunk are inserted first (columns = consecutive positions of ONE
// stream sharing its ring), then attn_kernel runs with skip_insert: column b sees the rows of columns < b.
template <int DH>
__global__ void __launch_bounds__(64) kv_insert_kernel(const AttnArgs a0) {
    __shared__ float q_s[DH];
    __shared__ __align__(16) uint16_t knew[DH], vnew[DH];
    AttnArgs a = a0;
    const int b = blockIdx.y, h = blockIdx.x, tid = threadIdx.x;
    griddep_launch();
    griddep_wait();
    a.qkv += (size_t)b * a.qkv_bstride; a.ctrl += b; if (a.rope_cs) a.rope_cs += (size_t)b * DH;
    const int pos = a.pos_const >= 0 ? a.pos_const : a.ctrl->offset;
    rope_rows<DH>(a, h, pos, tid, q_s, knew, vnew);
    __syncthreads();
    const size_t o = ((size_t)h * a.cap + (pos % a.cap)) * DH;
    const size_t vo = ((size_t)h * kVictimRows + b) * DH;                // old row of the slot -> victim row of this column
    if (tid < DH / 4) {
        if (a.victim_k) reinterpret_cast<uint2 *>(a.victim_k + vo)[tid] = reinterpret_cast<const uint2 *>(a.kc + o)[tid];
        reinterpret_cast<uint2 *>(a.kc + o)[tid] = reinterpret_cast<const uint2 *>(knew)[tid];
    } else if (tid >= 32 && tid < 32 + DH / 4) {
        if (a.victim_v) reinterpret_cast<uint2 *>(a.victim_v + vo)[tid - 32] = reinterpret_cast<const uint2 *>(a.vc + o)[tid - 32];
        reinterpret_cast<uint2 *>(a.vc + o)[tid - 32] = reinterpret_cast<const uint2 *>(vnew)[tid - 32];
    }
}

}  // namespace msx
