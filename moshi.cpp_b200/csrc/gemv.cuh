// gemv.cuh — fused dequant-GEMV for the single-stream decode step (sm_100a).
//
// Replaces, for one activation column, the reference's
//     torch_nn_linear = ggml_mul_mat(W_quant, x)            src/torch.h:79-87
// fused with the ops the reference emits as separate graph nodes around it:
//     moshi_rms_norm (ggml_rms_norm + ggml_mul alpha)        src/moshi/modules/transformer.h:15-23
//     residual ggml_add                                      transformer.h:934, 968
//     silu(left) * right of the gated MLP                    src/moshi/modules/gating.h:18-33
//     ggml_argmax (greedy sampling)                          src/moshi/utils/sampling.h:57-63
//     ggml_add(depformer_in(x), last_token_embedding)        src/moshi/models/lm.h:464-467
//
// Numerics mirror ggml's CPU mul_mat: the activation column is re-quantised (Q8_K per 256 for Q4_K
// weights, Q8_0 per 32 for Q8_0 weights), block dot products are exact integer dp4a sums, block
// scales are formed in fp32.  The exact per-block products scale*isum are accumulated in DOUBLE and
// rounded to fp32 once, so the result does not depend on how rows are tiled over lanes/warps/CTAs
// (ggml's own fp32 summation order is ISA-dependent; see DESIGN.md "Order-independent arithmetic").  HBM-bound: each CTA streams a contiguous range of repacked rows with
// 128-bit loads; the quantised activations live in shared memory (piece-major, conflict-free).
#pragma once
#include "common.cuh"

namespace msx {

// dp4a with unsigned bytes in a, signed bytes in b (no CUDA intrinsic overload for the mixed form)
__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}


// Geometry of the GEMV body comes in at run time; block-level synchronisation uses named barrier 1 over exactly the
// participating threads.
struct BlockGeom { int nthr; int nwarps; };
__device__ __forceinline__ void block_sync(const BlockGeom &bg) { asm volatile("bar.sync 1, %0;" ::"r"(bg.nthr) : "memory"); }

// Activation vectors are produced by other CTAs (previous kernel, or previous phase of the persistent
// kernel): read them through L2 (ld.global.cg) so a stale L1 line can never be observed.  Generic
// addresses that point into shared memory (fused attention output) take the plain path.
__device__ __forceinline__ float4 ld_act4(const float *p) {
    if (__isShared(p)) return *reinterpret_cast<const float4 *>(p);
    return __ldcg(reinterpret_cast<const float4 *>(p));
}

// quantised activations + block-reduce scratch
__host__ __device__ inline int gemv_q_bytes(int type, int K) {
    // Q4_K: x8[K] + bsums int2[K/64] + dx float[K/256];  Q8_0: x8[K] + dx float[K/32]
    int b = (type == 12) ? K + (K / 64) * 8 + (K / 256) * 4 : K + (K / 32) * 4;
    return (b + 15) / 16 * 16 + 128;  // + block-reduce scratch (16 doubles)
}
// Weight ring: every warp owns kStages slots; a slot holds one step of the warp (kR rows x 32 lanes):
// per lane 2 x 16 B of quants per row + 4 B scales + 4 B d/dmin per row, all lane-private.
constexpr int kStages = 4;
constexpr int kR = 2;                                    // rows per lane group per step
constexpr int kSlotBytes = 32 * kR * (32 + 4 + 4);       // 2560 B
__host__ __device__ inline int gemv_smem_bytes_n(int type, int K, int nwarps) { return gemv_q_bytes(type, K) + nwarps * kStages * kSlotBytes; }
__host__ __device__ inline int gemv_smem_bytes(int type, int K) { return gemv_smem_bytes_n(type, K, 16); }

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- prologue: (optional RMSNorm) + activation quantisation into shared memory -------------------
// One pass over x: every warp owns 256-element blocks (lane = 8 consecutive elements) and keeps up to
// kKeep of them in registers between the sum-of-squares reduction and the quantisation, so the activation
// vector is read from L2 exactly once (one dependent round trip instead of two).
// Σx² is accumulated in double like ggml_compute_forward_rms_norm_f32 (ggml_float).

// 8 consecutive activations starting at e0: plain f32 vector, or the sum of `nparts` double partial vectors
// (attention context written by the split-KV phase of the persistent kernel)
template <bool LEAN = false>
__device__ __forceinline__ void load_x8(const MatvecArgs &a, const float *xin, int e0, float (&v)[8]) {
    if (!LEAN && a.xparts) {
        double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = 0; c < a.nparts; c++) {
            const double2 *p = reinterpret_cast<const double2 *>(a.xparts + (size_t)c * a.part_stride + e0);
#pragma unroll
            for (int i = 0; i < 4; i++) { const double2 t = __ldcg(p + i); s[2 * i] += t.x; s[2 * i + 1] += t.y; }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = (float)s[i];
        return;
    }
    const float4 p0 = ld_act4(xin + e0), p1 = ld_act4(xin + e0 + 4);
    v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
}
__device__ __forceinline__ void load_alpha8(const MatvecArgs &a, int e0, float (&al)[8]) {
    const float4 a0 = *reinterpret_cast<const float4 *>(a.alpha + e0), a1 = *reinterpret_cast<const float4 *>(a.alpha + e0 + 4);
    al[0] = a0.x; al[1] = a0.y; al[2] = a0.z; al[3] = a0.w; al[4] = a1.x; al[5] = a1.y; al[6] = a1.z; al[7] = a1.w;
}
__device__ __forceinline__ uint2 pack8(const int (&q)[8]) {
    uint2 pk;
    pk.x = (uint32_t)(q[0] & 0xff) | ((uint32_t)(q[1] & 0xff) << 8) | ((uint32_t)(q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24);
    pk.y = (uint32_t)(q[4] & 0xff) | ((uint32_t)(q[5] & 0xff) << 8) | ((uint32_t)(q[6] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
    return pk;
}

// quantize_row_q8_K for block b (all 32 lanes participate): max-|x| carrier, iscale = -127/max,
// q = min(127, rne(iscale*x)), d = 1/iscale; 32-wide sub-block sums for the dmin term
__device__ __forceinline__ void quantize_block_q8k(int b, int lane, const float (&v)[8], int P, int8_t *x8, int *bs, float *dx) {
    float amax = 0.f, mx = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) { float ax = fabsf(v[i]); if (ax > amax) { amax = ax; mx = v[i]; } }
    // |x| bit patterns are monotonic as unsigned integers: one REDUX instead of a 5-level shuffle butterfly
    const float wmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(amax)));
    const unsigned hit = __ballot_sync(0xffffffffu, amax == wmax);          // carrier = first element attaining the max
    const float carrier = __shfl_sync(0xffffffffu, mx, __ffs(hit) - 1);
    int q[8];
    float d = 0.f;
    if (wmax == 0.f) {
#pragma unroll
        for (int i = 0; i < 8; i++) q[i] = 0;
    } else {
        const float iscale = -127.f / carrier;
#pragma unroll
        for (int i = 0; i < 8; i++) { int t = __float2int_rn(iscale * v[i]); q[i] = t < 127 ? t : 127; }
        d = 1.f / iscale;
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += q[i];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if ((lane & 3) == 0) bs[b * 8 + (lane >> 2)] = s;
    if (lane == 0) dx[b] = d;
    // piece-major store: pair p = 4b + lane/8, piece = (lane%8)/2, 8 bytes at (lane&1)*8
    const int p = 4 * b + (lane >> 3), piece = (lane & 7) >> 1;
    *reinterpret_cast<uint2 *>(x8 + ((size_t)piece * P + p) * 16 + (lane & 1) * 8) = pack8(q);
}

// quantize_row_q8_0 for the 8 blocks of 32 inside 256-block b: d = amax/127, q = roundf(x/d), d kept as fp16
__device__ __forceinline__ void quantize_block_q8_0(int b, int lane, const float (&v)[8], bool act, int P, int8_t *x8, float *dx) {
    float amax = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) amax = fmaxf(amax, fabsf(v[i]));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
    const float d = amax / 127.f;
    const float id = d ? 1.0f / d : 0.0f;
    if (act) {
        int q[8];
#pragma unroll
        for (int i = 0; i < 8; i++) q[i] = (int)roundf(v[i] * id);
        const int blk = b * 8 + (lane >> 2), piece = (lane & 3) >> 1;
        if ((lane & 3) == 0) dx[blk] = __half2float(__float2half_rn(d));
        *reinterpret_cast<uint2 *>(x8 + ((size_t)piece * P + blk) * 16 + (lane & 1) * 8) = pack8(q);
    }
}

// Latency structure (measured on B200: one L2 round trip for activations written by other SMs costs ~0.7 us):
// every warp owns the 256-element blocks b = warp, warp + nwarps, ...; they are processed in chunks of kChunk
// blocks whose loads are ALL issued before the first use, so a chunk costs one round trip.  With RMSNorm,
// if the warp's blocks fit in one chunk (the common case) x and alpha stay in registers between the
// sum-of-squares pass and the quantisation pass: the activation vector is read exactly once.
constexpr int kChunk = 3;

template <int WT, int LAYOUT = 0, bool LEAN = false>
__device__ __forceinline__ void gemv_prologue(const MatvecArgs &a, const float *xin, const bool norm_out_cta, const int PRO, int K, int8_t *x8,
                                              int *bs, float *dx, double *red, const BlockGeom &bg) {
    const int lane = threadIdx.x & 31, warp = uniform_warp_id();
    const int nblk = (K + 255) >> 8;
    const int P = WT == 12 ? (K >> 6) : (K >> 5);
    const int nb_w = warp < nblk ? (nblk - warp + bg.nwarps - 1) / bg.nwarps : 0;    // blocks owned by this warp
    const bool keep = nb_w <= kChunk;
    float v[kChunk][8], al[kChunk][8];
    float scale = 1.f;
    auto e0_of = [&](int i) { return (warp + i * bg.nwarps) * 256 + lane * 8; };
    if (PRO == PRO_RMS) {
        double ss = 0.0;
#pragma unroll 1
        for (int c0 = 0; c0 < nb_w; c0 += kChunk) {
#pragma unroll
            for (int j = 0; j < kChunk; j++) {
                const int e0 = e0_of(c0 + j);
                if (c0 + j < nb_w && e0 < K) { load_x8<LEAN>(a, xin, e0, v[j]); if (keep) load_alpha8(a, e0, al[j]); }
                else {
#pragma unroll
                    for (int i = 0; i < 8; i++) v[j][i] = 0.f;
                }
            }
#pragma unroll
            for (int j = 0; j < kChunk; j++)
#pragma unroll
                for (int i = 0; i < 8; i++) ss += (double)(v[j][i] * v[j][i]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) red[warp] = ss;
        block_sync(bg);
        double tot = 0.0;
#pragma unroll 1
        for (int w = 0; w < bg.nwarps; w++) tot += red[w];
        // sum / K like ggml; for power-of-two K scaling by 2^-log2(K) is the same number
        const float mean = (K & (K - 1)) == 0 ? (float)scalbn(tot, -(31 - __clz(K))) : (float)(tot / K);
        scale = 1.0f / sqrtf(mean + a.eps);
    }
#pragma unroll 1
    for (int c0 = 0; c0 < nb_w; c0 += kChunk) {
        if (!(PRO == PRO_RMS && keep)) {
#pragma unroll
            for (int j = 0; j < kChunk; j++) {
                const int e0 = e0_of(c0 + j);
                if (c0 + j < nb_w && e0 < K) { load_x8<LEAN>(a, xin, e0, v[j]); if (PRO == PRO_RMS) load_alpha8(a, e0, al[j]); }
                else {
#pragma unroll
                    for (int i = 0; i < 8; i++) v[j][i] = 0.f;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kChunk; j++) {
            if (c0 + j < nb_w) {                         // warp-uniform
                const int b = warp + (c0 + j) * bg.nwarps, e0 = e0_of(c0 + j);
                const bool act = e0 < K;
                if (PRO == PRO_RMS && act) {
#pragma unroll
                    for (int i = 0; i < 8; i++) v[j][i] = __fmul_rn(al[j][i], __fmul_rn(v[j][i], scale));   // alpha * (x * scale)
                    if (!LEAN && a.norm_out && norm_out_cta) {
                        *reinterpret_cast<float4 *>(a.norm_out + e0) = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
                        *reinterpret_cast<float4 *>(a.norm_out + e0 + 4) = make_float4(v[j][4], v[j][5], v[j][6], v[j][7]);
                    }
                }
                if (WT == 12) quantize_block_q8k(b, lane, v[j], P, x8, bs, dx);
                else quantize_block_q8_0(b, lane, v[j], act, P, x8, dx);
            }
        }
    }
}

// ---- epilogue ------------------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void gemv_epilogue(const MatvecArgs &a, const int EPI, int r0, const float (&acc)[R], int emb_token,
                                              unsigned long long &best) {
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int row = r0 + r;
        if (row >= a.w.rows) break;
        if (EPI == EPI_STORE) a.out[row] = acc[r];
        else if (EPI == EPI_RESID) a.out[row] = __ldcg(a.out + row) + acc[r];
        else if (EPI == EPI_GATE) {
            if ((r & 1) == 0) { const float g = acc[r]; a.out[row >> 1] = (g / (1.0f + (float)exp((double)(-g)))) * acc[r + 1]; }
        } else if (EPI == EPI_ARGMAX) {
            a.out[row] = acc[r];
            const unsigned long long k = argmax_key(acc[r], row);
            best = k > best ? k : best;
        } else if (EPI == EPI_ADD_EMB) {
            float e = 0.f;
            if (a.emb_step == 0) {   // moshi_scaled_embedding_step: -1 -> zeros, other negatives -> row 0
                e = emb_element(a.emb, emb_token < 0 ? 0 : emb_token, row);
                e = e * (emb_token == -1 ? 0.f : 1.f);
            } else {
                e = emb_element(a.emb, emb_token, row);
            }
            a.out[row] = acc[r] + e;
        } else if (EPI == EPI_ADD_VEC) {
            a.out[row] = acc[r] + __ldcg(a.addvec + row);
        }
    }
}

// token that feeds depformer step `k` (host override > forced > greedy result of the previous step)
__device__ __forceinline__ int depformer_prev_token(const Ctrl *c, int k) {
    // read through L2: the keys are written by other CTAs' atomics earlier in the same (persistent) launch
    if (k == 0) { const int o = __ldcg(&c->text_override); return o != INT32_MIN ? o : __ldcg(&c->out_tokens[0]); }
    const int f = __ldcg(&c->force[k - 1]);
    return f != INT32_MIN ? f : argmax_key_index(__ldcg(&c->audio_key[k - 1]));
}

// ---- the kernel ----------------------------------------------------------------------------------
// WT: 12 = Q4_K, 8 = Q8_0.  LANES: lanes cooperating on one row (32 or 16).
// ---- main loop building blocks ---------------------------------------------------------------------
// A "step" is one (tile, it) pair of a warp: kR rows x LANES pairs (Q4_K: 64 weights, Q8_0: 32 weights per
// lane).  The weights of a step are copied global -> shared with cp.async into a lane-private part of the
// warp's ring slot (no registers, no barriers: each lane later reads back exactly what it copied); kStages-1
// steps are in flight while one is computed, and the first kStages-1 steps are issued BEFORE the PDL wait
// and the activation prologue, so the HBM stream starts at kernel entry.
__host__ __device__ constexpr int tile_rows(int lanes) { return kR * (32 / lanes); }   // rows a warp retires per tile

// lane-private layout inside a slot: [row r][lane]: w0 at (r*32+lane)*16, w1 at 1024*kR/2.. see offsets below
__device__ __forceinline__ uint8_t *slot_w0(uint8_t *slot, int r, int lane) { return slot + (r * 32 + lane) * 16; }
__device__ __forceinline__ uint8_t *slot_w1(uint8_t *slot, int r, int lane) { return slot + 32 * kR * 16 + (r * 32 + lane) * 16; }
__device__ __forceinline__ uint8_t *slot_sc(uint8_t *slot, int r, int lane) { return slot + 32 * kR * 32 + (r * 32 + lane) * 4; }
__device__ __forceinline__ uint8_t *slot_dd(uint8_t *slot, int r, int lane) { return slot + 32 * kR * 36 + (r * 32 + lane) * 4; }

// position of a step inside the warp's work list, advanced incrementally (no div/mod in the loop)
struct StepPos {
    int it;            // K-iteration inside the tile
    int row0;          // first row of the lane group's rows in this tile
};
template <int LANES>
__device__ __forceinline__ void advance(StepPos &sp, int nit, int row_stride) {
    if (++sp.it == nit) { sp.it = 0; sp.row0 += row_stride; }
}

__device__ __forceinline__ void cp_async16_s(unsigned smem_addr, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4_s(unsigned smem_addr, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}

// slot_u32: shared-space address of the slot; lane-private offsets: w0 at (r*32+lane)*16, w1 at +1024,
// sc at 2048 + (r*32+lane)*4, dd at 2304 + (r*32+lane)*4
template <int WT, int LANES>
__device__ __forceinline__ void issue_step(const QLinear &w, unsigned slot_u32, uint8_t *slot, const StepPos &sp, int l, int lane, int P) {
    const int p = sp.it * LANES + l;
    if (p < P) {
        const int gsz = min(LANES, P - sp.it * LANES);
        const size_t row_qs = WT == 12 ? (size_t)(w.K >> 1) : (size_t)w.K;
#pragma unroll
        for (int r = 0; r < kR; r++) {
            const int row = min(sp.row0 + r, w.rows - 1);
            const uint8_t *q = w.qs + (size_t)row * row_qs + (size_t)(sp.it * (LANES * 32) + l * 16);
            const unsigned o16 = slot_u32 + (r * 32 + lane) * 16, o4 = slot_u32 + 32 * kR * 32 + (r * 32 + lane) * 4;
            cp_async16_s(o16, q);
            cp_async16_s(o16 + 32 * kR * 16, q + gsz * 16);
            if (WT == 12) {
                cp_async4_s(o4, w.sc + (size_t)row * P + p);
                cp_async4_s(o4 + 32 * kR * 4, reinterpret_cast<const uint32_t *>(w.dd) + (size_t)row * (w.K >> 8) + (p >> 2));
            } else {
                // fp16 scale: 4-byte aligned copy of the pair containing it
                const uint16_t *d = reinterpret_cast<const uint16_t *>(w.dd) + (size_t)row * P + p;
                cp_async4_s(o4 + 32 * kR * 4, reinterpret_cast<const void *>(reinterpret_cast<uintptr_t>(d) & ~(uintptr_t)3));
                *reinterpret_cast<uint32_t *>(slot_sc(slot, r, lane)) = (uint32_t)((reinterpret_cast<uintptr_t>(d) >> 1) & 1);   // which half
            }
        }
    }
}

// acc[r] += exact block terms of this step (double accumulation, see file header)
template <int WT, int LANES>
__device__ __forceinline__ void compute_step(const uint8_t *slot, int K, int it, int l, int lane, const int8_t *x8, const int *bs,
                                             const float *dx, double (&acc)[kR]) {
    const int P = WT == 12 ? (K >> 6) : (K >> 5);
    const int p = it * LANES + l;
    if (p >= P) return;
    if (WT == 12) {
        const int4 xa0 = *reinterpret_cast<const int4 *>(x8 + ((size_t)0 * P + p) * 16);
        const int4 xa1 = *reinterpret_cast<const int4 *>(x8 + ((size_t)1 * P + p) * 16);
        const int4 xb0 = *reinterpret_cast<const int4 *>(x8 + ((size_t)2 * P + p) * 16);
        const int4 xb1 = *reinterpret_cast<const int4 *>(x8 + ((size_t)3 * P + p) * 16);
        const int2 b2 = *reinterpret_cast<const int2 *>(bs + 2 * p);
        const float dxv = dx[p >> 2];
#pragma unroll
        for (int r = 0; r < kR; r++) {
            const int4 a0 = *reinterpret_cast<const int4 *>(slot_w0(const_cast<uint8_t *>(slot), r, lane));
            const int4 a1 = *reinterpret_cast<const int4 *>(slot_w1(const_cast<uint8_t *>(slot), r, lane));
            const uint32_t scv = *reinterpret_cast<const uint32_t *>(slot_sc(const_cast<uint8_t *>(slot), r, lane));
            const uint32_t ddv = *reinterpret_cast<const uint32_t *>(slot_dd(const_cast<uint8_t *>(slot), r, lane));
            int dl = 0, dh = 0;   // dh accumulates 16 * (hi nibble) products: exact multiple of 16
            dl = __dp4a(a0.x & 0x0F0F0F0F, xa0.x, dl); dh = dp4a_us((unsigned)a0.x & 0xF0F0F0F0u, xb0.x, dh);
            dl = __dp4a(a0.y & 0x0F0F0F0F, xa0.y, dl); dh = dp4a_us((unsigned)a0.y & 0xF0F0F0F0u, xb0.y, dh);
            dl = __dp4a(a0.z & 0x0F0F0F0F, xa0.z, dl); dh = dp4a_us((unsigned)a0.z & 0xF0F0F0F0u, xb0.z, dh);
            dl = __dp4a(a0.w & 0x0F0F0F0F, xa0.w, dl); dh = dp4a_us((unsigned)a0.w & 0xF0F0F0F0u, xb0.w, dh);
            dl = __dp4a(a1.x & 0x0F0F0F0F, xa1.x, dl); dh = dp4a_us((unsigned)a1.x & 0xF0F0F0F0u, xb1.x, dh);
            dl = __dp4a(a1.y & 0x0F0F0F0F, xa1.y, dl); dh = dp4a_us((unsigned)a1.y & 0xF0F0F0F0u, xb1.y, dh);
            dl = __dp4a(a1.z & 0x0F0F0F0F, xa1.z, dl); dh = dp4a_us((unsigned)a1.z & 0xF0F0F0F0u, xb1.z, dh);
            dl = __dp4a(a1.w & 0x0F0F0F0F, xa1.w, dl); dh = dp4a_us((unsigned)a1.w & 0xF0F0F0F0u, xb1.w, dh);
            const int isum = (int)(scv & 0xff) * dl + (int)((scv >> 8) & 0xff) * (dh >> 4);
            const int imin = (int)((scv >> 16) & 0xff) * b2.x + (int)(scv >> 24) * b2.y;
            const float2 dm = __half22float2(*reinterpret_cast<const __half2 *>(&ddv));
            // exact products (24-bit x <24-bit) accumulated in double: order-independent result
            acc[r] = fma((double)(dm.x * dxv), (double)isum, acc[r]);
            acc[r] = fma(-(double)(dm.y * dxv), (double)imin, acc[r]);
        }
    } else {
        const int4 xa = *reinterpret_cast<const int4 *>(x8 + ((size_t)0 * P + p) * 16);
        const int4 xb = *reinterpret_cast<const int4 *>(x8 + ((size_t)1 * P + p) * 16);
        const float dxv = dx[p];
#pragma unroll
        for (int r = 0; r < kR; r++) {
            const int4 a0 = *reinterpret_cast<const int4 *>(slot_w0(const_cast<uint8_t *>(slot), r, lane));
            const int4 a1 = *reinterpret_cast<const int4 *>(slot_w1(const_cast<uint8_t *>(slot), r, lane));
            const uint32_t half_sel = *reinterpret_cast<const uint32_t *>(slot_sc(const_cast<uint8_t *>(slot), r, lane));
            const uint32_t ddv = *reinterpret_cast<const uint32_t *>(slot_dd(const_cast<uint8_t *>(slot), r, lane));
            int sum = 0;
            sum = __dp4a(a0.x, xa.x, sum); sum = __dp4a(a0.y, xa.y, sum); sum = __dp4a(a0.z, xa.z, sum); sum = __dp4a(a0.w, xa.w, sum);
            sum = __dp4a(a1.x, xb.x, sum); sum = __dp4a(a1.y, xb.y, sum); sum = __dp4a(a1.z, xb.z, sum); sum = __dp4a(a1.w, xb.w, sum);
            const float dw = __half2float(__ushort_as_half((unsigned short)((half_sel ? (ddv >> 16) : ddv) & 0xffff)));
            acc[r] = fma((double)(dw * dxv), (double)sum, acc[r]);
        }
    }
}

// PRO / EPI are run-time (warp-uniform) selectors on purpose: one copy of the main loop per (WT, LANES).
// `a` may live in kernel-parameter space (standalone kernels) or in shared memory (phase descriptor of the
// persistent kernel); it is never copied to local memory.  x_over replaces a.x when non-null.
// TPPUSH: the tensor-parallel peer-memory epilogue (EPI_STORE_F64 with a.tp) lives in its own instantiation so that the
// single-GPU kernels carry none of its code (measured: 4 % slower frames when it was compiled into the common kernel).
// LEAN: store / residual / silu-gate epilogues only — the kernel of 4 of every 5 launches of a frame carries no embedding,
// arg-max, fp64-partial or add-vector code (measured: linear_in 13.96 -> 13.17 us).  Kernel SIZE matters as much: one body per
// kernel, because consecutive launches of a frame alternate between variants and a large kernel image thrashes the
// instruction caches (five specialised bodies in one kernel made the frame 10 % slower although each was faster in isolation).
struct NoMid { __device__ __forceinline__ void operator()() const {} };
// `mid` runs after the first weight steps have been requested and before the activation prologue: the PDL wait of the
// standalone kernels, or wait + local attention of the fused kernel (its weights stream in under the attention)
template <int WT, int LANES, bool PDL = false, bool TPPUSH = false, bool LEAN = false, typename Mid = NoMid>
__device__ __forceinline__ void gemv_body(const MatvecArgs &a, const float *x_over, const bool norm_out_cta, const int PRO, const int EPI,
                                          uint8_t *smem, const int cta, const int n_cta, const BlockGeom bg,
                                          Mid mid = Mid()) {
    const float *xin = x_over ? x_over : a.x;
    constexpr int TR = tile_rows(LANES);
    const int K = a.w.K;
    const int lane = threadIdx.x & 31, warp = uniform_warp_id();
    const int sub = lane / LANES, l = lane % LANES;
    const int P = WT == 12 ? (K >> 6) : (K >> 5);
    const int nit = (P + LANES - 1) / LANES;

    const int n_tiles = (a.w.rows + TR - 1) / TR;
    const int t_begin = (int)((long long)cta * n_tiles / n_cta);
    const int t_end = (int)((long long)(cta + 1) * n_tiles / n_cta);
    const int my_tiles = (t_end - t_begin - warp + bg.nwarps - 1) / bg.nwarps;      // tiles t_begin + warp + i * nwarps
    const int n_steps = my_tiles > 0 ? my_tiles * nit : 0;
    const int row_stride = bg.nwarps * TR;                                          // rows between consecutive tiles of this warp
    StepPos cons{0, (t_begin + warp) * TR + sub * kR};                              // step being computed
    StepPos prod = cons;                                                            // step being fetched

    // weights do not depend on the activations: get the first kStages-1 steps in flight before the prologue
    // (and, in the standalone kernels, before waiting for the previous kernel: PDL)
    uint8_t *ring = smem + gemv_q_bytes(WT, K) + (size_t)warp * kStages * kSlotBytes;
    const unsigned ring_u32 = (unsigned)__cvta_generic_to_shared(ring);
#pragma unroll
    for (int i = 0; i < kStages - 1; i++) {
        if (i < n_steps) { issue_step<WT, LANES>(a.w, ring_u32 + i * kSlotBytes, ring + i * kSlotBytes, prod, l, lane, P); advance<LANES>(prod, nit, row_stride); }
        cp_async_commit();
    }
    if (PDL && PRO == PRO_RMS) {
        // the RMSNorm weights do not depend on the predecessor either, and after a frame's 4 GB of weights they come from HBM:
        // pull this warp's blocks into L2 while the predecessor drains, so the loads after the wait are L2 hits
        for (int b = warp; b * 256 < K; b += bg.nwarps) {
            const float *ap = a.alpha + b * 256 + lane * 8;
            if (b * 256 + lane * 8 < K) asm volatile("prefetch.global.L2 [%0];" ::"l"(ap));
        }
    }
    if (PDL) griddep_wait();
    mid();

    int8_t *x8 = reinterpret_cast<int8_t *>(smem);
    int *bs = nullptr; float *dx = nullptr; double *red = nullptr;
    if (WT == 12) {
        bs = reinterpret_cast<int *>(smem + K);
        dx = reinterpret_cast<float *>(smem + K + (K >> 6) * 8);
        red = reinterpret_cast<double *>(smem + gemv_q_bytes(12, K) - 128);
        gemv_prologue<12, 0, LEAN>(a, xin, norm_out_cta, PRO, K, x8, bs, dx, red, bg);
    } else {
        dx = reinterpret_cast<float *>(smem + K);
        red = reinterpret_cast<double *>(smem + gemv_q_bytes(8, K) - 128);
        gemv_prologue<8, 0, LEAN>(a, xin, norm_out_cta, PRO, K, x8, nullptr, dx, red, bg);
    }
    block_sync(bg);

    int emb_token = 0;
    if (!LEAN && EPI == EPI_ADD_EMB) emb_token = depformer_prev_token(a.ctrl, a.emb_step);
    unsigned long long best = 0ull;
    uint32_t tp_epoch = 0, tp_parity = 0;      // tp_epoch + 1 = sequence number of this reduce
    if (TPPUSH) { tp_epoch = tp_seq(a.tp, a.tp_idx) - 1u; tp_parity = (uint32_t)a.tp_idx & 1u; }

    double acc[kR];
#pragma unroll
    for (int r = 0; r < kR; r++) acc[r] = 0.0;
    // residual epilogue: the old x[row] values of a tile are fetched when the tile STARTS, so the L2 round
    // trip (~0.7 us) hides under the tile's dot products instead of stalling the warp at every tile end
    float resid[kR] = {0.f, 0.f};
    auto fetch_resid = [&]() {
        if (EPI == EPI_RESID && l == 0) {
#pragma unroll
            for (int r = 0; r < kR; r++) { const int row = cons.row0 + r; if (row < a.w.rows) resid[r] = __ldcg(a.out + row); }
        }
    };
    // end of a tile: reduce the lane partials, run the epilogue, reset
    auto finish_tile = [&]() {
        float accf[kR];
#pragma unroll
        for (int r = 0; r < kR; r++) {
#pragma unroll
            for (int o = LANES / 2; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
            if (!LEAN && EPI == EPI_STORE_F64 && l == 0 && cons.row0 + r < a.w.rows) {
                if (TPPUSH) {    // partial sum + sequence number straight into every rank's inbox (own rank included)
                    const size_t o = ((size_t)tp_parity * a.tp->world + a.tp->rank) * a.tp->dim + cons.row0 + r;
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(acc[r]);
                    const uint4 pk = make_uint4((uint32_t)bits, tp_epoch + 1u, (uint32_t)(bits >> 32), tp_epoch + 1u);
                    for (int q = 0; q < a.tp->world; q++) a.tp->inbox[q][o] = pk;
                } else a.out_f64[cons.row0 + r] = acc[r];
            }
            accf[r] = (float)acc[r];
            acc[r] = 0.0;
        }
        if (l == 0) {
            if (EPI == EPI_RESID) {                        // residual already in registers
#pragma unroll
                for (int r = 0; r < kR; r++) { const int row = cons.row0 + r; if (row < a.w.rows) a.out[row] = resid[r] + accf[r]; }
            } else if (LEAN) {               // STORE / GATE only (RESID is handled above)
#pragma unroll
                for (int r = 0; r < kR; r++) {
                    const int row = cons.row0 + r;
                    if (row >= a.w.rows) break;
                    if (EPI == EPI_STORE) a.out[row] = accf[r];
                    else if ((r & 1) == 0) { const float g = accf[r]; a.out[row >> 1] = (g / (1.0f + (float)exp((double)(-g)))) * accf[r + 1]; }
                }
            } else if (EPI != EPI_STORE_F64) gemv_epilogue<kR>(a, EPI, cons.row0, accf, emb_token, best);
        }
    };
#pragma unroll 1
    for (int s = 0; s < n_steps; s++) {
        cp_async_wait<kStages - 2>();            // the group of step s has landed (this lane's own copies)
        if (cons.it == 0) fetch_resid();
        compute_step<WT, LANES>(ring + (s & (kStages - 1)) * kSlotBytes, K, cons.it, l, lane, x8, bs, dx, acc);
        // refill the slot consumed one step ago with step s + kStages - 1
        if (s + kStages - 1 < n_steps) {
            const int slot = (s + kStages - 1) & (kStages - 1);
            issue_step<WT, LANES>(a.w, ring_u32 + slot * kSlotBytes, ring + slot * kSlotBytes, prod, l, lane, P);
            advance<LANES>(prod, nit, row_stride);
        }
        cp_async_commit();
        if (cons.it == nit - 1) finish_tile();
        advance<LANES>(cons, nit, row_stride);
    }

    if (!LEAN && EPI == EPI_ARGMAX) {
        // CTA-level max, then one atomic per CTA
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o); best = t > best ? t : best; }
        unsigned long long *sbest = reinterpret_cast<unsigned long long *>(red);   // prologue scratch is free by now
        block_sync(bg);
        if (lane == 0) sbest[warp] = best;
        block_sync(bg);
        if (threadIdx.x == 0) {
            unsigned long long bb = 0;
            for (int w = 0; w < bg.nwarps; w++) bb = sbest[w] > bb ? sbest[w] : bb;
            if (bb) atomicMax(a.key, bb);
        }
    }
}

// 512 threads, one CTA per SM: half as many CTAs repeat the activation prologue, and each prologue has 16
// warps to spread its 256-element blocks over (K = 11264: 3 block iterations per warp instead of 6)
constexpr int kGemvThreads = 512;
template <int WT, int LANES, bool TPPUSH = false, bool LEAN = false>
__global__ void __launch_bounds__(kGemvThreads, 1) dq_matvec_kernel(const MatvecArgs a, const int pro, const int epi) {
    extern __shared__ __align__(16) uint8_t smem[];
    griddep_launch();      // the next kernel may start launching: it prefetches its weights and then waits for us
    gemv_body<WT, LANES, true, TPPUSH, LEAN>(a, nullptr, blockIdx.x == 0, pro, epi, smem, blockIdx.x, gridDim.x, BlockGeom{kGemvThreads, kGemvThreads / 32});
}

// Several GEMVs of one shape over the SAME input in one launch (the depformer_in[k] . t_out of all codebook steps, which
// do not depend on the serial chain): CTA blockIdx.x works on matrix blockIdx.x / per as CTA blockIdx.x % per of `per`.
constexpr int kGemvMultiMax = 40;
struct MatvecMulti {
    const uint8_t *qs[kGemvMultiMax];
    const uint32_t *sc[kGemvMultiMax];
    const void *dd[kGemvMultiMax];
    float *out[kGemvMultiMax];
    int32_t per = 1;
};
template <int WT, int LANES>
__global__ void __launch_bounds__(kGemvThreads, 1) dq_matvec_multi_kernel(const MatvecArgs a, const __grid_constant__ MatvecMulti m, const int pro, const int epi) {
    extern __shared__ __align__(16) uint8_t smem[];
    griddep_launch();
    const int mat = blockIdx.x / m.per, cta = blockIdx.x - mat * m.per;
    MatvecArgs b = a;
    b.w.qs = m.qs[mat]; b.w.sc = m.sc[mat]; b.w.dd = m.dd[mat]; b.out = m.out[mat];
    gemv_body<WT, LANES, true, false, true>(b, nullptr, false, pro, epi, smem, cta, m.per, BlockGeom{kGemvThreads, kGemvThreads / 32});
}

}  // namespace msx
