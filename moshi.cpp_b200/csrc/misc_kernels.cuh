// misc_kernels.cuh — embedding sum, frame bookkeeping, load-time repack and test-only dequant kernels.
#pragma once
#include "common.cuh"
#include "gemv.cuh"

namespace msx {

// ---- token embedding sum (lm.h:555-584 + lm_utils.h:157-182) --------------------------------------
// x = E_text(tok[0]) + E_0(tok[1]) + ... + E_{n_q-1}(tok[n_q]), added left to right like the graph.
// token == -1 -> zeros (scale 0), any other negative -> row 0 (scale 1).
struct EmbedArgs {
    const EmbTable *tables = nullptr;   // device array [n_q + 1]
    int32_t n_tables = 0;
    int32_t dim = 0;
    const Ctrl *ctrl = nullptr;
    float *x = nullptr;
    // RoPE table of this step (cos | sin, [dh]) computed once per frame instead of once per attention CTA
    float *rope_cs = nullptr;
    const float *rope_freq = nullptr;
    int32_t dh = 0;
    // TTS family: demuxed text embedding = pre1 + pre2 * right_scale(token) (out1 / out2 projections computed by
    // earlier launches, lm_utils.h:42-66) instead of table 0; cond_sum is added last (lm.h:575-577)
    const float *text_pre1 = nullptr, *text_pre2 = nullptr;
    int32_t num_embeddings = 0;
    const float *cond_sum = nullptr;
    const float *embed_in = nullptr;    // [dim]: used instead of the embedding sum when ctrl->embed_override (lm.h:694-709)
};

// `b` = stream of a batch (batched steps launch grid.y = streams; per-stream buffers are b * dim / b * dh apart)
__device__ __forceinline__ void embed_body(const EmbedArgs &a0, int cta, int n_cta, int nthr, int b = 0) {
    EmbedArgs a = a0;
    a.ctrl += b; a.x += (size_t)b * a.dim; if (a.rope_cs) a.rope_cs += (size_t)b * a.dh;
    const Ctrl *c = a.ctrl;
    const int32_t *toks = c->feed_n ? c->feed + (size_t)(c->frame % c->feed_n) * c->n_in : c->tokens;
    if (cta == 0 && a.rope_cs && (int)threadIdx.x < a.dh / 2) {
        // ggml_timestep_embedding on the f32 position; cos/sin through double (see attention.cuh)
        const float arg = (float)c->offset * a.rope_freq[threadIdx.x];
        a.rope_cs[threadIdx.x] = (float)cos((double)arg);
        a.rope_cs[a.dh / 2 + threadIdx.x] = (float)sin((double)arg);
    }
    if (a.embed_in && c->embed_override) {
        for (int i = cta * nthr + threadIdx.x; i < a.dim; i += n_cta * nthr) a.x[i] = __ldcg(a.embed_in + i);
        return;
    }
    for (int i = cta * nthr + threadIdx.x; i < a.dim; i += n_cta * nthr) {
        float acc = 0.f;
        for (int t = 0; t < a.n_tables; t++) {
            const int tok = toks[t];
            float e;
            if (t == 0 && a.text_pre1) {
                int tk = tok < 0 ? 0 : tok;
                const float rs = (tk / a.num_embeddings - 1) < 0 ? 0.f : 1.f;
                e = __fadd_rn(__ldcg(a.text_pre1 + i), __fmul_rn(__ldcg(a.text_pre2 + i), rs));
            } else {
                e = emb_element(a.tables[t], tok < 0 ? 0 : tok, i);
                e = e * (tok == -1 ? 0.f : 1.f);
            }
            acc = (t == 0) ? e : acc + e;
        }
        if (a.cond_sum) acc = a.cond_sum[i] + acc;
        a.x[i] = acc;
    }
}
__global__ void __launch_bounds__(kThreads) embed_kernel(const EmbedArgs a) { griddep_launch(); griddep_wait(); embed_body(a, blockIdx.x, gridDim.x, kThreads, blockIdx.y); }

// ---- frame bookkeeping ------------------------------------------------------------------------------
// end of the temporal graph: greedy text token out of the arg-max key, position advances
// (states->offset += T, transformer.h:1269-1270)
__device__ __forceinline__ void finalize_temporal_body(Ctrl *c, int has_depformer) {
    if (threadIdx.x == 0) {
        c->out_tokens[0] = argmax_key_index(c->text_key);
        c->text_key = 0ull;
        c->offset += 1;
        if (!has_depformer) {
            if (c->feed_n) { if (c->trace) c->trace[c->frame] = c->out_tokens[0]; c->frame += 1; }
        }
    }
}
__global__ void finalize_temporal_kernel(Ctrl *c, int has_depformer, uint32_t *tp_frame_ctr = nullptr) {
    griddep_launch(); griddep_wait();
    finalize_temporal_body(c + blockIdx.x, has_depformer);
    if (tp_frame_ctr && blockIdx.x == 0 && threadIdx.x == 0) *tp_frame_ctr += 1u;     // tensor parallel: next frame's sequence numbers
}

// end of the depformer graph: collect the dep_q greedy tokens (lm.h:548-552)
// single warp does the whole job (dep_q <= 40: two passes of 32 lanes)
__device__ __forceinline__ void finalize_depformer_body(Ctrl *c, int dep_q) {
    if (threadIdx.x >= 32) return;
    for (int k = threadIdx.x; k < dep_q; k += 32) {
        c->out_tokens[1 + k] = argmax_key_index(c->audio_key[k]);
        c->audio_key[k] = 0ull;
    }
    __syncwarp();
    if (c->feed_n) {
        if (c->trace) for (int k = threadIdx.x; k <= dep_q; k += 32) c->trace[(size_t)c->frame * (dep_q + 1) + k] = c->out_tokens[k];
        __syncwarp();
        if (threadIdx.x == 0) c->frame += 1;
    }
}
__global__ void finalize_depformer_kernel(Ctrl *c, int dep_q) { griddep_launch(); griddep_wait(); finalize_depformer_body(c + blockIdx.x, dep_q); }

// ---- tensor parallelism: residual += all-reduced double partial sums, rounded once (== the single-GPU fp64 accumulate) ---
__global__ void tp_apply_kernel(float *x, const double *partial, int n) {
    griddep_launch();
    griddep_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = __ldcg(x + i) + (float)__ldcg(partial + i);
}

// peer-memory variant: poll the inbox entries (data and sequence number arrive together), add them in rank order
// (identical on all ranks).  The sequence number depends only on the frame counter (stable for the whole temporal graph) and
// on the launch's own reduce index, so the polling — which does not touch x — runs BEFORE the PDL wait and overlaps the tail
// of the local GEMV; only the read-modify-write of x waits for the predecessor.
__device__ __forceinline__ uint4 ld_volatile_v4(const uint4 *p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
constexpr int kTpApplyPer = 4;     // elements per thread (dim <= 4096 with 1024 threads)
__global__ void __launch_bounds__(1024) tp_apply_p2p_kernel(float *x, const TpCtx *tp, int idx) {
    griddep_launch();
    const int world = tp->world, rank = tp->rank, dim = tp->dim;
    const uint32_t seq = tp_seq(tp, idx), parity = (uint32_t)idx & 1u;
    const uint4 *in = tp->inbox[rank] + (size_t)parity * world * dim;
    double s[kTpApplyPer];
    long long t0 = 0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
#pragma unroll
    for (int j = 0; j < kTpApplyPer; j++) {
        s[j] = 0.0;
        const int i = threadIdx.x + j * blockDim.x;
        if (i >= dim) continue;
        for (int r = 0; r < world; r++) {
            uint4 v = ld_volatile_v4(in + (size_t)r * dim + i);
            while (v.y != seq || v.w != seq) {
                long long t1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
                if (t1 - t0 > 20000000000ll) { *tp->error = 2; break; }     // watchdog: a peer never arrived (20 s)
                v = ld_volatile_v4(in + (size_t)r * dim + i);
            }
            s[j] += __longlong_as_double((long long)(((unsigned long long)v.z << 32) | v.x));
        }
    }
    griddep_wait();                  // x may still be read by the kernel before our predecessor
#pragma unroll
    for (int j = 0; j < kTpApplyPer; j++) {
        const int i = threadIdx.x + j * blockDim.x;
        if (i < dim) x[i] = __ldcg(x + i) + (float)s[j];
    }
}

// ---- depformer step input: depformer_in[k] . t_out (hoisted, `d`) + embedding of the previous token (lm.h:464-467, 494-516;
// the EPI_ADD_EMB epilogue as a kernel of its own: -1 -> zeros and other negatives -> row 0 for the text table)
__global__ void __launch_bounds__(256) dep_embed_add_kernel(const Ctrl *ctrl, const float *d, const EmbTable emb, const int step, float *out, const int n) {
    griddep_launch();
    griddep_wait();
    const int token = depformer_prev_token(ctrl, step);
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n; row += gridDim.x * blockDim.x) {
        float e;
        if (step == 0) {
            e = emb_element(emb, token < 0 ? 0 : token, row);
            e = e * (token == -1 ? 0.f : 1.f);
        } else e = emb_element(emb, token, row);
        out[row] = __ldcg(d + row) + e;
    }
}

// ---- quantise-on-load: f32 / f16 / bf16 rows -> GGUF Q8_0 blocks (ggml quantize_row_q8_0_ref: d = amax / 127, id = 1 / d,
// q = roundf(x * id), d stored as fp16).  One warp per 32-element block; the blocks then go through the normal repack.
// Reference: src/loader.h:149-233 (quantise while loading), ggml_cast in transformer.h:807.
__global__ void quantize_rows_q8_0_kernel(const uint8_t *src, int src_type, long long n_blocks, uint8_t *dst) {
    const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= n_blocks) return;
    const long long e = b * 32 + lane;
    float x;
    if (src_type == 0) x = reinterpret_cast<const float *>(src)[e];
    else if (src_type == 1) x = __half2float(reinterpret_cast<const __half *>(src)[e]);
    else x = bf16_bits_to_f32(reinterpret_cast<const uint16_t *>(src)[e]);
    const float amax = warp_max(fabsf(x));
    const float d = amax / 127.f;
    const float id = d ? 1.0f / d : 0.0f;
    const int q = (int)roundf(x * id);
    uint8_t *blk = dst + b * 34;
    if (lane == 0) *reinterpret_cast<__half *>(blk) = __float2half_rn(d);
    reinterpret_cast<int8_t *>(blk + 2)[lane] = (int8_t)q;
}

// ---- quantise-on-load, q4_k models: Q4_0 for the embedding tables (lm_utils.h:131-147), Q4_K for the linears ----------
__device__ __forceinline__ float load_src_element(const uint8_t *src, int src_type, long long e) {
    if (src_type == 0) return reinterpret_cast<const float *>(src)[e];
    if (src_type == 1) return __half2float(reinterpret_cast<const __half *>(src)[e]);
    return bf16_bits_to_f32(reinterpret_cast<const uint16_t *>(src)[e]);
}

// ggml quantize_row_q4_0_ref: the FIRST element of largest magnitude keeps its sign, d = that / -8,
// q = min(15, trunc(x * (1 / d) + 8.5)); low nibbles = elements 0..15, high = 16..31.  One warp per block.
__global__ void quantize_rows_q4_0_kernel(const uint8_t *src, int src_type, long long n_blocks, uint8_t *dst) {
    const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= n_blocks) return;
    const float x = load_src_element(src, src_type, b * 32 + lane);
    const float amax = warp_max(fabsf(x));
    const unsigned first = __ballot_sync(0xffffffffu, fabsf(x) == amax);
    const float carrier = amax > 0.f ? __shfl_sync(0xffffffffu, x, __ffs(first) - 1) : 0.f;
    const float d = carrier * -0.125f;
    const float id = d ? __fdiv_rn(1.0f, d) : 0.0f;
    int q = (int)(int8_t)(int)__fadd_rn(__fmul_rn(x, id), 8.5f);
    q = min(q, 15);
    const int hi = __shfl_down_sync(0xffffffffu, q, 16);
    uint8_t *blk = dst + b * 18;
    if (lane == 0) *reinterpret_cast<__half *>(blk) = __float2half_rn(d);
    if (lane < 16) blk[2 + lane] = (uint8_t)(q | (hi << 4));
}

// ggml quantize_row_q4_K_ref + make_qkx2_quants(32, 15, ..., rmin -1, rdelta 0.1, nstep 20): eight threads per
// 256-element block, one per 32-element sub-block, each running the scalar search with the same sequential fp32
// arithmetic (explicit _rn intrinsics: no contraction) so that the blocks equal the CPU quantiser's bit for bit.
constexpr int kQ4kQuantThreads = 128;
__device__ __forceinline__ int clamp_nibble(int v) { return max(0, min(15, v)); }
__device__ __forceinline__ uint32_t spread_nibbles4(uint32_t h) {          // 4 nibbles (16 bits) -> 4 bytes
    uint32_t v = h & 0xFFFFu;
    v = (v | (v << 8)) & 0x00FF00FFu;
    return (v | (v << 4)) & 0x0F0F0F0Fu;
}
__global__ void __launch_bounds__(kQ4kQuantThreads) quantize_rows_q4_K_kernel(const uint8_t *src, int src_type, long long n_blocks,
                                                                              uint8_t *dst) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = (t >> 3) < n_blocks;
    const long long b = live ? (t >> 3) : n_blocks - 1;      // idle tail threads redo the last block (shuffles stay full)
    const int j = (int)(t & 7);
    float x[32];
#pragma unroll
    for (int l = 0; l < 32; l++) x[l] = load_src_element(src, src_type, b * 256 + j * 32 + l);
    float sum_x2 = 0.f;
#pragma unroll
    for (int l = 0; l < 32; l++) sum_x2 = __fadd_rn(sum_x2, __fmul_rn(x[l], x[l]));
    const float av_x = __fsqrt_rn(__fmul_rn(sum_x2, 0.03125f));
    auto weight = [&](int l) { return __fadd_rn(av_x, fabsf(x[l])); };

    // ---- make_qkx2_quants ----
    float lo = x[0], hi = x[0], sum_w = weight(0), sum_x = __fmul_rn(sum_w, x[0]);
#pragma unroll
    for (int l = 1; l < 32; l++) {
        lo = x[l] < lo ? x[l] : lo;
        hi = x[l] > hi ? x[l] : hi;
        const float w = weight(l);
        sum_w = __fadd_rn(sum_w, w);
        sum_x = __fadd_rn(sum_x, __fmul_rn(w, x[l]));
    }
    if (lo > 0.f) lo = 0.f;
    uint32_t L[4] = {0u, 0u, 0u, 0u};
    float scale = 0.f;
    if (hi != lo) {
        float iscale = __fdiv_rn(15.f, __fsub_rn(hi, lo));
        scale = __fdiv_rn(1.f, iscale);
        float best = 0.f;
#pragma unroll
        for (int l = 0; l < 32; l++) {
            const int q = clamp_nibble(__float2int_rn(__fmul_rn(iscale, __fsub_rn(x[l], lo))));
            L[l >> 3] |= (uint32_t)q << (4 * (l & 7));
            const float diff = __fsub_rn(__fadd_rn(__fmul_rn(scale, (float)q), lo), x[l]);
            best = __fadd_rn(best, __fmul_rn(weight(l), __fmul_rn(diff, diff)));
        }
#pragma unroll 1
        for (int is = 0; is <= 20; is++) {
            iscale = __fdiv_rn(__fadd_rn(__fadd_rn(-1.f, __fmul_rn(0.1f, (float)is)), 15.f), __fsub_rn(hi, lo));
            uint32_t A[4] = {0u, 0u, 0u, 0u};
            float sum_l = 0.f, sum_l2 = 0.f, sum_xl = 0.f;
#pragma unroll
            for (int l = 0; l < 32; l++) {
                const int q = clamp_nibble(__float2int_rn(__fmul_rn(iscale, __fsub_rn(x[l], lo))));
                A[l >> 3] |= (uint32_t)q << (4 * (l & 7));
                const float wl = __fmul_rn(weight(l), (float)q);
                sum_l = __fadd_rn(sum_l, wl);
                sum_l2 = __fadd_rn(sum_l2, __fmul_rn(wl, (float)q));
                sum_xl = __fadd_rn(sum_xl, __fmul_rn(wl, x[l]));
            }
            const float D = __fsub_rn(__fmul_rn(sum_w, sum_l2), __fmul_rn(sum_l, sum_l));
            if (D > 0.f) {
                float this_scale = __fdiv_rn(__fsub_rn(__fmul_rn(sum_w, sum_xl), __fmul_rn(sum_x, sum_l)), D);
                float this_min = __fdiv_rn(__fsub_rn(__fmul_rn(sum_l2, sum_x), __fmul_rn(sum_l, sum_xl)), D);
                if (this_min > 0.f) { this_min = 0.f; this_scale = __fdiv_rn(sum_xl, sum_l2); }
                float err = 0.f;
#pragma unroll
                for (int l = 0; l < 32; l++) {
                    const float q = (float)((A[l >> 3] >> (4 * (l & 7))) & 15u);
                    const float diff = __fsub_rn(__fadd_rn(__fmul_rn(this_scale, q), this_min), x[l]);
                    err = __fadd_rn(err, __fmul_rn(weight(l), __fmul_rn(diff, diff)));
                }
                if (err < best) {
                    L[0] = A[0]; L[1] = A[1]; L[2] = A[2]; L[3] = A[3];
                    best = err; scale = this_scale; lo = this_min;
                }
            }
        }
    }
    const float neg_min = -lo;

    // ---- block level: 6-bit scales / mins against the block maxima, then the final rounding of every element ----
    float max_scale = fmaxf(scale, 0.f), max_min = fmaxf(neg_min, 0.f);
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        max_scale = fmaxf(max_scale, __shfl_xor_sync(0xffffffffu, max_scale, o));
        max_min = fmaxf(max_min, __shfl_xor_sync(0xffffffffu, max_min, o));
    }
    const float inv_scale = max_scale > 0.f ? __fdiv_rn(63.f, max_scale) : 0.f;
    const float inv_min = max_min > 0.f ? __fdiv_rn(63.f, max_min) : 0.f;
    const int ls = min(63, __float2int_rn(__fmul_rn(inv_scale, scale)) & 255);
    const int lm = min(63, __float2int_rn(__fmul_rn(inv_min, neg_min)) & 255);
    const __half d16 = __float2half_rn(__fdiv_rn(max_scale, 63.f)), dmin16 = __float2half_rn(__fdiv_rn(max_min, 63.f));
    const float d = __fmul_rn(__half2float(d16), (float)ls);
    if (d != 0.f) {
        const float dm = __fmul_rn(__half2float(dmin16), (float)lm);
        L[0] = L[1] = L[2] = L[3] = 0u;
#pragma unroll
        for (int l = 0; l < 32; l++) {
            const int q = clamp_nibble(__float2int_rn(__fdiv_rn(__fadd_rn(x[l], dm), d)));
            L[l >> 3] |= (uint32_t)q << (4 * (l & 7));
        }
    }
    const int lane = threadIdx.x & 31, base = lane & ~7;
    uint32_t lsv[8], lmv[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        lsv[k] = (uint32_t)__shfl_sync(0xffffffffu, ls, base + k);
        lmv[k] = (uint32_t)__shfl_sync(0xffffffffu, lm, base + k);
    }
    uint32_t P[4];                        // partner sub-block's nibbles (2p <-> 2p+1 share 32 bytes of qs)
#pragma unroll
    for (int k = 0; k < 4; k++) P[k] = __shfl_xor_sync(0xffffffffu, L[k], 1);
    if (!live) return;
    uint8_t *blk = dst + b * 144;
    if (j == 0) {
        uint32_t sb[12];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            sb[k] = lsv[k] | ((lsv[k + 4] >> 4) << 6);
            sb[k + 4] = lmv[k] | ((lmv[k + 4] >> 4) << 6);
            sb[k + 8] = (lsv[k + 4] & 0xFu) | ((lmv[k + 4] & 0xFu) << 4);
        }
        uint4 head;
        head.x = (uint32_t)__half_as_ushort(d16) | ((uint32_t)__half_as_ushort(dmin16) << 16);
        head.y = sb[0] | (sb[1] << 8) | (sb[2] << 16) | (sb[3] << 24);
        head.z = sb[4] | (sb[5] << 8) | (sb[6] << 16) | (sb[7] << 24);
        head.w = sb[8] | (sb[9] << 8) | (sb[10] << 16) | (sb[11] << 24);
        *reinterpret_cast<uint4 *>(blk) = head;
    }
    // bytes [16*(j&1), +16) of pair j>>1: byte l = even sub-block's nibble l | odd sub-block's nibble l << 4
    const int half = j & 1;
    const uint32_t e0 = half ? P[2] : L[0], e1 = half ? P[3] : L[1];       // even sub-block's nibbles 16*half .. +15
    const uint32_t o0 = half ? L[2] : P[0], o1 = half ? L[3] : P[1];       // odd sub-block's
    uint4 q;
    q.x = spread_nibbles4(e0) | (spread_nibbles4(o0) << 4);
    q.y = spread_nibbles4(e0 >> 16) | (spread_nibbles4(o0 >> 16) << 4);
    q.z = spread_nibbles4(e1) | (spread_nibbles4(o1) << 4);
    q.w = spread_nibbles4(e1 >> 16) | (spread_nibbles4(o1 >> 16) << 4);
    *reinterpret_cast<uint4 *>(blk + 16 + (j >> 1) * 32 + half * 16) = q;
}

// ---- load-time repack (GGUF row-major blocks -> device tiles, see common.cuh QLinear) --------------
// perm_half > 0 interleaves rows for the gated MLP: stored row v <- source row (v&1 ? perm_half + v/2 : v/2)
__device__ __forceinline__ int src_row_of(int v, int perm_half) { return perm_half > 0 ? ((v & 1) ? perm_half + (v >> 1) : (v >> 1)) : v; }

__global__ void repack_q4k_kernel(const uint8_t *src, uint8_t *qs, uint32_t *sc, uint32_t *dd,
                                  int rows, int K, int gs, int perm_half) {
    const int P = K >> 6, NSB = K >> 8;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (row, pair)
    if (idx >= (long long)rows * P) return;
    const int v = (int)(idx / P), p = (int)(idx % P);
    const int b = p >> 2, j = p & 3;
    const uint8_t *blk = src + ((size_t)src_row_of(v, perm_half) * NSB + b) * 144;
    const uint8_t *s12 = blk + 4;
    // get_scale_min_k4 for sub-blocks 2j and 2j+1
    uint32_t scv[2], mv[2];
#pragma unroll
    for (int t = 0; t < 2; t++) {
        const int s = 2 * j + t;
        if (s < 4) { scv[t] = s12[s] & 63; mv[t] = s12[s + 4] & 63; }
        else { scv[t] = (s12[s + 4] & 0xF) | ((s12[s - 4] >> 6) << 4); mv[t] = (s12[s + 4] >> 4) | ((s12[s] >> 6) << 4); }
    }
    sc[(size_t)v * P + p] = scv[0] | (scv[1] << 8) | (mv[0] << 16) | (mv[1] << 24);
    if (j == 0) dd[(size_t)v * NSB + b] = (uint32_t)blk[0] | ((uint32_t)blk[1] << 8) | ((uint32_t)blk[2] << 16) | ((uint32_t)blk[3] << 24);
    const int G = p / gs, q = p % gs, gsz = min(gs, P - G * gs);
    uint8_t *d0 = qs + (size_t)v * (K >> 1) + (size_t)G * gs * 32 + q * 16;
    uint8_t *d1 = d0 + gsz * 16;
    const uint8_t *s0 = blk + 16 + j * 32;
    *reinterpret_cast<int4 *>(d0) = *reinterpret_cast<const int4 *>(s0);        // 144-byte blocks keep 16 B alignment
    *reinterpret_cast<int4 *>(d1) = *reinterpret_cast<const int4 *>(s0 + 16);
}

__global__ void repack_q8_0_kernel(const uint8_t *src, uint8_t *qs, uint16_t *dd, int rows, int K, int gs, int perm_half) {
    const int P = K >> 5;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (row, block)
    if (idx >= (long long)rows * P) return;
    const int v = (int)(idx / P), p = (int)(idx % P);
    const uint8_t *blk = src + ((size_t)src_row_of(v, perm_half) * P + p) * 34;
    dd[(size_t)v * P + p] = (uint16_t)(blk[0] | (blk[1] << 8));
    const int G = p / gs, q = p % gs, gsz = min(gs, P - G * gs);
    uint8_t *d0 = qs + (size_t)v * K + (size_t)G * gs * 32 + q * 16;
    uint8_t *d1 = d0 + gsz * 16;
    const uint16_t *s16 = reinterpret_cast<const uint16_t *>(blk + 2);           // 34-byte blocks: 2 B alignment only
    uint16_t *o0 = reinterpret_cast<uint16_t *>(d0), *o1 = reinterpret_cast<uint16_t *>(d1);
#pragma unroll
    for (int i = 0; i < 8; i++) { o0[i] = s16[i]; o1[i] = s16[8 + i]; }
}

// ---- test-only: dequantise the REPACKED tiles back to f32 in source element order -------------------
__global__ void dequant_repacked_kernel(const QLinear w, float *out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)w.rows * w.K) return;
    const int row = (int)(idx / w.K), e = (int)(idx % w.K);
    if (w.type == 12) {
        const int P = w.K >> 6, NSB = w.K >> 8;
        const int p = e >> 6, ee = e & 63, half = ee >> 5, b = ee & 31, chunk = b >> 4;
        const int G = p / w.gs, q = p % w.gs, gsz = min(w.gs, P - G * w.gs);
        const uint8_t byte = w.qs[(size_t)row * (w.K >> 1) + (size_t)G * w.gs * 32 + (chunk ? gsz * 16 : 0) + q * 16 + (b & 15)];
        const int nib = half ? (byte >> 4) : (byte & 15);
        const uint32_t s = w.sc[(size_t)row * P + p];
        const uint32_t dd = reinterpret_cast<const uint32_t *>(w.dd)[(size_t)row * NSB + (p >> 2)];
        const float2 dm = __half22float2(*reinterpret_cast<const __half2 *>(&dd));
        const float scv = (float)(half ? ((s >> 8) & 0xff) : (s & 0xff));
        const float mv = (float)(half ? (s >> 24) : ((s >> 16) & 0xff));
        out[idx] = __fsub_rn(__fmul_rn(__fmul_rn(dm.x, scv), (float)nib), __fmul_rn(dm.y, mv));
    } else {
        const int P = w.K >> 5;
        const int p = e >> 5, b = e & 31, chunk = b >> 4;
        const int G = p / w.gs, q = p % w.gs, gsz = min(w.gs, P - G * w.gs);
        const int8_t v = (int8_t)w.qs[(size_t)row * w.K + (size_t)G * w.gs * 32 + (chunk ? gsz * 16 : 0) + q * 16 + (b & 15)];
        const float d = __half2float(reinterpret_cast<const __half *>(w.dd)[(size_t)row * P + p]);
        out[idx] = (float)v * d;
    }
}

// test-only: gather GGUF-format rows through the embedding path
__global__ void dequant_rows_kernel(const EmbTable t, const int32_t *ids, int n_rows, float *out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_rows * t.K) return;
    const int r = (int)(idx / t.K), i = (int)(idx % t.K);
    out[idx] = emb_element(t, ids[r], i);
}

}  // namespace msx
