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

constexpr int kRowsPerTile = 4;   // rows a warp retires per iteration

__host__ __device__ inline int gemv_smem_bytes(int type, int K) {
    // Q4_K: x8[K] + bsums int2[K/64] + dx float[K/256];  Q8_0: x8[K] + dx float[K/32]
    int b = (type == 12) ? K + (K / 64) * 8 + (K / 256) * 4 : K + (K / 32) * 4;
    return (b + 15) / 16 * 16 + 64;   // + block-reduce scratch
}

// ---- prologue: (optional RMSNorm) + activation quantisation into shared memory -------------------
// Σx² is accumulated in double like ggml_compute_forward_rms_norm_f32 (ggml_float), so the fp32
// `scale` is bit-identical to the CPU path in all but pathological cases.
template <int PRO>
__device__ __forceinline__ float rms_scale(const GemvArgs &a, int K, double *red) {
    if (PRO != PRO_RMS) return 1.f;
    double ss = 0.0;
    for (int i = threadIdx.x * 4; i < K; i += kThreads * 4) {
        float4 v = *reinterpret_cast<const float4 *>(a.x + i);
        ss += (double)(v.x * v.x); ss += (double)(v.y * v.y); ss += (double)(v.z * v.z); ss += (double)(v.w * v.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) tot += red[w];
    const float mean = (float)(tot / K);
    return 1.0f / sqrtf(mean + a.eps);
}

template <int PRO>
__device__ __forceinline__ void load8(const GemvArgs &a, int e0, float scale, float (&v)[8]) {
    float4 p0 = *reinterpret_cast<const float4 *>(a.x + e0);
    float4 p1 = *reinterpret_cast<const float4 *>(a.x + e0 + 4);
    v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
    if (PRO == PRO_RMS) {
        float4 a0 = *reinterpret_cast<const float4 *>(a.alpha + e0);
        float4 a1 = *reinterpret_cast<const float4 *>(a.alpha + e0 + 4);
        const float al[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = __fmul_rn(al[i], __fmul_rn(v[i], scale));   // alpha * (x * scale)
    }
}

// quantize_row_q8_K: per 256 block, max-|x| carrier, iscale = -127/max, q = min(127, rne(iscale*x)), d = 1/iscale
template <int PRO>
__device__ __forceinline__ void quantize_act_q8k(const GemvArgs &a, int K, int8_t *x8, int *bs, float *dx, double *red) {
    const float scale = rms_scale<PRO>(a, K, red);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int P = K >> 6;
    for (int b = warp; b < (K >> 8); b += kWarps) {
        const int e0 = b * 256 + lane * 8;
        float v[8];
        load8<PRO>(a, e0, scale, v);
        if (PRO == PRO_RMS && a.norm_out && blockIdx.x == 0) {
            *reinterpret_cast<float4 *>(a.norm_out + e0) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4 *>(a.norm_out + e0 + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
        float amax = 0.f, mx = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) { float ax = fabsf(v[i]); if (ax > amax) { amax = ax; mx = v[i]; } }
        const float wmax = warp_max(amax);
        // carrier = first element (lowest index) attaining the maximum magnitude
        const unsigned hit = __ballot_sync(0xffffffffu, amax == wmax);
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
        if ((lane & 3) == 0) bs[b * 8 + (lane >> 2)] = s;     // sum over one 32-wide sub-block
        if (lane == 0) dx[b] = d;
        // piece-major store: pair p = 4b + lane/8, piece = (lane%8)/2, 8 bytes at (lane&1)*8
        const int p = 4 * b + (lane >> 3), piece = (lane & 7) >> 1;
        uint2 pk;
        pk.x = (uint32_t)(q[0] & 0xff) | ((uint32_t)(q[1] & 0xff) << 8) | ((uint32_t)(q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24);
        pk.y = (uint32_t)(q[4] & 0xff) | ((uint32_t)(q[5] & 0xff) << 8) | ((uint32_t)(q[6] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
        *reinterpret_cast<uint2 *>(x8 + ((size_t)piece * P + p) * 16 + (lane & 1) * 8) = pk;
    }
}

// quantize_row_q8_0: per 32 block, d = amax/127, q = roundf(x/d), d kept as fp16
template <int PRO>
__device__ __forceinline__ void quantize_act_q8_0(const GemvArgs &a, int K, int8_t *x8, float *dx, double *red) {
    const float scale = rms_scale<PRO>(a, K, red);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int P = K >> 5;
    for (int b = warp; b < ((K + 255) >> 8); b += kWarps) {
        const int e0 = b * 256 + lane * 8;
        const bool act = e0 < K;                       // K % 32 == 0: groups of 4 lanes are uniform
        float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (act) load8<PRO>(a, e0, scale, v);
        if (PRO == PRO_RMS && a.norm_out && blockIdx.x == 0 && act) {
            *reinterpret_cast<float4 *>(a.norm_out + e0) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4 *>(a.norm_out + e0 + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
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
            uint2 pk;
            pk.x = (uint32_t)(q[0] & 0xff) | ((uint32_t)(q[1] & 0xff) << 8) | ((uint32_t)(q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24);
            pk.y = (uint32_t)(q[4] & 0xff) | ((uint32_t)(q[5] & 0xff) << 8) | ((uint32_t)(q[6] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
            *reinterpret_cast<uint2 *>(x8 + ((size_t)piece * P + blk) * 16 + (lane & 1) * 8) = pk;
        }
    }
}

// ---- epilogue ------------------------------------------------------------------------------------
template <int EPI, int R>
__device__ __forceinline__ void gemv_epilogue(const GemvArgs &a, int r0, const float (&acc)[R], int emb_token,
                                              unsigned long long &best) {
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int row = r0 + r;
        if (row >= a.w.rows) break;
        if (EPI == EPI_STORE) a.out[row] = acc[r];
        else if (EPI == EPI_RESID) a.out[row] = a.out[row] + acc[r];
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
        }
    }
}

// token that feeds depformer step `k` (host override > forced > greedy result of the previous step)
__device__ __forceinline__ int depformer_prev_token(const Ctrl *c, int k) {
    if (k == 0) return c->text_override != INT32_MIN ? c->text_override : c->out_tokens[0];
    const int f = c->force[k - 1];
    return f != INT32_MIN ? f : argmax_key_index(c->audio_key[k - 1]);
}

// ---- the kernel ----------------------------------------------------------------------------------
// WT: 12 = Q4_K, 8 = Q8_0.  LANES: lanes cooperating on one row (32 or 16).
template <int WT, int LANES, int PRO, int EPI>
__global__ void __launch_bounds__(kThreads, 2) gemv_kernel(const GemvArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int R = kRowsPerTile * LANES / 32;        // rows per lane group per iteration
    const int K = a.w.K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LANES, l = lane % LANES;

    int8_t *x8 = reinterpret_cast<int8_t *>(smem);
    int *bs = nullptr; float *dx = nullptr; double *red = nullptr;
    if (WT == 12) {
        bs = reinterpret_cast<int *>(smem + K);
        dx = reinterpret_cast<float *>(smem + K + (K >> 6) * 8);
        red = reinterpret_cast<double *>(smem + gemv_smem_bytes(12, K) - 64);
        quantize_act_q8k<PRO>(a, K, x8, bs, dx, red);
    } else {
        dx = reinterpret_cast<float *>(smem + K);
        red = reinterpret_cast<double *>(smem + gemv_smem_bytes(8, K) - 64);
        quantize_act_q8_0<PRO>(a, K, x8, dx, red);
    }
    __syncthreads();

    int emb_token = 0;
    if (EPI == EPI_ADD_EMB) emb_token = depformer_prev_token(a.ctrl, a.emb_step);
    unsigned long long best = 0ull;

    const int n_tiles = (a.w.rows + kRowsPerTile - 1) / kRowsPerTile;
    const int t_begin = (int)((long long)blockIdx.x * n_tiles / gridDim.x);
    const int t_end = (int)((long long)(blockIdx.x + 1) * n_tiles / gridDim.x);

    if (WT == 12) {
        const int P = K >> 6, NSB = K >> 8;
        const int nit = (P + LANES - 1) / LANES;
        const size_t row_qs = (size_t)(K >> 1);
        for (int tile = t_begin + warp; tile < t_end; tile += kWarps) {
            const int r0 = tile * kRowsPerTile + sub * R;
            double acc[R];
#pragma unroll
            for (int r = 0; r < R; r++) acc[r] = 0.0;
            for (int it = 0; it < nit; it++) {
                const int p = it * LANES + l;
                const int gsz = min(LANES, P - it * LANES);
                if (p < P) {
                    int4 w0[R], w1[R]; uint32_t sc[R], dd[R];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const int row = min(r0 + r, a.w.rows - 1);
                        const uint8_t *q = a.w.qs + (size_t)row * row_qs + (size_t)it * (LANES * 32) + l * 16;
                        w0[r] = ldg_stream(q);
                        w1[r] = ldg_stream(q + gsz * 16);
                        sc[r] = __ldg(a.w.sc + (size_t)row * P + p);
                        dd[r] = __ldg(reinterpret_cast<const uint32_t *>(a.w.dd) + (size_t)row * NSB + (p >> 2));
                    }
                    const int4 xa0 = *reinterpret_cast<const int4 *>(x8 + ((size_t)0 * P + p) * 16);
                    const int4 xa1 = *reinterpret_cast<const int4 *>(x8 + ((size_t)1 * P + p) * 16);
                    const int4 xb0 = *reinterpret_cast<const int4 *>(x8 + ((size_t)2 * P + p) * 16);
                    const int4 xb1 = *reinterpret_cast<const int4 *>(x8 + ((size_t)3 * P + p) * 16);
                    const int2 b2 = *reinterpret_cast<const int2 *>(bs + 2 * p);
                    const float dxv = dx[p >> 2];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        int dl = 0, dh = 0;   // dh accumulates 16 * (hi nibble) products: exact multiple of 16
                        dl = __dp4a(w0[r].x & 0x0F0F0F0F, xa0.x, dl); dh = dp4a_us((unsigned)w0[r].x & 0xF0F0F0F0u, xb0.x, dh);
                        dl = __dp4a(w0[r].y & 0x0F0F0F0F, xa0.y, dl); dh = dp4a_us((unsigned)w0[r].y & 0xF0F0F0F0u, xb0.y, dh);
                        dl = __dp4a(w0[r].z & 0x0F0F0F0F, xa0.z, dl); dh = dp4a_us((unsigned)w0[r].z & 0xF0F0F0F0u, xb0.z, dh);
                        dl = __dp4a(w0[r].w & 0x0F0F0F0F, xa0.w, dl); dh = dp4a_us((unsigned)w0[r].w & 0xF0F0F0F0u, xb0.w, dh);
                        dl = __dp4a(w1[r].x & 0x0F0F0F0F, xa1.x, dl); dh = dp4a_us((unsigned)w1[r].x & 0xF0F0F0F0u, xb1.x, dh);
                        dl = __dp4a(w1[r].y & 0x0F0F0F0F, xa1.y, dl); dh = dp4a_us((unsigned)w1[r].y & 0xF0F0F0F0u, xb1.y, dh);
                        dl = __dp4a(w1[r].z & 0x0F0F0F0F, xa1.z, dl); dh = dp4a_us((unsigned)w1[r].z & 0xF0F0F0F0u, xb1.z, dh);
                        dl = __dp4a(w1[r].w & 0x0F0F0F0F, xa1.w, dl); dh = dp4a_us((unsigned)w1[r].w & 0xF0F0F0F0u, xb1.w, dh);
                        const int isum = (int)(sc[r] & 0xff) * dl + (int)((sc[r] >> 8) & 0xff) * (dh >> 4);
                        const int imin = (int)((sc[r] >> 16) & 0xff) * b2.x + (int)(sc[r] >> 24) * b2.y;
                        const float2 dm = __half22float2(*reinterpret_cast<const __half2 *>(&dd[r]));
                        // exact products (24-bit x <24-bit) accumulated in double: order-independent result
                        acc[r] = fma((double)(dm.x * dxv), (double)isum, acc[r]);
                        acc[r] = fma(-(double)(dm.y * dxv), (double)imin, acc[r]);
                    }
                }
            }
            float accf[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
#pragma unroll
                for (int o = LANES / 2; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
                accf[r] = (float)acc[r];
            }
            if (l == 0) gemv_epilogue<EPI, R>(a, r0, accf, emb_token, best);
        }
    } else {   // Q8_0
        const int P = K >> 5;
        const int nit = (P + LANES - 1) / LANES;
        const size_t row_qs = (size_t)K;
        for (int tile = t_begin + warp; tile < t_end; tile += kWarps) {
            const int r0 = tile * kRowsPerTile + sub * R;
            double acc[R];
#pragma unroll
            for (int r = 0; r < R; r++) acc[r] = 0.0;
            for (int it = 0; it < nit; it++) {
                const int p = it * LANES + l;
                const int gsz = min(LANES, P - it * LANES);
                if (p < P) {
                    int4 w0[R], w1[R]; float dw[R];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const int row = min(r0 + r, a.w.rows - 1);
                        const uint8_t *q = a.w.qs + (size_t)row * row_qs + (size_t)it * (LANES * 32) + l * 16;
                        w0[r] = ldg_stream(q);
                        w1[r] = ldg_stream(q + gsz * 16);
                        dw[r] = __half2float(__ldg(reinterpret_cast<const __half *>(a.w.dd) + (size_t)row * P + p));
                    }
                    const int4 xa = *reinterpret_cast<const int4 *>(x8 + ((size_t)0 * P + p) * 16);
                    const int4 xb = *reinterpret_cast<const int4 *>(x8 + ((size_t)1 * P + p) * 16);
                    const float dxv = dx[p];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        int s = 0;
                        s = __dp4a(w0[r].x, xa.x, s); s = __dp4a(w0[r].y, xa.y, s); s = __dp4a(w0[r].z, xa.z, s); s = __dp4a(w0[r].w, xa.w, s);
                        s = __dp4a(w1[r].x, xb.x, s); s = __dp4a(w1[r].y, xb.y, s); s = __dp4a(w1[r].z, xb.z, s); s = __dp4a(w1[r].w, xb.w, s);
                        acc[r] = fma((double)(dw[r] * dxv), (double)s, acc[r]);
                    }
                }
            }
            float accf[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
#pragma unroll
                for (int o = LANES / 2; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
                accf[r] = (float)acc[r];
            }
            if (l == 0) gemv_epilogue<EPI, R>(a, r0, accf, emb_token, best);
        }
    }

    if (EPI == EPI_ARGMAX) {
        // CTA-level max, then one atomic per CTA
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o); best = t > best ? t : best; }
        __shared__ unsigned long long sbest[kWarps];
        if (lane == 0) sbest[warp] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long b = 0;
#pragma unroll
            for (int w = 0; w < kWarps; w++) b = sbest[w] > b ? sbest[w] : b;
            if (b) atomicMax(a.key, b);
        }
    }
}

}  // namespace msx
