// tts_kernels.cuh — the pieces only the TTS-family models need (sm_100a): cross-attention over a fixed
// conditioning memory, LayerNorm, demuxed two-stream text embeddings, low-rank embeddings.
//
// Reference:
//     moshi_streaming_multihead_cross_attention          src/moshi/modules/transformer.h:714-762
//     moshi_smha_state / init (k_cross, v_cross, f32)     transformer.h:319-396
//     torch_nn_layer_norm (ggml_norm, eps 0.0)            src/torch.h:49-60, src/moshi/models/lm_default.h:34
//     moshi_scaled_embedding_demux_*                      src/moshi/models/lm_utils.h:14-125
//     moshi_scaled_embedding (low_rank), _chained         lm_utils.h:126-217
// All of them are latency-sized (one token, a memory of a few hundred rows); they are written for exact
// agreement with the oracle (double accumulation of exact products), not for bandwidth.
#pragma once
#include "common.cuh"
#include "gemv.cuh"

namespace msx {

// ---- LayerNorm: y = (x - mean) / sqrt(var + eps) * w (+ b); mean / var in double like ggml_norm ----------
struct LayerNormArgs {
    const float *x = nullptr, *w = nullptr, *b = nullptr;
    float *y = nullptr;
    int32_t n = 0;
    float eps = 0.f;
};
constexpr int kLnThreads = 512;
__global__ void __launch_bounds__(kLnThreads) layer_norm_kernel(const LayerNormArgs a) {
    __shared__ double red[kLnThreads / 32];
    __shared__ float bc;
    griddep_launch();
    griddep_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto block_sum = [&](double v) -> double {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if (lane == 0) red[warp] = v;
        __syncthreads();
        double t = 0.0;
        for (int w = 0; w < kLnThreads / 32; w++) t += red[w];
        return t;
    };
    double s = 0.0;
    for (int i = threadIdx.x; i < a.n; i += kLnThreads) s += (double)__ldcg(a.x + i);
    const float mean = (float)(block_sum(s) / a.n);
    double s2 = 0.0;
    for (int i = threadIdx.x; i < a.n; i += kLnThreads) { const float v = __fsub_rn(__ldcg(a.x + i), mean); s2 += (double)__fmul_rn(v, v); }
    const float variance = (float)(block_sum(s2) / a.n);
    if (threadIdx.x == 0) bc = 1.0f / sqrtf(variance + a.eps);
    __syncthreads();
    const float scale = bc;
    for (int i = threadIdx.x; i < a.n; i += kLnThreads) {
        float v = __fmul_rn(__fsub_rn(__ldcg(a.x + i), mean), scale);
        v = __fmul_rn(v, a.w[i]);
        a.y[i] = a.b ? __fadd_rn(v, a.b[i]) : v;
    }
}

// ---- cross attention: softmax(K q / sqrt(Dh)) V over the f32 memory kv[tc][2*dim] = k | v, one CTA per head --
struct CrossAttnArgs {
    const float *q = nullptr;      // [dim]
    const float *kv = nullptr;     // [tc][2*dim]
    float *ctx = nullptr;          // [dim]
    int32_t tc = 0, dim = 0;
};
constexpr int kCrossThreads = 256;
template <int DH>
__global__ void __launch_bounds__(kCrossThreads) cross_attn_kernel(const CrossAttnArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    float *sc = reinterpret_cast<float *>(smem);                        // [tc]
    float *q_s = sc + ((a.tc + 3) & ~3);                                // [DH]
    double *dred = reinterpret_cast<double *>(q_s + DH);                // [8]
    double *part = dred + 8;                                            // [kCrossThreads / 32][DH]
    griddep_launch();
    griddep_wait();
    const int h = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < DH) q_s[tid] = __ldcg(a.q + h * DH + tid);
    __syncthreads();
    const float scale = 1.f / sqrtf((float)DH);
    // scores: one warp per memory row, lanes over DH
    float lmax = -INFINITY;
    for (int i = warp; i < a.tc; i += kCrossThreads / 32) {
        const float *K = a.kv + (size_t)i * 2 * a.dim + h * DH;
        double d = 0.0;
        for (int e = lane; e < DH; e += 32) d += (double)__fmul_rn(__ldg(K + e), q_s[e]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        const float s = __fmul_rn((float)d, scale);
        if (lane == 0) sc[i] = s;
        lmax = fmaxf(lmax, s);
    }
    float *fred = reinterpret_cast<float *>(dred);
    if (lane == 0) fred[warp] = lmax;
    __syncthreads();
    float gmax = fred[0];
    for (int w = 1; w < kCrossThreads / 32; w++) gmax = fmaxf(gmax, fred[w]);
    __syncthreads();
    double ls = 0.0;
    for (int i = tid; i < a.tc; i += kCrossThreads) { const float e = (float)exp((double)(sc[i] - gmax)); sc[i] = e; ls += (double)e; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, o);
    if (lane == 0) dred[warp] = ls;
    __syncthreads();
    double sum = 0.0;
    for (int w = 0; w < kCrossThreads / 32; w++) sum += dred[w];
    const float inv = (float)(1.0 / sum);
    // context: thread = (row group g, dim e)
    constexpr int NG = kCrossThreads / DH >= 1 ? kCrossThreads / DH : 1;
    const int g = tid / DH, e = tid % DH;
    double acc = 0.0;
    if (g < NG)
        for (int i = g; i < a.tc; i += NG) acc += (double)__fmul_rn(__ldg(a.kv + (size_t)i * 2 * a.dim + a.dim + h * DH + e), __fmul_rn(sc[i], inv));
    if (g < NG) part[g * DH + e] = acc;
    __syncthreads();
    if (tid < DH) {
        double t = 0.0;
        for (int gg = 0; gg < NG; gg++) t += part[gg * DH + tid];
        a.ctx[h * DH + tid] = (float)t;
    }
}
template <int DH>
__host__ inline int cross_attn_smem(int tc) { return (((tc + 3) & ~3) + DH) * 4 + 64 + (kCrossThreads / DH >= 1 ? kCrossThreads / DH : 1) * DH * 8; }

// ---- demux row fetch: token -> (left, right) rows of one table as f32 (lm_utils.h:70-86) -------------------
struct DemuxRowsArgs {
    EmbTable table;
    const Ctrl *ctrl = nullptr;
    int32_t num_embeddings = 0;
    float *left = nullptr, *right = nullptr;   // [K]
};
__device__ __forceinline__ void demux_split(int token, int num_embeddings, int &left, int &right, float &right_scale) {
    if (token < 0) token = 0;
    left = token % num_embeddings;
    right = token / num_embeddings - 1;
    right_scale = right < 0 ? 0.f : 1.f;
    if (right < 0) right = 0;
}
__device__ __forceinline__ int temporal_text_token(const Ctrl *c) {
    const int32_t *toks = c->feed_n ? c->feed + (size_t)(c->frame % c->feed_n) * c->n_in : c->tokens;
    return toks[0];
}
__global__ void demux_rows_kernel(const DemuxRowsArgs a) {
    griddep_launch();
    griddep_wait();
    int l, r; float rs;
    demux_split(temporal_text_token(a.ctrl), a.num_embeddings, l, r, rs);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.table.K; i += gridDim.x * blockDim.x) {
        a.left[i] = emb_element(a.table, l, i);
        a.right[i] = emb_element(a.table, r, i);
    }
}

// ---- small linear on GGUF-format rows (Q4_0 / Q8_0): low_rank and depformer demux projections --------------
// x = an embedding row picked by the token that feeds depformer step `step` (or a plain f32 vector), quantised to
// Q8_0 like ggml's mul_mat; y = W x.  modes: 0 scaled embedding (-1 -> zeros, other negatives -> row 0),
// 1 chained (token as is), 2 demux left, 3 demux right (y * right_scale added to out).
struct SmallLinearArgs {
    EmbTable table;
    EmbTable w;                      // [rows][K]
    const float *x = nullptr;        // non-null: use this vector instead of a table row
    const Ctrl *ctrl = nullptr;
    int32_t step = 0, mode = 0, num_embeddings = 0;
    float *out = nullptr;
    const float *addvec = nullptr;   // non-null: the result (after mode 3's accumulate) + addvec goes to dst
    float *dst = nullptr;
};
constexpr int kSmallThreads = 256;
constexpr int kSmallMaxK = 2048;
__global__ void __launch_bounds__(kSmallThreads) small_linear_kernel(const SmallLinearArgs a) {
    __shared__ float xs[kSmallMaxK];
    __shared__ __align__(16) int8_t q8[kSmallMaxK];
    __shared__ float dx[kSmallMaxK / 32];
    griddep_launch();
    griddep_wait();
    const int K = a.w.K, nb = K >> 5;
    float out_scale = 1.f;
    if (a.x) {
        for (int i = threadIdx.x; i < K; i += kSmallThreads) xs[i] = __ldcg(a.x + i);
    } else {
        int token = depformer_prev_token(a.ctrl, a.step);
        float in_scale = 1.f;
        int row = token;
        if (a.mode == 0) { in_scale = token == -1 ? 0.f : 1.f; row = token < 0 ? 0 : token; }
        else if (a.mode >= 2) { int l, r; float rs; demux_split(token, a.num_embeddings, l, r, rs); row = a.mode == 2 ? l : r; if (a.mode == 3) out_scale = rs; }
        for (int i = threadIdx.x; i < K; i += kSmallThreads) xs[i] = __fmul_rn(emb_element(a.table, row, i), in_scale);
    }
    __syncthreads();
    // quantize_row_q8_0: one thread per 32-block (K <= 2048 -> <= 64 blocks)
    if ((int)threadIdx.x < nb) {
        const int b = threadIdx.x;
        float amax = 0.f;
        for (int j = 0; j < 32; j++) amax = fmaxf(amax, fabsf(xs[b * 32 + j]));
        const float d = amax / 127.f;
        const float id = d ? 1.0f / d : 0.0f;
        for (int j = 0; j < 32; j++) q8[b * 32 + j] = (int8_t)(int)roundf(xs[b * 32 + j] * id);
        dx[b] = __half2float(__float2half_rn(d));
    }
    __syncthreads();
    const int row = blockIdx.x * kSmallThreads + threadIdx.x;
    if (row >= a.w.rows) return;
    const uint8_t *wr = a.w.data + (size_t)row * a.w.row_bytes;
    double acc = 0.0;
    for (int b = 0; b < nb; b++) {
        int sumi = 0;
        float dw;
        if (a.w.type == 2) {            // Q4_0: {fp16 d, 16 bytes: low nibbles = elements 0..15, high = 16..31}, value = nibble - 8
            const uint8_t *blk = wr + (size_t)b * 18;
            dw = __half2float(*reinterpret_cast<const __half *>(blk));
            for (int j = 0; j < 16; j++) {
                const int v = blk[2 + j];
                sumi += ((v & 15) - 8) * (int)q8[b * 32 + j] + ((v >> 4) - 8) * (int)q8[b * 32 + 16 + j];
            }
        } else {                        // Q8_0
            const uint8_t *blk = wr + (size_t)b * 34;
            dw = __half2float(*reinterpret_cast<const __half *>(blk));
            for (int j = 0; j < 32; j++) sumi += (int)(int8_t)blk[2 + j] * (int)q8[b * 32 + j];
        }
        acc = fma((double)__fmul_rn(dw, dx[b]), (double)sumi, acc);
    }
    const float y = (float)acc;
    const float r = a.mode == 3 ? __fadd_rn(a.out[row], __fmul_rn(y, out_scale)) : y;
    if (a.addvec) a.dst[row] = __fadd_rn(__ldcg(a.addvec + row), r);
    else a.out[row] = r;
}

// ---- voice conditioners (one-off per voice; src/moshi.cpp:296-366 voice_condition) --------------------------------
// FloatTensor: an unquantised GGUF tensor kept in its file type (f32 / f16 / bf16), [ne1][ne0] row-major.
struct FloatTensor {
    const uint8_t *data = nullptr;
    int32_t type = 0;                // 0 f32, 1 f16, 30 bf16
    int32_t ne0 = 0, ne1 = 1;
};
__device__ __forceinline__ float float_tensor_at(const FloatTensor &t, long long i) {
    if (t.type == 0) return reinterpret_cast<const float *>(t.data)[i];
    if (t.type == 1) return __half2float(reinterpret_cast<const __half *>(t.data)[i]);
    return bf16_bits_to_f32(reinterpret_cast<const uint16_t *>(t.data)[i]);
}
// ggml_mul_mat with an f32 / f16 / bf16 weight: the activation is rounded to the weight's type (vec_dot_type), products
// are exact in double, summed in double, rounded once (order-independent like every other reduction here).
// y[col][r] = sum_k W[r][k] * x[col * xcol + k * xstride]; row < 0: x is a vector, row >= 0: x = row `row` of `table`.
struct CondLinearArgs {
    FloatTensor w, table;
    int32_t row = -1;
    const float *x = nullptr;
    long long xstride = 1, xcol = 0;
    float *y = nullptr;              // [ncols][w.ne1]
};
__global__ void __launch_bounds__(256) cond_linear_kernel(const CondLinearArgs a) {
    const int lane = threadIdx.x & 31, r = blockIdx.x * 8 + (threadIdx.x >> 5), col = blockIdx.y;
    if (r >= a.w.ne1) return;
    double acc = 0.0;
    for (int k = lane; k < a.w.ne0; k += 32) {
        float xv = a.row >= 0 ? float_tensor_at(a.table, (long long)a.row * a.table.ne0 + k) : a.x[col * a.xcol + k * a.xstride];
        if (a.w.type == 1) xv = __half2float(__float2half_rn(xv));
        else if (a.w.type == 30) xv = __bfloat162float(__float2bfloat16_rn(xv));
        acc += (double)float_tensor_at(a.w, (long long)r * a.w.ne0 + k) * (double)xv;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) a.y[(size_t)col * a.w.ne1 + r] = (float)acc;
}
// condition_cross[t][d] = (t < T ? speaker[t][d] : learnt_padding[d]) + timestep_embedding(t)[d], t in [0, 5T):
// ggml_timestep_embedding = [cos(t f_j) | sin(t f_j)], f_j = expf(-logf(P) j / half) (host-computed like the RoPE table);
// cos / sin in double, rounded once.
__global__ void cond_cross_kernel(const float *speaker, FloatTensor pad, const float *freq, float *out, int T, int dim) {
    const int t = blockIdx.x, half = dim >> 1;
    for (int d = threadIdx.x; d < dim; d += blockDim.x) {
        const float base = t < T ? speaker[(size_t)t * dim + d] : float_tensor_at(pad, d);
        float pos = 0.f;
        if (d < 2 * half) {
            const float arg = __fmul_rn((float)t, freq[d < half ? d : d - half]);
            pos = d < half ? (float)cos((double)arg) : (float)sin((double)arg);
        }
        out[(size_t)t * dim + d] = __fadd_rn(base, pos);
    }
}
__global__ void cond_add_kernel(const float *a, const float *b, float *y, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = __fadd_rn(a[i], b[i]);
}

}  // namespace msx
