/*
 * oracle/ggml_ref.c — TEST INFRASTRUCTURE ONLY (see ggml_ref.h for the rules and parity status).
 *
 * Restates, on the CPU and in plain C:
 *   (1) the ggml CPU-backend numerics the reference's graphs execute (ggml is NOT in the
 *       reference tree: un-vendored, unpinned — README.md:183-198).  Algorithms follow the
 *       published ggml sources (ggml-quants.c / ggml-cpu ops, generic non-SIMD paths):
 *         block formats Q4_K / Q8_0 / Q4_0, quantize_row_q8_K / q8_0,
 *         vec_dot_q4_K_q8_K / q8_0_q8_0 / q4_0_q8_0, rms_norm, soft_max_ext, silu, argmax,
 *         timestep_embedding, bf16 KV "weights" with bf16-rounded second operand.
 *   (2) the reference's own graph semantics for the LM decode step:
 *         src/moshi/models/lm.h:446-690, 778-979; src/moshi/modules/transformer.h:15-23,
 *         149-249, 449-712, 910-1039, 1217-1289; src/moshi/modules/rope.h:8-128;
 *         src/moshi/modules/gating.h:12-37; src/torch.h:79-118, 162-237;
 *         src/moshi/models/lm_utils.h:126-217; src/moshi/utils/sampling.h:46-64.
 */
#include "ggml_ref.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#ifdef _OPENMP
#include <omp.h>
#ifdef __AVX2__
#include <immintrin.h>
#endif
#endif

#define QK_K 256
#define QK8_0 32
#define QK4_0 32

/* host threads of the row-parallel loops: set explicitly by the benchmark's CPU legs (torchrun exports OMP_NUM_THREADS=1) */
void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * scalar conversions (ggml-impl.h: ggml_compute_fp32_to_bf16, GGML_FP16_TO_FP32)
 * ---------------------------------------------------------------------------------------------- */
float orc_fp16_to_fp32(uint16_t h) { _Float16 f; memcpy(&f, &h, 2); return (float)f; }
uint16_t orc_fp32_to_fp16(float f) { _Float16 h = (_Float16)f; uint16_t u; memcpy(&u, &h, 2); return u; }
uint16_t orc_fp32_to_bf16(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 64); /* quiet NaN */
    return (uint16_t)((u + (0x7fffu + ((u >> 16) & 1u))) >> 16);
}
float orc_bf16_to_fp32(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
static inline float bf16_round(float f) { return orc_bf16_to_fp32(orc_fp32_to_bf16(f)); }

/* ggml-quants.c nearest_int(): round-to-nearest-even via the 1.5*2^23 magic constant */
static inline int nearest_int(float fval) {
    float val = fval + 12582912.f;
    int i; memcpy(&i, &val, sizeof(int));
    return (i & 0x007fffff) - 0x00400000;
}

/* ------------------------------------------------------------------------------------------------
 * block formats (ggml-common.h block_q4_K / block_q8_0 / block_q4_0; cross-checked with gguf/quants.py)
 * ---------------------------------------------------------------------------------------------- */
#pragma pack(push, 1)
typedef struct { uint16_t d, dmin; uint8_t scales[12]; uint8_t qs[QK_K / 2]; } block_q4_K; /* 144 B */
typedef struct { uint16_t d; int8_t qs[QK8_0]; } block_q8_0;                               /* 34 B  */
typedef struct { uint16_t d; uint8_t qs[QK4_0 / 2]; } block_q4_0;                          /* 18 B  */
#pragma pack(pop)

int64_t orc_row_size(int type, int64_t k) {
    switch (type) {
        case ORC_F32: return 4 * k;
        case ORC_F16: case ORC_BF16: return 2 * k;
        case ORC_Q4_0: return k / QK4_0 * (int64_t)sizeof(block_q4_0);
        case ORC_Q8_0: return k / QK8_0 * (int64_t)sizeof(block_q8_0);
        case ORC_Q4_K: return k / QK_K * (int64_t)sizeof(block_q4_K);
    }
    return -1;
}

/* ggml-quants.c get_scale_min_k4 */
static inline void get_scale_min_k4(int j, const uint8_t *q, uint8_t *d, uint8_t *m) {
    if (j < 4) { *d = q[j] & 63; *m = q[j + 4] & 63; }
    else {
        *d = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4);
        *m = (q[j + 4] >> 4) | ((q[j - 0] >> 6) << 4);
    }
}

/* dequantize_row_q4_K / q8_0 / q4_0 (ggml-quants.c). Compiled with -ffp-contract=off so that
 * d1*q - m1 keeps two roundings, like gguf-py's (d*sc)*q - (dmin*m). */
void orc_dequantize_row(int type, const void *src, float *y, int64_t k) {
    if (type == ORC_F32) { memcpy(y, src, 4 * k); return; }
    if (type == ORC_F16) { const uint16_t *s = src; for (int64_t i = 0; i < k; i++) y[i] = orc_fp16_to_fp32(s[i]); return; }
    if (type == ORC_BF16) { const uint16_t *s = src; for (int64_t i = 0; i < k; i++) y[i] = orc_bf16_to_fp32(s[i]); return; }
    if (type == ORC_Q8_0) {
        const block_q8_0 *x = src;
        for (int64_t i = 0; i < k / QK8_0; i++) {
            const float d = orc_fp16_to_fp32(x[i].d);
            for (int j = 0; j < QK8_0; j++) y[i * QK8_0 + j] = x[i].qs[j] * d;
        }
        return;
    }
    if (type == ORC_Q4_0) {
        const block_q4_0 *x = src;
        for (int64_t i = 0; i < k / QK4_0; i++) {
            const float d = orc_fp16_to_fp32(x[i].d);
            for (int j = 0; j < QK4_0 / 2; j++) {
                const int x0 = (x[i].qs[j] & 0x0F) - 8;
                const int x1 = (x[i].qs[j] >> 4) - 8;
                y[i * QK4_0 + j] = x0 * d;
                y[i * QK4_0 + j + QK4_0 / 2] = x1 * d;
            }
        }
        return;
    }
    if (type == ORC_Q4_K) {
        const block_q4_K *x = src;
        for (int64_t i = 0; i < k / QK_K; i++) {
            const uint8_t *q = x[i].qs;
            const float d = orc_fp16_to_fp32(x[i].d);
            const float min = orc_fp16_to_fp32(x[i].dmin);
            int is = 0; uint8_t sc, m;
            for (int j = 0; j < QK_K; j += 64) {
                get_scale_min_k4(is + 0, x[i].scales, &sc, &m);
                const float d1 = d * sc; const float m1 = min * m;
                get_scale_min_k4(is + 1, x[i].scales, &sc, &m);
                const float d2 = d * sc; const float m2 = min * m;
                for (int l = 0; l < 32; ++l) *y++ = d1 * (q[l] & 0xF) - m1;
                for (int l = 0; l < 32; ++l) *y++ = d2 * (q[l] >> 4) - m2;
                q += 32; is += 2;
            }
        }
        return;
    }
    fprintf(stderr, "orc_dequantize_row: unsupported type %d\n", type); abort();
}

/* quantize_row_q8_0_ref: d = amax/127, q = roundf(x/d), d stored fp16 */
void orc_quantize_row_q8_0(const float *x, void *dst, int64_t k) {
    block_q8_0 *y = dst;
    for (int64_t i = 0; i < k / QK8_0; i++) {
        float amax = 0.0f;
        for (int j = 0; j < QK8_0; j++) { const float v = fabsf(x[i * QK8_0 + j]); if (v > amax) amax = v; }
        const float d = amax / ((1 << 7) - 1);
        const float id = d ? 1.0f / d : 0.0f;
        y[i].d = orc_fp32_to_fp16(d);
        for (int j = 0; j < QK8_0; ++j) y[i].qs[j] = (int8_t)roundf(x[i * QK8_0 + j] * id);
    }
}

/* ggml_timestep_embedding frequencies: freq_j = expf(-logf(max_period) * j / half), libm single precision like the host
 * side of both implementations (used by the voice conditioners' position embedding, src/moshi.cpp:338-341) */
void orc_timestep_freq(int half, int max_period, float *out) {
    for (int j = 0; j < half; j++) out[j] = (float)expf(-logf((float)max_period) * j / half);
}

/* quantize_row_q4_0_ref (ggml-quants.c; cross-checked bit for bit with gguf/quants.py Q4_0.quantize_blocks):
 * the element of largest magnitude keeps its sign, d = that / -8, q = min(15, trunc(x / d + 8.5)). */
void orc_quantize_row_q4_0(const float *x, void *dst, int64_t k) {
    block_q4_0 *y = dst;
    for (int64_t i = 0; i < k / QK4_0; i++) {
        const float *xb = x + i * QK4_0;
        float amax = 0.0f, carrier = 0.0f;
        for (int j = 0; j < QK4_0; j++) if (amax < fabsf(xb[j])) { amax = fabsf(xb[j]); carrier = xb[j]; }
        const float d = carrier / -8;
        const float id = d ? 1.0f / d : 0.0f;
        y[i].d = orc_fp32_to_fp16(d);
        for (int j = 0; j < QK4_0 / 2; j++) {
            const int lo = (int8_t)(xb[j] * id + 8.5f), hi = (int8_t)(xb[QK4_0 / 2 + j] * id + 8.5f);
            y[i].qs[j] = (uint8_t)((lo < 15 ? lo : 15) | ((hi < 15 ? hi : 15) << 4));
        }
    }
}

/* make_qkx2_quants (ggml-quants.c) as quantize_row_q4_K_ref calls it: n = 32, nmax = 15, rmin = -1, rdelta = 0.1,
 * nstep = 20, squared error.  Weighted least squares for x ~ scale * L + min over a grid of 21 candidate inverse
 * scales; returns scale, *neg_min = -min (>= 0).  All sums run sequentially in fp32 like the scalar C code. */
static float q4k_fit_sub_block(const float *x, const float *w, uint8_t *L, float *neg_min) {
    enum { N = 32, NMAX = 15 };
    uint8_t Laux[N];
    float lo = x[0], hi = x[0], sum_w = w[0], sum_x = sum_w * x[0];
    for (int i = 1; i < N; i++) {
        if (x[i] < lo) lo = x[i];
        if (x[i] > hi) hi = x[i];
        sum_w += w[i];
        sum_x += w[i] * x[i];
    }
    if (lo > 0) lo = 0;
    if (hi == lo) { memset(L, 0, N); *neg_min = -lo; return 0.f; }
    float iscale = NMAX / (hi - lo), scale = 1 / iscale, best = 0;
    for (int i = 0; i < N; i++) {
        int l = nearest_int(iscale * (x[i] - lo));
        L[i] = (uint8_t)(l < 0 ? 0 : l > NMAX ? NMAX : l);
        float diff = scale * L[i] + lo - x[i];
        best += w[i] * (diff * diff);
    }
    for (int is = 0; is <= 20; is++) {
        iscale = (-1.f + 0.1f * is + NMAX) / (hi - lo);
        float sum_l = 0, sum_l2 = 0, sum_xl = 0;
        for (int i = 0; i < N; i++) {
            int l = nearest_int(iscale * (x[i] - lo));
            l = l < 0 ? 0 : l > NMAX ? NMAX : l;
            Laux[i] = (uint8_t)l;
            sum_l += w[i] * l;
            sum_l2 += w[i] * l * l;
            sum_xl += w[i] * l * x[i];
        }
        const float D = sum_w * sum_l2 - sum_l * sum_l;
        if (D > 0) {
            float this_scale = (sum_w * sum_xl - sum_x * sum_l) / D;
            float this_min = (sum_l2 * sum_x - sum_l * sum_xl) / D;
            if (this_min > 0) { this_min = 0; this_scale = sum_xl / sum_l2; }
            float err = 0;
            for (int i = 0; i < N; i++) {
                float diff = this_scale * Laux[i] + this_min - x[i];
                err += w[i] * (diff * diff);
            }
            if (err < best) { memcpy(L, Laux, N); best = err; scale = this_scale; lo = this_min; }
        }
    }
    *neg_min = -lo;
    return scale;
}

/* quantize_row_q4_K_ref (ggml-quants.c): per 32-element sub-block a (scale, min) fit with weights rms(x) + |x|,
 * the 8 scales / mins quantised to 6 bits against the block maxima (d = max_scale / 63, dmin = max_min / 63, fp16),
 * then every element re-rounded against the quantised scale / min: q = clamp(nearest_int((x + dmin*m) / (d*sc)), 0, 15).
 * Reference call site: ggml_cast in src/loader.h:183 (quantise while loading). */
void orc_quantize_row_q4_K(const float *x, void *dst, int64_t k) {
    block_q4_K *y = dst;
    for (int64_t i = 0; i < k / QK_K; i++, x += QK_K) {
        uint8_t L[QK_K];
        float scales[QK_K / 32], mins[QK_K / 32], w[32];
        float max_scale = 0, max_min = 0;
        for (int j = 0; j < QK_K / 32; j++) {
            const float *xb = x + 32 * j;
            float sum_x2 = 0;
            for (int l = 0; l < 32; l++) sum_x2 += xb[l] * xb[l];
            const float av_x = sqrtf(sum_x2 / 32);
            for (int l = 0; l < 32; l++) w[l] = av_x + fabsf(xb[l]);
            scales[j] = q4k_fit_sub_block(xb, w, L + 32 * j, &mins[j]);
            if (scales[j] > max_scale) max_scale = scales[j];
            if (mins[j] > max_min) max_min = mins[j];
        }
        const float inv_scale = max_scale > 0 ? 63.f / max_scale : 0.f;
        const float inv_min = max_min > 0 ? 63.f / max_min : 0.f;
        memset(y[i].scales, 0, sizeof(y[i].scales));
        for (int j = 0; j < QK_K / 32; j++) {
            int ls = nearest_int(inv_scale * scales[j]), lm = nearest_int(inv_min * mins[j]);
            ls = (uint8_t)ls; lm = (uint8_t)lm;
            if (ls > 63) ls = 63;
            if (lm > 63) lm = 63;
            if (j < 4) { y[i].scales[j] = (uint8_t)ls; y[i].scales[j + 4] = (uint8_t)lm; }
            else {
                y[i].scales[j + 4] = (uint8_t)((ls & 0xF) | ((lm & 0xF) << 4));
                y[i].scales[j - 4] |= (uint8_t)((ls >> 4) << 6);
                y[i].scales[j] |= (uint8_t)((lm >> 4) << 6);
            }
        }
        y[i].d = orc_fp32_to_fp16(max_scale / 63.f);
        y[i].dmin = orc_fp32_to_fp16(max_min / 63.f);
        const float fd = orc_fp16_to_fp32(y[i].d), fdmin = orc_fp16_to_fp32(y[i].dmin);
        for (int j = 0; j < QK_K / 32; j++) {
            uint8_t sc, m;
            get_scale_min_k4(j, y[i].scales, &sc, &m);
            const float d = fd * sc;
            if (!d) continue;
            const float dm = fdmin * m;
            for (int l = 0; l < 32; l++) {
                int q = nearest_int((x[32 * j + l] + dm) / d);
                L[32 * j + l] = (uint8_t)(q < 0 ? 0 : q > 15 ? 15 : q);
            }
        }
        uint8_t *q = y[i].qs;
        for (int j = 0; j < QK_K; j += 64, q += 32)
            for (int l = 0; l < 32; l++) q[l] = (uint8_t)(L[j + l] | (L[j + l + 32] << 4));
    }
}

/* quantize_row_q8_K_ref: iscale = -127/max(|x|-carrier), q = min(127, nearest_int(iscale*x)), d = 1/iscale */
void orc_quantize_row_q8_K(const float *x, int8_t *qs, float *dout, int16_t *bsums, int64_t k) {
    for (int64_t i = 0; i < k / QK_K; i++) {
        float max = 0, amax = 0;
        for (int j = 0; j < QK_K; ++j) { float ax = fabsf(x[j]); if (ax > amax) { amax = ax; max = x[j]; } }
        if (!amax) {
            dout[i] = 0; memset(qs, 0, QK_K); for (int j = 0; j < QK_K / 16; j++) bsums[j] = 0;
            x += QK_K; qs += QK_K; bsums += QK_K / 16; continue;
        }
        const float iscale = -127.f / max;
        for (int j = 0; j < QK_K; ++j) { int v = nearest_int(iscale * x[j]); qs[j] = (int8_t)(v < 127 ? v : 127); }
        for (int j = 0; j < QK_K / 16; ++j) { int sum = 0; for (int ii = 0; ii < 16; ++ii) sum += qs[j * 16 + ii]; bsums[j] = (int16_t)sum; }
        dout[i] = 1 / iscale;
        x += QK_K; qs += QK_K; bsums += QK_K / 16;
    }
}

/* ggml_vec_dot_q4_K_q8_K: integer inner sums exactly as ggml (sub-block dots scaled by the 6-bit scales,
 * bsums x mins); the per-block fp32 scales d = fp16(d_w)*d_x and dmin = fp16(dmin_w)*d_x are formed in
 * fp32 like ggml.  ACCUMULATION ORDER: ggml adds the block terms in fp32 in an ISA-dependent order
 * (generic: 8 lane partials; AVX2/AVX512/NEON: other groupings).  The oracle adds the exact products
 * d*isum and dmin*imin in double and rounds once, which makes the result independent of summation
 * order (see DESIGN.md "Order-independent arithmetic"); it differs from any fp32 order by <~1e-6 rel. */
static float vec_dot_q4_K_q8_K(int64_t n, const block_q4_K *x, const int8_t *yqs, const float *yd, const int16_t *ybsums) {
    const int nb = (int)(n / QK_K);
    uint8_t scales[8], mins[8];
    double sumd = 0.0;
    for (int i = 0; i < nb; ++i) {
        const uint8_t *q4 = x[i].qs;
        const int8_t *q8 = yqs + (int64_t)i * QK_K;
        const int16_t *bs = ybsums + (int64_t)i * (QK_K / 16);
#ifdef __AVX2__
        {   /* get_scale_min_k4 for all eight sub-blocks with word operations (the 6-bit fields of the 12 packed bytes) */
            uint32_t u[3]; memcpy(u, x[i].scales, 12);
            const uint32_t sc_lo = u[0] & 0x3f3f3f3fu, m_lo = u[1] & 0x3f3f3f3fu;
            const uint32_t sc_hi = (u[2] & 0x0f0f0f0fu) | ((u[0] >> 2) & 0x30303030u);
            const uint32_t m_hi = ((u[2] >> 4) & 0x0f0f0f0fu) | ((u[1] >> 2) & 0x30303030u);
            memcpy(scales, &sc_lo, 4); memcpy(scales + 4, &sc_hi, 4); memcpy(mins, &m_lo, 4); memcpy(mins + 4, &m_hi, 4);
        }
        int sumi;
        {   /* sum_j bsums[j] * mins[j / 2]: pair sums of the 16 int16 block sums (<= 2 * 16 * 127) times the 8 mins */
            const __m256i b16 = _mm256_loadu_si256((const __m256i *)bs);
            const __m128i pair = _mm_hadd_epi16(_mm256_castsi256_si128(b16), _mm256_extracti128_si256(b16, 1));
            const __m128i m16 = _mm_cvtepu8_epi16(_mm_loadl_epi64((const __m128i *)mins));
            __m128i s4 = _mm_madd_epi16(pair, m16);
            s4 = _mm_add_epi32(s4, _mm_shuffle_epi32(s4, 0x4E));
            s4 = _mm_add_epi32(s4, _mm_shuffle_epi32(s4, 0xB1));
            sumi = _mm_cvtsi128_si32(s4);
        }
#else
        for (int j = 0; j < 8; j++) get_scale_min_k4(j, x[i].scales, &scales[j], &mins[j]);
        int sumi = 0;
        for (int j = 0; j < QK_K / 16; ++j) sumi += bs[j] * mins[j / 2];
#endif
        int32_t isum = 0;
#ifdef __AVX2__
        /* the same exact integers with the instructions ggml's AVX2 kernel uses (maddubs on unsigned nibbles x signed q8,
         * madd with the sub-block scale): pair sums <= 2 * 15 * 127 fit int16, everything after that is int32 */
        {
            const __m256i m4 = _mm256_set1_epi8(0xF);
            __m256i acc = _mm256_setzero_si256();
            for (int j = 0; j < QK_K / 64; ++j) {
                const __m256i q4b = _mm256_loadu_si256((const __m256i *)q4);
                const __m256i lo8 = _mm256_and_si256(q4b, m4), hi8 = _mm256_and_si256(_mm256_srli_epi16(q4b, 4), m4);
                const __m256i plo = _mm256_maddubs_epi16(lo8, _mm256_loadu_si256((const __m256i *)q8));
                const __m256i phi = _mm256_maddubs_epi16(hi8, _mm256_loadu_si256((const __m256i *)(q8 + 32)));
                acc = _mm256_add_epi32(acc, _mm256_madd_epi16(plo, _mm256_set1_epi16((short)scales[2 * j])));
                acc = _mm256_add_epi32(acc, _mm256_madd_epi16(phi, _mm256_set1_epi16((short)scales[2 * j + 1])));
                q4 += 32; q8 += 64;
            }
            __m128i s4 = _mm_add_epi32(_mm256_castsi256_si128(acc), _mm256_extracti128_si256(acc, 1));
            s4 = _mm_add_epi32(s4, _mm_shuffle_epi32(s4, 0x4E));
            s4 = _mm_add_epi32(s4, _mm_shuffle_epi32(s4, 0xB1));
            isum = _mm_cvtsi128_si32(s4);
        }
#else
        for (int j = 0; j < QK_K / 64; ++j) {
            int32_t lo = 0, hi = 0;
            for (int l = 0; l < 32; ++l) { lo += (q4[l] & 0xF) * q8[l]; hi += (q4[l] >> 4) * q8[32 + l]; }
            isum += (int32_t)scales[2 * j] * lo + (int32_t)scales[2 * j + 1] * hi;
            q4 += 32; q8 += 64;
        }
#endif
        const float d = orc_fp16_to_fp32(x[i].d) * yd[i];
        const float dmin = orc_fp16_to_fp32(x[i].dmin) * yd[i];
        sumd += (double)d * (double)isum;
        sumd -= (double)dmin * (double)sumi;
    }
    return (float)sumd;
}

/* ggml_vec_dot_q8_0_q8_0: sumi * (fp16(d_w) * fp16(d_x)) per block; block terms accumulated in double (see above) */
static float vec_dot_q8_0_q8_0(int64_t n, const block_q8_0 *x, const block_q8_0 *y) {
    const int nb = (int)(n / QK8_0);
    double sumd = 0.0;
    for (int ib = 0; ib < nb; ++ib) {
        int sumi = 0;
#ifdef __AVX2__
        {   /* |x| (unsigned) x sign(x)-adjusted y: pair sums <= 2 * 127 * 127 fit int16 (ggml's mul_sum_i8_pairs) */
            const __m256i xv = _mm256_loadu_si256((const __m256i *)x[ib].qs), yv = _mm256_loadu_si256((const __m256i *)y[ib].qs);
            const __m256i p16 = _mm256_maddubs_epi16(_mm256_sign_epi8(xv, xv), _mm256_sign_epi8(yv, xv));
            const __m256i p32 = _mm256_madd_epi16(p16, _mm256_set1_epi16(1));
            __m128i s4 = _mm_add_epi32(_mm256_castsi256_si128(p32), _mm256_extracti128_si256(p32, 1));
            s4 = _mm_add_epi32(s4, _mm_shuffle_epi32(s4, 0x4E));
            s4 = _mm_add_epi32(s4, _mm_shuffle_epi32(s4, 0xB1));
            sumi = _mm_cvtsi128_si32(s4);
        }
#else
        for (int j = 0; j < QK8_0; j++) sumi += x[ib].qs[j] * y[ib].qs[j];
#endif
        const float dd = orc_fp16_to_fp32(x[ib].d) * orc_fp16_to_fp32(y[ib].d);
        sumd += (double)dd * (double)sumi;
    }
    return (float)sumd;
}

/* ggml_vec_dot_q4_0_q8_0 */
static float vec_dot_q4_0_q8_0(int64_t n, const block_q4_0 *x, const block_q8_0 *y) {
    const int nb = (int)(n / QK4_0);
    double sumd = 0.0;
    for (int ib = 0; ib < nb; ++ib) {
        int sumi0 = 0, sumi1 = 0;
        for (int j = 0; j < QK4_0 / 2; ++j) {
            const int v0 = (x[ib].qs[j] & 0x0F) - 8;
            const int v1 = (x[ib].qs[j] >> 4) - 8;
            sumi0 += v0 * y[ib].qs[j];
            sumi1 += v1 * y[ib].qs[j + QK4_0 / 2];
        }
        const float dd = orc_fp16_to_fp32(x[ib].d) * orc_fp16_to_fp32(y[ib].d);
        sumd += (double)dd * (double)(sumi0 + sumi1);
    }
    return (float)sumd;
}

/* ggml_compute_forward_mul_mat for one f32 activation column: src1 is converted to the weight
 * type's vec_dot_type (Q8_K for Q4_K, Q8_0 for Q8_0/Q4_0, bf16/f16 for bf16/f16), then rows of
 * src0 are dotted against it. */
void orc_mul_mat_vec(int type, const void *w, int64_t k, int64_t rows, const float *x, float *y) {
    const int64_t rs = orc_row_size(type, k);
    if (type == ORC_Q4_K) {
        int8_t *qs = malloc(k); float *d = malloc(sizeof(float) * (k / QK_K)); int16_t *bs = malloc(sizeof(int16_t) * (k / 16));
        orc_quantize_row_q8_K(x, qs, d, bs, k);
        #pragma omp parallel for schedule(static)
        for (int64_t r = 0; r < rows; r++)
            y[r] = vec_dot_q4_K_q8_K(k, (const block_q4_K *)((const char *)w + r * rs), qs, d, bs);
        free(qs); free(d); free(bs);
    } else if (type == ORC_Q8_0 || type == ORC_Q4_0) {
        block_q8_0 *xq = malloc(sizeof(block_q8_0) * (k / QK8_0));
        orc_quantize_row_q8_0(x, xq, k);
        #pragma omp parallel for schedule(static)
        for (int64_t r = 0; r < rows; r++) {
            const char *wr = (const char *)w + r * rs;
            y[r] = type == ORC_Q8_0 ? vec_dot_q8_0_q8_0(k, (const block_q8_0 *)wr, xq)
                                    : vec_dot_q4_0_q8_0(k, (const block_q4_0 *)wr, xq);
        }
        free(xq);
    } else if (type == ORC_F32 || type == ORC_BF16 || type == ORC_F16) {
        /* f32: plain dot (ggml_vec_dot_f32, ggml_float accumulate); bf16/f16: x rounded to that type */
        float *xr = malloc(sizeof(float) * k);
        for (int64_t i = 0; i < k; i++)
            xr[i] = type == ORC_F32 ? x[i] : type == ORC_BF16 ? bf16_round(x[i]) : orc_fp16_to_fp32(orc_fp32_to_fp16(x[i]));
        #pragma omp parallel for schedule(static)
        for (int64_t r = 0; r < rows; r++) {
            const char *wr = (const char *)w + r * rs;
            double acc = 0;
            for (int64_t i = 0; i < k; i++) {
                float wv = type == ORC_F32 ? ((const float *)wr)[i]
                         : type == ORC_BF16 ? orc_bf16_to_fp32(((const uint16_t *)wr)[i]) : orc_fp16_to_fp32(((const uint16_t *)wr)[i]);
                acc += (double)(wv * xr[i]);
            }
            y[r] = (float)acc;
        }
        free(xr);
    } else { fprintf(stderr, "orc_mul_mat_vec: unsupported type %d\n", type); abort(); }
}

/* T2 "ideal": dequantised weights, fp32 activations, double accumulation */
void orc_mul_mat_vec_ideal(int type, const void *w, int64_t k, int64_t rows, const float *x, float *y) {
    const int64_t rs = orc_row_size(type, k);
    #pragma omp parallel
    {
        float *row = malloc(sizeof(float) * k);
        #pragma omp for schedule(static)
        for (int64_t r = 0; r < rows; r++) {
            orc_dequantize_row(type, (const char *)w + r * rs, row, k);
            double acc = 0;
            for (int64_t i = 0; i < k; i++) acc += (double)row[i] * (double)x[i];
            y[r] = (float)acc;
        }
        free(row);
    }
}

/* ggml_compute_forward_rms_norm_f32 then ggml_mul(alpha, y)  (transformer.h:15-23) */
void orc_rms_norm(const float *x, const float *alpha, float eps, float *y, int64_t n) {
    double sum = 0.0;
    for (int64_t i = 0; i < n; i++) sum += (double)(x[i] * x[i]);
    const float mean = (float)(sum / n);
    const float scale = 1.0f / sqrtf(mean + eps);
    for (int64_t i = 0; i < n; i++) { float v = x[i] * scale; y[i] = alpha ? alpha[i] * v : v; }
}

/* ------------------------------------------------------------------------------------------------
 * model
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int type; int64_t ne0, ne1; const void *data; } orc_tensor;

typedef struct {
    orc_tensor norm1, norm2;
    orc_tensor *in_proj, *out_proj, *lin_in, *lin_out; /* [n_weights] */
    orc_tensor norm_cross_w, norm_cross_b, cross_in, cross_out;   /* cross-attention layers only */
} orc_layer;

struct orc_model {
    orc_config cfg;
    int ideal;
    int dep_num_weights, dep_cap;
    orc_tensor text_emb, *emb /*[n_q]*/;
    orc_tensor text_out1, text_out2;                  /* demux text embedding (lm_utils.h:14-19) */
    orc_tensor dep_text_out1, dep_text_out2, dep_text_lr, *dep_emb_lr /*[dep_q-1]*/;
    orc_layer *layers /*[num_layers]*/;
    orc_tensor out_norm, text_linear;
    orc_tensor *dep_in /*[dep_num_weights]*/, dep_text_emb, *dep_emb /*[dep_q-1]*/;
    orc_layer *dep_layers;
    orc_tensor *linears /*[dep_q]*/;
    orc_tensor *extra_heads;
};

static int dep_num_weights(const orc_config *c) {
    /* lm_default.h:72-83 */
    int n = c->dep_q;
    if (c->schedule_len) { int mx = c->schedule[0]; for (int i = 0; i < c->schedule_len; i++) if (c->schedule[i] > mx) mx = c->schedule[i]; n = mx + 1; }
    return n;
}

orc_model *orc_model_new(const orc_config *cfg) {
    orc_model *m = calloc(1, sizeof(*m));
    m->cfg = *cfg;
    const orc_config *c = &m->cfg;
    m->dep_num_weights = dep_num_weights(c);
    m->dep_cap = c->dep_context ? c->dep_context : c->schedule_len; /* transformer.h:327, lm_default.h:92 */
    m->emb = calloc(c->n_q, sizeof(orc_tensor));
    m->layers = calloc(c->num_layers, sizeof(orc_layer));
    for (int i = 0; i < c->num_layers; i++) {
        orc_layer *l = &m->layers[i];
        l->in_proj = calloc(1, sizeof(orc_tensor)); l->out_proj = calloc(1, sizeof(orc_tensor));
        l->lin_in = calloc(1, sizeof(orc_tensor)); l->lin_out = calloc(1, sizeof(orc_tensor));
    }
    if (c->dep_q > 0) {
        int nw = m->dep_num_weights;
        m->dep_in = calloc(nw, sizeof(orc_tensor));
        m->dep_emb = calloc(c->dep_q > 1 ? c->dep_q - 1 : 1, sizeof(orc_tensor));
        m->dep_emb_lr = calloc(c->dep_q > 1 ? c->dep_q - 1 : 1, sizeof(orc_tensor));
        m->dep_layers = calloc(c->dep_layers, sizeof(orc_layer));
        for (int i = 0; i < c->dep_layers; i++) {
            orc_layer *l = &m->dep_layers[i];
            l->in_proj = calloc(nw, sizeof(orc_tensor)); l->out_proj = calloc(nw, sizeof(orc_tensor));
            l->lin_in = calloc(nw, sizeof(orc_tensor)); l->lin_out = calloc(nw, sizeof(orc_tensor));
        }
        m->linears = calloc(c->dep_q, sizeof(orc_tensor));
    }
    if (c->extra_heads > 0) m->extra_heads = calloc(c->extra_heads, sizeof(orc_tensor));
    return m;
}

void orc_model_free(orc_model *m) {
    if (!m) return;
    const orc_config *c = &m->cfg;
    for (int i = 0; i < c->num_layers; i++) { orc_layer *l = &m->layers[i]; free(l->in_proj); free(l->out_proj); free(l->lin_in); free(l->lin_out); }
    if (m->dep_layers) for (int i = 0; i < c->dep_layers; i++) { orc_layer *l = &m->dep_layers[i]; free(l->in_proj); free(l->out_proj); free(l->lin_in); free(l->lin_out); }
    free(m->emb); free(m->layers); free(m->dep_in); free(m->dep_emb); free(m->dep_emb_lr); free(m->dep_layers); free(m->linears); free(m->extra_heads);
    free(m);
}

void orc_model_set_ideal(orc_model *m, int ideal) { m->ideal = ideal; }

/* name resolution follows get_weights(): lm.h:370-395, transformer.h:764-779, 1042-1080, gating.h:39-43 */
static orc_tensor *find_slot(orc_model *m, const char *name) {
    const orc_config *c = &m->cfg;
    int a, b; char tail[64];
    if (!strcmp(name, "lm.text_emb.weight")) return &m->text_emb;
    if (!strcmp(name, "lm.out_norm.alpha")) return &m->out_norm;
    if (!strcmp(name, "lm.text_linear.weight")) return &m->text_linear;
    if (!strcmp(name, "lm.depformer_text_emb.weight")) return c->dep_q > 0 ? &m->dep_text_emb : NULL;
    if (!strcmp(name, "lm.text_emb.out1.weight")) return c->demux_second_stream ? &m->text_out1 : NULL;
    if (!strcmp(name, "lm.text_emb.out2.weight")) return c->demux_second_stream ? &m->text_out2 : NULL;
    if (!strcmp(name, "lm.depformer_text_emb.out1.weight")) return (c->demux_second_stream && c->dep_q > 0) ? &m->dep_text_out1 : NULL;
    if (!strcmp(name, "lm.depformer_text_emb.out2.weight")) return (c->demux_second_stream && c->dep_q > 0) ? &m->dep_text_out2 : NULL;
    if (!strcmp(name, "lm.depformer_text_emb.low_rank.weight")) return (c->dep_low_rank && !c->demux_second_stream && c->dep_q > 0) ? &m->dep_text_lr : NULL;
    if (sscanf(name, "lm.depformer_emb.%d.low_rank.weigh%1[t]", &a, tail) == 2) return (c->dep_low_rank && m->dep_emb_lr && a < c->dep_q - 1) ? &m->dep_emb_lr[a] : NULL;
    if (sscanf(name, "lm.emb.%d.weigh%1[t]", &a, tail) == 2) return a < c->n_q ? &m->emb[a] : NULL;
    if (sscanf(name, "lm.depformer_in.%d.weigh%1[t]", &a, tail) == 2) return (m->dep_in && a < m->dep_num_weights) ? &m->dep_in[a] : NULL;
    if (sscanf(name, "lm.depformer_emb.%d.weigh%1[t]", &a, tail) == 2) return (m->dep_emb && a < c->dep_q - 1) ? &m->dep_emb[a] : NULL;
    if (sscanf(name, "lm.linears.%d.weigh%1[t]", &a, tail) == 2) return (m->linears && a < c->dep_q) ? &m->linears[a] : NULL;
    if (sscanf(name, "lm.extra_heads.%d.weigh%1[t]", &a, tail) == 2) return (m->extra_heads && a < c->extra_heads) ? &m->extra_heads[a] : NULL;
    if (sscanf(name, "lm.transformer.layers.%d.%63s", &a, tail) == 2 && a < c->num_layers) {
        orc_layer *l = &m->layers[a];
        if (!strcmp(tail, "norm1.alpha")) return &l->norm1;
        if (!strcmp(tail, "norm2.alpha")) return &l->norm2;
        if (!strcmp(tail, "self_attn.in_projs.0.weight")) return &l->in_proj[0];
        if (!strcmp(tail, "self_attn.out_projs.0.weight")) return &l->out_proj[0];
        if (!strcmp(tail, "gating.linear_in.weight")) return &l->lin_in[0];
        if (!strcmp(tail, "gating.linear_out.weight")) return &l->lin_out[0];
        if (c->cross_attention) {
            if (!strcmp(tail, "norm_cross.weight")) return &l->norm_cross_w;
            if (!strcmp(tail, "norm_cross.bias")) return &l->norm_cross_b;
            if (!strcmp(tail, "cross_attention.in_projs.0.weight")) return &l->cross_in;
            if (!strcmp(tail, "cross_attention.out_projs.0.weight")) return &l->cross_out;
        }
        return NULL;
    }
    if (m->dep_layers && sscanf(name, "lm.depformer.layers.%d.%63s", &a, tail) == 2 && a < c->dep_layers) {
        orc_layer *l = &m->dep_layers[a];
        if (!strcmp(tail, "norm1.alpha")) return &l->norm1;
        if (!strcmp(tail, "norm2.alpha")) return &l->norm2;
        if (sscanf(tail, "self_attn.in_projs.%d.weigh%1[t]", &b, tail + 60) == 2) return b < m->dep_num_weights ? &l->in_proj[b] : NULL;
        if (sscanf(tail, "self_attn.out_projs.%d.weigh%1[t]", &b, tail + 60) == 2) return b < m->dep_num_weights ? &l->out_proj[b] : NULL;
        if (sscanf(tail, "gating.%d.linear_in.weigh%1[t]", &b, tail + 60) == 2) return b < m->dep_num_weights ? &l->lin_in[b] : NULL;
        if (sscanf(tail, "gating.%d.linear_out.weigh%1[t]", &b, tail + 60) == 2) return b < m->dep_num_weights ? &l->lin_out[b] : NULL;
        /* single-weight depformer: "gating.linear_in.weight" (transformer.h:1064-1066) */
        if (!strcmp(tail, "gating.linear_in.weight")) return &l->lin_in[0];
        if (!strcmp(tail, "gating.linear_out.weight")) return &l->lin_out[0];
        return NULL;
    }
    return NULL;
}

int orc_model_set_tensor(orc_model *m, const char *name, int type, int64_t ne0, int64_t ne1, const void *data) {
    orc_tensor *t = find_slot(m, name);
    if (!t) return 0;
    t->type = type; t->ne0 = ne0; t->ne1 = ne1; t->data = data;
    return 1;
}

int orc_model_missing(orc_model *m, char *buf, int buflen) {
    const orc_config *c = &m->cfg; int n = 0; if (buf && buflen) buf[0] = 0;
#define CHK(t, nm) do { if (!(t).data) { if (!n && buf) snprintf(buf, buflen, "%s", nm); n++; } } while (0)
    CHK(m->text_emb, "text_emb"); CHK(m->out_norm, "out_norm"); CHK(m->text_linear, "text_linear");
    for (int i = 0; i < c->n_q; i++) CHK(m->emb[i], "emb");
    if (c->demux_second_stream) { CHK(m->text_out1, "text_emb.out1"); CHK(m->text_out2, "text_emb.out2"); }
    if (c->cross_attention) for (int i = 0; i < c->num_layers; i++) { orc_layer *l = &m->layers[i]; CHK(l->norm_cross_w, "norm_cross.weight"); CHK(l->cross_in, "cross_attention.in_proj"); CHK(l->cross_out, "cross_attention.out_proj"); }
    if (c->dep_q > 0 && c->demux_second_stream) { CHK(m->dep_text_out1, "depformer_text_emb.out1"); CHK(m->dep_text_out2, "depformer_text_emb.out2"); }
    if (c->dep_q > 0 && c->dep_low_rank) { if (!c->demux_second_stream) CHK(m->dep_text_lr, "depformer_text_emb.low_rank"); for (int i = 0; i < c->dep_q - 1; i++) CHK(m->dep_emb_lr[i], "depformer_emb.low_rank"); }
    for (int i = 0; i < c->num_layers; i++) { orc_layer *l = &m->layers[i]; CHK(l->norm1, "norm1"); CHK(l->norm2, "norm2"); CHK(l->in_proj[0], "in_proj"); CHK(l->out_proj[0], "out_proj"); CHK(l->lin_in[0], "lin_in"); CHK(l->lin_out[0], "lin_out"); }
    if (c->dep_q > 0) {
        CHK(m->dep_text_emb, "dep_text_emb");
        for (int i = 0; i < m->dep_num_weights; i++) CHK(m->dep_in[i], "dep_in");
        for (int i = 0; i < c->dep_q - 1; i++) CHK(m->dep_emb[i], "dep_emb");
        for (int i = 0; i < c->dep_q; i++) CHK(m->linears[i], "linears");
        for (int i = 0; i < c->dep_layers; i++) { orc_layer *l = &m->dep_layers[i]; CHK(l->norm1, "dnorm1"); CHK(l->norm2, "dnorm2");
            for (int w = 0; w < m->dep_num_weights; w++) { CHK(l->in_proj[w], "din_proj"); CHK(l->out_proj[w], "dout_proj"); CHK(l->lin_in[w], "dlin_in"); CHK(l->lin_out[w], "dlin_out"); } }
    }
    for (int i = 0; i < c->extra_heads; i++) CHK(m->extra_heads[i], "extra_heads");
#undef CHK
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * state (transformer.h:149-172 KV ring, bf16, zero-initialised; lm.h:423-436 transformer_out)
 * ---------------------------------------------------------------------------------------------- */
struct orc_state {
    orc_model *m;
    int offset;                 /* moshi_streaming_transformer_state_t::offset */
    uint16_t *k, *v;            /* [L][H][cap][Dh] */
    uint16_t *dk, *dv;          /* [Ld][Hd][capd][Dhd] */
    float *transformer_out;     /* [dim] */
    float *cond_sum;            /* [dim] or NULL (lm.h:575-577) */
    float *kv_cross; int tc;    /* [L][tc][2*dim] f32: k | v of the cross-attention memory (transformer.h:343-396) */
    /* sampling (sampling.h:46-64): temp <= 0 greedy; noise = Exp(1) draws in candidate order */
    float temp_text, temp_audio; int top_k_text, top_k_audio;
    const float *noise_text, *noise_audio;
};

static size_t kv_elems(const orc_config *c) { return (size_t)c->num_layers * c->context * c->dim; }
static size_t dkv_elems(const orc_model *m) { return (size_t)m->cfg.dep_layers * m->dep_cap * m->cfg.dep_dim; }

orc_state *orc_state_new(orc_model *m) {
    orc_state *s = calloc(1, sizeof(*s));
    s->m = m;
    s->k = calloc(kv_elems(&m->cfg), 2); s->v = calloc(kv_elems(&m->cfg), 2);
    if (m->cfg.dep_q > 0) { s->dk = calloc(dkv_elems(m), 2); s->dv = calloc(dkv_elems(m), 2); }
    s->transformer_out = calloc(m->cfg.dim, sizeof(float));
    return s;
}
void orc_state_free(orc_state *s) { if (!s) return; free(s->k); free(s->v); free(s->dk); free(s->dv); free(s->transformer_out); free(s->cond_sum); free(s->kv_cross); free(s); }
void orc_state_reset(orc_state *s) {
    s->offset = 0;
    memset(s->k, 0, kv_elems(&s->m->cfg) * 2); memset(s->v, 0, kv_elems(&s->m->cfg) * 2);
    if (s->dk) { memset(s->dk, 0, dkv_elems(s->m) * 2); memset(s->dv, 0, dkv_elems(s->m) * 2); }
    memset(s->transformer_out, 0, sizeof(float) * s->m->cfg.dim);
}
int orc_state_offset(orc_state *s) { return s->offset; }
void orc_state_get_kv(orc_state *s, int layer, int head, int slot, uint16_t *k, uint16_t *v) {
    const orc_config *c = &s->m->cfg; int Dh = c->dim / c->num_heads;
    size_t o = (((size_t)layer * c->num_heads + head) * c->context + slot) * Dh;
    memcpy(k, s->k + o, 2 * Dh); memcpy(v, s->v + o, 2 * Dh);
}

/* ------------------------------------------------------------------------------------------------
 * building blocks
 * ---------------------------------------------------------------------------------------------- */
static void linear(const orc_model *m, const orc_tensor *w, const float *x, float *y) {
    /* torch_nn_linear (torch.h:79-87); LM linears carry no bias */
    if (m->ideal) orc_mul_mat_vec_ideal(w->type, w->data, w->ne0, w->ne1, x, y);
    else orc_mul_mat_vec(w->type, w->data, w->ne0, w->ne1, x, y);
}

static void embed_row(const orc_tensor *t, int token, float *row);
#define embed_row_k embed_row
/* torch_nn_linear_view (torch.h:103-118): rows [row0, row0 + rows) of w */
static void linear_rows(const orc_model *m, const orc_tensor *w, int64_t row0, int64_t rows, const float *x, float *y) {
    const char *d = (const char *)w->data + row0 * orc_row_size(w->type, w->ne0);
    if (m->ideal) orc_mul_mat_vec_ideal(w->type, d, w->ne0, rows, x, y);
    else orc_mul_mat_vec(w->type, d, w->ne0, rows, x, y);
}

/* torch_nn_layer_norm (torch.h:49-60) = ggml_norm (mean, then variance of the centred values, both in double,
 * scale = 1/sqrtf(var + eps)) * weight (+ bias) */
static void layer_norm(const float *x, const float *w, const float *b, float eps, float *y, int64_t n) {
    double sum = 0.0;
    for (int64_t i = 0; i < n; i++) sum += (double)x[i];
    const float mean = (float)(sum / n);
    double sum2 = 0.0;
    for (int64_t i = 0; i < n; i++) { const float v = x[i] - mean; y[i] = v; sum2 += (double)(v * v); }
    const float variance = (float)(sum2 / n);
    const float scale = 1.0f / sqrtf(variance + eps);
    for (int64_t i = 0; i < n; i++) { float v = y[i] * scale; v = v * w[i]; y[i] = b ? v + b[i] : v; }
}

/* moshi_scaled_embedding_demux (lm_utils.h:42-125): token -> (left, right) rows of one table, out1(left) + out2(right)*scale */
static void embed_demux(const orc_model *m, const orc_tensor *table, const orc_tensor *out1, const orc_tensor *out2,
                        int num_embeddings, int token, float *y /*[out rows]*/) {
    if (token < 0) token = 0;
    const int left = token % num_embeddings;
    int right = token / num_embeddings - 1;
    const int right_zero = right < 0;
    if (right < 0) right = 0;
    const int64_t k = table->ne0, n = out1->ne1;
    float *rl = malloc(sizeof(float) * k), *rr = malloc(sizeof(float) * k), *yr = malloc(sizeof(float) * n);
    orc_dequantize_row(table->type, (const char *)table->data + (int64_t)left * orc_row_size(table->type, k), rl, k);
    orc_dequantize_row(table->type, (const char *)table->data + (int64_t)right * orc_row_size(table->type, k), rr, k);
    linear(m, out2, rr, yr);
    linear(m, out1, rl, y);
    const float sc = right_zero ? 0.f : 1.f;
    for (int64_t i = 0; i < n; i++) y[i] = y[i] + yr[i] * sc;
    free(rl); free(rr); free(yr);
}

/* moshi_scaled_embedding with low_rank (lm_utils.h:155-168, 209-217): y = low_rank(get_rows(w)[*scale]) */
static void embed_low_rank(const orc_model *m, const orc_tensor *table, const orc_tensor *lr, int token, int scaled, float *y) {
    const int64_t k = table->ne0;
    float *row = malloc(sizeof(float) * k);
    if (scaled) embed_row_k(table, token, row);
    else orc_dequantize_row(table->type, (const char *)table->data + (int64_t)token * orc_row_size(table->type, k), row, k);
    linear(m, lr, row, y);
    free(row);
}

/* moshi_scaled_embedding_step + get_rows*scale (lm_utils.h:157-182): -1 -> zeros, other negatives -> row 0 */
static void embed_row(const orc_tensor *t, int token, float *row /*[ne0]*/) {
    int is_zero = token == -1;
    if (token < 0) token = 0;
    orc_dequantize_row(t->type, (const char *)t->data + (int64_t)token * orc_row_size(t->type, t->ne0), row, t->ne0);
    const float scale = is_zero ? 0.f : 1.f;
    for (int64_t i = 0; i < t->ne0; i++) row[i] = row[i] * scale;
}

/* moshi_get_timestep_embedding (rope.h:8-20) = ggml_timestep_embedding on ts = arange(T)+offset:
 *   freq_j = expf(-logf(max_period) * j / half);  arg = ts * freq_j;  rotr = cos(arg), roti = sin(arg) */
/* cos/sin/exp are evaluated in double and rounded to float (= the correctly rounded fp32 value except
 * with probability ~1e-8), so that the result does not depend on which libm computes it. */
static void rope_table(float offset_f32, int Dh, int max_period, float *rotr, float *roti) {
    const int half = Dh / 2;
    for (int j = 0; j < half; j++) {
        float freq = (float)expf(-logf((float)max_period) * j / half);
        float arg = offset_f32 * freq;
        rotr[j] = (float)cos((double)arg); roti[j] = (float)sin((double)arg);
    }
}

/* moshi_apply_rope (rope.h:33-128): interleaved pairs rotated, output [re half | im half] */
static void apply_rope(const float *u, float *out, int Dh, const float *rotr, const float *roti) {
    const int half = Dh / 2;
    for (int j = 0; j < half; j++) {
        const float r = u[2 * j], i = u[2 * j + 1];
        out[j] = r * rotr[j] - i * roti[j];
        out[half + j] = r * roti[j] + i * rotr[j];
    }
}

/* torch_nn_functional_scaled_dot_product_attention_custom (torch.h:225-237) for T=1 on a bf16 ring:
 *   scores = mul_mat(K_bf16, q)  -> q rounded to bf16; soft_max_ext(scale, bias); mul_mat(V^T_bf16, p) -> p rounded to bf16.
 *   bias window for T=1: slot i visible iff i <= pos or pos >= cap-1 (torch.h:170-223). */
static void attention_head(const uint16_t *K, const uint16_t *V /*[cap][Dh]*/, int cap, int Dh, int pos,
                           const float *q, float *ctx, float *scratch /*[cap]*/, int ideal) {
    const int n_valid = (pos >= cap - 1) ? cap : pos + 1;
    const float scale = 1.f / sqrtf((float)Dh);
    float maxv = -INFINITY;
    for (int i = 0; i < n_valid; i++) {
        double acc = 0;
        for (int d = 0; d < Dh; d++) {
            float qq = ideal ? q[d] : bf16_round(q[d]);
            acc += (double)(orc_bf16_to_fp32(K[(size_t)i * Dh + d]) * qq);
        }
        float s = (float)acc;
        s = s * scale + 0.0f;
        scratch[i] = s; if (s > maxv) maxv = s;
    }
    double sum = 0;
    for (int i = 0; i < n_valid; i++) { float e = (float)exp((double)(scratch[i] - maxv)); scratch[i] = e; sum += (double)e; }
    const float inv = (float)(1.0 / sum);
    for (int i = 0; i < n_valid; i++) { float p = scratch[i] * inv; scratch[i] = ideal ? p : bf16_round(p); }
    for (int d = 0; d < Dh; d++) {
        double acc = 0;
        for (int i = 0; i < n_valid; i++) acc += (double)(orc_bf16_to_fp32(V[(size_t)i * Dh + d]) * scratch[i]);
        ctx[d] = (float)acc;
    }
}

/* moshi_streaming_transformer_layer (transformer.h:910-1039), T = 1 */
/* SDPA over the f32 cross-attention memory, no mask (transformer.h:714-762; torch.h:225-237 with f32 "weights") */
static void cross_attention_head(const float *kv /*[tc][2*dim]*/, int tc, int dim, int h, int Dh, const float *q, float *ctx, float *scratch) {
    const float scale = 1.f / sqrtf((float)Dh);
    float maxv = -INFINITY;
    for (int i = 0; i < tc; i++) {
        const float *K = kv + (size_t)i * 2 * dim + h * Dh;
        double acc = 0;
        for (int d = 0; d < Dh; d++) acc += (double)(K[d] * q[d]);
        float s = (float)acc;
        s = s * scale;
        scratch[i] = s; if (s > maxv) maxv = s;
    }
    double sum = 0;
    for (int i = 0; i < tc; i++) { float e = (float)exp((double)(scratch[i] - maxv)); scratch[i] = e; sum += (double)e; }
    const float inv = (float)(1.0 / sum);
    for (int i = 0; i < tc; i++) scratch[i] = scratch[i] * inv;
    for (int d = 0; d < Dh; d++) {
        double acc = 0;
        for (int i = 0; i < tc; i++) acc += (double)(kv[(size_t)i * 2 * dim + dim + h * Dh + d] * scratch[i]);
        ctx[d] = (float)acc;
    }
}

static void transformer_layer(const orc_model *m, const orc_layer *l, int w, int dim, int H, int cap,
                              int max_period, int pos, uint16_t *Kc, uint16_t *Vc /*[H][cap][Dh]*/, float *x,
                              const float *kv_cross /*[tc][2*dim] or NULL*/, int tc) {
    const int Dh = dim / H;
    const int F = (int)l->lin_out[w].ne0;
    float *nx = malloc(sizeof(float) * dim), *p = malloc(sizeof(float) * 3 * dim), *ctx = malloc(sizeof(float) * dim);
    float *upd = malloc(sizeof(float) * dim), *g = malloc(sizeof(float) * 2 * F), *mm = malloc(sizeof(float) * F);
    float *scratch = malloc(sizeof(float) * cap), *rotr = malloc(sizeof(float) * Dh), *roti = rotr + Dh / 2;
    float *qr = malloc(sizeof(float) * Dh), *kr = malloc(sizeof(float) * Dh);

    orc_rms_norm(x, (const float *)l->norm1.data, 1e-8f, nx, dim);
    linear(m, &l->in_proj[w], nx, p);
    const int slot = pos % cap;
    if (max_period) rope_table((float)pos, Dh, max_period, rotr, roti);
    for (int h = 0; h < H; h++) {
        const float *q = p + h * Dh, *k = p + dim + h * Dh, *v = p + 2 * dim + h * Dh;
        if (max_period) { apply_rope(q, qr, Dh, rotr, roti); apply_rope(k, kr, Dh, rotr, roti); }
        else { memcpy(qr, q, sizeof(float) * Dh); memcpy(kr, k, sizeof(float) * Dh); }
        uint16_t *Kh = Kc + (size_t)h * cap * Dh, *Vh = Vc + (size_t)h * cap * Dh;
        for (int d = 0; d < Dh; d++) { Kh[(size_t)slot * Dh + d] = orc_fp32_to_bf16(kr[d]); Vh[(size_t)slot * Dh + d] = orc_fp32_to_bf16(v[d]); }
        attention_head(Kh, Vh, cap, Dh, pos, qr, ctx + h * Dh, scratch, m->ideal);
    }
    linear(m, &l->out_proj[w], ctx, upd);
    for (int i = 0; i < dim; i++) x[i] = x[i] + upd[i];

    if (l->cross_in.data && kv_cross && tc > 0) {
        /* transformer.h:936-943: nx = layer_norm(x) (eps 0.0, lm_default.h:34); q = in_proj rows [0, dim) */
        float *cs = malloc(sizeof(float) * tc);
        layer_norm(x, (const float *)l->norm_cross_w.data, (const float *)l->norm_cross_b.data, 0.0f, nx, dim);
        linear_rows(m, &l->cross_in, 0, dim, nx, p);
        for (int h = 0; h < H; h++) cross_attention_head(kv_cross, tc, dim, h, Dh, p + h * Dh, ctx + h * Dh, cs);
        linear(m, &l->cross_out, ctx, upd);
        for (int i = 0; i < dim; i++) x[i] = x[i] + upd[i];
        free(cs);
    }

    orc_rms_norm(x, (const float *)l->norm2.data, 1e-8f, nx, dim);
    linear(m, &l->lin_in[w], nx, g);
    /* gating.h:12-37: silu(left) * right; ggml silu = x/(1+expf(-x)), exp via double (see rope_table) */
    for (int i = 0; i < F; i++) { float a = g[i]; float s = a / (1.0f + (float)exp((double)(-a))); mm[i] = s * g[F + i]; }
    linear(m, &l->lin_out[w], mm, upd);
    for (int i = 0; i < dim; i++) x[i] = x[i] + upd[i];

    free(nx); free(p); free(ctx); free(upd); free(g); free(mm); free(scratch); free(rotr); free(qr); free(kr);
}

/* moshi_sample_token (sampling.h:4-64) for use_sampling && temp > 0:
 *   probs = soft_max(logits * (1/temp))  [max, exp(x - max) through double, sum in double, * (float)(1/sum)]
 *   top-k by descending probability (ties: ascending id), q_j = p_j / e_j, first arg-max -> token */
typedef struct { float p; int id; } orc_cand;
static int cand_cmp(const void *a, const void *b) {
    const orc_cand *x = a, *y = b;
    if (x->p > y->p) return -1; if (x->p < y->p) return 1;
    return x->id < y->id ? -1 : (x->id > y->id ? 1 : 0);
}
static int sample_top_k(const float *logits, int n, float temp, int k, const float *noise) {
    const float inv_temp = 1.f / temp;
    orc_cand *c = malloc(sizeof(orc_cand) * n);
    float mx = -INFINITY;
    for (int i = 0; i < n; i++) { float v = logits[i] * inv_temp; if (v > mx) mx = v; }
    double sum = 0;
    for (int i = 0; i < n; i++) { float e = (float)exp((double)(logits[i] * inv_temp - mx)); c[i].p = e; c[i].id = i; sum += (double)e; }
    const float inv = (float)(1.0 / sum);
    for (int i = 0; i < n; i++) c[i].p = c[i].p * inv;
    qsort(c, n, sizeof(orc_cand), cand_cmp);
    if (k > n) k = n;
    int best = 0; float bq = c[0].p / noise[0];
    for (int j = 1; j < k; j++) { float q = c[j].p / noise[j]; if (q > bq) { bq = q; best = j; } }
    int tok = c[best].id;
    free(c);
    return tok;
}
void orc_state_set_sampling(orc_state *s, float temp_text, float temp_audio, int top_k_text, int top_k_audio) {
    s->temp_text = temp_text; s->temp_audio = temp_audio; s->top_k_text = top_k_text; s->top_k_audio = top_k_audio;
}
void orc_state_set_noise(orc_state *s, const float *noise_text, const float *noise_audio) { s->noise_text = noise_text; s->noise_audio = noise_audio; }

static int argmax_first(const float *v, int n) { int bi = 0; float bv = v[0]; for (int i = 1; i < n; i++) if (v[i] > bv) { bv = v[i]; bi = i; } return bi; }

static int temporal_from_x(orc_model *m, orc_state *s, float *x, float *text_logits, float *transformer_out) {
    const orc_config *c = &m->cfg; const int dim = c->dim;
    const int pos = s->offset; s->offset += 1;          /* transformer.h:1269-1270 */
    const size_t lstride = (size_t)c->context * dim;
    for (int l = 0; l < c->num_layers; l++)
        transformer_layer(m, &m->layers[l], 0, dim, c->num_heads, c->context, c->max_period, pos,
                          s->k + l * lstride, s->v + l * lstride, x,
                          s->kv_cross ? s->kv_cross + (size_t)l * s->tc * 2 * dim : NULL, s->tc);
    orc_rms_norm(x, (const float *)m->out_norm.data, 1e-8f, s->transformer_out, dim);   /* lm.h:671-672 */
    float *logits = text_logits ? text_logits : malloc(sizeof(float) * c->text_card);
    linear(m, &m->text_linear, s->transformer_out, logits);
    int tok = (s->temp_text > 0.f && s->noise_text) ? sample_top_k(logits, (int)m->text_linear.ne1, s->temp_text, s->top_k_text, s->noise_text)
                                                   : argmax_first(logits, (int)m->text_linear.ne1);
    if (transformer_out) memcpy(transformer_out, s->transformer_out, sizeof(float) * dim);
    if (!text_logits) free(logits);
    return tok;
}

int orc_step_temporal(orc_model *m, orc_state *s, const int32_t *tokens, float *text_logits, float *transformer_out) {
    const orc_config *c = &m->cfg; const int dim = c->dim;
    float *x = malloc(sizeof(float) * dim), *row = malloc(sizeof(float) * dim);
    /* lm.h:555-584: text emb first, then audio codebooks added left to right */
    if (c->demux_second_stream) embed_demux(m, &m->text_emb, &m->text_out1, &m->text_out2, c->text_card + 1, tokens[0], x);
    else embed_row(&m->text_emb, tokens[0], x);
    for (int q = 0; q < c->n_q; q++) { embed_row(&m->emb[q], tokens[q + 1], row); for (int i = 0; i < dim; i++) x[i] = x[i] + row[i]; }
    if (s->cond_sum) for (int i = 0; i < dim; i++) x[i] = s->cond_sum[i] + x[i];      /* lm.h:575-577 */
    const int tok = temporal_from_x(m, s, x, text_logits, transformer_out);
    free(x); free(row);
    return tok;
}

int orc_step_temporal_embedding(orc_model *m, orc_state *s, const float *xin, float *text_logits, float *transformer_out) {
    float *x = malloc(sizeof(float) * m->cfg.dim);
    memcpy(x, xin, sizeof(float) * m->cfg.dim);
    const int tok = temporal_from_x(m, s, x, text_logits, transformer_out);
    free(x);
    return tok;
}

void orc_step_depformer(orc_model *m, orc_state *s, int text_token, const int32_t *force, int32_t *audio_tokens, float *audio_logits) {
    const orc_config *c = &m->cfg; const int dd = c->dep_dim;
    float *y = malloc(sizeof(float) * dd), *e = malloc(sizeof(float) * dd), *logits = malloc(sizeof(float) * c->card);
    const size_t lstride = (size_t)m->dep_cap * dd;
    int prev = text_token;
    for (int k = 0; k < c->dep_q; k++) {
        int w = c->schedule_len ? c->schedule[k] : k;                 /* lm.h:457-462 */
        int wl = m->dep_num_weights == 1 ? 0 : w;                     /* transformer.h:74-83 */
        if (k == 0) {                                                  /* lm.h:494-501, scaled (-1 -> 0) */
            if (c->demux_second_stream) embed_demux(m, &m->dep_text_emb, &m->dep_text_out1, &m->dep_text_out2, c->text_card + 1, prev, e);
            else if (m->dep_text_lr.data) embed_low_rank(m, &m->dep_text_emb, &m->dep_text_lr, prev, 1, e);
            else embed_row(&m->dep_text_emb, prev, e);
        } else {                                                       /* chained get_rows (+ low_rank), lm_utils.h:209-217 */
            const orc_tensor *t = &m->dep_emb[k - 1];
            if (m->dep_emb_lr[k - 1].data) embed_low_rank(m, t, &m->dep_emb_lr[k - 1], prev, 0, e);
            else orc_dequantize_row(t->type, (const char *)t->data + (int64_t)prev * orc_row_size(t->type, t->ne0), e, t->ne0);
        }
        linear(m, &m->dep_in[m->dep_num_weights == 1 ? 0 : w], s->transformer_out, y);
        for (int i = 0; i < dd; i++) y[i] = y[i] + e[i];
        for (int l = 0; l < c->dep_layers; l++)
            transformer_layer(m, &m->dep_layers[l], wl, dd, c->dep_heads, m->dep_cap, c->dep_max_period, k,
                              s->dk + l * lstride, s->dv + l * lstride, y, NULL, 0);
        linear(m, &m->linears[k], y, logits);                          /* lm.h:472, no final norm */
        int tok = (s->temp_audio > 0.f && s->noise_audio)
                      ? sample_top_k(logits, (int)m->linears[k].ne1, s->temp_audio, s->top_k_audio,
                                     s->noise_audio + (size_t)k * (s->top_k_audio < c->card ? s->top_k_audio : c->card))
                      : argmax_first(logits, (int)m->linears[k].ne1);
        audio_tokens[k] = tok;
        if (audio_logits) memcpy(audio_logits + (size_t)k * c->card, logits, sizeof(float) * c->card);
        prev = (force && force[k] >= 0) ? force[k] : tok;
    }
    free(y); free(e); free(logits);
}

void orc_state_set_condition(orc_state *s, const float *sum, const float *cross, int tc) {
    orc_model *m = s->m; const orc_config *c = &m->cfg; const int dim = c->dim;
    free(s->cond_sum); s->cond_sum = NULL; free(s->kv_cross); s->kv_cross = NULL; s->tc = 0;
    if (sum) { s->cond_sum = malloc(sizeof(float) * dim); memcpy(s->cond_sum, sum, sizeof(float) * dim); }
    if (cross && tc > 0 && c->cross_attention) {
        /* init(): kv = in_proj rows [dim, 3*dim) applied to every condition column (transformer.h:343-396) */
        s->tc = tc;
        s->kv_cross = malloc(sizeof(float) * (size_t)c->num_layers * tc * 2 * dim);
        for (int l = 0; l < c->num_layers; l++)
            for (int i = 0; i < tc; i++)
                linear_rows(m, &m->layers[l].cross_in, dim, 2 * dim, cross + (size_t)i * dim, s->kv_cross + ((size_t)l * tc + i) * 2 * dim);
    }
}

float orc_vad(orc_model *m, orc_state *s) {
    if (m->cfg.extra_heads <= 2) return 0.f;
    const orc_tensor *t = &m->extra_heads[2]; int n = (int)t->ne1;
    float *l = malloc(sizeof(float) * n); linear(m, t, s->transformer_out, l);
    float mx = l[0]; for (int i = 1; i < n; i++) if (l[i] > mx) mx = l[i];
    double sum = 0; for (int i = 0; i < n; i++) { l[i] = expf(l[i] - mx); sum += l[i]; }
    float r = l[0] * (float)(1.0 / sum); free(l); return r;
}

/* ------------------------------------------------------------------------------------------------
 * LMGen host logic (lm.h:715-743 state, lm.h:778-979 step). Greedy, no state machine, no prefixes.
 * ---------------------------------------------------------------------------------------------- */
struct orc_lmgen {
    orc_model *m; orc_state *s;
    int offset, CT, ncb;
    int32_t *cache;     /* [CT][ncb], init -2 (lm_ungenerated_token_id) */
    int32_t initial[ORC_MAX_CODEBOOKS];
    int max_delay;
};

orc_lmgen *orc_lmgen_new(orc_model *m) {
    const orc_config *c = &m->cfg;
    orc_lmgen *g = calloc(1, sizeof(*g));
    g->m = m; g->s = orc_state_new(m); g->ncb = c->n_q + 1;
    int md = c->delays[0]; for (int i = 0; i < c->n_delays; i++) if (c->delays[i] > md) md = c->delays[i];
    g->max_delay = md;
    g->CT = md + 2 + (c->personaplex ? 1 : 0);
    g->cache = malloc(sizeof(int32_t) * g->CT * g->ncb);
    for (int i = 0; i < g->CT * g->ncb; i++) g->cache[i] = -2;
    g->initial[0] = c->text_card;
    for (int i = 1; i < g->ncb; i++) g->initial[i] = c->card;
    return g;
}
void orc_lmgen_free(orc_lmgen *g) { if (!g) return; orc_state_free(g->s); free(g->cache); free(g); }
orc_state *orc_lmgen_state(orc_lmgen *g) { return g->s; }
int orc_lmgen_offset(orc_lmgen *g) { return g->offset; }

int orc_lmgen_step(orc_lmgen *g, const int32_t *in_tokens, int n_in, int depformer_replace_tokens,
                   int32_t *out_text, int32_t *out_audio) {
    orc_model *m = g->m; const orc_config *c = &m->cfg;
    const int CT = g->CT, ncb = g->ncb;
    int dep_q = c->dep_q; if (c->personaplex) dep_q = 8;               /* lm.h:802-805 */
    const int dep_q_1 = dep_q + 1;
    const int needed = ncb - dep_q - 1;
    int provided = 0;
    if (needed > 0) {
        if (n_in == ncb) {
            for (int i = 0; i < ncb; i++) g->cache[((g->offset + c->delays[i]) % CT) * ncb + i] = in_tokens[i];
            provided = 1;
        } else {
            for (int i = 0; i < needed; i++) g->cache[((g->offset + c->delays[dep_q_1 + i]) % CT) * ncb + dep_q_1 + i] = in_tokens[i];
        }
    }
    const int pos = g->offset % CT;
    int32_t input[ORC_MAX_CODEBOOKS];
    for (int i = 0; i < ncb; i++) input[i] = (g->offset <= c->delays[i]) ? g->initial[i] : g->cache[pos * ncb + i];

    int text_token = orc_step_temporal(m, g->s, input, NULL, NULL);

    int32_t audio[ORC_MAX_STEPS];
    if (c->dep_q > 0) {
        if (!depformer_replace_tokens) orc_step_depformer(m, g->s, text_token, NULL, audio, NULL);
        else for (int i = 0; i < c->dep_q; i++) audio[i] = -1;
        if (c->delay_steps) for (int q = 0; q < c->dep_q; q++) if (g->offset < c->delays[q + 1] + c->delay_steps) audio[q] = -1;
    }
    g->offset++;
    if (!provided) {
        const int p = g->offset % CT;
        g->cache[p * ncb + 0] = text_token;
        if (c->dep_q > 0) for (int q = 0; q < c->dep_q; q++) g->cache[p * ncb + q + 1] = audio[q];
    }
    for (int q = 0; q < c->dep_q; q++) out_audio[q] = audio[q];
    if (g->offset <= g->max_delay || depformer_replace_tokens) return 0;
    *out_text = g->cache[((g->offset - g->max_delay + c->delays[0]) % CT) * ncb + 0];
    for (int i = 1; i < dep_q_1; i++) out_audio[i - 1] = g->cache[((g->offset - g->max_delay + c->delays[i]) % CT) * ncb + i];
    for (int q = 0; q < c->dep_q; q++) if (out_audio[q] == -1) return 0;
    return 1;
}
