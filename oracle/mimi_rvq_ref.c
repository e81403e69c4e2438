// mimi_rvq_ref.c — TEST-ONLY CPU restatement of the Mimi split residual vector quantiser (first slice: the
// codes <-> latent boundary either side of the LM step).  Parity unpinned like the rest of the oracle (ggml is not in the image);
// the ggml numerics are restated from the graph the reference builds:
//   moshi_EuclideanCodebook_encode  src/moshi/quantization/core_vq.h:28-55   c = sum_rows((b - a)^2); argmax(1 / (c + 1))
//   moshi_residual_vq_encode/decode core_vq.h:136-193                          residual -= codebook[idx]; sum of rows in layer order
//   moshi_rvq_encode/decode         src/moshi/quantization/vq.h:18-47          input_proj / output_proj = conv1d, kernel 1, no bias
//   moshi_split_rvq_encode/decode   vq.h:69-117                                rvq_first (n_q_semantic layers) | rvq_rest
//   torch_nn_conv1d                 src/torch.h:18-37                          ggml_conv_1d: F16 kernel, im2col rounds x to F16
// ggml ops: sub / mul / add / div are f32 element-wise; sum_rows accumulates a row in double in index order and rounds once
// (ggml_vec_sum_f32_ggf); argmax returns the FIRST maximum; mul_mat of F16 x F16 = exact products (restated as a double sum, like
// the other F16 mat-vecs of this oracle: ggml_ref.c orc_mul_mat_vec).
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include "ggml_ref.h"

// y[t][o] = sum_i f16(w[o][i]) * f16(x[t][i])      (conv1d, kernel size 1)
void orc_conv1d_k1_f16(const uint16_t *w, int n_in, int n_out, const float *x, int T, float *y) {
#pragma omp parallel for
    for (int t = 0; t < T; t++)
        for (int o = 0; o < n_out; o++) {
            double acc = 0.0;
            for (int i = 0; i < n_in; i++) {
                const float xv = orc_fp16_to_fp32(orc_fp32_to_fp16(x[(size_t)t * n_in + i]));
                acc += (double)orc_fp16_to_fp32(w[(size_t)o * n_in + i]) * (double)xv;
            }
            y[(size_t)t * n_out + o] = (float)acc;
        }
}

// nearest centroid of every row of x [T][D] in codebook [bins][D] (core_vq.h:28-55)
static int nearest(const float *cb, int bins, int D, const float *a) {
    int best = 0; float best_r = -INFINITY;
    for (int j = 0; j < bins; j++) {
        double s = 0.0;
        for (int d = 0; d < D; d++) { const float diff = cb[(size_t)j * D + d] - a[d]; const float sq = diff * diff; s += (double)sq; }
        const float c = (float)s;
        const float r = 1.0f / (c + 1.0f);
        if (r > best_r) { best_r = r; best = j; }          // first maximum
    }
    return best;
}

// residual VQ over n_q layers: codes [n_q][T]; x [T][D] is consumed (left holding the final residual)
void orc_residual_vq_encode(const float *codebooks, int n_q, int bins, int D, float *x, int T, int32_t *codes) {
    for (int q = 0; q < n_q; q++) {
        const float *cb = codebooks + (size_t)q * bins * D;
#pragma omp parallel for
        for (int t = 0; t < T; t++) {
            float *a = x + (size_t)t * D;
            const int j = nearest(cb, bins, D, a);
            codes[(size_t)q * T + t] = j;
            for (int d = 0; d < D; d++) a[d] = a[d] - cb[(size_t)j * D + d];
        }
    }
}

// sum of the centroids of layers [0, n_q) in layer order: out [T][D]
void orc_residual_vq_decode(const float *codebooks, int n_q, int bins, int D, const int32_t *codes, int T, float *out) {
    for (int t = 0; t < T; t++)
        for (int d = 0; d < D; d++) {
            float s = 0.f;
            for (int q = 0; q < n_q; q++) {
                const float v = codebooks[((size_t)q * bins + codes[(size_t)q * T + t]) * D + d];
                s = q == 0 ? v : s + v;
            }
            out[(size_t)t * D + d] = s;
        }
}

// mimi_quantizer_encode (compression.h:216-222): latent x [T][dim] -> codes [n_q][T]; layer 0.. n_sem-1 from rvq_first, the rest from rvq_rest
void orc_split_rvq_encode(const float *cb_first, const float *cb_rest, const uint16_t *in_first, const uint16_t *in_rest, int n_sem, int n_q,
                          int bins, int D, int dim, const float *x, int T, int32_t *codes) {
    float *p = (float *)malloc((size_t)T * D * sizeof(float));
    orc_conv1d_k1_f16(in_first, dim, D, x, T, p);
    orc_residual_vq_encode(cb_first, n_sem, bins, D, p, T, codes);
    if (n_q > n_sem) {
        orc_conv1d_k1_f16(in_rest, dim, D, x, T, p);
        orc_residual_vq_encode(cb_rest, n_q - n_sem, bins, D, p, T, codes + (size_t)n_sem * T);
    }
    free(p);
}

// mimi_decode_latent (compression.h:93-99): codes [K][T] -> latent [T][dim]
void orc_split_rvq_decode(const float *cb_first, const float *cb_rest, const uint16_t *out_first, const uint16_t *out_rest, int n_sem, int K,
                          int bins, int D, int dim, const int32_t *codes, int T, float *y) {
    float *q = (float *)malloc((size_t)T * D * sizeof(float));
    const int k1 = K < n_sem ? K : n_sem;
    orc_residual_vq_decode(cb_first, k1, bins, D, codes, T, q);
    orc_conv1d_k1_f16(out_first, D, dim, q, T, y);
    if (K > n_sem) {
        float *y2 = (float *)malloc((size_t)T * dim * sizeof(float));
        orc_residual_vq_decode(cb_rest, K - n_sem, bins, D, codes + (size_t)n_sem * T, T, q);
        orc_conv1d_k1_f16(out_rest, D, dim, q, T, y2);
        for (size_t i = 0; i < (size_t)T * dim; i++) y[i] = y[i] + y2[i];
        free(y2);
    }
    free(q);
}
