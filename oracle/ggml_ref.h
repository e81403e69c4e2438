/*
 * oracle/ggml_ref.h — TEST INFRASTRUCTURE ONLY (the parity oracle).
 *
 * CPU restatement of the arithmetic that moshi.cpp delegates to ggml's CPU backend for
 * the per-frame LM decode step, plus the reference's own host logic around it.
 * Nothing in the product path (moshi.cpp_b200/, include/) may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / reported CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" against a real ggml build — ggml is an un-vendored,
 * version-unpinned external dependency of the reference (cmake/FindGGML.cmake:11-34,
 * README.md:183-198) and is absent from this environment; the reference ships no golden
 * vectors.  What IS pinned: block formats / dequantisation (Q4_K, Q8_0, Q4_0) and Q8_0
 * quantisation are checked bit-exact against gguf-py 0.19.0 (ggml's own python
 * implementation: gguf/quants.py), see tests/test_oracle_pins.py and tests/golden/.
 */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum orc_type {           /* numeric values follow ggml's enum ggml_type (gguf/constants.py) */
    ORC_F32 = 0, ORC_F16 = 1, ORC_Q4_0 = 2, ORC_Q8_0 = 8, ORC_Q4_K = 12, ORC_BF16 = 30,
};

#define ORC_MAX_CODEBOOKS 40
#define ORC_MAX_STEPS 40

typedef struct orc_config {
    /* mirrors moshi_config_t (include/moshi/moshi.h:111-156) — only the LM fields */
    int32_t dim, num_heads, num_layers, context, max_period;
    int32_t n_q, dep_q, card, text_card;
    int32_t dep_dim, dep_heads, dep_layers, dep_context, dep_max_period; /* 0 = pos_emb "none" */
    int32_t n_delays; int32_t delays[ORC_MAX_CODEBOOKS];
    int32_t schedule_len; int32_t schedule[ORC_MAX_STEPS]; /* depformer_weights_per_step_schedule */
    int32_t personaplex;
    int32_t extra_heads;        /* extra_heads_num_heads */
    int32_t delay_steps;        /* moshi_lm_set_delay_steps */
    /* TTS-family switches (moshi.h:111-156): */
    int32_t cross_attention;    /* temporal layers carry norm_cross + cross_attention (lm_default.h:18-34) */
    int32_t demux_second_stream;/* text embeddings are moshi_scaled_embedding_demux_t (lm_default.h:175-184, 205-211) */
    int32_t dep_low_rank;       /* depformer_low_rank_embeddings: != 0 -> depformer embeddings carry a low_rank linear */
} orc_config;

typedef struct orc_model orc_model;
typedef struct orc_state orc_state;
typedef struct orc_lmgen orc_lmgen;

/* ---- T0: block formats ---------------------------------------------------------- */
/* OpenMP threads used by the row-parallel loops (benchmark CPU legs set this explicitly) */
void orc_set_threads(int n);
int orc_max_threads(void);
int64_t orc_row_size(int type, int64_t k);                       /* bytes of one row of k elements */
void orc_dequantize_row(int type, const void *src, float *dst, int64_t k);
void orc_quantize_row_q8_0(const float *x, void *dst, int64_t k);
void orc_timestep_freq(int half, int max_period, float *out);
void orc_quantize_row_q4_0(const float *x, void *dst, int64_t k);
void orc_quantize_row_q4_K(const float *x, void *dst, int64_t k);
void orc_quantize_row_q8_K(const float *x, int8_t *qs, float *d, int16_t *bsums, int64_t k);
float orc_fp16_to_fp32(uint16_t h);
uint16_t orc_fp32_to_fp16(float f);
uint16_t orc_fp32_to_bf16(float f);
float orc_bf16_to_fp32(uint16_t h);

/* ---- T1: ggml-CPU-faithful ops -------------------------------------------------- */
/* y[rows] = W[rows][k] * x[k]; activations re-quantised like ggml's CPU mul_mat */
void orc_mul_mat_vec(int type, const void *w, int64_t k, int64_t rows, const float *x, float *y);
/* T2: same contraction on dequantised weights with double accumulation, no act. quant */
void orc_mul_mat_vec_ideal(int type, const void *w, int64_t k, int64_t rows, const float *x, float *y);
void orc_rms_norm(const float *x, const float *alpha, float eps, float *y, int64_t n);

/* ---- model / state -------------------------------------------------------------- */
orc_model *orc_model_new(const orc_config *cfg);
void orc_model_free(orc_model *m);
/* name as in the GGUF; data must outlive the model. returns 0 if name unknown */
int orc_model_set_tensor(orc_model *m, const char *name, int type, int64_t ne0, int64_t ne1, const void *data);
int orc_model_missing(orc_model *m, char *buf, int buflen);     /* #unset required tensors */
void orc_model_set_ideal(orc_model *m, int ideal);              /* 1 = T2 numerics */

orc_state *orc_state_new(orc_model *m);
void orc_state_free(orc_state *s);
void orc_state_reset(orc_state *s);
int orc_state_offset(orc_state *s);

/* one temporal step (lm.h:659-690 + transformer.h:1217-1289): tokens[n_q+1] -> logits, token */
int orc_step_temporal(orc_model *m, orc_state *s, const int32_t *tokens,
                      float *text_logits /*[text_card] or NULL*/, float *transformer_out /*[dim] or NULL*/);
/* PersonaPlex voice-embedding prompt (lm.h:694-709, 1005-1036): the same step with the embedding sum REPLACED by a given
 * f32 row x[dim] (no tokens, no condition_sum) */
int orc_step_temporal_embedding(orc_model *m, orc_state *s, const float *x, float *text_logits, float *transformer_out);
/* depformer chain (lm.h:446-553); force[k] >= 0 replaces the greedy choice that feeds step k+1 */
void orc_step_depformer(orc_model *m, orc_state *s, int text_token, const int32_t *force /*[dep_q] or NULL*/,
                        int32_t *audio_tokens /*[dep_q]*/, float *audio_logits /*[dep_q][card] or NULL*/);
/* sampling: temp <= 0 keeps greedy; noise_text[top_k_text], noise_audio[dep_q][min(top_k_audio, card)] must outlive the steps */
void orc_state_set_sampling(orc_state *s, float temp_text, float temp_audio, int top_k_text, int top_k_audio);
void orc_state_set_noise(orc_state *s, const float *noise_text, const float *noise_audio);
/* TTS conditioning (moshi.cpp:851-883, transformer.h:343-396, lm.h:575-577): sum[dim] is added to the embedding sum of
 * every frame, cross[Tc][dim] is projected ONCE through every layer's cross_attention.in_proj rows [dim, 3*dim) into
 * f32 k_cross / v_cross.  Either may be NULL. */
void orc_state_set_condition(orc_state *s, const float *sum, const float *cross, int tc);
/* STT VAD head (lm.h:966-976): softmax(extra_heads[2] . transformer_out)[0] */
float orc_vad(orc_model *m, orc_state *s);
/* debugging / parity: copy KV row */
void orc_state_get_kv(orc_state *s, int layer, int head, int slot, uint16_t *k, uint16_t *v);

/* ---- host logic: LMGen (lm.h:715-743, 778-979), greedy only --------------------- */
orc_lmgen *orc_lmgen_new(orc_model *m);
void orc_lmgen_free(orc_lmgen *g);
/* in_tokens: n_in user codes (or n_q+1 when "provided"); returns 1 when out tokens valid */
int orc_lmgen_step(orc_lmgen *g, const int32_t *in_tokens, int n_in, int depformer_replace_tokens,
                   int32_t *out_text, int32_t *out_audio /*[dep_q]*/);
orc_state *orc_lmgen_state(orc_lmgen *g);
int orc_lmgen_offset(orc_lmgen *g);

/* ---- Mimi split residual vector quantiser (mimi_rvq_ref.c; reference src/moshi/quantization/{core_vq,vq}.h) ---- */
void orc_conv1d_k1_f16(const uint16_t *w, int n_in, int n_out, const float *x, int T, float *y);
void orc_residual_vq_encode(const float *codebooks, int n_q, int bins, int D, float *x, int T, int32_t *codes);
void orc_residual_vq_decode(const float *codebooks, int n_q, int bins, int D, const int32_t *codes, int T, float *out);
void orc_split_rvq_encode(const float *cb_first, const float *cb_rest, const uint16_t *in_first, const uint16_t *in_rest, int n_sem, int n_q,
                          int bins, int D, int dim, const float *x, int T, int32_t *codes);
void orc_split_rvq_decode(const float *cb_first, const float *cb_rest, const uint16_t *out_first, const uint16_t *out_rest, int n_sem, int K,
                          int bins, int D, int dim, const int32_t *codes, int T, float *y);

#ifdef __cplusplus
}

#endif
