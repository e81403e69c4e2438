/*
 * include/moshi_b200.h — C ABI of the B200-native LM decode step ("msx").
 *
 * This is the drop-in boundary for the hot path of Codes4Fun/moshi.cpp: it replaces the ggml op
 * graphs that the reference builds and runs for one frame of the LM
 *     temporal graph   src/moshi/models/lm.h:853-875  (moshi_lmmodel_forward_text_build/_step)
 *     depformer graph  src/moshi/models/lm.h:478-553  (moshi_lmmodel_depformer_step)
 * together with the runtime objects those graphs live in
 *     GraphContext / ScratchContext / StateContext   src/context.h:227-780
 *     WeightLoader (GGUF side)                        src/loader.h:85-99, 235-271
 * The reference itself has no C ABI (its API is C++ linkage, include/moshi/moshi.h:14-22); the
 * C++ mirror of that API (moshi.cpp_b200/host/moshi_api.h: moshi_lm_*) is implemented on top of
 * the entry points below.  Plain pointers and sizes only; no C++/torch types.
 *
 * All functions return 0 on success and a negative msx_status on failure; msx_last_error() gives
 * the message for the calling thread.  There is NO CPU fallback: every compute entry point fails
 * with MSX_ERR_CUDA when no sm_100 device / kernel image is available.
 */
#ifndef MOSHI_B200_H
#define MOSHI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSX_API __attribute__((visibility("default")))

#define MSX_MAX_CODEBOOKS 40
#define MSX_MAX_STEPS 40
#define MSX_NO_TOKEN INT32_MIN /* "no override" marker for text_override / force[] */

typedef enum msx_status {
    MSX_OK = 0,
    MSX_ERR_ARG = -1,     /* bad argument / config */
    MSX_ERR_IO = -2,      /* file missing / unreadable (reference: NULL from moshi_lm_from_files, moshi.cpp:621-627) */
    MSX_ERR_FORMAT = -3,  /* not GGUF / unsupported tensor type / shape mismatch / tensor missing */
    MSX_ERR_CUDA = -4,    /* CUDA runtime error, no device, out of memory */
    MSX_ERR_STATE = -5,   /* call sequence error */
} msx_status;

/* LM fields of moshi_config_t (include/moshi/moshi.h:111-156). hidden sizes come from the weights. */
typedef struct msx_config {
    int32_t dim, num_heads, num_layers, context, max_period;
    int32_t n_q, dep_q, card, text_card;
    int32_t dep_dim, dep_heads, dep_layers, dep_context, dep_max_period; /* 0 = depformer_pos_emb "none" */
    int32_t n_delays;
    int32_t delays[MSX_MAX_CODEBOOKS];
    int32_t schedule_len;
    int32_t schedule[MSX_MAX_STEPS]; /* depformer_weights_per_step_schedule */
    int32_t personaplex;             /* model_type == "personaplex" */
    int32_t extra_heads;             /* extra_heads_num_heads */
    /* TTS family (moshi.h:111-156): */
    int32_t cross_attention;         /* temporal layers carry norm_cross + cross_attention (lm_default.h:18-34) */
    int32_t demux_second_stream;     /* two-stream text embeddings (lm_utils.h:14-125) */
    int32_t dep_low_rank;            /* depformer_low_rank_embeddings (0 = none) */
} msx_config;

typedef struct msx_model msx_model;   /* device-resident, repacked weights (one GPU)      */
typedef struct msx_stream msx_stream; /* one conversation: KV rings, offset, step graphs  */
typedef struct msx_gen msx_gen;       /* LMGen host state: token delay ring (lm.h:715-743) */

MSX_API const char *msx_last_error(void);
MSX_API int msx_device_count(void);
MSX_API const char *msx_version(void);

/* ---- weights: replaces WeightLoader::from_gguf + load_gguf + get_weights("lm.") --------------
 * (src/loader.h:85-99, 235-271; src/moshi/models/lm.h:370-395).  Tensor names / shapes / types as
 * the reference resolves them; Q4_K / Q8_0 / Q4_0 linears and Q4_0 / Q8_0 / F32 / F16 / BF16
 * embedding tables are repacked into device tiles at load. */
MSX_API int msx_model_load_gguf(const char *path, const msx_config *cfg, int device, msx_model **out);
MSX_API void msx_model_free(msx_model *m);
MSX_API int msx_model_config(const msx_model *m, msx_config *out);
/* bytes of GGUF weight blocks one frame must read (every hot-path linear once; SURVEY.md §8d "W") */
MSX_API int64_t msx_model_weight_bytes_per_frame(const msx_model *m);
MSX_API int64_t msx_model_device_bytes(const msx_model *m);
MSX_API int msx_model_device(const msx_model *m);

/* ---- per-conversation state: replaces StateContext + moshi_lmmodel_states (lm.h:423-444) ------
 * context_override > 0 shrinks the temporal ring capacity (tools' "-c N", moshi-sts.cpp:254-264). */
MSX_API int msx_stream_create(msx_model *m, int context_override, msx_stream **out);
/* flags: MSX_STREAM_STEP_KERNEL = run each stack of the frame (temporal transformer + text head; depformer chain) as ONE
 * persistent cooperative kernel (csrc/step_kernel.cuh: TMA weight ring across phases, flag-in-data activation exchange) instead
 * of PDL-chained launches: 2 launches per frame instead of 373.  Results are bit-identical.  Opt-in: on B200 the launch chain
 * is currently ~15 % faster per frame (profiles/r2_step_kernel.md); models the kernel does not take (cross-attention, tensor
 * parallel, temperature sampling, mixed weight types) silently use the launch chain. */
#define MSX_STREAM_STEP_KERNEL 1
MSX_API int msx_stream_create_ex(msx_model *m, int context_override, int flags, msx_stream **out);
MSX_API void msx_stream_free(msx_stream *s);
MSX_API int msx_stream_reset(msx_stream *s);   /* offset = 0, KV rings zeroed */
MSX_API int msx_stream_offset(const msx_stream *s);
/* bytes of KV cache the NEXT temporal step must read: n_valid * 2 * dim * 2 B * num_layers */
MSX_API int64_t msx_stream_kv_bytes_next(const msx_stream *s);

/* One temporal-transformer step (lm.h:679-690 + graph compute, lm.h:872-878).
 * tokens[n_q+1] = {text, audio codebooks}; -1 embeds as zeros, other negatives as row 0
 * (lm_utils.h:172-182).  Returns the greedy text token (sampling.h:57-63, temp <= 0).
 * text_logits (host, [text_card]) and transformer_out (host, [dim]) may be NULL. */
MSX_API int msx_step_temporal(msx_stream *s, const int32_t *tokens, int32_t *text_token,
                              float *text_logits, float *transformer_out);
/* Depformer chain for the frame (lm.h:478-553): dep_q serial codebook steps on the device.
 * force (host, [dep_q]) optionally replaces the token fed to step k+1 (MSX_NO_TOKEN / NULL = greedy).
 * audio_logits (host, [dep_q][card]) may be NULL. */
/* PersonaPlex voice-embedding prompt (lm.h:694-709): the temporal step with the embedding sum replaced by x[dim] */
MSX_API int msx_step_temporal_embedding(msx_stream *s, const float *x, int32_t *text_token, float *text_logits, float *transformer_out);
MSX_API int msx_step_depformer(msx_stream *s, int32_t text_token, const int32_t *force,
                               int32_t *audio_tokens, float *audio_logits);
/* Fused frame: temporal -> greedy text -> depformer with a single host synchronisation.
 * out_tokens[1 + dep_q] = {text, audio...}. */
MSX_API int msx_step(msx_stream *s, const int32_t *tokens, int32_t *out_tokens);
/* Sampling (moshi_sample_token, sampling.h:46-64).  temp <= 0 = greedy (the default); temp > 0 = softmax(l/temp),
 * top-k, multinomial by arg-max of p_j / e_j.  The Exp(1) draws e are an INPUT: the reference draws them on the
 * host with libc rand() (context.h:464-480); msx_gen_step does the same, tests feed fixed numbers.
 * set_sampling re-captures the stream's graphs; top_k <= 256.  set_noise must precede every sampled step:
 * noise_text[min(top_k_text, text_card)], noise_audio[dep_q][min(top_k_audio, card)], candidate order = descending p. */
MSX_API int msx_stream_set_sampling(msx_stream *s, float temp_text, float temp_audio, int top_k_text, int top_k_audio);
MSX_API int msx_stream_set_noise(msx_stream *s, const float *noise_text, const float *noise_audio);
/* STT VAD head (lm.h:966-976): softmax(extra_heads[2] . transformer_out)[0]; 0 if < 3 extra heads */
/* ---- tensor parallelism over the temporal transformer (BASELINE.json config 4) -------------
 * One process per GPU.  Rank r of `world` loads only its shard: the q/k/v rows and KV ring of heads
 * [r*H/world, (r+1)*H/world), the matching out_proj columns, a hidden slice of the gated MLP; embeddings, text head and
 * depformer are replicated.  out_proj / linear_out produce partial sums kept in DOUBLE, all-reduced with
 * ncclAllReduce(sum, f64) inside the captured graph and rounded once into the residual stream, so every rank computes
 * the single-GPU numbers.  msx_tp_unique_id on one rank, ship the 128 bytes to the others (torch.distributed, MPI, a
 * file), then msx_stream_create_tp on every rank (collective).  NCCL is bound with dlopen("libnccl.so.2"). */
MSX_API int msx_model_load_gguf_tp(const char *path, const msx_config *cfg, int device, int tp_rank, int tp_world, msx_model **out);
/* the general loader: tensor-parallel shard (rank 0 of 1 = everything) and quantise-on-load (reference:
 * moshi_lm_quantize, src/moshi.cpp:654-673; type rules src/loader.h:149-233, src/moshi/models/lm_utils.h:131-147).
 * quantize = 8 (GGML_TYPE_Q8_0): f32 / f16 / bf16 linears AND embedding tables of the file become Q8_0 on the GPU while
 * loading (ggml quantize_row_q8_0 arithmetic).  quantize = 12 (GGML_TYPE_Q4_K): linears become Q4_K (ggml
 * quantize_row_q4_K: the make_qkx2_quants scale / min search per 32-element sub-block, run with the scalar code's
 * sequential fp32 arithmetic), embedding tables and K % 256 != 0 projections Q4_0 (quantize_row_q4_0).  0 = take the
 * file as it is.  Already quantised tensors are never touched. */
MSX_API int msx_model_load_gguf_ex(const char *path, const msx_config *cfg, int device, int tp_rank, int tp_world, int quantize, msx_model **out);
/* GGUF -> GGUF quantiser: every f32 / f16 / bf16 2-D "lm." tensor of in_path is quantised on the GPU with the rules of
 * msx_model_load_gguf_ex (linears -> quantize, Q4_K needs K % 256 == 0 else Q4_0, needs K % 32 == 0 else unchanged; embedding
 * tables of a q4_k model -> Q4_0), everything else is copied; out_path is GGUF v3 without key/value pairs, like the
 * reference's own files.  Replaces `-q <quant> -g out.gguf` (moshi_lm_quantize + moshi_lm_save_gguf, src/moshi.cpp:654-695,
 * WeightLoader::save_gguf src/loader.h:227-233). */
MSX_API int msx_gguf_quantize(const char *in_path, const char *out_path, int quantize, int device);
/* the same from the reference's own starting point, a model.safetensors with torch names (bf16 / f16 / f32): tensors are
 * renamed "lm." + name, *.in_proj_weight / *.out_proj.weight are split into the per-step *.in_projs.{i}.weight /
 * *.out_projs.{i}.weight the loader expects, vectors (norm alpha, biases) become F32, 2-D tensors follow the quantisation
 * rules above (WeightLoader::from_safetensor + fetch + save_gguf, src/loader.h:77-83, 149-233;
 * src/moshi/modules/transformer.h:764-849). */
MSX_API int msx_safetensors_to_gguf(const char *in_path, const char *out_path, int quantize, int device);
MSX_API int msx_tp_unique_id(uint8_t *out128);
MSX_API int msx_stream_create_tp(msx_model *model, int context_override, const uint8_t *nccl_id128, msx_stream **out);
/* Fused GEMV -> all-reduce over peer memory (replaces the NCCL launches of a tensor-parallel stream): every rank exports
 * a 64-byte CUDA-IPC handle of its inbox arena, the handles are exchanged out of band, and msx_stream_tp_connect maps the
 * peers and re-captures the graphs: out_proj / linear_out push their fp64 partial sums straight into every rank's inbox
 * over NVLink from the GEMV epilogue, the last CTA publishes an epoch flag (st.release.sys), and the residual kernel
 * waits for the flags and adds the rows in rank order.  Same numbers as the NCCL path and as one GPU. */
MSX_API int msx_stream_tp_export(msx_stream *s, uint8_t *handle64);
MSX_API int msx_stream_tp_connect(msx_stream *s, const uint8_t *handles /* [world][64] */);

/* TTS conditioning (reference: moshi_lm_start -> init(), moshi.cpp:851-883; transformer.h:343-396; lm.h:575-577).
 * cond_sum[dim] (or NULL) is added to the embedding sum of every frame; cond_cross[tc][dim] (or NULL) is projected once
 * through every layer's cross_attention.in_proj rows [dim, 3*dim) into the f32 k_cross / v_cross memory.  The
 * conditioners that PRODUCE these tensors (src/moshi.cpp:296-366) are outside the per-frame path. */
MSX_API int msx_stream_set_condition(msx_stream *s, const float *cond_sum, const float *cond_cross, int tc);
/* the conditioners themselves (voice_condition, src/moshi.cpp:296-366; weights tts.h:16-35, optional in the GGUF):
 * cond_sum = cfg.output_proj . cfg.embed[2] + control.output_proj . control.embed[0]; cond_cross [5 * frames][dim] = the
 * speaker embedding projected by speaker_wavs.output_proj in rows [0, frames), speaker_wavs.learnt_padding in the rest,
 * plus ggml_timestep_embedding(row, dim, 10000); then msx_stream_set_condition.  speaker_wavs: the voice file's tensor as
 * stored, [channels][frames] f32.  sum_out [dim] / cross_out [5 * frames][dim] (host, nullable) receive the tensors.
 * MSX_ERR_STATE without cross-attention (reference: -1) or without conditioner tensors (reference: -2). */
MSX_API int msx_stream_set_voice(msx_stream *s, const float *speaker_wavs, int channels, int frames, float *sum_out, float *cross_out);
/* moshi_lm_set_voice_condition + moshi_lm_load_voice_condition (src/moshi.cpp:729-760): reads "speaker_wavs" ([1,] channels,
 * frames; f32 / f16 / bf16) from a voice .safetensors and calls msx_stream_set_voice */
MSX_API int msx_stream_load_voice(msx_stream *s, const char *safetensors_path);
MSX_API int msx_model_has_conditioners(const msx_model *model);   /* 1 when the GGUF carried the conditioner tensors */
MSX_API int msx_vad(msx_stream *s, float *vad);

/* Device-resident replay for throughput measurement: frames[n_frames][n_q+1] (host) are uploaded
 * once, then n_steps fused frames run back to back with no host round trip; step i consumes
 * frames[i % n_frames].  elapsed_ms (may be NULL) is measured with CUDA events on the stream the
 * kernels run on.  out_tokens (host, [n_steps][1+dep_q]) may be NULL. */
MSX_API int msx_run_resident(msx_stream *s, const int32_t *frames, int n_frames, int n_steps,
                             int32_t *out_tokens, float *elapsed_ms);
/* Several streams on ONE GPU: enqueue without waiting, then join.  elapsed_ms = device time of this stream's run. */
MSX_API int msx_run_resident_async(msx_stream *s, const int32_t *frames, int n_frames, int n_steps);
MSX_API int msx_stream_wait(msx_stream *s, float *elapsed_ms);
/* kernels launched per fused frame (for bench.py "gpu_launches") */
MSX_API int msx_stream_launches_per_frame(const msx_stream *s);
/* the resident loop with one event between the two graphs of a frame: total device milliseconds of the temporal stack and of the
 * depformer over n_steps frames of the real pipelined run */
MSX_API int msx_run_resident_split(msx_stream *s, const int32_t *frames, int n_frames, int n_steps, float *temporal_ms, float *depformer_ms);
/* Measurement aid: runs ONE fused frame eagerly (no CUDA graph) with a CUDA event recorded on the
 * launching stream after every kernel, and returns the summed device time and launch count per
 * kernel family (msx_family_name(i), i < msx_family_count()).  Results are identical to msx_step. */
MSX_API int msx_profile_frame(msx_stream *s, const int32_t *tokens, int32_t *out_tokens,
                              float *family_ms, int32_t *family_launches, int max_families);
/* CUDA-event stopwatch on the stream's own CUDA stream: start .. stop spans everything enqueued between them */
MSX_API int msx_timer_start(msx_stream *s);
MSX_API int msx_timer_stop(msx_stream *s, float *elapsed_ms);
/* one frame with the step kernels' in-kernel timeline: rows [n_phases][n_cta][9] = {family, start, end, after prologue, after main loop, input loaded, rms scale known, epilogue stored, -} ns */
MSX_API int msx_step_timeline(msx_stream *s, const int32_t *tokens, long long *rows, int max_rows, int *n_phases, int *n_cta);
MSX_API int msx_family_count(void);
MSX_API const char *msx_family_name(int i);
/* KV read-back for parity tests: bf16 bits of K and V for (layer, head, slot), Dh values each */
MSX_API int msx_stream_get_kv(msx_stream *s, int layer, int head, int slot, uint16_t *k, uint16_t *v);

/* ---- LMGen: the reference's per-frame host logic (moshi_lmgen_step, lm.h:778-979) --------------
 * Token delay ring, initial tokens, delayed emission; greedy sampling; no TTS state machine.
 * in_tokens: n_in user codes (n_q - dep_q of them, or n_q + 1 when all streams are "provided").
 * Returns 1 when out_text / out_audio[dep_q] are valid, 0 during warm-up (offset <= max_delay),
 * negative msx_status on error. */
MSX_API int msx_gen_create(msx_stream *s, int delay_steps, msx_gen **out);
/* Same host logic over a caller-supplied model step (used by the CPU-only host-logic tests):
 * fn(user, tokens[n_q+1], depformer_replace_tokens, out[1+dep_q]) fills {text, audio...}, returns 0. */
typedef int (*msx_step_fn)(void *user, const int32_t *tokens, int depformer_replace_tokens, int32_t *out);
MSX_API int msx_gen_create_with_callback(const msx_config *cfg, int delay_steps, msx_step_fn fn, void *user, msx_gen **out);
MSX_API void msx_gen_free(msx_gen *g);
/* seed of this generator's Exp(1) draws (sampled generation).  Every generator owns a private glibc random state (random_r)
 * that yields the sequence srand(seed) / rand() would: one generator reproduces the reference's draws (which never seeds
 * except in --bench: srand(0)); several generators or threads do not disturb each other. */
MSX_API void msx_gen_seed(msx_gen *g, unsigned seed);
/* T prompt frames with all n_q+1 tokens given (rows [T][n_q+1]) as one batched-T prefill: delay-ring bookkeeping of
 * moshi_lmgen_step's "provided" branch + msx_stream_prefill; same final state as T msx_gen_step calls with n_in == n_q+1 */
MSX_API int msx_gen_prefill(msx_gen *g, const int32_t *rows, int n_frames);
/* PersonaPlex voice prompt, embedding variant (moshi_lmgen_step_voice_prompt, lm.h:1005-1050): replay one embedding
 * row (text forced to 3, depformer run, offset++); afterwards install the voice's token ring, cache[CT][n_q+1] with
 * CT = msx_gen_cache_rows() */
MSX_API int msx_gen_prompt_embedding(msx_gen *g, const float *row);
MSX_API int msx_gen_set_cache(msx_gen *g, const int32_t *cache);
MSX_API int msx_gen_cache_rows(const msx_gen *g);
/* TTS hooks of the generator.  text hook: called between the temporal and the depformer graph with the sampled text
 * token, returns the token to use instead (state machine / text prefix, lm.h:877-899).  audio hook: may overwrite the
 * dep_q audio tokens of this frame (audio prefix, lm.h:922-931); returns the number of following frames whose output
 * is to be swallowed (state->skip), or -1 to leave it alone. */
typedef int32_t (*msx_text_hook)(void *user, int32_t offset, int32_t text_token);
typedef int (*msx_audio_hook)(void *user, int32_t offset, int32_t *audio_tokens, int n);
MSX_API int msx_gen_set_text_hook(msx_gen *g, msx_text_hook fn, void *user);
MSX_API int msx_gen_set_audio_hook(msx_gen *g, msx_audio_hook fn, void *user);
MSX_API int msx_gen_step(msx_gen *g, const int32_t *in_tokens, int n_in, int depformer_replace_tokens,
                         int32_t *out_text, int32_t *out_audio);
MSX_API int msx_gen_offset(const msx_gen *g);
MSX_API int msx_gen_max_delay(const msx_gen *g);

/* ---- unit-level entry points (parity tests of single kernels) -----------------------------------
 * Repack one row-major GGUF tensor [rows][k] of `type` (ggml type id), run y = W.x on the device
 * with the same fused kernels the step uses.  prologue: 0 = quantise x, 1 = rms_norm(x)*alpha then
 * quantise.  All pointers are host pointers. */
/* Batched-T prompt prefill: T frames whose n_q+1 tokens are all given (PersonaPlex voice / system
 * prompt rows, lm.h:983-1134) are run 8 positions at a time as the 8 columns of the tensor-core GEMM.  Only the temporal
 * KV rings and the position advance — exactly what T "provided" steps leave behind (their logits, sampled tokens and
 * depformer output are discarded, lm.h:933-943).  tokens [T][n_q+1].  64 positions (q4_k models with tensor-core layouts; 8
 * otherwise) go through each weight pass, anywhere on the ring: the K / V rows of a pass are inserted first and the old rows of
 * the slots they overwrite are kept aside for the columns that still see them (the T > 1 window of torch.h:170-223 with
 * serial-step semantics), so a prompt may be longer than the ring. */
MSX_API int msx_stream_prefill(msx_stream *s, const int32_t *tokens, int n_frames);
/* profiling tool: tokens [1 + pass][n_q+1] = one prompt row, then ONE full prefill pass launched eagerly with an event after every
 * launch -> per-family kernel time (family order of msx_family_name) */
MSX_API int msx_stream_prefill_profile(msx_stream *s, const int32_t *tokens, float *family_ms, int32_t *family_launches, int max_families);

/* ---- lock-step batch of independent streams (BASELINE.json config 5) ------------------
 * n_streams conversations share every weight read: one activation-quantisation launch and one
 * tensor-core dequant-GEMM launch per linear layer serve all of them; KV rings, positions and delay state
 * stay private, so stream i of a batch computes exactly what a single msx_stream would.  q4_k and q8_0 models
 * (q8_0: K % 128 == 0); not the TTS-family layers.  1..8 streams: mma.sync GEMM; 9..64 streams: tcgen05 kind::i8
 * GEMM with 16 / 32 / 64 columns (q4_k models whose every linear has rows % 128 == 0 and K % 256 == 0 — the 7B
 * family does; MSX_ERR_ARG otherwise).  Each stream owns a full KV ring (7B, 3000 slots: 1.57 GB). */
typedef struct msx_batch msx_batch;
MSX_API int msx_batch_create(msx_model *model, int n_streams, int context_override, msx_batch **out);
MSX_API void msx_batch_free(msx_batch *b);
MSX_API int msx_batch_size(const msx_batch *b);
MSX_API int msx_batch_offset(const msx_batch *b, int stream);
MSX_API int msx_batch_launches_per_frame(const msx_batch *b);
MSX_API int64_t msx_batch_kv_bytes_next(const msx_batch *b);
/* stream < 0: every stream; otherwise restart one stream (KV cleared, position 0) while the others go on */
MSX_API int msx_batch_reset_stream(msx_batch *b, int stream);
/* tokens [n][n_q+1] -> out_tokens [n][1+dep_q] (greedy), one fused frame for every stream */
MSX_API int msx_batch_step(msx_batch *b, const int32_t *tokens, int32_t *out_tokens);
/* sampling for every stream of the batch (sampling.h:4-64): softmax(l / temp) -> top-k -> arg-max of p / Exp(1); the Exp(1)
 * draws of the next frame come from the host: noise_text [n][kt], noise_audio [n][dep_q][ka], kt / ka = min(top_k, card, 256) */
MSX_API int msx_batch_set_sampling(msx_batch *b, float temp_text, float temp_audio, int top_k_text, int top_k_audio);
MSX_API int msx_batch_set_noise(msx_batch *b, const float *noise_text, const float *noise_audio);
MSX_API int msx_batch_get_logits(msx_batch *b, int stream, float *text_logits, float *audio_logits);
/* frames [n][n_frames][n_q+1] copied to the device once, n_steps frames replayed; out_tokens [n][n_steps][1+dep_q] or NULL */
MSX_API int msx_batch_run_resident(msx_batch *b, const int32_t *frames, int n_frames, int n_steps, int32_t *out_tokens, float *elapsed_ms);
MSX_API int msx_batch_profile_frame(msx_batch *b, const int32_t *tokens, float *family_ms, int32_t *family_launches, int max_families);
/* Batched generator: the LMGen host logic (token delay ring, delayed emit; lm.h:715-743, 778-979) for every stream of a batch
 * around ONE msx_batch_step.  in_tokens [n][n_in] -> out_text [n], out_audio [n][dep_q], valid [n]. */
typedef struct msx_bgen msx_bgen;
MSX_API int msx_bgen_create(msx_batch *b, int delay_steps, msx_bgen **out);
MSX_API void msx_bgen_free(msx_bgen *g);
MSX_API int msx_bgen_offset(const msx_bgen *g, int stream);
MSX_API int msx_bgen_reset_stream(msx_bgen *g, int stream);     /* a new conversation takes over this slot */
MSX_API int msx_bgen_step(msx_bgen *g, const int32_t *in_tokens, int n_in, int32_t *out_text, int32_t *out_audio, int32_t *valid);
/* test / measurement hooks of the batched GEMM */
MSX_API int msx_test_gemm_batch(int device, int type, const void *w, int64_t k, int64_t rows, const float *x, int nb, const float *alpha, float *y);
MSX_API int msx_bench_gemm_batch(int device, const void *w, int64_t k, int64_t rows, int nb, int n_mats, int iters, int epilogue, int with_quant,
                                 float *avg_us);
/* same, plus per-CTA globaltimer stamps [iters][148][8]: entry, ring primed, predecessor done, image landed, main loop done */
MSX_API int msx_bench_gemm_batch_ex(int device, const void *w, int64_t k, int64_t rows, int nb, int n_mats, int iters, int epilogue, int with_quant,
                                    float *avg_us, long long *stamps_out);

MSX_API int msx_test_gemv(int device, int type, const void *w, int64_t k, int64_t rows,
                          const float *x, const float *alpha, int prologue, float *y);
/* GEMV micro-benchmark: back-to-back launches over n_mats rotating copies of the matrix; average us per launch */
MSX_API int msx_bench_gemv(int device, int type, const void *w, int64_t k, int64_t rows, int n_mats, int iters,
                           int prologue, int epilogue, float *avg_us);
/* bit-exact dequantisation through the device embedding-gather path: out[n_rows][k] */
MSX_API int msx_test_dequant_rows(int device, int type, const void *table, int64_t k, int64_t table_rows,
                                  const int32_t *row_ids, int n_rows, float *out);
/* device Q4_K -> f32 of the REPACKED tiles (checks the repack is lossless) */
/* the on-load quantisers alone: rows x k values of GGML type src_type (0 f32, 1 f16, 30 bf16) -> GGUF blocks of dst_type
 * (8 Q8_0, 2 Q4_0, 12 Q4_K) in `out` (host) */
MSX_API int msx_test_quantize_rows(int device, int src_type, int dst_type, const void *x, int64_t k, int64_t rows, void *out);
MSX_API int msx_test_dequant_repacked(int device, int type, const void *w, int64_t k, int64_t rows, float *out);

/* ---- Mimi codec, first slice: the split residual vector quantiser (codes <-> latent) ---------------------------
 * Replaces mimi_quantizer_encode / mimi_decode_latent (reference src/moshi/models/compression.h:93-99, 216-222) and the
 * graphs under them (src/moshi/quantization/vq.h:18-117, core_vq.h:14-193; 1 x 1 convolutions torch.h:18-37).
 * cb_first [n_sem][bins][D] f32, cb_rest [n_rest][bins][D] f32 (EuclideanCodebook embeddings); in_* [D][dim], out_* [dim][D]
 * as F16 bit patterns (ggml_conv_1d kernels).  Codes are bit-exact against the CPU oracle (ggml's distance arithmetic and
 * summation order).  The rest of Mimi (SEANet, codec transformers, resampling convolutions) is not built. */
typedef struct msx_rvq msx_rvq;
MSX_API int msx_rvq_create(int device, int n_sem, int n_rest, int bins, int D, int dim, const float *cb_first, const float *cb_rest,
                           const uint16_t *in_first, const uint16_t *in_rest, const uint16_t *out_first, const uint16_t *out_rest, msx_rvq **out);
MSX_API void msx_rvq_free(msx_rvq *q);
/* x [T][dim] (host) -> codes [n_q][T] (host); n_q <= n_sem + n_rest */
MSX_API int msx_rvq_encode(msx_rvq *q, const float *x, int T, int n_q, int32_t *codes);
/* codes [K][T] (host, each in [0, bins)) -> latent [T][dim] (host) */
MSX_API int msx_rvq_decode(msx_rvq *q, const int32_t *codes, int K, int T, float *y);
/* device-timed repeats on resident buffers: milliseconds per encode / decode of T frames with n_q codebooks */
MSX_API int msx_rvq_bench(msx_rvq *q, int T, int n_q, int reps, float *encode_ms, float *decode_ms);

#ifdef __cplusplus
}
#endif
#endif /* MOSHI_B200_H */
