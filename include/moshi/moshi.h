// moshi.h — the LM part of the reference's public API (/root/reference/include/moshi/moshi.h:14-203) on the B200 engine.
//
// Source-compatible with the reference header for everything the four tools call on the LM path (moshi-sts, personaplex,
// moshi-tts, moshi-stt): same function names, argument types, struct layouts (moshi_config_t in the reference's field order,
// with config_fuser_t / config_tts_t / config_stt_t / config_model_id_t / config_lm_gen_t), same ownership rules (handles
// are released with unref(); voice_prefix / audio_prompt steal the caller's deques) and the same error returns (NULL / -1 / -2
// for missing files, false for an unknown quantisation).  Implemented in moshi.cpp_b200/host/moshi_api.cpp (libmoshi.so) on
// top of the C ABI in include/moshi_b200.h.
//
// Differences, all at the edges of the path (SURVEY.md section 2 rows 13-25 are out of scope):
//   * no <ggml.h>: `ggml_backend` is an opaque type (ggml-backend.h shim) and moshi_alloc ignores its arguments;
//   * the Mimi codec entry points (moshi.h:31-59) are not declared — nothing behind them is built;
//   * tokenizer_t is a small built-in vocabulary reader, not sentencepiece (see tokenizer_alloc);
//   * a few additions, marked "addition", that the reference does not have.
#pragma once
#include <cstdint>
#include <deque>
#include <string>
#include <vector>

#include "ggml-backend.h"
#include "ptrs.h"

#if defined(MOSHI_BUILD)
#define MOSHI_API __attribute__((visibility("default"))) extern
#else
#define MOSHI_API extern
#endif

// MARK: Moshi Context  (moshi.h:24-29)
struct moshi_context_t;
MOSHI_API moshi_context_t *moshi_alloc(ggml_backend *backend, ggml_backend *backend_cpu);
MOSHI_API moshi_context_t *moshi_alloc_b200(int cuda_device);     // addition: choose the GPU explicitly
MOSHI_API void unref(moshi_context_t *moshi);

// MARK: Tokenizer  (moshi.h:61-78)
struct Entry {
    std::vector<int> tokens;
    std::string text;
    int padding = 0;
    int64_t time = 0;
};
// The reference wraps sentencepiece (src/moshi.cpp, tokenizer_t::sp).  sentencepiece is not part of this build: tokenizer_alloc
// reads a plain vocabulary file instead — one piece per line, the line number is the token id, "▁" (or a leading space)
// marks a word start — and encodes by greedy longest match, word by word; a sentencepiece `.model` protobuf gives NULL.
struct tokenizer_t;
MOSHI_API tokenizer_t *tokenizer_alloc(const char *filepath, bool insert_bos = true);
MOSHI_API void unref(tokenizer_t *tok);
MOSHI_API bool tokenizer_empty(tokenizer_t *tok);
MOSHI_API int tokenizer_send(tokenizer_t *tok, std::string text);
MOSHI_API int tokenizer_receive(tokenizer_t *tok, Entry *entry);
MOSHI_API std::string tokenizer_id_to_piece(tokenizer_t *tok, int token);

// MARK: Config  (moshi.h:80-158; field order and defaults of the reference, src/config.h:148-346)
struct config_fuser_t {
    bool cross_attention_pos_emb = false;
    float cross_attention_pos_emb_scale = 1.f;
    std::vector<std::string> sum;
    std::vector<std::string> cross;
};
struct config_tts_t {
    float audio_delay = 0.f;
    int64_t second_stream_ahead = 0;
};
struct config_stt_t {
    float audio_delay_seconds = 5.0f;
    float audio_silence_prefix_seconds = 1.0f;
};
struct config_model_id_t {
    std::string sig;
    int64_t epoch = 0;
};
struct config_lm_gen_t {
    float temp = 0.f;
    float temp_text = 0.f;
    int64_t top_k = 0;
    int64_t top_k_text = 0;
};
struct moshi_config_t {
    int64_t card = 0;
    int64_t n_q = 0;
    int64_t dep_q = 0;
    std::vector<int64_t> delays;
    int64_t dim = 0;
    int64_t text_card = 0;
    int64_t existing_text_padding_id = 3;
    int64_t num_heads = 0;
    int64_t num_layers = 0;
    float hidden_scale = 4.125f;
    bool causal = true;
    int64_t context = 0;
    int64_t max_period = 10000;
    std::string gating;
    std::string norm;
    std::string positional_embedding;
    int64_t depformer_dim = 0;
    int64_t depformer_num_heads = 0;
    int64_t depformer_num_layers = 0;
    bool depformer_multi_linear = true;
    int64_t depformer_context = 0;
    int64_t depformer_max_period = 0;
    std::string depformer_gating;
    std::string depformer_pos_emb;
    bool depformer_weights_per_step = true;
    int64_t depformer_low_rank_embeddings = 0;
    bool demux_second_stream = false;
    config_fuser_t fuser;
    bool cross_attention = false;
    int64_t extra_heads_num_heads = 0;
    config_tts_t tts_config;
    config_stt_t stt_config;
    config_model_id_t model_id;
    std::vector<int64_t> depformer_weights_per_step_schedule;
    std::string model_type;
    config_lm_gen_t lm_gen_config;
    std::string tokenizer_name;
    std::string mimi_name;
    std::string moshi_name = "model.safetensors";
};
MOSHI_API int moshi_get_config(moshi_config_t *config, const char *filename);   // 0 ok, -1 on error

// MARK: LM  (moshi.h:160-176)
struct moshi_lm_t;
MOSHI_API moshi_lm_t *moshi_lm_from_files(moshi_context_t *moshi, moshi_config_t *config, const char *filepath);   // NULL if the file is missing
MOSHI_API void unref(moshi_lm_t *lm);
MOSHI_API void moshi_lm_set_delay_steps(moshi_lm_t *lm, int delay_steps);
MOSHI_API int moshi_lm_get_max_delay(moshi_lm_t *lm);
MOSHI_API int moshi_lm_get_delay_steps(moshi_lm_t *lm);
MOSHI_API bool moshi_lm_quantize(moshi_lm_t *lm, const char *quant);   // "q8_0" / "q4_k": float tensors are quantised on the GPU while loading
MOSHI_API int moshi_lm_load(moshi_lm_t *lm);                           // 0 ok
MOSHI_API void moshi_lm_save_gguf(moshi_lm_t *lm, const char *filepath);

// MARK: Generator  (moshi.h:178-203)
struct moshi_lm_gen_t;
MOSHI_API moshi_lm_gen_t *moshi_lm_generator(moshi_lm_t *lm);
MOSHI_API void unref(moshi_lm_gen_t *gen);
MOSHI_API int moshi_lm_set_voice_condition(moshi_context_t *moshi, moshi_lm_gen_t *gen, const char *filepath);   // -1 / -2 like moshi.cpp:729-745
MOSHI_API int moshi_lm_load_voice_condition(moshi_context_t *moshi, moshi_lm_gen_t *gen);
MOSHI_API int moshi_lm_voice_prefix(moshi_lm_gen_t *gen, std::deque<int> &text_prefix, std::deque<std::vector<int>> &audio_prefix);   // steals both deques
MOSHI_API int moshi_lm_personaplex_audio_prompt(moshi_lm_gen_t *gen, std::deque<std::vector<int16_t>> &audio_prompt);               // steals the deque
MOSHI_API int moshi_lm_personaplex_load_voice(moshi_context_t *moshi, moshi_lm_gen_t *gen, const char *filename);
MOSHI_API int moshi_lm_personaplex_system_prompt(moshi_context_t *moshi, moshi_lm_gen_t *gen, tokenizer_t *tok, const char *prompt);
MOSHI_API void moshi_lm_start(moshi_context_t *moshi, moshi_lm_gen_t *gen, float depth_temperature, float text_temperature, bool logging = false);
MOSHI_API void moshi_lm_send(moshi_lm_gen_t *gen, Entry *entry);
MOSHI_API int moshi_lm_receive(moshi_lm_gen_t *gen, int &text_token, std::vector<int16_t> &audio_tokens);
MOSHI_API void moshi_lm_send2(moshi_lm_gen_t *gen, std::vector<int16_t> &audio_tokens);
MOSHI_API void moshi_lm_receive2(moshi_lm_gen_t *gen, int &text_token, float &vad);
MOSHI_API int moshi_lm_is_active(moshi_lm_gen_t *gen);
MOSHI_API int moshi_lm_is_empty(moshi_lm_gen_t *gen);
MOSHI_API void moshi_lm_machine_reset(moshi_lm_gen_t *gen);

// MARK: additions (not in the reference)
// the system prompt as token ids, for callers that tokenise elsewhere (the reference tokenises inside, moshi.cpp:838-849)
MOSHI_API int moshi_lm_personaplex_system_prompt_tokens(moshi_lm_gen_t *gen, const std::vector<int> &text_tokens);
// the tensors moshi_lm_personaplex_load_voice reads from a voice file ("voice.embeddings" as n_rows x dim f32, "voice.cache"
// as the token ring [CT][n_q+1] row-major; the file stores it transposed, lm.h:1047-1051)
MOSHI_API int moshi_lm_personaplex_voice_tensors(moshi_lm_gen_t *gen, const float *embeddings, int n_rows, const int32_t *cache, int cache_rows);
// conditioning tensors of a TTS utterance computed elsewhere: cond_sum[dim] (or NULL), cond_cross[tc][dim] (or NULL); call
// before moshi_lm_start.  Marks the generator as a TTS generator (state machine on) like a loaded voice does (moshi.cpp:857-871).
MOSHI_API int moshi_lm_set_condition(moshi_lm_gen_t *gen, const float *cond_sum, const float *cond_cross, int tc);
MOSHI_API const char *moshi_b200_last_error();

// TTS text scheduling (src/moshi/models/lm.h:5-194) with C linkage, so that host-only tests can drive it
struct moshi_tts_machine_t;
extern "C" {
#define MOSHI_C_API __attribute__((visibility("default")))
MOSHI_C_API moshi_tts_machine_t *moshi_tts_machine_new(int text_card, int second_stream_ahead, int max_padding, int initial_padding);
MOSHI_C_API void moshi_tts_machine_free(moshi_tts_machine_t *m);
MOSHI_C_API void moshi_tts_machine_push(moshi_tts_machine_t *m, const int *tokens, int n_tokens, int padding);
MOSHI_C_API int moshi_tts_machine_process(moshi_tts_machine_t *m, int step, int token);
MOSHI_C_API int moshi_tts_machine_end_step(moshi_tts_machine_t *m);
MOSHI_C_API int moshi_tts_machine_is_empty(moshi_tts_machine_t *m);
MOSHI_C_API void moshi_tts_machine_reset(moshi_tts_machine_t *m);
}
