// moshi_api.h — C++ mirror of the LM part of the reference's public API (include/moshi/moshi.h:111-203),
// implemented on top of the C ABI in include/moshi_b200.h.  Same function names, argument meaning and
// error behaviour as the reference, so a tool written against moshi.h keeps compiling for the LM path:
//   moshi_get_config, moshi_lm_from_files, moshi_lm_quantize, moshi_lm_load, moshi_lm_save_gguf, moshi_lm_set_delay_steps,
//   moshi_lm_get_max_delay, moshi_lm_get_delay_steps, moshi_lm_generator, moshi_lm_start, moshi_lm_send2,
//   moshi_lm_receive, moshi_lm_receive2, moshi_lm_personaplex_audio_prompt, moshi_lm_personaplex_system_prompt,
//   unref(...).
//   TTS: Entry, moshi_lm_send, moshi_lm_is_active, moshi_lm_is_empty, moshi_lm_machine_reset, moshi_lm_voice_prefix,
//   and moshi_lm_set_condition (the conditioning TENSORS; the conditioners that compute them from a voice file,
//   moshi.cpp:296-366, are outside the per-frame path).
// Out of scope: Mimi codec, tokenizer (sentencepiece), voice-file loading,
// quantise-on-load — those entry points are not declared here.
#pragma once
#include <cstdint>
#include <deque>
#include <string>
#include <vector>

#if defined(MOSHI_BUILD)
#define MOSHI_API __attribute__((visibility("default"))) extern
#else
#define MOSHI_API extern
#endif

// the reference passes ggml backends to moshi_alloc (moshi.h:28); here they are opaque and ignored
struct ggml_backend;

struct moshi_context_t;
MOSHI_API moshi_context_t *moshi_alloc(ggml_backend *backend, ggml_backend *backend_cpu);
MOSHI_API moshi_context_t *moshi_alloc_b200(int cuda_device);     // addition: choose the GPU explicitly
MOSHI_API void unref(moshi_context_t *moshi);

// LM fields of the reference's moshi_config_t (moshi.h:111-156); same names and types
struct moshi_config_t {
    int64_t card = 0, n_q = 0, dep_q = 0;
    std::vector<int64_t> delays;
    int64_t dim = 0, text_card = 0, existing_text_padding_id = 3, num_heads = 0, num_layers = 0;
    float hidden_scale = 4.125f;
    bool causal = true;
    int64_t context = 0, max_period = 10000;
    std::string gating, norm, positional_embedding;
    int64_t depformer_dim = 0, depformer_num_heads = 0, depformer_num_layers = 0;
    bool depformer_multi_linear = true;
    int64_t depformer_context = 0, depformer_max_period = 0;
    std::string depformer_gating, depformer_pos_emb;
    bool depformer_weights_per_step = true;
    int64_t depformer_low_rank_embeddings = 0;
    bool demux_second_stream = false;
    bool cross_attention = false;
    int64_t extra_heads_num_heads = 0;
    std::vector<int64_t> depformer_weights_per_step_schedule;
    std::string model_type, tokenizer_name, mimi_name, moshi_name = "model.safetensors";
    struct { float audio_delay = 0.f; int64_t second_stream_ahead = 0; } tts_config;   // config_tts_t (moshi.h:90-97)
};
MOSHI_API int moshi_get_config(moshi_config_t *config, const char *filename);   // 0 ok, -1 on error (config.h:148-346)

struct moshi_lm_t;
MOSHI_API moshi_lm_t *moshi_lm_from_files(moshi_context_t *moshi, moshi_config_t *config, const char *filepath);   // NULL if the file is missing
MOSHI_API void unref(moshi_lm_t *lm);
MOSHI_API void moshi_lm_set_delay_steps(moshi_lm_t *lm, int delay_steps);
MOSHI_API int moshi_lm_get_max_delay(moshi_lm_t *lm);
MOSHI_API int moshi_lm_get_delay_steps(moshi_lm_t *lm);
MOSHI_API bool moshi_lm_quantize(moshi_lm_t *lm, const char *quant);   // "q8_0" / "q4_k": float tensors are quantised on the GPU while loading
MOSHI_API int moshi_lm_load(moshi_lm_t *lm);                           // 0 ok
MOSHI_API void moshi_lm_save_gguf(moshi_lm_t *lm, const char *filepath);  // the (quantised) weights as a GGUF (moshi.h:175)

struct moshi_lm_gen_t;
MOSHI_API moshi_lm_gen_t *moshi_lm_generator(moshi_lm_t *lm);
MOSHI_API void unref(moshi_lm_gen_t *gen);
MOSHI_API int moshi_lm_personaplex_audio_prompt(moshi_lm_gen_t *gen, std::deque<std::vector<int16_t>> &audio_prompt);   // steals the deque
// voice prompt, embedding variant: the tensors moshi_lm_personaplex_load_voice reads from a voice file ("voice.embeddings"
// as n_rows x dim f32, "voice.cache" as the token ring [CT][n_q+1] row-major; the file stores it transposed, lm.h:1047-1051)
MOSHI_API int moshi_lm_personaplex_voice_tensors(moshi_lm_gen_t *gen, const float *embeddings, int n_rows, const int32_t *cache, int cache_rows);
// moshi.cpp:789-836: a voice file (.safetensors: "embeddings" [N, 1, 1, dim] float + "cache" [n_q+1, CT] I32; .gguf:
// "voice.embeddings" / "voice.cache") -> the two tensors above.  -1 for an unknown extension or an unreadable file.
MOSHI_API int moshi_lm_personaplex_load_voice(moshi_context_t *moshi, moshi_lm_gen_t *gen, const char *filepath);
// the reference tokenises `prompt` with sentencepiece (moshi.cpp:838-849); without a tokenizer the caller passes ids
MOSHI_API int moshi_lm_personaplex_system_prompt_tokens(moshi_lm_gen_t *gen, const std::vector<int> &text_tokens);
MOSHI_API void moshi_lm_start(moshi_context_t *moshi, moshi_lm_gen_t *gen, float depth_temperature, float text_temperature, bool logging = false);
// TTS word queue entry (moshi.h:63-68)
struct Entry {
    std::vector<int> tokens;
    std::string text;
    int padding = 0;
    int64_t time = 0;
};
// conditioning tensors of the utterance: cond_sum[dim] (or NULL), cond_cross[tc][dim] (or NULL); call before moshi_lm_start.
// Marks the generator as a TTS generator (state machine on), like a loaded voice does in the reference (moshi.cpp:857-871).
MOSHI_API int moshi_lm_set_condition(moshi_lm_gen_t *gen, const float *cond_sum, const float *cond_cross, int tc);
// reference entry points that need the safetensors loader + conditioners: -1 without cross-attention, -2 otherwise (moshi.cpp:729-760)
MOSHI_API int moshi_lm_set_voice_condition(moshi_context_t *moshi, moshi_lm_gen_t *gen, const char *filepath);
MOSHI_API int moshi_lm_load_voice_condition(moshi_context_t *moshi, moshi_lm_gen_t *gen);
MOSHI_API int moshi_lm_voice_prefix(moshi_lm_gen_t *gen, std::deque<int> &text_prefix, std::deque<std::vector<int>> &audio_prefix);   // steals both deques
MOSHI_API void moshi_lm_send(moshi_lm_gen_t *gen, Entry *entry);
MOSHI_API int moshi_lm_is_active(moshi_lm_gen_t *gen);
MOSHI_API int moshi_lm_is_empty(moshi_lm_gen_t *gen);
MOSHI_API void moshi_lm_machine_reset(moshi_lm_gen_t *gen);
MOSHI_API void moshi_lm_send2(moshi_lm_gen_t *gen, std::vector<int16_t> &audio_tokens);
MOSHI_API int moshi_lm_receive(moshi_lm_gen_t *gen, int &text_token, std::vector<int16_t> &audio_tokens);
MOSHI_API void moshi_lm_receive2(moshi_lm_gen_t *gen, int &text_token, float &vad);
MOSHI_API const char *moshi_b200_last_error();

// ---- TTS text scheduling (src/moshi/models/lm.h:5-194), exposed with C linkage so that host-only tests can drive it --
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
