// moshi_api.cpp — see moshi_api.h.  Mirrors src/moshi.cpp:600-953 (LM + generator) over the msx C ABI.
// Two pieces below are integer host logic that has to match the reference token for token and therefore follows it statement by
// statement (a restatement, not an independent design): moshi_tts_machine_t::process <- StateMachine::process (lm.h:104-193), and the
// PersonaPlex PROMPT_TOKENS table <- lm.h:983-987.  Everything else (JSON reader, voice loader, prompt batching, handles) is original.
#include "moshi_api.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <unistd.h>
#include <deque>
#include <unordered_map>

#include "../../include/moshi_b200.h"
#include "../csrc/gguf_file.h"
#include "../csrc/safetensors_file.h"

// ---- context -------------------------------------------------------------------------------------------
struct moshi_context_t { int device = 0; };
moshi_context_t *moshi_alloc(ggml_backend *, ggml_backend *) { return new moshi_context_t{0}; }
moshi_context_t *moshi_alloc_b200(int cuda_device) { return new moshi_context_t{cuda_device}; }
void unref(moshi_context_t *m) { delete m; }
const char *moshi_b200_last_error() { return msx_last_error(); }

// ---- config: minimal JSON reader for the flat keys of config.json (reference: src/config.h:148-346) -----
namespace {
struct J {
    const char *p, *e;
    void ws() { while (p < e && isspace((unsigned char)*p)) p++; }
    bool lit(const char *s) { size_t n = strlen(s); if ((size_t)(e - p) >= n && !strncmp(p, s, n)) { p += n; return true; } return false; }
    bool str(std::string &out) {
        ws(); if (p >= e || *p != '"') return false; p++; out.clear();
        while (p < e && *p != '"') { if (*p == '\\' && p + 1 < e) p++; out.push_back(*p++); }
        if (p >= e) return false; p++; return true;
    }
    bool num(double &v) { ws(); char *end = nullptr; v = strtod(p, &end); if (end == p) return false; p = end; return true; }
    bool skip() {   // any value
        ws(); if (p >= e) return false;
        if (*p == '"') { std::string s; return str(s); }
        if (*p == '{' || *p == '[') {
            char open = *p, close = open == '{' ? '}' : ']'; p++; ws();
            if (p < e && *p == close) { p++; return true; }
            while (p < e) {
                if (open == '{') { std::string k; if (!str(k)) return false; ws(); if (p >= e || *p != ':') return false; p++; }
                if (!skip()) return false; ws();
                if (p < e && *p == ',') { p++; continue; }
                if (p < e && *p == close) { p++; return true; }
                return false;
            }
            return false;
        }
        if (lit("true") || lit("false") || lit("null")) return true;
        double d; return num(d);
    }
};
}  // namespace

int moshi_get_config(moshi_config_t *c, const char *filename) {
    FILE *f = fopen(filename, "rb");
    if (!f) { fprintf(stderr, "error: failed to open %s\n", filename); return -1; }
    std::string raw; char buf[4096]; size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) raw.append(buf, n);
    fclose(f);
    if (raw.empty()) { fprintf(stderr, "error: empty file %s\n", filename); return -1; }
    *c = moshi_config_t();
    J j{raw.data(), raw.data() + raw.size()};
    j.ws();
    if (j.p >= j.e || *j.p != '{') { fprintf(stderr, "error: did not find expected json object"); return -1; }
    j.p++;
    auto i64 = [&](int64_t &dst) { j.ws(); if (j.lit("null")) return true; double d; if (!j.num(d)) return false; dst = (int64_t)d; return true; };
    auto boolean = [&](bool &dst) { j.ws(); if (j.lit("true")) { dst = true; return true; } if (j.lit("false")) { dst = false; return true; } return j.lit("null"); };
    auto arr = [&](std::vector<int64_t> &dst) {
        j.ws(); dst.clear();
        if (j.lit("null")) return true;
        if (j.p >= j.e || *j.p != '[') return false; j.p++; j.ws();
        if (j.p < j.e && *j.p == ']') { j.p++; return true; }
        while (j.p < j.e) { double d; if (!j.num(d)) return false; dst.push_back((int64_t)d); j.ws();
            if (*j.p == ',') { j.p++; continue; } if (*j.p == ']') { j.p++; return true; } return false; }
        return false;
    };
    auto string_or_null = [&](std::string &dst) { j.ws(); if (j.lit("null")) { dst.clear(); return true; } return j.str(dst); };
    auto f32 = [&](float &dst) { j.ws(); if (j.lit("null")) return true; double d = 0; if (!j.num(d)) return false; dst = (float)d; return true; };
    auto str_arr = [&](std::vector<std::string> &dst) {
        j.ws(); dst.clear();
        if (j.lit("null")) return true;
        if (j.p >= j.e || *j.p != '[') return false; j.p++; j.ws();
        if (j.p < j.e && *j.p == ']') { j.p++; return true; }
        while (j.p < j.e) { std::string v; if (!j.str(v)) return false; dst.push_back(v); j.ws();
            if (j.p < j.e && *j.p == ',') { j.p++; continue; } if (j.p < j.e && *j.p == ']') { j.p++; return true; } return false; }
        return false;
    };
    // a nested object: `field(key)` consumes the value of one key
    auto object = [&](auto field) {
        j.ws();
        if (j.lit("null")) return true;
        if (j.p >= j.e || *j.p != '{') return false;
        j.p++;
        while (true) {
            j.ws();
            if (j.p < j.e && *j.p == '}') { j.p++; return true; }
            std::string kk;
            if (!j.str(kk)) return false;
            j.ws(); if (j.p >= j.e || *j.p != ':') return false; j.p++;
            if (!field(kk)) return false;
            j.ws();
            if (j.p < j.e && *j.p == ',') j.p++;
        }
    };
    while (true) {
        j.ws();
        if (j.p < j.e && *j.p == '}') break;
        std::string k;
        if (!j.str(k)) { fprintf(stderr, "error: reading config %s\n", filename); return -1; }
        j.ws(); if (j.p >= j.e || *j.p != ':') return -1; j.p++;
        bool ok = true;
        if (k == "card") ok = i64(c->card); else if (k == "n_q") ok = i64(c->n_q); else if (k == "dep_q") ok = i64(c->dep_q);
        else if (k == "delays") ok = arr(c->delays); else if (k == "dim") ok = i64(c->dim); else if (k == "text_card") ok = i64(c->text_card);
        else if (k == "existing_text_padding_id") ok = i64(c->existing_text_padding_id);
        else if (k == "num_heads") ok = i64(c->num_heads); else if (k == "num_layers") ok = i64(c->num_layers);
        else if (k == "hidden_scale") { double d = 0; j.ws(); ok = j.num(d); c->hidden_scale = (float)d; }
        else if (k == "causal") ok = boolean(c->causal); else if (k == "context") ok = i64(c->context); else if (k == "max_period") ok = i64(c->max_period);
        else if (k == "gating") ok = string_or_null(c->gating); else if (k == "norm") ok = string_or_null(c->norm);
        else if (k == "positional_embedding") ok = string_or_null(c->positional_embedding);
        else if (k == "depformer_dim") ok = i64(c->depformer_dim); else if (k == "depformer_num_heads") ok = i64(c->depformer_num_heads);
        else if (k == "depformer_num_layers") ok = i64(c->depformer_num_layers); else if (k == "depformer_multi_linear") ok = boolean(c->depformer_multi_linear);
        else if (k == "depformer_context") ok = i64(c->depformer_context); else if (k == "depformer_max_period") ok = i64(c->depformer_max_period);
        else if (k == "depformer_gating") ok = string_or_null(c->depformer_gating); else if (k == "depformer_pos_emb") ok = string_or_null(c->depformer_pos_emb);
        else if (k == "depformer_weights_per_step") ok = boolean(c->depformer_weights_per_step);
        else if (k == "depformer_low_rank_embeddings") ok = i64(c->depformer_low_rank_embeddings);
        else if (k == "demux_second_stream") ok = boolean(c->demux_second_stream); else if (k == "cross_attention") ok = boolean(c->cross_attention);
        else if (k == "extra_heads_num_heads") ok = i64(c->extra_heads_num_heads);
        else if (k == "depformer_weights_per_step_schedule") ok = arr(c->depformer_weights_per_step_schedule);
        else if (k == "model_type") ok = string_or_null(c->model_type); else if (k == "tokenizer_name") ok = string_or_null(c->tokenizer_name);
        else if (k == "mimi_name") ok = string_or_null(c->mimi_name); else if (k == "moshi_name") ok = string_or_null(c->moshi_name);
        else if (k == "tts_config")                                // config_tts_parse (config.h:54-70)
            ok = object([&](const std::string &kk) {
                if (kk == "second_stream_ahead") return i64(c->tts_config.second_stream_ahead);
                if (kk == "audio_delay") return f32(c->tts_config.audio_delay);
                return j.skip(); });
        else if (k == "stt_config")                                // config_stt_parse (config.h:75-92)
            ok = object([&](const std::string &kk) {
                if (kk == "audio_delay_seconds") return f32(c->stt_config.audio_delay_seconds);
                if (kk == "audio_silence_prefix_seconds") return f32(c->stt_config.audio_silence_prefix_seconds);
                return j.skip(); });
        else if (k == "model_id")                                  // config_model_id_parse (config.h:96-114)
            ok = object([&](const std::string &kk) {
                if (kk == "sig") return string_or_null(c->model_id.sig);
                if (kk == "epoch") return i64(c->model_id.epoch);
                return j.skip(); });
        else if (k == "lm_gen_config")                             // config_lm_gen_parse (config.h:118-145)
            ok = object([&](const std::string &kk) {
                if (kk == "temp") return f32(c->lm_gen_config.temp);
                if (kk == "temp_text") return f32(c->lm_gen_config.temp_text);
                if (kk == "top_k") return i64(c->lm_gen_config.top_k);
                if (kk == "top_k_text") return i64(c->lm_gen_config.top_k_text);
                return j.skip(); });
        else if (k == "fuser")                                     // config_fuser_parse (config.h:20-50)
            ok = object([&](const std::string &kk) {
                if (kk == "cross_attention_pos_emb") return boolean(c->fuser.cross_attention_pos_emb);
                if (kk == "cross_attention_pos_emb_scale") return f32(c->fuser.cross_attention_pos_emb_scale);
                if (kk == "sum") return str_arr(c->fuser.sum);
                if (kk == "cross") return str_arr(c->fuser.cross);
                return j.skip(); });
        else ok = j.skip();
        if (!ok) { fprintf(stderr, "error: reading config %s\n", filename); return -1; }
        j.ws();
        if (j.p < j.e && *j.p == ',') { j.p++; continue; }
        if (j.p < j.e && *j.p == '}') break;
        fprintf(stderr, "error: reading config %s\n", filename); return -1;
    }
    return 0;
}

// ---- LM ------------------------------------------------------------------------------------------------
struct moshi_lm_t {
    std::string filepath;
    int device = 0;
    msx_config cfg{};
    msx_model *model = nullptr;
    int delay_steps = 0;
    int second_stream_ahead = 0;
    std::string want_quant;
};

static bool to_msx(const moshi_config_t &c, msx_config *m) {
    memset(m, 0, sizeof(*m));
    if ((int)c.delays.size() > MSX_MAX_CODEBOOKS || (int)c.depformer_weights_per_step_schedule.size() > MSX_MAX_STEPS) return false;
    m->dim = (int)c.dim; m->num_heads = (int)c.num_heads; m->num_layers = (int)c.num_layers; m->context = (int)c.context; m->max_period = (int)c.max_period;
    m->n_q = (int)c.n_q; m->dep_q = (int)c.dep_q; m->card = (int)c.card; m->text_card = (int)c.text_card;
    m->dep_dim = (int)c.depformer_dim; m->dep_heads = (int)c.depformer_num_heads; m->dep_layers = (int)c.depformer_num_layers;
    m->dep_context = (int)c.depformer_context;
    m->dep_max_period = c.depformer_pos_emb == "rope" ? (int)c.depformer_max_period : 0;     // lm_default.h:96-101
    m->n_delays = (int)c.delays.size();
    for (size_t i = 0; i < c.delays.size(); i++) m->delays[i] = (int)c.delays[i];
    m->schedule_len = (int)c.depformer_weights_per_step_schedule.size();
    for (int i = 0; i < m->schedule_len; i++) m->schedule[i] = (int)c.depformer_weights_per_step_schedule[i];
    m->personaplex = c.model_type == "personaplex";
    m->extra_heads = (int)c.extra_heads_num_heads;
    m->cross_attention = c.cross_attention; m->demux_second_stream = c.demux_second_stream;
    m->dep_low_rank = (int)c.depformer_low_rank_embeddings;
    return true;
}

moshi_lm_t *moshi_lm_from_files(moshi_context_t *moshi, moshi_config_t *config, const char *filepath) {
    if (!moshi || !config || !filepath) return nullptr;
    FILE *f = fopen(filepath, "rb");                       // reference: WeightLoader::from_gguf returns NULL (moshi.cpp:621-627)
    if (!f) return nullptr;
    fclose(f);
    auto lm = new moshi_lm_t;
    lm->filepath = filepath; lm->device = moshi->device;
    lm->second_stream_ahead = (int)config->tts_config.second_stream_ahead;      // moshi.cpp:633
    if (!to_msx(*config, &lm->cfg)) { delete lm; return nullptr; }
    return lm;
}
void unref(moshi_lm_t *lm) { if (lm) { msx_model_free(lm->model); delete lm; } }
void moshi_lm_set_delay_steps(moshi_lm_t *lm, int d) { lm->delay_steps = d; }
int moshi_lm_get_max_delay(moshi_lm_t *lm) { int m = lm->cfg.delays[0]; for (int i = 0; i < lm->cfg.n_delays; i++) m = std::max(m, lm->cfg.delays[i]); return m; }
int moshi_lm_get_delay_steps(moshi_lm_t *lm) { return lm->delay_steps; }
static bool is_safetensors(const std::string &path) {
    const std::string ext = ".safetensors";
    return path.size() > ext.size() && path.compare(path.size() - ext.size(), ext.size(), ext) == 0;
}
bool moshi_lm_quantize(moshi_lm_t *lm, const char *quant) {
    // reference: q4_0 / q4_k / q8_0 accepted, anything else false (moshi.cpp:654-673).  q8_0 and q4_k are applied while
    // loading when the GGUF holds f32 / f16 / bf16 tensors; a q4_0 model must already be a q4_0 file (and its linears are
    // rejected by the loader: the GEMV paths take q4_k / q8_0).
    const std::string q = quant ? quant : "";
    if (q != "q4_0" && q != "q4_k" && q != "q8_0") return false;
    lm->want_quant = q;
    return true;
}
int moshi_lm_load(moshi_lm_t *lm) {
    if (lm->model) return 0;
    // moshi_lm_quantize("q8_0" / "q4_k") on an unquantised (f32 / f16 / bf16) GGUF: quantise while loading like the
    // reference (loader.h:149-233); already-quantised tensors are taken as they are
    const int q = lm->want_quant == "q8_0" ? 8 : lm->want_quant == "q4_k" ? 12 : 0;
    if (is_safetensors(lm->filepath)) {
        // the reference's WeightLoader::from_safetensor path (moshi.cpp:611-636): quantise into a scratch GGUF on the GPU,
        // load that, drop it.  The linears need q4_k / q8_0, so -q is mandatory for a safetensors checkpoint.
        std::string tmp = "/tmp/moshi_b200_XXXXXX";
        const int fd = mkstemp(&tmp[0]);
        if (fd < 0) return -2;
        close(fd);
        int e = msx_safetensors_to_gguf(lm->filepath.c_str(), tmp.c_str(), q, lm->device);
        if (!e) e = msx_model_load_gguf_ex(tmp.c_str(), &lm->cfg, lm->device, 0, 1, 0, &lm->model);
        unlink(tmp.c_str());
        return e;
    }
    return msx_model_load_gguf_ex(lm->filepath.c_str(), &lm->cfg, lm->device, 0, 1, q, &lm->model);
}

// moshi.cpp:693-695: the weights as loaded (i.e. after moshi_lm_quantize) written as a GGUF; here a file-to-file pass on
// the GPU with the same quantisers and type rules as the loader.  The reference returns void; failures go to stderr.
void moshi_lm_save_gguf(moshi_lm_t *lm, const char *filepath) {
    const int q = lm->want_quant == "q8_0" ? 8 : lm->want_quant == "q4_k" ? 12 : 0;
    if ((is_safetensors(lm->filepath) ? msx_safetensors_to_gguf : msx_gguf_quantize)(lm->filepath.c_str(), filepath, q, lm->device) != 0)
        fprintf(stderr, "moshi_lm_save_gguf: %s\n", msx_last_error());
}

// ---- TTS text scheduling: TokenIds / State / StateMachine (src/moshi/models/lm.h:5-194) -------------------------
// The model proposes PAD or NEW_WORD; the machine decides what is actually fed: queued word tokens, forced
// padding after a word, at most max_padding pads in a row, the look-ahead word on the second text stream.
struct moshi_tts_machine_t {
    static constexpr int kNewWord = 0, kPad = 3, kZero = -1;     // TokenIds (lm.h:5-18)
    int card, second_stream_ahead, max_padding, initial_padding;
    int remaining_padding, forced_padding, end_step = -1;
    std::deque<Entry> entries;
    std::deque<int> queued, lookahead_queued;

    moshi_tts_machine_t(int text_card, int ahead, int max_pad, int init_pad)
        : card(text_card), second_stream_ahead(ahead), max_padding(max_pad), initial_padding(init_pad),
          remaining_padding(init_pad), forced_padding(init_pad) {}
    void reset() {                                               // reset_state (lm.h:95-102)
        remaining_padding = initial_padding; forced_padding = initial_padding; end_step = -1;
        entries.clear(); queued.clear(); lookahead_queued.clear();
    }
    bool is_empty() const { return entries.empty() && queued.empty() && lookahead_queued.empty(); }
    std::vector<int> tokens_ahead(int lookahead) const {         // State::get_tokens_ahead (lm.h:28-39)
        for (const Entry &e : entries) {
            if (e.tokens.empty()) continue;
            if (--lookahead != 0) continue;
            return e.tokens;
        }
        return {};
    }
    int process(int step, int token) {                           // StateMachine::process (lm.h:104-193)
        if (token != kNewWord && token != kPad) token = kPad;
        if (!queued.empty()) token = kPad;                       // text tokens still to be fed
        else if (forced_padding > 0) token = kPad;
        else if (remaining_padding <= 0) token = kNewWord;       // not allowed to pad any longer
        if (token == kNewWord) {
            if (!entries.empty()) {
                Entry entry = entries.front();
                entries.pop_front();
                if (!entry.tokens.empty()) {
                    for (int t : entry.tokens) queued.push_back(t);
                    if (second_stream_ahead)
                        for (int t : tokens_ahead(second_stream_ahead)) lookahead_queued.push_back(t);
                    remaining_padding = max_padding;
                } else token = kPad;
                forced_padding = entry.padding;
            } else {
                token = kPad;
                if (second_stream_ahead && end_step < 0) token = kNewWord;
                if (end_step < 0) end_step = step;               // consumed past the last word
            }
        }
        int output = 0;
        if (token == kPad) {
            if (remaining_padding > 0) remaining_padding -= 1;
            if (forced_padding > 0) forced_padding -= 1;
            if (!queued.empty()) { output = queued.front(); queued.pop_front(); }
            else output = kPad;
        } else if (token == kNewWord) output = kNewWord;
        else if (token == kZero) output = token;
        if (second_stream_ahead) {
            int second = -1;
            if (output == kNewWord) {
                second = kNewWord;
                if (!queued.empty()) { output = queued.front(); queued.pop_front(); }
                else output = kPad;
            } else if (!lookahead_queued.empty()) { second = lookahead_queued.front(); lookahead_queued.pop_front(); }
            output = (second + 1) * card + output;
        }
        return output;
    }
};
moshi_tts_machine_t *moshi_tts_machine_new(int text_card, int ahead, int max_padding, int initial_padding) { return new moshi_tts_machine_t(text_card, ahead, max_padding, initial_padding); }
void moshi_tts_machine_free(moshi_tts_machine_t *m) { delete m; }
void moshi_tts_machine_push(moshi_tts_machine_t *m, const int *tokens, int n, int padding) { Entry e; e.tokens.assign(tokens, tokens + n); e.padding = padding; m->entries.push_back(e); }
int moshi_tts_machine_process(moshi_tts_machine_t *m, int step, int token) { return m->process(step, token); }
int moshi_tts_machine_end_step(moshi_tts_machine_t *m) { return m->end_step; }
int moshi_tts_machine_is_empty(moshi_tts_machine_t *m) { return m->is_empty() ? 1 : 0; }
void moshi_tts_machine_reset(moshi_tts_machine_t *m) { m->reset(); }

// ---- generator -----------------------------------------------------------------------------------------
struct moshi_lm_gen_t {
    moshi_lm_t *lm = nullptr;
    msx_stream *stream = nullptr;
    msx_gen *gen = nullptr;
    std::vector<int32_t> audio_tokens;                         // moshi_lm_send2 -> next receive
    std::deque<std::vector<int16_t>> prompt_audio;             // personaplex voice prompt (codes)
    std::vector<int> text_prompt_tokens;                       // personaplex system prompt
    std::vector<float> prompt_embeddings; int prompt_rows = 0;  // personaplex voice prompt (embedding variant)
    std::vector<int32_t> prompt_cache; int prompt_cache_rows = 0;
    // TTS (voice_t + machine of the reference's moshi_lm_gen_t, moshi.cpp:586-606)
    bool has_voice = false;
    std::vector<float> cond_sum, cond_cross; int tc = 0;
    std::string voice_path; bool voice_from_file = false;      // moshi_lm_set_voice_condition / _load_voice_condition
    std::deque<int> text_prefixes;
    std::deque<std::vector<int>> audio_prefixes;
    moshi_tts_machine_t *machine = nullptr;
};
moshi_lm_gen_t *moshi_lm_generator(moshi_lm_t *lm) { auto g = new moshi_lm_gen_t; g->lm = lm; return g; }
void unref(moshi_lm_gen_t *g) { if (g) { msx_gen_free(g->gen); msx_stream_free(g->stream); delete g->machine; delete g; } }

int moshi_lm_set_condition(moshi_lm_gen_t *gen, const float *cond_sum, const float *cond_cross, int tc) {
    if (!gen) return -1;
    const msx_config &c = gen->lm->cfg;
    if (cond_cross && !c.cross_attention) return -1;           // uses_cross check of moshi_lm_set_voice_condition
    gen->cond_sum.clear(); gen->cond_cross.clear(); gen->tc = 0;
    if (cond_sum) gen->cond_sum.assign(cond_sum, cond_sum + c.dim);
    if (cond_cross && tc > 0) { gen->cond_cross.assign(cond_cross, cond_cross + (size_t)tc * c.dim); gen->tc = tc; }
    gen->has_voice = true; gen->voice_from_file = false;
    return 0;
}
// moshi.cpp:729-760: -1 without cross-attention or when the voice file cannot be opened, -2 when the model carries no
// conditioners; the conditioners themselves run on the GPU at moshi_lm_start (msx_stream_load_voice)
int moshi_lm_set_voice_condition(moshi_context_t *, moshi_lm_gen_t *gen, const char *filepath) {
    if (!gen->lm->cfg.cross_attention) return -1;
    FILE *f = filepath ? fopen(filepath, "rb") : nullptr;
    if (!f) return -1;
    fclose(f);
    gen->voice_path = filepath;
    return 0;
}
int moshi_lm_load_voice_condition(moshi_context_t *, moshi_lm_gen_t *gen) {
    if (!gen->lm->cfg.cross_attention) return -1;
    if (!gen->lm->model || !msx_model_has_conditioners(gen->lm->model)) return -2;
    if (gen->voice_path.empty()) return -1;
    gen->has_voice = gen->voice_from_file = true;
    return 0;
}
int moshi_lm_voice_prefix(moshi_lm_gen_t *gen, std::deque<int> &text_prefix, std::deque<std::vector<int>> &audio_prefix) {
    gen->text_prefixes.clear(); gen->audio_prefixes.clear();
    gen->text_prefixes.swap(text_prefix);                      // the reference swaps (steals) both deques (moshi.cpp:768-769)
    gen->audio_prefixes.swap(audio_prefix);
    gen->has_voice = true;
    return 0;
}
void moshi_lm_send(moshi_lm_gen_t *gen, Entry *entry) { if (gen->machine && entry) gen->machine->entries.push_back(*entry); }
int moshi_lm_is_active(moshi_lm_gen_t *gen) {                  // moshi.cpp:940-945
    if (!gen->machine || !gen->gen) return 0;
    const int final_padding = 4;
    const int end_offset = gen->machine->end_step + gen->lm->delay_steps + final_padding;
    return (msx_gen_offset(gen->gen) < end_offset || gen->machine->end_step == -1) ? 1 : 0;
}
int moshi_lm_is_empty(moshi_lm_gen_t *gen) { return (!gen->machine || gen->machine->is_empty()) ? 1 : 0; }
void moshi_lm_machine_reset(moshi_lm_gen_t *gen) { if (gen->machine) gen->machine->reset(); }

// on_text_hook / on_audio_hook of moshi_lmgen_step (lm.h:877-899, 922-931)
static int32_t tts_text_hook(void *user, int32_t offset, int32_t text_token) {
    auto gen = static_cast<moshi_lm_gen_t *>(user);
    if (!gen->text_prefixes.empty()) { const int t = gen->text_prefixes.front(); gen->text_prefixes.pop_front(); return t; }
    return gen->machine->process(offset, text_token);
}
static int tts_audio_hook(void *user, int32_t, int32_t *audio, int n) {
    auto gen = static_cast<moshi_lm_gen_t *>(user);
    if (gen->audio_prefixes.empty()) return -1;
    const std::vector<int> &codes = gen->audio_prefixes.front();
    for (int q = 0; q < n && q < (int)codes.size(); q++) if (codes[q] != -2) audio[q] = codes[q];   // lm_ungenerated_token_id
    gen->audio_prefixes.pop_front();
    return 2;                                                   // skip_prefix default (lm.h:787)
}

int moshi_lm_personaplex_audio_prompt(moshi_lm_gen_t *gen, std::deque<std::vector<int16_t>> &audio_prompt) {
    gen->prompt_audio.clear();
    gen->prompt_audio.swap(audio_prompt);                      // the reference swaps (steals) the caller's deque (moshi.cpp:782)
    return 0;
}
int moshi_lm_personaplex_voice_tensors(moshi_lm_gen_t *gen, const float *embeddings, int n_rows, const int32_t *cache, int cache_rows) {
    if (!gen || !embeddings || n_rows <= 0 || !cache) return -1;
    const msx_config &c = gen->lm->cfg;
    gen->prompt_embeddings.assign(embeddings, embeddings + (size_t)n_rows * c.dim);
    gen->prompt_rows = n_rows;
    gen->prompt_cache.assign(cache, cache + (size_t)cache_rows * (c.n_q + 1));
    gen->prompt_cache_rows = cache_rows;
    return 0;
}
// element i of a float / integer tensor stored as `dtype` (safetensors names) -> double
static bool voice_element(const uint8_t *p, const std::string &dt, size_t i, double &out) {
    if (dt == "F32") { float v; memcpy(&v, p + i * 4, 4); out = v; }
    else if (dt == "BF16") { uint16_t h; memcpy(&h, p + i * 2, 2); const uint32_t u = (uint32_t)h << 16; float v; memcpy(&v, &u, 4); out = v; }
    else if (dt == "F16") {
        uint16_t h; memcpy(&h, p + i * 2, 2);
        const int e = (h >> 10) & 31, m = h & 1023;
        const double mag = e == 0 ? std::ldexp((double)m, -24) : e == 31 ? (m ? NAN : INFINITY) : std::ldexp((double)(m | 1024), e - 25);
        out = (h & 0x8000) ? -mag : mag;
    }
    else if (dt == "I32") { int32_t v; memcpy(&v, p + i * 4, 4); out = v; }
    else if (dt == "I64") { int64_t v; memcpy(&v, p + i * 8, 8); out = (double)v; }
    else return false;
    return true;
}
int moshi_lm_personaplex_load_voice(moshi_context_t *, moshi_lm_gen_t *gen, const char *filepath) {
    if (!gen || !filepath) return -1;
    const std::string filename = filepath;
    const auto ext_index = filename.find_last_of('.');
    if (ext_index == std::string::npos) return -1;
    const std::string ext = filename.substr(ext_index);
    const msx_config &c = gen->lm->cfg;
    const int ncb = c.n_q + 1;
    // (data, dtype, element count, innermost extent) of the two tensors
    const uint8_t *emb = nullptr, *cache = nullptr;
    std::string emb_dt, cache_dt;
    int64_t emb_n = 0, cache_n = 0, cache_inner = 0;
    msx::SafeTensorsFile st;
    msx::GgufFile gg;
    std::string err;
    if (ext == ".safetensors") {
        if (!st.open(filename, err)) return -1;
        for (const auto &t : st.tensors()) {
            int64_t n = 1;
            for (int64_t v : t.shape) n *= v;
            if (t.name == "embeddings") { emb = t.data; emb_dt = t.dtype; emb_n = n; }
            else if (t.name == "cache" && !t.shape.empty()) { cache = t.data; cache_dt = t.dtype; cache_n = n; cache_inner = t.shape.back(); }
        }
    } else if (ext == ".gguf") {
        if (!gg.open(filename, err)) return -1;
        auto dt = [](int type) { return type == 0 ? "F32" : type == 1 ? "F16" : type == 30 ? "BF16" : type == 26 ? "I32" : ""; };
        if (const msx::GgufTensor *t = gg.find("voice.embeddings")) if (t->data) { emb = t->data; emb_dt = dt(t->type); emb_n = t->ne[0] * t->ne[1] * t->ne[2] * t->ne[3]; }
        if (const msx::GgufTensor *t = gg.find("voice.cache")) if (t->data) { cache = t->data; cache_dt = dt(t->type); cache_n = t->ne[0] * t->ne[1] * t->ne[2] * t->ne[3]; cache_inner = t->ne[0]; }
    } else return -1;
    if (!emb || !cache || emb_n <= 0 || emb_n % c.dim || cache_n <= 0 || cache_n != cache_inner * ncb) return -1;
    std::vector<float> rows((size_t)emb_n);
    for (size_t i = 0; i < rows.size(); i++) { double v; if (!voice_element(emb, emb_dt, i, v)) return -1; rows[i] = (float)v; }
    // the file holds cache[codebook][time] (lm.h:1047-1051: state->cache[i][j] = cache[i + j * height]); ours is [time][codebook]
    const int CT = (int)cache_inner;
    std::vector<int32_t> ring((size_t)cache_n);
    for (int j = 0; j < ncb; j++)
        for (int i = 0; i < CT; i++) { double v; if (!voice_element(cache, cache_dt, (size_t)j * CT + i, v)) return -1; ring[(size_t)i * ncb + j] = (int32_t)v; }
    return moshi_lm_personaplex_voice_tensors(gen, rows.data(), (int)(emb_n / c.dim), ring.data(), CT);
}
// ---- tokenizer: plain vocabulary file, greedy longest match (the reference wraps sentencepiece, which this build lacks) ------
struct tokenizer_t {
    std::vector<std::string> pieces;                     // id -> piece ("\xe2\x96\x81" = U+2581 marks a word start)
    std::unordered_map<std::string, int> ids;
    size_t longest = 1;
    bool insert_bos = true;
    std::deque<Entry> pending;
    std::vector<int> encode(const std::string &text) const {
        std::vector<int> out;
        std::string norm;                                 // sentencepiece normalisation, reduced: spaces -> U+2581, one in front
        norm = "\xe2\x96\x81";
        for (char ch : text) { if (ch == ' ') norm += "\xe2\x96\x81"; else norm.push_back(ch); }
        size_t i = 0;
        while (i < norm.size()) {
            size_t n = std::min(longest, norm.size() - i);
            int id = -1;
            for (; n > 0; n--) { auto it = ids.find(norm.substr(i, n)); if (it != ids.end()) { id = it->second; break; } }
            if (id < 0) { id = 0; n = 1; while (i + n < norm.size() && ((unsigned char)norm[i + n] & 0xC0) == 0x80) n++; }   // <unk>, one UTF-8 character
            out.push_back(id);
            i += n;
        }
        return out;
    }
};
tokenizer_t *tokenizer_alloc(const char *filepath, bool insert_bos) {
    FILE *f = filepath ? fopen(filepath, "rb") : nullptr;
    if (!f) return nullptr;
    std::string raw; char buf[4096]; size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) raw.append(buf, n);
    fclose(f);
    if (raw.find('\0') != std::string::npos) return nullptr;       // a sentencepiece .model protobuf, not a vocabulary listing
    tokenizer_t *t = new tokenizer_t;
    t->insert_bos = insert_bos;
    size_t a = 0;
    while (a < raw.size()) {
        size_t b = raw.find('\n', a); if (b == std::string::npos) b = raw.size();
        std::string piece = raw.substr(a, b - a);
        if (!piece.empty() && piece.back() == '\r') piece.pop_back();
        const size_t tab = piece.find('\t'); if (tab != std::string::npos) piece.resize(tab);      // "piece<TAB>score" listings
        if (!piece.empty() && piece[0] == ' ') piece = "\xe2\x96\x81" + piece.substr(1);
        t->ids.emplace(piece, (int)t->pieces.size());
        t->longest = std::max(t->longest, piece.size());
        t->pieces.push_back(piece);
        a = b + 1;
    }
    if (t->pieces.empty()) { delete t; return nullptr; }
    return t;
}
void unref(tokenizer_t *tok) { delete tok; }
bool tokenizer_empty(tokenizer_t *tok) { return !tok || tok->pending.empty(); }
// text -> one Entry per word (the TTS word queue, moshi.h:63-68): word tokens, the word itself, padding 0
int tokenizer_send(tokenizer_t *tok, std::string text) {
    if (!tok) return -1;
    size_t a = 0; int n = 0;
    while (a < text.size()) {
        while (a < text.size() && isspace((unsigned char)text[a])) a++;
        size_t b = a;
        while (b < text.size() && !isspace((unsigned char)text[b])) b++;
        if (b > a) {
            Entry e; e.text = text.substr(a, b - a); e.tokens = tok->encode(e.text); e.padding = 0;
            if (tok->insert_bos && tok->pending.empty() && n == 0) e.tokens.insert(e.tokens.begin(), 1);
            tok->pending.push_back(std::move(e)); n++;
        }
        a = b;
    }
    return n;
}
int tokenizer_receive(tokenizer_t *tok, Entry *entry) {
    if (!tok || !entry || tok->pending.empty()) return 0;
    *entry = std::move(tok->pending.front()); tok->pending.pop_front();
    return 1;
}
std::string tokenizer_id_to_piece(tokenizer_t *tok, int token) {
    if (!tok || token < 0 || (size_t)token >= tok->pieces.size()) return std::string();
    return tok->pieces[(size_t)token];
}

// moshi.cpp:838-849: "<system> " + prompt + " <system>" through the tokenizer
int moshi_lm_personaplex_system_prompt(moshi_context_t *, moshi_lm_gen_t *gen, tokenizer_t *tok, const char *prompt) {
    if (!gen || !tok || !prompt) return -1;
    gen->text_prompt_tokens = tok->encode(std::string("<system> ") + prompt + " <system>");
    return 0;
}
int moshi_lm_personaplex_system_prompt_tokens(moshi_lm_gen_t *gen, const std::vector<int> &text_tokens) {
    gen->text_prompt_tokens = text_tokens;
    return 0;
}

// moshi_lmgen_step_system_prompts (lm.h:983-1134): every prompt frame is a full 17-token row replayed through the step
static const int PROMPT_TOKENS[17] = {3, 948, 243, 1178, 546, 1736, 1030, 1978, 2008, 430, 1268, 381, 1611, 1095, 1495, 56, 472};
static void personaplex_prompts(moshi_lm_gen_t *gen) {
    const msx_config &c = gen->lm->cfg;
    const int ncb = c.n_q + 1;
    if (ncb != 17) return;                                     // the reference's table has 17 entries
    int32_t row[MSX_MAX_CODEBOOKS], text, audio[MSX_MAX_STEPS];
    // The table holds codes of the PersonaPlex-7B codebooks (card 2048).  A model with smaller tables (test models) cannot embed
    // them — the reference would index past its tables — so they are folded into this model's range; for card 2048 a no-op.
    int32_t prompt_row[17];
    for (int i = 0; i < 17; i++) prompt_row[i] = PROMPT_TOKENS[i] % (i == 0 ? c.text_card + 1 : c.card);
    // every prompt frame is a full token row: collect them all, then run them as ONE batched-T prefill (8 frames per
    // weight pass); models / situations the prefill does not cover fall back to one decode step per frame like the reference
    std::vector<int32_t> rows;
    auto push_row = [&]() { rows.insert(rows.end(), row, row + ncb); };
    auto flush = [&]() {
        if (rows.empty()) return;
        const int T = (int)(rows.size() / ncb);
        if (msx_gen_prefill(gen->gen, rows.data(), T) != 0)
            for (int f = 0; f < T; f++) msx_gen_step(gen->gen, rows.data() + (size_t)f * ncb, ncb, 0, &text, audio);
        rows.clear();
    };
    if (gen->prompt_rows > 0) {                                // embedding variant (lm.h:1005-1051)
        for (int i = 0; i < gen->prompt_rows; i++) msx_gen_prompt_embedding(gen->gen, gen->prompt_embeddings.data() + (size_t)i * c.dim);
        if (gen->prompt_cache_rows == msx_gen_cache_rows(gen->gen)) msx_gen_set_cache(gen->gen, gen->prompt_cache.data());
    } else
    while (!gen->prompt_audio.empty()) {                       // voice prompt: codes of the 8 moshi codebooks
        for (int i = 0; i < ncb; i++) row[i] = prompt_row[i];
        const auto &codes = gen->prompt_audio.front();
        for (int j = 0; j < 8 && j < (int)codes.size(); j++) row[j + 1] = codes[j];
        push_row();
        gen->prompt_audio.pop_front();
    }
    auto silence = [&](int n) { for (int f = 0; f < n; f++) { for (int i = 0; i < ncb; i++) row[i] = prompt_row[i]; push_row(); } };
    silence(6);
    for (int tok : gen->text_prompt_tokens) { for (int i = 0; i < ncb; i++) row[i] = prompt_row[i]; row[0] = tok; push_row(); }
    silence(6);
    flush();
}

void moshi_lm_start(moshi_context_t *, moshi_lm_gen_t *gen, float depth_temperature, float text_temperature, bool) {
    // like the reference: use_sampling = true, top_k = 250 (audio) / 25 (text) (moshi.cpp:862-877); a temperature
    // of 0 selects the greedy path (sampling.h:57-63).  Exp(1) noise comes from libc rand() inside msx_gen_step.
    if (gen->gen) { msx_gen_free(gen->gen); gen->gen = nullptr; }
    if (gen->stream) { msx_stream_free(gen->stream); gen->stream = nullptr; }
    if (msx_stream_create(gen->lm->model, 0, &gen->stream) != 0) { fprintf(stderr, "moshi_b200: %s\n", msx_last_error()); return; }
    if (depth_temperature > 0.f || text_temperature > 0.f)
        if (msx_stream_set_sampling(gen->stream, text_temperature, depth_temperature, 25, 250) != 0) { fprintf(stderr, "moshi_b200: %s\n", msx_last_error()); return; }
    if (gen->has_voice && !gen->lm->cfg.personaplex) {
        // TTS: conditioning memory + state machine (moshi.cpp:857-871: max_padding 8, initial_padding 2)
        if (gen->voice_from_file) {
            if (msx_stream_load_voice(gen->stream, gen->voice_path.c_str()) != 0) { fprintf(stderr, "moshi_b200: %s\n", msx_last_error()); return; }
        } else if (!gen->cond_sum.empty() || gen->tc > 0)
            if (msx_stream_set_condition(gen->stream, gen->cond_sum.empty() ? nullptr : gen->cond_sum.data(),
                                         gen->tc > 0 ? gen->cond_cross.data() : nullptr, gen->tc) != 0) { fprintf(stderr, "moshi_b200: %s\n", msx_last_error()); return; }
        delete gen->machine;
        gen->machine = new moshi_tts_machine_t(gen->lm->cfg.text_card + 1, gen->lm->second_stream_ahead, 8, 2);
    }
    if (msx_gen_create(gen->stream, gen->lm->delay_steps, &gen->gen) != 0) { fprintf(stderr, "moshi_b200: %s\n", msx_last_error()); return; }
    if (gen->machine) {
        msx_gen_set_text_hook(gen->gen, tts_text_hook, gen);
        msx_gen_set_audio_hook(gen->gen, tts_audio_hook, gen);
    }
    gen->audio_tokens.assign(gen->lm->cfg.n_q, 0);
    if (gen->lm->cfg.personaplex) personaplex_prompts(gen);
}

void moshi_lm_send2(moshi_lm_gen_t *gen, std::vector<int16_t> &audio_tokens) {
    gen->audio_tokens.assign(audio_tokens.begin(), audio_tokens.end());
}

int moshi_lm_receive(moshi_lm_gen_t *gen, int &text_token, std::vector<int16_t> &audio_tokens) {
    if (!gen->gen) return 0;
    const msx_config &c = gen->lm->cfg;
    const int replace = msx_gen_offset(gen->gen) < gen->lm->delay_steps;                 // moshi.cpp:905
    int32_t out[MSX_MAX_STEPS] = {0}, text = 0;
    const int rc = msx_gen_step(gen->gen, gen->audio_tokens.data(), (int)gen->audio_tokens.size(), replace, &text, out);
    audio_tokens.resize(c.dep_q);
    if (rc < 0) {
        // a failed step (bad token id, CUDA error) must not feed whatever is in `out` into the next frame; the reference would
        // have asserted — here the previous tokens stay, the error is reported and the caller sees a negative code
        fprintf(stderr, "moshi_lm_receive: %s\n", msx_last_error());
        return rc;
    }
    if (rc == 1) { text_token = text; for (int i = 0; i < c.dep_q; i++) audio_tokens[i] = (int16_t)out[i]; }
    // like the reference, the generated row becomes the "sent" tokens of the next call unless send2 overwrites them
    gen->audio_tokens.assign(out, out + c.dep_q);
    return rc == 1 ? 1 : 0;
}

void moshi_lm_receive2(moshi_lm_gen_t *gen, int &text_token, float &vad) {
    if (!gen->gen) return;
    int32_t out[MSX_MAX_STEPS] = {0}, text = 0;
    const int rc = msx_gen_step(gen->gen, gen->audio_tokens.data(), (int)gen->audio_tokens.size(), 0, &text, out);
    if (rc < 0) { fprintf(stderr, "moshi_lm_receive2: %s\n", msx_last_error()); return; }
    if (rc == 1) {                       // the reference evaluates the VAD head only on frames that emit (lm.h:950-977)
        text_token = text;
        msx_vad(gen->stream, &vad);
    }
}
