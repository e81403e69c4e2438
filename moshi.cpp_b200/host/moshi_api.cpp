// moshi_api.cpp — see moshi_api.h.  Mirrors src/moshi.cpp:600-953 (LM + generator) over the msx C ABI.
#include "moshi_api.h"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstring>

#include "../../include/moshi_b200.h"

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
    return true;
}

moshi_lm_t *moshi_lm_from_files(moshi_context_t *moshi, moshi_config_t *config, const char *filepath) {
    if (!moshi || !config || !filepath) return nullptr;
    FILE *f = fopen(filepath, "rb");                       // reference: WeightLoader::from_gguf returns NULL (moshi.cpp:621-627)
    if (!f) return nullptr;
    fclose(f);
    if (config->cross_attention || config->demux_second_stream) {
        fprintf(stderr, "moshi_b200: cross-attention / demux (TTS) models are not supported yet\n");
        return nullptr;
    }
    auto lm = new moshi_lm_t;
    lm->filepath = filepath; lm->device = moshi->device;
    if (!to_msx(*config, &lm->cfg)) { delete lm; return nullptr; }
    return lm;
}
void unref(moshi_lm_t *lm) { if (lm) { msx_model_free(lm->model); delete lm; } }
void moshi_lm_set_delay_steps(moshi_lm_t *lm, int d) { lm->delay_steps = d; }
int moshi_lm_get_max_delay(moshi_lm_t *lm) { int m = lm->cfg.delays[0]; for (int i = 0; i < lm->cfg.n_delays; i++) m = std::max(m, lm->cfg.delays[i]); return m; }
int moshi_lm_get_delay_steps(moshi_lm_t *lm) { return lm->delay_steps; }
bool moshi_lm_quantize(moshi_lm_t *lm, const char *quant) {
    // reference: q4_0 / q4_k / q8_0 accepted, anything else false (moshi.cpp:654-673).  Quantise-on-load from
    // safetensors is not built: GGUF files must already carry the requested type.
    const std::string q = quant ? quant : "";
    if (q != "q4_0" && q != "q4_k" && q != "q8_0") return false;
    lm->want_quant = q;
    return true;
}
int moshi_lm_load(moshi_lm_t *lm) {
    if (lm->model) return 0;
    return msx_model_load_gguf(lm->filepath.c_str(), &lm->cfg, lm->device, &lm->model);
}

// ---- generator -----------------------------------------------------------------------------------------
struct moshi_lm_gen_t {
    moshi_lm_t *lm = nullptr;
    msx_stream *stream = nullptr;
    msx_gen *gen = nullptr;
    std::vector<int32_t> audio_tokens;                         // moshi_lm_send2 -> next receive
    std::deque<std::vector<int16_t>> prompt_audio;             // personaplex voice prompt (codes)
    std::vector<int> text_prompt_tokens;                       // personaplex system prompt
};
moshi_lm_gen_t *moshi_lm_generator(moshi_lm_t *lm) { auto g = new moshi_lm_gen_t; g->lm = lm; return g; }
void unref(moshi_lm_gen_t *g) { if (g) { msx_gen_free(g->gen); msx_stream_free(g->stream); delete g; } }

int moshi_lm_personaplex_audio_prompt(moshi_lm_gen_t *gen, std::deque<std::vector<int16_t>> &audio_prompt) {
    gen->prompt_audio.clear();
    gen->prompt_audio.swap(audio_prompt);                      // the reference swaps (steals) the caller's deque (moshi.cpp:782)
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
    auto step = [&]() { msx_gen_step(gen->gen, row, ncb, 0, &text, audio); };
    while (!gen->prompt_audio.empty()) {                       // voice prompt: codes of the 8 moshi codebooks
        for (int i = 0; i < ncb; i++) row[i] = PROMPT_TOKENS[i];
        const auto &codes = gen->prompt_audio.front();
        for (int j = 0; j < 8 && j < (int)codes.size(); j++) row[j + 1] = codes[j];
        step();
        gen->prompt_audio.pop_front();
    }
    auto silence = [&](int n) { for (int f = 0; f < n; f++) { for (int i = 0; i < ncb; i++) row[i] = PROMPT_TOKENS[i]; step(); } };
    silence(6);
    for (int tok : gen->text_prompt_tokens) { for (int i = 0; i < ncb; i++) row[i] = PROMPT_TOKENS[i]; row[0] = tok; step(); }
    silence(6);
}

void moshi_lm_start(moshi_context_t *, moshi_lm_gen_t *gen, float depth_temperature, float text_temperature, bool) {
    // like the reference: use_sampling = true, top_k = 250 (audio) / 25 (text) (moshi.cpp:862-877); a temperature
    // of 0 selects the greedy path (sampling.h:57-63).  Exp(1) noise comes from libc rand() inside msx_gen_step.
    if (gen->gen) { msx_gen_free(gen->gen); gen->gen = nullptr; }
    if (gen->stream) { msx_stream_free(gen->stream); gen->stream = nullptr; }
    if (msx_stream_create(gen->lm->model, 0, &gen->stream) != 0) { fprintf(stderr, "moshi_b200: %s\n", msx_last_error()); return; }
    if (depth_temperature > 0.f || text_temperature > 0.f)
        if (msx_stream_set_sampling(gen->stream, text_temperature, depth_temperature, 25, 250) != 0) { fprintf(stderr, "moshi_b200: %s\n", msx_last_error()); return; }
    if (msx_gen_create(gen->stream, gen->lm->delay_steps, &gen->gen) != 0) { fprintf(stderr, "moshi_b200: %s\n", msx_last_error()); return; }
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
    int32_t out[MSX_MAX_STEPS], text = 0;
    const int rc = msx_gen_step(gen->gen, gen->audio_tokens.data(), (int)gen->audio_tokens.size(), replace, &text, out);
    audio_tokens.resize(c.dep_q);
    if (rc == 1) { text_token = text; for (int i = 0; i < c.dep_q; i++) audio_tokens[i] = (int16_t)out[i]; }
    // like the reference, the generated row becomes the "sent" tokens of the next call unless send2 overwrites them
    gen->audio_tokens.assign(out, out + c.dep_q);
    return rc == 1 ? 1 : 0;
}

void moshi_lm_receive2(moshi_lm_gen_t *gen, int &text_token, float &vad) {
    if (!gen->gen) return;
    int32_t out[MSX_MAX_STEPS], text = 0;
    if (msx_gen_step(gen->gen, gen->audio_tokens.data(), (int)gen->audio_tokens.size(), 0, &text, out) == 1) text_token = text;
    msx_vad(gen->stream, &vad);
}
