// generator.inl — msx_gen: LMGen host logic over a stream (delay ring, provided / replaced tokens, prompt prefill, hooks).
// Reference: moshi_lmgen_state_t lm.h:715-743, moshi_lmgen_step lm.h:778-979.  Included by engine.cu.

// -------------------------------------------------------------------------------------------------
// LMGen host logic (reference lm.h:715-743 state, lm.h:778-979 step; greedy, no state machine)
// -------------------------------------------------------------------------------------------------
// Per-generator libc random state.  The reference draws its Exp(1) noise with rand() (context.h:464-480), i.e. glibc's
// process-global TYPE_3 additive-feedback generator; random_r on a private 128-byte state produces the same sequence for the
// same seed (rand() before any srand() behaves like seed 1), so a single generator reproduces the reference's draws while
// several generators / threads no longer perturb each other.
struct GenRng {
    struct random_data rd;
    char state[128];
    GenRng() { seed(1u); }
    void seed(unsigned v) { memset(&rd, 0, sizeof(rd)); memset(state, 0, sizeof(state)); initstate_r(v, state, sizeof(state), &rd); }
    int next() { int32_t r = 0; random_r(&rd, &r); return (int)r; }
    float exp1() { return -logf(next() / (float)RAND_MAX); }
};

struct msx_gen {
    GenRng rng;
    msx_stream *s = nullptr;
    msx_config cfg{};
    msx_step_fn fn = nullptr;         // host-logic tests: the model step is a caller-supplied callback
    void *user = nullptr;
    int offset = 0, CT = 0, ncb = 0, max_delay = 0, delay_steps = 0;
    std::vector<int32_t> cache;       // [CT][ncb], init -2 = lm_ungenerated_token_id
    std::vector<int32_t> initial;     // {text_card, card, card, ...}
    // TTS hooks (lm.h:877-899 on_text_hook, 915-931 on_audio_hook) and the frames to swallow after an audio prefix
    msx_text_hook text_hook = nullptr; void *text_user = nullptr;
    msx_audio_hook audio_hook = nullptr; void *audio_user = nullptr;
    int skip = 0;
};

static void gen_init_impl(msx_gen *g, int delay_steps);
static void gen_init(msx_gen *g, int delay_steps) { gen_init_impl(g, delay_steps); }

extern "C" int msx_gen_create(msx_stream *s, int delay_steps, msx_gen **out) {
    if (!s || !out) return fail(MSX_ERR_ARG, "null argument");
    const msx_config &c = s->m->cfg;
    auto *g = new msx_gen;
    g->s = s; g->cfg = c;
    gen_init(g, delay_steps);
    *out = g;
    return 0;
}

extern "C" int msx_gen_create_with_callback(const msx_config *cfg, int delay_steps, msx_step_fn fn, void *user, msx_gen **out) {
    if (!cfg || !fn || !out) return fail(MSX_ERR_ARG, "null argument");
    if (cfg->n_q < 0 || cfg->n_q + 1 > MSX_MAX_CODEBOOKS || cfg->dep_q < 0 || cfg->dep_q > MSX_MAX_STEPS || cfg->n_delays < cfg->n_q + 1)
        return fail(MSX_ERR_ARG, "bad codebook counts / delays");
    auto *g = new msx_gen;
    g->cfg = *cfg; g->fn = fn; g->user = user;
    gen_init(g, delay_steps);
    *out = g;
    return 0;
}

static void gen_init_impl(msx_gen *g, int delay_steps) {
    const msx_config &c = g->cfg;
    g->ncb = c.n_q + 1; g->delay_steps = delay_steps;
    int md = c.delays[0];
    for (int i = 0; i < c.n_delays; i++) md = std::max(md, c.delays[i]);     // lm_default.h:177-183
    g->max_delay = md;
    g->CT = md + 2 + (c.personaplex ? 1 : 0);                                // lm.h:727-729
    g->cache.assign((size_t)g->CT * g->ncb, -2);
    g->initial.assign(g->ncb, c.card);
    g->initial[0] = c.text_card;
}
extern "C" void msx_gen_free(msx_gen *g) { delete g; }
extern "C" void msx_gen_seed(msx_gen *g, unsigned seed) { if (g) g->rng.seed(seed); }
extern "C" int msx_gen_offset(const msx_gen *g) { return g ? g->offset : -1; }
extern "C" int msx_gen_max_delay(const msx_gen *g) { return g ? g->max_delay : -1; }

// moshi_lmgen_step_voice_prompt, embedding variant (lm.h:1005-1050): one prompt frame = temporal step on the row,
// text token forced to 3, depformer step (its tokens are dropped), offset++
extern "C" int msx_gen_prompt_embedding(msx_gen *g, const float *row) {
    if (!g || !row || !g->s) return fail(MSX_ERR_ARG, "null argument / callback generator");
    int32_t text = 0, audio[MSX_MAX_STEPS];
    msx_stream *s = g->s; const msx_config &c = g->cfg;
    if (s->temp_text > 0.f || s->temp_audio > 0.f) {
        // the sampled graphs read this frame's Exp(1) noise: draw it exactly like msx_gen_step (and like the reference's
        // moshi_lmgen_step_voice_prompt, whose two graphs consume kt + dep_q * ka draws per prompt frame)
        const int kt = std::min(std::min(s->top_k_text, c.text_card), kSampleMaxK), ka = std::min(std::min(s->top_k_audio, c.card), kSampleMaxK);
        std::vector<float> nt(kt), na((size_t)std::max(1, c.dep_q) * ka);
        if (s->temp_text > 0.f) for (int i = 0; i < kt; i++) nt[i] = g->rng.exp1();
        if (s->temp_audio > 0.f) for (size_t i = 0; i < (size_t)c.dep_q * ka; i++) na[i] = g->rng.exp1();
        if (int e = msx_stream_set_noise(s, nt.data(), na.data())) return e;
    }
    if (int e = msx_step_temporal_embedding(g->s, row, &text, nullptr, nullptr)) return e;
    if (g->cfg.dep_q > 0) if (int e = msx_step_depformer(g->s, 3, nullptr, audio, nullptr)) return e;
    g->offset++;
    return 0;
}
// the token delay ring as stored with the voice ("voice.cache"): cache[CT][n_q+1] row-major here
extern "C" int msx_gen_set_cache(msx_gen *g, const int32_t *cache) {
    if (!g || !cache) return fail(MSX_ERR_ARG, "null argument");
    for (size_t i = 0; i < g->cache.size(); i++) g->cache[i] = cache[i];
    return 0;
}
extern "C" int msx_gen_cache_rows(const msx_gen *g) { return g ? g->CT : -1; }

// T prompt frames with ALL n_q+1 tokens given, as one batched-T prefill: the host side of moshi_lmgen_step's "provided"
// branch (ring writes lm.h:812-818, input gather 826-833, no output write-back 933-943, offset++) for every frame, then
// msx_stream_prefill on the gathered inputs.  The libc rand() draws the per-frame sampler would have consumed are
// consumed here too, so a sampled conversation continues with the reference's random sequence.
extern "C" int msx_stream_prefill(msx_stream *s, const int32_t *tokens, int T);
extern "C" int msx_gen_prefill(msx_gen *g, const int32_t *rows, int T) {
    if (!g || !rows || T <= 0 || !g->s) return fail(MSX_ERR_ARG, "bad argument / callback generator");
    const msx_config &c = g->cfg;
    const int CT = g->CT, ncb = g->ncb;
    std::vector<int32_t> inputs((size_t)T * ncb), cache = g->cache;
    int offset = g->offset;
    for (int f = 0; f < T; f++) {
        for (int i = 0; i < ncb; i++) cache[(size_t)((offset + c.delays[i]) % CT) * ncb + i] = rows[(size_t)f * ncb + i];
        const int pos = offset % CT;
        for (int i = 0; i < ncb; i++) inputs[(size_t)f * ncb + i] = (offset <= c.delays[i]) ? g->initial[i] : cache[(size_t)pos * ncb + i];
        offset++;
    }
    if (int e = msx_stream_prefill(g->s, inputs.data(), T)) return e;
    g->cache.swap(cache); g->offset = offset;
    if (g->s->temp_text > 0.f || g->s->temp_audio > 0.f) {
        const int kt = std::min(std::min(g->s->top_k_text, c.text_card), kSampleMaxK), ka = std::min(std::min(g->s->top_k_audio, c.card), kSampleMaxK);
        const long draws = (long)T * ((g->s->temp_text > 0.f ? kt : 0) + (g->s->temp_audio > 0.f ? (long)c.dep_q * ka : 0));
        for (long i = 0; i < draws; i++) (void)g->rng.next();
    }
    return 0;
}

extern "C" int msx_gen_set_text_hook(msx_gen *g, msx_text_hook fn, void *user) {
    if (!g) return fail(MSX_ERR_ARG, "null generator");
    g->text_hook = fn; g->text_user = user;
    return 0;
}
extern "C" int msx_gen_set_audio_hook(msx_gen *g, msx_audio_hook fn, void *user) {
    if (!g) return fail(MSX_ERR_ARG, "null generator");
    g->audio_hook = fn; g->audio_user = user;
    return 0;
}

// moshi_lmgen_step (lm.h:778-979) around the model call: gen_prepare = ring write of the incoming tokens + input gather,
// gen_finish = audio hooks, ring write of the generated tokens, delayed emit.  Shared by msx_gen_step (one stream) and
// msx_bgen_step (a batch of streams, batch.inl).
struct GenPrep { bool provided = false; int32_t input[MSX_MAX_CODEBOOKS]; };
static int gen_prepare(msx_gen *g, const int32_t *in_tokens, int n_in, GenPrep *p) {
    const msx_config &c = g->cfg;
    const int CT = g->CT, ncb = g->ncb;
    int dep_q = c.dep_q;
    if (c.personaplex) dep_q = 8;                                             // lm.h:802-805
    const int dep_q_1 = dep_q + 1;
    const int needed = ncb - dep_q - 1;
    p->provided = false;
    if (needed > 0) {
        if (!in_tokens || n_in < needed) return fail(MSX_ERR_ARG, "not enough input tokens");   // reference: assert (lm.h:810)
        if (n_in == ncb) {
            for (int i = 0; i < ncb; i++) g->cache[(size_t)((g->offset + c.delays[i]) % CT) * ncb + i] = in_tokens[i];
            p->provided = true;
        } else {
            for (int i = 0; i < needed; i++)
                g->cache[(size_t)((g->offset + c.delays[dep_q_1 + i]) % CT) * ncb + dep_q_1 + i] = in_tokens[i];
        }
    }
    const int pos = g->offset % CT;
    for (int i = 0; i < ncb; i++) p->input[i] = (g->offset <= c.delays[i]) ? g->initial[i] : g->cache[(size_t)pos * ncb + i];
    return 0;
}
static int gen_finish(msx_gen *g, const GenPrep &p, int32_t *out, int depformer_replace_tokens, int32_t *out_text, int32_t *out_audio) {
    const msx_config &c = g->cfg;
    const int CT = g->CT, ncb = g->ncb;
    int dep_q = c.dep_q;
    if (c.personaplex) dep_q = 8;
    const int dep_q_1 = dep_q + 1;
    const int text_token = out[0];
    int32_t *audio = out + 1;
    if (c.dep_q > 0 && g->delay_steps)
        for (int q = 0; q < c.dep_q; q++)
            if (g->offset < c.delays[q + 1] + g->delay_steps) audio[q] = -1;  // lm.h:915-921
    if (c.dep_q > 0 && g->audio_hook) {                                       // audio prefix (lm.h:922-931)
        const int sk = g->audio_hook(g->audio_user, g->offset, audio, c.dep_q);
        if (sk >= 0) g->skip = sk;
    }
    g->offset++;
    if (!p.provided) {
        const int pp = g->offset % CT;
        g->cache[(size_t)pp * ncb + 0] = text_token;
        for (int q = 0; q < c.dep_q; q++) g->cache[(size_t)pp * ncb + q + 1] = audio[q];
    }
    for (int q = 0; q < c.dep_q; q++) out_audio[q] = audio[q];
    if (g->skip > 0) { --g->skip; return 0; }                                 // lm.h:944-947
    if (g->offset <= g->max_delay || depformer_replace_tokens) return 0;
    *out_text = g->cache[(size_t)((g->offset - g->max_delay + c.delays[0]) % CT) * ncb + 0];
    for (int i = 1; i < dep_q_1; i++)
        out_audio[i - 1] = g->cache[(size_t)((g->offset - g->max_delay + c.delays[i]) % CT) * ncb + i];
    for (int q = 0; q < c.dep_q; q++)
        if (out_audio[q] == -1) return 0;
    return 1;
}

extern "C" int msx_gen_step(msx_gen *g, const int32_t *in_tokens, int n_in, int depformer_replace_tokens, int32_t *out_text, int32_t *out_audio) {
    if (!g || !out_text || !out_audio) return fail(MSX_ERR_ARG, "null argument");
    msx_stream *s = g->s;
    const msx_config &c = g->cfg;
    GenPrep p;
    if (int e = gen_prepare(g, in_tokens, n_in, &p)) return e;
    const int32_t *input = p.input;

    int32_t out[1 + MSX_MAX_STEPS];
    for (int i = 0; i < 1 + MSX_MAX_STEPS; i++) out[i] = -1;
    if (!g->fn && (s->temp_text > 0.f || s->temp_audio > 0.f)) {
        // Exp(1) draws exactly like GraphContext::_exponential_compute (context.h:464-480): one libc rand() per
        // top-k candidate, text graph first, then the depformer codebooks in order
        const int kt = std::min(std::min(s->top_k_text, c.text_card), kSampleMaxK), ka = std::min(std::min(s->top_k_audio, c.card), kSampleMaxK);
        std::vector<float> nt(kt), na((size_t)std::max(1, c.dep_q) * ka);
        if (s->temp_text > 0.f) for (int i = 0; i < kt; i++) nt[i] = g->rng.exp1();
        if (s->temp_audio > 0.f && !depformer_replace_tokens) for (size_t i = 0; i < (size_t)c.dep_q * ka; i++) na[i] = g->rng.exp1();
        if (int e = msx_stream_set_noise(s, nt.data(), na.data())) return e;
    }
    if (g->fn) {
        if (int e = g->fn(g->user, input, depformer_replace_tokens, out)) return fail(MSX_ERR_STATE, "step callback failed: " + std::to_string(e));
        if (g->text_hook) out[0] = g->text_hook(g->text_user, g->offset, out[0]);               // (callback mode: applied after the fact)
        if (depformer_replace_tokens) for (int q = 0; q < c.dep_q; q++) out[1 + q] = -1;      // lm.h:909-913
    } else if (g->text_hook) {
        // the text token is rewritten between the two graphs (state machine / text prefix, lm.h:877-899)
        if (int e = msx_step_temporal(s, input, &out[0], nullptr, nullptr)) return e;
        out[0] = g->text_hook(g->text_user, g->offset, out[0]);
        if (c.dep_q > 0 && !depformer_replace_tokens)
            if (int e = msx_step_depformer(s, out[0], nullptr, out + 1, nullptr)) return e;
    } else if (c.dep_q > 0 && !depformer_replace_tokens) {
        if (int e = msx_step(s, input, out)) return e;                        // temporal + depformer, one sync
    } else {
        if (int e = msx_step_temporal(s, input, &out[0], nullptr, nullptr)) return e;
    }
    return gen_finish(g, p, out, depformer_replace_tokens, out_text, out_audio);
}
