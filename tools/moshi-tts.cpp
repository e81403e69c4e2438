// moshi-tts — text-to-speech LM loop (reference: tools/moshi-tts.cpp:745-828): words are queued as Entry objects
// (tokenizer_send / tokenizer_receive / moshi_lm_send, at least 4 ahead for the state machine's look-ahead), frames are pulled
// with moshi_lm_receive while moshi_lm_is_active.  -v gives the voice (.safetensors with speaker embeddings, through the
// model's own conditioners); without a voice file a fixed synthetic conditioning memory is used (bench runs).
#include "lm_tool.h"

int main(int argc, char **argv) {
    LmToolArgs a = lm_tool_parse(argc, argv, "text-to-speech step (words in, audio codes out)");
    if (!a.ok) return 2;
    LmToolModel m;
    if (const int rc = lm_tool_open(a, &m)) return rc < 0 ? 0 : 1;
    unref_ptr<tokenizer_t> tok = tokenizer_alloc((m.dir + m.config.tokenizer_name).c_str());
    if (!a.voice.empty()) {
        const int v = moshi_lm_set_voice_condition(m.moshi, m.gen, a.voice.c_str());
        const int l = v == 0 ? moshi_lm_load_voice_condition(m.moshi, m.gen) : 0;
        if (v != 0 || l != 0) { fprintf(stderr, "error: voice condition (%d, %d): %s\n", v, l, moshi_b200_last_error()); return 1; }
    } else {
        const int tc = 125, dim = (int)m.config.dim;                   // 5 x 25 rows, the shape a 25-frame speaker embedding gives
        std::vector<float> sum((size_t)dim), cross((size_t)tc * dim);
        uint32_t l = 7;
        auto rnd = [&]() { l = l * 1664525u + 1013904223u; return ((l >> 8) % 2001) / 1000.f - 1.f; };
        for (float &v : sum) v = 0.2f * rnd();
        for (float &v : cross) v = rnd();
        if (moshi_lm_set_condition(m.gen, sum.data(), m.config.cross_attention ? cross.data() : nullptr, tc) != 0) { fprintf(stderr, "error: %s\n", moshi_b200_last_error()); return 1; }
    }
    srand((unsigned)a.seed);
    moshi_lm_start(m.moshi, m.gen, a.depth_temperature, a.text_temperature);
    MimiTokenWriter out;
    if (!a.output.empty() && !out.open(a.output)) { fprintf(stderr, "error: cannot open %s\n", a.output.c_str()); return 1; }

    // the text: -p, or for --bench a script of pseudo-random "words" (token ids straight into Entry objects)
    std::deque<Entry> script;
    if (!a.prompt.empty() && tok) { tokenizer_send(tok, a.prompt); Entry e; while (tokenizer_receive(tok, &e)) script.push_back(e); }
    else {
        if (!a.prompt.empty()) fprintf(stderr, "warning: no tokenizer vocabulary next to the model: synthetic words are spoken instead\n");
        uint32_t l = 99;
        const int words = a.bench ? 24 : 6;
        for (int w = 0; w < words; w++) {
            Entry e;
            l = l * 1664525u + 1013904223u;
            const int nt = 1 + (int)((l >> 10) % 3);
            for (int i = 0; i < nt; i++) { l = l * 1664525u + 1013904223u; e.tokens.push_back(4 + (int)((l >> 8) % (uint32_t)(m.config.text_card - 4))); }
            e.text = "w" + std::to_string(w); e.padding = (int)((l >> 20) & 1);
            script.push_back(e);
        }
    }
    std::vector<int16_t> codes;
    int text_token = 0;
    long frames = 0, tokens_sent = 0;
    bool active = true;
    LmToolClock clock;
    while (active) {
        active = false;
        for (int i = 0; i < 4 && !script.empty(); i++) {                // at least 4 entries ahead (moshi-tts.cpp:769-779)
            moshi_lm_send(m.gen, &script.front()); script.pop_front();
            tokens_sent++; active = true;
        }
        const int rc = moshi_lm_receive(m.gen, text_token, codes);
        if (rc < 0) return 1;
        if (rc) {
            frames++;
            out.put(codes);
            if (a.print_tokens) { printf("%d:", text_token); for (int16_t t : codes) printf(" %d", t); printf("\n"); }
        }
        if (moshi_lm_is_active(m.gen)) active = true;
        if (a.bench && frames >= a.frames) break;
    }
    printf("token count: %4ld tokens\nframe count: %4ld frames\n", tokens_sent, frames);
    lm_tool_report("moshi-tts", frames, clock.seconds());
    return 0;
}
