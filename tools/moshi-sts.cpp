// moshi-sts — speech-to-speech LM loop on the B200 engine (reference: tools/moshi-sts.cpp:90-836; main loop :731-808).
// Per frame: user codes -> moshi_lm_send2 -> moshi_lm_receive -> text token + dep_q audio codes.  With a PersonaPlex model
// (-v voice, -p prompt) it is the reference's `personaplex` tool; see personaplex.cpp for that entry point.
#include "lm_tool.h"

int sts_main(int argc, char **argv, bool personaplex_tool) {
    LmToolArgs a = lm_tool_parse(argc, argv, personaplex_tool ? "PersonaPlex full-duplex step" : "speech-to-speech step (user codes in, text + audio codes out)");
    if (!a.ok) return 2;
    LmToolModel m;
    if (const int rc = lm_tool_open(a, &m)) return rc < 0 ? 0 : 1;
    unref_ptr<tokenizer_t> tok = tokenizer_alloc((m.dir + m.config.tokenizer_name).c_str());       // NULL without a vocabulary listing: ids are printed
    const bool pplex = m.config.model_type == "personaplex";
    // the user stream: Mimi's 8 codebooks; PersonaPlex runs dep_q = 16 depformer steps but still takes 8 user codes (lm.h:802-805)
    const int n_user = pplex ? 8 : (int)(m.config.n_q - m.config.dep_q);
    if (pplex) {
        // tools/personaplex.cpp / moshi-sts.cpp:575-660: voice prompt (embeddings + token ring), then the text system prompt
        if (!a.voice.empty() && moshi_lm_personaplex_load_voice(m.moshi, m.gen, a.voice.c_str()) != 0) { fprintf(stderr, "error: could not load voice %s\n", a.voice.c_str()); return 1; }
        if (!a.prompt.empty()) {
            if (tok) moshi_lm_personaplex_system_prompt(m.moshi, m.gen, tok, a.prompt.c_str());
            else fprintf(stderr, "warning: no tokenizer vocabulary next to the model: the system prompt is skipped\n");
        }
    }
    srand((unsigned)a.seed);
    moshi_lm_start(m.moshi, m.gen, a.depth_temperature, a.text_temperature);
    MimiTokenReader in; MimiTokenWriter out;
    if (!a.input.empty() && !in.open(a.input, n_user)) { fprintf(stderr, "error: cannot open %s\n", a.input.c_str()); return 1; }
    if (!a.output.empty() && !out.open(a.output)) { fprintf(stderr, "error: cannot open %s\n", a.output.c_str()); return 1; }
    if (a.input.empty() && !a.bench) { fprintf(stderr, "error: give -i FILE.mimi or --bench (audio capture is not part of this build)\n"); return 2; }
    const std::vector<int16_t> silence = lm_tool_silence_codes(n_user, (int)m.config.card);
    std::vector<int16_t> tokens;
    int text_token = 0;
    long frames = 0;
    printf("ready\n");
    LmToolClock clock;
    while (true) {
        if (!a.input.empty()) { if (!in.next(tokens)) break; }
        else tokens = silence;
        moshi_lm_send2(m.gen, tokens);
        const int rc = moshi_lm_receive(m.gen, text_token, tokens);
        if (rc < 0) return 1;
        if (rc) {
            out.put(tokens);
            frames++;
            if (a.print_tokens) { printf("%d:", text_token); for (int16_t t : tokens) printf(" %d", t); printf("\n"); }
            else lm_tool_print_piece(tok, text_token);
            if (a.bench && frames >= a.frames) break;
        }
    }
    lm_tool_report(personaplex_tool ? "personaplex" : "moshi-sts", frames, clock.seconds());
    return 0;
}

#ifndef MOSHI_TOOL_NO_MAIN
int main(int argc, char **argv) { return sts_main(argc, argv, false); }
#endif
