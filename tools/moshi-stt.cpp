// moshi-stt — speech-to-text LM loop (reference: tools/moshi-stt.cpp:544-726): per frame n_q user codes -> moshi_lm_send2 ->
// moshi_lm_receive2 -> text token + voice-activity probability of extra head 2.
#include "lm_tool.h"

int main(int argc, char **argv) {
    LmToolArgs a = lm_tool_parse(argc, argv, "speech-to-text step (audio codes in, text tokens + VAD out)");
    if (!a.ok) return 2;
    LmToolModel m;
    if (const int rc = lm_tool_open(a, &m)) return rc < 0 ? 0 : 1;
    unref_ptr<tokenizer_t> tok = tokenizer_alloc((m.dir + m.config.tokenizer_name).c_str());
    // greedy like the reference's STT (it starts the generator with temperature 0)
    moshi_lm_start(m.moshi, m.gen, 0.f, 0.f);
    const int n_q = (int)m.config.n_q;
    MimiTokenReader in;
    if (!a.input.empty() && !in.open(a.input, n_q)) { fprintf(stderr, "error: cannot open %s\n", a.input.c_str()); return 1; }
    if (a.input.empty() && !a.bench) { fprintf(stderr, "error: give -i FILE.mimi or --bench (audio capture is not part of this build)\n"); return 2; }
    const std::vector<int16_t> silence = lm_tool_silence_codes(n_q, (int)m.config.card);
    std::vector<int16_t> tokens;
    long frames = 0;
    // the reference appends audio_delay_seconds of silence so that the delayed text stream can finish (stt_config)
    long tail = (long)(m.config.stt_config.audio_delay_seconds * 12.5f) + 1;
    LmToolClock clock;
    while (true) {
        if (!a.input.empty()) { if (!in.next(tokens)) { if (tail-- <= 0) break; tokens = silence; } }
        else { if (frames >= a.frames) break; tokens = silence; }
        moshi_lm_send2(m.gen, tokens);
        int text_token = 0; float vad = 0.f;
        moshi_lm_receive2(m.gen, text_token, vad);
        frames++;
        if (a.debug || a.print_tokens) printf("%s%f %d\n", vad > 0.5f ? "*" : "", vad, text_token);
        else lm_tool_print_piece(tok, text_token);
    }
    lm_tool_report("moshi-stt", frames, clock.seconds());
    return 0;
}
