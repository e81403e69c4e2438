// moshi_sts_bench.cpp — the `--bench` entry point of the reference's speech-to-speech tools
// (tools/moshi-sts.cpp:731-808, tools/personaplex.cpp) reduced to the LM path: the Mimi encoder/decoder and
// SDL/FFmpeg I/O are out of scope, so user audio codes are synthetic (seeded LCG) instead of encoded silence.
//   moshi-sts-bench <model.gguf> <config.json> [frames=125] [device=0] [--print-tokens] [-q q8_0|q4_k] [-g out.gguf] [--voice v.safetensors] [--pplex-voice v.safetensors|v.gguf]
// -q quantises an unquantised (f32 / f16 / bf16) GGUF while loading, -g writes the quantised weights as a GGUF and exits
// (tools/moshi-sts.cpp `-q`, `-g`: moshi_lm_quantize, moshi_lm_save_gguf).
// For a TTS model (model_type "tts": no user stream, cross-attention conditioning) it runs the moshi-tts --bench loop
// instead (tools/moshi-tts.cpp:770-781): a synthetic conditioning memory, a script of LCG "words" sent as Entry
// objects, receive() while moshi_lm_is_active().
// Uses only the moshi_lm_* API (host/moshi_api.h), exactly like the reference tool uses include/moshi/moshi.h.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "moshi_api.h"

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s model.gguf config.json [frames] [device] [--print-tokens]\n", argv[0]); return 2; }
    int frames = 125, device = 0, positional = 0;
    bool print_tokens = false;
    const char *quant = nullptr, *save_path = nullptr, *voice_path = nullptr, *pplex_voice = nullptr;
    for (int i = 3; i < argc; i++) {
        if (!strcmp(argv[i], "--print-tokens")) print_tokens = true;
        else if (!strcmp(argv[i], "-q") && i + 1 < argc) quant = argv[++i];
        else if (!strcmp(argv[i], "-g") && i + 1 < argc) save_path = argv[++i];
        else if (!strcmp(argv[i], "--voice") && i + 1 < argc) voice_path = argv[++i];
        else if (!strcmp(argv[i], "--pplex-voice") && i + 1 < argc) pplex_voice = argv[++i];
        else if (argv[i][0] != '-') { (positional == 0 ? frames : device) = atoi(argv[i]); positional++; }
    }

    moshi_config_t config;
    if (moshi_get_config(&config, argv[2]) != 0) return 1;
    moshi_context_t *moshi = moshi_alloc_b200(device);
    moshi_lm_t *lm = moshi_lm_from_files(moshi, &config, argv[1]);
    if (!lm) { fprintf(stderr, "error: could not open %s\n", argv[1]); return 1; }
    if (quant && !moshi_lm_quantize(lm, quant)) { fprintf(stderr, "error: unknown quantisation %s\n", quant); return 1; }
    if (save_path) { moshi_lm_save_gguf(lm, save_path); unref(lm); return 0; }
    if (moshi_lm_load(lm) != 0) { fprintf(stderr, "error: %s\n", moshi_b200_last_error()); return 1; }
    moshi_lm_gen_t *gen = moshi_lm_generator(lm);
    if (config.model_type == "tts") {
        // conditioning tensors the conditioners would produce (moshi.cpp:296-366): seeded, deterministic
        const int tc = 11, dim = (int)config.dim;
        std::vector<float> sum(dim), cross((size_t)tc * dim);
        uint32_t l2 = 7;
        auto rnd = [&]() { l2 = l2 * 1664525u + 1013904223u; return ((l2 >> 8) % 2001) / 1000.f - 1.f; };
        for (float &v : sum) v = 0.2f * rnd();
        for (float &v : cross) v = rnd();
        if (voice_path) {            // tools/moshi-tts.cpp: a voice file through the model's own conditioners
            const int a = moshi_lm_set_voice_condition(moshi, gen, voice_path);
            const int b = a == 0 ? moshi_lm_load_voice_condition(moshi, gen) : 0;
            if (a != 0 || b != 0) { fprintf(stderr, "error: voice condition (%d, %d)\n", a, b); return 1; }
        } else if (moshi_lm_set_condition(gen, sum.data(), config.cross_attention ? cross.data() : nullptr, tc) != 0) { fprintf(stderr, "error: set_condition\n"); return 1; }
        moshi_lm_start(moshi, gen, 0.f, 0.f);
        uint32_t l3 = 99;
        for (int w = 0; w < 6; w++) {                                // six "words" of 1-3 tokens, padding 0-1
            Entry e;
            l3 = l3 * 1664525u + 1013904223u;
            const int nt = 1 + (int)((l3 >> 8) % 3);
            for (int i = 0; i < nt; i++) { l3 = l3 * 1664525u + 1013904223u; e.tokens.push_back(4 + (int)((l3 >> 8) % (config.text_card - 4))); }
            e.padding = w % 2;
            moshi_lm_send(gen, &e);
        }
        std::vector<int16_t> audio;
        int text = -1, f = 0, emitted = 0;
        const auto t0 = std::chrono::steady_clock::now();
        while (moshi_lm_is_active(gen) && f < frames) {
            const int ok = moshi_lm_receive(gen, text, audio);
            if (ok) emitted++;
            if (print_tokens) { printf("%d %d %d", f, ok, ok ? text : -1); if (ok) for (int16_t a : audio) printf(" %d", (int)a); printf("\n"); }
            f++;
        }
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "tts frames %d emitted %d empty %d  %.2f frames/s\n", f, emitted, moshi_lm_is_empty(gen), f / sec);
        unref(gen); unref(lm); unref(moshi);
        return 0;
    }
    if (config.model_type == "personaplex") {
        std::deque<std::vector<int16_t>> voice;                      // 4 frames of synthetic voice-prompt codes
        for (int f = 0; f < 4; f++) { std::vector<int16_t> c(8); for (int j = 0; j < 8; j++) c[j] = (int16_t)((f * 131 + j * 17) % config.card); voice.push_back(c); }
        if (pplex_voice) {           // personaplex.cpp: a pre-computed voice (prompt embeddings + token ring) instead of prompt audio
            if (moshi_lm_personaplex_load_voice(moshi, gen, pplex_voice) != 0) { fprintf(stderr, "error: could not load voice %s\n", pplex_voice); return 1; }
        } else
        moshi_lm_personaplex_audio_prompt(gen, voice);
        moshi_lm_personaplex_system_prompt_tokens(gen, {5, 17, 99, 250});
    }
    moshi_lm_start(moshi, gen, 0.f, 0.f);                            // temperature 0 = greedy

    const int n_user = (int)(config.n_q - (config.model_type == "personaplex" ? 8 : config.dep_q));
    std::vector<int16_t> user(n_user), audio;
    uint32_t lcg = 42;
    int text = -1, emitted = 0;
    const auto t0 = std::chrono::steady_clock::now();
    for (int f = 0; f < frames; f++) {
        for (int j = 0; j < n_user; j++) { lcg = lcg * 1664525u + 1013904223u; user[j] = (int16_t)((lcg >> 8) % config.card); }
        moshi_lm_send2(gen, user);
        if (config.dep_q > 0) {
            const int ok = moshi_lm_receive(gen, text, audio);
            if (ok) emitted++;
            if (print_tokens) { printf("%d %d %d", f, ok, ok ? text : -1); if (ok) for (int16_t a : audio) printf(" %d", (int)a); printf("\n"); }
        } else {
            float vad = 0.f;
            moshi_lm_receive2(gen, text, vad);
            if (print_tokens) printf("%d %d %.6f\n", f, text, vad);
        }
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "frames %d emitted %d  %.2f frames/s (%.1fx real time at 12.5 Hz)\n", frames, emitted, frames / sec, frames / sec / 12.5);
    unref(gen); unref(lm); unref(moshi);
    return 0;
}
