// lm_tool.h — shared plumbing of the four LM tools (moshi-sts, personaplex, moshi-tts, moshi-stt) built on include/moshi/moshi.h.
//
// The reference's tools (tools/moshi-sts.cpp, personaplex.cpp, moshi-tts.cpp, moshi-stt.cpp) wrap the LM step between the Mimi
// codec, SDL capture / playback and FFmpeg files.  Those are outside this repository's scope,
// so the tools here keep the command line (-m -q -g -c -s -t -d -b -v -p -i -o) and the LM call sequence of the reference's
// main loops and replace the audio side by Mimi TOKEN frames: `-i file.mimi` reads frames of n_q little-endian int16 codes
// (the format the reference's moshi-tts writes with `-o x.mimi`), `--bench` feeds a fixed "silence" code row like the
// reference's --bench feeds encoded silence, `-o file.mimi` writes the generated codes.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <moshi/moshi.h>

struct LmToolArgs {
    std::string model = ".", quant, save_gguf, input, output, voice, prompt;
    int device = 0, context = -1, seed = 0, frames = 125, delay = 0;
    float depth_temperature = 0.8f, text_temperature = 0.7f;
    bool bench = false, debug = false, print_tokens = false;
    bool ok = true;
};

inline void lm_tool_usage(const char *prog, const char *what) {
    fprintf(stderr,
            "usage: %s [options]      %s\n"
            "  -m PATH        model directory (config.json + *.gguf) or a .gguf file with config.json next to it\n"
            "  -q QUANT       quantise an unquantised file while loading: q8_0 | q4_k\n"
            "  -g FILE        write the (quantised) weights as a GGUF and exit\n"
            "  -c N           context (ring slots) instead of the config's\n"
            "  -s SEED        seed of the sampler       -t DEPTH,TEXT  temperatures (0 = greedy)\n"
            "  -d N           CUDA device\n"
            "  -b, --bench    run --frames frames on synthetic input and print the frame rate\n"
            "  --frames N     frames of a bench run (default 125 = 10 s)\n"
            "  -i FILE.mimi   input token frames (n_q int16 per frame)      -o FILE.mimi   output token frames\n"
            "  -v FILE        voice (.safetensors / .gguf)                  -p TEXT        text / system prompt\n"
            "  --print-tokens print the generated tokens\n", prog, what);
}

inline LmToolArgs lm_tool_parse(int argc, char **argv, const char *what) {
    LmToolArgs a;
    a.seed = (int)time(nullptr);
    for (int i = 1; i < argc; i++) {
        const std::string s = argv[i];
        auto need = [&](const char *name) -> const char * {
            if (i + 1 >= argc) { fprintf(stderr, "error: \"%s\" requires a value\n", name); a.ok = false; return ""; }
            return argv[++i];
        };
        if (s == "-h" || s == "--help") { lm_tool_usage(argv[0], what); exit(0); }
        else if (s == "-m" || s == "--model") a.model = need("-m");
        else if (s == "-q" || s == "--quantize") a.quant = need("-q");
        else if (s == "-g" || s == "--gguf") a.save_gguf = need("-g");
        else if (s == "-c" || s == "--context") a.context = atoi(need("-c"));
        else if (s == "-s" || s == "--seed") a.seed = atoi(need("-s"));
        else if (s == "-d" || s == "--device") a.device = atoi(need("-d"));
        else if (s == "-t" || s == "--temperature") {
            const char *v = need("-t");
            if (sscanf(v, "%f,%f", &a.depth_temperature, &a.text_temperature) < 2) a.text_temperature = a.depth_temperature;
        }
        else if (s == "-b" || s == "--bench") a.bench = true;
        else if (s == "--frames") a.frames = atoi(need("--frames"));
        else if (s == "--delay") a.delay = atoi(need("--delay"));
        else if (s == "-i" || s == "--input") a.input = need("-i");
        else if (s == "-o" || s == "--output") a.output = need("-o");
        else if (s == "-v" || s == "--voice") a.voice = need("-v");
        else if (s == "-p" || s == "--prompt") a.prompt = need("-p");
        else if (s == "--debug") a.debug = true;
        else if (s == "--print-tokens") a.print_tokens = true;
        else if (s == "--threads" || s == "-r") (void)need(s.c_str());      // accepted for command-line compatibility, unused
        else { fprintf(stderr, "error: unknown option %s\n", s.c_str()); a.ok = false; }
    }
    return a;
}

inline bool lm_tool_file_exists(const std::string &p) { FILE *f = fopen(p.c_str(), "rb"); if (f) fclose(f); return f != nullptr; }
inline bool lm_tool_ends_with(const std::string &s, const char *t) { const size_t n = strlen(t); return s.size() >= n && s.compare(s.size() - n, n, t) == 0; }

// -m: a directory with config.json and the weights, or the weights file itself
inline bool lm_tool_locate(const LmToolArgs &a, moshi_config_t *config, std::string *weights, std::string *dir) {
    std::string d = a.model, w;
    if (lm_tool_ends_with(d, ".gguf") || lm_tool_ends_with(d, ".safetensors")) {
        w = d;
        const size_t slash = d.find_last_of('/');
        d = slash == std::string::npos ? "." : d.substr(0, slash);
    }
    if (!d.empty() && d.back() != '/') d += '/';
    std::string cfg_path = d + "config.json";
    for (const char *alt : {"moshi-config.json", "personaplex-config.json"}) if (!lm_tool_file_exists(cfg_path)) cfg_path = d + alt;
    if (moshi_get_config(config, cfg_path.c_str()) != 0) { fprintf(stderr, "error: no readable config.json in %s\n", d.c_str()); return false; }
    if (w.empty()) {
        std::string base = config->moshi_name;
        const size_t dot = base.find_last_of('.');
        const std::string gguf = d + (dot == std::string::npos ? base : base.substr(0, dot)) + ".gguf";
        w = lm_tool_file_exists(gguf) ? gguf : d + base;
    }
    *weights = w; *dir = d;
    return true;
}

struct LmToolModel {
    moshi_config_t config;
    unref_ptr<moshi_context_t> moshi;
    unref_ptr<moshi_lm_t> lm;
    unref_ptr<moshi_lm_gen_t> gen;
    std::string dir;
};

// config -> context -> lm (-q, -g) -> load -> generator: the set-up every reference tool performs before its main loop.
// Returns 0 to continue, 1 on error, -1 when `-g` wrote the GGUF and the tool should exit successfully.
inline int lm_tool_open(const LmToolArgs &a, LmToolModel *m) {
    std::string weights;
    if (!lm_tool_locate(a, &m->config, &weights, &m->dir)) return 1;
    if (a.context > 0) m->config.context = a.context;                          // tools/moshi-sts.cpp `-c`
    m->moshi = moshi_alloc_b200(a.device);
    m->lm = moshi_lm_from_files(m->moshi, &m->config, weights.c_str());
    if (!m->lm) { fprintf(stderr, "error: could not open %s\n", weights.c_str()); return 1; }
    if (!a.quant.empty() && !moshi_lm_quantize(m->lm, a.quant.c_str())) { fprintf(stderr, "error: unknown quantisation %s\n", a.quant.c_str()); return 1; }
    if (!a.save_gguf.empty()) { moshi_lm_save_gguf(m->lm, a.save_gguf.c_str()); return -1; }
    if (a.delay > 0) moshi_lm_set_delay_steps(m->lm, a.delay);
    if (moshi_lm_load(m->lm) != 0) { fprintf(stderr, "error: %s\n", moshi_b200_last_error()); return 1; }
    m->gen = moshi_lm_generator(m->lm);
    return 0;
}

// token-frame files: n int16 codes per frame, frames back to back
struct MimiTokenReader {
    FILE *f = nullptr; int n = 0;
    bool open(const std::string &path, int n_codes) { f = fopen(path.c_str(), "rb"); n = n_codes; return f != nullptr; }
    bool next(std::vector<int16_t> &codes) { codes.resize(n); return f && fread(codes.data(), 2, (size_t)n, f) == (size_t)n; }
    ~MimiTokenReader() { if (f) fclose(f); }
};
struct MimiTokenWriter {
    FILE *f = nullptr;
    bool open(const std::string &path) { f = fopen(path.c_str(), "wb"); return f != nullptr; }
    void put(const std::vector<int16_t> &codes) { if (f) fwrite(codes.data(), 2, codes.size(), f); }
    ~MimiTokenWriter() { if (f) fclose(f); }
};
// what --bench feeds instead of encoded silence (no Mimi encoder here): one fixed pseudo-random row of codes
inline std::vector<int16_t> lm_tool_silence_codes(int n, int card) {
    std::vector<int16_t> v((size_t)n);
    uint32_t l = 12345u;
    for (auto &c : v) { l = l * 1664525u + 1013904223u; c = (int16_t)((l >> 8) % (uint32_t)card); }
    return v;
}

struct LmToolClock {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double seconds() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};
inline void lm_tool_report(const char *tool, long frames, double seconds) {
    printf("\n%s: %ld frames in %.3f s = %.1f frames/s = %.1fx real time (12.5 Hz)\n", tool, frames, seconds, frames / seconds, frames / seconds / 12.5);
}
inline void lm_tool_print_piece(tokenizer_t *tok, int text_token) {
    if (!tok || text_token == 0 || text_token == 3) return;
    std::string piece = tokenizer_id_to_piece(tok, text_token), text;
    for (size_t i = 0; i < piece.size(); i++) {
        if ((unsigned char)piece[i] == 0xE2 && i + 2 < piece.size()) { text += ' '; i += 2; continue; }     // U+2581
        text += piece[i];
    }
    fputs(text.c_str(), stdout); fflush(stdout);
}
