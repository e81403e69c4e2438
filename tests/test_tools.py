"""The four LM tools of the reference (tools/moshi-sts.cpp, personaplex.cpp, moshi-tts.cpp, moshi-stt.cpp) rebuilt on
include/moshi/moshi.h + libmoshi.so: command line, error behaviour (CPU) and their main loops against the Python binding (GPU)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from moshi_cpp_b200 import binding as msx, configs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tools():
    msx.build_host()
    d = {t: os.path.join(msx.BIN_DIR, t) for t in msx.TOOLS}
    assert all(os.path.exists(p) for p in d.values())
    return d


def model_dir(tmp_path, gguf_for, preset, quant="q4_k"):
    """a model directory like the reference's: config.json + the weights named after config['moshi_name'] (.gguf)"""
    path, cfg = gguf_for(preset, quant)
    d = tmp_path / preset
    d.mkdir(exist_ok=True)
    with open(d / "config.json", "w") as f:
        json.dump(configs.to_config_json(cfg), f)
    link = d / "model.gguf"
    if not link.exists():
        os.symlink(path, link)
    return str(d), cfg


def silence_codes(n, card):
    l, out = 12345, []
    for _ in range(n):
        l = (l * 1664525 + 1013904223) & 0xFFFFFFFF
        out.append((l >> 8) % card)
    return np.array(out, dtype=np.int32)


def test_tools_command_line_and_errors(tools, gguf_for, tmp_path):
    for name, exe in tools.items():
        r = subprocess.run([exe, "--help"], capture_output=True, text=True)
        assert r.returncode == 0 and "usage:" in r.stderr and "-m PATH" in r.stderr, name
        r = subprocess.run([exe, "--no-such-flag"], capture_output=True, text=True)
        assert r.returncode == 2 and "unknown option" in r.stderr, name
        r = subprocess.run([exe, "-m", str(tmp_path / "nowhere"), "--bench"], capture_output=True, text=True)
        assert r.returncode == 1 and "config.json" in r.stderr, name                # no config -> error, like the reference's exit(1)
    d, cfg = model_dir(tmp_path, gguf_for, "tiny")
    os.remove(os.path.join(d, "model.gguf"))
    r = subprocess.run([tools["moshi-sts"], "-m", d, "--bench"], capture_output=True, text=True)
    assert r.returncode == 1 and "could not open" in r.stderr                       # reference: moshi_lm_from_files -> NULL
    (tmp_path / "b").mkdir()
    d, cfg = model_dir(tmp_path / "b", gguf_for, "tiny")
    r = subprocess.run([tools["moshi-sts"], "-m", d], capture_output=True, text=True)
    if msx.lib().msx_device_count() == 0:
        assert r.returncode == 1 and "error:" in r.stderr                           # no CPU fallback behind the tools either
        r = subprocess.run([tools["moshi-sts"], "-m", d, "-q", "q5_1", "--bench"], capture_output=True, text=True)
        assert r.returncode == 1 and "unknown quantisation" in r.stderr             # moshi_lm_quantize -> false (moshi.cpp:667-668)
    else:
        assert r.returncode == 2 and "--bench" in r.stderr                          # no capture device in this build


def test_reference_style_client_compiles_against_the_header(tmp_path):
    """a translation unit written the way the reference's tools use include/moshi/moshi.h (unref_ptr handles, the full
    moshi_config_t incl. fuser / stt_config / model_id / lm_gen_config, the tokenizer_t* system prompt, Entry, deques)
    compiles against include/moshi/moshi.h and links against libmoshi.so"""
    msx.build_host()
    src = tmp_path / "client.cpp"
    src.write_text(r'''
#include <moshi/moshi.h>
#include <cstdio>
int main(int argc, char **argv) {
    unref_ptr<moshi_context_t> moshi = moshi_alloc(NULL, NULL);
    moshi_config_t config;
    if (argc < 2 || moshi_get_config(&config, argv[1]) != 0) return 3;
    printf("%d %d %f %f %s %d %d %f\n", (int)config.fuser.sum.size(), (int)config.fuser.cross_attention_pos_emb,
           config.stt_config.audio_delay_seconds, config.tts_config.audio_delay, config.model_id.sig.c_str(), (int)config.model_id.epoch,
           (int)config.lm_gen_config.top_k, config.lm_gen_config.temp_text);
    unref_ptr<moshi_lm_t> lm = moshi_lm_from_files(moshi, &config, "/nonexistent/model.gguf");
    if (lm) return 4;                                     // missing file -> NULL
    unref_ptr<tokenizer_t> tok = tokenizer_alloc(argc > 2 ? argv[2] : "/nonexistent", true);
    if (tok) {
        tokenizer_send(tok, "hello world");
        Entry e; int n = 0;
        while (tokenizer_receive(tok, &e)) { n += (int)e.tokens.size(); printf("[%s]", e.text.c_str()); }
        printf(" %d %s\n", n, tokenizer_id_to_piece(tok, 2).c_str());
    }
    std::deque<int> text_prefix; std::deque<std::vector<int>> audio_prefix; std::deque<std::vector<int16_t>> audio_prompt;
    std::vector<int16_t> codes; int text = 0; float vad = 0;
    // the calls below are only type-checked (no generator without a model file)
    if (argc > 99) {
        moshi_lm_gen_t *gen = moshi_lm_generator(lm);
        moshi_lm_quantize(lm, "q4_k"); moshi_lm_load(lm); moshi_lm_save_gguf(lm, "x.gguf"); moshi_lm_set_delay_steps(lm, 2);
        moshi_lm_get_max_delay(lm); moshi_lm_get_delay_steps(lm);
        moshi_lm_set_voice_condition(moshi, gen, "v"); moshi_lm_load_voice_condition(moshi, gen); moshi_lm_voice_prefix(gen, text_prefix, audio_prefix);
        moshi_lm_personaplex_audio_prompt(gen, audio_prompt); moshi_lm_personaplex_load_voice(moshi, gen, "v");
        moshi_lm_personaplex_system_prompt(moshi, gen, tok, "You are helpful.");
        moshi_lm_start(moshi, gen, 0.8f, 0.7f); Entry e; moshi_lm_send(gen, &e);
        moshi_lm_send2(gen, codes); moshi_lm_receive(gen, text, codes); moshi_lm_receive2(gen, text, vad);
        moshi_lm_is_active(gen); moshi_lm_is_empty(gen); moshi_lm_machine_reset(gen); unref(gen);
    }
    return 0;
}
''')
    exe = tmp_path / "client"
    subprocess.check_call(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-o", str(exe), str(src),
                           "-L" + os.path.dirname(msx.HOST_SO_PATH), "-lmoshi", "-lmoshi_b200", "-Wl,-rpath," + os.path.dirname(msx.HOST_SO_PATH)])
    cfg = tmp_path / "config.json"
    cfg.write_text(json.dumps({"card": 2048, "n_q": 16, "dep_q": 8, "delays": [0] * 17, "dim": 512, "text_card": 1000,
                               "fuser": {"cross_attention_pos_emb": True, "cross_attention_pos_emb_scale": 1.0, "sum": ["control", "cfg"], "prepend": [], "cross": ["speaker_wavs"]},
                               "stt_config": {"audio_delay_seconds": 0.5, "audio_silence_prefix_seconds": 0.0},
                               "tts_config": {"audio_delay": 1.28, "second_stream_ahead": 2},
                               "model_id": {"sig": "1e68beda", "epoch": 240}, "lm_gen_config": {"temp": 0.6, "temp_text": 0.6, "top_k": 250, "top_k_text": 50}}))
    vocab = tmp_path / "vocab.txt"
    vocab.write_text("<unk>\n<s>\n</s>\n▁hello\n▁wor\nld\n")
    r = subprocess.run([str(exe), str(cfg), str(vocab)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.splitlines()[0] == "2 1 0.500000 1.280000 1e68beda 240 250 0.600000"
    assert r.stdout.splitlines()[1] == "[hello][world] 4 </s>"              # <s> + ▁hello | ▁wor + ld


@pytest.mark.gpu
@pytest.mark.parametrize("tool,preset", [("moshi-sts", "tiny"), ("personaplex", "tiny_pplex")])
def test_sts_tools_match_the_binding(tools, gguf_for, tmp_path, tool, preset):
    """--bench loop of moshi-sts / personaplex (send2 -> receive per frame on the fixed silence codes, greedy) == the same frames
    through the Python generator; then the same through a .mimi token file in and out"""
    d, cfg = model_dir(tmp_path, gguf_for, preset)
    import re
    n_user = cfg["n_q"] - (8 if cfg["model_type"] == "personaplex" else cfg["dep_q"])      # PersonaPlex: dep_q 16, 8 user codebooks (lm.h:802-805)
    frames = 24
    r = subprocess.run([tools[tool], "-m", d, "--bench", "--frames", str(frames), "-t", "0", "--print-tokens"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    rows = [l for l in r.stdout.splitlines() if re.match(r"^-?\d+:", l)]
    got = [[int(l.split(":")[0])] + [int(v) for v in l.split(":")[1].split()] for l in rows]
    assert len(got) == frames and "frames/s" in r.stdout
    gm = msx.Model(os.path.join(d, "model.gguf"), cfg); gs = msx.Stream(gm); gen = msx.Gen(gs)
    user = silence_codes(n_user, cfg["card"])
    if cfg["model_type"] == "personaplex":      # moshi_lm_start replays the system prompts first (lm.h:1120-1134): 6 + 6 silence rows
        PT = [3, 948, 243, 1178, 546, 1736, 1030, 1978, 2008, 430, 1268, 381, 1611, 1095, 1495, 56, 472]
        row = np.array([PT[0]] + [v % cfg["card"] for v in PT[1:]], dtype=np.int32)
        for _ in range(12):
            gen.step(row)
    exp = []
    while len(exp) < frames:
        ok, t, a = gen.step(user)
        if ok:
            exp.append([t] + [int(v) for v in a])
    w = min(len(got[0]), len(exp[0]))
    assert [g[:w] for g in got] == [e[:w] for e in exp]
    # token files: the generated codes of run 1 become the user stream of run 2 (any int16 frames do)
    inp = tmp_path / "in.mimi"; outp = tmp_path / "out.mimi"
    np.tile(user.astype(np.int16), (10, 1)).tofile(inp)
    r = subprocess.run([tools[tool], "-m", d, "-i", str(inp), "-o", str(outp), "-t", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = np.fromfile(outp, dtype=np.int16)
    n_emit = 10 - (0 if cfg["model_type"] == "personaplex" else max(cfg["delays"]))      # the PersonaPlex prompt frames already filled the delay window
    assert out.size == n_emit * (len(got[0]) - 1)
    assert out.reshape(n_emit, -1).tolist() == [g[1:] for g in got[:n_emit]]


@pytest.mark.gpu
def test_tts_and_stt_tools_run(tools, gguf_for, tmp_path):
    d, cfg = model_dir(tmp_path, gguf_for, "tiny_tts")
    r = subprocess.run([tools["moshi-tts"], "-m", d, "--bench", "--frames", "40", "-t", "0", "-o", str(tmp_path / "tts.mimi")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "token count:" in r.stdout and "frames/s" in r.stdout
    codes = np.fromfile(tmp_path / "tts.mimi", dtype=np.int16)
    assert codes.size == 40 * cfg["dep_q"] and codes.min() >= 0 and codes.max() < cfg["card"]
    d, cfg = model_dir(tmp_path, gguf_for, "tiny_stt", "q8_0")
    r = subprocess.run([tools["moshi-stt"], "-m", d, "--bench", "--frames", "30", "--debug"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l and (l[0].isdigit() or l[0] == "*")]
    assert len(lines) == 30 and all(0.0 <= float(l.lstrip("*").split()[0]) <= 1.0 for l in lines)
