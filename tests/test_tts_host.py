"""TTS host logic: the C++ state machine of the moshi_lm_* mirror against an independent
restatement of lm.h:55-194, and the moshi-tts --bench style loop of the C++ API against the same loop written in
Python over the C ABI."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from moshi_cpp_b200 import binding as msx, configs
from tts_machine_ref import Machine


@pytest.fixture(scope="module")
def host():
    msx.build()
    msx.build_host()
    L = C.CDLL(msx.HOST_SO_PATH)
    L.moshi_tts_machine_new.restype = C.c_void_p
    L.moshi_tts_machine_new.argtypes = [C.c_int] * 4
    for f in ("free", "reset"):
        getattr(L, "moshi_tts_machine_" + f).argtypes = [C.c_void_p]
    L.moshi_tts_machine_push.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.moshi_tts_machine_process.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.moshi_tts_machine_end_step.argtypes = [C.c_void_p]
    L.moshi_tts_machine_is_empty.argtypes = [C.c_void_p]
    return L


@pytest.mark.parametrize("ahead", [0, 1, 2])
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_state_machine_matches_restatement(host, ahead, seed):
    rng = np.random.default_rng(seed * 10 + ahead)
    card = 501
    h = host.moshi_tts_machine_new(card, ahead, 8, 2)
    ref = Machine(card, ahead, 8, 2)
    step = 0
    for round_ in range(3):
        for w in range(int(rng.integers(0, 7))):
            toks = [int(t) for t in rng.integers(4, 500, size=int(rng.integers(0, 4)))]     # empty entries = pure pauses
            pad = int(rng.integers(0, 3))
            arr = (C.c_int * max(1, len(toks)))(*toks)
            host.moshi_tts_machine_push(h, arr, len(toks), pad)
            ref.push(toks, pad)
        for _ in range(40):
            tok = int(rng.choice([0, 3, 3, 3, 17, -1, 499]))          # what the model proposes: mostly PAD / NEW_WORD
            a = host.moshi_tts_machine_process(h, step, tok)
            b = ref.process(step, tok)
            assert a == b, f"step {step}: {a} != {b}"
            assert host.moshi_tts_machine_end_step(h) == ref.end_step
            assert bool(host.moshi_tts_machine_is_empty(h)) == ref.is_empty()
            step += 1
        if round_ == 1:
            host.moshi_tts_machine_reset(h); ref.reset()
    host.moshi_tts_machine_free(h)


@pytest.mark.gpu
@pytest.mark.parametrize("voice", [False, True])
def test_tts_loop_cpp_api_vs_c_abi(gguf_for, tmp_path, voice):
    """moshi_lm_set_condition / start / send(Entry) / receive-while-is_active through the C++ mirror (the tool) ==
    the same loop in Python: delay ring (lm.h:796-979, no user stream), state machine between the two graphs."""
    tool = msx.STS_BENCH
    msx.build_host()
    path, cfg = gguf_for("tiny_tts_voice" if voice else "tiny_tts", "q4_k")
    cj = tmp_path / "config.json"
    with open(cj, "w") as f:
        json.dump(configs.to_config_json(cfg), f)
    extra = []
    if voice:      # moshi_lm_set_voice_condition + moshi_lm_load_voice_condition (moshi.cpp:729-760): a voice file, the model's conditioners
        from moshi_cpp_b200 import synth
        wavs = np.random.default_rng(8).standard_normal((synth.COND_CHANNELS, 7)).astype(np.float32)
        vp = str(tmp_path / "voice.safetensors")
        synth.write_safetensors(vp, [("speaker_wavs", "F32", [1, synth.COND_CHANNELS, 7], wavs.tobytes())])
        extra = ["--voice", vp]
        bad = subprocess.run([tool, gguf_for("tiny_tts", "q4_k")[0], str(cj), "4", "0", "--voice", vp], capture_output=True, text=True, timeout=300)
        assert bad.returncode == 1 and "(0, -2)" in bad.stderr          # no conditioner tensors in that GGUF
    r = subprocess.run([tool, path, str(cj), "60", "0", "--print-tokens"] + extra, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = [[int(v) for v in l.split()] for l in r.stdout.strip().splitlines()]

    # the tool's deterministic inputs
    def lcg(state):
        while True:
            state[0] = (state[0] * 1664525 + 1013904223) & 0xFFFFFFFF
            yield state[0]
    dim, tc = cfg["dim"], 11
    g2 = lcg([7])
    rnd = lambda: np.float32(((next(g2) >> 8) % 2001) / np.float32(1000.0)) - np.float32(1.0)
    cs = np.array([np.float32(0.2) * rnd() for _ in range(dim)], dtype=np.float32)
    cc = np.array([rnd() for _ in range(tc * dim)], dtype=np.float32).reshape(tc, dim)
    g3 = lcg([99])
    machine = Machine(cfg["text_card"] + 1, 2, 8, 2)
    for w in range(6):
        nt = 1 + (next(g3) >> 8) % 3
        toks = [4 + (next(g3) >> 8) % (cfg["text_card"] - 4) for _ in range(nt)]
        machine.push(toks, w % 2)

    gm = msx.Model(path, cfg); gs = msx.Stream(gm)
    if voice:
        gs.load_voice(vp)
    else:
        gs.set_condition(cs, cc)
    ncb, dep_q, delays = cfg["n_q"] + 1, cfg["dep_q"], cfg["delays"]
    max_delay = max(delays); CT = max_delay + 2
    cache = np.full((CT, ncb), -2, dtype=np.int64)
    initial = [cfg["text_card"]] + [cfg["card"]] * cfg["n_q"]
    offset, f, exp = 0, 0, []
    while (offset < machine.end_step + 0 + 4 or machine.end_step == -1) and f < 60:      # moshi_lm_is_active, delay_steps 0
        inp = [initial[i] if offset <= delays[i] else int(cache[offset % CT][i]) for i in range(ncb)]
        t, _, _ = gs.step_temporal(np.array(inp, dtype=np.int32), want_logits=False)
        t = machine.process(offset, int(t))
        a, _ = gs.step_depformer(t, want_logits=False)
        offset += 1
        cache[offset % CT][0] = t
        cache[offset % CT][1:1 + dep_q] = a
        row = [f, 0, -1]
        if offset > max_delay:
            ot = int(cache[(offset - max_delay + delays[0]) % CT][0])
            oa = [int(cache[(offset - max_delay + delays[i]) % CT][i]) for i in range(1, dep_q + 1)]
            if all(v != -1 for v in oa):
                row = [f, 1, ot] + oa
        exp.append(row)
        f += 1
    assert len(got) == len(exp) and len(got) > 20
    assert got == exp
    assert any(r[1] == 1 and r[2] >= cfg["text_card"] + 1 for r in got), "second-stream (look-ahead) tokens must appear"
