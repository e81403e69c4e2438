"""The C++ mirror of the reference's moshi_lm_* API (moshi.cpp_b200/host) exercised through the
moshi-sts --bench style tool, which only uses that API (like tools/moshi-sts.cpp uses include/moshi/moshi.h)."""
import json
import os
import subprocess

import numpy as np
import pytest

from moshi_cpp_b200 import binding as msx, configs

PROMPT_TOKENS = [3, 948, 243, 1178, 546, 1736, 1030, 1978, 2008, 430, 1268, 381, 1611, 1095, 1495, 56, 472]   # lm.h:983-987


@pytest.fixture(scope="module")
def tool():
    msx.build_host()
    assert os.path.exists(msx.STS_BENCH)
    return msx.STS_BENCH


def write_config(path, cfg):
    with open(path, "w") as f:
        json.dump(configs.to_config_json(cfg), f)


def lcg_user_codes(n_frames, n_user, card):
    lcg, out = 42, []
    for _ in range(n_frames):
        row = []
        for _ in range(n_user):
            lcg = (lcg * 1664525 + 1013904223) & 0xFFFFFFFF
            row.append((lcg >> 8) % card)
        out.append(row)
    return out


def test_tool_error_paths(tool, gguf_for, tmp_path):
    path, cfg = gguf_for("tiny", "q4_k")
    cj = tmp_path / "config.json"; write_config(cj, cfg)
    r = subprocess.run([tool, str(tmp_path / "missing.gguf"), str(cj)], capture_output=True, text=True)
    assert r.returncode == 1 and "could not open" in r.stderr          # reference: moshi_lm_from_files -> NULL
    r = subprocess.run([tool, path, str(tmp_path / "missing.json")], capture_output=True, text=True)
    assert r.returncode == 1 and "failed to open" in r.stderr           # reference message (config.h:160-163)
    if msx.lib().msx_device_count() == 0:
        r = subprocess.run([tool, path, str(cj), "4"], capture_output=True, text=True)
        assert r.returncode == 1 and "error:" in r.stderr               # no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("container", ["safetensors", "gguf"])
def test_personaplex_voice_file(tool, gguf_for, tmp_path, container):
    """moshi_lm_personaplex_load_voice (moshi.cpp:789-836): prompt embeddings [N,1,1,dim] + the token ring stored
    [codebook][time] (lm.h:1005-1051) from a .safetensors / .gguf voice file == the same tensors fed through the C ABI."""
    from moshi_cpp_b200 import synth
    path, cfg = gguf_for("tiny_pplex", "q4_k")
    cj = tmp_path / "config.json"; write_config(cj, cfg)
    rng = np.random.default_rng(12)
    ncb, dim = cfg["n_q"] + 1, cfg["dim"]
    CT = max(cfg["delays"]) + 3                                   # max_delay + 2, + 1 for PersonaPlex (lm.h:715-743)
    emb = (0.3 * rng.standard_normal((5, dim))).astype(np.float32)
    ring = rng.integers(0, cfg["card"], size=(CT, ncb)).astype(np.int32)
    ring[:, 0] = rng.integers(0, cfg["text_card"], size=CT)
    if container == "safetensors":
        vp = str(tmp_path / "voice.safetensors")
        bits = ((emb.view(np.uint32) + 0x7FFF + ((emb.view(np.uint32) >> 16) & 1)) >> 16).astype(np.uint16)
        emb = (bits.astype(np.uint32) << 16).view(np.float32)     # the file carries bf16 embeddings
        synth.write_safetensors(vp, [("embeddings", "BF16", [5, 1, 1, dim], bits.tobytes()),
                                     ("cache", "I32", [1, ncb, CT], np.ascontiguousarray(ring.T).tobytes())])
    else:
        vp = str(tmp_path / "voice.gguf")
        import gguf
        w = gguf.GGUFWriter(vp, "voice")
        w.add_tensor("voice.embeddings", emb.reshape(5, 1, 1, dim))
        w.add_tensor("voice.cache", np.ascontiguousarray(ring.T).reshape(1, ncb, CT))
        w.write_header_to_file(); w.write_kv_data_to_file(); w.write_tensors_to_file(); w.close()
    frames = 30
    r = subprocess.run([tool, path, str(cj), str(frames), "0", "--print-tokens", "--pplex-voice", vp], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l.split() for l in r.stdout.strip().splitlines()]
    assert len(lines) == frames
    gs = msx.Stream(msx.Model(path, cfg)); gg = msx.Gen(gs)
    for row_ in emb:
        gg.prompt_embedding(row_)
    gg.set_cache(ring)
    def row(text=None):
        t = [PROMPT_TOKENS[0]] + [v % cfg["card"] for v in PROMPT_TOKENS[1:]]     # folded into the test model's tables (moshi_api.cpp)
        if text is not None: t[0] = text
        return t
    for _ in range(6): gg.step(row())
    for tok in (5, 17, 99, 250): gg.step(row(text=tok))
    for _ in range(6): gg.step(row())
    users = lcg_user_codes(frames, cfg["n_q"] - 8, cfg["card"])
    n_ok = 0
    for f in range(frames):
        ok, text, audio = gg.step(users[f])
        assert int(lines[f][1]) == ok, f"frame {f}"
        if ok:
            n_ok += 1
            assert int(lines[f][2]) == text and [int(v) for v in lines[f][3:]] == list(audio), f"frame {f}"
    assert n_ok > 20
    bad = subprocess.run([tool, path, str(cj), "4", "0", "--pplex-voice", str(tmp_path / "voice.pt")], capture_output=True, text=True)
    assert bad.returncode == 1 and "could not load voice" in bad.stderr     # unknown extension -> -1 (moshi.cpp:808-810)


@pytest.mark.gpu
@pytest.mark.parametrize("quant", ["q4_k", "q8_0"])
def test_tool_quantize_and_save_gguf(tool, tmp_path, quant):
    """`moshi-sts -q <quant> -g out.gguf` then running the saved file == `-q <quant>` on the unquantised file
    (moshi_lm_quantize / moshi_lm_save_gguf / moshi_lm_load, moshi.cpp:654-695)."""
    from moshi_cpp_b200 import synth
    cfg = configs.get("tiny")
    src = str(tmp_path / "tiny-bf16.gguf"); out = str(tmp_path / f"tiny-{quant}.gguf")
    synth.write_gguf(src, cfg, "bf16", seed=5)
    cj = tmp_path / "config.json"; write_config(cj, cfg)
    r = subprocess.run([tool, src, str(cj), "-q", quant, "-g", out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and os.path.getsize(out) < os.path.getsize(src) * 0.6, r.stderr
    a = subprocess.run([tool, src, str(cj), "30", "0", "--print-tokens", "-q", quant], capture_output=True, text=True, timeout=300)
    b = subprocess.run([tool, out, str(cj), "30", "0", "--print-tokens"], capture_output=True, text=True, timeout=300)
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
    ta = [l for l in a.stdout.splitlines() if l[:1].isdigit()]
    tb = [l for l in b.stdout.splitlines() if l[:1].isdigit()]
    assert len(ta) == 30 and ta == tb
    if quant == "q4_k":      # the same checkpoint as model.safetensors (torch names; here no fused tensors): -q on the fly == the saved file
        import gguf
        dt = {0: "F32", 1: "F16", 30: "BF16"}
        st = [(t.name[3:], dt[int(t.tensor_type)], [int(v) for v in t.shape][::-1], np.ascontiguousarray(t.data).tobytes())
              for t in gguf.GGUFReader(src).tensors]
        sp = str(tmp_path / "model.safetensors")
        synth.write_safetensors(sp, st)
        c = subprocess.run([tool, sp, str(cj), "30", "0", "--print-tokens", "-q", quant], capture_output=True, text=True, timeout=300)
        assert c.returncode == 0, c.stderr
        assert [l for l in c.stdout.splitlines() if l[:1].isdigit()] == ta
    r = subprocess.run([tool, src, str(cj), "-q", "q3_x"], capture_output=True, text=True)
    assert r.returncode == 1 and "unknown quantisation" in r.stderr     # moshi_lm_quantize -> false (moshi.cpp:667-668)


@pytest.mark.gpu
@pytest.mark.parametrize("preset", ["tiny", "tiny_pplex", "tiny_stt"])
def test_cpp_api_matches_c_abi(tool, gguf_for, tmp_path, preset):
    quant = "q8_0" if preset == "tiny_stt" else "q4_k"
    path, cfg = gguf_for(preset, quant)
    cj = tmp_path / "config.json"; write_config(cj, cfg)
    frames = 40
    r = subprocess.run([tool, path, str(cj), str(frames), "0", "--print-tokens"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l.split() for l in r.stdout.strip().splitlines()]
    assert len(lines) == frames
    # the same thing through the C ABI from python
    gm = msx.Model(path, cfg); gs = msx.Stream(gm); gg = msx.Gen(gs)
    pplex = cfg["model_type"] == "personaplex"
    if pplex:   # moshi_lmgen_step_system_prompts (lm.h:1120-1134): voice codes, 6 silence, text prompt, 6 silence
        def row(text=None, codes=None):
            t = [PROMPT_TOKENS[0]] + [v % cfg["card"] for v in PROMPT_TOKENS[1:]]     # folded into the test model's tables (moshi_api.cpp)
            if text is not None: t[0] = text
            if codes is not None: t[1:9] = codes
            return t
        for f in range(4):
            gg.step(row(codes=[(f * 131 + j * 17) % cfg["card"] for j in range(8)]))
        for _ in range(6): gg.step(row())
        for tok in (5, 17, 99, 250): gg.step(row(text=tok))
        for _ in range(6): gg.step(row())
    n_user = cfg["n_q"] - (8 if pplex else cfg["dep_q"])
    users = lcg_user_codes(frames, n_user, cfg["card"])
    for f in range(frames):
        ok, text, audio = gg.step(users[f])
        if cfg["dep_q"] > 0:
            assert int(lines[f][1]) == ok, f"frame {f}"
            if ok:
                assert int(lines[f][2]) == text and [int(v) for v in lines[f][3:]] == list(audio), f"frame {f}"
        else:
            if ok:                        # the VAD head is evaluated on emitting frames only (lm.h:950-977)
                assert int(lines[f][1]) == text
                assert abs(float(lines[f][2]) - gs.vad()) < 1e-5
