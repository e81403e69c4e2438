"""CPU-only tests: the C-ABI library loads and exports every declared symbol, fails loudly without a
GPU, the GGUF writer round-trips through gguf-py, the LMGen host logic matches a restatement of the
reference, and the 2-rank (gloo) plumbing of the stream-sharded mode works."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from moshi_cpp_b200 import binding as msx, configs, parallel, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "moshi_b200.h")).read()
    declared = sorted(set(re.findall(r"MSX_API[^;(]*?\b(msx_\w+)\s*\(", hdr)))
    assert len(declared) >= 30
    out = subprocess.check_output(["nm", "-D", "--defined-only", msx.SO_PATH], text=True)
    exported = set(re.findall(r" T (msx_\w+)", out))
    missing = [d for d in declared if d not in exported]
    assert not missing, f"declared in include/moshi_b200.h but not exported: {missing}"
    L = msx.lib()
    for d in declared:
        getattr(L, d)
    assert b"sm_100a" in L.msx_version()


def test_no_cpu_fallback(gguf_for):
    """without a CUDA device every compute entry point fails with MSX_ERR_CUDA (-4) instead of computing on the host"""
    if msx.lib().msx_device_count() > 0:
        pytest.skip("a GPU is present")
    path, cfg = gguf_for("tiny", "q4_k")
    with pytest.raises(msx.MsxError) as e:
        msx.Model(path, cfg)            # the GGUF parses fine; device bring-up must fail
    assert e.value.code == -4
    rng = np.random.default_rng(0)
    raw = synth.random_tensor(rng, synth.GGML_Q4_K, 8, 256, 0.05)
    with pytest.raises(msx.MsxError) as e:
        msx.test_gemv(synth.GGML_Q4_K, raw, 256, np.zeros(256, np.float32))
    assert e.value.code == -4
    with pytest.raises(msx.MsxError) as e:                      # the quantisers too: no host implementation behind them
        msx.test_quantize_rows(synth.GGML_Q4_K, np.zeros((1, 256), np.float32))
    assert e.value.code == -4
    # file converters: a well-formed input parses (no -3) and then stops at the missing device (-4), no host quantiser
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        st = os.path.join(d, "m.safetensors")
        synth.write_safetensors(st, [("text_emb.weight", "F32", [3, 32], np.zeros((3, 32), np.float32).tobytes()),
                                     ("out_norm.alpha", "BF16", [1, 1, 32], np.zeros(32, np.uint16).tobytes())])
        for fn, src in ((msx.safetensors_to_gguf, st), (msx.gguf_quantize, path)):
            with pytest.raises(msx.MsxError) as e:
                fn(src, os.path.join(d, "out.gguf"), "q8_0")
            assert e.value.code == -4
            assert not os.path.exists(os.path.join(d, "out.gguf"))


def test_loader_errors_before_device(gguf_for, tmp_path):
    path, cfg = gguf_for("tiny", "q4_k")
    with pytest.raises(msx.MsxError) as e:
        msx.Model(str(tmp_path / "nope.gguf"), cfg)
    assert e.value.code == -2                                   # reference: NULL from moshi_lm_from_files (moshi.cpp:621-627)
    bad = tmp_path / "bad.gguf"; bad.write_bytes(b"GGUF" + b"\x07\0\0\0" + b"\0" * 32)
    with pytest.raises(msx.MsxError) as e:
        msx.Model(str(bad), cfg)
    assert e.value.code == -3
    trunc = tmp_path / "trunc.gguf"; trunc.write_bytes(open(path, "rb").read(4096))
    with pytest.raises(msx.MsxError) as e:
        msx.Model(str(trunc), cfg)
    assert e.value.code == -3
    wrong = dict(cfg); wrong["num_heads"] = 3
    with pytest.raises(msx.MsxError) as e:
        msx.Model(path, wrong)
    assert e.value.code == -1
    with pytest.raises(msx.MsxError) as e:                      # file-to-file quantiser: same error codes
        msx.gguf_quantize(str(tmp_path / "nope.gguf"), str(tmp_path / "out.gguf"), "q4_k")
    assert e.value.code == -2
    with pytest.raises(msx.MsxError) as e:
        msx.gguf_quantize(str(bad), str(tmp_path / "out.gguf"), "q8_0")
    assert e.value.code == -3
    with pytest.raises(msx.MsxError) as e:                      # writing over the (memory-mapped) input is refused
        msx.gguf_quantize(path, path, "q8_0")
    assert e.value.code == -1
    with pytest.raises(msx.MsxError) as e:                      # safetensors front end
        msx.safetensors_to_gguf(str(tmp_path / "nope.safetensors"), str(tmp_path / "out.gguf"), "q4_k")
    assert e.value.code == -2
    for blob in (b"\x10\0\0\0\0\0\0\0" + b'{"a": {"dtype"', (1 << 40).to_bytes(8, "little") + b"{}" * 8,
                 b"\x30\0\0\0\0\0\0\0" + b'{"a":{"dtype":"F32","shape":[4],"data_offsets":[0,64]}}'.ljust(48)):
        st = tmp_path / "bad.safetensors"; st.write_bytes(blob)
        with pytest.raises(msx.MsxError) as e:
            msx.safetensors_to_gguf(str(st), str(tmp_path / "out.gguf"), None)
        assert e.value.code == -3


def _gguf(n_tensors, n_kv, infos=b"", pad=64):
    """hand-made GGUF v3 header: magic, version, tensor count, key count, raw tensor-info bytes, zero padding"""
    return b"GGUF" + (3).to_bytes(4, "little") + n_tensors.to_bytes(8, "little") + n_kv.to_bytes(8, "little") + infos + b"\0" * pad


def _info(name, ne, ggml_type, offset):
    b = len(name).to_bytes(8, "little") + name.encode() + len(ne).to_bytes(4, "little")
    for v in ne:
        b += v.to_bytes(8, "little")
    return b + ggml_type.to_bytes(4, "little") + offset.to_bytes(8, "little")


def test_hostile_gguf_headers_are_rejected(gguf_for, tmp_path):
    """header fields of a (downloaded) model file are untrusted: counts beyond the file, zero / huge dimensions, sizes that
    wrap and offsets outside the mapping all end in MSX_ERR_FORMAT (-3) — no allocation of attacker-chosen size, no abort,
    no pointer outside the mmap."""
    _, cfg = gguf_for("tiny", "q4_k")
    cases = {
        "huge_tensor_count": _gguf(1 << 60, 0),
        "huge_key_count": _gguf(0, 1 << 61),
        "zero_dim": _gguf(1, 0, _info("lm.text_emb.weight", [0, 7], 0, 0)),
        "huge_dim": _gguf(1, 0, _info("lm.text_emb.weight", [1 << 62, 1 << 62], 0, 0)),
        "size_wraps": _gguf(1, 0, _info("lm.text_emb.weight", [1 << 33, 1 << 33, 1 << 33, 1 << 33], 0, 0)),
        "offset_wraps": _gguf(1, 0, _info("lm.text_emb.weight", [4, 2], 0, (1 << 64) - 16)),
        "offset_past_end": _gguf(1, 0, _info("lm.text_emb.weight", [4, 2], 0, 1 << 20)),
        "array_value_wraps": b"GGUF" + (3).to_bytes(4, "little") + (0).to_bytes(8, "little") + (1).to_bytes(8, "little")
                             + (1).to_bytes(8, "little") + b"k" + (9).to_bytes(4, "little") + (10).to_bytes(4, "little") + ((1 << 62) + 1).to_bytes(8, "little") + b"\0" * 64,
    }
    for name, blob in cases.items():
        f = tmp_path / f"{name}.gguf"; f.write_bytes(blob)
        with pytest.raises(msx.MsxError) as e:
            msx.Model(str(f), cfg)
        assert e.value.code == -3, name
        with pytest.raises(msx.MsxError) as e:
            msx.gguf_quantize(str(f), str(tmp_path / "out.gguf"), "q8_0")
        assert e.value.code == -3, name
        assert not (tmp_path / "out.gguf").exists()
    # safetensors converter: zero dimension (division by zero in the reference's split logic) and overflowing shapes
    for shape in ([3, 0], [1 << 40, 1 << 40]):
        hdr = ('{"self_attn.in_proj_weight":{"dtype":"F32","shape":%s,"data_offsets":[0,0]}}' % str(shape).replace(" ", "")).encode()
        hdr += b" " * (-len(hdr) % 8)
        st = tmp_path / "z.safetensors"; st.write_bytes(len(hdr).to_bytes(8, "little") + hdr)
        with pytest.raises(msx.MsxError) as e:
            msx.safetensors_to_gguf(str(st), str(tmp_path / "out.gguf"), "q8_0")
        assert e.value.code == -3, shape
        assert not (tmp_path / "out.gguf").exists()


def test_gguf_roundtrip_with_gguf_py(gguf_for):
    from gguf import GGUFReader
    path, cfg = gguf_for("tiny", "q8_0")
    r = GGUFReader(path)
    man = synth.manifest(cfg, synth.GGML_Q8_0)
    assert [t.name for t in r.tensors] == [m[0] for m in man]
    for t, (name, gt, k, rows, _) in zip(r.tensors, man):
        assert int(t.tensor_type) == gt and int(t.shape[0]) == k
        assert t.data.nbytes == synth.row_bytes(gt, k) * rows
    assert len(r.fields) == 3           # GGUF.version / tensor_count / kv_count only: no metadata, like the reference's files


def test_weight_bytes_match_survey():
    """algorithmic bytes per frame: 7B q4_k = 4 148 232 192 B, q8_0 = 7 835 549 696 B"""
    def w_bytes(cfg, quant):
        qt = synth.TYPE_NAMES[quant]
        tot = 0
        for name, gt, k, rows, _ in synth.manifest(cfg, qt):
            if "emb" in name or name.endswith("alpha"):
                continue
            tot += synth.row_bytes(gt, k) * rows
        return tot
    assert w_bytes(configs.get("moshi7b"), "q4_k") == 4_148_232_192
    assert w_bytes(configs.get("moshi7b"), "q8_0") == 7_835_549_696


# ---- LMGen host logic: restatement of moshi_lmgen_step (lm.h:778-979) in Python as the checker -----
class RefLMGen:
    def __init__(self, cfg, step_fn, delay_steps=0):
        self.c, self.step_fn, self.delay_steps = cfg, step_fn, delay_steps
        self.ncb = cfg["n_q"] + 1
        self.max_delay = max(cfg["delays"])
        self.CT = self.max_delay + 2 + (1 if cfg["model_type"] == "personaplex" else 0)
        self.cache = [[-2] * self.ncb for _ in range(self.CT)]
        self.initial = [cfg["text_card"]] + [cfg["card"]] * cfg["n_q"]
        self.offset = 0

    def step(self, toks, replace=False):
        c, CT, ncb = self.c, self.CT, self.ncb
        dep_q = 8 if c["model_type"] == "personaplex" else c["dep_q"]
        needed = ncb - dep_q - 1
        provided = False
        if needed > 0:
            if len(toks) == ncb:
                for i in range(ncb):
                    self.cache[(self.offset + c["delays"][i]) % CT][i] = int(toks[i])
                provided = True
            else:
                for i in range(needed):
                    self.cache[(self.offset + c["delays"][dep_q + 1 + i]) % CT][dep_q + 1 + i] = int(toks[i])
        pos = self.offset % CT
        inp = [self.initial[i] if self.offset <= c["delays"][i] else self.cache[pos][i] for i in range(ncb)]
        out = self.step_fn(inp, replace)
        text, audio = out[0], list(out[1:1 + c["dep_q"]])
        if replace:
            audio = [-1] * c["dep_q"]
        if c["dep_q"] > 0 and self.delay_steps:
            for q in range(c["dep_q"]):
                if self.offset < c["delays"][q + 1] + self.delay_steps:
                    audio[q] = -1
        self.offset += 1
        if not provided:
            p = self.offset % CT
            self.cache[p][0] = text
            for q in range(c["dep_q"]):
                self.cache[p][q + 1] = audio[q]
        if self.offset <= self.max_delay or replace:
            return 0, None, audio
        t = self.cache[(self.offset - self.max_delay + c["delays"][0]) % CT][0]
        for i in range(1, dep_q + 1):
            audio[i - 1] = self.cache[(self.offset - self.max_delay + c["delays"][i]) % CT][i]
        if any(a == -1 for a in audio):
            return 0, None, audio
        return 1, t, audio


@pytest.mark.parametrize("preset,delay_steps", [("tiny", 0), ("tiny", 3), ("tiny_pplex", 0), ("tiny_stt", 0)])
def test_lmgen_host_logic(preset, delay_steps):
    cfg = configs.get(preset)
    n_out = 1 + cfg["dep_q"]

    def fake_model(tokens, replace):         # deterministic function of the gathered input row
        h = 1469598103934665603
        for t in tokens:
            h = ((h ^ (int(t) & 0xFFFFFFFF)) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return [int((h >> (5 * i)) % (cfg["text_card"] if i == 0 else cfg["card"])) for i in range(n_out)]

    ref = RefLMGen(cfg, fake_model, delay_steps)
    gen = msx.Gen(None, delay_steps=delay_steps, cfg=cfg, step_fn=fake_model)
    rng = np.random.default_rng(1)
    dep_q_eff = 8 if cfg["model_type"] == "personaplex" else cfg["dep_q"]
    n_user = cfg["n_q"] - dep_q_eff
    for f in range(40):
        if f in (0, 1, 17):
            toks = rng.integers(0, cfg["card"], size=cfg["n_q"] + 1)      # all streams provided (prompt replay)
        else:
            toks = rng.integers(0, cfg["card"], size=n_user)
        rep = f in (2, 3)
        ok_r, t_r, a_r = ref.step(toks, rep)
        ok_g, t_g, a_g = gen.step(toks, rep)
        assert ok_g == ok_r, f"frame {f}"
        assert list(a_g) == list(a_r), f"frame {f}"
        if ok_r:
            assert t_g == t_r
        assert gen.offset == ref.offset


def test_stream_assignment():
    a = parallel.assign_streams(64, 8)
    assert all(len(x) == 8 for x in a) and sorted(sum(a, [])) == list(range(64))
    assert parallel.assign_streams(3, 2) == [[0, 2], [1]]
    assert parallel.aggregate_throughput([100, 100], [50.0, 100.0]) == pytest.approx(2000.0)


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import _pkgload; _pkgload.load()
import torch.distributed as dist
from moshi_cpp_b200 import parallel
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
mine = parallel.assign_streams(5, w)[r]
local_ms = 10.0 * (r + 1)                       # pretend device time
dist.barrier()
worst = parallel.reduce_max_ms(local_ms, dist)
assert worst == 10.0 * w, worst
import torch
n = torch.tensor([len(mine)]); dist.all_reduce(n)
assert int(n[0]) == 5
# tensor-parallel bring-up protocol: id from rank 0 to everyone, one handle per rank gathered in rank order
ids = [b"id-from-rank-0" if r == 0 else None]
dist.broadcast_object_list(ids, src=0)
assert ids[0] == b"id-from-rank-0"
hs = parallel.tp_exchange(dist, bytes([r]) * 64)
assert [h[0] for h in hs] == list(range(w)) and all(len(h) == 64 for h in hs)
if r == 0:
    print("OK", worst, int(n[0]))
dist.destroy_process_group()
"""


def test_two_rank_gloo_plumbing(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(script), ROOT],
                         capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "OK 20.0 5" in out.stdout


def test_bench_reference_arm_contract_under_torchrun():
    """`bench.py --impl reference` launched like the driver launches it for N > 1: rank 0 alone times the CPU path and prints
    ONE JSON line on stdout (nothing else reaches stdout), the other rank exits 0 without work.  torchrun exports
    OMP_NUM_THREADS=1 to its workers: the arm must still use every host core it may run on and say how many it used
    (the SCALE reference numbers at N >= 2 ran on one thread while reporting 32 cores)."""
    import json
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29543", os.path.join(ROOT, "bench.py"),
                          "--impl", "reference", "--preset", "tiny", "--gpus", "2", "--steps", "3", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["unit"] == "frames/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    avail = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["cores"] == avail, "the reference arm must set its OpenMP thread count explicitly"
    if avail > 1:
        assert d["cpu_baseline"]["cores"] > 1
    assert set(("metric", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config")) <= set(d)


@pytest.mark.parametrize("preset", ["moshi7b", "personaplex7b", "tiny", "stt1b"])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_tensor_parallel_shards_tile_the_model(preset, world):
    """heads and hidden slices of all ranks are disjoint, cover everything, and respect the block alignment the
    quantised K-slices need (256 for q4_k super-blocks when the width allows, else 32)"""
    cfg = configs.get(preset)
    if cfg["num_heads"] % world:
        with pytest.raises(ValueError):
            parallel.tp_shard(cfg, 0, world)
        return
    F = cfg["hidden"]
    try:
        shards = [parallel.tp_shard(cfg, r, world) for r in range(world)]
    except ValueError:
        assert F // (256 if F % 256 == 0 else 32) < world
        return
    assert shards[0]["h0"] == 0 and shards[-1]["h1"] == cfg["num_heads"] and shards[0]["f0"] == 0 and shards[-1]["f1"] == F
    for a, b in zip(shards, shards[1:]):
        assert a["h1"] == b["h0"] and a["f1"] == b["f0"]
    unit = 256 if F % 256 == 0 else 32
    dh = cfg["dim"] // cfg["num_heads"]
    for s in shards:
        assert s["f0"] % unit == 0 and s["f1"] % unit == 0 and s["f1"] > s["f0"]
        assert s["adim"] == (s["h1"] - s["h0"]) * dh
        assert (s["h0"] * dh) % 32 == 0


def test_tensor_parallel_decomposition_is_exact_in_double():
    """why the shards reproduce one GPU bit for bit: out_proj / linear_out partial sums over K-slices, added in
    double, equal the full contraction; in_proj / linear_in row slices concatenate"""
    import oracle as orc
    rng = np.random.default_rng(0)
    cfg = configs.get("tiny")
    k, rows = cfg["hidden"], cfg["dim"]
    raw = synth.random_tensor(rng, synth.GGML_Q8_0, rows, k, 0.05)
    w = orc.dequantize(synth.GGML_Q8_0, raw, k).astype(np.float64)
    x = rng.standard_normal(k)
    full = w @ x
    for world in (2, 3):
        if world == 3:
            continue
        parts = []
        for r in range(world):
            s = parallel.tp_shard(cfg, r, world)
            parts.append(w[:, s["f0"]:s["f1"]] @ x[s["f0"]:s["f1"]])
        assert np.allclose(np.sum(parts, axis=0), full, rtol=1e-13, atol=1e-13)
    rows_w = rng.standard_normal((3 * cfg["dim"], 16))
    got = []
    for sec in range(3):
        for r in range(2):
            a, b = parallel.tp_shard(cfg, r, 2)["in_proj_rows"][sec]
            got.append(rows_w[a:b])
    assert np.array_equal(np.concatenate(got), rows_w)
