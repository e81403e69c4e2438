"""Regenerates tests/golden/golden.{npz,json}.

Ground truth for the block formats is gguf-py 0.19.0 (gguf/quants.py — ggml's own python
implementation); the GEMV and LMGen traces are the oracle's outputs, pinned so that any change to
the oracle's numerics shows up as a golden-vector diff.  (The reference ships no golden vectors and
cannot be built or imported here: SURVEY.md §4, §8c.)
    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import _pkgload  # noqa: E402
_pkgload.load()
import oracle  # noqa: E402
from gguf import GGMLQuantizationType as QT  # noqa: E402
from gguf.quants import dequantize  # noqa: E402
from moshi_cpp_b200 import configs, synth  # noqa: E402

rng = np.random.default_rng(2026)
arrays, meta = {}, {"dequant": {}, "gemv": {}}
for quant, k, rows in (("q4_k", 512, 6), ("q8_0", 96, 5), ("q4_0", 64, 7)):
    gt = synth.TYPE_NAMES[quant]
    raw = synth.random_tensor(rng, gt, rows, k, 0.05)
    if quant == "q4_k":
        blk = raw.reshape(-1, 144); blk[:, 4:16] = rng.integers(0, 256, size=(blk.shape[0], 12), dtype=np.uint8)
        raw = blk.reshape(rows, -1)
    name = f"dq_{quant}"
    arrays[name + "_raw"] = raw
    arrays[name + "_f32"] = dequantize(raw, QT(gt)).reshape(rows, k).astype(np.float32)
    meta["dequant"][name] = {"type": gt, "k": k}
for quant, k, rows in (("q4_k", 1024, 16), ("q8_0", 256, 12)):
    gt = synth.TYPE_NAMES[quant]
    raw = synth.random_tensor(rng, gt, rows, k, 1 / np.sqrt(k))
    x = rng.standard_normal(k).astype(np.float32)
    name = f"gemv_{quant}"
    arrays[name + "_raw"] = raw; arrays[name + "_x"] = x
    arrays[name + "_y"] = oracle.mul_mat_vec(gt, raw, k, x)
    meta["gemv"][name] = {"type": gt, "k": k}
cfg = configs.get("tiny")
path = synth.cached_gguf("tiny", "q4_k")
om = oracle.Model(path, cfg); og = oracle.LMGen(om)
r2 = np.random.default_rng(42)
trace = []
for f in range(24):
    user = r2.integers(0, cfg["card"], size=cfg["n_q"] - cfg["dep_q"]).astype(np.int32)
    ok, t, a = og.step(user)
    trace.append([int(ok), int(t)] + [int(v) for v in a])
meta["lmgen_tiny_q4k"] = {"frames": 24, "trace": trace, "model_seed": 1234, "token_seed": 42}
np.savez_compressed(os.path.join(HERE, "golden.npz"), **arrays)
with open(os.path.join(HERE, "golden.json"), "w") as f:
    json.dump(meta, f)
print("wrote", {k: v.shape for k, v in arrays.items()})

# ---- quantisers (second file, so the vectors above never move): gguf-py's Q4_0 / Q8_0 blocks are reference outputs
# (ggml's own python implementation); the Q4_K blocks and the voice-conditioner tensors are the oracle's, pinned as a
# regression (gguf-py has no Q4_K quantiser and the reference ships no voice fixtures)
from gguf.quants import quantize  # noqa: E402
rq = np.random.default_rng(2027)
xq = (rq.standard_normal((6, 512)) * 0.05).astype(np.float32)
xq[0, :32] = 0.0; xq[1] = np.abs(xq[1]); xq[2, 256:] *= 40.0; xq[3, ::2] = xq[3, 1::2]
qa = {"x": xq,
      "q4_0": np.ascontiguousarray(quantize(xq, QT.Q4_0)).view(np.uint8).reshape(6, -1),
      "q8_0": np.ascontiguousarray(quantize(xq, QT.Q8_0)).view(np.uint8).reshape(6, -1),
      "q4_k_oracle": np.stack([oracle.quantize_q4_K(r) for r in xq])}
vcfg = configs.get("tiny_tts_voice")
vpath = synth.cached_gguf("tiny_tts_voice", "q4_k")
wav = rq.standard_normal((synth.COND_CHANNELS, 3)).astype(np.float32)
vs, vc = oracle.voice_condition(vpath, vcfg, wav)
qa["voice_wavs"] = wav; qa["voice_sum_oracle"] = vs; qa["voice_cross_oracle"] = vc
np.savez_compressed(os.path.join(HERE, "golden_quant.npz"), **qa)
print("wrote", {k: v.shape for k, v in qa.items()})
