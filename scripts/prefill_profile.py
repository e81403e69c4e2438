"""Per-family kernel time of one 64-position prefill pass (eager launches, CUDA event after each).
usage: prefill_profile.py [preset] [quant]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkgload; _pkgload.load()
from moshi_cpp_b200 import configs, synth, binding as msx
preset = sys.argv[1] if len(sys.argv) > 1 else "moshi7b"
quant = sys.argv[2] if len(sys.argv) > 2 else "q4_k"
cfg = configs.get(preset); path = synth.cached_gguf(preset, quant)
m = msx.Model(path, cfg); s = msx.Stream(m)
rng = np.random.default_rng(0)
rows = rng.integers(0, cfg["card"], size=(1 + 64, cfg["n_q"] + 1)).astype(np.int32)
rows[:, 0] = rng.integers(0, cfg["text_card"], size=rows.shape[0])
s.prefill(rows[:64])                       # warm
for rep in range(2):
    fam = s.prefill_profile(rows)
tot = sum(v[0] for v in fam.values())
L = cfg["num_layers"]
print(f"[{preset} {quant}] one 64-position prefill pass, eager: {tot:.3f} ms")
for k, (ms, n) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:12s} {ms * 1e3:9.1f} us  {n:4d} launches  {ms * 1e3 / n:7.2f} us each  ({ms / tot:.3f})")
