"""Per-family kernel time of one eagerly launched batched frame (CUDA event after every launch).
usage: batch_profile.py [preset] [quant] [n_streams ...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkgload; _pkgload.load()
from moshi_cpp_b200 import configs, synth, binding as msx
preset = sys.argv[1] if len(sys.argv) > 1 else "moshi7b"
quant = sys.argv[2] if len(sys.argv) > 2 else "q4_k"
ns = [int(a) for a in sys.argv[3:]] or [8, 16, 64]
cfg = configs.get(preset); path = synth.cached_gguf(preset, quant)
m = msx.Model(path, cfg)
rng = np.random.default_rng(0)
for n in ns:
    b = msx.Batch(m, n, 256)
    toks = rng.integers(0, cfg["card"], size=(n, cfg["n_q"] + 1)).astype(np.int32)
    toks[:, 0] = rng.integers(0, cfg["text_card"], size=n)
    for rep in range(3):
        fam = b.profile_frame(toks)
    tot = sum(v[0] for v in fam.values())
    print(f"[{preset} {quant}] {n} streams, one eager frame: {tot:.3f} ms, {sum(v[1] for v in fam.values())} launches")
    for k, (ms, c) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
        print(f"  {k:14s} {ms * 1e3:9.1f} us  {c:4d} launches  {ms * 1e3 / c:7.2f} us each  ({ms / tot:.3f})")
    b.close()
