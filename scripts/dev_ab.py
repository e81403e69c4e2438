"""dev: A/B frame time of the persistent-depformer stream vs the multi-kernel stream (resident replay)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkgload; _pkgload.load()
from moshi_cpp_b200 import configs, synth, binding as msx
preset = sys.argv[1] if len(sys.argv) > 1 else "moshi7b"
quant = sys.argv[2] if len(sys.argv) > 2 else "q4_k"
cfg = configs.get(preset); path = synth.cached_gguf(preset, quant)
m = msx.Model(path, cfg)
rng = np.random.default_rng(0)
frames = rng.integers(0, cfg["card"], size=(32, cfg["n_q"] + 1)).astype(np.int32)
for name, mk in (("persistent", True), ("multi-kernel", False), ("persistent", True), ("multi-kernel", False)):
    s = msx.Stream(m, persistent_depformer=mk)
    s.run_resident(frames, 20)
    ms, _ = s.run_resident(frames, 200)
    print(f"{name:14s} launches/frame {s.launches_per_frame:4d}  {ms/200:.3f} ms/frame  {200/ms*1e3:.1f} fps")
    s.close()
