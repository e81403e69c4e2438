"""Profiling driver: load the model, warm up, then run a few frames between cudaProfilerStart/Stop.
Use under ncu with --profile-from-start off."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkgload; _pkgload.load()
from moshi_cpp_b200 import configs, synth, binding as msx

preset = sys.argv[1] if len(sys.argv) > 1 else "moshi7b"
quant = sys.argv[2] if len(sys.argv) > 2 else "q4_k"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
prefill = int(sys.argv[4]) if len(sys.argv) > 4 else 8
cfg = configs.get(preset)
path = synth.cached_gguf(preset, quant)
m = msx.Model(path, cfg); s = msx.Stream(m)
rng = np.random.default_rng(42)
frames = rng.integers(0, cfg["card"], size=(64, cfg["n_q"] + 1)).astype(np.int32)
s.run_resident(frames, prefill)
rt = None
for name in ("libcudart.so", "libcudart.so.12"):
    try:
        rt = ctypes.CDLL(name); break
    except OSError:
        pass
if rt: rt.cudaProfilerStart()
for i in range(n):
    s.step(frames[i])
if rt: rt.cudaProfilerStop()
print("done", s.offset)
