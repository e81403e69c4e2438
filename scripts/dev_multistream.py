"""dev: aggregate frames/s of S independent streams running concurrently on ONE GPU (own CUDA stream + graphs each)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkgload; _pkgload.load()
from moshi_cpp_b200 import configs, synth, binding as msx
cfg = configs.get("moshi7b"); path = synth.cached_gguf("moshi7b", "q4_k")
m = msx.Model(path, cfg)
rng = np.random.default_rng(0)
frames = rng.integers(0, cfg["card"], size=(32, cfg["n_q"] + 1)).astype(np.int32)
for S in (1, 2, 4, 8):
    streams = [msx.Stream(m) for _ in range(S)]
    for s in streams: s.run_resident(frames, 10)
    K = 150
    t0 = time.perf_counter()
    for s in streams: s.run_resident_async(frames, K)
    ms = [s.wait() for s in streams]
    wall = (time.perf_counter() - t0) * 1e3
    print(f"S={S}: per-stream device ms/frame {max(ms)/K:.3f}  aggregate {S*K/(max(ms)*1e-3):.1f} fps (wall {S*K/(wall*1e-3):.1f})")
    for s in streams: s.close()
