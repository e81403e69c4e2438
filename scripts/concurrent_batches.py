"""Several lock-step batches stepped concurrently from host threads (each batch has its own CUDA stream): do the latency-bound
launch chains of independent batches fill each other's gaps?  usage: concurrent_batches.py [preset] [quant]"""
import os, sys, time, threading
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
K = 60
for n_batches, n in [(1, 8), (2, 8), (4, 8), (8, 8), (1, 64), (2, 32), (2, 64)]:
    batches = [msx.Batch(m, n, 512) for _ in range(n_batches)]
    frames = [rng.integers(0, cfg["card"], size=(n, 64, cfg["n_q"] + 1)).astype(np.int32) for _ in range(n_batches)]
    for f in frames:
        f[:, :, 0] = rng.integers(0, cfg["text_card"], size=(n, 64))
    for b, f in zip(batches, frames):
        b.run_resident(f, 5)
    ms = [0.0] * n_batches
    def work(i):
        ms[i], _ = batches[i].run_resident(frames[i], K)
    th = [threading.Thread(target=work, args=(i,)) for i in range(n_batches)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    wall = (time.perf_counter() - t0) * 1e3
    print(f"{n_batches} x {n:2d} streams: wall {wall / K:7.3f} ms per step of all batches (device per batch {max(ms) / K:7.3f}) -> {n_batches * n * K / wall * 1e3:8.0f} frames/s aggregate", flush=True)
    for b in batches: b.close()
