"""dev: batched-stream step timing (moshi7b q4_k): aggregate frames/s, per-family times, GEMM micro-benchmark"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkgload; _pkgload.load()
from moshi_cpp_b200 import binding as msx, configs, synth

preset = os.environ.get("PRESET", "moshi7b")
cfg = configs.get(preset)
path = synth.cached_gguf(preset, "q4_k")
gm = msx.Model(path, cfg)
rng = np.random.default_rng(42)
if os.environ.get("MICRO", "1") == "1":
    for (k, rows, epi, name) in [(4096, 12288, 0, "in_proj"), (4096, 4096, 1, "out_proj"), (4096, 22528, 2, "linear_in"), (11264, 4096, 1, "linear_out")]:
        raw = synth.random_tensor(rng, synth.GGML_Q4_K, rows, k, 1.0 / np.sqrt(k))
        nbytes = raw.size
        for nb in (8,):
            us = msx.bench_gemm_batch(raw, k, nb, 6, 48, epi, True)
            us2 = msx.bench_gemm_batch(raw, k, nb, 6, 48, epi, False)
            print(f"{name:10s} K={k:5d} rows={rows:5d} nb={nb}: quant+gemm {us:6.2f} us  gemm only {us2:6.2f} us  -> {nbytes/us2/1e6:6.2f} TB/s", flush=True)
for n in [int(v) for v in os.environ.get("NS", "8,4,1").split(",")]:
    b = msx.Batch(gm, n)
    frames = rng.integers(0, cfg["card"], size=(n, 64, cfg["n_q"] + 1)).astype(np.int32)
    b.run_resident(frames, 20)
    ms, _ = b.run_resident(frames, 100)
    per = ms / 100
    print(f"batch n={n}: {per:.3f} ms/frame -> {n*1000/per:.1f} frames/s aggregate ({1000/per:.1f} per stream), launches/frame {b.launches_per_frame}", flush=True)
    if n == 8:
        fam = b.profile_frame(frames[:, 0])
        tot = sum(v[0] for v in fam.values())
        print("  eager per-family:", {k: (round(v[0] * 1000 / max(1, v[1]), 1), v[1]) for k, v in fam.items()}, f"total {tot:.3f} ms")
    b.close()
