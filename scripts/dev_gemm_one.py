"""dev: one batched GEMM shape under ncu"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkgload; _pkgload.load()
from moshi_cpp_b200 import binding as msx, synth
k, rows, epi = int(os.environ.get("K", 4096)), int(os.environ.get("ROWS", 22528)), int(os.environ.get("EPI", 2))
rng = np.random.default_rng(1)
raw = synth.random_tensor(rng, synth.GGML_Q4_K, rows, k, 1.0 / np.sqrt(k))
print(msx.bench_gemm_batch(raw, k, 8, 3, 6, epi, True))
