"""Mimi split-RVQ quantiser: GPU (msx_rvq_*, device-timed on resident buffers + end to end with host buffers) next to the CPU oracle.
usage: rvq_bench.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkgload; _pkgload.load()
from moshi_cpp_b200 import binding as msx
import oracle
from test_mimi_rvq import make_quantiser
rng = np.random.default_rng(0)
qz = make_quantiser(rng, 1, 31, 2048, 256, 512)
g = msx.RVQ(*qz); o = oracle.SplitRVQ(*qz)
oracle.set_threads(os.cpu_count() or 1)
print(f"Mimi quantiser: 1 + 31 codebooks of 2048 x 256 f32 (2 MB each), latent 512; CPU oracle on {oracle.lib().orc_max_threads()} threads")
for T, n_q in [(1, 8), (1, 16), (1, 32), (375, 8), (375, 32)]:
    x = rng.standard_normal((T, 512)).astype(np.float32) * 3.0
    codes = g.encode(x, n_q)
    enc_ms, dec_ms = g.bench(T, n_q, 20)
    t = time.perf_counter()
    for _ in range(5):
        g.encode(x, n_q)
    e2e_enc = (time.perf_counter() - t) / 5 * 1e3
    t = time.perf_counter()
    for _ in range(5):
        g.decode(codes)
    e2e_dec = (time.perf_counter() - t) / 5 * 1e3
    t = time.perf_counter(); c2 = o.encode(x, n_q); cpu_enc = (time.perf_counter() - t) * 1e3
    t = time.perf_counter(); o.decode(codes); cpu_dec = (time.perf_counter() - t) * 1e3
    cb_bytes = n_q * 2048 * 256 * 4
    print(f"T = {T:3d} frames, {n_q:2d} codebooks: encode {enc_ms:8.3f} ms device ({T / enc_ms * 1e3:9.0f} frames/s, codebooks read at {cb_bytes / enc_ms / 1e6:7.1f} GB/s), "
          f"{e2e_enc:8.3f} ms with host buffers | decode {dec_ms:7.3f} ms device, {e2e_dec:7.3f} ms host | CPU oracle encode {cpu_enc:9.2f} ms, decode {cpu_dec:7.2f} ms | codes identical: {bool(np.array_equal(codes, c2))}")
