"""dev: in-kernel timeline of the batched GEMM (globaltimer stamps per CTA)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkgload; _pkgload.load()
from moshi_cpp_b200 import binding as msx, synth
rng = np.random.default_rng(1)
for (k, rows, epi, name) in [(4096, 22528, 2, "linear_in"), (4096, 4096, 1, "out_proj"), (11264, 4096, 1, "linear_out")]:
    raw = synth.random_tensor(rng, synth.GGML_Q4_K, rows, k, 1.0 / np.sqrt(k))
    for wq in (True, False):
        us, st = msx.bench_gemm_batch_stamps(raw, k, 8, 6, 24, epi, wq)
        print(f"{name} with_quant={wq}: {us:.2f} us/iter")
        for it in (10, 11, 12):
            s = st[it]; live = s[:, 0] > 0
            t0 = s[live, 0].min()
            prev_end = st[it - 1][st[it - 1][:, 0] > 0, 4].max()
            rel = lambda j, f: f(s[live, j] - t0) / 1000.0
            print(f"  it {it}: prev main-loop end {(prev_end - t0)/1000:+.2f} | entry max {rel(0, np.max):.2f} | ring primed med {rel(1, np.median):.2f} | dep done med {rel(2, np.median):.2f} max {rel(2, np.max):.2f} | image med {rel(3, np.median):.2f} max {rel(3, np.max):.2f} | loop end min {rel(4, np.min):.2f} med {rel(4, np.median):.2f} max {rel(4, np.max):.2f} us")
