"""dev: is the GEMV compute- or memory-bound?  Same kernel over 1 matrix (L2 resident after the first pass) vs 8 rotating"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkgload; _pkgload.load()
import ctypes as C
from moshi_cpp_b200 import binding as msx, synth
L = msx.lib()
L.msx_bench_gemv.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
rng = np.random.default_rng(1)
for (k, rows, pro, epi, name) in [(4096, 12288, 1, 0, "in_proj"), (4096, 22528, 1, 2, "linear_in"), (11264, 4096, 0, 1, "linear_out")]:
    raw = synth.random_tensor(rng, synth.GGML_Q4_K, rows, k, 1.0 / np.sqrt(k))
    for mma in ("0", "1"):
        os.environ["MSX_MMA"] = mma
        out = []
        for nm in (1, 8):
            us = C.c_float(0)
            L.msx_bench_gemv(0, synth.GGML_Q4_K, raw.ctypes.data, k, rows, nm, 200, pro, epi, C.byref(us))
            out.append(us.value)
        print(f"{name:10s} mma={mma}: L2-resident {out[0]:6.2f} us   HBM {out[1]:6.2f} us", flush=True)
