"""dev: one long GEMV launch (many rows) for per-instruction stall sampling under ncu"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkgload; _pkgload.load()
import ctypes as C
from moshi_cpp_b200 import binding as msx, synth
L = msx.lib()
L.msx_bench_gemv.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
k, rows = 4096, 22528 * int(os.environ.get("MULT", 6))
rng = np.random.default_rng(1)
base = synth.random_tensor(rng, synth.GGML_Q4_K, 22528, k, 1.0 / np.sqrt(k))
raw = np.ascontiguousarray(np.tile(base, (rows // 22528, 1)))
us = C.c_float(0)
rc = L.msx_bench_gemv(0, synth.GGML_Q4_K, raw.ctypes.data, k, rows, 2, 4, 1, 2, C.byref(us))
print(rc, us.value, raw.size / us.value / 1e6, "TB/s")
