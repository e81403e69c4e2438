"""dev: fixed cost of a GEMV launch: tiny row counts (prologue + one tile) for the K values in use."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkgload; _pkgload.load()
from moshi_cpp_b200 import synth, binding as msx
L = msx.lib()
L.msx_bench_gemv.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
gt = synth.TYPE_NAMES["q4_k"]
rng = np.random.default_rng(0)
for k in (1024, 4096, 11264):
    for rows in (8, 1184, 4736):
        for pro in (0, 1):
            raw = synth.random_tensor(rng, gt, rows, k, 1 / np.sqrt(k))
            us = C.c_float(0)
            rc = L.msx_bench_gemv(0, gt, raw.ctypes.data, k, rows, 4, 200, pro, 0, C.byref(us))
            assert rc == 0, L.msx_last_error()
            print(f"K={k:6d} rows={rows:5d} {'RMS  ' if pro else 'PLAIN'}  {us.value:6.2f} us")
