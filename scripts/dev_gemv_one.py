import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkgload; _pkgload.load()
from moshi_cpp_b200 import synth, binding as msx
L = msx.lib()
L.msx_bench_gemv.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
k, rows, pro = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
gt = synth.TYPE_NAMES["q4_k"]
raw = synth.random_tensor(np.random.default_rng(0), gt, rows, k, 1 / np.sqrt(k))
us = C.c_float(0)
assert L.msx_bench_gemv(0, gt, raw.ctypes.data, k, rows, 2, 6, pro, 0, C.byref(us)) == 0
print(us.value)
