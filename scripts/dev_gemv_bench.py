"""dev: GEMV kernel micro-benchmark over the 7B / depformer shapes (cold weights: rotating matrices > L2)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkgload; _pkgload.load()
from moshi_cpp_b200 import synth, binding as msx
L = msx.lib()
L.msx_bench_gemv.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
quant = sys.argv[1] if len(sys.argv) > 1 else "q4_k"
gt = synth.TYPE_NAMES[quant]
rng = np.random.default_rng(0)
shapes = [("in_proj", 4096, 12288, 1, 0), ("out_proj", 4096, 4096, 0, 1), ("linear_in", 4096, 22528, 1, 2), ("linear_out", 11264, 4096, 0, 1),
          ("text_head", 4096, 32000, 1, 3), ("dep_in_proj", 1024, 3072, 1, 0), ("dep_out_proj", 1024, 1024, 0, 1),
          ("dep_linear_in", 1024, 5632, 1, 2), ("dep_linear_out", 2816, 1024, 0, 1)]
tot_big = 0.0
for name, k, rows, pro, epi in shapes:
    raw = synth.random_tensor(rng, gt, rows, k, 1 / np.sqrt(k))
    nbytes = raw.nbytes
    n_mats = max(2, min(24, int(400e6 // nbytes) + 1))
    us = C.c_float(0)
    rc = L.msx_bench_gemv(0, gt, raw.ctypes.data, k, rows, n_mats, 200, pro, epi, C.byref(us))
    assert rc == 0, L.msx_last_error()
    gbs = nbytes / (us.value * 1e-6) / 1e9
    print(f"{name:15s} K={k:6d} rows={rows:6d} {nbytes/1e6:7.2f} MB  {us.value:7.2f} us  {gbs:7.0f} GB/s  ({gbs/6530*100:4.1f}% of 6530)")
    if name in ("in_proj", "out_proj", "linear_in", "linear_out"): tot_big += us.value
print(f"temporal layer GEMVs: {tot_big:.1f} us/layer -> {tot_big*32/1e3:.2f} ms per 32 layers (roofline 0.565 ms)")
