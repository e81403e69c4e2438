"""Load-time quantiser throughput: one linear_in-sized tensor (22528 x 4096, bf16 source) through msx_test_quantize_rows
(H2D copy + kernel + D2H of the blocks).  python scripts/dev_quant_speed.py"""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkgload; _pkgload.load()
from moshi_cpp_b200 import binding as msx

rows, k = 22528, 4096
rng = np.random.default_rng(0)
x = (rng.standard_normal((rows, k), dtype=np.float32) * 0.02)
bits = (x.view(np.uint32) >> 16).astype(np.uint16)
for name, t in (("q8_0", 8), ("q4_0", 2), ("q4_k", 12)):
    msx.test_quantize_rows(t, bits[:64], src_type=30)
    t0 = time.perf_counter()
    msx.test_quantize_rows(t, bits, src_type=30)
    dt = time.perf_counter() - t0
    print(f"{name}: {rows * k / 1e6:.0f} M weights in {dt * 1e3:.1f} ms  ({rows * k / dt / 1e9:.2f} G weights/s incl. copies) -> 7.37 G weights ~ {7.37e9 / (rows * k / dt):.1f} s")
