"""dev: A/B an experimental library build (MSX_LIB_EXPERIMENT) on GEMV shapes and the full frame"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkgload; _pkgload.load()
import ctypes as C
from moshi_cpp_b200 import binding as msx, configs, synth
L = msx.lib()
L.msx_bench_gemv.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
rng = np.random.default_rng(1)
print("library:", os.environ.get("MSX_LIB_EXPERIMENT") or "default")
for q, gt in (("q4_k", synth.GGML_Q4_K), ("q8_0", synth.GGML_Q8_0)):
    for (k, rows, pro, epi, name) in [(4096, 12288, 1, 0, "in_proj"), (4096, 22528, 1, 2, "linear_in"), (11264, 4096, 0, 1, "linear_out"), (1024, 3072, 1, 0, "dep_in_proj")]:
        raw = synth.random_tensor(rng, gt, rows, k, 1.0 / np.sqrt(k))
        us = C.c_float(0)
        rc = L.msx_bench_gemv(0, gt, raw.ctypes.data, k, rows, 8, 200, pro, epi, C.byref(us))
        print(f"  {q} {name:12s}: {us.value:6.2f} us ({raw.size/us.value/1e6:5.2f} TB/s) rc={rc}", flush=True)
cfg = configs.get("moshi7b"); path = synth.cached_gguf("moshi7b", "q4_k")
frames = rng.integers(0, cfg["card"], size=(64, cfg["n_q"] + 1)).astype(np.int32)
gm = msx.Model(path, cfg); st = msx.Stream(gm)
st.run_resident(frames, 20)
ms, tk = st.run_resident(frames, 200, want_tokens=True)
print(f"  frame: {ms/200:.3f} ms  {200/(ms*1e-3):.1f} fps  token checksum {int(tk.astype(np.int64).sum())}")
