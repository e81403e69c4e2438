"""dev: single stream through the tensor-core unit kernel (MSX_MMA=1) vs the dp4a GEMV: frame time + per-shape replay"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkgload; _pkgload.load()
import ctypes as C
from moshi_cpp_b200 import binding as msx, configs, synth
L = msx.lib()
L.msx_bench_gemv.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
rng = np.random.default_rng(1)
shapes = [(4096, 12288, 1, 0, "in_proj"), (4096, 4096, 0, 1, "out_proj"), (4096, 22528, 1, 2, "linear_in"), (11264, 4096, 0, 1, "linear_out"),
          (4096, 1024, 0, 0, "dep_in"), (1024, 3072, 1, 0, "dep_in_proj"), (1024, 5632, 1, 2, "dep_linear_in"), (2816, 1024, 0, 1, "dep_linear_out"), (1024, 2048, 0, 3, "dep_head")]
for (k, rows, pro, epi, name) in shapes:
    raw = synth.random_tensor(rng, synth.GGML_Q4_K, rows, k, 1.0 / np.sqrt(k))
    res = []
    for mma in ("0", "1"):
        os.environ["MSX_MMA"] = mma
        us = C.c_float(0)
        rc = L.msx_bench_gemv(0, synth.GGML_Q4_K, raw.ctypes.data, k, rows, 8, 200, pro, epi, C.byref(us))
        res.append(us.value if rc == 0 else float("nan"))
    print(f"{name:14s} K={k:5d} rows={rows:5d}: dp4a {res[0]:6.2f} us   mma {res[1]:6.2f} us   ({raw.size/res[1]/1e6:5.2f} TB/s)", flush=True)
preset = os.environ.get("PRESET", "moshi7b")
cfg = configs.get(preset); path = synth.cached_gguf(preset, "q4_k")
frames = rng.integers(0, cfg["card"], size=(64, cfg["n_q"] + 1)).astype(np.int32)
toks = {}
for mma in ("0", "1"):
    os.environ["MSX_MMA"] = mma
    gm = msx.Model(path, cfg); st = msx.Stream(gm)
    st.run_resident(frames, 20)
    ms, tk = st.run_resident(frames, 200, want_tokens=True)
    toks[mma] = tk
    print(f"MSX_MMA={mma}: {ms/200:.3f} ms/frame  {200/(ms*1e-3):.1f} frames/s  launches {st.launches_per_frame}", flush=True)
    if mma == "1":
        _, fam = st.profile_frame(frames[0])
        print("  eager per-family us:", {k: round(v[0] * 1000 / max(1, v[1]), 1) for k, v in fam.items()})
    del st, gm
print("tokens identical:", bool(np.array_equal(toks["0"], toks["1"])))
