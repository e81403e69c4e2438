import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkgload; _pkgload.load()
from moshi_cpp_b200 import configs, synth, binding as msx
preset = sys.argv[1] if len(sys.argv) > 1 else "moshi7b"
cfg = configs.get(preset); path = synth.cached_gguf(preset, "q4_k")
m = msx.Model(path, cfg); s = msx.Stream(m, persistent_depformer=True)
rng = np.random.default_rng(0)
frames = rng.integers(0, cfg["card"], size=(8, cfg["n_q"] + 1)).astype(np.int32)
for i in range(4): s.step(frames[i])
L = msx.lib()
L.msx_debug_depformer_timeline.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
for rep in range(2):
    s.step_temporal(frames[4 + rep], want_logits=False)
    st = np.zeros((400, 8), dtype=np.int64); n = C.c_int(0)
    rc = L.msx_debug_depformer_timeline(s.h, 5, st.ctypes.data, 400, C.byref(n))
    assert rc == 0, L.msx_last_error()
st = st[: n.value]
t0 = st[0, 0]
print("phases", n.value, "total us", (st[-1, 0] - t0) / 1e3)
pro = (st[:, 1] - st[:, 0]) / 1e3; main = (st[:, 2] - st[:, 1]) / 1e3; epi = (st[:, 3] - st[:, 2]) / 1e3; bar = (st[:, 4] - st[:, 3]) / 1e3
for i in range(0, 30):
    print(f"p{i:3d} prologue {pro[i]:6.2f} main {main[i]:6.2f} tail {epi[i]:6.2f} barrier {bar[i]:6.2f}  | phase total {(st[i,4]-st[i,0])/1e3:6.2f}")
v = slice(0, n.value - 1)
print("TOTAL depformer kernel us", (st[-1, 0] - t0) / 1e3)
print("mean: prologue %.2f main %.2f tail %.2f barrier %.2f" % (pro[v].mean(), main[v].mean(), epi[v].mean(), bar[v].mean()))
L.msx_debug_barrier_timeline.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip())
for mode in (0, 2):
    st = np.zeros((64, 8), dtype=np.int64)
    assert L.msx_debug_barrier_timeline(s.h, 64, st.ctypes.data, mode) == 0
    print("mode", mode, "empty phase mean us %.2f" % ((st[40, 4] - st[1, 0]) / 39e3), " barrier only mean %.2f" % ((st[1:40, 4] - st[1:40, 3]).mean() / 1e3))

L.msx_debug_repeat_phase.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
names = {0: "dep_in(K4096,PLAIN,ADD_EMB)", 1: "in_proj(RMS,STORE)", 2: "attn+out_proj", 3: "lin_in(RMS,GATE)", 4: "lin_out(K2816,RESID)", 25: "head(ARGMAX)"}
for idx in (1, 2, 3, 4, 0):
    st = np.zeros((48, 8), dtype=np.int64)
    assert L.msx_debug_repeat_phase(s.h, idx, 48, st.ctypes.data) == 0, L.msx_last_error()
    v = slice(8, 47)
    print(f"   inside prologue: to-x-loaded {((st[v,5]-st[v,0]).mean())/1e3:5.2f}  reduce+sync {((st[v,6]-st[v,5]).mean())/1e3:5.2f}  scale+quantize {((st[v,7]-st[v,6]).mean())/1e3:5.2f}  final sync {((st[v,1]-st[v,7]).mean())/1e3:5.2f}")
    print(f"repeat {names[idx]:28s}: phase {((st[v,4]-st[v,0]).mean())/1e3:5.2f} us = prologue {((st[v,1]-st[v,0]).mean())/1e3:5.2f} main {((st[v,2]-st[v,1]).mean())/1e3:5.2f} tail {((st[v,3]-st[v,2]).mean())/1e3:5.2f} barrier {((st[v,4]-st[v,3]).mean())/1e3:5.2f}")
