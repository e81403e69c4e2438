"""dev: batched-T prompt prefill vs one decode step per prompt frame (moshi7b q4_k)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkgload; _pkgload.load()
from moshi_cpp_b200 import binding as msx, configs, synth
preset = os.environ.get("PRESET", "moshi7b")
cfg = configs.get(preset); path = synth.cached_gguf(preset, "q4_k")
gm = msx.Model(path, cfg)
rng = np.random.default_rng(0)
T = int(os.environ.get("T", 512))
rows = rng.integers(0, cfg["card"], size=(T, cfg["n_q"] + 1)).astype(np.int32)
rows[:, 0] = rng.integers(0, cfg["text_card"], size=T)
a = msx.Stream(gm); ga = msx.Gen(a)
ga.prefill(rows[:16]); a.reset(); ga = msx.Gen(a)     # warm-up (builds the units layout and the prefill graph)
t0 = time.perf_counter(); ga.prefill(rows); t1 = time.perf_counter()
b = msx.Stream(gm); g = msx.Gen(b)
for f in range(8):
    g.step(rows[f])
b.reset(); g = msx.Gen(b)
t2 = time.perf_counter()
for f in range(T):
    g.step(rows[f])
t3 = time.perf_counter()
ka, va = a.get_kv(cfg["num_layers"] - 1, 3, T - 1); kb, vb = b.get_kv(cfg["num_layers"] - 1, 3, T - 1)
print(f"{preset}: {T} prompt frames: batched-T prefill {1e3*(t1-t0):.1f} ms ({1e3*(t1-t0)/T:.3f} ms/frame, {T/(t1-t0):.0f} frames/s)  "
      f"serial provided steps {1e3*(t3-t2):.1f} ms ({1e3*(t3-t2)/T:.3f} ms/frame)  speed-up {(t3-t2)/(t1-t0):.1f}x  last KV row identical: {bool(np.array_equal(ka, kb) and np.array_equal(va, vb))}")
def bf(x): return (x.astype(np.uint32) << 16).view(np.float32)
first = None
for slot in range(T):
    ka, va = a.get_kv(cfg["num_layers"] - 1, 3, slot); kb, vb = b.get_kv(cfg["num_layers"] - 1, 3, slot)
    if not (np.array_equal(ka, kb) and np.array_equal(va, vb)):
        first = slot; break
print("first differing slot (last layer, head 3):", first)
for slot in [0, 4, 8, 16, 32, 64, 128, 256, T - 1]:
    ka, va = a.get_kv(cfg["num_layers"] - 1, 3, slot); kb, vb = b.get_kv(cfg["num_layers"] - 1, 3, slot)
    d = float(np.max(np.abs(bf(va) - bf(vb))) / max(1e-30, np.max(np.abs(bf(vb)))))
    k0a, v0a = a.get_kv(0, 3, slot); k0b, v0b = b.get_kv(0, 3, slot)
    d0 = float(np.max(np.abs(bf(v0a) - bf(v0b))) / max(1e-30, np.max(np.abs(bf(v0b)))))
    print(f"  slot {slot:4d}: layer0 max-rel {d0:.2e}   last layer max-rel {d:.2e}")
