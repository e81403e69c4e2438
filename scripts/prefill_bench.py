"""Prompt prefill throughput: T prompt rows through msx_stream_prefill vs the same rows as serial steps.
usage: prefill_bench.py [preset] [quant] [T]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkgload; _pkgload.load()
from moshi_cpp_b200 import configs, synth, binding as msx
preset = sys.argv[1] if len(sys.argv) > 1 else "moshi7b"
quant = sys.argv[2] if len(sys.argv) > 2 else "q4_k"
T = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
cfg = configs.get(preset); path = synth.cached_gguf(preset, quant)
m = msx.Model(path, cfg); s = msx.Stream(m)
rng = np.random.default_rng(0)
rows = rng.integers(0, cfg["card"], size=(T, cfg["n_q"] + 1)).astype(np.int32)
rows[:, 0] = rng.integers(0, cfg["text_card"], size=T)
s.prefill(rows[:128]); s.reset()                      # warm: prefill context, graphs
best = 1e9
for rep in range(3):
    s.reset()
    t = time.perf_counter(); s.prefill(rows); dt = time.perf_counter() - t
    best = min(best, dt)
kv = [s.get_kv(l, h, sl) for l in (0, cfg["num_layers"] - 1) for h in (0, cfg["num_heads"] - 1) for sl in (0, T // 2, T - 1)]
print(f"[{preset} {quant}] prefill of {T} prompt frames: {best * 1e3:.1f} ms = {best / T * 1e3:.4f} ms per frame = {T / best:.0f} prompt frames/s", flush=True)
s.reset()
n = min(T, 256)
frames = rows[:n]
ms, _ = s.run_resident(frames, n)
print(f"[{preset} {quant}] the same rows as full decode steps: {ms / n:.4f} ms per frame ({n / ms * 1e3:.0f} frames/s) -> prefill is {ms / n / (best / T * 1e3):.1f}x faster per prompt frame")
# serial temporal-only steps for a KV comparison on a short prefix
s.reset()
for f in range(min(T, 96)):
    s.step_temporal(rows[f], want_logits=False)
ok = all(np.array_equal(s.get_kv(l, h, 0)[0], kv[i * 3][0]) for i, (l, h) in enumerate([(l, h) for l in (0, cfg["num_layers"] - 1) for h in (0, cfg["num_heads"] - 1)]))
print(f"KV rows of slot 0 identical to serial steps: {ok}")
