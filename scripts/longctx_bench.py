"""dev: frame time as the KV ring fills (n_valid 0 -> 3000) for moshi7b q4_k."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkgload; _pkgload.load()
from moshi_cpp_b200 import configs, synth, binding as msx
cfg = configs.get("moshi7b"); path = synth.cached_gguf("moshi7b", "q4_k")
m = msx.Model(path, cfg); s = msx.Stream(m)
rng = np.random.default_rng(0)
frames = rng.integers(0, cfg["card"], size=(32, cfg["n_q"] + 1)).astype(np.int32)
s.run_resident(frames, 20)
for target in (100, 500, 1000, 2000, 3000, 3300):
    while s.offset < target - 100:
        s.run_resident(frames, min(400, target - 100 - s.offset))
    ms, _ = s.run_resident(frames, 100)
    kvb = s.kv_bytes_next
    print(f"offset ~{s.offset:5d}: {ms/100:.3f} ms/frame  {100/ms*1e3:.1f} fps   KV bytes/frame {kvb/1e6:.0f} MB  -> step {(m.weight_bytes_per_frame + kvb)/(ms/100*1e-3)/1e9:.0f} GB/s")
_, fam = s.profile_frame(frames[0])
tot = sum(v[0] for v in fam.values())
print({k: (round(v[0] * 1e3 / v[1], 1), f"{100*v[0]/tot:.0f}%") for k, v in fam.items() if k in ("attn", "in_proj", "linear_in")})
