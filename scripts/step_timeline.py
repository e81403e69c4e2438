"""In-kernel timeline of the persistent step kernels (all CTAs): per-family phase time split into wait+prologue / main loop /
epilogue for CTA 0, plus the spread over CTAs (skew) of every stage."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkgload; _pkgload.load()
from moshi_cpp_b200 import configs, synth, binding as msx
preset = sys.argv[1] if len(sys.argv) > 1 else "moshi7b"
quant = sys.argv[2] if len(sys.argv) > 2 else "q4_k"
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cfg = configs.get(preset); path = synth.cached_gguf(preset, quant)
m = msx.Model(path, cfg); s = msx.Stream(m, step_kernel=True)
rng = np.random.default_rng(0)
frames = rng.integers(0, cfg["card"], size=(32, cfg["n_q"] + 1)).astype(np.int32)
s.run_resident(frames, warm)
for rep in range(3):
    fams, st = s.step_timeline(frames[rep])
st = st.astype(np.float64) / 1e3                     # us
t0 = st[0, :, 0].min()
n_t = sum(1 for f in fams if not f.startswith("dep_"))
print(f"[{preset} {quant}] offset {s.offset}: temporal kernel {st[:n_t, :, 1].max() - st[0, :, 0].min():.1f} us"
      + (f", depformer kernel {st[n_t:, :, 1].max() - st[n_t, :, 0].min():.1f} us" if n_t < len(fams) else ""))
print(f"{'family':15s} {'n':>3s} {'avg us':>7s} | CTA0: {'pro':>5s} {'main':>5s} {'epi':>5s} | all CTAs: {'start spread':>12s} {'pro-end spread':>14s} {'main-end spread':>15s} {'end spread':>10s} {'last end - first start':>22s}")
agg = {}
for i, f in enumerate(fams):
    a = agg.setdefault(f, [])
    r = st[i]
    gem = r[0, 2] > 0
    act = r[:, 2] > 0 if gem else np.ones(r.shape[0], bool)
    a.append([r[0, 1] - r[0, 0], (r[0, 2] - r[0, 0]) if gem else 0, (r[0, 3] - r[0, 2]) if gem else 0, (r[0, 1] - r[0, 3]) if gem else 0,
              r[:, 0].max() - r[:, 0].min(), (r[act, 2].max() - r[act, 2].min()) if gem else 0, (r[act, 3].max() - r[act, 3].min()) if gem else 0,
              r[:, 1].max() - r[:, 1].min(), r[:, 1].max() - r[:, 0].min()])
for f, a in agg.items():
    a = np.array(a).mean(axis=0)
    print(f"{f:15s} {len(agg[f]):3d} {a[0]:7.2f} |       {a[1]:5.2f} {a[2]:5.2f} {a[3]:5.2f} |           {a[4]:12.2f} {a[5]:14.2f} {a[6]:15.2f} {a[7]:10.2f} {a[8]:22.2f}")
# finer CTA-0 breakdown of the GEMV phases: start -> input loaded -> rms -> prologue end -> main end -> stored -> end
print("CTA 0 stages (avg us):   load   rms  quant  main  reduce+store  tail")
for f in agg:
    idx = [i for i, g in enumerate(fams) if g == f and st[i, 0, 2] > 0]
    if not idx: continue
    r = st[idx, 0, :]
    print(f"  {f:15s} {np.mean(r[:,4]-r[:,0]):6.2f} {np.mean(r[:,5]-r[:,4]):5.2f} {np.mean(r[:,2]-r[:,5]):6.2f} {np.mean(r[:,3]-r[:,2]):5.2f} {np.mean(r[:,6]-r[:,3]):10.2f} {np.mean(r[:,1]-r[:,6]):8.2f}")
# detail of one temporal layer (layer 3) over CTAs: when does each CTA finish main / the phase
if n_t > 20:
    for i in range(16, 21):
        r = st[i] - t0
        print(f"  phase {i} {fams[i]:12s} start [{r[:,0].min():8.2f} {r[:,0].max():8.2f}]  pro-end [{r[:,2].min():8.2f} {r[:,2].max():8.2f}]  main-end [{r[:,3].min():8.2f} {r[:,3].max():8.2f}]  end [{r[:,1].min():8.2f} {r[:,1].max():8.2f}]")
