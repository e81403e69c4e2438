"""dev helper: per-frame GPU-vs-oracle error statistics (teacher forced)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import _pkgload; _pkgload.load()
from moshi_cpp_b200 import configs, synth, binding as msx
import oracle as orc

preset, quant, frames = sys.argv[1], sys.argv[2], int(sys.argv[3])
cfg = configs.get(preset)
path = synth.cached_gguf(preset, quant)
gm = msx.Model(path, cfg); gs = msx.Stream(gm)
om = orc.Model(path, cfg); os_ = orc.State(om)
oi = orc.Model(path, cfg, ideal=True); ois = orc.State(oi)
rng = np.random.default_rng(42)
n_q, dep_q = cfg["n_q"], cfg["dep_q"]
toks = np.array([cfg["text_card"]] + [cfg["card"]] * n_q, dtype=np.int32)
def mr(a, b): return float(np.max(np.abs(a.astype(np.float64) - b)) / np.max(np.abs(b)))
for f in range(frames):
    t_ref, lg_ref, to_ref = os_.step_temporal(toks)
    t_gpu, lg_gpu, to_gpu = gs.step_temporal(toks)
    t_id, lg_id, to_id = ois.step_temporal(toks)
    line = f"f{f:3d} text gpu-vs-T1 {mr(lg_gpu, lg_ref):.1e}  T1-vs-T2 {mr(lg_ref, lg_id):.1e} tout {mr(to_gpu,to_ref):.1e} tok {t_gpu==t_ref}"
    if dep_q:
        a_ref, al_ref = os_.step_depformer(t_ref)
        a_gpu, al_gpu = gs.step_depformer(t_ref, force=a_ref)
        a_id, al_id = ois.step_depformer(t_ref, force=a_ref)
        line += " | audio " + " ".join(f"{mr(al_gpu[k], al_ref[k]):.0e}" for k in range(dep_q))
        line += " | T1vsT2 " + " ".join(f"{mr(al_ref[k], al_id[k]):.0e}" for k in range(dep_q))
        nxt = [t_ref] + list(a_ref)
    else:
        nxt = [t_ref]
    print(line, flush=True)
    user = list(rng.integers(0, cfg["card"], size=n_q + 1 - len(nxt)))
    toks = np.array(nxt + user, dtype=np.int32)
