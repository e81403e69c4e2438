"""T1-vs-T2 gap table (CPU only): how far the reference-faithful arithmetic (T1 = oracle/ggml_ref.c: Q8_K / Q8_0 activation
re-quantisation, integer block dots, bf16 ring cache) sits from exact arithmetic on the same weights (T2 = the oracle's "ideal"
mode: dequantised weights, no activation quantisation, f32 ring).  The GPU path is asserted bit-identical to T1 in
tests/test_gpu_parity.py, so this table is the error bar that belongs next to every "GPU == oracle" statement: it bounds what a
different-but-valid ggml build (other SIMD summation order, other rounding of the same steps) could differ by.

Teacher-forced: both modes see the same input tokens every frame (T1's own outputs + seeded user codes).
usage: t1_t2_gap.py [frames] > profiles/r2_parity_gap.md"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import _pkgload; _pkgload.load()
from moshi_cpp_b200 import configs, synth
import oracle as orc

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 12
CASES = [("tiny", "q4_k"), ("tiny", "q8_0"), ("tiny_pplex", "q4_k"), ("tiny_stt", "q8_0"), ("moshi7b_l2", "q4_k"), ("moshi7b_l2", "q8_0"),
         ("moshi7b_l4", "q4_k")]


def mr(a, b):
    return float(np.max(np.abs(a.astype(np.float64) - b)) / max(np.max(np.abs(b)), 1e-30))


def margin(lg):
    s = np.sort(lg.astype(np.float64))
    return float((s[-1] - s[-2]) / max(np.max(np.abs(lg)), 1e-30))


print("| preset | weights | frames | text logits: max-rel T1 vs T2 (median / worst) | audio logits: max-rel T1 vs T2 (median / worst) | "
      "greedy tokens T1 == T2 | largest T2 top-1 margin among the differing tokens |")
print("|---|---|---|---|---|---|---|")
for preset, quant in CASES:
    cfg = configs.get(preset)
    path = synth.cached_gguf(preset, quant)
    m1 = orc.Model(path, cfg); s1 = orc.State(m1)
    m2 = orc.Model(path, cfg, ideal=True); s2 = orc.State(m2)
    rng = np.random.default_rng(42)
    n_q, dep_q = cfg["n_q"], cfg["dep_q"]
    toks = np.array([cfg["text_card"]] + [cfg["card"]] * n_q, dtype=np.int32)
    et, ea, same, total, worst_margin = [], [], 0, 0, None
    for f in range(frames):
        t1, l1, _ = s1.step_temporal(toks)
        t2, l2, _ = s2.step_temporal(toks)
        et.append(mr(l1, l2)); total += 1; same += int(t1 == t2)
        if t1 != t2:
            mg = margin(l2); worst_margin = mg if worst_margin is None else max(worst_margin, mg)
        nxt = [t1]
        if dep_q:
            a1, al1 = s1.step_depformer(t1)
            a2, al2 = s2.step_depformer(t1, force=a1)
            for k in range(dep_q):
                ea.append(mr(al1[k], al2[k])); total += 1; same += int(a1[k] == a2[k])
                if a1[k] != a2[k]:
                    mg = margin(al2[k]); worst_margin = mg if worst_margin is None else max(worst_margin, mg)
            nxt += list(a1)
        user = list(rng.integers(0, cfg["card"], size=n_q + 1 - len(nxt)))
        toks = np.array(nxt + user, dtype=np.int32)
    fa = f"{np.median(ea):.1e} / {np.max(ea):.1e}" if ea else "—"
    print(f"| {preset} | {quant} | {frames} | {np.median(et):.1e} / {np.max(et):.1e} | {fa} | {same} / {total} | "
          f"{'—' if worst_margin is None else f'{worst_margin:.1e}'} |", flush=True)
print()
print("T1 = `oracle/ggml_ref.c` (what the GPU path reproduces bit for bit), T2 = the same graph in exact-weight f32 arithmetic "
      "(`oracle.Model(..., ideal=True)`). Random-init weights; max-rel = max |a - b| / max |b| over one logit vector. "
      "Where the greedy tokens differ, the T2 top-1 margin (relative to max |logit|) is below the gap, i.e. the arg-max was a "
      "near-tie that the int8 activation noise decides.")
