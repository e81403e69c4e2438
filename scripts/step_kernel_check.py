"""A/B of the persistent step kernel (MSX_STREAM_STEP_KERNEL) against the PDL-chained launches through the same C ABI, then timing.
usage: step_kernel_check.py [stage ...]   stages: tiny tiny8 pplex stt l2 time time8 oracle"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import _pkgload; _pkgload.load()
from moshi_cpp_b200 import configs, synth, binding as msx


def ab(preset, quant, n_frames, context=0):
    cfg = configs.get(preset); path = synth.cached_gguf(preset, quant)
    m = msx.Model(path, cfg)
    a = msx.Stream(m, context=context, step_kernel=True); b = msx.Stream(m, context=context)
    print(f"[{preset} {quant}] launches/frame: step-kernel {a.launches_per_frame}, chain {b.launches_per_frame}", flush=True)
    rng = np.random.default_rng(7)
    toks = np.array([cfg["text_card"]] + [cfg["card"]] * cfg["n_q"], dtype=np.int32)
    bad = 0
    for f in range(n_frames):
        ta, la, oa = a.step_temporal(toks); tb, lb, ob = b.step_temporal(toks)
        e_t = float(np.max(np.abs(la - lb))); e_o = float(np.max(np.abs(oa - ob)))
        msg = f"  frame {f}: text {ta} vs {tb}  |dlogits| {e_t:.3g}  |dtout| {e_o:.3g}"
        out_a = [ta]
        if cfg["dep_q"] > 0:
            aa, ala = a.step_depformer(tb); ab_, alb = b.step_depformer(tb)
            e_a = float(np.max(np.abs(ala - alb)))
            msg += f"  audio equal {np.array_equal(aa, ab_)}  |daudio logits| {e_a:.3g}"
            if not np.array_equal(aa, ab_) or e_a != 0.0:
                bad += 1
                k = int(np.argmax(np.max(np.abs(ala - alb), axis=1) > 0))
                msg += f"  first differing step {k}: {aa.tolist()} vs {ab_.tolist()}"
            out_a += list(ab_)
        if ta != tb or e_t != 0.0 or e_o != 0.0: bad += 1
        if f < 4 or bad: print(msg, flush=True)
        if bad > 3: break
        nxt = list(out_a) + list(rng.integers(0, cfg["card"], cfg["n_q"] + 1 - len(out_a)))
        toks = np.array(nxt[: cfg["n_q"] + 1], dtype=np.int32)
    print(f"[{preset} {quant}] {'OK bit-identical' if bad == 0 else 'MISMATCH'} over {f + 1} frames", flush=True)
    a.close(); b.close(); m.close()
    return bad == 0


def oracle_check(preset, quant, n_frames):
    import oracle
    cfg = configs.get(preset); path = synth.cached_gguf(preset, quant)
    m = msx.Model(path, cfg); s = msx.Stream(m, step_kernel=True)
    om = oracle.Model(path, cfg); os_ = oracle.State(om)
    rng = np.random.default_rng(3)
    toks = np.array([cfg["text_card"]] + [cfg["card"]] * cfg["n_q"], dtype=np.int32)
    ok = True
    for f in range(n_frames):
        tr, lr, _ = os_.step_temporal(toks); tg, lg, _ = s.step_temporal(toks)
        ar, alr = os_.step_depformer(tr); ag, alg = s.step_depformer(tr)
        e = max(float(np.max(np.abs(lg - lr))), float(np.max(np.abs(alg - alr))))
        same = tr == tg and np.array_equal(ar, ag)
        print(f"  oracle frame {f}: tokens equal {same}  max |dlogits| {e:.3g}", flush=True)
        ok &= same and e == 0.0
        toks = np.array([tr] + list(ar) + list(rng.integers(0, cfg["card"], cfg["n_q"] - cfg["dep_q"])), dtype=np.int32)
    print(f"[{preset} {quant}] oracle: {'OK' if ok else 'MISMATCH'}", flush=True)
    return ok


def timing(preset, quant, n=300, context=0):
    cfg = configs.get(preset); path = synth.cached_gguf(preset, quant)
    m = msx.Model(path, cfg)
    rng = np.random.default_rng(0)
    frames = rng.integers(0, cfg["card"], size=(32, cfg["n_q"] + 1)).astype(np.int32)
    gb = m.weight_bytes_per_frame / 1e9
    for name, sk in (("step-kernel", True), ("chain", False), ("step-kernel", True)):
        s = msx.Stream(m, context=context, step_kernel=sk)
        s.run_resident(frames, 30)
        ms, _ = s.run_resident(frames, n)
        print(f"[{preset} {quant}] {name:12s} launches/frame {s.launches_per_frame:4d}  {ms / n:.4f} ms/frame  {n / ms * 1e3:.1f} fps  "
              f"weights {gb / (ms / n) * 1e3:.0f} GB/s", flush=True)
        s.close()
    m.close()


if __name__ == "__main__":
    stages = sys.argv[1:] or ["tiny", "tiny8", "pplex", "stt", "l2", "time"]
    t0 = time.time()
    for st in stages:
        try:
            if st == "tiny": ab("tiny", "q4_k", 40)
            elif st == "tiny8": ab("tiny", "q8_0", 40)
            elif st == "pplex": ab("tiny_pplex", "q4_k", 30)
            elif st == "stt": ab("tiny_stt", "q8_0", 30)
            elif st == "l2": ab("moshi7b_l2", "q4_k", 6)
            elif st == "l2q8": ab("moshi7b_l2", "q8_0", 4)
            elif st == "oracle": oracle_check("tiny", "q4_k", 4)
            elif st == "time": timing("moshi7b", "q4_k")
            elif st == "time8": timing("moshi7b", "q8_0")
            elif st == "timel2": timing("moshi7b_l2", "q4_k")
        except Exception as e:
            print(f"stage {st} FAILED: {e}", flush=True)
        print(f"-- {st} done at {time.time() - t0:.0f} s", flush=True)
