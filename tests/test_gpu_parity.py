"""GPU parity tests: the CUDA path (through the C ABI, include/moshi_b200.h) against the CPU oracle
(oracle/ggml_ref.c) on identical random-init GGUF weights and synthetic tokens.

Bars (BASELINE.json north_star): q4_k / q8_0 dequantisation bit-exact; per-step logits within
max-rel 2e-3; greedy text and audio tokens identical wherever the top-2 margin exceeds that tolerance.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LOGIT_TOL = 2e-3      # max |gpu - oracle| / max |oracle|   (north_star bar)
# Both sides accumulate exact block terms in double and round once (DESIGN.md "Order-independent
# arithmetic"), so the results are expected to be bit-identical; a last-bit difference can only come
# from a double-rounding coincidence (~1e-8 per output).
GEMV_TOL = 1e-6
EXACT_FRACTION = 0.999


def assert_bitwise_mostly(a, b, what=""):
    same = float(np.mean(a.view(np.uint32) == b.view(np.uint32)))
    assert same >= EXACT_FRACTION, f"{what}: only {same:.5f} of the values are bit-identical"


def max_rel(a, b):
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))) / max(1e-30, float(np.max(np.abs(b)))))


def top2_margin(logits):
    s = np.sort(logits.astype(np.float64))
    return float(s[-1] - s[-2])


@pytest.fixture(scope="module")
def msx():
    from moshi_cpp_b200 import binding
    assert binding.lib().msx_device_count() > 0, "no CUDA device: these tests must run on the B200 box"
    return binding


@pytest.fixture(scope="module")
def orc():
    import oracle
    oracle.lib()
    return oracle


# ---------------------------------------------------------------------------------------------------
# T0: block formats
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("quant,k,rows", [("q4_k", 4096, 64), ("q4_k", 11264, 8), ("q4_k", 1024, 40), ("q4_k", 2816, 24),
                                          ("q8_0", 4096, 32), ("q8_0", 1024, 48), ("q8_0", 2816, 16), ("q8_0", 96, 8)])
def test_repacked_dequant_bit_exact(msx, quant, k, rows):
    from gguf.quants import dequantize
    from gguf import GGMLQuantizationType as QT
    from moshi_cpp_b200 import synth
    rng = np.random.default_rng(7)
    gt = synth.TYPE_NAMES[quant]
    raw = synth.random_tensor(rng, gt, rows, k, 0.02)
    ref = dequantize(raw, QT(gt)).reshape(rows, k)
    got = msx.test_dequant_repacked(gt, raw, k)
    assert np.array_equal(ref.view(np.uint32), got.view(np.uint32))


@pytest.mark.parametrize("quant", ["q4_0", "q8_0", "f32", "f16", "bf16"])
def test_embedding_rows_bit_exact(msx, orc, quant):
    from moshi_cpp_b200 import synth
    rng = np.random.default_rng(11)
    gt = synth.TYPE_NAMES[quant]
    k, rows = 1024, 300
    raw = synth.random_tensor(rng, gt, rows, k, 0.25)
    ids = np.array([0, 1, 17, 299, 5, 5], dtype=np.int32)
    got = msx.test_dequant_rows(gt, raw, k, ids)
    ref = orc.dequantize(gt, raw[ids], k)
    assert np.array_equal(ref.view(np.uint32), got.view(np.uint32))


# ---------------------------------------------------------------------------------------------------
# fused GEMV (activation quantisation + dp4a dot + block scales) vs the ggml-CPU-faithful oracle
# ---------------------------------------------------------------------------------------------------
GEMV_SHAPES = [(4096, 512), (4096, 12288), (11264, 256), (1024, 3072), (2816, 1024), (1024, 2048), (4096, 6), (512, 10), (768, 512)]


@pytest.mark.parametrize("quant", ["q4_k", "q8_0"])
@pytest.mark.parametrize("k,rows", GEMV_SHAPES)
@pytest.mark.parametrize("rms", [False, True])
def test_gemv_vs_oracle(msx, orc, quant, k, rows, rms):
    from moshi_cpp_b200 import synth
    if quant == "q4_k" and k % 256:
        pytest.skip("q4_k needs K % 256 == 0")
    rng = np.random.default_rng(k * 31 + rows)
    gt = synth.TYPE_NAMES[quant]
    raw = synth.random_tensor(rng, gt, rows, k, 1.0 / np.sqrt(k))
    x = rng.standard_normal(k).astype(np.float32) * 1.7
    x[3] = 0.0
    alpha = (1.0 + 0.1 * rng.standard_normal(k)).astype(np.float32) if rms else None
    xin = orc.rms_norm(x, alpha) if rms else x
    ref = orc.mul_mat_vec(gt, raw, k, xin)
    got = msx.test_gemv(gt, raw, k, x, alpha)
    assert max_rel(got, ref) < GEMV_TOL
    assert_bitwise_mostly(got, ref, "gemv")
    # and both sit within quantisation noise of the exact contraction
    ideal = orc.mul_mat_vec(gt, raw, k, xin, ideal=True)
    assert max_rel(got, ideal) < 3e-2


def test_gemv_zero_and_constant_blocks(msx, orc):
    """edge cases of quantize_row_q8_K: an all-zero 256-block (d = 0) and ties for the max-|x| carrier"""
    from moshi_cpp_b200 import synth
    rng = np.random.default_rng(5)
    k, rows = 1024, 64
    raw = synth.random_tensor(rng, synth.GGML_Q4_K, rows, k, 0.03)
    x = rng.standard_normal(k).astype(np.float32)
    x[256:512] = 0.0
    x[512:768] = 0.5
    x[768] = -2.0; x[900] = 2.0
    ref = orc.mul_mat_vec(synth.GGML_Q4_K, raw, k, x)
    got = msx.test_gemv(synth.GGML_Q4_K, raw, k, x)
    assert max_rel(got, ref) < GEMV_TOL


# ---------------------------------------------------------------------------------------------------
# whole-step parity
# ---------------------------------------------------------------------------------------------------
def run_teacher_forced(msx, orc, path, cfg, n_frames, seed=42, check_kv=True, step_kernel=False):
    """Drive both implementations with the ORACLE's token history; compare logits every step."""
    gm = msx.Model(path, cfg); gs = msx.Stream(gm, step_kernel=step_kernel)
    if step_kernel:
        assert gs.launches_per_frame == (2 if cfg["dep_q"] > 0 else 1), "the persistent step kernel did not take this model"
    om = orc.Model(path, cfg); os_ = orc.State(om)
    rng = np.random.default_rng(seed)
    n_q, dep_q = cfg["n_q"], cfg["dep_q"]
    toks = np.array([cfg["text_card"]] + [cfg["card"]] * n_q, dtype=np.int32)
    worst_text, worst_audio, n_flip = 0.0, 0.0, 0
    n_exact, n_cmp = 0, 0
    for f in range(n_frames):
        t_ref, lg_ref, to_ref = os_.step_temporal(toks)
        t_gpu, lg_gpu, to_gpu = gs.step_temporal(toks)
        scale = float(np.max(np.abs(lg_ref)))
        worst_text = max(worst_text, max_rel(lg_gpu, lg_ref))
        assert max_rel(lg_gpu, lg_ref) < LOGIT_TOL, f"frame {f}: text logits"
        assert max_rel(to_gpu, to_ref) < LOGIT_TOL, f"frame {f}: transformer_out"
        n_cmp += 1; n_exact += int(np.array_equal(lg_gpu.view(np.uint32), lg_ref.view(np.uint32)))
        if t_gpu != t_ref:
            n_flip += 1
            assert top2_margin(lg_ref) <= LOGIT_TOL * scale, f"frame {f}: text token {t_gpu} != {t_ref} with a clear margin"
        if dep_q > 0:
            a_ref, al_ref = os_.step_depformer(t_ref)
            a_gpu, al_gpu = gs.step_depformer(t_ref, force=a_ref)     # feed the oracle's tokens forward
            for k in range(dep_q):
                worst_audio = max(worst_audio, max_rel(al_gpu[k], al_ref[k]))
                assert max_rel(al_gpu[k], al_ref[k]) < LOGIT_TOL, f"frame {f} codebook {k}: audio logits"
                n_cmp += 1; n_exact += int(np.array_equal(al_gpu[k].view(np.uint32), al_ref[k].view(np.uint32)))
                if a_gpu[k] != a_ref[k]:
                    n_flip += 1
                    assert top2_margin(al_ref[k]) <= LOGIT_TOL * float(np.max(np.abs(al_ref[k]))), f"frame {f} cb {k}: token mismatch with a clear margin"
            nxt = [t_ref] + list(a_ref)
        else:
            nxt = [t_ref]
        user = list(rng.integers(0, cfg["card"], size=n_q + 1 - len(nxt)))
        toks = np.array(nxt + user, dtype=np.int32)
        if f % 7 == 3:
            toks[1 + (f % n_q)] = -1          # zero embedding
        if f % 11 == 5:
            toks[0] = -2                      # "ungenerated" -> row 0
    if check_kv:
        # KV rows of the last written slot
        slot = (n_frames - 1) % cfg["context"]
        for layer in (0, cfg["num_layers"] - 1):
            for head in (0, cfg["num_heads"] - 1):
                kg, vg = gs.get_kv(layer, head, slot); ko, vo = os_.get_kv(layer, head, slot)
                assert np.array_equal(kg, ko) and np.array_equal(vg, vo), "KV ring rows must be bit-identical"
    # stronger than the north_star bar: logits are bit-identical step after step
    assert n_exact >= 0.98 * n_cmp, f"only {n_exact}/{n_cmp} logit vectors bit-identical"
    return worst_text, worst_audio, n_flip


@pytest.mark.parametrize("preset,quant,frames", [("tiny", "q4_k", 80), ("tiny", "q8_0", 60), ("tiny_pplex", "q4_k", 40), ("tiny_stt", "q8_0", 50)])
def test_step_parity_small(msx, orc, gguf_for, preset, quant, frames):
    path, cfg = gguf_for(preset, quant)
    wt, wa, flips = run_teacher_forced(msx, orc, path, cfg, frames)
    print(f"{preset}/{quant}: worst text max-rel {wt:.2e}, audio {wa:.2e}, margin-excused flips {flips}")


@pytest.mark.parametrize("preset,quant,frames", [("tiny", "q4_k", 60), ("tiny", "q8_0", 40), ("tiny_pplex", "q4_k", 30), ("tiny_stt", "q8_0", 30),
                                                 ("tiny_sched", "q4_k", 20), ("moshi7b_l2", "q4_k", 4), ("moshi7b_l2", "q8_0", 3)])
def test_step_kernel_parity(msx, orc, gguf_for, preset, quant, frames):
    """MSX_STREAM_STEP_KERNEL (one persistent cooperative kernel per stack: TMA weight ring, flag-in-data activation exchange)
    against the oracle: same bar as the launch chain — logits bit-identical step after step, KV rows bit-identical."""
    path, cfg = gguf_for(preset, quant)
    wt, wa, flips = run_teacher_forced(msx, orc, path, cfg, frames, step_kernel=True)
    print(f"step kernel {preset}/{quant}: worst text max-rel {wt:.2e}, audio {wa:.2e}, margin-excused flips {flips}")


def test_depformer_weight_schedule(msx, orc, gguf_for):
    """non-identity depformer_weights_per_step_schedule (lm.h:457-462, transformer.h:74-83): in_projs / out_projs / gating /
    depformer_in of step k come from weight set schedule[k]; embeddings and linears stay per step"""
    path, cfg = gguf_for("tiny_sched", "q4_k")
    assert cfg["schedule"] and cfg["schedule"] != list(range(cfg["dep_q"]))
    wt, wa, flips = run_teacher_forced(msx, orc, path, cfg, 40)
    print(f"tiny_sched: worst text max-rel {wt:.2e}, audio {wa:.2e}, flips {flips}")


@pytest.mark.parametrize("quant", ["q4_k", "q8_0"])
def test_step_parity_7b_shapes(msx, orc, gguf_for, quant):
    """Full-size 7B layer / head / depformer shapes (2 temporal layers so the oracle finishes in seconds)."""
    path, cfg = gguf_for("moshi7b_l2", quant)
    wt, wa, flips = run_teacher_forced(msx, orc, path, cfg, 4)
    print(f"moshi7b_l2/{quant}: worst text max-rel {wt:.2e}, audio {wa:.2e}, flips {flips}")


def test_greedy_tokens_250_frames(msx, orc, gguf_for):
    """LMGen end to end (delay ring + step) for 250 frames, free running on both sides."""
    path, cfg = gguf_for("tiny", "q4_k")
    gm = msx.Model(path, cfg); gs = msx.Stream(gm); gg = msx.Gen(gs)
    om = orc.Model(path, cfg); og = orc.LMGen(om)
    rng = np.random.default_rng(42)
    n_user = cfg["n_q"] - cfg["dep_q"]
    emitted = 0
    for f in range(250):
        user = rng.integers(0, cfg["card"], size=n_user).astype(np.int32)
        ok_o, t_o, a_o = og.step(user)
        ok_g, t_g, a_g = gg.step(user)
        assert ok_g == ok_o, f"frame {f}: emit flag"
        assert gg.offset == og.offset
        if ok_o:
            emitted += 1
            assert t_g == t_o and np.array_equal(a_g, a_o), f"frame {f}: tokens differ"
        else:
            assert np.array_equal(a_g, a_o)
    assert emitted == 250 - max(cfg["delays"])


def test_gen_provided_and_replace(msx, orc, gguf_for):
    """all-streams-provided input (prompt replay, lm.h:812-817) and depformer_replace_tokens"""
    path, cfg = gguf_for("tiny", "q4_k")
    gm = msx.Model(path, cfg); gs = msx.Stream(gm); gg = msx.Gen(gs, delay_steps=2)
    ocfg = dict(cfg)
    om = orc.Model(path, ocfg, delay_steps=2); og = orc.LMGen(om)
    rng = np.random.default_rng(3)
    for f in range(12):
        if f < 4:
            toks = rng.integers(0, cfg["card"], size=cfg["n_q"] + 1).astype(np.int32)
        else:
            toks = rng.integers(0, cfg["card"], size=cfg["n_q"] - cfg["dep_q"]).astype(np.int32)
        rep = f < 2
        r_o = og.step(toks, replace=rep); r_g = gg.step(toks, replace=rep)
        assert r_o[0] == r_g[0] and np.array_equal(r_o[2], r_g[2]), f"frame {f}"
        if r_o[0]:
            assert r_o[1] == r_g[1]


def test_vad_head(msx, orc, gguf_for):
    path, cfg = gguf_for("tiny_stt", "q8_0")
    gm = msx.Model(path, cfg); gs = msx.Stream(gm)
    om = orc.Model(path, cfg); os_ = orc.State(om)
    rng = np.random.default_rng(9)
    for f in range(5):
        toks = rng.integers(0, cfg["card"], size=cfg["n_q"] + 1).astype(np.int32)
        os_.step_temporal(toks); gs.step_temporal(toks, want_logits=False)
        assert abs(gs.vad() - os_.vad()) < 1e-4


def test_resident_replay_matches_host_steps(msx, gguf_for):
    """msx_run_resident (tokens resident in HBM, no host round trips) produces the same tokens as msx_step"""
    path, cfg = gguf_for("tiny", "q4_k")
    gm = msx.Model(path, cfg)
    rng = np.random.default_rng(1)
    frames = rng.integers(0, cfg["card"], size=(16, cfg["n_q"] + 1)).astype(np.int32)
    frames[:, 0] = rng.integers(0, cfg["text_card"], size=16)
    s1 = msx.Stream(gm)
    host = np.stack([s1.step(frames[i % 16]) for i in range(40)])
    s2 = msx.Stream(gm)
    ms, dev = s2.run_resident(frames, 40, want_tokens=True)
    assert ms > 0 and np.array_equal(host, dev)
    assert s2.offset == 40
    # and the stream keeps working in host mode afterwards
    a = s1.step(frames[3]); b = s2.step(frames[3])
    assert np.array_equal(a, b)


def test_stream_reset_and_context_override(msx, gguf_for):
    path, cfg = gguf_for("tiny", "q4_k")
    gm = msx.Model(path, cfg)
    rng = np.random.default_rng(2)
    frames = rng.integers(0, cfg["card"], size=(30, cfg["n_q"] + 1)).astype(np.int32)
    s = msx.Stream(gm)
    a = np.stack([s.step(f) for f in frames])
    s.reset()
    assert s.offset == 0
    b = np.stack([s.step(f) for f in frames])
    assert np.array_equal(a, b)
    # smaller ring (tools' "-c N"): identical until the smaller ring wraps
    s8 = msx.Stream(gm, context=8)
    c = np.stack([s8.step(f) for f in frames[:8]])
    assert np.array_equal(a[:8], c)


def test_errors(msx, gguf_for, tmp_path):
    path, cfg = gguf_for("tiny", "q4_k")
    with pytest.raises(msx.MsxError) as e:
        msx.Model(str(tmp_path / "missing.gguf"), cfg)
    assert e.value.code == -2
    bad = tmp_path / "bad.gguf"
    bad.write_bytes(b"NOPE" + b"\0" * 64)
    with pytest.raises(msx.MsxError) as e:
        msx.Model(str(bad), cfg)
    assert e.value.code == -3
    wrong = dict(cfg); wrong["text_card"] = 999        # embedding table shape no longer matches
    with pytest.raises(msx.MsxError) as e:
        msx.Model(path, wrong)
    assert e.value.code == -3


def test_out_of_range_tokens_are_rejected(msx, gguf_for):
    """token ids index embedding tables on the device: ids outside a table are refused on the host with MSX_ERR_ARG (-1) on
    every input path (per-frame steps, depformer overrides, resident replay, batch, prefill) and the stream keeps working
    (a stray id used to read device memory out of bounds)."""
    path, cfg = gguf_for("tiny", "q4_k")
    gm = msx.Model(path, cfg); gs = msx.Stream(gm)
    good = np.array([cfg["text_card"]] + [cfg["card"]] * cfg["n_q"], dtype=np.int32)
    t0, _, _ = gs.step_temporal(good)
    for pos, val in ((0, cfg["text_card"] + 1), (0, -3), (3, cfg["card"] + 1), (cfg["n_q"], 1 << 30), (5, -7)):
        bad = good.copy(); bad[pos] = val
        for call in (lambda: gs.step_temporal(bad), lambda: gs.step(bad), lambda: gs.run_resident(np.stack([good, bad]), 2),
                     lambda: gs.prefill(np.stack([good, bad]))):
            with pytest.raises(msx.MsxError) as e:
                call()
            assert e.value.code == -1
    for call in (lambda: gs.step_depformer(cfg["text_card"] + 1), lambda: gs.step_depformer(3, force=[cfg["card"] + 1] * cfg["dep_q"]),
                 lambda: gs.step_depformer(3, force=[-1] * cfg["dep_q"])):
        with pytest.raises(msx.MsxError) as e:
            call()
        assert e.value.code == -1
    batch = msx.Batch(gm, 2)
    bad = good.copy(); bad[2] = cfg["card"] + 5
    with pytest.raises(msx.MsxError) as e:
        batch.step(np.stack([good, bad]))
    assert e.value.code == -1
    gs.reset()
    t1, _, _ = gs.step_temporal(good)
    assert t1 == t0                                           # nothing was consumed or corrupted by the refused calls
    for tok in (-1, -2):                                      # the two negative ids the reference defines stay legal
        ok = good.copy(); ok[1] = tok
        gs.step_temporal(ok)


@pytest.mark.parametrize("preset,quant", [("tiny", "q4_k"), ("tiny_pplex", "q8_0"), ("moshi7b_l2", "q4_k")])
def test_step_kernel_matches_launch_chain(msx, gguf_for, preset, quant):
    """The persistent step kernels and the one-launch-per-op path share their arithmetic and accumulate order-independently:
    logits and tokens must be bit-identical (extra check; the parity evidence is test_step_kernel_parity against the oracle)."""
    path, cfg = gguf_for(preset, quant)
    gm = msx.Model(path, cfg)
    a = msx.Stream(gm, step_kernel=True); b = msx.Stream(gm)
    assert a.launches_per_frame == 2 and a.launches_per_frame < b.launches_per_frame
    rng = np.random.default_rng(5)
    for f in range(30 if preset != "moshi7b_l2" else 4):
        toks = rng.integers(0, cfg["card"], size=cfg["n_q"] + 1).astype(np.int32)
        ta, la, oa = a.step_temporal(toks); tb, lb, ob = b.step_temporal(toks)
        assert ta == tb and np.array_equal(la.view(np.uint32), lb.view(np.uint32))
        xa, xla = a.step_depformer(ta); xb, xlb = b.step_depformer(tb)
        assert np.array_equal(xa, xb), f"frame {f}"
        assert np.array_equal(xla.view(np.uint32), xlb.view(np.uint32)), f"frame {f}"
    # sampling is not in the step kernel: the stream falls back to the launch chain and keeps working
    a.set_sampling(0.8, 0.8)
    assert a.launches_per_frame > 2



# ---------------------------------------------------------------------------------------------------
# parity where the benchmark runs: full-size models, long rings, prefill and batches against the ORACLE
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("step_kernel", [False, True])
def test_full_7b_q4k_250_frames_vs_oracle(msx, orc, gguf_for, step_kernel):
    """BASELINE.json config 3 at FULL size (32 layers, the GGUF bench.py times): 250 frames of free-running greedy LMGen —
    text and audio tokens identical to the oracle wherever the top-2 margin exceeds the 2e-3 logit tolerance (north_star) —
    then teacher-forced logits <= 2e-3 max-rel for a few more frames."""
    path, cfg = gguf_for("moshi7b", "q4_k")
    gm = msx.Model(path, cfg); gs = msx.Stream(gm, step_kernel=step_kernel); gg = msx.Gen(gs)
    om = orc.Model(path, cfg); og = orc.LMGen(om)
    rng = np.random.default_rng(42)
    n_user = cfg["n_q"] - cfg["dep_q"]
    emitted = 0
    for f in range(250):
        user = rng.integers(0, cfg["card"], size=n_user).astype(np.int32)
        ok_o, t_o, a_o = og.step(user); ok_g, t_g, a_g = gg.step(user)
        assert ok_g == ok_o and gg.offset == og.offset, f"frame {f}: emit flag / offset"
        assert t_g == t_o and np.array_equal(a_g, a_o), f"frame {f}: tokens differ from the oracle"
        emitted += int(ok_o)
    assert emitted == 250 - max(cfg["delays"])
    # teacher-forced logits on the same state (the oracle's LMGen owns an orc.State: compare through fresh streams instead)
    wt, wa, flips = run_teacher_forced(msx, orc, path, cfg, 3, step_kernel=step_kernel)
    assert flips == 0
    print(f"moshi7b q4_k (step_kernel={step_kernel}): 250 free-running frames identical; teacher-forced worst max-rel text {wt:.2e} audio {wa:.2e}")


def test_full_7b_q4k_tcgen05_prefill_and_wide_batch_vs_oracle(msx, orc, gguf_for):
    """The tcgen05 paths at FULL size (32 layers, the GGUF bench.py times): (a) a 70-row prompt through msx_stream_prefill
    (one 64-column pass + a 6-column tail on tc_matmul_q4k_kernel) against 70 serial oracle steps — KV rows bit-identical in the first
    and last layer, next-frame logits within tolerance; (b) one frame of a 16-stream batch (tc_matmul_q4k_kernel<16>, all 32 layers +
    depformer) against 16 oracle states."""
    path, cfg = gguf_for("moshi7b", "q4_k")
    gm = msx.Model(path, cfg); om = orc.Model(path, cfg)
    rng = np.random.default_rng(77)
    T = 70
    gs = msx.Stream(gm, context=128); os_ = orc.State(om)
    rows = rng.integers(0, cfg["card"], size=(T, cfg["n_q"] + 1)).astype(np.int32)
    rows[:, 0] = rng.integers(0, cfg["text_card"], size=T)
    gs.prefill(rows)
    for f in range(T):
        os_.step_temporal(rows[f])
    for layer in (0, cfg["num_layers"] - 1):
        for head in (0, cfg["num_heads"] - 1):
            for slot in (0, 63, 64, T - 1):
                kg, vg = gs.get_kv(layer, head, slot); ko, vo = os_.get_kv(layer, head, slot)
                assert np.array_equal(kg, ko) and np.array_equal(vg, vo), f"KV row layer {layer} head {head} slot {slot}"
    toks = rng.integers(0, cfg["card"], size=cfg["n_q"] + 1).astype(np.int32)
    t_ref, lg_ref, _ = os_.step_temporal(toks); t_gpu, lg_gpu, _ = gs.step_temporal(toks)
    assert max_rel(lg_gpu, lg_ref) < LOGIT_TOL
    assert_bitwise_mostly(lg_gpu, lg_ref, "text logits after the 70-row prompt")
    gs.close()
    n = 16
    batch = msx.Batch(gm, n, 64)
    btoks = rng.integers(0, cfg["card"], size=(n, cfg["n_q"] + 1)).astype(np.int32)
    btoks[:, 0] = rng.integers(0, cfg["text_card"], size=n)
    out = batch.step(btoks)
    for s in range(n):
        st = orc.State(om)
        t_ref, lg_ref, _ = st.step_temporal(btoks[s]); a_ref, al_ref = st.step_depformer(t_ref)
        btl, bal = batch.logits(s)
        assert max_rel(btl, lg_ref) < LOGIT_TOL and max_rel(bal, al_ref) < LOGIT_TOL, f"stream {s}: logits vs the oracle"
        assert_bitwise_mostly(btl, lg_ref, f"text logits stream {s}")
        if out[s, 0] == t_ref:
            assert np.array_equal(out[s, 1:], a_ref), f"stream {s}: audio tokens"
        else:
            assert top2_margin(lg_ref) <= LOGIT_TOL * float(np.max(np.abs(lg_ref)))


@pytest.mark.parametrize("step_kernel", [False, True])
def test_long_ring_7b_shapes_q8_0_vs_oracle(msx, orc, gguf_for, step_kernel):
    """BASELINE.json config 4's operating point at 7B layer shapes: PersonaPlex (dep_q 16) q8_0, the ring filled past its
    capacity (1100 slots, > 1024: the split long-ring attention path with every slot valid), then 8 frames against the
    oracle: logits, tokens and KV rows."""
    path, cfg = gguf_for("pplex7b_l2_c1100", "q8_0")
    gm = msx.Model(path, cfg); gs = msx.Stream(gm, step_kernel=step_kernel)
    om = orc.Model(path, cfg); os_ = orc.State(om)
    rng = np.random.default_rng(17)
    n_q = cfg["n_q"]
    fill = cfg["context"] + 7
    for f in range(fill):                                    # temporal steps only: the depformer does not touch the ring
        toks = rng.integers(0, cfg["card"], size=n_q + 1).astype(np.int32)
        toks[0] = rng.integers(0, cfg["text_card"])
        os_.step_temporal(toks)
        gs.step_temporal(toks, want_logits=False)
    assert gs.offset == fill
    for f in range(8):
        toks = rng.integers(0, cfg["card"], size=n_q + 1).astype(np.int32)
        toks[0] = rng.integers(0, cfg["text_card"])
        t_ref, lg_ref, to_ref = os_.step_temporal(toks); t_gpu, lg_gpu, to_gpu = gs.step_temporal(toks)
        assert max_rel(lg_gpu, lg_ref) < LOGIT_TOL and max_rel(to_gpu, to_ref) < LOGIT_TOL, f"frame {f}: text logits with a full ring"
        assert t_gpu == t_ref or top2_margin(lg_ref) <= LOGIT_TOL * float(np.max(np.abs(lg_ref)))
        a_ref, al_ref = os_.step_depformer(t_ref); a_gpu, al_gpu = gs.step_depformer(t_ref, force=a_ref)
        assert max_rel(al_gpu, al_ref) < LOGIT_TOL, f"frame {f}: audio logits"
        assert_bitwise_mostly(lg_gpu, lg_ref, f"text logits frame {f}")
    slot = (fill + 7) % cfg["context"]
    for layer in (0, cfg["num_layers"] - 1):
        for head in (0, 13, cfg["num_heads"] - 1):
            for sl in (slot, 0, cfg["context"] - 1):
                kg, vg = gs.get_kv(layer, head, sl); ko, vo = os_.get_kv(layer, head, sl)
                assert np.array_equal(kg, ko) and np.array_equal(vg, vo), f"KV row layer {layer} head {head} slot {sl}"


@pytest.mark.parametrize("preset,T,quant", [("tiny", 19, "q4_k"), ("tiny_pplex", 13, "q8_0"), ("moshi7b_l2", 11, "q4_k"), ("moshi7b_l2", 139, "q4_k"),
                                            ("moshi7b_l2", 70, "q8_0"), ("tiny", 61, "q4_k"), ("tiny_pplex", 45, "q8_0")])
def test_batched_prefill_vs_oracle(msx, orc, gguf_for, preset, T, quant):
    """Prompt prefill against the ORACLE's T serial steps: KV rows bit-identical, logits of the frames that follow within tolerance
    (the prefill was only compared with the serial GPU path).  Q4_K models run 64 positions per weight pass on the
    tcgen05 kind::i8 GEMM (tc_gemm.cuh: T = 139 = two full passes + a tail of 11 columns), Q8_0 models 8 per pass on mma.sync."""
    path, cfg = gguf_for(preset, quant)
    gm = msx.Model(path, cfg); gs = msx.Stream(gm)
    om = orc.Model(path, cfg); os_ = orc.State(om)
    rng = np.random.default_rng(31)
    rows = rng.integers(0, cfg["card"], size=(T, cfg["n_q"] + 1)).astype(np.int32)
    rows[:, 0] = rng.integers(0, cfg["text_card"], size=T)
    rows[2, 3] = -1
    gs.prefill(rows)
    for f in range(T):
        os_.step_temporal(rows[f])
    assert gs.offset == os_.offset == T
    for layer in (0, cfg["num_layers"] - 1):
        for head in (0, cfg["num_heads"] - 1):
            for slot in (0, (T // 2) % cfg["context"], (T - 1) % cfg["context"]):
                kg, vg = gs.get_kv(layer, head, slot); ko, vo = os_.get_kv(layer, head, slot)
                assert np.array_equal(kg, ko) and np.array_equal(vg, vo), f"KV row layer {layer} head {head} slot {slot}"
    toks = rng.integers(0, cfg["card"], size=cfg["n_q"] + 1).astype(np.int32)
    for f in range(3):
        t_ref, lg_ref, _ = os_.step_temporal(toks); t_gpu, lg_gpu, _ = gs.step_temporal(toks)
        assert max_rel(lg_gpu, lg_ref) < LOGIT_TOL, f"frame {f} after the prompt"
        assert_bitwise_mostly(lg_gpu, lg_ref, f"text logits frame {f} after the prompt")
        a_ref, al_ref = os_.step_depformer(t_ref); a_gpu, al_gpu = gs.step_depformer(t_ref, force=a_ref)
        assert max_rel(al_gpu, al_ref) < LOGIT_TOL
        toks = np.concatenate([[t_ref], a_ref, rng.integers(0, cfg["card"], size=cfg["n_q"] - cfg["dep_q"])]).astype(np.int32)


def test_prefill_past_the_first_lap_7b_shapes(msx, orc, gguf_for):
    """A prompt longer than the KV ring at 7B layer shapes (ring of 100 slots, 230 rows): 64-column passes on the tcgen05 GEMM that
    start before / straddle / lie beyond the end of the first lap.  The columns of a pass overwrite the oldest slots before anybody
    attends, and every column reads the saved old rows in place of the slots its successors took (T > 1 window, torch.h:170-223):
    the whole ring and the logits that follow equal 230 serial oracle steps."""
    path, cfg0 = gguf_for("moshi7b_l2", "q4_k")
    cfg = dict(cfg0, context=100)
    gm = msx.Model(path, cfg); gs = msx.Stream(gm)
    om = orc.Model(path, cfg); os_ = orc.State(om)
    rng = np.random.default_rng(5)
    T = 230
    rows = rng.integers(0, cfg["card"], size=(T, cfg["n_q"] + 1)).astype(np.int32)
    rows[:, 0] = rng.integers(0, cfg["text_card"], size=T)
    gs.prefill(rows[:150]); gs.prefill(rows[150:])
    for f in range(T):
        os_.step_temporal(rows[f])
    assert gs.offset == os_.offset == T
    for layer in range(cfg["num_layers"]):
        for head in (0, 17, cfg["num_heads"] - 1):
            for slot in range(0, 100, 7):
                kg, vg = gs.get_kv(layer, head, slot); ko, vo = os_.get_kv(layer, head, slot)
                assert np.array_equal(kg, ko) and np.array_equal(vg, vo), f"KV row layer {layer} head {head} slot {slot}"
    toks = rng.integers(0, cfg["card"], size=cfg["n_q"] + 1).astype(np.int32)
    t_ref, lg_ref, _ = os_.step_temporal(toks); t_gpu, lg_gpu, _ = gs.step_temporal(toks)
    assert max_rel(lg_gpu, lg_ref) < LOGIT_TOL
    assert_bitwise_mostly(lg_gpu, lg_ref, "text logits after a prompt of 2.3 ring lengths")


@pytest.mark.parametrize("preset,n,quant", [("tiny", 5, "q4_k"), ("tiny_pplex", 3, "q8_0"), ("moshi7b_l2", 8, "q4_k"),
                                            ("moshi7b_l2", 13, "q4_k"), ("moshi7b_l2", 32, "q4_k"), ("moshi7b_l2", 64, "q4_k")])
def test_batch_streams_vs_oracle(msx, orc, gguf_for, preset, n, quant):
    """Every stream of a lock-step batch (tensor-core dequant-GEMM, BASELINE.json config 5) against its own ORACLE state:
    different inputs per stream, teacher-forced with the oracle's tokens, logits within tolerance and (almost always) bit-identical.
    Up to 8 streams run on mma.sync (mma_gemm.cuh); 9..64 streams on the tcgen05 kind::i8 GEMM with 16 / 32 / 64 columns
    (tc_gemm.cuh; text head, depformer_in and the codebook heads included)."""
    path, cfg = gguf_for(preset, quant)
    gm = msx.Model(path, cfg); batch = msx.Batch(gm, n)
    om = orc.Model(path, cfg); states = [orc.State(om) for _ in range(n)]
    rng = np.random.default_rng(23)
    toks = rng.integers(0, cfg["card"], size=(n, cfg["n_q"] + 1)).astype(np.int32)
    toks[:, 0] = rng.integers(0, cfg["text_card"], size=n)
    for f in range(6 if preset != "moshi7b_l2" else 2):
        out = batch.step(toks)
        nxt = toks.copy()
        for s in range(n):
            t_ref, lg_ref, _ = states[s].step_temporal(toks[s]); a_ref, al_ref = states[s].step_depformer(t_ref)
            btl, bal = batch.logits(s)
            assert max_rel(btl, lg_ref) < LOGIT_TOL and max_rel(bal, al_ref) < LOGIT_TOL, f"frame {f} stream {s}: logits vs the oracle"
            assert_bitwise_mostly(btl, lg_ref, f"text logits frame {f} stream {s}")
            if out[s, 0] != t_ref:
                assert top2_margin(lg_ref) <= LOGIT_TOL * float(np.max(np.abs(lg_ref)))
                pytest.skip("a margin-excused text flip: the batch's depformer ran on another token")
            assert np.array_equal(out[s, 1:], a_ref), f"frame {f} stream {s}: audio tokens"
            nxt[s] = np.concatenate([[t_ref], a_ref, rng.integers(0, cfg["card"], size=cfg["n_q"] - cfg["dep_q"])])
        toks = nxt.astype(np.int32)

@pytest.mark.parametrize("preset,quant", [("tiny", "q4_k"), ("moshi7b_l2", "q4_k")])
def test_top_k_sampling_matches_oracle(msx, orc, gguf_for, preset, quant):
    """temp > 0: softmax(l/temp) -> top-k (25 text / 250 audio) -> arg-max of p/Exp(1) with the SAME host noise
    on both sides (reference: sampling.h:4-64, context.h:464-480) -> identical tokens, free running."""
    path, cfg = gguf_for(preset, quant)
    gm = msx.Model(path, cfg); gs = msx.Stream(gm)
    om = orc.Model(path, cfg); os_ = orc.State(om)
    kt, ka = min(25, cfg["text_card"]), min(250, cfg["card"])
    gs.set_sampling(0.7, 0.8, 25, 250); os_.set_sampling(0.7, 0.8, 25, 250)
    rng = np.random.default_rng(11)
    toks = np.array([cfg["text_card"]] + [cfg["card"]] * cfg["n_q"], dtype=np.int32)
    n_diff_from_greedy = 0
    for f in range(10 if preset == "tiny" else 3):
        nt = rng.exponential(size=kt).astype(np.float32); na = rng.exponential(size=(cfg["dep_q"], ka)).astype(np.float32)
        os_.set_noise(nt, na); gs.set_noise(nt, na)
        t_ref, lg_ref, _ = os_.step_temporal(toks)
        a_ref, al_ref = os_.step_depformer(t_ref)
        t_gpu, lg_gpu, _ = gs.step_temporal(toks)
        a_gpu, al_gpu = gs.step_depformer(t_gpu)
        assert t_gpu == t_ref, f"frame {f}: sampled text token"
        assert np.array_equal(a_gpu, a_ref), f"frame {f}: sampled audio tokens"
        assert np.array_equal(lg_gpu.view(np.uint32), lg_ref.view(np.uint32))
        n_diff_from_greedy += int(t_ref != int(np.argmax(lg_ref))) + int(np.sum(a_ref != np.argmax(al_ref, axis=1)))
        toks = np.array([t_ref] + list(a_ref) + list(rng.integers(0, cfg["card"], size=cfg["n_q"] - cfg["dep_q"])), dtype=np.int32)
    assert n_diff_from_greedy > 0, "sampling never left the greedy path: the test would not exercise top-k"
    # back to greedy: graphs are re-captured
    gs.set_sampling(0.0, 0.0); os_.set_sampling(0.0, 0.0)
    t_ref, lg_ref, _ = os_.step_temporal(toks); t_gpu, lg_gpu, _ = gs.step_temporal(toks)
    assert t_gpu == t_ref == int(np.argmax(lg_ref))


# ---------------------------------------------------------------------------------------------------
# batched streams (SURVEY.md §8e row 1 / BASELINE.json config 5): tensor-core dequant-GEMM, 8 columns
# ---------------------------------------------------------------------------------------------------
GEMM_SHAPES = [(4096, 512), (4096, 12288), (11264, 256), (1024, 3072), (2816, 1024), (1024, 2048), (4096, 6), (256, 40), (768, 520)]


@pytest.mark.parametrize("quant", ["q4_k", "q8_0"])
@pytest.mark.parametrize("k,rows", GEMM_SHAPES)
@pytest.mark.parametrize("nb,rms", [(8, False), (8, True), (3, True), (1, False)])
def test_batched_gemm_vs_oracle(msx, orc, quant, k, rows, nb, rms):
    """every column of the batched quantise + mma GEMM equals the oracle's ggml-faithful mul_mat of that column"""
    from moshi_cpp_b200 import synth
    rng = np.random.default_rng(k * 17 + rows + nb)
    GT = synth.TYPE_NAMES[quant]
    raw = synth.random_tensor(rng, GT, rows, k, 1.0 / np.sqrt(k))
    x = rng.standard_normal((nb, k)).astype(np.float32) * 1.3
    x[0, 5] = 0.0
    if nb > 1:
        x[1, :256] = 0.0                                  # an all-zero Q8_K block (d = 0)
    alpha = (1.0 + 0.1 * rng.standard_normal(k)).astype(np.float32) if rms else None
    got = msx.test_gemm_batch(GT, raw, k, x, alpha)
    for b in range(nb):
        xin = orc.rms_norm(x[b], alpha) if rms else x[b]
        ref = orc.mul_mat_vec(GT, raw, k, xin)
        assert max_rel(got[b], ref) < GEMV_TOL, f"column {b}"
        assert_bitwise_mostly(got[b], ref, f"gemm column {b}")


@pytest.mark.parametrize("preset,n,quant", [("tiny", 8, "q4_k"), ("tiny", 3, "q4_k"), ("tiny_pplex", 5, "q4_k"), ("moshi7b_l2", 8, "q4_k"),
                                            ("tiny", 8, "q8_0"), ("moshi7b_l2", 5, "q8_0"), ("moshi7b_l2", 24, "q4_k")])
def test_batch_equals_independent_streams(msx, gguf_for, preset, n, quant):
    """n streams stepped as one batch (different inputs, free running) produce the tokens and logits of n
    single msx_streams: the batched kernels share the single-stream arithmetic (double accumulation of exact
    block terms), only the summation order of the fp64 partials differs."""
    path, cfg = gguf_for(preset, quant)
    gm = msx.Model(path, cfg)
    batch = msx.Batch(gm, n)
    singles = [msx.Stream(gm) for _ in range(n)]
    rng = np.random.default_rng(9)
    frames = 10 if preset != "moshi7b_l2" else 3
    for f in range(frames):
        toks = rng.integers(0, cfg["card"], size=(n, cfg["n_q"] + 1)).astype(np.int32)
        toks[:, 0] = rng.integers(0, cfg["text_card"], size=n)
        if f == 0:
            toks[0, :] = -1                                   # zero embedding rows (lm_utils.h:172-182)
        out = batch.step(toks)
        for s in range(n):
            t, tl, _ = singles[s].step_temporal(toks[s])
            a, al = singles[s].step_depformer(t)
            btl, bal = batch.logits(s)
            assert max_rel(btl, tl) < 2e-3 and max_rel(bal, al) < 2e-3, f"frame {f} stream {s}"
            assert out[s, 0] == t and np.array_equal(out[s, 1:], a), f"frame {f} stream {s}"
            assert_bitwise_mostly(btl, tl, f"text logits frame {f} stream {s}")


def test_wide_batch_needs_tensor_core_layouts(msx, gguf_for):
    """more than 8 streams: only models whose every linear has 128-row tiles and a 256-multiple inner dimension (tiny's
    1000-row text head does not); 65 streams never"""
    path, cfg = gguf_for("tiny", "q4_k")
    gm = msx.Model(path, cfg)
    with pytest.raises(Exception):
        msx.Batch(gm, 16)
    path, cfg = gguf_for("moshi7b_l2", "q4_k")
    gm = msx.Model(path, cfg)
    with pytest.raises(Exception):
        msx.Batch(gm, 65)
    path, cfg = gguf_for("moshi7b_l2", "q8_0")
    with pytest.raises(Exception):
        msx.Batch(msx.Model(path, cfg), 16)


def test_batch_stream_restart_and_resident(msx, gguf_for):
    """streams of a batch sit at their own positions: restarting one leaves the others untouched; the resident
    replay gives the tokens of the per-call path"""
    path, cfg = gguf_for("tiny", "q4_k")
    gm = msx.Model(path, cfg)
    n = 4
    batch = msx.Batch(gm, n); ref = msx.Batch(gm, n)
    rng = np.random.default_rng(3)
    frames = rng.integers(0, cfg["card"], size=(n, 12, cfg["n_q"] + 1)).astype(np.int32)
    outs = [batch.step(frames[:, f]) for f in range(6)]
    batch.reset(2)
    assert batch.offset(2) == 0 and batch.offset(1) == 6
    # stream 2 replays its first frames, the others continue
    nxt = frames[:, 6:9].copy(); nxt[2] = frames[2, 0:3]
    outs2 = [batch.step(nxt[:, f]) for f in range(3)]
    for f in range(3):
        assert np.array_equal(outs2[f][2], outs[f][2]), "restarted stream repeats its history"
    ms, tok = ref.run_resident(frames, 9, want_tokens=True)
    for f in range(6):
        assert np.array_equal(tok[:, f], outs[f])
    for f in range(3):
        for s in (0, 1, 3):
            assert np.array_equal(tok[s, 6 + f], outs2[f][s])


def test_batch_rejects_unsupported_models_and_bad_sizes(msx, gguf_for):
    path, cfg = gguf_for("tiny_tts", "q4_k")                 # cross-attention / demux layers are single-stream only
    gm = msx.Model(path, cfg)
    with pytest.raises(msx.MsxError):
        msx.Batch(gm, 4)
    path, cfg = gguf_for("tiny", "q4_k")
    gm = msx.Model(path, cfg)
    for n in (0, 9):
        with pytest.raises(msx.MsxError):
            msx.Batch(gm, n)


# ---------------------------------------------------------------------------------------------------
# TTS family: cross-attention memory, condition_sum, demuxed text
# embeddings, low-rank depformer embeddings
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("preset,quant,tc", [("tiny_tts", "q4_k", 37), ("tiny_tts", "q8_0", 5), ("tiny_lowrank", "q4_k", 0)])
def test_tts_family_step_parity(msx, orc, gguf_for, preset, quant, tc):
    """teacher-forced like run_teacher_forced, with a conditioning memory (cond_cross [Tc][dim] -> per-layer f32 k/v,
    transformer.h:343-396), condition_sum (lm.h:575-577) and two-stream text tokens (second + 1) * (text_card + 1) + text
    as StateMachine::process emits them (lm.h:171-188)"""
    path, cfg = gguf_for(preset, quant)
    gm = msx.Model(path, cfg); gs = msx.Stream(gm)
    om = orc.Model(path, cfg); os_ = orc.State(om)
    rng = np.random.default_rng(17)
    if tc:
        cs = (0.2 * rng.standard_normal(cfg["dim"])).astype(np.float32)
        cc = rng.standard_normal((tc, cfg["dim"])).astype(np.float32)
        gs.set_condition(cs, cc); os_.set_condition(cs, cc)
    n_q, dep_q, ne = cfg["n_q"], cfg["dep_q"], cfg["text_card"] + 1
    toks = np.array([cfg["text_card"]] + [cfg["card"]] * n_q, dtype=np.int32)
    n_exact = n_cmp = 0
    for f in range(24):
        t_ref, lg_ref, to_ref = os_.step_temporal(toks)
        t_gpu, lg_gpu, to_gpu = gs.step_temporal(toks)
        assert max_rel(lg_gpu, lg_ref) < LOGIT_TOL and max_rel(to_gpu, to_ref) < LOGIT_TOL, f"frame {f}"
        n_cmp += 1; n_exact += int(np.array_equal(lg_gpu.view(np.uint32), lg_ref.view(np.uint32)))
        # the text token the state machine would hand to the depformer: sometimes with a second-stream word
        text = int(t_ref)
        if cfg["demux"] and f % 3 != 0:
            text = (int(rng.integers(0, cfg["text_card"])) + 1) * ne + text
        if f % 5 == 4:
            text = -1 if not cfg["demux"] else 0
        a_ref, al_ref = os_.step_depformer(text)
        a_gpu, al_gpu = gs.step_depformer(text, force=a_ref)
        for k in range(dep_q):
            assert max_rel(al_gpu[k], al_ref[k]) < LOGIT_TOL, f"frame {f} codebook {k}"
            n_cmp += 1; n_exact += int(np.array_equal(al_gpu[k].view(np.uint32), al_ref[k].view(np.uint32)))
        user = list(rng.integers(0, cfg["card"], size=n_q - dep_q))
        toks = np.array([text] + list(a_ref) + user, dtype=np.int32)
    assert n_exact >= 0.98 * n_cmp, f"only {n_exact}/{n_cmp} logit vectors bit-identical"


def test_condition_changes_output_and_can_be_replaced(msx, orc, gguf_for):
    path, cfg = gguf_for("tiny_tts", "q4_k")
    gm = msx.Model(path, cfg); gs = msx.Stream(gm)
    om = orc.Model(path, cfg); os_ = orc.State(om)
    rng = np.random.default_rng(2)
    toks = np.array([cfg["text_card"]] + [cfg["card"]] * cfg["n_q"], dtype=np.int32)
    _, l0, _ = gs.step_temporal(toks)
    for tc in (3, 64, 9):                                   # memory re-sized between utterances
        cc = rng.standard_normal((tc, cfg["dim"])).astype(np.float32)
        gs.reset(); os_.reset()
        gs.set_condition(None, cc); os_.set_condition(None, cc)
        _, l1, _ = gs.step_temporal(toks); _, r1, _ = os_.step_temporal(toks)
        assert max_rel(l1, r1) < LOGIT_TOL
        assert max_rel(l1, l0) > 1e-3, "the conditioning memory must influence the logits"
    # a model without cross-attention refuses a memory (reference: moshi_lm_set_voice_condition returns -1)
    path2, cfg2 = gguf_for("tiny", "q4_k")
    s2 = msx.Stream(msx.Model(path2, cfg2))
    with pytest.raises(msx.MsxError):
        s2.set_condition(None, np.zeros((2, cfg2["dim"]), dtype=np.float32))


@pytest.mark.parametrize("quant,frames", [("q4_k", 5), ("q8_0", 12)])
def test_voice_conditioners_match_oracle(msx, orc, gguf_for, quant, frames, tmp_path):
    """SURVEY.md §8f rank 1, the conditioners (voice_condition, moshi.cpp:296-366): cfg / control look-up tables and their
    projections, the projected speaker embedding + learnt padding + sinusoidal positions -> condition_sum / condition_cross
    on the GPU == the oracle's restatement; the frames that follow agree with the oracle fed the oracle's own tensors;
    a voice .safetensors (bf16 [1, C, T]) gives the same as its values passed directly."""
    from moshi_cpp_b200 import synth
    path, cfg = gguf_for("tiny_tts_voice", quant)
    gm = msx.Model(path, cfg); gs = msx.Stream(gm)
    assert msx.lib().msx_model_has_conditioners(gm.h) == 1
    om = orc.Model(path, cfg); os_ = orc.State(om)
    rng = np.random.default_rng(31)
    wavs = rng.standard_normal((synth.COND_CHANNELS, frames)).astype(np.float32)
    cs, cc = gs.set_voice(wavs)
    rs, rc = orc.voice_condition(path, cfg, wavs)
    assert cc.shape == rc.shape == (5 * frames, cfg["dim"])
    assert np.abs(cs - rs).max() <= 1e-6 and np.abs(cc - rc).max() <= 1e-6
    exact = (cs.view(np.uint32) == rs.view(np.uint32)).mean(), (cc.view(np.uint32) == rc.view(np.uint32)).mean()
    assert min(exact) > 0.999, exact
    assert np.abs(cc[frames:2 * frames] - cc[2 * frames:3 * frames]).max() > 1e-3   # padding rows differ by their positions only
    os_.set_condition(rs, rc)
    toks = np.array([cfg["text_card"]] + [cfg["card"]] * cfg["n_q"], dtype=np.int32)
    for f in range(4):
        t_ref, lg_ref, _ = os_.step_temporal(toks)
        t_gpu, lg_gpu, _ = gs.step_temporal(toks)
        assert max_rel(lg_gpu, lg_ref) < LOGIT_TOL and t_gpu == t_ref, f"frame {f}"
        a_ref, _ = os_.step_depformer(t_ref)
        a_gpu, _ = gs.step_depformer(t_ref)
        assert np.array_equal(a_ref, a_gpu)
        toks = np.array([t_ref] + list(a_ref) + [0] * (cfg["n_q"] - len(a_ref)), dtype=np.int32)
    # the voice file form
    bits = ((wavs.view(np.uint32) + 0x7FFF + ((wavs.view(np.uint32) >> 16) & 1)) >> 16).astype(np.uint16)
    vp = str(tmp_path / "voice.safetensors")
    synth.write_safetensors(vp, [("speaker_wavs", "BF16", [1, synth.COND_CHANNELS, frames], bits.tobytes())])
    g2 = msx.Stream(gm); g3 = msx.Stream(gm)
    g2.load_voice(vp)
    g3.set_voice((bits.astype(np.uint32) << 16).view(np.float32))
    toks = np.array([cfg["text_card"]] + [cfg["card"]] * cfg["n_q"], dtype=np.int32)
    _, l2, _ = g2.step_temporal(toks); _, l3, _ = g3.step_temporal(toks)
    assert np.array_equal(l2.view(np.uint32), l3.view(np.uint32))
    # error behaviour: wrong channel count; a TTS model without conditioner tensors (reference -2); no cross-attention (-1)
    with pytest.raises(msx.MsxError) as e:
        gs.set_voice(np.zeros((synth.COND_CHANNELS + 1, 3), np.float32))
    assert e.value.code == -1
    p2, c2 = gguf_for("tiny_tts", quant)
    with pytest.raises(msx.MsxError) as e:
        msx.Stream(msx.Model(p2, c2)).set_voice(wavs)
    assert e.value.code == -5
    p3, c3 = gguf_for("tiny", "q4_k")
    with pytest.raises(msx.MsxError) as e:
        msx.Stream(msx.Model(p3, c3)).load_voice(vp)
    assert e.value.code == -5
    with pytest.raises(msx.MsxError) as e:
        gs.load_voice(str(tmp_path / "missing.safetensors"))
    assert e.value.code == -2


def test_voice_embedding_prompt_matches_oracle(msx, orc, gguf_for):
    """PersonaPlex voice-embedding prompt (lm.h:694-709, 1005-1036): f32 rows fed straight into the
    temporal transformer, text forced to 3, depformer run; then normal token frames continue on the same KV state"""
    path, cfg = gguf_for("tiny_pplex", "q4_k")
    gm = msx.Model(path, cfg); gs = msx.Stream(gm)
    om = orc.Model(path, cfg); os_ = orc.State(om)
    rng = np.random.default_rng(21)
    for f in range(5):
        row = (0.5 * rng.standard_normal(cfg["dim"])).astype(np.float32)
        t_ref, lg_ref, to_ref = os_.step_temporal_embedding(row)
        t_gpu, lg_gpu, to_gpu = gs.step_temporal_embedding(row)
        assert max_rel(lg_gpu, lg_ref) < LOGIT_TOL and max_rel(to_gpu, to_ref) < LOGIT_TOL, f"prompt frame {f}"
        assert_bitwise_mostly(lg_gpu, lg_ref, f"prompt frame {f}")
        a_ref, al_ref = os_.step_depformer(3)
        a_gpu, al_gpu = gs.step_depformer(3, force=a_ref)
        assert max_rel(al_gpu, al_ref) < LOGIT_TOL
    toks = np.array([cfg["text_card"]] + [cfg["card"]] * cfg["n_q"], dtype=np.int32)
    for f in range(6):                                    # token frames attend to the prompt's KV rows
        t_ref, lg_ref, _ = os_.step_temporal(toks); t_gpu, lg_gpu, _ = gs.step_temporal(toks)
        assert max_rel(lg_gpu, lg_ref) < LOGIT_TOL and t_gpu == t_ref, f"frame {f}"
        a_ref, _ = os_.step_depformer(t_ref); a_gpu, _ = gs.step_depformer(t_ref, force=a_ref)
        assert np.array_equal(a_gpu, a_ref)
        toks = np.concatenate([[t_ref], a_ref, rng.integers(0, cfg["card"], size=cfg["n_q"] - cfg["dep_q"])]).astype(np.int32)
    assert gs.offset == 11 and os_.offset == 11


@pytest.mark.parametrize("preset,T,quant", [("tiny", 19, "q4_k"), ("tiny_pplex", 8, "q4_k"), ("moshi7b_l2", 11, "q4_k"), ("tiny_pplex", 13, "q8_0")])
def test_batched_prefill_equals_serial_prompt_steps(msx, gguf_for, preset, T, quant):
    """T fully-given prompt frames run 8 positions at a time through the tensor-core GEMM leave the
    same KV rings / position as T one-frame steps, so the conversation that follows is identical (bit for bit)."""
    path, cfg = gguf_for(preset, quant)
    gm = msx.Model(path, cfg)
    a, b = msx.Stream(gm), msx.Stream(gm)
    rng = np.random.default_rng(31)
    rows = rng.integers(0, cfg["card"], size=(T, cfg["n_q"] + 1)).astype(np.int32)
    rows[:, 0] = rng.integers(0, cfg["text_card"], size=T)
    rows[2, 3] = -1
    a.prefill(rows)
    for f in range(T):
        b.step_temporal(rows[f], want_logits=False)
    assert a.offset == b.offset == T
    for layer in (0, cfg["num_layers"] - 1):
        for head in (0, cfg["num_heads"] - 1):
            for slot in (0, T // 2, T - 1):
                ka, va = a.get_kv(layer, head, slot); kb, vb = b.get_kv(layer, head, slot)
                assert np.array_equal(ka, kb) and np.array_equal(va, vb), f"KV row layer {layer} head {head} slot {slot}"
    toks = rng.integers(0, cfg["card"], size=cfg["n_q"] + 1).astype(np.int32)
    for f in range(4):
        ta, la, _ = a.step_temporal(toks); tb, lb, _ = b.step_temporal(toks)
        assert ta == tb and np.array_equal(la.view(np.uint32), lb.view(np.uint32)), f"frame {f} after the prompt"
        xa, _ = a.step_depformer(ta); xb, _ = b.step_depformer(tb)
        assert np.array_equal(xa, xb)
        toks = np.concatenate([[ta], xa, rng.integers(0, cfg["card"], size=cfg["n_q"] - cfg["dep_q"])]).astype(np.int32)
    # a prompt that runs past the end of the ring (first lap 8 positions per pass, then one per pass) == serial steps
    more = rng.integers(0, cfg["card"], size=(cfg["context"] + 5, cfg["n_q"] + 1)).astype(np.int32)
    more[:, 0] = rng.integers(0, cfg["text_card"], size=len(more))
    a.prefill(more)
    for f in range(len(more)):
        b.step_temporal(more[f], want_logits=False)
    assert a.offset == b.offset
    for slot in (0, 3, cfg["context"] - 1):
        ka, va = a.get_kv(0, 0, slot); kb, vb = b.get_kv(0, 0, slot)
        assert np.array_equal(ka, kb) and np.array_equal(va, vb), f"KV row slot {slot} after the wrap"
    ta, la, _ = a.step_temporal(toks); tb, lb, _ = b.step_temporal(toks)
    assert ta == tb and np.array_equal(la.view(np.uint32), lb.view(np.uint32))


def test_generator_prefill_equals_provided_steps(msx, gguf_for):
    """msx_gen_prefill == the same rows through msx_gen_step with all 17 tokens provided (delay ring + model state)"""
    path, cfg = gguf_for("tiny_pplex", "q4_k")
    gm = msx.Model(path, cfg)
    sa, sb = msx.Stream(gm), msx.Stream(gm)
    ga, gb = msx.Gen(sa), msx.Gen(sb)
    rng = np.random.default_rng(8)
    rows = rng.integers(0, cfg["card"], size=(13, cfg["n_q"] + 1)).astype(np.int32)
    ga.prefill(rows)
    for f in range(13):
        gb.step(rows[f])
    assert ga.offset == gb.offset == 13
    n_user = cfg["n_q"] - 8
    for f in range(12):
        user = rng.integers(0, cfg["card"], size=n_user).astype(np.int32)
        ra, rb = ga.step(user), gb.step(user)
        assert ra[0] == rb[0] and ra[1] == rb[1] and np.array_equal(ra[2], rb[2]), f"frame {f}"


def _quant_cases(rng, k, rows):
    """rows of k values with the edge cases the quantisers branch on"""
    x = (rng.standard_normal((rows, k)) * 0.05).astype(np.float32)
    x[0, :32] = 0.0                                    # all-zero sub-block
    x[1] = 0.0                                         # all-zero row
    x[2, :64] = 0.125                                  # constant positive (max == min after the clamp at 0 is false: 0 .. 0.125)
    x[3, :32] = -0.25                                  # constant negative: hi == lo, scale 0, min 0.25
    x[4] = np.abs(x[4])                                # all positive: min clamps to 0
    x[5] *= np.repeat(np.float32(2.0) ** (np.arange(k // 32) % 12 - 6), 32)   # sub-block scales spread over 2^11
    x[6, ::2] = x[6, 1::2]                             # ties for the carrier / max
    x[7] *= 1e-6                                       # tiny magnitudes (fp16 d goes subnormal)
    x[8] *= 300.0                                      # large magnitudes
    return x


@pytest.mark.parametrize("dst,k", [("q4_k", 256), ("q4_k", 4096), ("q4_0", 128), ("q4_0", 4096), ("q8_0", 1024)])
@pytest.mark.parametrize("src", ["f32", "bf16", "f16"])
def test_quantize_rows_bit_exact(msx, orc, dst, k, src):
    """the GPU on-load quantisers produce the CPU quantiser's GGUF blocks bit for bit
    (quantize_row_q4_K incl. the make_qkx2_quants search, quantize_row_q4_0, quantize_row_q8_0); the Q4_0 / Q8_0 oracles
    are pinned against gguf-py in tests/test_oracle_pins.py."""
    rng = np.random.default_rng(101)
    x = _quant_cases(rng, k, 40)
    if src == "bf16":
        bits = ((x.view(np.uint32) + 0x7FFF + ((x.view(np.uint32) >> 16) & 1)) >> 16).astype(np.uint16)
        dev_in, st = bits, 30
        x = (bits.astype(np.uint32) << 16).view(np.float32)
    elif src == "f16":
        h = x.astype(np.float16)
        dev_in, st = h.view(np.uint16), 1
        x = h.astype(np.float32)
    else:
        dev_in, st = x, 0
    gt = {"q4_k": 12, "q4_0": 2, "q8_0": 8}[dst]
    got = msx.test_quantize_rows(gt, dev_in, src_type=st)
    fn = {"q4_k": orc.quantize_q4_K, "q4_0": orc.quantize_q4_0, "q8_0": orc.quantize_q8_0}[dst]
    ref = np.stack([fn(x[r]) for r in range(x.shape[0])])
    bad = np.argwhere(got != ref)
    assert bad.size == 0, f"{len(bad)} bytes differ, first at row {bad[0][0]} byte {bad[0][1]}"


def _expected_quantised(orc, name, gt, k, n_dims, raw, quant):
    """(type, bytes) the reference's type rules (loader.h:161-172, lm_utils.h:131-147) + the CPU quantisers give for one tensor"""
    from moshi_cpp_b200 import synth
    if quant is None or n_dims != 2 or gt not in (synth.GGML_F32, synth.GGML_F16, synth.GGML_BF16):
        return gt, raw
    f32 = ((raw.view(np.uint16).astype(np.uint32) << 16).view(np.float32) if gt == synth.GGML_BF16
           else raw.view(np.float16).astype(np.float32) if gt == synth.GGML_F16 else raw.view(np.float32))
    if quant == "q8_0":
        return synth.GGML_Q8_0, orc.quantize_q8_0(f32)
    if _TABLE_RE.search(name) or k % 256:
        return synth.GGML_Q4_0, orc.quantize_q4_0(f32)
    return synth.GGML_Q4_K, orc.quantize_q4_K(f32)


@pytest.mark.parametrize("preset,quant", [("tiny", "q4_k"), ("tiny_pplex", "q8_0"), ("tiny_tts", "q4_k"), ("tiny_stt", None)])
def test_safetensors_to_gguf(msx, orc, preset, quant, tmp_path):
    """The reference's starting point is model.safetensors with torch names (WeightLoader::from_safetensor, loader.h:77-83):
    fused in_proj_weight / out_proj.weight of all depformer steps, bf16 norm vectors of shape [1,1,d].  The converter
    must give the GGUF the reference would save (`-q <quant> -g`): "lm." names, per-step splits
    (transformer.h:764-849), F32 vectors, quantised 2-D tensors — and that file must load and step."""
    import gguf, re
    from moshi_cpp_b200 import configs, synth
    cfg = configs.get(preset)
    fp = str(tmp_path / "src.gguf"); sp = str(tmp_path / "model.safetensors"); qp = str(tmp_path / "out.gguf")
    synth.write_gguf(fp, cfg, "bf16", seed=92)
    src = {t.name: t for t in gguf.GGUFReader(fp).tensors}
    dt = {synth.GGML_F32: "F32", synth.GGML_F16: "F16", synth.GGML_BF16: "BF16"}
    fused, st, expect = {}, [], {}
    for name, t in src.items():
        raw = np.ascontiguousarray(t.data).view(np.uint8).reshape(-1)
        gt = int(t.tensor_type)
        shape = [int(v) for v in t.shape][::-1]                   # torch order
        m = re.match(r"lm\.(.*)\.(in|out)_projs\.(\d+)\.weight$", name)
        if m:
            fused.setdefault((m.group(1), m.group(2)), {})[int(m.group(3))] = (gt, shape, raw)
            expect[name] = _expected_quantised(orc, name, gt, shape[-1], 2, raw, quant) + (shape,)
        elif len(shape) == 1 and gt == synth.GGML_F32:             # norm vectors: bf16 [1,1,d] in the checkpoint
            f = raw.view(np.float32)
            bits = ((f.view(np.uint32) + 0x7FFF + ((f.view(np.uint32) >> 16) & 1)) >> 16).astype(np.uint16)
            st.append((name[3:], "BF16", [1, 1, shape[0]], bits.tobytes()))
            expect[name] = (synth.GGML_F32, (bits.astype(np.uint32) << 16).view(np.uint8), [1, 1, shape[0]])
        else:
            st.append((name[3:], dt[gt], shape, raw.tobytes()))
            expect[name] = _expected_quantised(orc, name, gt, shape[-1], len(shape), raw, quant) + (shape,)
    for (stem, kind), parts in fused.items():
        gt, shape, _ = parts[0]
        blob = b"".join(parts[i][2].tobytes() for i in range(len(parts)))
        st.append((f"{stem}.in_proj_weight" if kind == "in" else f"{stem}.out_proj.weight", dt[gt], [shape[0] * len(parts), shape[1]], blob))
    assert any(len(p) > 1 for p in fused.values()) or cfg["dep_q"] == 0      # per-step depformer weights really are fused
    st.append(("num_batches_tracked", "I64", [1], (7).to_bytes(8, "little")))   # bookkeeping tensors are left out, like save_gguf does
    synth.write_safetensors(sp, st)
    msx.safetensors_to_gguf(sp, qp, quant)
    out = {t.name: t for t in gguf.GGUFReader(qp).tensors}
    assert set(out) == set(expect)
    for name, (want_t, want, shape) in expect.items():
        t = out[name]
        assert int(t.tensor_type) == want_t, name
        assert [int(v) for v in t.shape][::-1] == shape, name
        assert np.array_equal(np.ascontiguousarray(t.data).view(np.uint8).reshape(-1), np.frombuffer(bytes(want), np.uint8)), name
    if quant is None:
        return
    gm = msx.Model(qp, cfg); gs = msx.Stream(gm)
    toks = np.array([cfg["text_card"]] + [cfg["card"]] * cfg["n_q"], dtype=np.int32)
    t0, lg, _ = gs.step_temporal(toks)
    assert np.isfinite(lg).all()


@pytest.mark.parametrize("preset,src,quant", [("tiny_lowrank", "bf16", "q4_k"), ("tiny", "f16", "q8_0"), ("tiny_stt", "bf16", None)])
def test_gguf_quantize_file(msx, orc, preset, src, quant, tmp_path):
    """`-q <quant> -g out.gguf` (moshi_lm_quantize + moshi_lm_save_gguf, moshi.cpp:654-695): the GPU file-to-file quantiser
    writes a GGUF that gguf-py parses, whose blocks equal the CPU quantisers' and whose types follow loader.h:161-172 /
    lm_utils.h:131-147; the written file then loads and steps like the on-load path."""
    import gguf
    from moshi_cpp_b200 import configs, synth
    cfg = configs.get(preset)
    fp = str(tmp_path / "src.gguf"); qp = str(tmp_path / "out.gguf")
    synth.write_gguf(fp, cfg, src, seed=91)
    msx.gguf_quantize(fp, qp, quant)
    a, b = gguf.GGUFReader(fp), gguf.GGUFReader(qp)
    assert [t.name for t in a.tensors] == [t.name for t in b.tensors]
    assert len(b.fields) <= 3                                # GGUF.version / tensor_count / kv_count only: no key/value pairs
    seen = set()
    for ta, tb in zip(a.tensors, b.tensors):
        k = int(ta.shape[0]); rows = int(ta.shape[1]) if len(ta.shape) > 1 else 1
        assert [int(v) for v in ta.shape] == [int(v) for v in tb.shape], ta.name
        raw = np.ascontiguousarray(ta.data).view(np.uint8).reshape(-1)
        want_t, want = _expected_quantised(orc, ta.name, int(ta.tensor_type), k, len(ta.shape), raw, quant)
        assert int(tb.tensor_type) == want_t, ta.name
        assert np.array_equal(np.ascontiguousarray(tb.data).view(np.uint8).reshape(-1), want), ta.name
        seen.add(want_t)
    if quant == "q4_k":
        assert {synth.GGML_Q4_K, synth.GGML_Q4_0, synth.GGML_F32} <= seen
    if quant is None:
        return
    ga = msx.Stream(msx.Model(fp, cfg, quantize=quant)); gb = msx.Stream(msx.Model(qp, cfg))
    toks = np.array([cfg["text_card"]] + [cfg["card"]] * cfg["n_q"], dtype=np.int32)
    for f in range(3):
        ta_, la, _ = ga.step_temporal(toks); tb_, lb, _ = gb.step_temporal(toks)
        assert ta_ == tb_ and np.array_equal(la.view(np.uint32), lb.view(np.uint32))
        if cfg["dep_q"] > 0:
            aa, _ = ga.step_depformer(ta_); ab, _ = gb.step_depformer(tb_)
            assert np.array_equal(aa, ab)
            toks = np.array([ta_] + list(aa) + [0] * (cfg["n_q"] - len(aa)), dtype=np.int32)


_TABLE_RE = __import__("re").compile(r"(^|\.)(text_emb|emb\.\d+|depformer_emb\.\d+|depformer_text_emb)\.weight$")


@pytest.mark.parametrize("preset,src,quant", [("tiny", "bf16", "q8_0"), ("tiny_stt", "f16", "q8_0"), ("tiny_lowrank", "bf16", "q8_0"),
                                              ("tiny", "bf16", "q4_k"), ("tiny_stt", "f16", "q4_k"), ("tiny_lowrank", "bf16", "q4_k")])
def test_quantize_on_load(msx, orc, preset, src, quant, tmp_path):
    """an unquantised GGUF loaded with quantize="q8_0" / "q4_k" (every linear and embedding table
    quantised on the GPU with the reference's type rules, loader.h:161-172, lm_utils.h:131-147) == the same weights
    quantised offline on the CPU (gguf-py's Q8_0; the oracle's quantize_row_q4_K / q4_0) and run through the oracle."""
    import gguf
    from moshi_cpp_b200 import configs, synth
    cfg = configs.get(preset)
    fp = str(tmp_path / f"{preset}-{src}.gguf")
    synth.write_gguf(fp, cfg, src, seed=77)
    rd = gguf.GGUFReader(fp)
    twin = []
    for t in rd.tensors:
        k = int(t.shape[0]); rows = int(t.shape[1]) if len(t.shape) > 1 else 1
        gt = int(t.tensor_type)
        raw = np.ascontiguousarray(t.data).view(np.uint8).reshape(-1)
        if gt in (synth.GGML_BF16, synth.GGML_F16) or (gt == synth.GGML_F32 and rows > 1):
            if gt == synth.GGML_BF16:
                f32 = (raw.view(np.uint16).astype(np.uint32) << 16).view(np.float32)
            elif gt == synth.GGML_F16:
                f32 = raw.view(np.float16).astype(np.float32)
            else:
                f32 = raw.view(np.float32)
            if quant == "q8_0":
                q = gguf.quants.quantize(f32.reshape(rows, k), gguf.GGMLQuantizationType.Q8_0)
                twin.append((t.name, synth.GGML_Q8_0, k, rows, np.ascontiguousarray(q).view(np.uint8).reshape(-1)))
            elif _TABLE_RE.search(t.name) or k % 256:
                twin.append((t.name, synth.GGML_Q4_0, k, rows, orc.quantize_q4_0(f32)))
            else:
                twin.append((t.name, synth.GGML_Q4_K, k, rows, orc.quantize_q4_K(f32)))
        else:
            twin.append((t.name, gt, k, rows, raw))
    qp = str(tmp_path / f"{preset}-{quant}-twin.gguf")
    synth.write_gguf_tensors(qp, twin)
    gm = msx.Model(fp, cfg, quantize=quant); gs = msx.Stream(gm)
    g2 = msx.Stream(msx.Model(qp, cfg))                     # the offline-quantised twin through the GPU path too
    om = orc.Model(qp, cfg); os_ = orc.State(om)
    assert gm.weight_bytes_per_frame == msx.Model(qp, cfg).weight_bytes_per_frame
    rng = np.random.default_rng(4)
    toks = np.array([cfg["text_card"]] + [cfg["card"]] * cfg["n_q"], dtype=np.int32)
    for f in range(8):
        t_ref, lg_ref, _ = os_.step_temporal(toks)
        t_gpu, lg_gpu, _ = gs.step_temporal(toks)
        t_g2, lg_g2, _ = g2.step_temporal(toks)
        assert np.array_equal(lg_gpu.view(np.uint32), lg_g2.view(np.uint32)), f"frame {f}: on-load quantisation != offline {quant} file"
        assert max_rel(lg_gpu, lg_ref) < LOGIT_TOL and t_gpu == t_ref, f"frame {f}"
        nxt = [t_ref]
        if cfg["dep_q"] > 0:
            a_ref, al_ref = os_.step_depformer(t_ref)
            a_gpu, al_gpu = gs.step_depformer(t_ref, force=a_ref)
            g2.step_depformer(t_ref, force=a_ref)
            assert max_rel(al_gpu, al_ref) < LOGIT_TOL
            nxt += list(a_ref)
        toks = np.array(nxt + list(rng.integers(0, cfg["card"], size=cfg["n_q"] + 1 - len(nxt))), dtype=np.int32)


def test_batch_sampling_equals_independent_streams(msx, gguf_for):
    """temperature > 0 in a batch: every stream samples with its own host noise exactly like a single msx_stream"""
    path, cfg = gguf_for("tiny", "q4_k")
    gm = msx.Model(path, cfg)
    n = 4
    batch = msx.Batch(gm, n); batch.set_sampling(0.7, 0.8, 25, 250)
    singles = [msx.Stream(gm) for _ in range(n)]
    for s in singles:
        s.set_sampling(0.7, 0.8, 25, 250)
    kt, ka = min(25, cfg["text_card"]), min(250, cfg["card"])
    rng = np.random.default_rng(12)
    toks = rng.integers(0, cfg["card"], size=(n, cfg["n_q"] + 1)).astype(np.int32)
    toks[:, 0] = rng.integers(0, cfg["text_card"], size=n)
    diff_from_greedy = 0
    for f in range(8):
        nt = rng.exponential(size=(n, kt)).astype(np.float32); na = rng.exponential(size=(n, cfg["dep_q"], ka)).astype(np.float32)
        batch.set_noise(nt, na)
        out = batch.step(toks)
        for i in range(n):
            singles[i].set_noise(nt[i], na[i])
            t, tl, _ = singles[i].step_temporal(toks[i])
            a, _ = singles[i].step_depformer(t)
            assert out[i, 0] == t and np.array_equal(out[i, 1:], a), f"frame {f} stream {i}"
            diff_from_greedy += int(t != int(np.argmax(tl)))
        toks = np.concatenate([out, rng.integers(0, cfg["card"], size=(n, cfg["n_q"] - cfg["dep_q"]))], axis=1).astype(np.int32)
    assert diff_from_greedy > 0, "sampling should not collapse to greedy"


def test_wide_batch_sampling_and_generator(msx, gguf_for):
    """20 conversations on the tcgen05 GEMM (32 columns, 12 of them dead) at 7B shapes: top-k sampling with per-stream host noise
    == 20 single streams; then the batched generator (delay rings, one slot replaced mid-run) == 20 (stream, generator) pairs"""
    path, cfg = gguf_for("moshi7b_l2", "q4_k")
    gm = msx.Model(path, cfg)
    n = 20
    batch = msx.Batch(gm, n, 64); batch.set_sampling(0.7, 0.8, 25, 250)
    singles = [msx.Stream(gm, context=64) for _ in range(n)]
    for s in singles:
        s.set_sampling(0.7, 0.8, 25, 250)
    kt, ka = min(25, cfg["text_card"]), min(250, cfg["card"])
    rng = np.random.default_rng(15)
    toks = rng.integers(0, cfg["card"], size=(n, cfg["n_q"] + 1)).astype(np.int32)
    toks[:, 0] = rng.integers(0, cfg["text_card"], size=n)
    for f in range(3):
        nt = rng.exponential(size=(n, kt)).astype(np.float32); na = rng.exponential(size=(n, cfg["dep_q"], ka)).astype(np.float32)
        batch.set_noise(nt, na)
        out = batch.step(toks)
        for i in range(n):
            singles[i].set_noise(nt[i], na[i])
            t, _, _ = singles[i].step_temporal(toks[i])
            a, _ = singles[i].step_depformer(t)
            assert out[i, 0] == t and np.array_equal(out[i, 1:], a), f"sampled frame {f} stream {i}"
        toks = np.concatenate([out, rng.integers(0, cfg["card"], size=(n, cfg["n_q"] - cfg["dep_q"]))], axis=1).astype(np.int32)
    batch.close()
    for s in singles:
        s.close()
    bg = msx.BatchGen(msx.Batch(gm, n, 64))
    streams = [msx.Stream(gm, context=64) for _ in range(n)]
    gens = [msx.Gen(s) for s in streams]
    n_user = cfg["n_q"] - cfg["dep_q"]
    emitted = 0
    for f in range(8):
        if f == 4:
            bg.reset(7)
            streams[7].reset(); gens[7] = msx.Gen(streams[7])
        user = rng.integers(0, cfg["card"], size=(n, n_user)).astype(np.int32)
        valid, text, audio = bg.step(user)
        for i in range(n):
            rc, t, a = gens[i].step(user[i])
            assert valid[i] == rc, f"frame {f} stream {i}"
            if rc:
                emitted += 1
                assert text[i] == t and np.array_equal(audio[i], a), f"frame {f} stream {i}"
    assert emitted > n * 4 and bg.offset(7) == 4 and bg.offset(0) == 8


@pytest.mark.parametrize("preset", ["tiny", "tiny_pplex"])
def test_batched_generator_equals_independent_generators(msx, gguf_for, preset):
    """config 5 end to end: n conversations with their own delay rings on one batch == n (msx_stream + msx_gen) pairs,
    including a conversation that is replaced mid-run"""
    path, cfg = gguf_for(preset, "q4_k")
    gm = msx.Model(path, cfg)
    n = 5
    bg = msx.BatchGen(msx.Batch(gm, n))
    streams = [msx.Stream(gm) for _ in range(n)]
    gens = [msx.Gen(s) for s in streams]
    rng = np.random.default_rng(6)
    n_user = cfg["n_q"] - (8 if cfg["model_type"] == "personaplex" else cfg["dep_q"])
    emitted = 0
    for f in range(30):
        if f == 11:                                   # slot 3 starts a new conversation
            bg.reset(3)
            streams[3].reset(); gens[3] = msx.Gen(streams[3])
        user = rng.integers(0, cfg["card"], size=(n, n_user)).astype(np.int32)
        valid, text, audio = bg.step(user)
        for i in range(n):
            rc, t, a = gens[i].step(user[i])
            assert valid[i] == rc, f"frame {f} stream {i}"
            if rc:
                emitted += 1
                assert text[i] == t and np.array_equal(audio[i], a), f"frame {f} stream {i}"
    assert emitted > n * 20 and bg.offset(3) == 19 and bg.offset(0) == 30
