#!/usr/bin/env python
"""bench.py — frames/s of the per-frame LM decode step (temporal transformer + depformer).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--preset moshi7b] [--quant q4_k]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one frame of one conversation stream (17 input tokens -> text token + dep_q audio tokens).
N > 1: one process per GPU, each with its own replica of the weights and its own independent stream
(the path shards by conversation stream; no data-path collective) -> weak scaling, value = N*K / max-over-ranks time.

value   device-resident replay: token frames already in HBM, K fused frames back to back, CUDA events on the
        launching stream (msx_run_resident).
e2e     the same metric through the public per-frame API (msx_gen_step = moshi_lm_send2 + moshi_lm_receive):
        host tokens in, H2D + launch + D2H + host sync every frame, timed with CUDA events on the same stream.
roofline  the dominant kernel family (gated-MLP linear_in dequant-GEMV) timed per launch inside a real frame
        (msx_profile_frame: CUDA event after every launch on the launching stream).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import _pkgload  # noqa: E402

_pkgload.load()
from moshi_cpp_b200 import configs, synth  # noqa: E402

SEED_MODEL, SEED_TOKENS = 1234, 42
FRAME_RATE = 12.5


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """SM clock / power / throttle reasons sampled every few milliseconds WHILE the timed regions run (NVML in a thread; the
    timed regions of a default run last tens of milliseconds, far below nvidia-smi's 200 ms polling period)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int, period_s: float = 0.004):
        self.gpu_index, self.period, self.samples, self.stop_flag, self.t, self.err = gpu_index, period_s, [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES remaps CUDA ordinals; NVML sees physical indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if gpu_index < len(ids) and ids[gpu_index].isdigit():
                    phys = int(ids[gpu_index])
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
        except Exception as e:                       # noqa: BLE001
            self.nv, self.h, self.err = None, None, f"NVML unavailable: {e}"

    def _loop(self):
        nv, h = self.nv, self.h
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((sm, pw, int(rs)))
            except Exception as e:                   # noqa: BLE001
                self.err = str(e); return
            time.sleep(self.period)

    def start(self):
        if self.nv is None:
            return
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def stop(self) -> dict:
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "NVML unavailable"]}
        self.stop_flag = True
        self.t.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "no samples"]}
        sm = [x[0] for x in self.samples]
        try:
            mx = float(self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM))
        except Exception:                            # noqa: BLE001
            mx = float(max(sm))
        bits = 0
        for x in self.samples:
            bits |= x[2]
        busy = [v for v in sm if v > 0.5 * max(sm)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": mx, "power_w_max": float(max(x[1] for x in self.samples)),
                "samples": len(sm), "reasons": sorted(n for b, n in self.REASONS.items() if bits & b), "how": "NVML, 4 ms period, timed regions only"}


def stored_bytes(quant: str, K: int, rows: int) -> int:
    """bytes one launch of the GEMV reads from the repacked planes: q4_k stores 148 B per 256 weights"""
    return rows * (K // 256) * 148 if quant == "q4_k" else rows * (K // 32) * 34


def committed_traffic(kernel: str, quant: str, K: int, rows: int):
    """DRAM bytes per launch measured by ncu for this kernel / shape (profiles/kernel_traffic.json, written from a --set full
    capture), or None when no capture matches the layout this build stores — a stale number is worse than none"""
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_traffic.json")) as f:
            for e in json.load(f)["captures"]:
                if e["kernel"] == kernel and e["quant"] == quant and e["K"] == K and e["rows"] == rows and \
                        e["stored_bytes_per_launch"] == stored_bytes(quant, K, rows):
                    return e["dram_bytes_read"] + e["dram_bytes_write"]
    except (OSError, KeyError, ValueError):
        pass
    return None


def make_frames(cfg, n=64):
    rng = np.random.default_rng(SEED_TOKENS)
    fr = rng.integers(0, cfg["card"], size=(n, cfg["n_q"] + 1)).astype(np.int32)
    fr[:, 0] = rng.integers(0, cfg["text_card"], size=n)
    return fr


def ensure_gguf(preset, quant, rank, world, barrier):
    path = os.path.join(os.environ.get("MSX_CACHE", "/tmp/msx_cache"), f"{preset}-{quant}-s{SEED_MODEL}.gguf")
    if rank == 0 and not os.path.exists(path):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        synth.write_gguf(path, configs.get(preset), quant, SEED_MODEL)
    barrier()
    return path


def user_codes(cfg, row):
    """the codes the caller supplies per frame: the user stream (n_q - dep_q codebooks), all n_q for STT, none for TTS"""
    n_user = cfg["n_q"] - cfg["dep_q"] if cfg["dep_q"] > 0 else cfg["n_q"]
    return row[len(row) - n_user:]


def tts_condition(cfg, tc=125):
    """synthetic conditioning of a cross-attention (TTS) model: condition_sum [dim] and a memory of tc = 5 x 25 rows, the
    shape voice_condition (moshi.cpp:296-366) produces for a 25-frame speaker embedding"""
    rng = np.random.default_rng(SEED_TOKENS + 1)
    return (0.2 * rng.standard_normal(cfg["dim"])).astype(np.float32), rng.standard_normal((tc, cfg["dim"])).astype(np.float32)


def host_threads() -> int:
    """cores this process may run on (torchrun pins nothing but exports OMP_NUM_THREADS=1, which must not decide this)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_fps(path, cfg, frames, max_frames, budget_s):
    """The oracle (CPU restatement of the reference's ggml-CPU path) on the host cores, greedy LMGen.
    -> (frames/s, frames timed, seconds, OpenMP threads actually used, emitted tokens per frame [(ok, text, audio...)])"""
    import oracle
    threads = oracle.set_threads(host_threads())       # explicit: OMP_NUM_THREADS=1 under torchrun would otherwise win
    om = oracle.Model(path, cfg)
    trace = []
    if cfg.get("cross_attention"):                      # TTS: the same two graphs per frame on a conditioned state
        st = oracle.State(om)
        st.set_condition(*tts_condition(cfg))
        toks = np.array([cfg["text_card"]] + [cfg["card"]] * cfg["n_q"], dtype=np.int32)

        def step(_):
            nonlocal toks
            t, _, _ = st.step_temporal(toks)
            a, _ = st.step_depformer(t)
            trace.append((1, int(t)) + tuple(int(v) for v in a))
            toks = np.array([t] + list(a) + [0] * (cfg["n_q"] - len(a)), dtype=np.int32)
    else:
        og = oracle.LMGen(om)

        def step(i):
            ok, t, a = og.step(user_codes(cfg, frames[i % len(frames)]))
            trace.append((int(ok), int(t)) + tuple(int(v) for v in a))
    step(0)                                            # warm-up frame (page-in of the mmapped weights)
    t0 = time.perf_counter(); n = 0
    while n < max_frames:
        step(n + 1); n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return n / dt, n, dt, threads, trace


def gpu_token_trace(msx, stream, cfg, frames, n):
    """the first n frames of the SAME protocol as cpu_reference_fps through the product's public per-frame API"""
    stream.reset()
    trace = []
    if cfg.get("cross_attention"):
        toks = np.array([cfg["text_card"]] + [cfg["card"]] * cfg["n_q"], dtype=np.int32)
        for _ in range(n):
            t, _, _ = stream.step_temporal(toks, want_logits=False)
            a, _ = stream.step_depformer(t, want_logits=False)
            trace.append((1, int(t)) + tuple(int(v) for v in a))
            toks = np.array([t] + list(a) + [0] * (cfg["n_q"] - len(a)), dtype=np.int32)
    else:
        gen = msx.Gen(stream)
        for i in range(n):
            ok, t, a = gen.step(user_codes(cfg, frames[i % len(frames)]))
            trace.append((int(ok), int(t)) + tuple(int(v) for v in a))
        gen.close()
    return trace


_JSON_FD = None


def emit(line: dict) -> None:
    """the ONE JSON line of the contract, on the process's real stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    # libraries print to stdout behind our back (NCCL's "NCCL version ..." banner under torchrun): keep the real stdout
    # for the JSON line and send everything else to stderr
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preset", default="moshi7b")
    ap.add_argument("--quant", default="q4_k")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the short single-stream runs of BASELINE.json's other model configs (extra key other_configs, N = 1 only)")
    ap.add_argument("--streams", type=int, default=8,
                    help="also time a lock-step batch of this many streams per GPU (BASELINE config 5; q4_k only, 0 = skip)")
    ap.add_argument("--tp", action="store_true",
                    help="tensor-parallel: the N ranks serve ONE stream (BASELINE config 4; strong scaling, fused peer-memory all-reduce)")
    ap.add_argument("--fill", type=int, default=0, help="--tp: frames replayed before timing (e.g. 3000 = full KV ring)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = configs.get(args.preset)
    kind = {"tts": "text-to-speech decode step (temporal transformer with cross-attention + depformer)",
            "stt": "speech-to-text step (temporal transformer)"}.get(cfg.get("model_type"), "speech-to-speech step (temporal transformer + depformer)")
    workload = f"{args.preset} {args.quant} {kind}, single stream per GPU"
    base_cfg = {"workload": workload, "streams_per_gpu": 1, "sharding": "independent conversation streams (replicas), no data-path collective",
                "l2": "inputs larger than L2 (each frame streams the full weight set, 4.1 GB >> 126 MB)",
                "model_seed": SEED_MODEL, "token_seed": SEED_TOKENS, "context": cfg["context"]}
    frames = make_frames(cfg)

    # ------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        path = ensure_gguf(args.preset, args.quant, 0, 1, lambda: None)
        budget = 150.0
        fps, n, dt, ncores, _ = cpu_reference_fps(path, cfg, frames, max(1, args.steps), budget)
        sample = (f"{n} of the requested {args.steps} frames (capped at {budget:.0f} s of CPU time), greedy LMGen, same GGUF and token seed; "
                  "CPU restatement of the reference's ggml-CPU path (oracle/ggml_ref.c, OpenMP) — ggml itself is not buildable here")
        line = {"impl": "reference", "metric": "frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
                "steps_measured": n, "warmup": 1, "ms_per_step": 1000.0 / fps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int8 dot / f32", "data": "synthetic", "config": base_cfg,
                "realtime_factor": fps / FRAME_RATE,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": ncores, "kind": "port", "sample": sample},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    # ------------------------------------------------------------------------------------------ our arm
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    from moshi_cpp_b200 import binding as msx
    path = ensure_gguf(args.preset, args.quant, rank, world, barrier)
    if args.tp and world > 1:
        # ---- config 4: one stream sharded over all ranks ------------------------------------------------
        ids = [msx.tp_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        model = msx.Model(path, cfg, device=local_rank, tp_rank=rank, tp_world=world)
        stream = msx.Stream(model, nccl_id=ids[0])
        hs = [None] * world
        dist.all_gather_object(hs, stream.tp_export())
        stream.tp_connect(hs)
        K, W = args.steps, max(3, args.warmup)
        if args.fill:
            stream.run_resident(frames, args.fill)
        stream.run_resident(frames, W)
        sampler = ClockSampler(local_rank); barrier(); torch.cuda.synchronize(local_rank); sampler.start()
        ms_res, _ = stream.run_resident(frames, K)
        torch.cuda.synchronize(local_rank); barrier()
        clocks = sampler.stop()
        t = torch.tensor([ms_res], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_res = float(t[0])
        if rank == 0:
            fps = K / (ms_res * 1e-3)
            cfg4 = dict(base_cfg, workload=f"{args.preset} {args.quant} single stream, tensor-parallel over {world} GPUs (heads / hidden shards, "
                        "fp64 partial sums, fused GEMV -> peer-memory all-reduce)", sharding=f"tensor parallel x{world}", kv_slots_at_timing=min(stream.offset, cfg["context"]))
            emit({"metric": "frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
                              "ms_per_step": ms_res / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                              "dtype": "int8 dot / f32, f64 partial sums", "data": "synthetic", "config": cfg4,
                              "realtime_factor": fps / FRAME_RATE, "gpu_launches": stream.launches_per_frame * K,
                  "launches_per_frame": stream.launches_per_frame, "clocks": clocks})
        dist.destroy_process_group()
        return 0
    t_load = time.perf_counter()
    model = msx.Model(path, cfg, device=local_rank)
    stream = msx.Stream(model)
    t_load = time.perf_counter() - t_load
    K, W = args.steps, max(3, args.warmup)
    if cfg.get("cross_attention"):
        stream.set_condition(*tts_condition(cfg))

    sampler = ClockSampler(local_rank)
    # ---- value: device-resident replay ---------------------------------------------------------
    stream.run_resident(frames, W)
    kv0 = stream.kv_bytes_next
    barrier(); torch.cuda.synchronize(local_rank)
    sampler.start()
    ms_res, _ = stream.run_resident(frames, K)
    torch.cuda.synchronize(local_rank); barrier()
    kv1 = stream.kv_bytes_next
    # ---- e2e: public per-frame API, host tokens in / out every frame ------------------------------
    stream.reset()
    gen = msx.Gen(stream)
    for i in range(W):
        gen.step(user_codes(cfg, frames[i % len(frames)]))
    barrier(); torch.cuda.synchronize(local_rank)
    stream.timer_start()
    t0 = time.perf_counter()
    for i in range(K):
        gen.step(user_codes(cfg, frames[(W + i) % len(frames)]))
    ms_e2e = stream.timer_stop()
    wall_e2e = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize(local_rank); barrier()
    clocks = sampler.stop()

    if dist is not None:
        t = torch.tensor([ms_res, ms_e2e], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_res, ms_e2e = float(t[0]), float(t[1])

    # ---- roofline of the dominant kernel, timed per launch inside real frames -----------------------
    fam_ms, fam_n = {}, {}
    n_prof = 8
    for i in range(n_prof):
        _, fam = stream.profile_frame(frames[i % len(frames)])
        if i < 2:
            continue                                   # first eager frames: module / clock warm-up
        for k, (ms, n) in fam.items():
            fam_ms[k] = fam_ms.get(k, 0.0) + ms; fam_n[k] = fam_n.get(k, 0) + n
    tot_ms = sum(fam_ms.values())
    shares = {k: round(v / tot_ms, 4) for k, v in sorted(fam_ms.items(), key=lambda kv: -kv[1])}
    dom = "linear_in"
    row_b = synth.row_bytes(synth.TYPE_NAMES[args.quant], cfg["dim"])
    dom_bytes = row_b * 2 * cfg["hidden"]              # GGUF bytes of one gating.linear_in matrix = algorithmic bytes per launch
    dom_ms = fam_ms[dom] / fam_n[dom]
    peaks, peak_src = load_peaks()
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    # same kernel, same shape, replayed from a CUDA graph over rotating copies of the matrix (> L2), i.e. with the
    # PDL overlap the real step has (the in-situ number above is taken with an event after every launch, which
    # serialises the launches and hides that overlap)
    import ctypes as C
    L = msx.lib()
    L.msx_bench_gemv.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
    gt = synth.TYPE_NAMES[args.quant]
    raw = synth.random_tensor(np.random.default_rng(7), gt, 2 * cfg["hidden"], cfg["dim"], 1.0 / np.sqrt(cfg["dim"]))
    us = C.c_float(0)
    replay_gbs = None
    if L.msx_bench_gemv(local_rank, gt, raw.ctypes.data, cfg["dim"], 2 * cfg["hidden"], 8, 200, 1, 2, C.byref(us)) == 0:
        replay_gbs = dom_bytes / (us.value * 1e-6) / 1e9
    roofline = {"bound": "hbm", "kernel": "dq_matvec_kernel<Q4_K,32> gating.linear_in (rms_norm + q8_K quant + dequant-GEMV + silu gate)",
                "achieved": achieved, "peak": peaks["hbm_gbs"], "peak_source": f"{peak_src} copy bandwidth (MEASURED_PEAKS.json)",
                "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel and shape, from the committed ncu --set full capture
                # (profiles/kernel_traffic.json); used only while the capture's stored-layout bytes equal this build's
                "traffic": committed_traffic("dq_matvec_kernel linear_in", args.quant, cfg["dim"], 2 * cfg["hidden"]),
                "bytes_per_launch": dom_bytes, "launch_us": dom_ms * 1e3, "launches_timed": fam_n[dom],
                "graph_replay": {"launch_us": us.value, "achieved": replay_gbs, "frac": (replay_gbs or 0) / peaks["hbm_gbs"],
                                 "how": "200 launches of the same kernel/shape in one CUDA graph over 8 rotating matrices (415 MB > L2)"},
                "family_time_share": shares}

    # ---- the two stacks inside the real pipelined run (graphs + PDL, one event between the two graphs of a frame) --------
    stacks = None
    if cfg["dep_q"] > 0 and not (cfg.get("cross_attention") or cfg.get("demux")):
        Ks = min(K, 200)
        kvs0 = stream.kv_bytes_next
        t_ms, d_ms = stream.run_resident_split(frames, Ks)
        kvs1 = stream.kv_bytes_next
        lin_b = lambda rows, k: rows * synth.row_bytes(synth.linear_type(synth.TYPE_NAMES[args.quant], k), k)
        d_, h_ = cfg["dim"], cfg["hidden"]
        temporal_w = cfg["num_layers"] * (lin_b(3 * d_, d_) + lin_b(d_, d_) + lin_b(2 * h_, d_) + lin_b(d_, h_)) + lin_b(cfg["text_card"], d_)
        dep_w = model.weight_bytes_per_frame - temporal_w
        t_gbs = (temporal_w + 0.5 * (kvs0 + kvs1)) / (t_ms / Ks * 1e-3) / 1e9
        d_gbs = dep_w / (d_ms / Ks * 1e-3) / 1e9
        stacks = {"temporal_ms": t_ms / Ks, "depformer_ms": d_ms / Ks, "temporal_bytes": temporal_w + 0.5 * (kvs0 + kvs1), "depformer_bytes": dep_w,
                  "temporal_gbs": t_gbs, "temporal_frac_of_peak": t_gbs / load_peaks()[0]["hbm_gbs"], "depformer_gbs": d_gbs,
                  "depformer_frac_of_peak": d_gbs / load_peaks()[0]["hbm_gbs"], "steps": Ks,
                  "how": "msx_run_resident_split: the timed graphs replayed with one CUDA event between the temporal and the depformer graph of every frame"}

    fps = world * K / (ms_res * 1e-3)
    fps_e2e = world * K / (ms_e2e * 1e-3)
    w_bytes = model.weight_bytes_per_frame
    kv_avg = 0.5 * (kv0 + kv1)
    step_gbs = (w_bytes + kv_avg) / (ms_res / K * 1e-3) / 1e9

    # ---- batched streams (config 5): n conversations per GPU, weights read once per frame for all -------------
    batched = None
    if args.streams > 1 and args.quant in ("q4_k", "q8_0") and not (cfg.get("cross_attention") or cfg.get("demux") or cfg.get("dep_low_rank")):
        nb = min(8, args.streams)
        t_b = time.perf_counter()
        batch = msx.Batch(model, nb)
        t_b = time.perf_counter() - t_b
        rngb = np.random.default_rng(43)
        bframes = rngb.integers(0, cfg["card"], size=(nb, len(frames), cfg["n_q"] + 1)).astype(np.int32)
        bframes[:, :, 0] = rngb.integers(0, cfg["text_card"], size=(nb, len(frames)))
        Kb = min(K, 200)
        batch.run_resident(bframes, W)
        bkv0 = batch.kv_bytes_next()
        barrier(); torch.cuda.synchronize(local_rank)
        ms_b, _ = batch.run_resident(bframes, Kb)
        torch.cuda.synchronize(local_rank); barrier()
        bkv1 = batch.kv_bytes_next()
        # end to end through the batched generator (per-conversation delay rings, msx_bgen_step): host user codes in,
        # delayed tokens out, n x 336 B H2D + n x 176 B D2H and a host sync every frame
        batch.reset()
        bgen = msx.BatchGen(batch)
        n_user_b = cfg["n_q"] - cfg["dep_q"] if cfg["dep_q"] > 0 else cfg["n_q"]
        for i in range(W):
            bgen.step(bframes[:, i % bframes.shape[1], -n_user_b:])
        barrier(); torch.cuda.synchronize(local_rank)
        t0 = time.perf_counter()
        for i in range(Kb):
            bgen.step(bframes[:, (W + i) % bframes.shape[1], -n_user_b:])
        ms_be = (time.perf_counter() - t0) * 1e3
        bgen.close()
        torch.cuda.synchronize(local_rank); barrier()
        if dist is not None:
            t = torch.tensor([ms_b, ms_be], device=f"cuda:{local_rank}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_b, ms_be = float(t[0]), float(t[1])
        gemm_us = msx.bench_gemm_batch(raw, cfg["dim"], nb, 8, 200, 2, False, device=local_rank) if args.quant == "q4_k" else None   # (micro-benchmark hook is q4_k only)
        b_bytes = w_bytes + 0.5 * (bkv0 + bkv1)
        batched = {
            "workload": f"{args.preset} {args.quant}, {nb} independent streams per GPU stepped as one batch (BASELINE.json config 5)",
            "streams_per_gpu": nb, "value": world * nb * Kb / (ms_b * 1e-3), "unit": "frames/s (all streams, all GPUs)",
            "ms_per_step": ms_b / Kb, "per_stream_fps": Kb / (ms_b * 1e-3), "per_stream_realtime_factor": Kb / (ms_b * 1e-3) / FRAME_RATE,
            "e2e": {"value": world * nb * Kb / (ms_be * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": 336 * nb, "d2h_bytes_per_step": 176 * nb,
                    "timing": "wall clock around msx_bgen_step (per-stream delay rings on the host, one msx_batch_step, host sync inside every call)"},
            "step_bytes": {"weights_read_once": w_bytes, "kv_avg": 0.5 * (bkv0 + bkv1), "gbs": b_bytes / (ms_b / Kb * 1e-3) / 1e9},
            "launches_per_frame": batch.launches_per_frame, "steps": Kb,
            "gemm": {"kernel": "dq_matmul_q4k_kernel gating.linear_in (mma.sync m16n8k32 u8 x s8, TMA unit ring, silu gate)",
                     "launch_us": gemm_us, "achieved": dom_bytes / (gemm_us * 1e-6) / 1e9 if gemm_us else None,
                     "frac": dom_bytes / (gemm_us * 1e-6) / 1e9 / peaks["hbm_gbs"] if gemm_us else None,
                     "how": "200 launches in one CUDA graph over 8 rotating matrices, activations pre-quantised"},
            "setup_s": t_b,
        }
        batch.close()
        # ---- wider batches on the tcgen05 kind::i8 GEMM (tc_gemm.cuh): 16 / 32 / 64 streams per GPU, full-length KV rings ----
        wide = []
        if args.quant == "q4_k":
            for nw in (16, 32, 64):
                try:
                    wb = msx.Batch(model, nw)
                except Exception as e:                   # noqa: BLE001  (a model without tensor-core layouts, or no room for the rings)
                    wide.append({"streams_per_gpu": nw, "error": str(e)[:160]})
                    break
                wfr = rngb.integers(0, cfg["card"], size=(nw, 64, cfg["n_q"] + 1)).astype(np.int32)
                wfr[:, :, 0] = rngb.integers(0, cfg["text_card"], size=(nw, 64))
                Kw = min(K, 60)
                wb.run_resident(wfr, 5)
                wkv0 = wb.kv_bytes_next()
                barrier(); torch.cuda.synchronize(local_rank)
                ms_w, _ = wb.run_resident(wfr, Kw)
                torch.cuda.synchronize(local_rank); barrier()
                wkv1 = wb.kv_bytes_next()
                # end to end: host user codes in, delayed tokens out through the batched generator, host sync every frame
                wb.reset()
                wgen = msx.BatchGen(wb)
                for i in range(5):
                    wgen.step(wfr[:, i % wfr.shape[1], -n_user_b:])
                barrier(); torch.cuda.synchronize(local_rank)
                t0 = time.perf_counter()
                for i in range(Kw):
                    wgen.step(wfr[:, (5 + i) % wfr.shape[1], -n_user_b:])
                ms_we = (time.perf_counter() - t0) * 1e3
                wgen.close()
                torch.cuda.synchronize(local_rank); barrier()
                if dist is not None:
                    t = torch.tensor([ms_w, ms_we], device=f"cuda:{local_rank}", dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms_w, ms_we = float(t[0]), float(t[1])
                wide.append({"streams_per_gpu": nw, "value": world * nw * Kw / (ms_w * 1e-3), "unit": "frames/s (all streams, all GPUs)",
                             "ms_per_step": ms_w / Kw, "per_stream_fps": Kw / (ms_w * 1e-3), "per_stream_realtime_factor": Kw / (ms_w * 1e-3) / FRAME_RATE,
                             "launches_per_frame": wb.launches_per_frame, "steps": Kw, "kv_avg": 0.5 * (wkv0 + wkv1),
                             "e2e": {"value": world * nw * Kw / (ms_we * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": 336 * nw,
                                     "d2h_bytes_per_step": 176 * nw, "timing": "wall clock around msx_bgen_step"}})
                wb.close()
        # ---- several batches at once: each msx_batch owns a CUDA stream, so the latency-bound launch chains of independent batches
        #      fill each other's gaps when they are stepped from their own host threads ----
        concurrent = []
        if args.quant == "q4_k" and wide and "error" not in wide[-1]:
            import threading
            for n_b, n_s, ctx_c in [(2, 8, 0), (2, 64, 1024)]:
                try:
                    cbs = [msx.Batch(model, n_s, ctx_c) for _ in range(n_b)]
                except Exception as e:                   # noqa: BLE001
                    concurrent.append({"batches": n_b, "streams_per_batch": n_s, "error": str(e)[:160]})
                    continue
                cfr = [rngb.integers(0, cfg["card"], size=(n_s, 64, cfg["n_q"] + 1)).astype(np.int32) for _ in range(n_b)]
                for f in cfr:
                    f[:, :, 0] = rngb.integers(0, cfg["text_card"], size=(n_s, 64))
                for cb, f in zip(cbs, cfr):
                    cb.run_resident(f, 5)
                Kc = min(K, 60)
                dev_ms = [0.0] * n_b

                def _work(i):
                    dev_ms[i], _ = cbs[i].run_resident(cfr[i], Kc)
                ths = [threading.Thread(target=_work, args=(i,)) for i in range(n_b)]
                barrier(); torch.cuda.synchronize(local_rank)
                t0 = time.perf_counter()
                for t_ in ths:
                    t_.start()
                for t_ in ths:
                    t_.join()
                wall_c = (time.perf_counter() - t0) * 1e3
                torch.cuda.synchronize(local_rank); barrier()
                if dist is not None:
                    t = torch.tensor([wall_c], device=f"cuda:{local_rank}", dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    wall_c = float(t[0])
                concurrent.append({"batches": n_b, "streams_per_batch": n_s, "streams_per_gpu": n_b * n_s, "context": ctx_c or cfg["context"],
                                   "value": world * n_b * n_s * Kc / (wall_c * 1e-3), "unit": "frames/s (all streams, all GPUs)",
                                   "ms_per_step_all_batches": wall_c / Kc, "device_ms_per_step_slowest_batch": max(dev_ms) / Kc, "steps": Kc,
                                   "timing": "wall clock around the host threads (one msx_batch_run_resident each); device time of the slowest batch beside it"})
                for cb in cbs:
                    cb.close()
        batched["concurrent_batches"] = concurrent
        batched["wide"] = {"kernel": "tc_matmul_q4k_kernel<16|32|64> (tcgen05.mma kind::i8, accumulators in tensor memory, exact Q4_K x Q8_K)",
                           "what": "the same lock-step batch with more conversations per GPU; device-timed resident replay", "runs": wide}

    # ---- persistent step kernel (opt-in path), for the record --------------------------------------------------
    step_kernel = None
    try:
        sk_stream = msx.Stream(model, step_kernel=True)
        if sk_stream.launches_per_frame <= 2:
            sk_stream.run_resident(frames, W)
            barrier(); torch.cuda.synchronize(local_rank)
            ms_sk, _ = sk_stream.run_resident(frames, K)
            step_kernel = {"ms_per_step": ms_sk / K, "frames_per_s": K / (ms_sk * 1e-3), "launches_per_frame": sk_stream.launches_per_frame,
                           "what": "MSX_STREAM_STEP_KERNEL: one persistent cooperative kernel per stack (TMA weight ring, flag-in-data "
                                   "activation exchange), same frames, this rank only"}
        sk_stream.close()
    except Exception as e:                           # noqa: BLE001
        step_kernel = {"error": str(e)[:200]}

    # ---- config 4 when the driver runs N > 1: the N ranks also serve ONE tensor-parallel stream -----------------
    tensor_parallel = None
    if dist is not None and not cfg.get("cross_attention") and cfg["num_heads"] % world == 0:
        try:
            ids = [msx.tp_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            tpm = msx.Model(path, cfg, device=local_rank, tp_rank=rank, tp_world=world)
            tps = msx.Stream(tpm, nccl_id=ids[0])
            hs = [None] * world
            dist.all_gather_object(hs, tps.tp_export())
            tps.tp_connect(hs)
            tps.run_resident(frames, W)
            barrier(); torch.cuda.synchronize(local_rank)
            ms_tp, tok_tp = tps.run_resident(frames, min(K, 200), want_tokens=True)
            torch.cuda.synchronize(local_rank); barrier()
            t = torch.tensor([ms_tp], device=f"cuda:{local_rank}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ref = msx.Stream(model)
            ref.run_resident(frames, W)
            ms_one, tok_one = ref.run_resident(frames, min(K, 200), want_tokens=True)
            ref.close()
            same = torch.tensor([int(np.array_equal(tok_tp, tok_one))], device=f"cuda:{local_rank}")
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            tensor_parallel = {"world": world, "ms_per_step": float(t[0]) / min(K, 200), "one_gpu_ms": ms_one / min(K, 200),
                               "speedup_vs_one_gpu": (ms_one / min(K, 200)) / (float(t[0]) / min(K, 200)),
                               "allreduce": f"{2 * cfg['num_layers']} fused GEMV -> peer-memory all-reduces per frame (f64 partial sums over NVLink)",
                               "tokens_identical_to_one_gpu": bool(int(same[0])), "steps": min(K, 200),
                               "workload": f"{args.preset} {args.quant} ONE stream, tensor-parallel x{world} (BASELINE.json config 4 mode)"}
            tps.close(); tpm.close()
        except Exception as e:                       # noqa: BLE001
            tensor_parallel = {"error": str(e)[:300]}
            barrier()

    # ---- BASELINE.json's other single-GPU configurations, for the record (N = 1, default workload only) ----------------
    # config 0's model (stt-1b, q8_0), config 1 (tts-1.6b q4_k decode against a 125-row conditioning memory) and the headline model
    # with its KV ring full (config 4's operating point on one GPU): 200 device-resident frames each, same timing rules.
    other = None
    if dist is None and not args.no_other_configs and args.preset == "moshi7b" and args.quant == "q4_k":
        other = {}
        try:
            stream.reset()
            stream.run_resident(frames, cfg["context"] + 8)                      # ring full
            ms_f, _ = stream.run_resident(frames, 200)
            other["moshi7b q4_k, KV ring full"] = {"ms_per_step": ms_f / 200, "frames_per_s": 200 / (ms_f * 1e-3), "kv_slots": min(stream.offset, cfg["context"]),
                                                   "kv_bytes_per_step": int(stream.kv_bytes_next)}
            stream.reset()
        except Exception as e:                       # noqa: BLE001
            other["moshi7b q4_k, KV ring full"] = {"error": str(e)[:200]}
        for o_preset, o_quant in (("stt1b", "q8_0"), ("tts1_6b", "q4_k")):
            key = f"{o_preset} {o_quant}"
            try:
                ocfg = configs.get(o_preset)
                opath = ensure_gguf(o_preset, o_quant, rank, world, barrier)
                om = msx.Model(opath, ocfg, device=local_rank); os_ = msx.Stream(om)
                if ocfg.get("cross_attention"):
                    os_.set_condition(*tts_condition(ocfg))
                ofr = make_frames(ocfg)
                os_.run_resident(ofr, W)
                torch.cuda.synchronize(local_rank)
                ms_o, _ = os_.run_resident(ofr, 200)
                other[key] = {"ms_per_step": ms_o / 200, "frames_per_s": 200 / (ms_o * 1e-3), "realtime_factor": 200 / (ms_o * 1e-3) / FRAME_RATE,
                              "launches_per_frame": os_.launches_per_frame, "weight_bytes_per_step": int(om.weight_bytes_per_frame)}
                os_.close(); om.close()
            except Exception as e:                   # noqa: BLE001
                other[key] = {"error": str(e)[:200]}

    # ---- CPU baseline (rank 0, every N) + parity of the timed model against it --------------------------------------
    cpu, parity = None, None
    if rank == 0 and not args.no_cpu_baseline:
        fps_cpu, n, dt, threads, cpu_trace = cpu_reference_fps(path, cfg, frames, 64, args.cpu_budget)
        cpu = {"value": fps_cpu, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": f"{n} frames ({dt:.1f} s) of the same {args.preset} {args.quant} GGUF and token seed through oracle/ggml_ref.c "
                         f"(CPU restatement of the reference's ggml-CPU path, OpenMP, {threads} threads set explicitly)"}
        gpu_trace = gpu_token_trace(msx, stream, cfg, frames, len(cpu_trace))
        bad = next((i for i, (a, b) in enumerate(zip(gpu_trace, cpu_trace)) if a != b), None)
        parity = {"frames": len(cpu_trace), "identical": bad is None, "first_mismatch": bad,
                  "what": "greedy text + audio tokens of the timed GGUF through the per-frame API vs the CPU oracle, same inputs"}
    barrier()

    if rank == 0:
        line = {
            "metric": "frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_res / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8 dot (q4_k x q8_K) / f32", "data": "synthetic",
            "config": base_cfg, "realtime_factor": fps / world / FRAME_RATE,
            "step_bytes": {"weights": w_bytes, "kv_avg": kv_avg, "gbs": step_gbs, "frac_of_peak": step_gbs / peaks["hbm_gbs"],
                           "frac_of_8tbs": step_gbs / 8000.0},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": 336, "d2h_bytes_per_step": 176,
                    "ms_per_step": ms_e2e / K, "wall_ms_per_step": wall_e2e / K},
            "gpu_launches": stream.launches_per_frame * K,
            "launches_per_frame": stream.launches_per_frame,
            "clocks": clocks, "load_s": t_load,
            "stacks": stacks,
            "batched_streams": batched,
            "parity_checked": parity["frames"] if parity else 0, "parity": parity,
            "step_kernel": step_kernel, "tensor_parallel": tensor_parallel, "other_configs": other,
        }
        emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
