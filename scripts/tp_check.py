"""Tensor-parallel single stream (BASELINE.json config 4): parity against the single-GPU stream and timing.
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 scripts/tp_check.py
Every rank holds a shard (heads / hidden slice) of the temporal transformer; rank 0 also runs the un-sharded model
on its own GPU and compares logits + tokens frame by frame."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkgload; _pkgload.load()
import torch
import torch.distributed as dist
from moshi_cpp_b200 import binding as msx, configs, synth


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    preset, quant = os.environ.get("PRESET", "tiny"), os.environ.get("QUANT", "q4_k")
    frames, ctx_fill = int(os.environ.get("FRAMES", 24)), int(os.environ.get("FILL", 0))
    dist.init_process_group("gloo")
    cfg = configs.get(preset)
    if rank == 0:
        path = synth.cached_gguf(preset, quant)
    dist.barrier()
    path = synth.cached_gguf(preset, quant)
    ids = [msx.tp_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    model = msx.Model(path, cfg, device=local, tp_rank=rank, tp_world=world)
    stream = msx.Stream(model, nccl_id=ids[0])
    if os.environ.get("TP_P2P", "0") == "1":       # fused GEMV -> peer-memory all-reduce instead of NCCL launches
        hs = [None] * world
        dist.all_gather_object(hs, stream.tp_export())
        stream.tp_connect(hs)
    ref = msx.Stream(msx.Model(path, cfg, device=local)) if (rank == 0 and os.environ.get("CHECK", "1") == "1") else None
    dist.barrier()
    rng = np.random.default_rng(5)
    n_q, dep_q = cfg["n_q"], cfg["dep_q"]
    toks = np.array([cfg["text_card"]] + [cfg["card"]] * n_q, dtype=np.int32)
    worst, exact, cmp_n, tok_bad = 0.0, 0, 0, 0
    for f in range(frames):
        t, lg, to = stream.step_temporal(toks)
        a, al = stream.step_depformer(t) if dep_q else (np.zeros(0, np.int32), None)
        if ref is not None:
            t2, lg2, to2 = ref.step_temporal(toks)
            a2, al2 = ref.step_depformer(t2) if dep_q else (np.zeros(0, np.int32), None)
            rel = float(np.max(np.abs(lg - lg2)) / max(1e-30, np.max(np.abs(lg2))))
            worst = max(worst, rel); cmp_n += 1; exact += int(np.array_equal(lg.view(np.uint32), lg2.view(np.uint32)))
            if dep_q:
                rel = float(np.max(np.abs(al - al2)) / max(1e-30, np.max(np.abs(al2))))
                worst = max(worst, rel); cmp_n += 1; exact += int(np.array_equal(al.view(np.uint32), al2.view(np.uint32)))
            tok_bad += int(t != t2) + int(np.sum(a != a2))
        user = rng.integers(0, cfg["card"], size=n_q - dep_q if dep_q else n_q)
        toks = np.concatenate([[t], a, user]).astype(np.int32)
    # all ranks must have produced identical tokens
    mine = torch.tensor([int(t)] + [int(v) for v in a], dtype=torch.int64)
    allt = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allt, mine)
    same = all(torch.equal(allt[0], x) for x in allt)
    # timing: resident replay, all ranks in lock-step
    res = {}
    if os.environ.get("TIME", "1") == "1":
        fr = rng.integers(0, cfg["card"], size=(64, n_q + 1)).astype(np.int32)
        fr[:, 0] = rng.integers(0, cfg["text_card"], size=64)
        if ctx_fill:
            stream.run_resident(fr, ctx_fill)
        stream.run_resident(fr, 20)
        dist.barrier()
        ms, _ = stream.run_resident(fr, 200)
        tms = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        res = {"ms_per_frame": float(tms[0]) / 200, "fps": 200 / (float(tms[0]) * 1e-3), "launches_per_frame": stream.launches_per_frame,
               "offset_at_timing": stream.offset}
        if ref is not None:
            if ctx_fill:
                ref.run_resident(fr, ctx_fill)
            ref.run_resident(fr, 20)
            ms1, _ = ref.run_resident(fr, 200)
            res["single_gpu_ms_per_frame"] = ms1 / 200
    if rank == 0:
        out = {"preset": preset, "quant": quant, "tp": world, "allreduce": "peer-memory (fused)" if os.environ.get("TP_P2P", "0") == "1" else "nccl", "frames": frames, "worst_max_rel_vs_single_gpu": worst,
               "logit_vectors_bit_identical": f"{exact}/{cmp_n}", "token_mismatches": tok_bad, "ranks_agree": bool(same), **res}
        print("TP_CHECK " + json.dumps(out), flush=True)
        ok = same and (ref is None or (tok_bad == 0 and worst < 2e-3))
    else:
        ok = True
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
