"""Tensor-parallel single stream (BASELINE.json config 4) on two GPUs of one box: every rank must reproduce the single-GPU
arithmetic bit for bit (logits and tokens), with both all-reduce implementations (NCCL inside the CUDA graph; fused
GEMV -> peer-memory push).  Skipped on boxes with one GPU.  Worker = scripts/tp_check.py, launched the way bench.py is
(torch.distributed.run, 127.0.0.1 rendezvous); rank 0 also runs the un-sharded model and compares frame by frame."""
import json, os, socket, subprocess, sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_tp(world, preset, quant, p2p, frames, fill=0):
    env = dict(os.environ, PRESET=preset, QUANT=quant, TP_P2P=str(p2p), FRAMES=str(frames), FILL=str(fill), TIME="0", CHECK="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "scripts", "tp_check.py")]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("TP_CHECK ")]
    assert r.returncode == 0 and lines, f"tp_check failed (rc {r.returncode}):\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}"
    return json.loads(lines[-1][len("TP_CHECK "):])


@pytest.mark.gpu
@pytest.mark.parametrize("p2p", [0, 1], ids=["nccl", "peer-memory"])
@pytest.mark.parametrize("preset,quant,frames", [("tiny", "q4_k", 30), ("tiny_pplex", "q8_0", 30), ("moshi7b_l2", "q4_k", 8)])
def test_tensor_parallel_two_gpus_bit_identical(preset, quant, frames, p2p):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on the box")
    out = _run_tp(2, preset, quant, p2p, frames)
    n = out["logit_vectors_bit_identical"].split("/")
    assert out["ranks_agree"] and out["token_mismatches"] == 0, out
    assert n[0] == n[1] and int(n[1]) > 0, out            # every text / audio logit vector identical to the single-GPU stream
    assert out["worst_max_rel_vs_single_gpu"] == 0.0, out
