"""GPU (needs >= 2 devices; skipped otherwise): data-parallel gradient equivalence and replica consistency over NCCL."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_allreduced_gradients_are_the_sum_of_replica_gradients(cuda_dev):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (run under gpurun --gpus 2)")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dp_grad_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("DPRESULT ")]
    assert line, r.stdout[-3000:]
    out = json.loads(line[-1][len("DPRESULT "):])
    print(out)
    # every element of both flat gradient buffers goes through exactly one all-reduce, and the result is the sum of the
    # replicas' local gradients (a two-term fp32 sum: exact)
    assert out["G"]["covered_exactly_once"] and out["D"]["covered_exactly_once"], out
    # G: translator / pose_encoder / image_encoder buckets; D: [conv_5, D_logit] / [conv_3, conv_4] / rest
    assert out["G"]["slices"] >= 3 and out["D"]["slices"] == 3, out
    assert out["G"]["rel_l2"] <= 1e-6 and out["D"]["rel_l2"] <= 1e-6, out
    assert out["replica_skew_after_2_steps"] == 0.0, out
