"""Library-owned NCCL communicator on >= 2 GPUs (skipped on a one-GPU box; tests/run_dp_library_comm.py is the body)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL refuses two ranks on one device)")
def test_library_communicator_two_ranks():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "run_dp_library_comm.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-3000:]
    r = json.loads([ln for ln in proc.stdout.splitlines() if ln.startswith("RESULT ")][-1][len("RESULT "):])
    assert r["world"] == 2 and r["allreduce_max_err"] < 1e-6
    assert r["library_equals_torch_path"] and r["replicas_identical_library"] and not r["replicas_identical_without_allreduce"]
