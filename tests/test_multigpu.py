"""Multi-GPU parity (one process per GPU, halos over NCCL): spawns tests/dist/nccl_worker.py under
torch.distributed.run on every visible GPU (2, 4 or 8) and checks its verdict.  Needs >= 2 CUDA devices."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
@pytest.mark.multigpu
def test_nccl_decomposed_runs_match_the_oracle():
    import torch
    n = torch.cuda.device_count()
    n = 8 if n >= 8 else 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(ROOT, "tests", "dist", "nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    sys.stdout.write(r.stdout[-6000:]); sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0 and "MULTIGPU OK" in r.stdout
