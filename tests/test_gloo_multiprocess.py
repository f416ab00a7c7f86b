"""CPU test of the N > 1 host path with real processes (torch.distributed, gloo backend, 127.0.0.1): the halo
plan of include/mglc.h driven across processes reproduces the reference's message_passing_sendrecv()."""
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


@pytest.mark.parametrize("world,dims", [(2, None), (4, None), (4, "1x2x2")])
def test_halo_plan_across_processes(world, dims):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(ROOT, "tests", "dist", "gloo_halo_worker.py")] + ([dims] if dims else [])
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and "GLOO HALO OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


@pytest.mark.parametrize("world,dims", [(2, None), (4, None), (4, "1x4"), (6, "3x2")])
def test_halo_plan_2d_across_processes(world, dims):
    """the 2-D drivers' message table (mglc_halo_plan_2d, which the CUDA drivers check their own table against) moved with real
    processes == the 2-D oracles' exchanges, for the lid and the thermal driver"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(ROOT, "tests", "dist", "gloo_halo2d_worker.py")] + ([dims] if dims else [])
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and "GLOO HALO 2D OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
