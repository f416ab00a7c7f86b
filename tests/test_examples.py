"""examples/*.c: plain-C drivers that bind libmglc.so through include/mglc.h alone (CPU: they compile as C99 with -Wall -Werror
and fail loudly without a device; the GPU run of laplace2d_driver.c is in test_zz_examples_laplace_gpu.py).
examples/lid2d_driver.c: a plain-C driver that binds libmglc.so through include/mglc.h alone, in the role of the reference's
own C program (MPI/Lid_driven_cavity/c/lid_driven_cavity.c).  CPU: it compiles as C99 against the header and fails loudly
without a device.  GPU: after 2000 iterations in strict arithmetic its `flow_binary` is byte-identical to the file the reference
program writes (SHA-256 committed in tests/golden/ref_lid2d.npz by make_golden_lid2d.py, from the reference's own run)."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "ref_lid2d.npz"))


def build(tmp_path, name="lid2d_driver"):
    exe = str(tmp_path / name)
    lib = os.path.join(ROOT, "mglc_b200")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", name + ".c"), "-L", lib, "-lmglc", f"-Wl,-rpath,{lib}", "-o", exe])
    return exe


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_c_driver_builds_against_the_header_and_fails_loudly_without_a_gpu(tmp_path):
    exe = build(tmp_path)
    r = subprocess.run([exe, "10", "1", "flow_binary", "3"], capture_output=True, text=True, cwd=tmp_path)      # c:13: SRT = 1, MRT = 2
    assert r.returncode == 2 and "model" in r.stderr
    if _has_gpu():
        pytest.skip("a CUDA device is present: the run itself is covered by the gpu test")
    r = subprocess.run([exe, "10"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "no CUDA device" in r.stderr and not (tmp_path / "flow_binary").exists()
    r = subprocess.run([exe, "10", "1", "flow_binary", "1"], capture_output=True, text=True, cwd=tmp_path)      # model = SRT
    assert r.returncode == 1 and "no CUDA device" in r.stderr


def test_laplace_driver_builds_against_the_header_and_fails_loudly_without_a_gpu(tmp_path):
    """examples/laplace2d_driver.c, in the role of the reference's MPI/Laplace/c/laplace2d.c"""
    exe = build(tmp_path, "laplace2d_driver")
    r = subprocess.run([exe, "2", "2"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 2 and "interior" in r.stderr
    if _has_gpu():
        pytest.skip("a CUDA device is present: the run itself is covered by the gpu test")
    r = subprocess.run([exe, "64", "48", "10", "dump.bin"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "no CUDA device" in r.stderr and not r.stdout and not (tmp_path / "dump.bin").exists()


@pytest.mark.gpu
def test_c_driver_writes_the_reference_programs_file(tmp_path):
    exe = build(tmp_path)
    r = subprocess.run([exe, "2000", "1", "flow_binary"], capture_output=True, text=True, cwd=tmp_path, timeout=300)
    assert r.returncode == 0, r.stderr
    itc, err = r.stdout.split()[-2:]
    assert int(itc) == 2000 and abs(float(err) - 1.0) < 1e-12      # the program's first check(): up = vp = 0, so error = 1
    raw = (tmp_path / "flow_binary").read_bytes()
    assert len(raw) == int(GOLD["c/output_binary_len"][0])
    assert hashlib.sha256(raw).digest() == GOLD["c/output_binary_sha256"].tobytes()
