"""CPU-only check of the product's 2-D lid-driven cavity KERNEL SOURCE (mglc_b200/csrc/lid2d_kernels.inl) against the oracle:
tests/host_shim/l2d_host.cpp compiles the same .inl for the host and sweeps (blockIdx, threadIdx) sequentially.  Both shipped
programs' roundings, every kind of subdomain (wall / neighbour on each side), the lid term in the top corners, wall halos poisoned.
The GPU parity tests proper are tests/test_lid2d_gpu.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "l2d_host.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", "-ffp-contract=off", "-Wl,-Bsymbolic", "-I/usr/local/cuda/include",
                           "-o", out, os.path.join(ROOT, "tests", "host_shim", "l2d_host.cpp")])
    S = C.CDLL(out)
    S.shim_l2d.argtypes = [C.c_int] * 5 + [ip] + [C.c_double] * 4 + [dp] * 5
    return S


def run(S, wd, R, mode, strict, fin, lid_in, fields=None):
    nx, ny = R.n
    wall = (C.c_int * 4)(R.coords[0] == wd.dims[0] - 1, R.coords[0] == 0, R.coords[1] == wd.dims[1] - 1, R.coords[1] == 0)
    fin = np.asfortranarray(fin)
    fout = np.zeros((9, nx + 2, ny + 2), order="F")
    lid_in = np.ascontiguousarray(lid_in, dtype=np.float64)
    lid_out = np.full(nx, np.nan)
    fl = np.zeros((3, nx * ny)) if fields is None else np.ascontiguousarray(np.stack([np.asfortranarray(a).ravel(order="F") for a in fields]))
    rc = S.shim_l2d(mode, int(strict), int(wd.variant == "f"), nx, ny, wall, wd.Snu, wd.Sq, 0.1, 1.0, fin.ctypes.data_as(dp),
                    lid_in.ctypes.data_as(dp), fout.ctypes.data_as(dp), lid_out.ctypes.data_as(dp), fl.ctypes.data_as(dp))
    assert rc == 0
    return fout, lid_out, [fl[q].reshape((nx, ny), order="F") for q in range(3)]


@pytest.mark.parametrize("variant", ["c", "f"])
@pytest.mark.parametrize("dims", [(1, 1), (2, 2), (3, 3), (1, 3), (3, 1)])
@pytest.mark.parametrize("strict", [True, False])
def test_fused_kernel_source_reproduces_one_oracle_step(shim, variant, dims, strict):
    wd = orc.Lid2DWorld((23, 19), dims[0] * dims[1], dims, variant=variant)
    wd.initial()
    wd.step(30)
    wd.collision(); wd.message_passing_sendrecv()
    snap = []
    for R in wd.ranks:
        fp = R.f_post.copy()
        if R.coords[0] == 0: fp[:, 0, :] = np.nan
        if R.coords[0] == dims[0] - 1: fp[:, -1, :] = np.nan
        if R.coords[1] == 0: fp[:, :, 0] = np.nan
        if R.coords[1] == dims[1] - 1: fp[:, :, -1] = np.nan
        snap.append((fp, R.rho[:, -1].copy()))
    wd.streaming(); wd.bounceback(); wd.macro()
    macros = [(R.f.copy(), R.rho.copy(), R.u.copy(), R.v.copy()) for R in wd.ranks]
    wd.collision()
    for R, (fp, lid), mac in zip(wd.ranks, snap, macros):
        fo, lid_out, _ = run(shim, wd, R, 0, strict, fp, lid)
        want = R.f_post[:, 1:-1, 1:-1]
        if strict:
            assert np.array_equal(fo[:, 1:-1, 1:-1], want)
        else:
            assert np.abs(fo[:, 1:-1, 1:-1] - want).max() < 1e-15
        if R.coords[1] == dims[1] - 1:
            assert np.array_equal(lid_out, mac[1][:, -1])           # the lid row of this step's rho, for the next step's lid term
        fo, _, fl = run(shim, wd, R, 1, strict, fp, lid)
        assert np.array_equal(fo[:, 1:-1, 1:-1], mac[0])
        for got, w in zip(fl, mac[1:]):
            assert np.array_equal(got, w)
    wd.close()


@pytest.mark.parametrize("variant", ["c", "f"])
def test_collision_kernel_source_and_a_fast_run(shim, variant):
    wd = orc.Lid2DWorld((33, 29), 1, variant=variant)
    wd.initial()
    wd.step(20)
    R = wd.ranks[0]
    f = np.full((9, 35, 31), np.nan, order="F")
    f[:, 1:-1, 1:-1] = R.f
    fields = [R.rho.copy(), R.u.copy(), R.v.copy()]
    lid = R.rho[:, -1].copy()
    wd.collision()
    fo, _, _ = run(shim, wd, R, 2, True, f, lid, fields)
    assert np.array_equal(fo[:, 1:-1, 1:-1], R.f_post[:, 1:-1, 1:-1])
    fo, _, _ = run(shim, wd, R, 2, False, f, lid, fields)
    assert np.abs(fo[:, 1:-1, 1:-1] - R.f_post[:, 1:-1, 1:-1]).max() < 1e-15
    # 150 rotated steps in the throughput arithmetic
    fp = fo
    wd.streaming(); wd.bounceback(); wd.macro()
    n = 150
    for _ in range(n - 1):
        fp, lid, _ = run(shim, wd, R, 0, False, fp, lid)
    _, _, fl = run(shim, wd, R, 1, False, fp, lid)
    wd.step(n - 1)
    for got, k in zip(fl, ("rho", "u", "v")):
        want = getattr(R, k)
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-12 and np.abs(got - want).max() < 1e-10, k
    wd.close()
