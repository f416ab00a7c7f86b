"""CPU-only check of the product's 2-D lid-driven cavity KERNEL SOURCE (mglc_b200/csrc/lid2d_kernels.inl) against the oracle:
tests/host_shim/l2d_host.cpp compiles the same .inl files (lid2d_kernels.inl, lid2d_exact.inl) for the host and sweeps
(blockIdx, threadIdx) sequentially.  All shipped arithmetics (C with MRT or with its SRT switch, Fortran + MPI, the incompressible sequential program), every kind of subdomain (wall / neighbour on each side), the lid term in the top corners, wall halos poisoned.
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


def run(S, wd, R, mode, strict, fin, lid_in, fields=None, fout=None):
    nx, ny = R.n
    wall = (C.c_int * 4)(R.coords[0] == wd.dims[0] - 1, R.coords[0] == 0, R.coords[1] == wd.dims[1] - 1, R.coords[1] == 0)
    fin = np.asfortranarray(fin)
    fout = np.zeros((9, nx + 2, ny + 2), order="F") if fout is None else np.asfortranarray(fout).copy(order="F")
    lid_in = np.ascontiguousarray(lid_in, dtype=np.float64)
    lid_out = np.full(nx, np.nan)
    fl = np.zeros((3, nx * ny)) if fields is None else np.ascontiguousarray(np.stack([np.asfortranarray(a).ravel(order="F") for a in fields]))
    rc = S.shim_l2d(mode, int(strict), orc.L2_VARIANTS[wd.variant], nx, ny, wall, wd.Snu, wd.Sq, 0.1, 1.0, fin.ctypes.data_as(dp),
                    lid_in.ctypes.data_as(dp), fout.ctypes.data_as(dp), lid_out.ctypes.data_as(dp), fl.ctypes.data_as(dp))
    assert rc == 0
    return fout, lid_out, [fl[q].reshape((nx, ny), order="F") for q in range(3)]


@pytest.mark.parametrize("variant", ["c", "f", "i", "s"])
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


@pytest.mark.parametrize("variant", ["c", "f", "i", "s"])
def test_collision_kernel_source_and_a_fast_run(shim, variant):
    # the single-relaxation-time operator is unstable at Re = 1000 on a 33-cell cavity (tau = 0.51): Re = 100 there
    wd = orc.Lid2DWorld((33, 29), 1, variant=variant, Re=100.0 if variant == "s" else 1000.0)
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


@pytest.mark.parametrize("variant", ["c", "f", "i", "s"])
@pytest.mark.parametrize("dims", [(1, 1), (2, 2), (3, 1)])
def test_per_subroutine_kernel_source(shim, variant, dims):
    """lid2d_exact.inl: k_l2_initial, k_l2_streaming, k_l2_bounceback, k_l2_macro of every block against the oracle's
    subroutines, bit for bit (the incompressible program: rho = 0 and f without the rho factor in initial(), the lid term without
    rho, u and v undivided)"""
    wd = orc.Lid2DWorld((17, 13), dims[0] * dims[1], dims, variant=variant)
    wd.initial()
    for R in wd.ranks:
        nx, ny = R.n
        z = np.zeros((9, nx + 2, ny + 2), order="F")
        fo, _, fl = run(shim, wd, R, 3, True, z, np.zeros(nx))
        assert np.array_equal(fo[:, 1:-1, 1:-1], R.f)
        for got, k in zip(fl, ("rho", "u", "v")):
            assert np.array_equal(got, getattr(R, k)), k
    wd.step(7)
    wd.collision(); wd.message_passing_sendrecv()
    rng = np.random.default_rng(5)
    for R in wd.ranks:                                    # wall halos hold junk, as in the reference (allocate, never written)
        for sl, on in (((slice(None), 0), R.coords[0] == 0), ((slice(None), -1), R.coords[0] == dims[0] - 1)):
            if on: R.f_post[sl] = rng.random(R.f_post[sl].shape)
        if R.coords[1] == 0: R.f_post[:, :, 0] = rng.random(R.f_post[:, :, 0].shape)
        if R.coords[1] == dims[1] - 1: R.f_post[:, :, -1] = rng.random(R.f_post[:, :, -1].shape)
    posts = [R.f_post.copy() for R in wd.ranks]
    rhos = [R.rho.copy() for R in wd.ranks]
    wd.streaming()
    streamed = [R.f.copy() for R in wd.ranks]
    wd.bounceback()
    bounced = [R.f.copy() for R in wd.ranks]
    wd.macro()
    for R, fp, rho, st, bb in zip(wd.ranks, posts, rhos, streamed, bounced):
        nx, ny = R.n
        fo, _, _ = run(shim, wd, R, 4, True, fp, np.zeros(nx))
        assert np.array_equal(fo[:, 1:-1, 1:-1], st)
        f_in = np.zeros((9, nx + 2, ny + 2), order="F")
        f_in[:, 1:-1, 1:-1] = st
        got, _, _ = run(shim, wd, R, 5, True, fp, np.zeros(nx), [rho, rho, rho], fout=f_in)
        assert np.array_equal(got[:, 1:-1, 1:-1], bb)
        f_in[:, 1:-1, 1:-1] = bb
        _, _, fl = run(shim, wd, R, 6, True, f_in, np.zeros(nx))
        for g_, k in zip(fl, ("rho", "u", "v")):
            assert np.array_equal(g_, getattr(R, k)), k
    wd.close()


def test_incompressible_first_step_uses_rho_zero(shim):
    """L2I:137: rho is 0 until the first macro(), so the first collision() relaxes towards meq(1) = 3|u|^2, meq(2) = -3|u|^2;
    the product's collision kernel takes rho from the field, not from the populations"""
    wd = orc.Lid2DWorld((12, 9), 1, variant="i")
    wd.initial()
    R = wd.ranks[0]
    assert not R.rho.any()
    f = np.zeros((9, 14, 11), order="F")
    f[:, 1:-1, 1:-1] = R.f
    fields = [R.rho.copy(), R.u.copy(), R.v.copy()]
    wd.collision()
    for strict in (True, False):
        fo, _, _ = run(shim, wd, R, 2, strict, f, np.zeros(12), fields)
        d = np.abs(fo[:, 1:-1, 1:-1] - R.f_post[:, 1:-1, 1:-1]).max()
        assert d == 0 if strict else d < 1e-15
    wd.close()
