"""CPU-only check of the product's 3-D D3Q19 KERNEL SOURCE (mglc_b200/csrc/lbm_kernels.inl) against the oracle:
tests/host_shim/lbm_host.cpp compiles the same .inl for the host and sweeps (blockIdx, threadIdx) sequentially.  On P emulated
subdomains it runs the fused kernel WITH the direct halo stores (PeerTable) and requires, after each launch, the interior of
every subdomain's new f_post AND exactly the halo entries the reference's message_passing_sendrecv() fills (5 populations per
face, 1 per edge, nothing else) to equal the oracle's, bit for bit in the strict build.  The GPU parity tests proper are
tests/test_lid_gpu.py."""
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
    out = str(tmp_path_factory.mktemp("shim") / "lbm_host.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", "-ffp-contract=off", "-Wl,-Bsymbolic", "-I/usr/local/cuda/include",
                           "-o", out, os.path.join(ROOT, "tests", "host_shim", "lbm_host.cpp")])
    S = C.CDLL(out)       # -Bsymbolic: libmglc.so (RTLD_GLOBAL) exports host stubs with the kernels' names; bind to the shim's own
    S.lbm_shim_create.restype = C.c_void_p
    S.lbm_shim_create.argtypes = [C.c_int] * 3 + [ip, C.c_int, dp, dp, C.c_int]
    S.lbm_shim_destroy.argtypes = [C.c_void_p]
    for fn in ("lbm_shim_put", "lbm_shim_get"):
        getattr(S, fn).argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
    for fn in ("lbm_shim_put_lid", "lbm_shim_get_lid", "lbm_shim_put_force", "lbm_shim_get_force"):
        getattr(S, fn).argtypes = [C.c_void_p, C.c_int, dp]
    S.lbm_shim_set_peer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    S.lbm_shim_fused.argtypes = [C.c_void_p, C.c_int, C.c_int]
    S.lbm_shim_th_fused.argtypes = [C.c_void_p, C.c_int, C.c_int]
    S.lbm_shim_stream_macro.argtypes = [C.c_void_p, C.c_int] + [dp] * 5
    return S


def ptr(a):
    return a.ctypes.data_as(dp)


def make_subs(S, wd, strict, collision="mrt"):
    subs = []
    for R in wd.ranks:
        wall = (C.c_int * 6)(*[int(R.coords[a] == (wd.dims[a] - 1 if plus else 0)) for a in range(3) for plus in (True, False)])
        par = (C.c_double * 5)(wd.Snu, wd.Sq, wd.U0, wd.rho0, float(collision == "bgk"))
        subs.append(S.lbm_shim_create(*R.n, wall, int(R.coords[2] == wd.dims[2] - 1), par, None, int(strict)))
    for r, R in enumerate(wd.ranks):                 # PeerTable: face d (0..5) = nbr_surface(d+1), edge a (7..18) = nbr_line(a)
        for b in (0, 1):
            for d in range(6):
                if R.nbr_surface[d + 1] >= 0:
                    S.lbm_shim_set_peer(subs[r], b, d, subs[R.nbr_surface[d + 1]])
            for a in range(7, 19):
                if R.nbr_line[a] >= 0:
                    S.lbm_shim_set_peer(subs[r], b, a, subs[R.nbr_line[a]])
    return subs


def halo_mask(shape):
    m = np.ones(shape, dtype=bool)
    m[:, 1:-1, 1:-1, 1:-1] = False
    return m


@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 1, 1), (2, 2, 1), (1, 2, 2), (2, 2, 2), (3, 1, 2)])
@pytest.mark.parametrize("collision", ["mrt", "bgk"])
def test_fused_kernel_with_direct_halo_stores_reproduces_the_oracle(shim, dims, collision):
    total = (11, 9, 7)
    P = dims[0] * dims[1] * dims[2]
    wd = orc.LidWorld(total, P, dims=dims, collision=collision)
    wd.initial()
    wd.step(4)
    wd.collision(); wd.message_passing_sendrecv()                       # the rotated loop's state: f_post with valid halos
    subs = make_subs(shim, wd, strict=True, collision=collision)
    for h, R in zip(subs, wd.ranks):
        fp = R.f_post.copy(order="F")
        # the reference leaves wall halos (and unused halo entries) undefined: poison everything a message did not fill
        tmp = orc.LidWorld(total, P, dims=dims)                          # which entries does an exchange fill? ask the oracle
        for Q in tmp.ranks:
            Q.f_post[...] = 0.0
            Q.f_post[:, 1:-1, 1:-1, 1:-1] = 1.0
        tmp.message_passing_sendrecv()
        filled = tmp.ranks[wd.ranks.index(R)].f_post == 1.0
        tmp.close()
        fp[~filled] = np.nan
        shim.lbm_shim_put(h, 19, 0, ptr(fp))
        shim.lbm_shim_put_lid(h, 0, ptr(np.asfortranarray(R.rho[:, :, -1])))
    cur = 0
    for step in range(2):
        for h in subs:
            shim.lbm_shim_fused(h, cur, 1)
        cur ^= 1
        wd.streaming(); wd.bounceback(); wd.macro(); wd.collision()
        for R in wd.ranks:                                               # which halo entries does THIS exchange write?
            R.f_post[halo_mask(R.f_post.shape)] = np.nan
        wd.message_passing_sendrecv()
        for h, R in zip(subs, wd.ranks):
            got = np.empty(R.f_post.shape, order="F")
            shim.lbm_shim_get(h, 19, cur, ptr(got))
            assert np.array_equal(got[:, 1:-1, 1:-1, 1:-1], R.f_post[:, 1:-1, 1:-1, 1:-1]), (dims, step, "interior")
            hm = halo_mask(got.shape)
            # exactly the reference's messages arrive, with the reference's values
            assert np.array_equal(np.isnan(got[hm]), np.isnan(R.f_post[hm])), (dims, step, "which halo entries")
            assert np.array_equal(got[hm][~np.isnan(got[hm])], R.f_post[hm][~np.isnan(R.f_post[hm])]), (dims, step, "halo values")
            if R.coords[2] == dims[2] - 1:
                lid = np.empty(R.n[:2], order="F")
                shim.lbm_shim_get_lid(h, cur, ptr(lid))
                assert np.array_equal(lid, R.rho[:, :, -1])
        # the next launch must not read anything the messages did not fill: re-poison the lattice just consumed
    # epilogue: streaming + bounceback + macro from the last lattice
    wd.streaming(); wd.bounceback(); wd.macro()
    for h, R in zip(subs, wd.ranks):
        f = np.empty(R.f.shape, order="F")
        fl = [np.empty(R.n, order="F") for _ in range(4)]
        shim.lbm_shim_stream_macro(h, cur, ptr(f), *[ptr(a) for a in fl])
        assert np.array_equal(f, R.f)
        for a, k in zip(fl, ("rho", "u", "v", "w")):
            assert np.array_equal(a, getattr(R, k)), k
    for h in subs:
        shim.lbm_shim_destroy(h)
    wd.close()


def test_fast_build_of_the_fused_kernel_tracks_the_oracle(shim):
    total = (17, 16, 15)
    wd = orc.LidWorld(total, 1)
    wd.initial()
    wd.collision()
    (h,) = make_subs(shim, wd, strict=False)
    R = wd.ranks[0]
    fp = R.f_post.copy(order="F")
    fp[halo_mask(fp.shape)] = np.nan
    shim.lbm_shim_put(h, 19, 0, ptr(fp))
    shim.lbm_shim_put_lid(h, 0, ptr(np.asfortranarray(R.rho[:, :, -1])))
    n = 60
    for s in range(n - 1):
        shim.lbm_shim_fused(h, s & 1, 0)
    f = np.empty(R.f.shape, order="F")
    fl = [np.empty(R.n, order="F") for _ in range(4)]
    shim.lbm_shim_stream_macro(h, (n - 1) & 1, ptr(f), *[ptr(a) for a in fl])
    wd.message_passing_sendrecv(); wd.streaming(); wd.bounceback(); wd.macro()
    wd.step(n - 1)
    for a, k in zip(fl, ("rho", "u", "v", "w")):
        b = getattr(R, k)
        assert np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300) <= 1e-12 and np.abs(a - b).max() <= 1e-10, k
    shim.lbm_shim_destroy(h)
    wd.close()


# ---------------- thermal double-distribution kernel (thermal_kernels.inl: k_th_fused) ----------------
EX = [0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0]
EY = [0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1]
EZ = [0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1]


def make_thermal_subs(S, wd, strict):
    p, d = wd.p, wd.dims
    rank_of = lambda c: (c[0] * d[1] + c[1]) * d[2] + c[2] if all(0 <= c[a] < d[a] for a in range(3)) else -1
    subs = []
    for R in wd.ranks:
        wall = (C.c_int * 6)(*[int(R.coords[a] == (d[a] - 1 if plus else 0)) for a in range(3) for plus in (True, False)])
        par = (C.c_double * 5)(p.Snu, p.Sq, 0.0, 1.0, 0.0)
        wallT = [(6.0 + p.paraA) / 21.0 * (p.Thot if k == orc.TH_CONST_HOT else p.Tcold) for k in wd.bcT]          # B3:1128-1163
        tpar = (C.c_double * 22)(p.Snu, p.Sq, p.Qd, p.Qnu, p.paraA, p.gBeta, p.Tref, p.omegaRatating, p.Thot, p.Tcold, *wallT,
                                 *[float(k) for k in wd.bcT])
        subs.append(S.lbm_shim_create(*R.n, wall, 0, par, tpar, int(strict)))
    for r, R in enumerate(wd.ranks):
        for b in (0, 1):
            for dd in range(6):
                e = [0, 0, 0]
                e[dd >> 1] = 1 if not (dd & 1) else -1
                n = rank_of([R.coords[a] + e[a] for a in range(3)])
                if n >= 0:
                    S.lbm_shim_set_peer(subs[r], b, dd, subs[n])
            for a in range(7, 19):
                n = rank_of([R.coords[0] + EX[a], R.coords[1] + EY[a], R.coords[2] + EZ[a]])
                if n >= 0:
                    S.lbm_shim_set_peer(subs[r], b, a, subs[n])
    return subs


@pytest.mark.parametrize("dims,bcT", [((1, 1, 1), None), ((2, 2, 2), None), ((1, 2, 2), (0, 0, 0, 0, 2, 1)), ((3, 1, 2), None)])
def test_thermal_fused_kernel_with_direct_halo_stores_reproduces_the_oracle(shim, dims, bcT):
    total = (11, 9, 7)
    P = dims[0] * dims[1] * dims[2]
    wd = orc.ThermalWorld(total, P, dims=dims, bcT=bcT)
    wd.initial()
    wd.step(4)
    wd.collision(); wd.collisionT()
    for R in wd.ranks:                          # everything a message does not fill is poison: the kernel must not read it
        R.f_post[halo_mask(R.f_post.shape)] = np.nan; R.g_post[halo_mask(R.g_post.shape)] = np.nan
    wd.f_message_passing_sendrecv(); wd.g_message_passing_sendrecv()
    subs = make_thermal_subs(shim, wd, strict=True)
    for h, R in zip(subs, wd.ranks):
        shim.lbm_shim_put(h, 19, 0, ptr(R.f_post.copy(order="F")))
        shim.lbm_shim_put(h, 7, 0, ptr(R.g_post.copy(order="F")))
        shim.lbm_shim_put_force(h, 0, ptr(np.concatenate([a.ravel(order="F") for a in (R.Fx, R.Fy, R.Fz)])))
    cur = 0
    for step in range(2):
        for h in subs:
            shim.lbm_shim_th_fused(h, cur, 1)
        cur ^= 1
        wd.streaming(); wd.bounceback(); wd.streamingT(); wd.bouncebackT(); wd.macro(); wd.macroT()
        wd.collision(); wd.collisionT()
        for R in wd.ranks:
            R.f_post[halo_mask(R.f_post.shape)] = np.nan; R.g_post[halo_mask(R.g_post.shape)] = np.nan
        wd.f_message_passing_sendrecv(); wd.g_message_passing_sendrecv()
        for h, R in zip(subs, wd.ranks):
            for nq, want in ((19, R.f_post), (7, R.g_post)):
                got = np.empty(want.shape, order="F")
                shim.lbm_shim_get(h, nq, cur, ptr(got))
                assert np.array_equal(got[:, 1:-1, 1:-1, 1:-1], want[:, 1:-1, 1:-1, 1:-1]), (dims, step, nq, "interior")
                hm = halo_mask(got.shape)
                assert np.array_equal(np.isnan(got[hm]), np.isnan(want[hm])), (dims, step, nq, "which halo entries")
                assert np.array_equal(got[hm][~np.isnan(got[hm])], want[hm][~np.isnan(want[hm])]), (dims, step, nq, "halo values")
            fc = np.empty(3 * R.Fx.size)
            shim.lbm_shim_get_force(h, cur, ptr(fc))
            assert np.array_equal(fc, np.concatenate([a.ravel(order="F") for a in (R.Fx, R.Fy, R.Fz)])), (dims, step, "force")
    for h in subs:
        shim.lbm_shim_destroy(h)
    wd.close()
