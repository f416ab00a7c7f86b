"""CPU-only check of the product's 2-D thermal KERNEL SOURCE (mglc_b200/csrc/thermal2d_kernels.inl + d2q9_thermal.inl) against
the oracle: tests/host_shim/t2d_host.cpp compiles the same .inl files for the host and sweeps (blockIdx, threadIdx)
sequentially -- exact for these kernels, which use no shared memory or synchronisation.  Covers the fused kernel's pull
addressing, the unified wall rule on every kind of subdomain (wall / neighbour on each side), the constant-temperature and
adiabatic g rules, the in-place force update and both arithmetic builds, before any GPU time is spent.  The GPU parity tests
proper are tests/test_thermal2d_gpu.py."""
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
    out = str(tmp_path_factory.mktemp("shim") / "t2d_host.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", "-ffp-contract=off", "-Wl,-Bsymbolic", "-I/usr/local/cuda/include", "-o", out,
                           os.path.join(ROOT, "tests", "host_shim", "t2d_host.cpp")])
    S = C.CDLL(out)       # -Bsymbolic: libmglc.so (RTLD_GLOBAL) exports host stubs with the kernels' names; bind to the shim's own
    S.shim_t2d.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, ip, dp, dp, ip, dp, dp, dp, dp, dp, dp]
    return S


def run_shim(S, w, R, mode, strict, fin, gin, Fy, fields=None):
    p = w.params
    nx, ny = R.n
    wall, par, wallT, bcT = shim_params(w, R)
    fin, gin = np.asfortranarray(fin), np.asfortranarray(gin)
    fout, gout = np.zeros((9, nx + 2, ny + 2), order="F"), np.zeros((5, nx + 2, ny + 2), order="F")
    Fy = np.asfortranarray(Fy.copy())
    fl = np.zeros((4, nx * ny)) if fields is None else np.ascontiguousarray(np.stack([np.asfortranarray(a).ravel(order="F") for a in fields]))
    rc = S.shim_t2d(mode, int(strict), nx, ny, wall, par, wallT, bcT, fin.ctypes.data_as(dp), gin.ctypes.data_as(dp),
                    fout.ctypes.data_as(dp), gout.ctypes.data_as(dp), Fy.ctypes.data_as(dp), fl.ctypes.data_as(dp))
    assert rc == 0
    return fout, gout, Fy, [fl[q].reshape((nx, ny), order="F") for q in range(4)]


def shim_params(w, R):
    """Geom2.wall, T2Params as thermal2d.cu fills them (periodic vertical sides are not walls and carry no thermal kind)"""
    p = w.params
    perx = orc.T2_PERIODIC in w.bcT
    wall = (C.c_int * 4)(R.coords[0] == w.dims[0] - 1 and not perx, R.coords[0] == 0 and not perx, R.coords[1] == w.dims[1] - 1, R.coords[1] == 0)
    par = (C.c_double * 25)(p.Snu, p.Sq, p.Qd, p.Qnu, p.paraA, p.gBeta, p.Tref, p.rho0, p.Thot, p.Tcold, float(perx), float(w.variant == "acc"),
                            *w.Uwall, float(w.cornersT and not perx), float(R.start[0]), float(R.start[1]), float(w.total[0]), float(w.total[1]))
    wallT = (C.c_double * 4)(*[(4.0 + p.paraA) / 10.0 * (p.Thot if k == orc.T2_CONST_HOT else p.Tcold) for k in w.bcT])
    bcT = (C.c_int * 4)(*[0 if k == orc.T2_PERIODIC else k for k in w.bcT])
    return wall, par, wallT, bcT


def padded(a):
    """interior array (q, nx, ny) -> halo'd (q, nx+2, ny+2) with NaN halos (the kernels must not read them)"""
    out = np.full((a.shape[0], a.shape[1] + 2, a.shape[2] + 2), np.nan, order="F")
    out[:, 1:-1, 1:-1] = a
    return out


CASES = [((1, 1), orc.T2_SIDE_HEATED), ((2, 2), orc.T2_SIDE_HEATED), ((3, 3), orc.T2_RAYLEIGH_BENARD), ((1, 3), (2, 1, 1, 2)), ((3, 1), (0, 0, 0, 0)),
         ((1, 1), orc.T2_RB_PERIODIC), ((1, 3), orc.T2_RB_PERIODIC)]      # the OpenACC program's set: periodic vertical walls, acc arithmetic


# the sheared Rayleigh-Benard programs (seq/R_B_2d.F90): every wall moving along itself, halves with opposite signs, its corner rule
SHEAR = dict(Uwall=[4e-3, -4e-3, -4e-3, 4e-3, 3e-3, 3.5e-3, 2.5e-3, 2e-3], cornersT=True)
SHEARED_CASES = [((1, 1), orc.T2_RAYLEIGH_BENARD), ((2, 2), orc.T2_RAYLEIGH_BENARD), ((3, 3), orc.T2_RAYLEIGH_BENARD), ((4, 1), orc.T2_SIDE_HEATED)]


def world_for(total, dims, bcT, **kw):
    acc = orc.T2_PERIODIC in bcT
    return orc.Thermal2DWorld(total, nprocs=dims[0] * dims[1], dims=dims, bcT=bcT, variant="acc" if acc else "mpi",
                              lengthUnit=float(total[0]) if acc else 0.0, **kw)


@pytest.mark.parametrize("dims,bcT,walls", [c + ({},) for c in CASES] + [c + (SHEAR,) for c in SHEARED_CASES])
@pytest.mark.parametrize("strict", [True, False])
def test_fused_kernel_source_reproduces_one_oracle_step(shim, dims, bcT, walls, strict):
    """k_t2_fused on every rank == streaming .. macroT of this step + collision/collisionT of the next (oracle), wall halos poisoned;
    with moving walls the kernel reads rho of the previous macro() at the wall cells and leaves the new one there"""
    w = world_for((23, 19), dims, bcT, Rayleigh=1e6, **walls)
    w.initial()
    w.step(30)
    w.collision(); w.message_passing_f(); w.collisionT(); w.message_passing_g()          # rotated-loop state: halos valid
    snap = []
    for R in w.ranks:
        fp, gp = R.f_post.copy(), R.g_post.copy()
        # poison the halo entries no message fills (physical walls and the unused corners): the kernel must never read them
        if R.coords[0] == 0: fp[:, 0, :] = gp[:, 0, :] = np.nan                   # (periodic sides: the wrap reads column nx, never the halo)
        if R.coords[0] == dims[0] - 1: fp[:, -1, :] = gp[:, -1, :] = np.nan
        if R.coords[1] == 0: fp[:, :, 0] = gp[:, :, 0] = np.nan
        if R.coords[1] == dims[1] - 1: fp[:, :, -1] = gp[:, :, -1] = np.nan
        snap.append((fp, gp, R.Fy.copy(), R.rho.copy()))
    # the oracle finishes the step and collides again
    w.streaming(); w.bounceback(); w.streamingT(); w.bouncebackT(); w.macro(); w.macroT()
    macros = [(R.f.copy(), R.g.copy(), R.rho.copy(), R.u.copy(), R.v.copy(), R.T.copy()) for R in w.ranks]
    w.collision(); w.collisionT()
    for R, (fp, gp, Fy, rho_prev), mac in zip(w.ranks, snap, macros):
        prev = [rho_prev, rho_prev * 0, rho_prev * 0, rho_prev * 0]
        fo, go, Fy2, fl0 = run_shim(shim, w, R, 0, strict, fp, gp, Fy, fields=prev)
        if walls:                                   # the wall cells of the rho field now hold this step's macro()
            ring = np.zeros(R.n, bool)
            if R.coords[0] == 0: ring[0, :] = True
            if R.coords[0] == dims[0] - 1: ring[-1, :] = True
            if R.coords[1] == 0: ring[:, 0] = True
            if R.coords[1] == dims[1] - 1: ring[:, -1] = True
            assert np.array_equal(fl0[0][ring], mac[2][ring]) and np.array_equal(fl0[0][~ring], rho_prev[~ring])
        want_f, want_g = R.f_post[:, 1:-1, 1:-1], R.g_post[:, 1:-1, 1:-1]
        if strict:
            assert np.array_equal(fo[:, 1:-1, 1:-1], want_f) and np.array_equal(go[:, 1:-1, 1:-1], want_g) and np.array_equal(Fy2, R.Fy)
        else:
            assert np.abs(fo[:, 1:-1, 1:-1] - want_f).max() < 2e-16 * 4 and np.abs(go[:, 1:-1, 1:-1] - want_g).max() < 1e-15
            assert np.array_equal(Fy2, R.Fy)          # the stored force is rounded like the reference in both builds
        # the epilogue kernel: bit-exact in both builds (copies, ordered adds, IEEE divisions)
        fo, go, _, fl = run_shim(shim, w, R, 1, strict, fp, gp, Fy, fields=prev)
        assert np.array_equal(fo[:, 1:-1, 1:-1], mac[0]) and np.array_equal(go[:, 1:-1, 1:-1], mac[1])
        for got, want in zip(fl, mac[2:]):
            assert np.array_equal(got, want)
            assert np.array_equal(np.signbit(got), np.signbit(want))
    w.close()


@pytest.mark.parametrize("strict", [True, False])
def test_collision_kernel_sources(shim, strict):
    w = orc.Thermal2DWorld((23, 19), Rayleigh=1e6)
    w.initial()
    w.step(40)
    R = w.ranks[0]
    f, g, fields = R.f.copy(), R.g.copy(), [R.rho.copy(), R.u.copy(), R.v.copy(), R.T.copy()]
    w.collision(); w.collisionT()
    fo, go, Fy, _ = run_shim(shim, w, R, 2, strict, padded(f), padded(g), np.zeros_like(R.Fy), fields)
    if strict:
        assert np.array_equal(fo[:, 1:-1, 1:-1], R.f_post[:, 1:-1, 1:-1]) and np.array_equal(go[:, 1:-1, 1:-1], R.g_post[:, 1:-1, 1:-1])
    else:
        assert np.abs(fo[:, 1:-1, 1:-1] - R.f_post[:, 1:-1, 1:-1]).max() < 1e-15
        assert np.abs(go[:, 1:-1, 1:-1] - R.g_post[:, 1:-1, 1:-1]).max() < 1e-15
    assert np.array_equal(Fy, R.Fy)
    w.close()


def test_rotated_loop_in_the_fast_build_tracks_the_oracle(shim):
    """200 steps of [exchange-free single rank] fused kernel in the throughput arithmetic vs the oracle: <= 1e-12 relative L2"""
    w = orc.Thermal2DWorld((33, 29), Rayleigh=1e6)
    w.initial()
    R = w.ranks[0]
    w.collision(); w.collisionT()
    fp, gp, Fy = R.f_post.copy(), R.g_post.copy(), R.Fy.copy()
    w.streaming(); w.bounceback(); w.streamingT(); w.bouncebackT(); w.macro(); w.macroT()
    n = 200
    for _ in range(n - 1):
        fp, gp, Fy, _ = run_shim(shim, w, R, 0, False, fp, gp, Fy)
    _, _, _, fl = run_shim(shim, w, R, 1, False, fp, gp, Fy)
    w.step(n - 1)
    for got, want, tol in zip(fl, (R.rho, R.u, R.v, R.T), (1e-13, 1e-12, 1e-12, 1e-12)):
        scale = max(np.linalg.norm(want), 1e-30)
        assert np.linalg.norm(got - want) / scale < tol
    w.close()


# ---------------- the copy-type kernels (thermal2d_exact.inl) on the CPU, P emulated subdomains ----------------
WHICH = {"f": 0, "f_post": 1, "g": 2, "g_post": 3, "rho": 4, "u": 5, "v": 6, "T": 7, "up": 8, "vp": 9, "Tp": 10, "Fx": 11, "Fy": 12}
EX9 = [0, 1, 0, -1, 0, 1, -1, -1, 1]
EY9 = [0, 0, 1, 0, -1, 1, 1, -1, -1]


class ShimWorld:
    """P subdomains of the product's exact kernels, driven the way thermal2d.cu drives them (same message table)"""

    def __init__(self, S, w):
        self.S, self.w, self.subs = S, w, []
        S.shim_sub_create.restype = C.c_void_p
        S.shim_sub_create.argtypes = [C.c_int, C.c_int, ip, dp, dp, ip]
        S.shim_sub_destroy.argtypes = [C.c_void_p]
        S.shim_sub_put.argtypes = [C.c_void_p, C.c_int, dp]
        S.shim_sub_get.argtypes = [C.c_void_p, C.c_int, dp]
        S.shim_sub_op.argtypes = [C.c_void_p] + [C.c_int] * 4
        S.shim_sub_pack.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp]
        S.shim_sub_unpack.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp]
        for R in w.ranks:
            wall, par, wallT, bcT = shim_params(w, R)
            self.subs.append(S.shim_sub_create(R.n[0], R.n[1], wall, par, wallT, bcT))

    def close(self):
        for h in self.subs:
            self.S.shim_sub_destroy(h)

    def shape(self, r, name):
        nx, ny = self.w.ranks[r].n
        return {"f": (9, nx, ny), "f_post": (9, nx + 2, ny + 2), "g": (5, nx, ny), "g_post": (5, nx + 2, ny + 2)}.get(name, (nx, ny))

    def put(self, r, name, a):
        a = np.asfortranarray(a, dtype=np.float64)
        assert a.shape == self.shape(r, name)
        self.S.shim_sub_put(self.subs[r], WHICH[name], a.ctypes.data_as(dp))

    def get(self, r, name):
        a = np.empty(self.shape(r, name), order="F")
        self.S.shim_sub_get(self.subs[r], WHICH[name], a.ctypes.data_as(dp))
        return a

    def op(self, op, *args):
        args = list(args) + [0] * (3 - len(args))
        for h in self.subs:
            assert self.S.shim_sub_op(h, op, *args) == 0

    def initial(self):
        w = self.w
        isT = [k in (orc.T2_CONST_HOT, orc.T2_CONST_COLD) for k in w.bcT]
        vert, hor = isT[0] or isT[1], isT[2] or isT[3]
        profile = 2 if hor else 1 if vert else 0
        for h, R in zip(self.subs, w.ranks):
            axis = 1 if profile == 2 else 0
            assert self.S.shim_sub_op(h, 0, profile, R.start[axis], w.total[axis]) == 0

    def exchange(self, which):
        """the message table of thermal2d.cu (t2_make_sub): dir 0..3 f faces, 4..7 f corners, 8..11 g faces"""
        w = self.w
        d0, d1 = w.dims
        rank = lambda c0, c1: c0 * d1 + c1 if 0 <= c0 < d0 and 0 <= c1 < d1 else -1
        bufs = {}
        for r, R in enumerate(w.ranks):
            for dr in range(12):
                if not (which & (2 if dr >= 8 else 1)):
                    continue
                d = dr - 8 if dr >= 8 else dr
                ox, oy = ((d == 0) - (d == 1), (d == 2) - (d == 3)) if d < 4 else (EX9[d + 1], EY9[d + 1])
                to = rank(R.coords[0] + ox, R.coords[1] + oy)
                if to < 0:
                    continue
                n1 = (R.n[1] if d < 2 else R.n[0]) if d < 4 else 1
                npop = (1 if dr >= 8 else 3) if d < 4 else 1
                buf = np.full(n1 * npop, np.nan)
                self.S.shim_sub_pack(self.subs[r], dr, n1, npop, buf.ctypes.data_as(dp))
                bufs[(to, dr)] = (buf, n1, npop)
        for (to, dr), (buf, n1, npop) in bufs.items():
            self.S.shim_sub_unpack(self.subs[to], dr, n1, npop, buf.ctypes.data_as(dp))


@pytest.mark.parametrize("dims,bcT,walls", [c + ({},) for c in CASES + [((2, 3), (2, 1, 1, 2))]] + [c + (SHEAR,) for c in SHEARED_CASES])
def test_exact_kernel_sources_follow_the_oracle_subroutine_by_subroutine(shim, dims, bcT, walls):
    total = (23, 19)
    w = world_for(total, dims, bcT, Rayleigh=1e6, Thot=0.75, Tcold=-0.25, **walls)
    sw = ShimWorld(shim, w)
    P = range(w.nprocs)

    def same(*names):
        for r in P:
            for k in names:
                assert np.array_equal(sw.get(r, k), getattr(w.ranks[r], k)), (k, r)
    w.initial(); sw.initial()
    same("f", "g", "f_post", "g_post", "rho", "u", "v", "T", "up", "vp", "Tp")
    # a seeded, fully non-trivial state (the collisions themselves are covered above)
    rng = np.random.default_rng(11)
    for r, R in enumerate(w.ranks):
        for k in ("f", "g", "f_post", "g_post", "Fx", "Fy", "rho"):
            a = getattr(R, k)
            a[...] = rng.random(a.shape) * (1e-3 if k in ("Fx", "Fy") else 1.0) + (0.5 if k == "rho" else 0.0)
            sw.put(r, k, a)
    same("f", "g", "f_post", "g_post", "Fx", "Fy", "rho")                 # the transposing upload / download kernels round-trip
    w.message_passing_f(); sw.exchange(1)
    same("f_post", "g_post")
    w.streaming(); sw.op(1)
    same("f")
    w.bounceback(); sw.op(3)
    same("f")
    w.message_passing_g(); sw.exchange(2)
    same("g_post", "f_post")
    w.streamingT(); sw.op(2)
    same("g")
    w.bouncebackT(); sw.op(4)
    same("g")
    w.macro(); sw.op(5); w.macroT(); sw.op(6)
    same("rho", "u", "v", "T")
    # both message sets in one exchange (what the fused step does)
    for r, R in enumerate(w.ranks):
        R.f_post[...] = rng.random(R.f_post.shape); R.g_post[...] = rng.random(R.g_post.shape)
        sw.put(r, "f_post", R.f_post); sw.put(r, "g_post", R.g_post)
    w.message_passing_f(); w.message_passing_g(); sw.exchange(3)
    same("f_post", "g_post")
    sw.close(); w.close()
