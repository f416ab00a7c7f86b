"""CPU-only check of the product's AA-pattern (single-lattice) D3Q19 KERNEL SOURCE and launch schedule against the oracle:
tests/host_shim/aa_host.cpp compiles mglc_b200/csrc/lbm_aa_kernels.inl + lbm_aa_exact.inl + aa_run() (lbm_aa.cuh) for the host
and sweeps (blockIdx, threadIdx) sequentially -- exact for these kernels.  Covers the in-place pull/push addressing, the wall
rule on both sides of an odd launch, both lid terms, every way a run can start and end (either layout), download of f from
either layout, the BGK operator and both arithmetic builds, before any GPU time is spent.  The GPU tests proper are
tests/test_aa_gpu.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "aa_host.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", "-ffp-contract=off", "-Wl,-Bsymbolic", "-I/usr/local/cuda/include",
                           "-o", out, os.path.join(ROOT, "tests", "host_shim", "aa_host.cpp")])
    S = C.CDLL(out)       # -Bsymbolic: libmglc.so (RTLD_GLOBAL) exports host stubs with the kernels' names; bind to the shim's own
    S.aa_shim_create.restype = C.c_void_p
    S.aa_shim_create.argtypes = [C.c_int] * 3 + [C.c_double] * 4 + [C.c_int] * 2
    S.aa_shim_destroy.argtypes = [C.c_void_p]
    S.aa_shim_upload.argtypes = [C.c_void_p] + [dp] * 5
    S.aa_shim_layout.argtypes = [C.c_void_p]
    S.aa_shim_step.restype = C.c_longlong
    S.aa_shim_step.argtypes = [C.c_void_p, C.c_int]
    S.aa_shim_download_macro.argtypes = [C.c_void_p] + [dp] * 4
    S.aa_shim_download_f.argtypes = [C.c_void_p, dp, C.c_longlong]
    S.aa_shim_create_block.restype = C.c_void_p
    S.aa_shim_create_block.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int), C.c_int] + [C.c_double] * 4 + [C.c_int] * 2
    S.aa_shim_set_peers.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    S.aa_shim_world_step.restype = C.c_longlong
    S.aa_shim_world_step.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]
    return S


class AaSim:
    def __init__(self, S, wd, strict, bgk=False):
        self.S, self.total = S, wd.total
        self.h = S.aa_shim_create(*wd.total, wd.Snu, wd.Sq, wd.U0, wd.rho0, int(bgk), int(strict))

    def upload(self, wd):
        R = wd.ranks[0]
        a = [np.asfortranarray(x) for x in (R.f, R.rho, R.u, R.v, R.w)]
        self.S.aa_shim_upload(self.h, *[x.ctypes.data_as(dp) for x in a])

    def step(self, n):
        return self.S.aa_shim_step(self.h, n)

    def layout(self):
        return self.S.aa_shim_layout(self.h)

    def macro(self):
        out = [np.empty(self.total, order="F") for _ in range(4)]
        self.S.aa_shim_download_macro(self.h, *[x.ctypes.data_as(dp) for x in out])
        return dict(zip(("rho", "u", "v", "w"), out))

    def f(self, chunk=97):
        out = np.empty((19,) + self.total, order="F")
        self.S.aa_shim_download_f(self.h, out.ctypes.data_as(dp), chunk)
        return out

    def close(self):
        self.S.aa_shim_destroy(self.h)


def seeded(wd, seed):
    """perturbed populations and fields that are NOT their moments: the first collision() must use the stored rho,u,v,w"""
    rng = np.random.default_rng(seed)
    R = wd.ranks[0]
    R.f[...] *= 1.0 + 0.05 * rng.uniform(-1, 1, R.f.shape)
    R.rho[...] = 1.0 + 0.02 * rng.uniform(-1, 1, R.rho.shape)
    for k in ("u", "v", "w"):
        getattr(R, k)[...] = 0.05 * rng.uniform(-1, 1, R.rho.shape)


SCHEDULES = [[1], [2], [3], [4], [1, 1, 1, 1], [2, 1, 2], [3, 3], [5, 2, 1], [1, 4, 1]]


@pytest.mark.parametrize("calls", SCHEDULES)
@pytest.mark.parametrize("collision", ["mrt", "bgk"])
def test_strict_build_is_bit_exact_for_every_way_a_run_starts_and_ends(shim, calls, collision):
    total = (9, 8, 7)
    wd = orc.LidWorld(total, 1, collision=collision)
    wd.initial()
    seeded(wd, 5)
    sim = AaSim(shim, wd, strict=True, bgk=collision == "bgk")
    sim.upload(wd)
    for n in calls:
        wd.step(n)
        launches = sim.step(n)
        assert launches >= n + 1
        m = sim.macro()
        for k in ("rho", "u", "v", "w"):
            assert np.array_equal(m[k], wd.gather(k)), (calls, n, k)
        assert np.array_equal(sim.f(), wd.gather("f")), (calls, n, "f", sim.layout())
    sim.close(); wd.close()


def test_layouts_alternate_as_documented(shim):
    wd = orc.LidWorld((5, 4, 6), 1)
    wd.initial()
    sim = AaSim(shim, wd, strict=True)
    sim.upload(wd)
    assert sim.layout() == 0
    assert sim.step(1) == 3 and sim.layout() == 1          # lid plane + collision + fields through the pull
    assert sim.step(1) == 2 and sim.layout() == 0          # odd launch + macro
    assert sim.step(2) == 4 and sim.layout() == 0          # lid plane + collision + odd + macro
    assert sim.step(3) == 5 and sim.layout() == 1          # lid plane + collision + odd + even + fields through the pull
    assert sim.step(4) == 5 and sim.layout() == 1          # odd + even + odd + even + fields through the pull
    sim.close(); wd.close()


def test_lid_driven_start_from_initial_strict(shim):
    """the reference's own start (rest fluid, moving lid): both lid terms of an odd launch matter from the first step"""
    total = (11, 10, 9)
    wd = orc.LidWorld(total, 1)
    wd.initial()
    sim = AaSim(shim, wd, strict=True)
    sim.upload(wd)
    for n in (1, 6, 13):
        wd.step(n); sim.step(n)
        m = sim.macro()
        for k in ("rho", "u", "v", "w"):
            assert np.array_equal(m[k], wd.gather(k)), (n, k)
    assert np.array_equal(sim.f(chunk=1000), wd.gather("f"))
    sim.close(); wd.close()


def test_fast_build_tracks_the_oracle(shim):
    total = (17, 16, 15)
    wd = orc.LidWorld(total, 1)
    wd.initial()
    sim = AaSim(shim, wd, strict=False)
    sim.upload(wd)
    for n in (1, 10, 89):
        wd.step(n); sim.step(n)
        m = sim.macro()
        for k in ("rho", "u", "v", "w"):
            a, b = m[k], wd.gather(k)
            rel = np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)
            assert rel <= 1e-12 and np.abs(a - b).max() <= 1e-10, (n, k, rel)
    sim.close(); wd.close()


# ---- decomposed lattices: the PEER build of the same kernels, blocks wired to each other's lattices ---------------------------
# direction d of a message (ex_sendrecv.f90:12-123): faces 0..5 = +x,-x,+y,-y,+z,-z, edges 7..18 = the population crossing it
EX = [0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0]
EY = [0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1]
EZ = [0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1]
FACE = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]


class AaWorld:
    """the blocks of the oracle world wd (nprocs > 1, same decomposition), each an in-place lattice of the host shim"""

    def __init__(self, S, wd, strict, bgk=False):
        self.S, self.wd = S, wd
        self.dims = tuple(wd.dims)
        self.subs = []
        for R in wd.ranks:
            co = R.coords
            wall = (C.c_int * 6)(*[int(co[q // 2] == (self.dims[q // 2] - 1 if q % 2 == 0 else 0)) for q in range(6)])
            self.subs.append(S.aa_shim_create_block(*R.n, wall, wall[4], wd.Snu, wd.Sq, wd.U0, wd.rho0, int(bgk), int(strict)))
        rank_of = {tuple(R.coords): r for r, R in enumerate(wd.ranks)}
        for r, R in enumerate(wd.ranks):
            nbr = (C.c_void_p * 19)()
            for d in range(19):
                if d == 6:
                    continue
                e = FACE[d] if d < 6 else (EX[d], EY[d], EZ[d])
                other = rank_of.get(tuple(c + x for c, x in zip(R.coords, e)))
                nbr[d] = self.subs[other] if other is not None else None
            S.aa_shim_set_peers(self.subs[r], nbr)
        self.arr = (C.c_void_p * len(self.subs))(*self.subs)

    def upload(self):
        for h, R in zip(self.subs, self.wd.ranks):
            a = [np.asfortranarray(x) for x in (R.f, R.rho, R.u, R.v, R.w)]
            self.S.aa_shim_upload(h, *[x.ctypes.data_as(dp) for x in a])

    def step(self, n, order=0):
        return self.S.aa_shim_world_step(self.arr, len(self.subs), n, order)

    def gather(self, what, chunk=97):
        lead = (19,) if what == "f" else ()
        glob = np.empty(lead + tuple(self.wd.total), order="F")
        for h, R in zip(self.subs, self.wd.ranks):
            n = tuple(R.n)
            if what == "f":
                out = np.empty((19,) + n, order="F")
                self.S.aa_shim_download_f(h, out.ctypes.data_as(dp), chunk)
            else:
                m = [np.empty(n, order="F") for _ in range(4)]
                self.S.aa_shim_download_macro(h, *[x.ctypes.data_as(dp) for x in m])
                out = m[("rho", "u", "v", "w").index(what)]
            sl = tuple(slice(s, s + k) for s, k in zip(R.start, n))
            glob[(slice(None),) * len(lead) + sl] = out
        return glob

    def close(self):
        for h in self.subs:
            self.S.aa_shim_destroy(h)


def seeded_world(wd, seed):
    rng = np.random.default_rng(seed)
    for R in wd.ranks:
        R.f[...] *= 1.0 + 0.05 * rng.uniform(-1, 1, R.f.shape)
        R.rho[...] = 1.0 + 0.02 * rng.uniform(-1, 1, R.rho.shape)
        for k in ("u", "v", "w"):
            getattr(R, k)[...] = 0.05 * rng.uniform(-1, 1, R.rho.shape)


@pytest.mark.parametrize("nprocs,total", [(2, (7, 6, 9)), (4, (7, 9, 8)), (8, (9, 8, 7)), (12, (9, 7, 10)), (3, (5, 4, 3))])
@pytest.mark.parametrize("order", [0, 1])
def test_decomposed_blocks_store_into_each_other_bit_exact(shim, nprocs, total, order):
    """P blocks (uneven, down to one cell thick), every launch storing into the neighbours: equal to the single-rank oracle
    whichever block of a launch runs first -- nothing a launch writes into a neighbour is read over there in the same launch"""
    ref = orc.LidWorld(total, 1)
    ref.initial()
    wd = orc.LidWorld(total, nprocs)            # only its decomposition and its state at upload time are used
    wd.initial()
    seeded_world(wd, 11)
    for k in ("rho", "u", "v", "w", "f"):       # the same perturbed state on the single-rank oracle
        getattr(ref.ranks[0], k)[...] = wd.gather(k)
    sim = AaWorld(shim, wd, strict=True)
    sim.upload()
    for n in (1, 2, 3, 5):
        ref.step(n)
        sim.step(n, order)
        for k in ("rho", "u", "v", "w", "f"):
            assert np.array_equal(sim.gather(k), ref.gather(k)), (nprocs, n, k)
    sim.close(); wd.close(); ref.close()


def test_decomposed_blocks_bgk_and_fast_build(shim):
    total = (10, 9, 8)
    for collision, strict in (("bgk", True), ("mrt", False)):
        ref = orc.LidWorld(total, 1, collision=collision)
        ref.initial()
        wd = orc.LidWorld(total, 4, collision=collision)
        wd.initial()
        sim = AaWorld(shim, wd, strict=strict, bgk=collision == "bgk")
        sim.upload()
        ref.step(21); sim.step(21)
        for k in ("rho", "u", "v", "w"):
            a, b = sim.gather(k), ref.gather(k)
            if strict:
                assert np.array_equal(a, b), (collision, k)
            else:
                assert np.abs(a - b).max() <= 1e-13, (collision, k)
        sim.close(); wd.close(); ref.close()


@pytest.mark.parametrize("nx", [31, 32, 33, 64, 65, 70, 129])
def test_rows_of_one_to_five_warps_strict(shim, nx):
    """the odd launch picks its path per warp from the warp's first x index: rows where the first and the last cell share a warp,
    sit in neighbouring warps, or have interior warps between them; one block and two blocks split along x"""
    total = (nx, 4, 3)
    wd = orc.LidWorld(total, 1)
    wd.initial()
    seeded(wd, nx)
    sim = AaSim(shim, wd, strict=True)
    sim.upload(wd)
    w2 = orc.LidWorld(total, 2, dims=(2, 1, 1))
    w2.initial()
    for k in ("rho", "u", "v", "w", "f"):
        full = wd.gather(k)
        for R in w2.ranks:
            sl = tuple(slice(s, s + n) for s, n in zip(R.start, R.n))
            getattr(R, k)[...] = full[(slice(None),) + sl] if k == "f" else full[sl]
    two = AaWorld(shim, w2, strict=True)
    two.upload()
    for n in (2, 3):
        wd.step(n); sim.step(n); two.step(n)
        for k in ("rho", "u", "v", "w"):
            assert np.array_equal(sim.macro()[k], wd.gather(k)), (nx, n, k)
            assert np.array_equal(two.gather(k), wd.gather(k)), (nx, n, k, "two blocks")
        assert np.array_equal(sim.f(), wd.gather("f")), (nx, n)
        assert np.array_equal(two.gather("f"), wd.gather("f")), (nx, n, "two blocks")
    sim.close(); two.close(); wd.close(); w2.close()


# ---- randomised decompositions (hypothesis): the park / push protocol must hold for any block grid, any remainder, both orders ----
try:
    from hypothesis import HealthCheck, given, settings, strategies as st
    HAVE_HYPOTHESIS = True
except Exception:                                   # pragma: no cover
    HAVE_HYPOTHESIS = False


if HAVE_HYPOTHESIS:
    @st.composite
    def block_grids(draw):
        dims = tuple(draw(st.integers(1, 3)) for _ in range(3))
        total = tuple(draw(st.integers(d, d + 5)) for d in dims)          # at least one cell per block, uneven remainders
        return dims, total

    @settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
    @given(grid=block_grids(), order=st.integers(0, 1), steps=st.lists(st.integers(1, 4), min_size=1, max_size=3),
           bgk=st.booleans(), seed=st.integers(0, 10**6))
    def test_any_block_grid_matches_the_single_rank_oracle(shim, grid, order, steps, bgk, seed):
        dims, total = grid
        nprocs = dims[0] * dims[1] * dims[2]
        collision = "bgk" if bgk else "mrt"
        ref = orc.LidWorld(total, 1, collision=collision)
        ref.initial()
        wd = orc.LidWorld(total, nprocs, dims=dims, collision=collision)
        wd.initial()
        seeded_world(wd, seed)
        for k in ("rho", "u", "v", "w", "f"):
            getattr(ref.ranks[0], k)[...] = wd.gather(k)
        sim = AaWorld(shim, wd, strict=True, bgk=bgk)
        sim.upload()
        for n in steps:
            ref.step(n); sim.step(n, order)
        for k in ("rho", "u", "v", "w", "f"):
            assert np.array_equal(sim.gather(k), ref.gather(k)), (dims, total, steps, k)
        sim.close(); wd.close(); ref.close()
