"""GPU parity tests of the INCOMPRESSIBLE 2-D lid-driven cavity (variant "i" of mglc_l2d_*: the reference's sequential program
MPI/Lid_driven_cavity/fortran/2d/seq/lid-driven_cavity_incompress.f90) against the CPU oracle (oracle/lid2d.c, variant L2_I,
pinned bit for bit to that program's source text by test_oracle_lid2d.py) and against the committed outputs of the program's own
run (tests/golden/ref_fortran_lid2d_incomp.npz).  Strict arithmetic: bit-exact.  Fast arithmetic: <= 1e-12 relative L2,
<= 1e-10 max pointwise.  The same kernel source is checked on the CPU by test_lid2d_kernels_host.py (host shim).
(The file sorts last on purpose: it was added after the other GPU files and a failure here must not hide their results under -x.)"""
import os

import numpy as np
import pytest

import mglc_b200 as mg
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
IGOLD = np.load(os.path.join(HERE, "golden", "ref_fortran_lid2d_incomp.npz"))
REL_L2, MAX_ABS = 1e-12, 1e-10


def close_enough(got, want):
    d = np.linalg.norm((got - want).ravel()) / max(np.linalg.norm(want.ravel()), 1e-300)
    return d <= REL_L2 and np.abs(got - want).max() <= MAX_ABS


def velocity_close_enough(sim, wd):
    """the tolerance on the velocity VECTOR field: relative L2 of (u, v) jointly (after a few steps v alone is non-zero only in
    the two lid corners, so its own norm is not a scale of the flow), max pointwise on each component"""
    du, dv = sim.gather("u") - wd.gather("u"), sim.gather("v") - wd.gather("v")
    scale = np.sqrt((wd.gather("u") ** 2).sum() + (wd.gather("v") ** 2).sum())
    return np.sqrt((du ** 2).sum() + (dv ** 2).sum()) <= REL_L2 * scale and max(np.abs(du).max(), np.abs(dv).max()) <= MAX_ABS


def pair(total, nprocs=1, dims=None, strict=True, seed=None):
    wd = orc.Lid2DWorld(total, nprocs, dims, variant="i")
    sim = mg.LidDrivenCavity2D(total, nprocs=nprocs, dims=dims, variant="i", strict=strict)
    assert sim.dims == wd.dims and (sim.tauf, sim.Snu, sim.Sq) == (wd.tauf, wd.Snu, wd.Sq)
    wd.initial(); sim.initial()
    if seed is not None:                   # perturbed populations, fields independent of them: every moment non-trivial
        rng = np.random.default_rng(seed)
        f = np.asfortranarray(wd.gather("f") * (1.0 + 0.05 * rng.uniform(-1, 1, (9,) + tuple(total))))
        rho = np.asfortranarray(1.0 + 0.02 * rng.uniform(-1, 1, total))
        u, v = (np.asfortranarray(0.05 * rng.uniform(-1, 1, total)) for _ in range(2))
        for k, a in (("f", f), ("rho", rho), ("u", u), ("v", v)):
            wd.scatter(k, a); sim.scatter(k, a)
    return wd, sim


def assert_rank_arrays_equal(wd, sim, names, interior_only_fpost=False):
    for r, R in enumerate(wd.ranks):
        for k in names:
            got, want = sim.download(r, k), getattr(R, k)
            if k == "f_post" and interior_only_fpost:
                got, want = got[:, 1:-1, 1:-1], want[:, 1:-1, 1:-1]
            assert np.array_equal(got, want), (k, r)


def test_shipped_constants():
    sim = mg.LidDrivenCavity2D(variant="i")
    assert sim.total == (257, 257)                                        # inc:7
    tau = 0.1 * 257.0 / 1000.0 * 3.0 + 0.5                                # inc:11
    assert (sim.tauf, sim.Snu, sim.Sq) == (tau, 1.0 / tau, 8.0 * (2.0 * tau - 1.0) / (8.0 * tau - 1.0))
    sim.close()


@pytest.mark.parametrize("total", [(257, 257), (37, 5), (8, 7)])
def test_initial_bit_exact(total):
    """inc:137-163: rho = 0 (not rho0), u = U0 on the lid row, f = omega*(...)"""
    for nprocs in (1, 2, 4):
        wd, sim = pair(total, nprocs)
        assert_rank_arrays_equal(wd, sim, ("f", "rho", "u", "v"))
        assert not sim.gather("rho").any()
        wd.close(); sim.close()


def test_program_run_matches_the_reference_text_outputs():
    """strict GPU run of the program's loop (inc:60-66) vs the outputs of its own source text after 0, 1, 2 and 20 iterations,
    bit for bit, on 1 subdomain and on 2 x 2; check() after 20 and 25"""
    total = tuple(int(x) for x in IGOLD["shape"])
    for nprocs in (1, 4):
        sim = mg.LidDrivenCavity2D(total, nprocs=nprocs, variant="i", strict=True)
        assert (sim.tauf, sim.Snu, sim.Sq) == tuple(IGOLD["params"])
        sim.initial()
        assert np.array_equal(sim.gather("f"), IGOLD["run0/f"])
        done = 0
        for n in (1, 2, 20):
            sim.step(n - done); done = n
            assert np.array_equal(sim.gather("f"), IGOLD[f"run{n}/f"]), n
            assert np.array_equal(np.stack([sim.gather(k) for k in ("rho", "u", "v")]), IGOLD[f"run{n}/ruv"]), n
        assert np.isclose(sim.check(), IGOLD["run20/check"][2], rtol=1e-13, atol=0)
        sim.step(5)
        assert np.isclose(sim.check(), IGOLD["run25/check"][2], rtol=1e-13, atol=0)
        sim.close()


def test_subroutines_on_the_reference_text_arrays():
    """collision / streaming / bounceback / macro / check on the seeded arrays the reference's text was evaluated on"""
    total = tuple(int(x) for x in IGOLD["shape"])
    sim = mg.LidDrivenCavity2D(total, variant="i", strict=True)

    def load():
        sim.upload(0, f=IGOLD["in/f0"], f_post=IGOLD["in/f_post"], rho=IGOLD["in/rho"], u=IGOLD["in/u"], v=IGOLD["in/v"])

    load(); sim.collision()
    assert np.array_equal(sim.download(0, "f_post")[:, 1:-1, 1:-1], IGOLD["collision/f_post"])
    load(); sim.streaming()
    assert np.array_equal(sim.download(0, "f"), IGOLD["streaming/f"])
    load(); sim.bounceback()
    assert np.array_equal(sim.download(0, "f"), IGOLD["bounceback/f"])
    load(); sim.macro()
    assert np.array_equal(np.stack(sim.download(0, "rho", "u", "v")), IGOLD["macro/ruv"])
    sim.close()


@pytest.mark.parametrize("total,nprocs,dims", [((34, 33), 1, None), ((34, 33), 4, None), ((23, 19), 6, None), ((40, 7), 3, (3, 1)), ((9, 31), 3, (1, 3))])
def test_each_subroutine_bit_exact_strict(total, nprocs, dims):
    wd, sim = pair(total, nprocs, dims, strict=True, seed=3)
    for R in wd.ranks:                                 # the reference leaves wall halos uninitialised: make them recognisable
        R.f_post[...] = -7.25
    for r in range(nprocs):
        sim.upload(r, f_post=wd.ranks[r].f_post)
    for it in range(3):
        wd.collision(); sim.collision()
        assert_rank_arrays_equal(wd, sim, ("f_post",))
        wd.message_passing_sendrecv(); sim.message_passing_sendrecv()
        assert_rank_arrays_equal(wd, sim, ("f_post",))
        wd.streaming(); sim.streaming()
        assert_rank_arrays_equal(wd, sim, ("f",))
        wd.bounceback(); sim.bounceback()
        assert_rank_arrays_equal(wd, sim, ("f",))
        wd.macro(); sim.macro()
        assert_rank_arrays_equal(wd, sim, ("rho", "u", "v"))
    assert np.isclose(sim.check(), wd.check(), rtol=1e-13, atol=0)
    assert np.isclose(sim.check(), wd.check(), rtol=1e-13, atol=0)     # up, vp were refreshed identically
    wd.close(); sim.close()


@pytest.mark.parametrize("total,nprocs,dims", [((34, 33), 1, None), ((23, 19), 4, None), ((23, 19), 6, None), ((130, 6), 2, None), ((9, 31), 3, (1, 3))])
def test_fused_step_strict_is_bit_exact(total, nprocs, dims):
    """step(N) = the rotated loop (collision, N-1 fused launches, stream+macro), starting from the program's rho = 0"""
    wd, sim = pair(total, nprocs, dims, strict=True)
    for n in (1, 2, 17):
        wd.step(n); sim.step(n)
        assert_rank_arrays_equal(wd, sim, ("f", "rho", "u", "v"))
        assert_rank_arrays_equal(wd, sim, ("f_post",), interior_only_fpost=True)
    assert_rank_arrays_equal(wd, sim, ("f_post",))
    wd.collision(); sim.collision()
    wd.message_passing_sendrecv(); sim.message_passing_sendrecv()
    wd.streaming(); sim.streaming(); wd.bounceback(); sim.bounceback(); wd.macro(); sim.macro()
    wd.step(3); sim.step(3)
    assert_rank_arrays_equal(wd, sim, ("f", "f_post", "rho", "u", "v"))
    wd.close(); sim.close()


def test_graph_replayed_steps_strict_are_bit_exact():
    wd, sim = pair((45, 38), 1, strict=True)
    for n in (136, 65):
        wd.step(n); sim.step(n)
        assert_rank_arrays_equal(wd, sim, ("f", "rho", "u", "v"))
    wd.close(); sim.close()


def test_shipped_case_fast_within_tolerance():
    """the shipped 257 x 257 grid at Re = 1000, N in {1, 10, 100, 1000}, fast arithmetic; check() every 1000 steps as the program does"""
    wd, sim = pair((257, 257), 1, strict=False)
    done = 0
    for n in (1, 10, 100, 1000):
        wd.step(n - done); sim.step(n - done); done = n
        assert close_enough(sim.gather("rho"), wd.gather("rho")), n
        assert velocity_close_enough(sim, wd), n
    assert sim.check() == wd.check() == 1.0                     # up = vp = 0 before the first check(): error1 == error2
    wd.step(10); sim.step(10)
    assert np.isclose(sim.check(), wd.check(), rtol=1e-9, atol=0)
    wd.close(); sim.close()


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("nprocs,dims", [(2, None), (4, None), (6, None), (3, (1, 3))])
def test_decomposed_equals_single_subdomain_bit_for_bit(strict, nprocs, dims):
    total = (67, 45)
    one = mg.LidDrivenCavity2D(total, variant="i", strict=strict)
    many = mg.LidDrivenCavity2D(total, nprocs=nprocs, dims=dims, variant="i", strict=strict)
    one.initial(); many.initial()
    one.step(40); many.step(40)
    for k in ("f", "rho", "u", "v"):
        assert np.array_equal(one.gather(k), many.gather(k)), k
    assert np.isclose(one.check(), many.check(), rtol=1e-12)
    one.close(); many.close()


def test_large_lattice_properties():
    """2048 x 2048 (no oracle run): total mass is constant to rounding (the lid term adds -U0/6 + U0/6 per lid cell), cells
    farther than N rows from the lid stay at rest exactly (their populations change uniformly: the first collision() sees
    rho = 0), and 4 subdomains reproduce 1"""
    total, n = (2048, 2048), 12
    one = mg.LidDrivenCavity2D(total, variant="i", strict=False)
    many = mg.LidDrivenCavity2D(total, nprocs=4, variant="i", strict=False)
    one.initial(); many.initial()
    one.step(n); many.step(n)
    rho, u, v = (one.gather(k) for k in ("rho", "u", "v"))
    assert abs(rho.sum() / rho.size - 1.0) < 1e-13
    far = slice(0, total[1] - n - 1)
    assert np.all(u[:, far] == 0.0) and np.all(v[:, far] == 0.0) and np.abs(rho[:, far] - 1.0).max() < 1e-14
    assert np.abs(u[:, -1]).max() > 0.0
    for k, a in (("rho", rho), ("u", u), ("v", v)):
        assert np.array_equal(many.gather(k), a), k
    one.close(); many.close()
