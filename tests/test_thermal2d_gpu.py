"""GPU parity tests of the 2-D thermal D2Q9 + D2Q5 path (libmglc.so through the C ABI) against the CPU oracle (oracle/thermal2d.c,
pinned bit for bit to the reference's Fortran text by test_oracle_thermal2d.py) and against the committed evaluations of the
reference's own source (tests/golden/ref_fortran_thermal2d.npz).
Strict arithmetic: everything bit-exact.  Fast arithmetic: copy-type subroutines, macro()/macroT() and the stored force
bit-exact, collision()/collisionT()/step() to the north-star tolerance (<= 1e-12 relative L2, <= 1e-10 max pointwise)."""
import os

import numpy as np
import pytest

import mglc_b200 as mg
from mglc_b200.thermal2d import PARAM_NAMES, RAYLEIGH_BENARD, RB_PERIODIC, SIDE_HEATED
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "ref_fortran_thermal2d.npz"))
REL_L2, MAX_ABS = 1e-12, 1e-10
FIELDS = ("rho", "u", "v", "T")


def close_enough(got, want, floor=0.0):
    d = np.linalg.norm((got - want).ravel()) / max(np.linalg.norm(want.ravel()), floor, 1e-300)
    return d <= REL_L2 and np.abs(got - want).max() <= MAX_ABS


def seeded_state(total, seed):
    """f, g = perturbed equilibria; rho/u/v/T independent of them: every moment non-trivial"""
    rng = np.random.default_rng(seed)
    wd = orc.Thermal2DWorld(total, 1)
    wd.initial()
    f = np.asfortranarray(wd.gather("f") * (1.0 + 0.05 * rng.uniform(-1, 1, (9,) + tuple(total))))
    g = np.asfortranarray(0.2 * (0.5 + rng.uniform(-0.3, 0.3, (5,) + tuple(total))))
    wd.close()
    out = dict(f=f, g=g, rho=1.0 + 0.02 * rng.uniform(-1, 1, total), T=rng.uniform(-0.1, 1.1, total))
    out["u"], out["v"] = (0.05 * rng.uniform(-1, 1, total) for _ in range(2))
    out["Fx"], out["Fy"] = np.zeros(total), 1e-4 * rng.uniform(-1, 1, total)
    return {k: np.asfortranarray(a) for k, a in out.items()}


def pair(total, nprocs=1, dims=None, bcT=None, strict=True, seed=None, variant="mpi", **params):
    lu = float(total[0]) if variant == "acc" else 0.0                 # the OpenACC program's lengthUnit = dble(nx), acc:57
    wd = orc.Thermal2DWorld(total, nprocs, dims, bcT=bcT, variant=variant, lengthUnit=lu, **params)
    sim = mg.BuoyancyDrivenCavity2D(total, nprocs=nprocs, dims=dims, bcT=bcT, strict=strict, variant=variant, **params)
    assert sim.dims == wd.dims and sim.bcT == wd.bcT
    for k in PARAM_NAMES:
        assert sim.params[k] == getattr(wd.params, k), k
    for r, R in enumerate(wd.ranks):
        inf = sim.info[r]
        assert inf["n"] == R.n and inf["start"] == R.start and inf["coords"] == R.coords and inf["nbr"] == R.nbr + R.cnr
    if seed is None:
        wd.initial(); sim.initial()
    else:
        for k, a in seeded_state(total, seed).items():
            wd.scatter(k, a); sim.scatter(k, a)
    return wd, sim


def assert_rank_arrays_equal(wd, sim, names, interior_only_post=False):
    for r, R in enumerate(wd.ranks):
        for k in names:
            got, want = sim.download(r, k), getattr(R, k)
            if k in ("f_post", "g_post") and interior_only_post:
                got, want = got[:, 1:-1, 1:-1], want[:, 1:-1, 1:-1]
            assert np.array_equal(got, want), (k, r)


def test_parameters_match_the_fortran_text():
    sim = mg.BuoyancyDrivenCavity2D()
    assert sim.total == (201, 201) and sim.bcT == SIDE_HEATED
    assert tuple(sim.params[k] for k in PARAM_NAMES[:9]) == tuple(GOLD["params/201"])
    sim.close()
    sim = mg.BuoyancyDrivenCavity2D((64, 64), Rayleigh=1e6)
    assert tuple(sim.params[k] for k in PARAM_NAMES[:9]) == tuple(GOLD["params/64_ra1e6"])
    sim.close()


@pytest.mark.parametrize("total,bcT", [((201, 201), SIDE_HEATED), ((37, 5), RAYLEIGH_BENARD), ((130, 2), SIDE_HEATED), ((2, 9), (0, 0, 0, 0))])
def test_initial_bit_exact(total, bcT):
    for nprocs in (1, 2, 4):
        if min(total) < 4 and nprocs == 4:
            continue
        wd, sim = pair(total, nprocs, bcT=bcT, Thot=0.75, Tcold=-0.25)
        assert_rank_arrays_equal(wd, sim, ("f", "g", "f_post", "g_post") + FIELDS)
        wd.close(); sim.close()


def test_collisions_on_the_golden_cells_of_the_reference():
    """the 48 seeded cells whose f_post, g_post, Fx, Fy the reference's own source text produced"""
    f, g, r = GOLD["cells/f"], GOLD["cells/g"], GOLD["cells/ruvT"]
    n, nx, ny = len(f), 67, 201                                   # total_ny = 201: the shipped tauf, paraA, gBeta
    pad = np.arange(nx * ny) % n
    shp = lambda a: np.asfortranarray(a[pad].reshape(nx, ny, order="F"))
    for strict in (True, False):
        sim = mg.BuoyancyDrivenCavity2D((nx, ny), strict=strict)
        sim.upload(0, f=np.asfortranarray(f[pad].T.reshape(9, nx, ny, order="F")), g=np.asfortranarray(g[pad].T.reshape(5, nx, ny, order="F")),
                   rho=shp(r[:, 0]), u=shp(r[:, 1]), v=shp(r[:, 2]), T=shp(r[:, 3]))
        sim.collision(); sim.collisionT()
        fp = sim.download(0, "f_post")[:, 1:-1, 1:-1].reshape(9, -1, order="F").T
        gp = sim.download(0, "g_post")[:, 1:-1, 1:-1].reshape(5, -1, order="F").T
        Fx, Fy = (sim.download(0, k).ravel(order="F") for k in ("Fx", "Fy"))
        assert np.array_equal(Fx, GOLD["collision/FxFy"][pad, 0]) and np.array_equal(Fy, GOLD["collision/FxFy"][pad, 1])
        if strict:
            assert np.array_equal(fp, GOLD["collision/f_post"][pad]) and np.array_equal(gp, GOLD["collisionT/g_post"][pad])
        else:
            assert close_enough(fp, GOLD["collision/f_post"][pad]) and close_enough(gp, GOLD["collisionT/g_post"][pad])
        # macro() / macroT() on the same cells with the golden forces: bit-exact in both builds, including signed zeros
        F2 = GOLD["cells/FxFy"]
        sim.upload(0, Fx=shp(F2[:, 0]), Fy=shp(F2[:, 1]))
        sim.macro(); sim.macroT()
        got = np.stack([sim.download(0, k).ravel(order="F") for k in ("rho", "u", "v")], axis=1)
        assert np.array_equal(got, GOLD["macro/ruv"][pad]) and np.array_equal(np.signbit(got), np.signbit(GOLD["macro/ruv"][pad]))
        assert np.array_equal(sim.download(0, "T").ravel(order="F"), GOLD["macro/T"][pad])
        sim.close()


def test_copy_subroutines_on_the_golden_block_of_the_reference():
    """streaming / bounceback / streamingT / bouncebackT on the 6 x 5 block the reference's source text was evaluated on"""
    nx, ny = 6, 5
    for case in range(5):
        c = GOLD["field/bb_cases"][case]
        coords, dims = tuple(int(x) for x in c[:2]), tuple(int(x) for x in c[2:])
        r = coords[0] * dims[1] + coords[1]
        for tag, bcT in (("side", SIDE_HEATED), ("rb", RAYLEIGH_BENARD)):
            sim = mg.BuoyancyDrivenCavity2D((nx * dims[0], ny * dims[1]), nprocs=dims[0] * dims[1], dims=dims, bcT=bcT, strict=True)
            assert sim.info[r]["n"] == (nx, ny) and sim.info[r]["coords"] == coords
            sim.upload(r, f_post=GOLD["field/f_post"], g_post=GOLD["field/g_post"])
            if case == 0:
                sim.streaming(); sim.streamingT()
                assert np.array_equal(sim.download(r, "f"), GOLD["field/streaming_f"])
                assert np.array_equal(sim.download(r, "g"), GOLD["field/streamingT_g"])
            sim.upload(r, f=GOLD["field/f0"], g=GOLD["field/g0"])
            sim.bounceback(); sim.bouncebackT()
            assert np.array_equal(sim.download(r, "f"), GOLD[f"field/bounceback_{case}"]), (case, tag)
            assert np.array_equal(sim.download(r, "g"), GOLD[f"field/bouncebackT_{tag}_{case}"]), (case, tag)
            sim.close()


@pytest.mark.parametrize("bcT", [SIDE_HEATED, RAYLEIGH_BENARD, (2, 1, 1, 2)])
@pytest.mark.parametrize("total,nprocs,dims", [((34, 33), 1, None), ((34, 33), 4, None), ((23, 19), 6, None), ((40, 7), 3, (3, 1)), ((9, 31), 3, (1, 3))])
def test_each_subroutine_bit_exact_strict(bcT, total, nprocs, dims):
    wd, sim = pair(total, nprocs, dims, bcT=bcT, strict=True, seed=3, Rayleigh=1e6)
    for R in wd.ranks:                                 # make never-written halo entries recognisable
        R.f_post[...] = -7.25; R.g_post[...] = 3.5
    for r in range(nprocs):
        sim.upload(r, f_post=wd.ranks[r].f_post, g_post=wd.ranks[r].g_post)
    for it in range(3):
        wd.collision(); sim.collision()
        assert_rank_arrays_equal(wd, sim, ("f_post", "Fx", "Fy"))
        wd.message_passing_f(); sim.message_passing_f()
        assert_rank_arrays_equal(wd, sim, ("f_post", "g_post"))
        wd.streaming(); sim.streaming()
        assert_rank_arrays_equal(wd, sim, ("f",))
        wd.bounceback(); sim.bounceback()
        assert_rank_arrays_equal(wd, sim, ("f",))
        wd.collisionT(); sim.collisionT()
        assert_rank_arrays_equal(wd, sim, ("g_post",))
        wd.message_passing_g(); sim.message_passing_g()
        assert_rank_arrays_equal(wd, sim, ("g_post", "f_post"))
        wd.streamingT(); sim.streamingT()
        assert_rank_arrays_equal(wd, sim, ("g",))
        wd.bouncebackT(); sim.bouncebackT()
        assert_rank_arrays_equal(wd, sim, ("g",))
        wd.macro(); sim.macro(); wd.macroT(); sim.macroT()
        assert_rank_arrays_equal(wd, sim, FIELDS)
    for _ in range(2):                                  # the second call checks that up, vp, Tp were refreshed identically
        a, b = sim.check(), wd.check()
        assert np.isclose(a[0], b[0], rtol=1e-13, atol=0) and np.isclose(a[1], b[1], rtol=1e-13, atol=0)
    assert np.allclose(sim.calNuRe()[1:], nure_of(wd)[1:], rtol=1e-12, atol=0)
    assert abs(sim.calNuRe()[0] - nure_of(wd)[0]) < 1e-12
    wd.close(); sim.close()


def nure_of(wd):
    """calNuRe()'s averages from the oracle's sums, NuRe.F90:31,41,62"""
    a, b, c = wd.nure_sums()
    N = float(wd.total[0] * wd.total[1])
    p = wd.params
    return a / N, b / N * p.lengthUnit / p.diffusivity + 1.0, np.sqrt(c / N) * p.lengthUnit / p.viscosity


@pytest.mark.parametrize("total,nprocs", [((34, 33), 1), ((34, 33), 4)])
def test_copy_type_subroutines_bit_exact_on_random_lattices(total, nprocs):
    """exchange / streaming / bounceback move doubles: `==` on every value, random f_post and g_post including the halos"""
    wd, sim = pair(total, nprocs, strict=False, seed=4)
    rng = np.random.default_rng(7)
    for r, R in enumerate(wd.ranks):
        R.f_post[...] = rng.random(R.f_post.shape); R.g_post[...] = rng.random(R.g_post.shape)
        sim.upload(r, f_post=R.f_post, g_post=R.g_post)
    wd.message_passing_f(); sim.message_passing_f(); wd.message_passing_g(); sim.message_passing_g()
    assert_rank_arrays_equal(wd, sim, ("f_post", "g_post"))
    wd.streaming(); sim.streaming(); wd.streamingT(); sim.streamingT()
    assert_rank_arrays_equal(wd, sim, ("f", "g"))
    wd.bounceback(); sim.bounceback(); wd.bouncebackT(); sim.bouncebackT()
    assert_rank_arrays_equal(wd, sim, ("f", "g"))
    wd.macro(); sim.macro(); wd.macroT(); sim.macroT()
    assert_rank_arrays_equal(wd, sim, FIELDS)
    wd.close(); sim.close()


@pytest.mark.parametrize("bcT", [SIDE_HEATED, RAYLEIGH_BENARD])
@pytest.mark.parametrize("total,nprocs,dims", [((34, 33), 1, None), ((23, 19), 4, None), ((23, 19), 6, None), ((130, 6), 2, None), ((9, 31), 3, (1, 3))])
def test_fused_step_strict_is_bit_exact(bcT, total, nprocs, dims):
    """step(N) = the rotated loop (2 collisions, N-1 fused launches, stream+macro): every array the reference holds afterwards"""
    wd, sim = pair(total, nprocs, dims, bcT=bcT, strict=True, Rayleigh=1e6)
    for n in (1, 2, 17):
        wd.step(n); sim.step(n)
        assert_rank_arrays_equal(wd, sim, ("f", "g", "Fx", "Fy") + FIELDS)
        assert_rank_arrays_equal(wd, sim, ("f_post", "g_post"), interior_only_post=True)
    # the exchanged halo entries too (wall halos are never written by either side: zero in both)
    assert_rank_arrays_equal(wd, sim, ("f_post", "g_post"))
    # calls compose with the per-subroutine entry points
    wd.collision(); sim.collision(); wd.message_passing_f(); sim.message_passing_f()
    wd.streaming(); sim.streaming(); wd.bounceback(); sim.bounceback()
    wd.collisionT(); sim.collisionT(); wd.message_passing_g(); sim.message_passing_g()
    wd.streamingT(); sim.streamingT(); wd.bouncebackT(); sim.bouncebackT()
    wd.macro(); sim.macro(); wd.macroT(); sim.macroT()
    wd.step(3); sim.step(3)
    assert_rank_arrays_equal(wd, sim, ("f", "g", "f_post", "g_post", "Fx", "Fy") + FIELDS)
    wd.close(); sim.close()


@pytest.mark.parametrize("variant,bcT", [("mpi", SIDE_HEATED), ("acc", RB_PERIODIC)])
def test_graph_replayed_steps_strict_are_bit_exact(variant, bcT):
    """a single small subdomain replays its fused launches as CUDA graphs of 64 kernels: 1 + 2 x 64 + 7 steps, then 64 + 1 more
    (the second call starts from the other ping-pong index), bit for bit"""
    wd, sim = pair((45, 38), 1, bcT=bcT, strict=True, variant=variant, Rayleigh=1e5)
    for n in (136, 65):
        l0 = sim.launch_count()
        wd.step(n); sim.step(n)
        assert sim.launch_count() - l0 == 2 + (n - 1) + 1          # kernels, whether launched directly or from a graph
        assert_rank_arrays_equal(wd, sim, ("f", "g", "Fx", "Fy") + FIELDS)
    wd.close(); sim.close()


def test_wall_halos_are_never_read_by_the_fused_step():
    wd, sim = pair((31, 17), 4, strict=True)
    for r in range(4):
        sim.upload(r, f_post=np.full(sim._shape(r, "f_post"), np.nan), g_post=np.full(sim._shape(r, "g_post"), np.nan))
    wd.step(20); sim.step(20)
    assert_rank_arrays_equal(wd, sim, ("f", "g") + FIELDS)
    wd.close(); sim.close()


def velocity_floor(wd):
    """The cavity starts from rest, so u and v are tiny at first while the populations they are differenced from are O(0.1):
    reordered arithmetic leaves an ABSOLUTE floor of ~1e-15 on them (measured on the CPU build of the same kernel source:
    7e-16 max after 10 steps, when |u| ~ 1e-6).  Their relative L2 is therefore taken against the larger of the field's own
    norm and the norm of a field at the reference's velocity unit sqrt(gBeta*L0*DeltaT) (module.F90:78) -- the same rule
    as the 3-D thermal test."""
    p = wd.params
    return np.sqrt(p.gBeta * p.lengthUnit) * np.sqrt(wd.total[0] * wd.total[1])


@pytest.mark.parametrize("bcT,Ra", [(SIDE_HEATED, 1e7), (RAYLEIGH_BENARD, 1e6)])
def test_shipped_case_fast_within_tolerance(bcT, Ra):
    """the shipped 201 x 201 grid (Ra = 1e7 side-heated; the Rayleigh-Benard macro set), N in {1, 10, 100, 2000}, fast arithmetic"""
    wd, sim = pair((201, 201), 1, bcT=bcT, strict=False, Rayleigh=Ra)
    done = 0
    for n in (1, 10, 100, 2000):
        wd.step(n - done); sim.step(n - done); done = n
        for k in FIELDS:
            assert close_enough(sim.gather(k), wd.gather(k), floor=velocity_floor(wd) if k in ("u", "v") else 0.0), (n, k)
    a, b = sim.check(), wd.check()
    assert np.isclose(a[0], b[0], rtol=1e-9) and np.isclose(a[1], b[1], rtol=1e-9)
    assert np.allclose(sim.calNuRe()[1:], nure_of(wd)[1:], rtol=1e-9)
    wd.close(); sim.close()


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("nprocs,dims", [(2, None), (4, None), (6, None), (3, (1, 3)), (4, (4, 1)), (9, None)])
def test_decomposed_equals_single_subdomain_bit_for_bit(strict, nprocs, dims):
    """the reference's seq == MPI contract on the device, in both arithmetic builds"""
    total = (67, 45)
    one = mg.BuoyancyDrivenCavity2D(total, strict=strict, Rayleigh=1e6)
    many = mg.BuoyancyDrivenCavity2D(total, nprocs=nprocs, dims=dims, strict=strict, Rayleigh=1e6)
    one.initial(); many.initial()
    one.step(40); many.step(40)
    for k in ("f", "g", "Fy") + FIELDS:
        assert np.array_equal(one.gather(k), many.gather(k)), k
    a, b = one.check(), many.check()
    assert np.isclose(a[0], b[0], rtol=1e-12) and np.isclose(a[1], b[1], rtol=1e-12)
    one.close(); many.close()


# ---------------- the OpenACC program (seq/bouyancy2d_acc.F90): periodic vertical walls, its own rounding of f_post(0) ----------------
def test_acc_shipped_constants_and_golden_cells():
    sim = mg.BuoyancyDrivenCavity2D(variant="acc", strict=True)
    assert sim.total == (513, 257) and sim.bcT == RB_PERIODIC and sim.params["lengthUnit"] == 513.0
    assert tuple(sim.params[k] for k in PARAM_NAMES[:9]) == tuple(GOLD["acc/params"])
    f, g, r = GOLD["cells/f"], GOLD["cells/g"], GOLD["cells/ruvT"]
    n, (nx, ny) = len(f), sim.total
    pad = np.arange(nx * ny) % n
    shp = lambda a: np.asfortranarray(a[pad].reshape(nx, ny, order="F"))
    sim.upload(0, f=np.asfortranarray(f[pad].T.reshape(9, nx, ny, order="F")), g=np.asfortranarray(g[pad].T.reshape(5, nx, ny, order="F")),
               rho=shp(r[:, 0]), u=shp(r[:, 1]), v=shp(r[:, 2]), T=shp(r[:, 3]))
    sim.collision(); sim.collisionT()
    fp = sim.download(0, "f_post")[:, 1:-1, 1:-1].reshape(9, -1, order="F").T
    gp = sim.download(0, "g_post")[:, 1:-1, 1:-1].reshape(5, -1, order="F").T
    assert np.array_equal(fp, GOLD["acc/collision_f_post"][pad]) and np.array_equal(gp, GOLD["acc/collisionT_g_post"][pad])
    assert not np.array_equal(fp, GOLD["collision/f_post"][pad])          # not the MPI program's bits
    sim.close()


def test_acc_periodic_walls_on_the_golden_block():
    nx, ny = 6, 5
    sim = mg.BuoyancyDrivenCavity2D((nx, ny), variant="acc", lengthUnit=513.0, strict=True, Rayleigh=1e5)
    assert sim.params["paraA"] == GOLD["acc/params"][3]
    sim.upload(0, f_post=GOLD["field/f_post"], g_post=GOLD["field/g_post"])
    sim.streaming(); sim.streamingT()
    assert np.array_equal(sim.download(0, "f"), GOLD["acc/streaming_f"]) and np.array_equal(sim.download(0, "g"), GOLD["acc/streamingT_g"])
    sim.upload(0, f=GOLD["field/f0"], g=GOLD["field/g0"])
    sim.bounceback(); sim.bouncebackT()
    assert np.array_equal(sim.download(0, "f"), GOLD["acc/bounceback_f"]) and np.array_equal(sim.download(0, "g"), GOLD["acc/bouncebackT_g"])
    sim.close()


@pytest.mark.parametrize("total,nprocs,dims", [((34, 33), 1, None), ((23, 19), 3, (1, 3)), ((130, 6), 2, (1, 2)), ((513, 257), 1, None)])
def test_acc_fused_step_strict_is_bit_exact(total, nprocs, dims):
    """the OpenACC program's loop (acc:165-187) through the fused kernel, split along y only (x must stay whole to wrap)"""
    wd, sim = pair(total, nprocs, dims, bcT=RB_PERIODIC, strict=True, variant="acc", Rayleigh=1e5)
    for n in (1, 2, 17):
        wd.step(n); sim.step(n)
        assert_rank_arrays_equal(wd, sim, ("f", "g", "Fx", "Fy") + FIELDS)
        assert_rank_arrays_equal(wd, sim, ("f_post", "g_post"), interior_only_post=True)
    wd.collision(); sim.collision(); wd.message_passing_f(); sim.message_passing_f()
    wd.streaming(); sim.streaming(); wd.bounceback(); sim.bounceback()
    wd.collisionT(); sim.collisionT(); wd.message_passing_g(); sim.message_passing_g()
    wd.streamingT(); sim.streamingT(); wd.bouncebackT(); sim.bouncebackT()
    wd.macro(); sim.macro(); wd.macroT(); sim.macroT()
    assert_rank_arrays_equal(wd, sim, ("f", "g", "Fx", "Fy") + FIELDS)
    wd.step(3); sim.step(3)
    assert_rank_arrays_equal(wd, sim, ("f", "g", "Fx", "Fy") + FIELDS)
    wd.close(); sim.close()


def test_acc_shipped_case_fast_within_tolerance():
    """513 x 257, Ra = 1e5, N in {1, 10, 100, 1000}: the OpenACC program's own configuration in the throughput arithmetic"""
    wd, sim = pair((513, 257), 1, bcT=RB_PERIODIC, strict=False, variant="acc", Rayleigh=1e5)
    done = 0
    for n in (1, 10, 100, 1000):
        wd.step(n - done); sim.step(n - done); done = n
        for k in FIELDS:
            assert close_enough(sim.gather(k), wd.gather(k), floor=velocity_floor(wd) if k in ("u", "v") else 0.0), (n, k)
    a, b = sim.check(), wd.check()
    assert np.isclose(a[0], b[0], rtol=1e-9) and np.isclose(a[1], b[1], rtol=1e-9)
    wd.close(); sim.close()


def test_large_lattice_properties():
    """4096 x 4096 (no oracle run).  Side-heated: total mass constant to rounding, T bounded, four subdomains == one bit for bit.
    Rayleigh-Benard set: the initial state does not depend on x, so for N steps every column farther than N cells from the side
    walls runs the same arithmetic on the same inputs (causality): those columns are bit-identical."""
    total, n = (4096, 4096), 12
    one = mg.BuoyancyDrivenCavity2D(total, strict=False)
    many = mg.BuoyancyDrivenCavity2D(total, nprocs=4, strict=False)
    one.initial(); many.initial()
    m0 = one.gather("rho").sum()
    one.step(n); many.step(n)
    rho, u, v, T = (one.gather(k) for k in FIELDS)
    assert abs(rho.sum() - m0) / m0 < 1e-13
    assert np.abs(v).max() > 0.0 and T.min() > -0.01 and T.max() < 1.01
    for k, a in zip(FIELDS, (rho, u, v, T)):
        assert np.array_equal(many.gather(k), a), k
    one.close(); many.close()
    rb = mg.BuoyancyDrivenCavity2D(total, bcT=RAYLEIGH_BENARD, strict=False)
    rb.initial()
    rb.step(n)
    for k in FIELDS:
        a = rb.gather(k)
        mid = a[n + 1:-n - 1, :]
        assert np.all(mid == mid[0:1, :]), k
    assert np.abs(rb.gather("v")).max() > 0.0
    rb.close()


def test_error_behaviour():
    with pytest.raises(mg.MglcError):
        mg.BuoyancyDrivenCavity2D((8, 8), nprocs=3, dims=(2, 2))
    with pytest.raises(mg.MglcError):
        mg.BuoyancyDrivenCavity2D((2, 8), nprocs=4, dims=(4, 1))
    with pytest.raises(mg.MglcError):
        mg.BuoyancyDrivenCavity2D((8, 8), bcT=(0, 0, 0, 7))
    with pytest.raises(mg.MglcError):                       # periodic vertical walls cannot be split along x
        mg.BuoyancyDrivenCavity2D((16, 16), nprocs=2, dims=(2, 1), bcT=RB_PERIODIC)
    with pytest.raises(mg.MglcError):                       # ... and come in pairs
        mg.BuoyancyDrivenCavity2D((16, 16), bcT=(3, 0, 2, 1))
    with pytest.raises(mg.MglcError):                       # the reference stops when paraA leaves (-4, 1): initial.F90:30
        mg.BuoyancyDrivenCavity2D((4001, 4001), Rayleigh=1e3, Mach=0.3)
    sim = mg.BuoyancyDrivenCavity2D((8, 8))
    with pytest.raises(mg.MglcError):
        sim.step(-1)
    with pytest.raises(ValueError):
        sim.upload(0, rho=np.zeros((3, 3)))
    sim.close()


# ---- the sheared Rayleigh-Benard programs: moving walls, corner cells of bouncebackT() (seq/R_B_2d.F90) ------------------------
SRUN_SHEAR = np.load(os.path.join(HERE, "golden", "ref_fortran_thermal2d_seq_run_sheared_rb.npz"))
SHEAR_WALLS = [4e-3, -4e-3, -4e-3, 4e-3, 3e-3, 3.5e-3, 2.5e-3, 2e-3]


def sheared_pair(total, nprocs, dims, strict, bcT=RAYLEIGH_BENARD, Uwall=SHEAR_WALLS, **params):
    wd = orc.Thermal2DWorld(total, nprocs, dims, bcT=bcT, Uwall=Uwall, cornersT=True, **params)
    sim = mg.BuoyancyDrivenCavity2D(total, nprocs=nprocs, dims=dims, bcT=bcT, strict=strict, Uwall=Uwall, cornersT=True, **params)
    wd.initial(); sim.initial()
    return wd, sim


@pytest.mark.parametrize("nprocs,dims", [(1, None), (2, None), (4, (2, 2)), (6, (3, 2)), (12, (4, 3))])
def test_sheared_rb_program_as_shipped_reproduces_its_fortran_run(nprocs, dims):
    """seq/R_B_2d.F90 AS SHIPPED, evaluated from its text on 9 x 7 (make_golden_thermal2d_seq_run.py): variant "sheared_rb" of the
    driver (its desc_init: Pr 5.3, plates, shearReynolds = 100, corner cells) through step(): f, g, fields bit for bit"""
    S = SRUN_SHEAR
    total = tuple(int(x) for x in S["shape"])
    sim = mg.BuoyancyDrivenCavity2D(total, nprocs=nprocs, dims=dims, strict=True, variant="sheared_rb")
    assert sim.Uwall == tuple(float(x) for x in S["uwall"]) and sim.cornersT
    assert tuple(sim.params[k] for k in ("tauf", "viscosity", "diffusivity", "paraA", "gBeta", "Snu", "Sq", "Qd", "Qnu")) == tuple(S["params"])
    sim.initial()

    def same(tag):
        assert np.array_equal(sim.gather("f"), S[tag + "/f"]), tag
        assert np.array_equal(sim.gather("g"), S[tag + "/g"]), tag
        assert np.array_equal(np.stack([sim.gather(k) for k in FIELDS]), S[tag + "/ruvT"]), tag
        assert np.array_equal(np.stack([sim.gather(k) for k in ("Fx", "Fy")]), S[tag + "/F"]), tag

    same("run0")
    done = 0
    for n in (1, 2, 20, 25):
        sim.step(n - done); done = n
        same(f"run{n}")
    sim.close()


@pytest.mark.parametrize("total,nprocs,dims,bcT", [((34, 33), 1, None, RAYLEIGH_BENARD), ((23, 19), 4, None, RAYLEIGH_BENARD),
                                                   ((23, 19), 6, None, SIDE_HEATED), ((130, 6), 2, None, RAYLEIGH_BENARD),
                                                   ((9, 31), 3, (1, 3), RAYLEIGH_BENARD)])
def test_sheared_walls_fused_step_and_subroutines_strict_bit_exact(total, nprocs, dims, bcT):
    """eight different wall velocities: the fused step (rho of the previous macro() carried at the wall cells), then the
    per-subroutine entry points (bounceback() with the rho field, bouncebackT() with the corner rule), then step() again"""
    wd, sim = sheared_pair(total, nprocs, dims, True, bcT=bcT, Rayleigh=1e6)
    assert_rank_arrays_equal(wd, sim, ("f", "g") + FIELDS)
    for n in (1, 2, 17):
        wd.step(n); sim.step(n)
        assert_rank_arrays_equal(wd, sim, ("f", "g", "Fx", "Fy") + FIELDS)
    wd.collision(); sim.collision(); wd.message_passing_f(); sim.message_passing_f()
    wd.streaming(); sim.streaming(); wd.bounceback(); sim.bounceback()
    assert_rank_arrays_equal(wd, sim, ("f",))
    wd.collisionT(); sim.collisionT(); wd.message_passing_g(); sim.message_passing_g()
    wd.streamingT(); sim.streamingT(); wd.bouncebackT(); sim.bouncebackT()
    assert_rank_arrays_equal(wd, sim, ("g",))
    wd.macro(); sim.macro(); wd.macroT(); sim.macroT()
    wd.step(3); sim.step(3)
    assert_rank_arrays_equal(wd, sim, ("f", "g", "f_post", "g_post", "Fx", "Fy") + FIELDS)
    wd.close(); sim.close()


def test_sheared_rb_graph_replay_and_fast_arithmetic():
    """a single small subdomain replays its fused launches from CUDA graphs (the rho field is an argument of the captured
    launches); fast arithmetic at 201 x 201 as shipped stays within the north-star tolerance over 1000 steps"""
    wd, sim = sheared_pair((45, 38), 1, None, True, Rayleigh=1e5)
    wd.step(136); sim.step(136)
    assert_rank_arrays_equal(wd, sim, ("f", "g", "Fx", "Fy") + FIELDS)
    wd.close(); sim.close()
    sim = mg.BuoyancyDrivenCavity2D(None, strict=False, variant="sheared_rb")
    wd = orc.Thermal2DWorld((201, 201), 1, None, bcT=RAYLEIGH_BENARD, Prandtl=5.3, Uwall=list(sim.Uwall), cornersT=True)
    wd.initial(); sim.initial()
    done = 0
    for n in (1, 10, 100, 1000):
        wd.step(n - done); sim.step(n - done); done = n
        for k in FIELDS:
            assert close_enough(sim.gather(k), wd.gather(k), floor=velocity_floor(wd) if k in ("u", "v") else 0.0), (n, k)
    # the sheared top wall drags the fluid: u just below the top-left half of the lid follows UwallTopLeft's sign
    assert np.sign(wd.gather("u")[20:80, -2].mean()) == np.sign(sim.Uwall[0])
    wd.close(); sim.close()


def test_moving_walls_with_periodic_vertical_walls_are_refused():
    from mglc_b200 import _lib as L
    with pytest.raises(L.MglcError):
        mg.BuoyancyDrivenCavity2D((33, 17), variant="acc", Uwall=[1e-3] * 8)
