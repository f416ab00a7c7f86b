"""GPU parity tests of the 2-D D2Q9 lid-driven cavity path (libmglc.so through the C ABI) against the CPU oracle
(oracle/lid2d.c, pinned bit for bit to the reference's compiled C program and to its Fortran text by test_oracle_lid2d.py) and
against the committed outputs of the reference's own program (tests/golden/ref_lid2d.npz).
Strict arithmetic: everything bit-exact.  Fast arithmetic: copy-type subroutines and macro() bit-exact, collision()/step() to
the north-star tolerance (<= 1e-12 relative L2, <= 1e-10 max pointwise)."""
import os

import numpy as np
import pytest

import mglc_b200 as mg
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "ref_lid2d.npz"))
REL_L2, MAX_ABS = 1e-12, 1e-10


def close_enough(got, want):
    d = np.linalg.norm((got - want).ravel()) / max(np.linalg.norm(want.ravel()), 1e-300)
    return d <= REL_L2 and np.abs(got - want).max() <= MAX_ABS


def seeded_state(total, seed):
    """f = a perturbed equilibrium, rho/u/v independent of it: every moment non-trivial"""
    rng = np.random.default_rng(seed)
    wd = orc.Lid2DWorld(total, 1)
    wd.initial()
    f = np.asfortranarray(wd.gather("f") * (1.0 + 0.05 * rng.uniform(-1, 1, (9,) + tuple(total))))
    wd.close()
    rho = np.asfortranarray(1.0 + 0.02 * rng.uniform(-1, 1, total))
    u, v = (np.asfortranarray(0.05 * rng.uniform(-1, 1, total)) for _ in range(2))
    return f, rho, u, v


def pair(total, nprocs=1, dims=None, variant="f", strict=True, seed=None):
    wd = orc.Lid2DWorld(total, nprocs, dims, variant=variant)
    sim = mg.LidDrivenCavity2D(total, nprocs=nprocs, dims=dims, variant=variant, strict=strict)
    assert sim.dims == wd.dims and (sim.tauf, sim.Snu, sim.Sq) == (wd.tauf, wd.Snu, wd.Sq)
    for r, R in enumerate(wd.ranks):
        inf = sim.info[r]
        assert inf["n"] == R.n and inf["start"] == R.start and inf["coords"] == R.coords and inf["nbr"] == R.nbr + R.cnr
    if seed is None:
        wd.initial(); sim.initial()
    else:
        f, rho, u, v = seeded_state(total, seed)
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


@pytest.mark.parametrize("variant,total", [("c", (200, 200)), ("f", (201, 201)), ("f", (37, 5)), ("c", (1, 9)), ("f", (130, 1))])
def test_initial_bit_exact(variant, total):
    for nprocs in (1, 2, 4):
        if any(n < 2 for n in total) and nprocs > 1:
            continue
        wd, sim = pair(total, nprocs, variant=variant)
        assert_rank_arrays_equal(wd, sim, ("f", "rho", "u", "v"))
        wd.close(); sim.close()


@pytest.mark.parametrize("variant", ["c", "f"])
def test_collision_golden_cells_of_the_reference(variant):
    """the 64 seeded cells whose f_post the reference itself produced (compiled C program / Fortran text)"""
    f, ruv = GOLD["cells/f"], GOLD["cells/ruv"]
    n, nx = len(f), 200 if variant == "c" else 201            # the shipped width: tau depends on total_nx
    pad = np.arange(nx) % n
    for strict in (True, False):
        sim = mg.LidDrivenCavity2D((nx, 1), variant=variant, strict=strict)
        assert (sim.tauf, sim.Snu, sim.Sq) == tuple(GOLD[variant + "/params"])
        sim.upload(0, f=f[pad].T.reshape(9, nx, 1), rho=ruv[pad, 0].reshape(nx, 1), u=ruv[pad, 1].reshape(nx, 1), v=ruv[pad, 2].reshape(nx, 1))
        sim.collision()
        got = sim.download(0, "f_post")[:, 1:-1, 1].T
        want = GOLD[variant + "/collision_f_post"][pad]
        assert np.array_equal(got, want) if strict else close_enough(got, want)
        sim.close()


@pytest.mark.parametrize("variant", ["c", "f"])
@pytest.mark.parametrize("total,nprocs,dims", [((34, 33), 1, None), ((34, 33), 4, None), ((23, 19), 6, None), ((40, 7), 3, (3, 1)), ((9, 31), 3, (1, 3))])
def test_each_subroutine_bit_exact_strict(variant, total, nprocs, dims):
    wd, sim = pair(total, nprocs, dims, variant=variant, strict=True, seed=3)
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


@pytest.mark.parametrize("total,nprocs", [((34, 33), 1), ((34, 33), 4)])
def test_copy_type_subroutines_bit_exact_on_random_f_post(total, nprocs):
    """exchange / streaming / bounceback move doubles: `==` on every value, random f_post including the halos"""
    wd, sim = pair(total, nprocs, strict=False, seed=4)
    rng = np.random.default_rng(7)
    for r, R in enumerate(wd.ranks):
        R.f_post[...] = rng.random(R.f_post.shape)
        sim.upload(r, f_post=R.f_post)
    wd.message_passing_sendrecv(); sim.message_passing_sendrecv()
    assert_rank_arrays_equal(wd, sim, ("f_post",))
    wd.streaming(); sim.streaming()
    assert_rank_arrays_equal(wd, sim, ("f",))
    wd.bounceback(); sim.bounceback()
    assert_rank_arrays_equal(wd, sim, ("f",))
    wd.macro(); sim.macro()
    assert_rank_arrays_equal(wd, sim, ("rho", "u", "v"))
    wd.close(); sim.close()


@pytest.mark.parametrize("variant", ["c", "f"])
@pytest.mark.parametrize("total,nprocs,dims", [((34, 33), 1, None), ((23, 19), 4, None), ((23, 19), 6, None), ((130, 6), 2, None), ((9, 31), 3, (1, 3))])
def test_fused_step_strict_is_bit_exact(variant, total, nprocs, dims):
    """step(N) = the rotated loop (collision, N-1 fused launches, stream+macro): f, f_post incl. exchanged halos, rho, u, v"""
    wd, sim = pair(total, nprocs, dims, variant=variant, strict=True)
    for n in (1, 2, 17):
        wd.step(n); sim.step(n)
        assert_rank_arrays_equal(wd, sim, ("f", "rho", "u", "v"))
        assert_rank_arrays_equal(wd, sim, ("f_post",), interior_only_fpost=True)
    # the exchanged halo entries too (wall halos are never written by either side: zero in both)
    assert_rank_arrays_equal(wd, sim, ("f_post",))
    # calls compose with the per-subroutine entry points
    wd.collision(); sim.collision()
    wd.message_passing_sendrecv(); sim.message_passing_sendrecv()
    wd.streaming(); sim.streaming(); wd.bounceback(); sim.bounceback(); wd.macro(); sim.macro()
    wd.step(3); sim.step(3)
    assert_rank_arrays_equal(wd, sim, ("f", "f_post", "rho", "u", "v"))
    wd.close(); sim.close()


@pytest.mark.parametrize("variant", ["c", "f"])
def test_graph_replayed_steps_strict_are_bit_exact(variant):
    """a single small subdomain replays its fused launches as CUDA graphs of 64 kernels (the lid-row side buffers alternate
    inside the graph): 1 + 2 x 64 + 7 steps, then 64 + 1 more from the other ping-pong index, bit for bit"""
    wd, sim = pair((45, 38), 1, variant=variant, strict=True)
    for n in (136, 65):
        wd.step(n); sim.step(n)
        assert_rank_arrays_equal(wd, sim, ("f", "rho", "u", "v"))
    wd.close(); sim.close()


def test_wall_halos_are_never_read_by_the_fused_step():
    wd, sim = pair((31, 17), 4, strict=True)
    for r in range(4):
        shape = sim._shape(r, "f_post")
        sim.upload(r, f_post=np.full(shape, np.nan))
    wd.step(20); sim.step(20)
    assert_rank_arrays_equal(wd, sim, ("f", "rho", "u", "v"))
    wd.close(); sim.close()


@pytest.mark.parametrize("variant,total", [("c", (200, 200)), ("f", (201, 201))])
def test_shipped_case_fast_within_tolerance(variant, total):
    """config-1 analogue in 2-D: the shipped grid and Reynolds number, N in {1, 10, 100, 2000}, fast arithmetic"""
    wd, sim = pair(total, 1, variant=variant, strict=False)
    done = 0
    for n in (1, 10, 100, 2000):
        wd.step(n - done); sim.step(n - done); done = n
        for k in ("rho", "u", "v"):
            assert close_enough(sim.gather(k), wd.gather(k)), (n, k)
    assert np.isclose(sim.check(), wd.check(), rtol=1e-9)
    wd.close(); sim.close()


def test_variant_c_run_matches_the_reference_programs_committed_outputs():
    """strict GPU run vs what the reference's own compiled C program wrote (ref_lid2d.npz), bit for bit"""
    sim = mg.LidDrivenCavity2D(variant="c", strict=True)
    assert sim.total == (200, 200)
    sim.initial()
    done = 0
    for n in (1, 10, 100, 1000):
        sim.step(n - done); done = n
        for k in ("rho", "u", "v"):
            a = np.ascontiguousarray(sim.gather(k))
            assert np.array_equal(a[100, :], GOLD[f"c/run{n}/{k}_col100"]), (n, k)
            assert np.array_equal(a[:, 199], GOLD[f"c/run{n}/{k}_row199"]), (n, k)
            assert np.array_equal(a[:, 0], GOLD[f"c/run{n}/{k}_row0"]), (n, k)
        f = np.ascontiguousarray(np.transpose(sim.gather("f"), (1, 2, 0)))
        assert np.array_equal(f[:3, :3, :], GOLD[f"c/run{n}/f_corner"]) and np.array_equal(f[-3:, -3:, :], GOLD[f"c/run{n}/f_topright"])
    assert np.isclose(sim.check(), GOLD["c/check_1000"][0], rtol=1e-12)
    sim.step(1000)
    assert np.isclose(sim.check(), GOLD["c/check_2000"][0], rtol=1e-12)
    for k in ("rho", "u", "v"):
        assert np.array_equal(np.ascontiguousarray(sim.gather(k)), GOLD[f"c/run2000/{k}_full"]), k
    sim.close()


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("nprocs,dims", [(2, None), (4, None), (6, None), (3, (1, 3)), (4, (4, 1)), (9, None)])
def test_decomposed_equals_single_subdomain_bit_for_bit(strict, nprocs, dims):
    """the reference's seq == MPI contract on the device, in both arithmetic builds"""
    total = (67, 45)
    one = mg.LidDrivenCavity2D(total, strict=strict)
    many = mg.LidDrivenCavity2D(total, nprocs=nprocs, dims=dims, strict=strict)
    one.initial(); many.initial()
    one.step(40); many.step(40)
    for k in ("f", "rho", "u", "v"):
        assert np.array_equal(one.gather(k), many.gather(k)), k
    assert np.isclose(one.check(), many.check(), rtol=1e-12)
    one.close(); many.close()


def test_large_lattice_properties():
    """4096 x 4096 (no oracle run): mass conservation, mirror symmetry of the lid-driven flow about x -> -x is broken by the lid,
    so use the exact invariants instead: total mass constant to rounding, causality (cells farther than N from the lid stay at
    rest exactly), and decomposition invariance on 4 subdomains"""
    total, n = (4096, 4096), 12
    one = mg.LidDrivenCavity2D(total, strict=False)
    many = mg.LidDrivenCavity2D(total, nprocs=4, strict=False)
    one.initial(); many.initial()
    m0 = one.gather("rho").sum()
    one.step(n); many.step(n)
    rho, u, v = (one.gather(k) for k in ("rho", "u", "v"))
    assert abs(rho.sum() - m0) / m0 < 1e-13
    assert np.all(u[:, : total[1] - n - 1] == 0.0) and np.all(v[:, : total[1] - n - 1] == 0.0) and np.abs(rho[:, : total[1] - n - 1] - 1.0).max() < 1e-14
    assert np.abs(u[:, -1]).max() > 0.0
    for k, a in (("rho", rho), ("u", u), ("v", v)):
        assert np.array_equal(many.gather(k), a), k
    one.close(); many.close()


def test_error_behaviour():
    with pytest.raises(mg.MglcError):
        mg.LidDrivenCavity2D((8, 8), nprocs=3, dims=(2, 2))
    with pytest.raises(mg.MglcError):
        mg.LidDrivenCavity2D((2, 8), nprocs=4, dims=(4, 1))
    sim = mg.LidDrivenCavity2D((8, 8))
    with pytest.raises(mg.MglcError):
        sim.step(-1)
    with pytest.raises(ValueError):
        sim.upload(0, rho=np.zeros((3, 3)))
    sim.close()
