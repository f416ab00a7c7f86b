"""GPU parity tests of the 2-D lid-driven cavity with the single-relaxation-time (BGK) collision: variant "s" of mglc_l2d_* =
the reference's C program MPI/Lid_driven_cavity/c/lid_driven_cavity.c with its own model switch set to SRT (c:13-14, collision
c:160-176), against the CPU oracle (oracle/lid2d.c, variant L2_S, pinned live to that program compiled with the switch) and
against the committed outputs of that compiled program (tests/golden/ref_lid2d_srt.npz).  Strict arithmetic: bit-exact.  Fast
arithmetic: <= 1e-12 relative L2, <= 1e-10 max pointwise.  The same kernel source is checked on the CPU by
test_lid2d_kernels_host.py (host shim).  BGK at the shipped Re = 1000 needs the shipped 200-cell cavity (tau = 0.56); the small
lattices below run at Re = 100.
(The file sorts last on purpose: it was added after the other GPU files and a failure here must not hide their results under -x.)"""
import os

import numpy as np
import pytest

import mglc_b200 as mg
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "ref_lid2d.npz"))
SGOLD = np.load(os.path.join(HERE, "golden", "ref_lid2d_srt.npz"))
REL_L2, MAX_ABS = 1e-12, 1e-10


def close_enough(got, want):
    d = np.linalg.norm((got - want).ravel()) / max(np.linalg.norm(want.ravel()), 1e-300)
    return d <= REL_L2 and np.abs(got - want).max() <= MAX_ABS


def velocity_close_enough(sim, wd):
    """relative L2 of the velocity VECTOR (u, v) jointly (after a few steps v alone is non-zero only in the two lid corners),
    max pointwise on each component"""
    du, dv = sim.gather("u") - wd.gather("u"), sim.gather("v") - wd.gather("v")
    scale = np.sqrt((wd.gather("u") ** 2).sum() + (wd.gather("v") ** 2).sum())
    return np.sqrt((du ** 2).sum() + (dv ** 2).sum()) <= REL_L2 * scale and max(np.abs(du).max(), np.abs(dv).max()) <= MAX_ABS


def pair(total, nprocs=1, dims=None, strict=True, seed=None, Re=100.0):
    wd = orc.Lid2DWorld(total, nprocs, dims, variant="s", Re=Re)
    sim = mg.LidDrivenCavity2D(total, nprocs=nprocs, dims=dims, variant="s", strict=strict, Re=Re)
    assert sim.dims == wd.dims and (sim.tauf, sim.Snu, sim.Sq) == (wd.tauf, wd.Snu, wd.Sq)
    wd.initial(); sim.initial()
    if seed is not None:
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


def test_collision_golden_cells_of_the_compiled_reference():
    """the seeded cells whose f_post the reference program (model = SRT) itself produced"""
    f, ruv = GOLD["cells/f"], GOLD["cells/ruv"]
    n, nx = len(f), 200                                        # the shipped width: tau depends on total_nx
    pad = np.arange(nx) % n
    for strict in (True, False):
        sim = mg.LidDrivenCavity2D((nx, 1), variant="s", strict=strict)
        assert (sim.tauf, sim.Snu, sim.Sq) == tuple(SGOLD["params"])
        sim.upload(0, f=f[pad].T.reshape(9, nx, 1), rho=ruv[pad, 0].reshape(nx, 1), u=ruv[pad, 1].reshape(nx, 1), v=ruv[pad, 2].reshape(nx, 1))
        sim.collision()
        got = sim.download(0, "f_post")[:, 1:-1, 1].T
        want = SGOLD["collision_f_post"][pad]
        assert np.array_equal(got, want) if strict else close_enough(got, want)
        sim.close()


def test_run_matches_the_compiled_reference_programs_committed_outputs():
    """strict GPU run vs what the reference's own C program, compiled with model = SRT, computed (ref_lid2d_srt.npz)"""
    sim = mg.LidDrivenCavity2D(variant="s", strict=True)
    assert sim.total == (200, 200)
    sim.initial()
    done = 0
    for n in (1, 10, 100, 1000):
        sim.step(n - done); done = n
        for k in ("rho", "u", "v"):
            a = np.ascontiguousarray(sim.gather(k))
            assert np.array_equal(a[100, :], SGOLD[f"run{n}/{k}_col100"]), (n, k)
            assert np.array_equal(a[:, 199], SGOLD[f"run{n}/{k}_row199"]), (n, k)
            assert np.array_equal(a[:, 0], SGOLD[f"run{n}/{k}_row0"]), (n, k)
            assert np.array_equal(a[::4, ::4], SGOLD[f"run{n}/{k}_stride4"]), (n, k)
        f = np.ascontiguousarray(np.transpose(sim.gather("f"), (1, 2, 0)))
        assert np.array_equal(f[:3, :3, :], SGOLD[f"run{n}/f_corner"]) and np.array_equal(f[-3:, -3:, :], SGOLD[f"run{n}/f_topright"])
    assert np.isclose(sim.check(), SGOLD["check_1000"][0], rtol=1e-12)
    sim.step(100)
    assert np.isclose(sim.check(), SGOLD["check_1100"][0], rtol=1e-12)
    sim.close()


@pytest.mark.parametrize("total,nprocs,dims", [((34, 33), 1, None), ((34, 33), 4, None), ((23, 19), 6, None), ((9, 31), 3, (1, 3))])
def test_each_subroutine_bit_exact_strict(total, nprocs, dims):
    wd, sim = pair(total, nprocs, dims, strict=True, seed=3)
    for R in wd.ranks:
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
    wd.close(); sim.close()


@pytest.mark.parametrize("total,nprocs,dims", [((34, 33), 1, None), ((23, 19), 4, None), ((23, 19), 6, None), ((130, 6), 2, None)])
def test_fused_step_strict_is_bit_exact(total, nprocs, dims):
    wd, sim = pair(total, nprocs, dims, strict=True)
    for n in (1, 2, 17):
        wd.step(n); sim.step(n)
        assert_rank_arrays_equal(wd, sim, ("f", "rho", "u", "v"))
        assert_rank_arrays_equal(wd, sim, ("f_post",), interior_only_fpost=True)
    assert_rank_arrays_equal(wd, sim, ("f_post",))
    wd.close(); sim.close()


def test_graph_replayed_steps_strict_are_bit_exact():
    wd, sim = pair((45, 38), 1, strict=True)
    for n in (136, 65):
        wd.step(n); sim.step(n)
        assert_rank_arrays_equal(wd, sim, ("f", "rho", "u", "v"))
    wd.close(); sim.close()


def test_shipped_case_fast_within_tolerance():
    """the shipped 200 x 200 grid at Re = 1000 (tau = 0.56), N in {1, 10, 100, 1000}, fast arithmetic"""
    wd, sim = pair((200, 200), 1, strict=False, Re=1000.0)
    done = 0
    for n in (1, 10, 100, 1000):
        wd.step(n - done); sim.step(n - done); done = n
        assert close_enough(sim.gather("rho"), wd.gather("rho")), n
        assert velocity_close_enough(sim, wd), n
    assert np.isclose(sim.check(), wd.check(), rtol=1e-9)
    wd.close(); sim.close()


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("nprocs,dims", [(4, None), (6, None), (3, (1, 3))])
def test_decomposed_equals_single_subdomain_bit_for_bit(strict, nprocs, dims):
    total = (67, 45)
    one = mg.LidDrivenCavity2D(total, variant="s", strict=strict, Re=100.0)
    many = mg.LidDrivenCavity2D(total, nprocs=nprocs, dims=dims, variant="s", strict=strict, Re=100.0)
    one.initial(); many.initial()
    one.step(40); many.step(40)
    for k in ("f", "rho", "u", "v"):
        assert np.array_equal(one.gather(k), many.gather(k)), k
    one.close(); many.close()


def test_c_driver_with_the_srt_switch_writes_the_reference_programs_file(tmp_path):
    """examples/lid2d_driver.c with model = 1 (the program's own SRT switch, c:13-14): after 2000 iterations in strict
    arithmetic its flow_binary has the SHA-256 of the file the reference program compiled with that switch wrote"""
    import hashlib
    import subprocess
    root = os.path.dirname(HERE)
    exe, lib = str(tmp_path / "lid2d_driver"), os.path.join(root, "mglc_b200")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                           os.path.join(root, "examples", "lid2d_driver.c"), "-L", lib, "-lmglc", f"-Wl,-rpath,{lib}", "-o", exe])
    r = subprocess.run([exe, "2000", "1", "flow_binary", "1"], capture_output=True, text=True, cwd=tmp_path, timeout=300)
    assert r.returncode == 0, r.stderr
    raw = (tmp_path / "flow_binary").read_bytes()
    assert len(raw) == int(SGOLD["output_binary_len"][0])
    assert hashlib.sha256(raw).digest() == SGOLD["output_binary_sha256"].tobytes()
