"""GPU parity tests of the Jacobi halo-exchange path (libmglc.so through the C ABI) against the CPU
oracle and the reference-generated golden vectors.  Everything here is bit-exact: the update is adds in
the reference's order and one multiply by a constant."""
import os

import numpy as np
import pytest

import mglc_b200 as mg
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "laplace2d_ref.npz"))


def load_random(wd, sim, total, seed, with_f=True):
    rng = np.random.default_rng(seed)
    glob = rng.random(tuple(n + 2 for n in total))
    src = rng.random(tuple(n + 2 for n in total)) if with_f else None
    for r, inf in enumerate(wd.info):
        sl = tuple(slice(s, s + n + 2) for s, n in zip(inf["start"], inf["n"]))
        wd.array(r, "A")[...] = glob[sl]
        wd.array(r, "A_new")[...] = glob[sl]
        if with_f:
            wd.array(r, "f")[...] = src[sl]
        assert sim.info[r]["n"] == inf["n"] and sim.info[r]["start"] == inf["start"]
        sim.upload(r, A=glob[sl], A_new=glob[sl], f=src[sl] if with_f else None)


@pytest.mark.parametrize("name,its", [("shipped_bc_19x14", 25), ("random_23x37", 7), ("random_130x9", 3)])
def test_2d_matches_reference_golden_vectors(name, its):
    """GPU vs the reference's own compiled jacobi()+swap() (MPI/Laplace/c/laplace2d.c)."""
    A0, want, errs = GOLD[name + "/A0"], GOLD[name + "/A"], GOLD[name + "/err"]
    sim = mg.Jacobi((A0.shape[0] - 2, A0.shape[1] - 2))
    sim.upload(0, A=A0, A_new=A0)
    assert sim.check_diff() >= 0.0                # A_p <- A
    for it in range(its):
        sim.step(1)
        assert sim.check_diff() == errs[it]
    assert np.array_equal(sim.download(0)[1:-1, 1:-1], want[1:-1, 1:-1])
    sim.close()


@pytest.mark.parametrize("total", [(10, 8), (6, 5, 4), (130, 3, 2)])
def test_init_bit_exact(total):
    for nprocs in (1, 2, 4):
        wd, sim = orc.JacobiWorld(total, nprocs), mg.Jacobi(total, nprocs=nprocs)
        wd.init(); sim.init()
        for r in range(nprocs):
            assert np.array_equal(sim.download(r, "A"), wd.array(r, "A"))
            assert np.array_equal(sim.download(r, "A_new"), wd.array(r, "A_new"))
        wd.close(); sim.close()


@pytest.mark.parametrize("total,with_f", [((37, 23), True), ((300, 5), False), ((33, 9, 70), True), ((200, 7, 3), False)])
def test_sweep_bit_exact(total, with_f):
    wd, sim = orc.JacobiWorld(total, 1), mg.Jacobi(total)
    load_random(wd, sim, total, 11, with_f)
    wd.jacobi(); sim.jacobi()
    assert np.array_equal(sim.download(0), wd.array(0))
    wd.close(); sim.close()


@pytest.mark.parametrize("kernel,shape,th,chunks", [("tma", 0, 0, 0), ("tma", 1, 0, 0), ("tma", 2, 0, 0), ("tma", 0, 14, 2), ("tma", 1, 13, 2), ("tma", 2, 5, 3),
                                                    ("tma", 1, 1, 9), ("tma", 0, 6, 1), ("reg", 0, 0, 0), ("reg", 1, 0, 0)])
@pytest.mark.parametrize("total,with_f", [((131, 17, 9), True), ((260, 35, 70), False), ((128, 16, 3), True), ((5, 4, 3), False)])
def test_3d_sweep_kernels_bit_exact(total, with_f, kernel, shape, th, chunks, monkeypatch):
    """Both 3-D sweep kernels -- the TMA pipeline (persistent CTAs, 128 x th tiles with th chosen per block or forced, any number
    of z chunks, each of its three CTA shapes) and the register-blocked LDG kernel -- on blocks that do not divide into tiles:
    every cell, several sweeps, bit for bit."""
    monkeypatch.setenv("MGLC_JACOBI_KERNEL", kernel)
    monkeypatch.setenv("MGLC_JACOBI_TMA_SHAPE", str(shape))
    monkeypatch.setenv("MGLC_JACOBI_PF", str(shape))              # register-blocked kernel: with / without the one-plane prefetch
    if th:
        monkeypatch.setenv("MGLC_JACOBI_TMA_TH", str(th))
        monkeypatch.setenv("MGLC_JACOBI_TMA_CHUNKS", str(chunks))
    wd, sim = orc.JacobiWorld(total, 1), mg.Jacobi(total)
    load_random(wd, sim, total, 5, with_f)
    for _ in range(3):
        wd.jacobi(); sim.jacobi()
        assert np.array_equal(sim.download(0), wd.array(0))
    wd.close(); sim.close()


@pytest.mark.parametrize("total,nprocs,dims", [((37, 23), 2, None), ((37, 23), 6, None), ((37, 23), 3, (1, 3)),
                                               ((17, 13, 11), 2, None), ((17, 13, 11), 8, None),
                                               ((17, 13, 11), 12, None), ((17, 13, 11), 3, (1, 1, 3))])
def test_exchange_and_steps_bit_exact(total, nprocs, dims):
    wd, sim = orc.JacobiWorld(total, nprocs, dims), mg.Jacobi(total, nprocs=nprocs, dims=dims)
    assert wd.dims == sim.dims
    load_random(wd, sim, total, 3)
    wd.exchange_message(); sim.exchange_message()
    for r in range(nprocs):
        assert np.array_equal(sim.download(r), wd.array(r)), r      # ghost layers included
    wd.step(7); sim.step(7)
    for r in range(nprocs):
        a, b = sim.download(r), wd.array(r)
        inner = tuple(slice(1, -1) for _ in total)
        assert np.array_equal(a[inner], b[inner]), r
    assert sim.check_diff() == wd.check_diff()
    wd.close(); sim.close()


def test_reference_problem_2d_1000_iterations():
    """The shipped 2-D problem (top boundary 1) at a size the oracle finishes in seconds."""
    total = (400, 300)
    wd, sim = orc.JacobiWorld(total, 1), mg.Jacobi(total, nprocs=4)
    wd.init(); sim.init()
    for _ in range(10):
        wd.step(100); sim.step(100)
        assert sim.check_diff() == wd.check_diff()
    assert np.array_equal(sim.gather(), wd.gather())
    wd.close(); sim.close()


def test_3d_256_matches_oracle_and_512_properties():
    total = (256, 256, 256)
    wd, sim = orc.JacobiWorld(total, 1), mg.Jacobi(total)
    wd.init(); sim.init()
    wd.step(12); sim.step(12)
    assert np.array_equal(sim.gather(), wd.gather())
    wd.close(); sim.close()
    # config 2 size: maximum principle, exact x-mirror symmetry (a+b == b+a), plane-wise monotone in z
    sim = mg.Jacobi((512, 512, 512))
    sim.init()
    sim.step(40)
    a = sim.gather()
    assert a.min() >= 0.0 and a.max() <= 1.0
    assert np.array_equal(a, a[::-1, :, :])
    col = a[256, 256, :]
    assert np.all(np.diff(col) >= 0.0) and col[-1] > 0.1 and col[0] == 0.0
    assert sim.launch_count() >= 40
    sim.close()


@pytest.mark.parametrize("halo", ["direct", "exchange"])
@pytest.mark.parametrize("total,nprocs,dims", [((37, 23), 6, None), ((41, 19), 3, (3, 1)), ((33, 21, 19), 8, None), ((150, 18, 9), 12, None),
                                               ((17, 13, 29), 3, (1, 1, 3))])
def test_fused_steps_with_direct_halo_stores_match_the_oracle(total, nprocs, dims, halo):
    """mglc_jacobi_step on P subdomains: the sweep stores its boundary values into the neighbours' ghost layers (default) or the
    halos go through exchange_message first (LAP:94-103 as written).  Either way every interior cell equals the 1-rank oracle
    bit for bit, across several step() calls with check_diff(), a download and an explicit exchange_message() in between."""
    wd, sim = orc.JacobiWorld(total, 1), mg.Jacobi(total, nprocs=nprocs, dims=dims)
    assert sim.direct_halo_available()
    sim.set_halo(halo)
    wd.init(); sim.init()
    rng = np.random.default_rng(9)
    glob = rng.random(tuple(n + 2 for n in total))
    wd.array(0, "A")[...] = glob; wd.array(0, "A_new")[...] = glob
    for r, inf in enumerate(sim.info):
        sl = tuple(slice(s, s + n + 2) for s, n in zip(inf["start"], inf["n"]))
        sim.upload(r, A=glob[sl], A_new=glob[sl])
    for n in (1, 4, 2):
        wd.step(n); sim.step(n)
        assert np.array_equal(sim.gather(), wd.gather())
        assert sim.check_diff() == wd.check_diff()
    sim.exchange_message(); wd.exchange_message()
    wd.step(3); sim.step(3)
    assert np.array_equal(sim.gather(), wd.gather())
    sim.sync()
    wd.close(); sim.close()
