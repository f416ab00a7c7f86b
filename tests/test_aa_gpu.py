"""GPU parity tests of the AA-pattern (single-lattice) D3Q19 lid-driven cavity path (mglc_aa_* through the C ABI) against the
CPU oracle (oracle/lid3d.c) and against the ping-pong path (mglc_lbm_step): strict arithmetic bit-exact for every way a run can
start and end, fast arithmetic within the north-star tolerance (<= 1e-12 relative L2, <= 1e-10 max pointwise)."""
import numpy as np
import pytest

import mglc_b200 as mg
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
REL_L2, MAX_ABS = 1e-12, 1e-10
FIELDS = ("rho", "u", "v", "w")


def close_enough(got, want):
    d = np.linalg.norm((got - want).ravel()) / max(np.linalg.norm(want.ravel()), 1e-300)
    return d <= REL_L2 and np.abs(got - want).max() <= MAX_ABS


def seeded(wd, seed):
    """perturbed populations and fields that are NOT their moments: the first collision() must use the stored rho,u,v,w"""
    rng = np.random.default_rng(seed)
    R = wd.ranks[0]
    R.f[...] *= 1.0 + 0.05 * rng.uniform(-1, 1, R.f.shape)
    R.rho[...] = 1.0 + 0.02 * rng.uniform(-1, 1, R.rho.shape)
    for k in ("u", "v", "w"):
        getattr(R, k)[...] = 0.05 * rng.uniform(-1, 1, R.rho.shape)


@pytest.mark.parametrize("calls", [[1], [2], [3], [4], [1, 1, 1, 1], [2, 1, 2], [3, 3], [5, 2, 1], [1, 4, 1]])
@pytest.mark.parametrize("collision", ["mrt", "bgk"])
def test_strict_is_bit_exact_for_every_way_a_run_starts_and_ends(calls, collision):
    total = (34, 9, 7)
    wd = orc.LidWorld(total, 1, collision=collision)
    wd.initial()
    seeded(wd, 5)
    sim = mg.LidDrivenCavityAA(total, arith="strict", collision=collision)
    assert sim.tauf == wd.tauf
    R = wd.ranks[0]
    sim.upload(R.f, R.rho, R.u, R.v, R.w)
    for n in calls:
        wd.step(n); sim.step(n)
        m = sim.download_macro()
        for k in FIELDS:
            assert np.array_equal(m[k], wd.gather(k)), (calls, n, k)
        assert np.array_equal(sim.download_f(), wd.gather("f")), (calls, n)
    assert np.isclose(sim.check(), wd.check(), rtol=1e-13, atol=0)
    sim.close(); wd.close()


@pytest.mark.parametrize("total", [(65, 65, 65), (130, 5, 3), (1, 1, 1), (3, 300, 2)])
def test_config1_strict_from_initial(total):
    """config 1 (65^3, Re = 1000, U0 = 0.1) and degenerate shapes from the reference's own initial(): bit-exact"""
    wd = orc.LidWorld(total, 1)
    wd.initial()
    sim = mg.LidDrivenCavityAA(total, arith="strict")
    sim.initial()
    assert np.array_equal(sim.download_f(), wd.gather("f"))
    done = 0
    for n in (1, 10, 25):
        wd.step(n - done); sim.step(n - done); done = n
        m = sim.download_macro()
        for k in FIELDS:
            assert np.array_equal(m[k], wd.gather(k)), (n, k)
    assert np.array_equal(sim.download_f(), wd.gather("f"))
    sim.close(); wd.close()


def test_config1_fast_within_tolerance():
    """config 1: 65^3, N in {1, 10, 100, 2000}, fast arithmetic, plus check() at step 2000"""
    total = (65, 65, 65)
    wd = orc.LidWorld(total, 1)
    wd.initial()
    sim = mg.LidDrivenCavityAA(total, arith="fast")
    sim.initial()
    done = 0
    for n in (1, 10, 100, 2000):
        wd.step(n - done); sim.step(n - done); done = n
        m = sim.download_macro()
        for k in FIELDS:
            assert close_enough(m[k], wd.gather(k)), (n, k)
    assert np.isclose(sim.check(), wd.check(), rtol=1e-9)
    sim.close(); wd.close()


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_same_fields_as_the_ping_pong_path(arith):
    """one lattice or two: bit for bit in the strict build; in the fast build the compiler may contract the shared arithmetic
    differently inside different kernels, so the two paths agree to the north-star tolerance"""
    total = (70, 33, 18)
    a = mg.LidDrivenCavityAA(total, arith=arith)
    b = mg.LidDrivenCavity(total, arith=arith)
    a.initial(); b.initial()
    for n in (7, 12):
        a.step(n); b.step(n)
        ma, mb = a.download_macro(), b.gather_macro()
        for k in FIELDS:
            assert np.array_equal(ma[k], mb[k]) if arith == "strict" else close_enough(ma[k], mb[k]), (arith, n, k)
    a.close(); b.close()


def test_large_lattice_properties():
    """512^3 on one lattice (no oracle run): total mass conserved to rounding, mirror symmetry about the y mid-plane (the lid
    moves along x), causality (cells farther than N from the lid are exactly at rest), and agreement with the ping-pong path"""
    total, n = (512, 512, 512), 10
    a = mg.LidDrivenCavityAA(total, arith="fast")
    a.initial()
    a.step(n)
    m = a.download_macro()
    bytes_one = a.device_bytes()
    a.close()
    assert abs(m["rho"].sum() - 512.0 ** 3) / 512.0 ** 3 < 1e-13
    assert np.abs(m["u"] - m["u"][:, ::-1, :]).max() < 1e-13 and np.abs(m["v"] + m["v"][:, ::-1, :]).max() < 1e-13
    assert np.all(m["u"][:, :, : 512 - n - 1] == 0.0) and np.abs(m["rho"][:, :, : 512 - n - 1] - 1.0).max() < 1e-14
    assert np.abs(m["u"][:, :, -1]).max() > 0.01
    b = mg.LidDrivenCavity(total, arith="fast")
    b.initial(); b.step(n)
    mb = b.gather_macro()
    assert bytes_one < 0.6 * b.device_bytes()          # one lattice + fields against two lattices + fields
    b.close()
    for k in FIELDS:
        assert close_enough(m[k], mb[k]), k


def test_error_behaviour():
    with pytest.raises(mg.MglcError):
        mg.LidDrivenCavityAA((0, 4, 4))
    with pytest.raises(mg.MglcError):
        mg.LidDrivenCavityAA((8, 8, 8), Re=-1.0)
    sim = mg.LidDrivenCavityAA((8, 8, 8))
    sim.initial()
    with pytest.raises(mg.MglcError):
        sim.step(-1)
    with pytest.raises(ValueError):
        sim.upload(rho=np.zeros((3, 3, 3)))
    sim.step(1)                                       # the lattice now sits between two streaming steps
    with pytest.raises(mg.MglcError):
        sim.upload(rho=np.ones((8, 8, 8)))            # fields alone cannot be replaced in that state
    sim.close()


# ---- decomposed lattices: blocks in one process storing into each other (mglc_aa_group_*) --------------------------------------
@pytest.mark.parametrize("nranks,dims,total", [(2, None, (34, 9, 7)), (4, None, (21, 18, 17)), (8, None, (19, 18, 17)),
                                               (12, (2, 2, 3), (15, 14, 13)), (3, (3, 1, 1), (3, 5, 4)), (6, (1, 2, 3), (130, 7, 9))])
@pytest.mark.parametrize("collision", ["mrt", "bgk"])
def test_group_blocks_strict_bit_exact(nranks, dims, total, collision):
    """uneven blocks (down to one cell thick) on one device: every way a run starts and ends, f from either layout, check()"""
    wd = orc.LidWorld(total, 1, collision=collision)
    wd.initial()
    seeded(wd, 7)
    sim = mg.LidDrivenCavityAA(total, arith="strict", collision=collision, nranks=nranks, dims=dims)
    assert len(sim.blocks) == nranks and sim.tauf == wd.tauf
    R = wd.ranks[0]
    sim.upload(R.f, R.rho, R.u, R.v, R.w)
    for n in (1, 2, 1, 3, 5):
        wd.step(n); sim.step(n)
        m = sim.download_macro()
        for k in FIELDS:
            assert np.array_equal(m[k], wd.gather(k)), (nranks, n, k)
        assert np.array_equal(sim.download_f(), wd.gather("f")), (nranks, n)
    assert np.isclose(sim.check(), wd.check(), rtol=1e-13, atol=0)
    sim.close(); wd.close()


def test_group_from_initial_and_reinitialised():
    """the reference's own start on 2x2x2 blocks; initial() again in the middle of a run starts over (no stale parked halos)"""
    total = (33, 31, 30)
    wd = orc.LidWorld(total, 1)
    wd.initial(); wd.step(25)
    sim = mg.LidDrivenCavityAA(total, arith="strict", nranks=8)
    sim.initial(); sim.step(3)
    sim.initial(); sim.step(24); sim.step(1)
    m = sim.download_macro()
    for k in FIELDS:
        assert np.array_equal(m[k], wd.gather(k)), k
    assert np.array_equal(sim.download_f(), wd.gather("f"))
    assert sim.launch_count() > 0
    sim.close(); wd.close()


def test_group_fast_equals_one_block_fast():
    """fast arithmetic: the decomposed run does the same per-cell arithmetic as the one-block run -- identical results"""
    total = (40, 37, 35)
    one = mg.LidDrivenCavityAA(total, arith="fast")
    grp = mg.LidDrivenCavityAA(total, arith="fast", nranks=4)
    one.initial(); grp.initial()
    one.step(60); grp.step(60)
    a, b = one.download_macro(), grp.download_macro()
    for k in FIELDS:
        assert np.array_equal(a[k], b[k]), k
    one.close(); grp.close()


def test_group_member_calls_are_refused():
    import ctypes as C
    from mglc_b200 import _lib as L
    sim = mg.LidDrivenCavityAA((12, 11, 10), nranks=2)
    h = sim.blocks[0]._h
    assert L.lib().mglc_aa_step(h, 1) == L.E_STATE
    assert L.lib().mglc_aa_initial(h) == L.E_STATE
    e = C.c_double()
    assert L.lib().mglc_aa_check(h, C.byref(e)) == L.E_STATE
    sim.initial(); sim.step(1)                          # POST layout now: a block upload is refused
    with pytest.raises(L.MglcError):
        sim.blocks[0].upload(rho=np.ones(sim.blocks[0].n, order="F"))
    sim.close()


def test_group_on_two_devices_if_present():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one device")
    total = (48, 33, 31)
    wd = orc.LidWorld(total, 1)
    wd.initial(); wd.step(20)
    sim = mg.LidDrivenCavityAA(total, arith="strict", nranks=2, devices=[0, 1])
    sim.initial(); sim.step(20)
    m = sim.download_macro()
    for k in FIELDS:
        assert np.array_equal(m[k], wd.gather(k)), k
    sim.close(); wd.close()
