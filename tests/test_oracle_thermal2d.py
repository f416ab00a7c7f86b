"""Pins the 2-D thermal (D2Q9 + D2Q5) oracle (oracle/thermal2d.c) to the reference's Fortran source text
(MPI/Buoyancy_driven_cavity/fortran/2d/mpi_blocked/, machine-evaluated by tests/golden/make_golden_thermal2d.py into
tests/golden/ref_fortran_thermal2d.npz), to analytic known answers, and to the reference's seq == MPI contract."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "ref_fortran_thermal2d.npz"))
PNAMES = ("tauf", "viscosity", "diffusivity", "paraA", "gBeta", "Snu", "Sq", "Qd", "Qnu")
EX = np.array([0, 1, 0, -1, 0, 1, -1, -1, 1])
EY = np.array([0, 0, 1, 0, -1, 1, 1, -1, -1])
SIDE, RB = orc.T2_SIDE_HEATED, orc.T2_RAYLEIGH_BENARD


def test_parameters_match_the_fortran_text():
    p = orc.t2_params(201)
    assert tuple(getattr(p, k) for k in PNAMES) == tuple(GOLD["params/201"])
    p = orc.t2_params(64, Rayleigh=1e6)
    assert tuple(getattr(p, k) for k in PNAMES) == tuple(GOLD["params/64_ra1e6"])
    # module.F90:73 / initial.F90:30: the shipped case must satisfy -4 < paraA < 1
    assert -4.0 < orc.t2_params(201).paraA < 1.0


def test_collision_cells_bit_exact():
    p = orc.t2_params(201)
    f, r = GOLD["cells/f"], GOLD["cells/ruvT"]
    for k in range(len(f)):
        fp, F2 = orc.t2_collide_cell(p, f[k], *r[k])
        assert np.array_equal(fp, GOLD["collision/f_post"][k]), k
        assert np.array_equal(F2, GOLD["collision/FxFy"][k]), k
        assert np.array_equal(np.signbit(fp), np.signbit(GOLD["collision/f_post"][k]))


def test_collisionT_cells_bit_exact():
    p = orc.t2_params(201)
    g, r = GOLD["cells/g"], GOLD["cells/ruvT"]
    for k in range(len(g)):
        gp = orc.t2_collideT_cell(p, g[k], r[k][1], r[k][2], r[k][3])
        assert np.array_equal(gp, GOLD["collisionT/g_post"][k]), k


def test_macro_macroT_cells_bit_exact_including_the_sign_of_zero():
    f, g, F2 = GOLD["cells/f"], GOLD["cells/g"], GOLD["cells/FxFy"]
    w = orc.Thermal2DWorld((len(f), 1))
    R = w.ranks[0]
    R.f[:, :, 0], R.g[:, :, 0] = f.T, g.T
    R.Fx[:, 0], R.Fy[:, 0] = F2[:, 0], F2[:, 1]
    w.macro(); w.macroT()
    got = np.stack([R.rho[:, 0], R.u[:, 0], R.v[:, 0]], axis=1)
    assert np.array_equal(got, GOLD["macro/ruv"])
    assert np.array_equal(np.signbit(got), np.signbit(GOLD["macro/ruv"]))
    assert np.array_equal(R.T[:, 0], GOLD["macro/T"])
    w.close()


def test_initial_weights_profile_and_equilibria():
    w = orc.Thermal2DWorld((201, 201))
    w.initial()
    R = w.ranks[0]
    p = w.params
    # rest fluid: f = rho0 * omega, g = T * omegaT
    assert np.array_equal(R.f[:, 5, 7], GOLD["initial/omega"] * 1.0)
    args = GOLD["initial/T_profile_args"]
    w2 = orc.Thermal2DWorld((201, 201), nprocs=6, dims=(3, 2))
    w2.initial()
    full = w2.gather("T")
    for (s, i), want in zip(args, GOLD["initial/T_profile_201"]):
        assert full[s + i - 1, 0] == want and R.T[s + i - 1, 100] == want
    assert np.array_equal(full, w.gather("T"))
    k = 17
    assert np.array_equal(R.g[:, k, 3], R.T[k, 3] * GOLD["initial/omegaT"] * (1.0 + 10.0 / (4.0 + p.paraA) * 0.0))
    w.close(); w2.close()
    # y profile (HorizontalWallsConstT) with non-default wall temperatures
    w = orc.Thermal2DWorld((9, 77), bcT=RB, Thot=0.5, Tcold=-0.5)
    w.initial()
    for (s, j), want in zip([(0, 1), (0, 39), (39, 38), (20, 7)], GOLD["initial/T_profile_y_77"]):
        assert w.ranks[0].T[4, s + j - 1] == want
    w.close()
    # the equilibrium formulas with non-zero velocity: re-evaluate the oracle's own loop through a one-row world
    f, r = GOLD["cells/f"], GOLD["cells/ruvT"]
    p = orc.t2_params(201)
    om, omT = GOLD["initial/omega"], GOLD["initial/omegaT"]
    for k in range(len(f)):
        rho, u, v, T = r[k]
        us2 = u * u + v * v
        for a in range(9):
            un = u * float(EX[a]) + v * float(EY[a])
            assert rho * om[a] * (1.0 + 3.0 * un + 4.5 * un * un - 1.5 * us2) == GOLD["initial/feq"][k][a]
        for a in range(5):
            un = u * float(EX[a]) + v * float(EY[a])
            assert T * omT[a] * (1.0 + 10.0 / (4.0 + p.paraA) * un) == GOLD["initial/geq"][k][a]


def _block_world(coords, dims, bcT):
    """a world whose rank at `coords` of a dims[0] x dims[1] grid is the golden 6 x 5 block"""
    nx, ny = 6, 5
    w = orc.Thermal2DWorld((nx * dims[0], ny * dims[1]), nprocs=dims[0] * dims[1], dims=dims, bcT=bcT)
    R = w.ranks[coords[0] * dims[1] + coords[1]]
    assert R.n == (nx, ny) and R.coords == tuple(coords)
    R.f_post[...] = GOLD["field/f_post"]; R.g_post[...] = GOLD["field/g_post"]
    R.f[...] = GOLD["field/f0"]; R.g[...] = GOLD["field/g0"]
    return w, R


def test_streaming_and_streamingT_bit_exact():
    w, R = _block_world((0, 0), (1, 1), SIDE)
    w.streaming(); w.streamingT()
    assert np.array_equal(R.f, GOLD["field/streaming_f"])
    assert np.array_equal(R.g, GOLD["field/streamingT_g"])
    w.close()


@pytest.mark.parametrize("case", range(5))
def test_bounceback_and_bouncebackT_bit_exact(case):
    c = GOLD["field/bb_cases"][case]
    coords, dims = tuple(int(x) for x in c[:2]), tuple(int(x) for x in c[2:])
    for tag, bcT in (("side", SIDE), ("rb", RB)):
        w, R = _block_world(coords, dims, bcT)
        w.bounceback(); w.bouncebackT()
        assert np.array_equal(R.f, GOLD[f"field/bounceback_{case}"]), (case, tag)
        assert np.array_equal(R.g, GOLD[f"field/bouncebackT_{tag}_{case}"]), (case, tag)
        w.close()
    assert not np.array_equal(GOLD[f"field/bouncebackT_side_{case}"], GOLD[f"field/bouncebackT_rb_{case}"])


def test_check_sums_bit_exact():
    w = orc.Thermal2DWorld((6, 5))
    R = w.ranks[0]
    for k in ("u", "v", "up", "vp", "T", "Tp"):
        getattr(R, k)[...] = GOLD[f"field/check_{k}"]
    eu, et = w.check()
    e1, e2, e5, e6 = GOLD["field/check_sums"]
    assert eu == np.sqrt(e1) / np.sqrt(e2) and et == e5 / e6
    assert np.array_equal(R.up, GOLD["field/check_up_after"]) and np.array_equal(R.Tp, R.T)
    w.close()


# ---------------- analytic known answers ----------------
def test_moment_matrices_invert():
    """M^-1 M = I for the D2Q9 and D2Q5 transforms as the collision routines spell them (probing with unit vectors,
    all relaxation rates zeroed by feeding the equilibrium-independent path: tau -> infinity is not reachable, so
    compare against explicit matrices built from the reference's rows)"""
    M9 = np.array([[1, 1, 1, 1, 1, 1, 1, 1, 1], [-4, -1, -1, -1, -1, 2, 2, 2, 2], [4, -2, -2, -2, -2, 1, 1, 1, 1],
                   [0, 1, 0, -1, 0, 1, -1, -1, 1], [0, -2, 0, 2, 0, 1, -1, -1, 1], [0, 0, 1, 0, -1, 1, 1, -1, -1],
                   [0, 0, -2, 0, 2, 1, 1, -1, -1], [0, 1, -1, 1, -1, 0, 0, 0, 0], [0, 0, 0, 0, 0, 1, -1, 1, -1]], float)
    p = orc.t2_params(201)
    rng = np.random.default_rng(5)
    f = rng.random(9)
    rho, u, v, T = 1.03, 0.02, -0.01, 0.4
    fp, F2 = orc.t2_collide_cell(p, f, rho, u, v, T)
    m = M9 @ f
    meq = rho * np.array([1, -2 + 3 * (u * u + v * v), 1 - 3 * (u * u + v * v), u, -u, v, -v, u * u - v * v, u * v])
    s = np.array([0, p.Snu, p.Snu, 0, p.Sq, 0, p.Sq, p.Snu, p.Snu])
    Fy = rho * p.gBeta * (T - p.Tref)
    src = np.array([0, (6 - 3 * p.Snu) * v * Fy, -(6 - 3 * p.Snu) * v * Fy, 0, 0, Fy, -(1 - 0.5 * p.Sq) * Fy, -(2 - p.Snu) * v * Fy,
                    (1 - 0.5 * p.Snu) * u * Fy])
    want = np.linalg.solve(M9, m - s * (m - meq) + src)
    assert np.allclose(fp, want, rtol=0, atol=1e-15)
    assert F2[0] == 0.0 and F2[1] == Fy
    N5 = np.array([[1, 1, 1, 1, 1], [0, 1, 0, -1, 0], [0, 0, 1, 0, -1], [-4, 1, 1, 1, 1], [0, 1, -1, 1, -1]], float)
    g = rng.random(5)
    gp = orc.t2_collideT_cell(p, g, u, v, T)
    n = N5 @ g
    q = np.array([0, p.Qd, p.Qd, p.Qnu, p.Qnu])
    want = np.linalg.solve(N5, n - q * (n - np.array([T, T * u, T * v, T * p.paraA, 0.0])))
    assert np.allclose(gp, want, rtol=0, atol=1e-15)


def test_collisions_conserve_mass_and_temperature_and_add_the_force():
    p = orc.t2_params(201)
    rng = np.random.default_rng(6)
    for _ in range(20):
        f, g = 0.1 + rng.random(9), rng.random(5)
        rho = f.sum()
        u, v, T = 0.05 * rng.uniform(-1, 1), 0.05 * rng.uniform(-1, 1), rng.random()
        fp, F2 = orc.t2_collide_cell(p, f, rho, u, v, T)
        assert abs(fp.sum() - f.sum()) < 1e-14
        # momentum after collision = momentum before + F (the moments 3 and 5 relax with s = 0 and gain (1 - 0/2) F)
        assert abs((fp * EX).sum() - (f * EX).sum() - F2[0]) < 1e-15 and abs((fp * EY).sum() - (f * EY).sum() - F2[1]) < 1e-15
        gp = orc.t2_collideT_cell(p, g, u, v, T)
        assert abs(gp.sum() - g.sum()) < 1e-15


def test_isothermal_rest_state_is_a_fixed_point():
    """T = Tref everywhere with adiabatic walls: no buoyancy, nothing moves, bit for bit"""
    w = orc.Thermal2DWorld((17, 13), bcT=(0, 0, 0, 0))
    w.initial()
    f0, g0 = w.gather("f").copy(), w.gather("g").copy()
    w.step(25)
    assert np.array_equal(w.gather("u"), np.zeros((17, 13))) and np.array_equal(w.gather("v"), np.zeros((17, 13)))
    assert np.allclose(w.gather("f"), f0, rtol=0, atol=1e-14) and np.array_equal(w.gather("g"), g0)
    w.close()


def test_total_mass_is_conserved_and_wall_halos_are_never_used():
    w = orc.Thermal2DWorld((33, 29))
    w.initial()
    m0 = w.gather("f").sum()
    w.step(50)
    assert abs(w.gather("f").sum() - m0) < 1e-10
    ref = {k: w.gather(k).copy() for k in ("rho", "u", "v", "T")}
    w.close()
    # poison every wall halo after each collision: bounceback()/bouncebackT() must overwrite whatever streaming pulled in
    w = orc.Thermal2DWorld((33, 29))
    w.initial()
    R = w.ranks[0]
    for _ in range(50):
        w.collision()
        R.f_post[:, 0, :] = R.f_post[:, -1, :] = R.f_post[:, :, 0] = R.f_post[:, :, -1] = np.nan
        w.streaming(); w.bounceback(); w.collisionT()
        R.g_post[:, 0, :] = R.g_post[:, -1, :] = R.g_post[:, :, 0] = R.g_post[:, :, -1] = np.nan
        w.streamingT(); w.bouncebackT(); w.macro(); w.macroT()
    for k, a in ref.items():
        assert np.array_equal(w.gather(k), a), k
    w.close()


def test_side_heated_cell_develops_the_expected_circulation():
    """hot left wall: fluid rises along it (v > 0 near i = 1), sinks along the cold wall; T stays within [Tcold, Thot] + overshoot"""
    w = orc.Thermal2DWorld((41, 41), Rayleigh=1e5)
    w.initial()
    w.step(2000)
    v, T = w.gather("v"), w.gather("T")
    assert v[1, 20] > 0 and v[-2, 20] < 0
    assert T.min() > -0.05 and T.max() < 1.05
    eu, et = w.check()
    assert np.isfinite(eu) and np.isfinite(et)
    w.close()


@pytest.mark.parametrize("nprocs,dims", [(2, None), (4, None), (6, (2, 3)), (6, (3, 2)), (9, (3, 3))])
@pytest.mark.parametrize("bcT", [SIDE, RB])
def test_decomposition_invariance_bit_exact(nprocs, dims, bcT):
    """the reference's seq == MPI contract: P ranks reproduce one rank bit for bit (uneven blocks: 37 x 31)"""
    one = orc.Thermal2DWorld((37, 31), bcT=bcT, Rayleigh=1e6)
    many = orc.Thermal2DWorld((37, 31), nprocs=nprocs, dims=dims, bcT=bcT, Rayleigh=1e6)
    one.initial(); many.initial()
    one.step(60); many.step(60)
    for k in ("rho", "u", "v", "T", "f", "g", "Fy"):
        assert np.array_equal(one.gather(k), many.gather(k)), k
    assert one.check() == pytest.approx(many.check(), rel=1e-13)
    a, b = one.nure_sums(), many.nure_sums()
    assert np.allclose(a, b, rtol=1e-12, atol=1e-12)      # the angular momentum is a cancelling sum
    one.close(); many.close()


def test_exchange_moves_exactly_the_reference_messages():
    """message_exchange.F90: f faces carry 3 populations over the interior range, corners 1 value; g faces carry 1"""
    w = orc.Thermal2DWorld((8, 6), nprocs=4, dims=(2, 2))
    rng = np.random.default_rng(9)
    for R in w.ranks:
        R.f_post[...] = rng.random(R.f_post.shape); R.g_post[...] = rng.random(R.g_post.shape)
    before = [(R.f_post.copy(), R.g_post.copy()) for R in w.ranks]
    w.message_passing_f(); w.message_passing_g()
    R00, R10, R01, R11 = w.ranks[0], w.ranks[2], w.ranks[1], w.ranks[3]
    nx, ny = R00.n
    # to the right (+x): populations 1, 5, 8 from i = nx to the neighbour's i = 0, rows 1..ny only
    for a in (1, 5, 8):
        assert np.array_equal(R10.f_post[a, 0, 1:-1], before[0][0][a, nx, 1:-1])
    assert np.array_equal(R10.g_post[1, 0, 1:-1], before[0][1][1, nx, 1:-1])
    # corner: population 5 from (nx, ny) of rank (0,0) to (0, 0) of rank (1,1)
    assert R11.f_post[5, 0, 0] == before[0][0][5, nx, ny]
    # everything not named by the reference is untouched: count changed entries on rank (1,1)
    changed_f = (R11.f_post != before[3][0]).sum()
    changed_g = (R11.g_post != before[3][1]).sum()
    assert changed_f == 3 * R11.n[1] + 3 * R11.n[0] + 1 and changed_g == R11.n[1] + R11.n[0]
    w.close()


# ---------------- the OpenACC program's flavour (seq/bouyancy2d_acc.F90) ----------------
def test_acc_parameters_and_collisions_bit_exact():
    """nx = 513, ny = 257, lengthUnit = dble(nx), Ra = 1e5 (acc:55-60): constants, collision (f_post(0) rounded term by term)
    and collisionT against the OpenACC program's own text"""
    p = orc.t2_params(257, Rayleigh=1e5, lengthUnit=513.0)
    assert tuple(getattr(p, k) for k in PNAMES) == tuple(GOLD["acc/params"])
    f, g, r = GOLD["cells/f"], GOLD["cells/g"], GOLD["cells/ruvT"]
    differs = 0
    for k in range(len(f)):
        fp, F2 = orc.t2_collide_cell(p, f[k], *r[k], variant="acc")
        assert np.array_equal(fp, GOLD["acc/collision_f_post"][k]) and np.array_equal(F2, GOLD["acc/collision_FxFy"][k]), k
        differs += not np.array_equal(fp, orc.t2_collide_cell(p, f[k], *r[k], variant="mpi")[0])
        assert np.array_equal(orc.t2_collideT_cell(p, g[k], r[k][1], r[k][2], r[k][3]), GOLD["acc/collisionT_g_post"][k]), k
    assert differs > 0          # the two programs really round f_post(0) differently


def test_acc_periodic_walls_bit_exact():
    """streaming / bounceback / streamingT / bouncebackT of the OpenACC program with its shipped macro set: vertical walls
    periodic (the SAME row of the opposite column, also for the diagonal populations), constant-temperature plates"""
    nx, ny = 6, 5
    w = orc.Thermal2DWorld((nx, ny), bcT=orc.T2_RB_PERIODIC, variant="acc", lengthUnit=513.0, Rayleigh=1e5)
    assert w.params.paraA == GOLD["acc/params"][3]
    R = w.ranks[0]
    R.f_post[...] = GOLD["field/f_post"]; R.g_post[...] = GOLD["field/g_post"]
    w.streaming(); w.streamingT()
    assert np.array_equal(R.f, GOLD["acc/streaming_f"]) and np.array_equal(R.g, GOLD["acc/streamingT_g"])
    R.f[...] = GOLD["field/f0"]; R.g[...] = GOLD["field/g0"]
    w.bounceback(); w.bouncebackT()
    assert np.array_equal(R.f, GOLD["acc/bounceback_f"]) and np.array_equal(R.g, GOLD["acc/bouncebackT_g"])
    w.close()


def test_acc_periodic_run_splits_along_y_bit_exactly():
    """(the reference's periodic rule takes the diagonal populations from the SAME row of the opposite column, so in the four
    corner cells one population is duplicated and one dropped: total mass drifts by O(1e-5) per step -- reproduced, not fixed)"""
    one = orc.Thermal2DWorld((33, 17), bcT=orc.T2_RB_PERIODIC, variant="acc", lengthUnit=33.0, Rayleigh=1e5)
    many = orc.Thermal2DWorld((33, 17), nprocs=3, dims=(1, 3), bcT=orc.T2_RB_PERIODIC, variant="acc", lengthUnit=33.0, Rayleigh=1e5)
    one.initial(); many.initial()
    m0 = one.gather("f").sum()
    one.step(80); many.step(80)
    assert 1e-6 < abs(one.gather("f").sum() - m0) < 1e-2
    for k in ("rho", "u", "v", "T", "f", "g"):
        assert np.array_equal(one.gather(k), many.gather(k)), k
    # (the same-row rule also breaks translation invariance: columns 1 and nx differ from the interior columns)
    T = one.gather("T")
    assert np.all(T[2:-2, :] == T[2:3, :]) or not np.all(T == T[0:1, :])
    one.close(); many.close()


# ---------------- the sequential side-heated program's own run, from its text (make_golden_thermal2d_seq_run.py) ----------------
SRUN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fortran_thermal2d_seq_run.npz"))


SRUN_RB = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fortran_thermal2d_seq_run_rb.npz"))


@pytest.mark.parametrize("macro_set", ["side_heated", "rb"])
@pytest.mark.parametrize("nprocs,dims", [(1, None), (2, None), (4, (2, 2)), (6, (3, 2)), (3, (1, 3))])
def test_oracle_reproduces_the_sequential_side_heated_programs_run(nprocs, dims, macro_set):
    """seq/steady.F90 with its shipped macro set (side-heated cell, no-slip walls at rest), evaluated from its text on 9 x 7:
    parameters, initial() (T linear in x) and its loop of eight subroutines for 1, 2, 20 and 25 iterations, check() after 20 and
    25 -- including its separate corner statements in bounceback().  The restatement of the MPI program (variant "mpi", the
    side-heated set: the row INTEGRATION.md maps this file to) reproduces f, g, rho, u, v, T, Fx, Fy bit for bit on 1 to 6 ranks."""
    # "rb": the same file with its Rayleigh-Benard macro set (steady.F90:22-24) switched on instead of the side-heated one
    SRUN = SRUN_RB if macro_set == "rb" else globals()["SRUN"]
    total = tuple(int(x) for x in SRUN["shape"])
    wd = orc.Thermal2DWorld(total, nprocs, dims, bcT=orc.T2_RAYLEIGH_BENARD if macro_set == "rb" else orc.T2_SIDE_HEATED, variant="mpi")
    assert tuple(getattr(wd.params, k) for k in ("tauf", "viscosity", "diffusivity", "paraA", "gBeta", "Snu", "Sq", "Qd", "Qnu")) == tuple(SRUN["params"])
    wd.initial()

    def same(tag):
        assert np.array_equal(wd.gather("f"), SRUN[tag + "/f"]), tag
        assert np.array_equal(wd.gather("g"), SRUN[tag + "/g"]), tag
        assert np.array_equal(np.stack([wd.gather(k) for k in ("rho", "u", "v", "T")]), SRUN[tag + "/ruvT"]), tag
        assert np.array_equal(np.stack([wd.gather(k) for k in ("Fx", "Fy")]), SRUN[tag + "/F"]), tag

    same("run0")
    done = 0
    for n in (1, 2, 20):
        wd.step(n - done); done = n
        same(f"run{n}")
    tol = 0 if nprocs == 1 else 1e-14
    eu, et = wd.check()
    assert abs(eu - SRUN["run20/check"][0]) <= tol * abs(eu) and abs(et - SRUN["run20/check"][1]) <= tol * abs(et)
    wd.step(5)
    eu, et = wd.check()
    assert abs(eu - SRUN["run25/check"][0]) <= tol * abs(eu) and abs(et - SRUN["run25/check"][1]) <= tol * abs(et)
    same("run25")
    wd.close()


SRUN_SHEAR = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fortran_thermal2d_seq_run_sheared_rb.npz"))


@pytest.mark.parametrize("nprocs,dims", [(1, None), (2, None), (4, (2, 2)), (6, (3, 2)), (3, (1, 3)), (12, (4, 3))])
def test_oracle_reproduces_the_sheared_rayleigh_benard_programs_run(nprocs, dims):
    """seq/R_B_2d.F90 AS SHIPPED (walls moving at shearReynolds = 100, Pr = 5.3, Rayleigh-Benard plates, its own corner cells in
    bouncebackT()), evaluated from its text on 9 x 7: initial() with the wall velocities, 25 iterations, check().  The
    restatement with Uwall and cornersT reproduces f, g, rho, u, v, T, Fx, Fy bit for bit on 1 to 12 ranks (the halves of the
    walls meet at nxHalf / nyHalf of the global lattice, inside a subdomain or on a subdomain boundary)."""
    S = SRUN_SHEAR
    total = tuple(int(x) for x in S["shape"])
    wd = orc.Thermal2DWorld(total, nprocs, dims, bcT=orc.T2_RAYLEIGH_BENARD, variant="mpi", Prandtl=float(S["prandtl"]),
                            Uwall=[float(x) for x in S["uwall"]], cornersT=True)
    assert tuple(getattr(wd.params, k) for k in ("tauf", "viscosity", "diffusivity", "paraA", "gBeta", "Snu", "Sq", "Qd", "Qnu")) == tuple(S["params"])
    # U0 = shearReynolds*viscosity/dble(ny), R_B_2d.F90:118
    assert S["uwall"][0] == 100.0 * wd.params.viscosity / float(total[1])
    wd.initial()

    def same(tag):
        assert np.array_equal(wd.gather("f"), S[tag + "/f"]), tag
        assert np.array_equal(wd.gather("g"), S[tag + "/g"]), tag
        assert np.array_equal(np.stack([wd.gather(k) for k in ("rho", "u", "v", "T")]), S[tag + "/ruvT"]), tag
        assert np.array_equal(np.stack([wd.gather(k) for k in ("Fx", "Fy")]), S[tag + "/F"]), tag

    same("run0")
    done = 0
    for n in (1, 2, 20):
        wd.step(n - done); done = n
        same(f"run{n}")
    tol = 0 if nprocs == 1 else 1e-14
    eu, et = wd.check()
    assert abs(eu - S["run20/check"][0]) <= tol * abs(eu) and abs(et - S["run20/check"][1]) <= tol * abs(et)
    wd.step(5)
    same("run25")
    wd.close()


def test_the_openmp_program_of_the_family_runs_identically():
    """seq/bouyancy2d_omp.F90 as shipped (its irregular-geometry / tilted-cell / tracer-particle branches compiled out or not
    reached), evaluated from ITS text: the same numbers as seq/R_B_2d.F90's run, array for array -- one descriptor covers both"""
    omp = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fortran_thermal2d_seq_run_sheared_omp.npz"))
    assert sorted(omp.keys()) == sorted(SRUN_SHEAR.keys())
    for k in omp.keys():
        assert np.array_equal(omp[k], SRUN_SHEAR[k]), k


def test_walls_at_rest_and_plain_corners_are_the_default():
    """Uwall = 0 and cornersT off change nothing (the shipped MPI program); each of the two options changes the run"""
    total = (9, 7)
    runs = {}
    for name, kw in (("plain", {}), ("zero", dict(Uwall=[0.0] * 8)), ("corners", dict(cornersT=True)), ("moving", dict(Uwall=[1e-3] * 8))):
        wd = orc.Thermal2DWorld(total, 1, None, bcT=orc.T2_RAYLEIGH_BENARD, **kw)
        wd.initial(); wd.step(6)
        runs[name] = wd.gather("f").copy(), wd.gather("g").copy()
        wd.close()
    assert np.array_equal(runs["plain"][0], runs["zero"][0]) and np.array_equal(runs["plain"][1], runs["zero"][1])
    assert not np.array_equal(runs["plain"][1], runs["corners"][1])
    assert not np.array_equal(runs["plain"][0], runs["moving"][0])


ARUN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fortran_thermal2d_acc_run.npz"))


@pytest.mark.parametrize("nprocs,dims", [(1, None), (2, (1, 2)), (3, (1, 3))])
def test_oracle_reproduces_the_openacc_programs_run(nprocs, dims):
    """seq/bouyancy2d_acc.F90 (the reference's only GPU code) with its shipped macro set -- Rayleigh-Benard plates, vertical walls
    periodic for f and g with its same-row rule, lengthUnit = dble(nx), f_post(0) rounded term by term -- evaluated from its
    text as a whole program on 9 x 7: parameters, initial() (T linear in y), 25 iterations of its eight subroutines, check().
    Variant "acc" of the restatement reproduces f, g, rho, u, v, T, Fx, Fy bit for bit on one rank and on ranks stacked along y."""
    total = tuple(int(x) for x in ARUN["shape"])
    wd = orc.Thermal2DWorld(total, nprocs, dims, bcT=orc.T2_RB_PERIODIC, variant="acc", lengthUnit=float(total[0]), Rayleigh=1e5)
    assert tuple(getattr(wd.params, k) for k in ("tauf", "viscosity", "diffusivity", "paraA", "gBeta", "Snu", "Sq", "Qd", "Qnu")) == tuple(ARUN["params"])
    wd.initial()

    def same(tag):
        assert np.array_equal(wd.gather("f"), ARUN[tag + "/f"]), tag
        assert np.array_equal(wd.gather("g"), ARUN[tag + "/g"]), tag
        assert np.array_equal(np.stack([wd.gather(k) for k in ("rho", "u", "v", "T")]), ARUN[tag + "/ruvT"]), tag
        assert np.array_equal(np.stack([wd.gather(k) for k in ("Fx", "Fy")]), ARUN[tag + "/F"]), tag

    same("run0")
    done = 0
    for n in (1, 2, 20):
        wd.step(n - done); done = n
        same(f"run{n}")
    tol = 0 if nprocs == 1 else 1e-14
    eu, et = wd.check()
    assert abs(eu - ARUN["run20/check"][0]) <= tol * abs(eu) and abs(et - ARUN["run20/check"][1]) <= tol * abs(et)
    wd.step(5)
    eu, et = wd.check()
    assert abs(eu - ARUN["run25/check"][0]) <= tol * abs(eu) and abs(et - ARUN["run25/check"][1]) <= tol * abs(et)
    same("run25")
    wd.close()
