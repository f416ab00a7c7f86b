"""CPU tests of the thermal oracle (oracle/thermal3d.c): pinned bit for bit to vectors obtained by
machine-evaluating the reference's own Fortran source (tests/golden/make_golden_fortran.py), plus
construction tests of the copy-only parts (streaming, boundary rules, exchange, decomposition invariance)
and analytic invariants."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fortran_kernels.npz"))
dp = C.POINTER(C.c_double)


def ptr(a):
    return a.ctypes.data_as(dp)


def test_module_parameters_match_reference_source():
    """tauf, paraA, gBeta, omegaRatating, Snu, Sq, Qd, Qnu ... of B3:26-47,73-74 at the shipped 51^3 / Ra=1e6."""
    P = dict(zip([str(n) for n in GOLD["th_params/names"]], GOLD["th_params/values"]))
    p = orc.th_params(total_nz=int(P["total_nx"]), Rayleigh=P["rayleigh"], Prandtl=P["prandtl"], Mach=P["mach"], Ekman=P["ekman"])
    for mine, ref in [("tauf", "tauf"), ("viscosity", "viscosity"), ("diffusivity", "diffusivity"),
                      ("omegaRatating", "omegaratating"), ("paraA", "paraa"), ("gBeta1", "gbeta1"), ("gBeta", "gbeta"),
                      ("Snu", "snu"), ("Sq", "sq"), ("Qd", "qd"), ("Qnu", "qnu")]:
        assert getattr(p, mine) == P[ref], mine


def test_collision_with_forcing_matches_reference_source_bit_for_bit():
    L, p = orc._th_lib(), orc.th_params()
    f, s, want, wantF = (GOLD["th_collision/" + k] for k in ("f", "ruvwT", "f_post", "F"))
    for c in range(f.shape[0]):
        out, Fv = np.empty(19), np.empty(3)
        L.th_collide_cell(ptr(np.ascontiguousarray(f[c])), *[float(x) for x in s[c]], C.byref(p), ptr(out), ptr(Fv))
        assert np.array_equal(out, want[c]), c
        assert np.array_equal(Fv, wantF[c]), c


def test_macro_collisionT_and_equilibria_match_reference_source_bit_for_bit():
    L, p = orc._th_lib(), orc.th_params()
    f, Fv, want = GOLD["th_macro/f"], GOLD["th_macro/F"], GOLD["th_macro/ruvw"]
    for c in range(f.shape[0]):
        out = np.empty(4)
        L.th_macro_cell(ptr(np.ascontiguousarray(f[c])), *[float(x) for x in Fv[c]], ptr(out))
        assert np.array_equal(out, want[c]), c
    g, s, want = GOLD["th_collisionT/g"], GOLD["th_collisionT/uvwT"], GOLD["th_collisionT/g_post"]
    for c in range(g.shape[0]):
        out = np.empty(7)
        L.th_collideT_cell(ptr(np.ascontiguousarray(g[c])), *[float(x) for x in s[c]], C.byref(p), ptr(out))
        assert np.array_equal(out, want[c]), c
    s = GOLD["th_feq/ruvwT"]
    for c in range(s.shape[0]):
        fo, go = np.empty(19), np.empty(7)
        L.th_feq_cell(*[float(x) for x in s[c, :4]], ptr(fo))
        L.th_geq_cell(float(s[c, 4]), *[float(x) for x in s[c, 1:4]], p.paraA, ptr(go))
        assert np.array_equal(fo, GOLD["th_feq/f"][c]) and np.array_equal(go, GOLD["th_feq/g"][c]), c


def test_d3q7_transform_roundtrip_and_conservation():
    """q0 = 0 conserves T = sum g; with all rates 0 collisionT is the identity (N^-1 N = I)."""
    L, p = orc._th_lib(), orc.th_params()
    rng = np.random.default_rng(1)
    for _ in range(20):
        g = rng.random(7)
        out = np.empty(7)
        L.th_collideT_cell(ptr(g), 0.03, -0.02, 0.01, 0.4, C.byref(p), ptr(out))
        assert abs(out.sum() - g.sum()) < 1e-15
        p0 = orc.ThParams.from_buffer_copy(bytes(p)); p0.Qd = 0.0; p0.Qnu = 0.0
        L.th_collideT_cell(ptr(g), 0.03, -0.02, 0.01, 0.4, C.byref(p0), ptr(out))
        assert np.abs(out - g).max() < 5e-16


def test_collision_conserves_mass_and_adds_force_to_momentum():
    L, p = orc._th_lib(), orc.th_params()
    f, s = GOLD["th_collision/f"], GOLD["th_collision/ruvwT"]
    ex, ey, ez = orc.EX, orc.EY, orc.EZ
    for c in range(8):
        out, Fv = np.empty(19), np.empty(3)
        L.th_collide_cell(ptr(np.ascontiguousarray(f[c])), *[float(x) for x in s[c]], C.byref(p), ptr(out), ptr(Fv))
        assert abs(out.sum() - f[c].sum()) < 2e-15
        for e, Fq in ((ex, Fv[0]), (ey, Fv[1]), (ez, Fv[2])):       # s3 = s5 = s7 = 0: m_post = m + F
            assert abs((out * e).sum() - (f[c] * e).sum() - Fq) < 2e-16 + 1e-15


def test_initial_state():
    wd = orc.ThermalWorld((7, 6, 5), 4)
    wd.initial()
    T = wd.gather("T")
    assert np.all(T[:, 0, :] == 1.0) and np.all(T[:, 1:, :] == 0.0)      # hot wall layer j=1; Tcold = 0 at j=ny
    assert np.all(wd.gather("rho") == 1.0) and np.all(wd.gather("u") == 0.0)
    g = wd.gather("g")
    assert np.allclose(g.sum(axis=0), T, atol=1e-15)
    assert np.allclose(wd.gather("f").sum(axis=0), 1.0, atol=1e-15)
    for R in wd.ranks:
        assert np.all(R.f_post == 0.0) and np.all(R.g_post == 0.0)
    wd.close()


def test_rest_state_with_uniform_temperature_is_stationary_without_gravity_coupling():
    """T = Tref everywhere, all walls adiabatic: nothing moves, T stays put."""
    wd = orc.ThermalWorld((6, 5, 4), 1, bcT=[0] * 6)
    wd.initial()
    wd.step(5)
    assert np.abs(wd.gather("u")).max() == 0.0 and np.abs(wd.gather("w")).max() == 0.0
    assert np.all(wd.gather("T") == 0.0)
    assert np.allclose(wd.gather("rho"), 1.0, atol=1e-14)
    wd.close()


def test_adiabatic_walls_conserve_heat_and_constT_walls_drive_it():
    wd = orc.ThermalWorld((8, 7, 6), 2, bcT=[0] * 6)
    wd.initial()
    rng = np.random.default_rng(3)
    T0 = rng.random(wd.total)
    for R in wd.ranks:
        sl = tuple(slice(s, s + n) for s, n in zip(R.start, R.n))
        R.T[...] = T0[sl]
        for a in range(7):
            R.g[a] = T0[sl] * ((1.0 - wd.p.paraA) / 7.0 if a == 0 else (wd.p.paraA + 6.0) / 42.0)
    wd.step(10)
    assert abs(wd.gather("T").sum() - T0.sum()) < 1e-11
    wd.close()
    wd = orc.ThermalWorld((8, 7, 6), 1)            # shipped BCs: hot wall at j=1
    wd.initial()
    wd.step(30)
    T = wd.gather("T")
    m = T.mean(axis=(0, 2))                          # heat diffuses in from the hot wall at j = 1
    assert m[0] > 0.5 and m[1] > 0.01 and np.all(np.abs(m[1:]) < np.abs(m[:-1]))
    assert np.abs(wd.gather("w")).max() > 0.0        # buoyancy set the fluid in motion
    wd.close()


@pytest.mark.parametrize("nprocs,dims", [(2, None), (4, None), (8, None), (3, (1, 3, 1)), (4, (1, 2, 2)), (6, None)])
def test_decomposition_invariance_bit_exact(nprocs, dims):
    total = (9, 8, 7)
    one, many = orc.ThermalWorld(total, 1), orc.ThermalWorld(total, nprocs, dims)
    one.initial(); many.initial()
    one.step(12); many.step(12)
    for k in ("rho", "u", "v", "w", "T", "f", "g", "Fx", "Fy", "Fz"):
        assert np.array_equal(one.gather(k), many.gather(k)), k
    assert one.check() == many.check() or np.allclose(one.check(), many.check(), rtol=1e-13)
    one.close(); many.close()


def test_check_matches_numpy():
    wd = orc.ThermalWorld((6, 5, 7), 1)
    wd.initial()
    wd.step(4)
    u, v, w, T = (wd.gather(k) for k in ("u", "v", "w", "T"))
    eu, et = wd.check()
    assert np.isclose(eu, 1.0) and np.isclose(et, 1.0)       # up = Tp = 0 initially
    wd.step(2)
    u2, v2, w2, T2 = (wd.gather(k) for k in ("u", "v", "w", "T"))
    eu, et = wd.check()
    assert np.isclose(eu, np.sqrt(((u2 - u) ** 2 + (v2 - v) ** 2 + (w2 - w) ** 2).sum()) / np.sqrt((u2 ** 2 + v2 ** 2 + w2 ** 2).sum()), rtol=1e-12)
    assert np.isclose(et, np.abs(T2 - T).sum() / np.abs(T2).sum(), rtol=1e-12)
    wd.close()


# ---------------- whole-array subroutines against the reference's own text (make_golden_thermal3d_fields.py) ----------------
import os as _os  # noqa: E402

FGOLD = np.load(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "ref_fortran_thermal3d_fields.npz"))
BC_SETS = {"cavity": None, "rb": (0, 0, 0, 0, 2, 1)}          # +x,-x,+y,-y,+z,-z: the shipped benchmark cavity; RB plates (bottom hot)


def test_streamingT_whole_array_matches_the_fortran_text():
    wd = orc.ThermalWorld((5, 4, 3), 1)
    R = wd.ranks[0]
    R.g_post[...] = FGOLD["g_post"]
    wd.streamingT()
    assert np.array_equal(R.g, FGOLD["streamingT_g"])
    wd.close()


@pytest.mark.parametrize("case", range(13))
def test_bounceback_and_bouncebackT_whole_array_match_the_fortran_text(case):
    """B3:900-980 and B3:1106-1207 with the benchmark-cavity and the RB-convection macro sets, for the single rank, an interior
    block, three mixed positions and the eight corner blocks of a 3 x 3 x 3 process grid"""
    c = FGOLD["bb_cases"][case]
    coords, dims = tuple(int(x) for x in c[:3]), tuple(int(x) for x in c[3:])
    total = (5 * dims[0], 4 * dims[1], 3 * dims[2])
    for tag, bcT in BC_SETS.items():
        wd = orc.ThermalWorld(total, dims[0] * dims[1] * dims[2], dims=dims, bcT=bcT)
        assert wd.p.paraA == FGOLD[f"paraA_{case}"][0]
        R = next(Q for Q in wd.ranks if Q.coords == coords)
        assert R.n == (5, 4, 3)
        R.f_post[...] = FGOLD["f_post"]; R.f[...] = FGOLD["f0"]; R.g_post[...] = FGOLD["g_post"]; R.g[...] = FGOLD["g0"]
        wd.bounceback(); wd.bouncebackT()
        assert np.array_equal(R.f, FGOLD[f"bounceback_{case}"]), (case, tag)
        assert np.array_equal(R.g, FGOLD[f"bouncebackT_{tag}_{case}"]), (case, tag)
        wd.close()


def test_check_sums_match_the_fortran_text():
    wd = orc.ThermalWorld((5, 4, 3), 1)
    R = wd.ranks[0]
    for k in ("u", "v", "w", "up", "vp", "wp", "T", "Tp"):
        getattr(R, k)[...] = FGOLD[f"check_{k}"]
    e1, e2, e5, e6 = FGOLD["check_sums"]
    eu, et = wd.check()
    assert eu == np.sqrt(e1) / np.sqrt(e2) and et == e5 / e6
    assert np.array_equal(R.Tp, R.T) and np.array_equal(R.wp, R.w)
    wd.close()


# ---------------- the sequential program's own run, from its text (make_golden_thermal3d_seq_run.py) ----------------
SRUN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fortran_thermal3d_seq_run.npz"))


SRUN_RB = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fortran_thermal3d_seq_run_rb.npz"))


@pytest.mark.parametrize("macro_set", ["cavity", "rb"])
@pytest.mark.parametrize("nprocs,dims", [(1, None), (2, None), (4, None), (8, (2, 2, 2)), (6, (3, 1, 2))])
def test_oracle_reproduces_the_sequential_thermal_programs_run(nprocs, dims, macro_set):
    """3d/seq/bouyancy3d.F90 with its shipped macro set, evaluated from its text on 6 x 5 x 4: parameters, initial() and its loop
    (collision, streaming, bounceback, collisionT, streamingT, bouncebackT, macro, macroT) for 1, 2, 10 and 12 iterations, check()
    after 10 and 12.  The restatement of the MPI program reproduces f, g, rho, u, v, w, T and the force fields bit for bit on
    1, 2, 4, 6 and 8 emulated ranks (the reference's seq == MPI contract, from the sequential program's own text)."""
    SRUN = SRUN_RB if macro_set == "rb" else globals()["SRUN"]          # the program's two macro sets (seq:74-85)
    total = tuple(int(x) for x in SRUN["shape"])
    wd = orc.ThermalWorld(total, nprocs, dims=dims, bcT=BC_SETS[macro_set])
    names_p = ("tauf", "viscosity", "diffusivity", "omegaRatating", "paraA", "gBeta1", "gBeta", "Snu", "Sq", "Qd", "Qnu")
    assert tuple(getattr(wd.p, k) for k in names_p) == tuple(SRUN["params"])
    wd.initial()

    def same(tag):
        assert np.array_equal(wd.gather("f"), SRUN[tag + "/f"]), tag
        assert np.array_equal(wd.gather("g"), SRUN[tag + "/g"]), tag
        assert np.array_equal(np.stack([wd.gather(k) for k in ("rho", "u", "v", "w", "T")]), SRUN[tag + "/ruvwT"]), tag

    same("run0")
    done = 0
    for n in (1, 2, 10):
        wd.step(n - done); done = n
        same(f"run{n}")
        assert np.array_equal(np.stack([wd.gather(k) for k in ("Fx", "Fy", "Fz")]), SRUN[f"run{n}/F"]), n
    eu, et = wd.check()
    tol = 0 if nprocs == 1 else 1e-14
    assert abs(eu - SRUN["run10/check"][0]) <= tol * abs(eu) and abs(et - SRUN["run10/check"][1]) <= tol * abs(et)
    wd.step(2)
    eu, et = wd.check()
    assert abs(eu - SRUN["run12/check"][0]) <= tol * abs(eu) and abs(et - SRUN["run12/check"][1]) <= tol * abs(et)
    same("run12")
    wd.close()
