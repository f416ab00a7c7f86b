"""CPU tests of the particle-laden D2Q9 oracle (oracle/particles2d.c): pinned bit for bit to vectors obtained by
machine-evaluating the reference's own Fortran source (tests/golden/make_golden_particles.py), plus construction
tests of the copy-only parts (masked streaming, halo exchanges, mask rebuild) and physical sanity."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fortran_particles.npz"))
dp = C.POINTER(C.c_double)


def ptr(a):
    return a.ctypes.data_as(dp)


def test_parameters_match_reference_source():
    P = dict(zip([str(n) for n in GOLD["params/names"]], GOLD["params/values"]))
    p = orc.p2_default_params()
    for mine, ref in [("Pi", "pi"), ("total_nx", "total_nx"), ("total_ny", "total_ny"), ("radius0", "radius0"), ("rho0", "rho0"),
                      ("rhoSolid", "rhosolid"), ("viscosity", "viscosity"), ("tauf", "tauf"), ("Snu", "snu"), ("Sq", "sq"),
                      ("gravity", "gravity"), ("thresholdWall", "thresholdwall"), ("stiffWall", "stiffwall"),
                      ("thresholdParticle", "thresholdparticle"), ("stiffParticle", "stiffparticle"), ("N", "cnummax")]:
        assert getattr(p, mine) == P[ref], mine


def test_fluid_cell_arithmetic_matches_reference_source_bit_for_bit():
    L, p = orc._p2_lib(), orc.p2_default_params()
    f, s = GOLD["collision/f"], GOLD["collision/ruv"]
    for c in range(f.shape[0]):
        out, mo = np.empty(9), np.empty(3)
        L.p2_collide_cell(ptr(np.ascontiguousarray(f[c])), *[float(x) for x in s[c]], p.Snu, p.Sq, ptr(out))
        assert np.array_equal(out, GOLD["collision/f_post"][c]), c
        L.p2_macro_cell(ptr(np.ascontiguousarray(f[c])), ptr(mo))
        assert np.array_equal(mo, GOLD["macro/ruv"][c]), c


def link_world():
    xc, yc, rad, Uc, Vc, om, rhoAvg = GOLD["link/particle"]
    wd = orc.ParticleWorld([xc], [yc], radius=[rad])
    wd.initial()
    wd.Uc[0], wd.Vc[0], wd.rationalOmega[0] = Uc, Vc, om
    wd._lib.p2_set_rhoAvg(wd._h, rhoAvg)
    return wd


def test_calQ_interpolated_bounceback_and_link_force_match_reference_source():
    wd = link_world()
    R = wd.ranks[0]
    for n, (i, j, a) in enumerate(GOLD["link/ija"]):
        i, j, a = int(i), int(j), int(a)
        q, x0, y0 = C.c_double(), C.c_double(), C.c_double()
        assert wd._lib.p2_calQ(wd._h, 0, float(i), float(j), a, C.byref(x0), C.byref(y0), C.byref(q)) == 0
        assert (q.value, x0.value, y0.value) == tuple(GOLD["link/q_x0_y0"][n]), n
        for s in range(3):       # f_post(:, x - s e_alpha)   (f_post has 2 halo layers: array index = i + 1)
            R.f_post[:, i - s * orc.EX9[a] + 1, j - s * orc.EY9[a] + 1] = GOLD["link/fpost_0_1_2"][n, s]
        R.f[:, i + 2, j + 2] = GOLD["link/f"][n]
        out = np.empty(3)
        assert wd._lib.p2_force_link_r(wd._h, 0, i, j, a, 0, ptr(out)) == 0
        assert np.array_equal(out, GOLD["link/force"][n]), n
        wd._lib.p2_bb_link_r(wd._h, 0, i, j, a, 0)
        assert R.f[orc.OPP9[a], i + 2, j + 2] == GOLD["link/bb"][n], n
    wd.close()


def test_particle_forces_and_kinematics_match_reference_source():
    X, Y, rads = GOLD["forces/xy_rad"]
    wd = orc.ParticleWorld(X, Y, radius=rads)
    wd.initial()
    wd._lib.p2_set_rhoAvg(wd._h, float(GOLD["forces/rhoAvg"][0]))
    wd.wallTotalForceX[:], wd.wallTotalForceY[:], wd.totalTorque[:] = GOLD["forces/hydro"]
    for c in range(len(X)):
        fx, fy = C.c_double(), C.c_double()
        assert wd._lib.p2_particle_forces(wd._h, c, C.byref(fx), C.byref(fy)) == 0
        assert fx.value == GOLD["forces/total"][0, c] and fy.value == GOLD["forces/total"][1, c], c
    p = orc.p2_default_params()
    for v, want in zip(GOLD["advance/in"], GOLD["advance/out"]):
        out = np.empty(5)
        wd._lib.p2_particle_advance(C.byref(p), *[float(x) for x in v], ptr(out))
        assert np.array_equal(out, want)
    wd.close()


def test_refill_matches_reference_source():
    rhoAvg, Uc, Vc, om = GOLD["refill/scal"]
    for n, (i, j) in enumerate(GOLD["refill/ij"]):
        i, j = int(i), int(j)
        xc, yc = GOLD["refill/center"][n]
        wd = orc.ParticleWorld([xc], [yc])
        wd.initial()
        wd.Uc[0], wd.Vc[0], wd.rationalOmega[0] = Uc, Vc, om
        wd._lib.p2_set_rhoAvg(wd._h, rhoAvg)
        R = wd.ranks[0]
        R.f[:, i - 3 + 2:i + 4 + 2, j - 3 + 2:j + 4 + 2] = GOLD["refill/patch"][n]
        assert wd._lib.p2_refill_cell_r(wd._h, 0, i, j, 0) == 0
        assert np.array_equal(R.f[:, i + 2, j + 2], GOLD["refill/f"][n]), n
        assert (R.rho[i - 1, j - 1], R.u[i - 1, j - 1], R.v[i - 1, j - 1]) == tuple(GOLD["refill/ruv"][n]), n
        wd.close()


def test_dims_create_2d_matches_reference_source():
    L = orc._p2_lib()
    for np_, d0, d1 in GOLD["dims/np_d0_d1"]:
        d = (C.c_int * 2)()
        L.p2_dims_create(int(np_), 201, 801, d)
        assert tuple(d) == (d0, d1), np_


SMALL = dict(total_nx=61, total_ny=90)
PX, PY = [20.3, 41.2], [60.0, 33.7]


def small_world(nprocs=1, dims=None):
    wd = orc.ParticleWorld(PX, PY, nprocs=nprocs, dims=dims, **SMALL)
    wd.initial()
    return wd


def test_initial_state():
    wd = small_world(4, (2, 2))
    obst = wd.gather("obst")
    ii, jj = np.meshgrid(np.arange(1, 62), np.arange(1, 91), indexing="ij")
    want = np.zeros_like(obst)
    for x, y in zip(PX, PY):
        want |= ((ii - x) ** 2 + (jj - y) ** 2 <= 100.0)
    assert np.array_equal(obst, want)
    rho = wd.gather("rho")
    assert np.all(rho[want == 1] == 1.01) and np.all(rho[want == 0] == 1.0)
    f = wd.gather("f")
    assert np.all(f[:, want == 1] == 0.0) and np.allclose(f[:, want == 0].sum(axis=0), 1.0, atol=1e-15)
    wd.close()


@pytest.mark.parametrize("nprocs,dims", [(2, (2, 1)), (2, (1, 2)), (4, (2, 2)), (6, (2, 3)), (6, (3, 2))])
def test_halo_exchanges_copy_the_neighbours_interior(nprocs, dims):
    wd = small_world(nprocs, dims)
    rng = np.random.default_rng(2)
    gf = rng.random((9, 61 + 6, 90 + 6))          # global arrays with room for the 3-deep rim
    for R in wd.ranks:
        nx, ny = R.n
        blk = gf[:, R.start[0]:R.start[0] + nx + 6, R.start[1]:R.start[1] + ny + 6]
        R.f[...] = blk + 100.0                                                    # halos wrong on purpose
        R.f[:, 3:nx + 3, 3:ny + 3] = blk[:, 3:nx + 3, 3:ny + 3]
        R.f_post[...] = R.f[:, 1:nx + 5, 1:ny + 5]
    wd.send_all_f(); wd.send_all_fp()
    for R in wd.ranks:
        nx, ny = R.n
        want = gf[:, R.start[0]:R.start[0] + nx + 6, R.start[1]:R.start[1] + ny + 6]
        got, gotp = R.f.copy(), R.f_post.copy()
        # layers that face a physical wall keep their old (+100) values; everything else equals the global array
        lo_x = 3 if R.coords[0] == 0 else 0
        hi_x = nx + 3 if R.coords[0] == wd.dims[0] - 1 else nx + 6
        lo_y = 3 if R.coords[1] == 0 else 0
        hi_y = ny + 3 if R.coords[1] == wd.dims[1] - 1 else ny + 6
        assert np.array_equal(got[:, lo_x:hi_x, lo_y:hi_y], want[:, lo_x:hi_x, lo_y:hi_y])
        plx, phx = max(lo_x, 1), min(hi_x, nx + 5)
        ply, phy = max(lo_y, 1), min(hi_y, ny + 5)
        assert np.array_equal(gotp[:, plx - 1:phx - 1, ply - 1:phy - 1], want[:, plx:phx, ply:phy])
    wd.close()


def test_streaming_skips_populations_coming_out_of_solid_nodes():
    wd = small_world(1)
    R = wd.ranks[0]
    rng = np.random.default_rng(3)
    R.f_post[...] = rng.random(R.f_post.shape)
    before = R.f.copy()
    wd.streaming()
    nx, ny = R.n
    for a in range(9):
        src = R.f_post[a, 2 - orc.EX9[a]:nx + 2 - orc.EX9[a], 2 - orc.EY9[a]:ny + 2 - orc.EY9[a]]
        up_solid = R.obst[1 - orc.EX9[a]:nx + 1 - orc.EX9[a], 1 - orc.EY9[a]:ny + 1 - orc.EY9[a]] == 1
        got = R.f[a, 3:nx + 3, 3:ny + 3]
        assert np.array_equal(got[~up_solid], src[~up_solid])
        assert np.array_equal(got[up_solid], before[a, 3:nx + 3, 3:ny + 3][up_solid])
    wd.close()


def test_particles_sediment_and_fluid_responds():
    wd = small_world(1)
    y0 = wd.yCenter.copy()
    wd.step(150)
    inf = wd.info()
    assert inf["error_flag"] == 0 and inf["itc"] == 150
    assert np.all(wd.yCenter < y0 - 0.5) and np.all(wd.Vc < 0.0)           # heavier than the fluid: they fall
    assert np.abs(wd.gather("v")).max() > 1e-4
    assert abs(inf["rhoAvg"] - 1.0) < 1e-3
    assert np.isclose(wd.check(), 1.0)                                        # up = vp = 0 before the first check
    wd.close()


@pytest.mark.parametrize("nprocs,dims", [(2, (1, 2)), (4, (2, 2)), (6, (2, 3))])
def test_decomposition_invariance_within_summation_order(nprocs, dims):
    """rhoAvg and the force sums are rank-ordered sums (Allreduce), so P ranks agree with one rank only to
    rounding; everything else is a copy.  Positions are chosen so particles cross subdomain boundaries."""
    one, many = small_world(1), small_world(nprocs, dims)
    one.step(120); many.step(120)
    for k in ("rho", "u", "v"):
        a, b = many.gather(k), one.gather(k)
        assert np.abs(a - b).max() < 1e-12, k
    assert np.array_equal(many.gather("obst"), one.gather("obst"))
    for k in ("xCenter", "yCenter", "Uc", "Vc", "rationalOmega"):
        assert np.allclose(getattr(many, k), getattr(one, k), rtol=0, atol=1e-12), k
    one.close(); many.close()


# ---- the product's calQ (mglc_b200/csrc/p2d_calq.inl, compiled for the host) ------------------------------------------

@pytest.fixture(scope="module")
def calq_shim(tmp_path_factory):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path_factory.mktemp("calq") / "calq_host.so")
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", out,
                           os.path.join(root, "tests", "host_shim", "calq_host.cpp")])
    S = C.CDLL(out)
    S.shim_calq.argtypes = [C.c_double] * 7 + [dp]
    return S


def test_product_calQ_matches_reference_source_vectors(calq_shim):
    """The square-root-free far-field halvings must not change a single bit of q, x0, y0 (vectors evaluated from the
    reference's Fortran text, P4/particle_bounceback.F90:98-141)."""
    xc, yc, rad = GOLD["link/particle"][:3]
    for n, (i, j, a) in enumerate(GOLD["link/ija"]):
        out = np.empty(3)
        assert calq_shim.shim_calq(xc, yc, rad, float(i), float(j), float(orc.EX9[int(a)]), float(orc.EY9[int(a)]), ptr(out)) == 0
        q, x0, y0 = GOLD["link/q_x0_y0"][n]
        assert (out[2], out[0], out[1]) == (q, x0, y0), n


def test_product_calQ_matches_oracle_on_random_and_grazing_links(calq_shim):
    rng = np.random.default_rng(2024)
    checked = grazing = 0
    for trial in range(400):
        rad = float(rng.uniform(2.5, 14.0))
        xc, yc = (float(v) for v in rng.uniform(30.0, 34.0, 2))
        wd = orc.ParticleWorld([xc], [yc], radius=[rad], total_nx=64, total_ny=64)
        wd.initial()
        lo, hi = int(np.floor(xc - rad)) - 1, int(np.ceil(xc + rad)) + 1
        for i in range(lo, hi + 1):
            for j in range(int(np.floor(yc - rad)) - 1, int(np.ceil(yc + rad)) + 2):
                if (i - xc) ** 2 + (j - yc) ** 2 <= rad * rad:
                    continue
                for a in range(1, 9):
                    ip, jp = i + int(orc.EX9[a]), j + int(orc.EY9[a])
                    if (ip - xc) ** 2 + (jp - yc) ** 2 > rad * rad:
                        continue
                    q, x0, y0 = C.c_double(), C.c_double(), C.c_double()
                    rc = wd._lib.p2_calQ(wd._h, 0, float(i), float(j), a, C.byref(x0), C.byref(y0), C.byref(q))
                    out = np.empty(3)
                    rc2 = calq_shim.shim_calq(xc, yc, rad, float(i), float(j), float(orc.EX9[a]), float(orc.EY9[a]), ptr(out))
                    assert (rc == 0) == (rc2 == 0)
                    if rc == 0:
                        assert (out[0], out[1], out[2]) == (x0.value, y0.value, q.value), (trial, i, j, a)
                    checked += 1
                    grazing += q.value < 0.02 or q.value > 0.98
        wd.close()
        if checked > 40000:
            break
    assert checked > 20000 and grazing > 200


# ---------------- the program's own run on one rank, from its text (make_golden_particles_run.py) ----------------
PRUN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fortran_particles_run.npz"))


def test_oracle_reproduces_the_particle_programs_run():
    """case4/mpi_particle evaluated from its text as a whole program on one rank (56 x 72, two particles 3.5 cells apart):
    initial() and 1, 2, 30 iterations of collision, streaming, bounceback, bounceback_particle (with calQ called as written),
    macro, calForce, updateCenter (mask rebuild, refill), then check().  The restatement reproduces every population incl. the
    ghost layers, rho, u, v, the solid mask, rhoAvg and the particle state bit for bit."""
    nx, ny, N = (int(x) for x in PRUN["shape"])
    x, y = PRUN["positions"]
    wd = orc.ParticleWorld(x, y, nprocs=1, total_nx=nx, total_ny=ny)
    names_p = ("Pi", "radius0", "rho0", "rhoSolid", "viscosity", "tauf", "Snu", "Sq", "gravity", "thresholdWall", "stiffWall",
               "thresholdParticle", "stiffParticle")
    assert tuple(getattr(wd.p, k) for k in names_p) == tuple(PRUN["params"])
    wd.initial()
    R = wd.ranks[0]
    keys = ("xCenter", "yCenter", "Uc", "Vc", "rationalOmega", "wallTotalForceX", "wallTotalForceY", "totalTorque", "xCenterOld", "yCenterOld")

    def same(tag, forces=True):
        assert np.array_equal(R.f, PRUN[tag + "/f"]), tag
        assert np.array_equal(R.f_post, PRUN[tag + "/f_post"]), tag
        assert np.array_equal(np.stack([R.rho, R.u, R.v]), PRUN[tag + "/ruv"]), tag
        assert np.array_equal(R.obst, PRUN[tag + "/obst"]), tag
        got = np.stack([getattr(wd, k) for k in keys])
        assert np.array_equal(got, PRUN[tag + "/particles"]), (tag, got - PRUN[tag + "/particles"])

    same("run0")
    done = 0
    for n in (1, 2, 30):
        wd.step(n - done); done = n
        same(f"run{n}")
        assert wd.info()["rhoAvg"] == PRUN[f"run{n}/rhoAvg"][0]
    assert PRUN["run30/particles"][5].any() and PRUN["run30/obst"].sum() > 500      # forces act; the particles are on the lattice
    assert (PRUN["run30/obst"] != PRUN["run0/obst"]).any()                            # nodes were uncovered / covered: refill ran
    assert wd.check() == PRUN["run30/check"][2]
    wd.close()


# ---- the options taken from the reference's other particle scenario (case1/mpi_complete, "P1") ---------------------------
GOLD1 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fortran_particles_case1.npz"))


def test_linear_interpolated_bounceback_matches_case1_source():
    """bb_linear = 1: P1/particle_bounceback.F90:67-75 (`#ifdef linear`), with calQ as that file writes it, bit for bit"""
    xc, yc, rad, Uc, Vc, om, rhoAvg = GOLD1["link/particle"]
    wd = orc.ParticleWorld([xc], [yc], radius=[rad], total_nx=140, total_ny=90, bb_linear=1)
    wd.initial()
    wd.Uc[0], wd.Vc[0], wd.rationalOmega[0] = Uc, Vc, om
    wd._lib.p2_set_rhoAvg(wd._h, float(rhoAvg))
    R = wd.ranks[0]
    for n, (i, j, a) in enumerate(GOLD1["link/ija"]):
        i, j, a = int(i), int(j), int(a)
        q, x0, y0 = C.c_double(), C.c_double(), C.c_double()
        assert wd._lib.p2_calQ(wd._h, 0, float(i), float(j), a, C.byref(x0), C.byref(y0), C.byref(q)) == 0
        assert (q.value, x0.value, y0.value) == tuple(GOLD1["link/q_x0_y0"][n]), n
        for s in range(2):
            R.f_post[:, i - s * orc.EX9[a] + 1, j - s * orc.EY9[a] + 1] = GOLD1["link/fpost_0_1"][n, s]
        wd._lib.p2_bb_link_r(wd._h, 0, i, j, a, 0)
        assert R.f[orc.OPP9[a], i + 2, j + 2] == GOLD1["link/bb_linear"][n], n
    wd.close()


@pytest.mark.parametrize("frame", ["moving", "stationary"])
def test_moving_walls_match_case1_source(frame):
    """moving_walls = 1: P1/fluid.F90:130-145 (`movingFrame`, walls seen from a frame moving with U0) and :151-168
    (`stationaryFrame` = the same expressions with U0 = 0), bit for bit on a seeded row of nodes along either wall"""
    nx, ny, Uwall, U0 = GOLD1["walls/params"]
    nx, ny = int(nx), int(ny)
    wd = orc.ParticleWorld([], [], total_nx=nx, total_ny=ny, moving_walls=1, Uwall=float(Uwall), Uframe=float(U0) if frame == "moving" else 0.0)
    wd.initial()
    R = wd.ranks[0]
    for row, j in enumerate((1, ny)):
        R.f_post[:, 2:nx + 2, j + 1] = GOLD1[f"walls/{frame}/f_post"][row].T
    wd.bounceback()
    want = GOLD1[f"walls/{frame}/f"]
    for row, j in enumerate((1, ny)):
        got = R.f[:, 3:nx + 3, j + 2].T
        pops = (2, 5, 6) if j == 1 else (4, 7, 8)
        inner = slice(1, nx - 1)                 # the reference scenario has no side walls; ours overwrite 5/8 and 6/7 in the two corner nodes
        for b in pops:
            assert np.array_equal(got[inner, b], want[row][inner, b]), (frame, j, b)
    wd.close()
