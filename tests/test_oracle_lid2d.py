"""Pins the 2-D lid-driven-cavity oracle (oracle/lid2d.c) to the reference:
  * variant "c" against the reference's own compiled C program -- live through oracle/_ref/liblid2d_ref.so (built by
    `make -C oracle ref` from /root/reference/MPI/Lid_driven_cavity/c/lid_driven_cavity.c, shipped to the GPU box) and
    through the committed outputs of that program (tests/golden/ref_lid2d.npz, made by make_golden_lid2d.py);
  * variant "f" against the Fortran source text of 2d_revised/mpi_blocked evaluated by tests/golden/fortran_eval.py,
    and the reference's seq == MPI contract (P emulated ranks == 1 rank, bit for bit)."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "ref_lid2d.npz"))
REF_SO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "liblid2d_ref.so")


def c_view(a):
    """oracle (nx, ny[, 9 leading]) column-major -> the C program's [NX][NY]([9])"""
    return np.ascontiguousarray(np.transpose(a, (1, 2, 0))) if a.ndim == 3 else np.ascontiguousarray(a)


def test_parameters_match_the_compiled_reference_and_the_fortran_text():
    wd = orc.Lid2DWorld((200, 200), variant="c")
    assert (wd.tauf, wd.Snu, wd.Sq) == tuple(GOLD["c/params"])
    wd.close()
    wd = orc.Lid2DWorld((201, 201), variant="f")
    assert (wd.tauf, wd.Snu, wd.Sq) == tuple(GOLD["f/params"])
    wd.close()


@pytest.mark.parametrize("variant", ["c", "f"])
def test_collision_cells_bit_exact(variant):
    f, ruv = GOLD["cells/f"], GOLD["cells/ruv"]
    _, snu, sq = GOLD[variant + "/params"]
    for k in range(len(f)):
        got = orc.l2_collide_cell(variant, f[k], *ruv[k], snu, sq)
        assert np.array_equal(got, GOLD[variant + "/collision_f_post"][k]), (variant, k)
    # the two programs really round differently (otherwise one variant would do)
    assert not np.array_equal(GOLD["c/collision_f_post"], GOLD["f/collision_f_post"])


def test_fortran_macro_and_feq_cells_bit_exact():
    f, ruv = GOLD["cells/f"], GOLD["cells/ruv"]
    wd = orc.Lid2DWorld((len(f), 1), variant="f")
    R = wd.ranks[0]
    R.f[:, :, 0] = f.T
    wd.macro()
    got = np.stack([R.rho[:, 0], R.u[:, 0], R.v[:, 0]], axis=1)
    assert np.array_equal(got, GOLD["f/macro_ruv"])
    wd.close()
    # initial(): f = feq(rho0, u) -- the lid row carries u = U0; compare the formula on the golden cells through a world
    # whose rho/u/v are overwritten before the population loop is re-run by hand
    W9 = [4.0 / 9.0] + [1.0 / 9.0] * 4 + [1.0 / 36.0] * 4
    EX, EY = orc.EX9, orc.EY9
    for k in range(len(f)):
        rho, u, v = ruv[k]
        us2 = u * u + v * v
        for a in range(9):
            un = u * float(EX[a]) + v * float(EY[a])
            assert rho * W9[a] * (1.0 + 3.0 * un + 4.5 * un * un - 1.5 * us2) == GOLD["f/feq"][k, a]


def test_variant_c_run_matches_committed_reference_outputs():
    wd = orc.Lid2DWorld((200, 200), variant="c")
    wd.initial()
    f0 = c_view(wd.gather("f"))
    assert np.array_equal(f0[:, -1, :], GOLD["c/initial_f_top"]) and np.array_equal(f0[7, 3, :], GOLD["c/initial_f_bulk"])
    done = 0
    for n in (1, 10, 100, 1000):
        wd.step(n - done); done = n
        for k in ("rho", "u", "v"):
            a = c_view(wd.gather(k))
            assert np.array_equal(a[100, :], GOLD[f"c/run{n}/{k}_col100"]), (n, k)
            assert np.array_equal(a[:, 199], GOLD[f"c/run{n}/{k}_row199"]), (n, k)
            assert np.array_equal(a[:, 0], GOLD[f"c/run{n}/{k}_row0"]), (n, k)
            assert np.array_equal(np.array([a.sum(), np.abs(a).sum(), (a * a).sum()]), GOLD[f"c/run{n}/{k}_sum"]), (n, k)
        f = c_view(wd.gather("f"))
        assert np.array_equal(f[:3, :3, :], GOLD[f"c/run{n}/f_corner"]) and np.array_equal(f[-3:, -3:, :], GOLD[f"c/run{n}/f_topright"])
    assert wd.check() == GOLD["c/check_1000"][0]
    wd.step(1000)
    assert wd.check() == GOLD["c/check_2000"][0]
    for k in ("rho", "u", "v"):
        assert np.array_equal(c_view(wd.gather(k)), GOLD[f"c/run2000/{k}_full"]), k
    wd.close()


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/liblid2d_ref.so not built (make -C oracle ref)")
def test_variant_c_live_against_the_compiled_reference(tmp_path, monkeypatch):
    """every array of the reference program after each of its own subroutines, on seeded non-trivial input"""
    monkeypatch.chdir(tmp_path)
    ref = orc.RefLid2D(REF_SO)
    ref.lib.initial()
    wd = orc.Lid2DWorld((200, 200), variant="c")
    wd.initial()
    assert np.array_equal(ref.f_F(), wd.gather("f"))
    rng = np.random.default_rng(5)
    f = np.asfortranarray(wd.gather("f") * (1.0 + 0.05 * rng.uniform(-1, 1, (9, 200, 200))))
    rho = np.asfortranarray(1.0 + 0.02 * rng.uniform(-1, 1, (200, 200)))
    u, v = (np.asfortranarray(0.05 * rng.uniform(-1, 1, (200, 200))) for _ in range(2))
    ref.f[...] = np.transpose(f, (1, 2, 0)); ref.rho[...] = rho; ref.u[...] = u; ref.v[...] = v
    for k, a in (("f", f), ("rho", rho), ("u", u), ("v", v)):
        wd.scatter(k, a)
    for it in range(6):
        ref.lib.collision(); wd.collision()
        assert np.array_equal(ref.f_F(True), wd.ranks[0].f_post[:, 1:-1, 1:-1]), ("collision", it)
        ref.lib.streaming(); ref.lib.boundary(); wd.message_passing_sendrecv(); wd.streaming(); wd.bounceback()
        assert np.array_equal(ref.f_F(), wd.gather("f")), ("streaming+boundary", it)
        ref.lib.macro(); wd.macro()
        for k in ("rho", "u", "v"):
            assert np.array_equal(ref.field_F(k), wd.gather(k)), (k, it)
    assert ref.lib.check(6) == wd.check()
    ref.step(20); wd.step(20)
    assert ref.lib.check(26) == wd.check()
    assert np.array_equal(ref.field_F("up"), wd.gather("up"))


@pytest.mark.parametrize("variant", ["c", "f"])
@pytest.mark.parametrize("nprocs,dims", [(2, None), (4, None), (6, None), (3, (1, 3)), (4, (4, 1)), (9, None)])
def test_decomposed_equals_single_rank_bit_for_bit(variant, nprocs, dims):
    total = (23, 19)
    one, many = orc.Lid2DWorld(total, 1, variant=variant), orc.Lid2DWorld(total, nprocs, dims, variant=variant)
    one.initial(); many.initial()
    one.step(30); many.step(30)
    for k in ("f", "rho", "u", "v"):
        assert np.array_equal(one.gather(k), many.gather(k)), k
    if variant == "f":
        assert np.isclose(one.check(), many.check(), rtol=1e-13)
    one.close(); many.close()


def test_dims_create_2d():
    lib = orc._l2_lib()
    import ctypes as C
    for n, want in [(1, (1, 1)), (2, (2, 1)), (4, (2, 2)), (6, (3, 2)), (8, (4, 2)), (9, (3, 3)), (12, (4, 3)), (7, (7, 1))]:
        d = (C.c_int * 2)()
        lib.l2_dims_create(n, d)
        assert tuple(d) == want


def test_mass_is_conserved_and_walls_never_leak_halo_values():
    wd = orc.Lid2DWorld((31, 17), 4, variant="f")
    wd.initial()
    for R in wd.ranks:
        R.f_post[...] = np.nan            # poison: wall halos must never reach f
    m0 = wd.gather("rho").sum()
    wd.step(50)
    rho = wd.gather("rho")
    assert np.isfinite(rho).all() and abs(rho.sum() - m0) / m0 < 1e-13
    wd.close()


# ---------------- the Fortran program's whole-array subroutines against its own text (make_golden_lid2d_fields.py) ----------------
FGOLD = np.load(os.path.join(HERE, "golden", "ref_fortran_lid2d_fields.npz"))


def test_fortran_streaming_whole_array():
    wd = orc.Lid2DWorld((6, 5), 1, variant="f")
    R = wd.ranks[0]
    R.f_post[...] = FGOLD["f_post"]
    wd.streaming()
    assert np.array_equal(R.f, FGOLD["streaming_f"])
    wd.close()


@pytest.mark.parametrize("case", range(7))
def test_fortran_bounceback_whole_array(case):
    """bounceback.f90:7-42 for the single rank, the interior block, the four corner blocks and the top-middle block of a 3 x 3
    grid: which walls a block owns, the moving top wall winning in the corners, the lid term with the previous macro()'s rho"""
    c = FGOLD["bb_cases"][case]
    coords, dims = tuple(int(x) for x in c[:2]), tuple(int(x) for x in c[2:])
    wd = orc.Lid2DWorld((6 * dims[0], 5 * dims[1]), dims[0] * dims[1], dims, variant="f")
    R = next(Q for Q in wd.ranks if Q.coords == coords)
    assert R.n == (6, 5)
    R.f_post[...] = FGOLD["f_post"]; R.f[...] = FGOLD["f0"]; R.rho[...] = FGOLD["rho"]
    wd.bounceback()
    assert np.array_equal(R.f, FGOLD[f"bounceback_{case}"])
    wd.close()


def test_fortran_check_sums():
    wd = orc.Lid2DWorld((6, 5), 1, variant="f")
    R = wd.ranks[0]
    for k in ("u", "v", "up", "vp"):
        getattr(R, k)[...] = FGOLD[f"check_{k}"]
    e1, e2 = FGOLD["check_sums"]
    assert wd.check() == np.sqrt(e1) / np.sqrt(e2)
    wd.close()


# ---------------- the incompressible sequential program (variant "i") against its own text (make_golden_lid2d_incomp.py) ----------------
IGOLD = np.load(os.path.join(HERE, "golden", "ref_fortran_lid2d_incomp.npz"))


def _incomp_world(nprocs=1, dims=None):
    nx, ny = (int(x) for x in IGOLD["shape"])
    wd = orc.Lid2DWorld((nx, ny), nprocs, dims, variant="i")
    assert (wd.tauf, wd.Snu, wd.Sq) == tuple(IGOLD["params"])
    return wd


def test_incompressible_subroutines_whole_array():
    """lid-driven_cavity_incompress.f90: collision :181-236, streaming :249-259, bounceback :270-293, macro :304-310,
    check :322-335 on seeded arrays, bit for bit"""
    wd = _incomp_world()
    R = wd.ranks[0]

    def load():
        R.f[...] = IGOLD["in/f0"]; R.f_post[...] = IGOLD["in/f_post"]
        for k in ("rho", "u", "v", "up", "vp"):
            getattr(R, k)[...] = IGOLD["in/" + k]

    load(); wd.collision()
    assert np.array_equal(R.f_post[:, 1:-1, 1:-1], IGOLD["collision/f_post"])
    load(); wd.streaming()
    assert np.array_equal(R.f, IGOLD["streaming/f"])
    load(); wd.bounceback()
    assert np.array_equal(R.f, IGOLD["bounceback/f"])
    load(); wd.macro()
    assert np.array_equal(np.stack([R.rho, R.u, R.v]), IGOLD["macro/ruv"])
    load()
    e1, e2, eu = IGOLD["check/e1_e2_errorU"]
    assert wd.check() == eu == e1 / e2
    assert np.array_equal(R.up, IGOLD["in/u"]) and np.array_equal(R.vp, IGOLD["in/v"])
    wd.close()


def test_incompressible_differs_from_the_compressible_program():
    """the variant is not a relabelling: same inputs, different f_post / u / lid populations"""
    a, b = _incomp_world(), orc.Lid2DWorld(tuple(int(x) for x in IGOLD["shape"]), 1, variant="f")
    for wd in (a, b):
        R = wd.ranks[0]
        R.f[...] = IGOLD["in/f0"]; R.f_post[...] = IGOLD["in/f_post"]
        for k in ("rho", "u", "v"):
            getattr(R, k)[...] = IGOLD["in/" + k]
        wd.collision()
    assert not np.array_equal(a.ranks[0].f_post, b.ranks[0].f_post)
    for wd in (a, b):
        wd.ranks[0].f_post[...] = IGOLD["in/f_post"]
        wd.bounceback(); wd.macro()
    assert not np.array_equal(a.ranks[0].f[7, :, -1], b.ranks[0].f[7, :, -1])
    assert np.array_equal(a.ranks[0].rho[:, :-1], b.ranks[0].rho[:, :-1]) and not np.array_equal(a.ranks[0].u[:, :-1], b.ranks[0].u[:, :-1])
    a.close(); b.close()


@pytest.mark.parametrize("nprocs,dims", [(1, None), (4, (2, 2)), (6, (3, 2))])
def test_incompressible_program_run(nprocs, dims):
    """initial() (rho = 0 until the first macro(): the first collision() relaxes towards meq(1) = 3|u|^2) and 1, 2, 20, 25
    iterations of the program's loop with check() after 20 and 25; on P emulated ranks the halo exchange of the MPI program
    makes the blocks reproduce the sequential program bit for bit"""
    wd = _incomp_world(nprocs, dims)
    wd.initial()
    assert np.array_equal(wd.gather("f"), IGOLD["run0/f"])
    assert np.array_equal(np.stack([wd.gather(k) for k in ("rho", "u", "v")]), IGOLD["run0/ruv"])
    assert not wd.gather("rho").any()
    done = 0
    for n in (1, 2, 20):
        wd.step(n - done); done = n
        assert np.array_equal(wd.gather("f"), IGOLD[f"run{n}/f"]), n
        assert np.array_equal(np.stack([wd.gather(k) for k in ("rho", "u", "v")]), IGOLD[f"run{n}/ruv"]), n
    e = wd.check()
    if nprocs == 1:
        assert e == IGOLD["run20/check"][2]
    else:
        assert abs(e - IGOLD["run20/check"][2]) <= 1e-14 * abs(e)        # rank sums add in a different order
    wd.step(5)
    e = wd.check()
    assert abs(e - IGOLD["run25/check"][2]) <= (0 if nprocs == 1 else 1e-14 * abs(e))
    wd.close()


# ---------------- the C program with its model switch set to SRT / BGK (variant "s"; make_golden_lid2d_srt.py) ----------------
SGOLD = np.load(os.path.join(HERE, "golden", "ref_lid2d_srt.npz"))
SRT_SO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "liblid2d_srt_ref.so")


def test_srt_collision_cells_bit_exact():
    """c:160-176 on the seeded cells, as the compiled program (model = SRT) computed them"""
    f, ruv = GOLD["cells/f"], GOLD["cells/ruv"]
    _, snu, sq = SGOLD["params"]
    assert tuple(SGOLD["params"]) == tuple(GOLD["c/params"])
    for k in range(len(f)):
        assert np.array_equal(orc.l2_collide_cell("s", f[k], *ruv[k], snu, sq), SGOLD["collision_f_post"][k]), k
    assert not np.array_equal(SGOLD["collision_f_post"], GOLD["c/collision_f_post"])


def test_srt_run_matches_committed_reference_outputs():
    wd = orc.Lid2DWorld((200, 200), variant="s")
    wd.initial()
    done = 0
    for n in (1, 10, 100, 1000):
        wd.step(n - done); done = n
        for k in ("rho", "u", "v"):
            a = c_view(wd.gather(k))
            assert np.array_equal(a[100, :], SGOLD[f"run{n}/{k}_col100"]), (n, k)
            assert np.array_equal(a[:, 199], SGOLD[f"run{n}/{k}_row199"]), (n, k)
            assert np.array_equal(a[:, 0], SGOLD[f"run{n}/{k}_row0"]), (n, k)
            assert np.array_equal(a[::4, ::4], SGOLD[f"run{n}/{k}_stride4"]), (n, k)
            assert np.array_equal(np.array([a.sum(), np.abs(a).sum(), (a * a).sum()]), SGOLD[f"run{n}/{k}_sum"]), (n, k)
        f = c_view(wd.gather("f"))
        assert np.array_equal(f[:3, :3, :], SGOLD[f"run{n}/f_corner"]) and np.array_equal(f[-3:, -3:, :], SGOLD[f"run{n}/f_topright"])
    assert wd.check() == SGOLD["check_1000"][0]
    wd.step(100)
    assert wd.check() == SGOLD["check_1100"][0]
    wd.close()


@pytest.mark.skipif(not os.path.exists(SRT_SO), reason="oracle/_ref/liblid2d_srt_ref.so not built (make -C oracle ref)")
def test_srt_live_against_the_compiled_reference(tmp_path, monkeypatch):
    """the reference program compiled with model = SRT: every array after each of its own subroutines, seeded input"""
    monkeypatch.chdir(tmp_path)
    ref = orc.RefLid2D(SRT_SO)
    ref.lib.initial()
    wd = orc.Lid2DWorld((200, 200), variant="s")
    wd.initial()
    assert np.array_equal(ref.f_F(), wd.gather("f"))
    rng = np.random.default_rng(6)
    f = np.asfortranarray(wd.gather("f") * (1.0 + 0.05 * rng.uniform(-1, 1, (9, 200, 200))))
    rho = np.asfortranarray(1.0 + 0.02 * rng.uniform(-1, 1, (200, 200)))
    u, v = (np.asfortranarray(0.05 * rng.uniform(-1, 1, (200, 200))) for _ in range(2))
    ref.f[...] = np.transpose(f, (1, 2, 0)); ref.rho[...] = rho; ref.u[...] = u; ref.v[...] = v
    for k, a in (("f", f), ("rho", rho), ("u", u), ("v", v)):
        wd.scatter(k, a)
    for it in range(6):
        ref.lib.collision(); wd.collision()
        assert np.array_equal(ref.f_F(True), wd.ranks[0].f_post[:, 1:-1, 1:-1]), ("collision", it)
        ref.lib.streaming(); ref.lib.boundary(); wd.message_passing_sendrecv(); wd.streaming(); wd.bounceback()
        assert np.array_equal(ref.f_F(), wd.gather("f")), ("streaming+boundary", it)
        ref.lib.macro(); wd.macro()
        for k in ("rho", "u", "v"):
            assert np.array_equal(ref.field_F(k), wd.gather(k)), (k, it)
    assert ref.lib.check(6) == wd.check()
    wd.close()


def test_srt_decomposed_equals_one_rank():
    one, many = orc.Lid2DWorld((31, 23), 1, variant="s"), orc.Lid2DWorld((31, 23), 6, variant="s")
    for wd in (one, many):
        wd.initial(); wd.step(40)
    for k in ("f", "rho", "u", "v"):
        assert np.array_equal(one.gather(k), many.gather(k)), k
    one.close(); many.close()


# ---------------- the compressible sequential program against variant "f" (make_golden_lid2d_incomp.py, second output) ----------------
QGOLD = np.load(os.path.join(HERE, "golden", "ref_fortran_lid2d_seq.npz"))


@pytest.mark.parametrize("nprocs,dims", [(1, None), (4, (2, 2)), (6, (2, 3))])
def test_variant_f_reproduces_the_sequential_fortran_program(nprocs, dims):
    """seq/lid-driven_cavity.f90 (= the subroutines of 2d_old/lid-driven_cavity.f90) evaluated from its text: every subroutine
    on seeded arrays and the program's loop for 25 iterations.  The MPI program's restatement (variant "f"), on 1, 4 and 6
    emulated ranks, reproduces the sequential program bit for bit: the reference's seq == MPI contract, from its own text."""
    nx, ny = (int(x) for x in QGOLD["shape"])
    wd = orc.Lid2DWorld((nx, ny), nprocs, dims, variant="f")
    assert (wd.tauf, wd.Snu, wd.Sq) == tuple(QGOLD["params"])
    if nprocs == 1:
        R = wd.ranks[0]

        def load():
            R.f[...] = QGOLD["in/f0"]; R.f_post[...] = QGOLD["in/f_post"]
            for k in ("rho", "u", "v", "up", "vp"):
                getattr(R, k)[...] = QGOLD["in/" + k]

        load(); wd.collision()
        assert np.array_equal(R.f_post[:, 1:-1, 1:-1], QGOLD["collision/f_post"])
        load(); wd.streaming()
        assert np.array_equal(R.f, QGOLD["streaming/f"])
        load(); wd.bounceback()
        assert np.array_equal(R.f, QGOLD["bounceback/f"])
        load(); wd.macro()
        assert np.array_equal(np.stack([R.rho, R.u, R.v]), QGOLD["macro/ruv"])
        load()
        e1, e2, eu = QGOLD["check/e1_e2_errorU"]
        assert wd.check() == eu == np.sqrt(e1) / np.sqrt(e2)
    wd.initial()
    assert np.array_equal(wd.gather("f"), QGOLD["run0/f"])
    assert np.array_equal(np.stack([wd.gather(k) for k in ("rho", "u", "v")]), QGOLD["run0/ruv"])
    done = 0
    for n in (1, 2, 20):
        wd.step(n - done); done = n
        assert np.array_equal(wd.gather("f"), QGOLD[f"run{n}/f"]), n
        assert np.array_equal(np.stack([wd.gather(k) for k in ("rho", "u", "v")]), QGOLD[f"run{n}/ruv"]), n
    e = wd.check()
    assert abs(e - QGOLD["run20/check"][2]) <= (0 if nprocs == 1 else 1e-14 * abs(e))
    wd.close()


def _flow_binary_bytes(wd):
    """the bytes of the C program's output_binary() (c:428-457: x, y, rho, u, v as double[NX][NY]) from an oracle world"""
    nx, ny = wd.total
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    x, y = i * (float(nx) / (nx - 1)), j * (float(nx) / (ny - 1))
    return b"".join(np.ascontiguousarray(a, dtype=np.float64).tobytes() for a in (x, y, c_view(wd.gather("rho")), c_view(wd.gather("u")), c_view(wd.gather("v"))))


@pytest.mark.parametrize("variant,gold", [("c", "c/"), ("s", "")])
def test_output_file_of_the_c_program_after_2000_iterations(variant, gold):
    """the SHA-256 of the file the compiled reference program wrote (MRT and SRT builds) == the same bytes assembled from the
    restatement's fields: what examples/lid2d_driver.c must reproduce on the GPU with model = 2 / model = 1"""
    import hashlib
    G = GOLD if variant == "c" else SGOLD
    wd = orc.Lid2DWorld((200, 200), variant=variant)
    wd.initial(); wd.step(2000)
    raw = _flow_binary_bytes(wd)
    assert len(raw) == int(G[gold + "output_binary_len"][0])
    assert hashlib.sha256(raw).digest() == G[gold + "output_binary_sha256"].tobytes()
    wd.close()
