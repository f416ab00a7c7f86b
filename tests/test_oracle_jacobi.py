"""CPU tests of the Jacobi oracle (oracle/jacobi.c).  The 2-D mode is pinned bit for bit to the
REFERENCE's own compiled code: golden vectors produced by MPI/Laplace/c/laplace2d.c's jacobi()+swap()
(tests/golden/make_golden_jacobi.py) and, when oracle/_ref/liblaplace2d_ref.so is present, a live
comparison on fresh random grids."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "laplace2d_ref.npz"))
REF_SO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "liblaplace2d_ref.so")


def oracle_from_c_grid(A0):
    """laplace2d.c grid [x][y] (boundary included) -> 1-rank oracle world whose ghost layer is that
    boundary.  Both index the same mathematical (i, j): the reference sums (i-1,j)+(i+1,j)+(i,j-1)+(i,j+1)
    in C (laplace2d.c:77-78) and in Fortran (jacobi2d_mpi.f90:178)."""
    nx, ny = A0.shape[0] - 2, A0.shape[1] - 2
    wd = orc.JacobiWorld((nx, ny), 1)
    wd.array(0, "A")[...] = A0
    wd.array(0, "A_new")[...] = A0
    wd.array(0, "A_p")[...] = A0
    return wd


@pytest.mark.parametrize("name,its", [("shipped_bc_19x14", 25), ("random_23x37", 7), ("random_130x9", 3)])
def test_2d_matches_reference_golden_vectors(name, its):
    A0, want, errs = GOLD[name + "/A0"], GOLD[name + "/A"], GOLD[name + "/err"]
    wd = oracle_from_c_grid(A0)
    for it in range(its):
        wd.step(1)
        assert wd.check_diff() == errs[it]          # jacobi() returns max |A_new - A| (laplace2d.c:79-80)
    assert np.array_equal(wd.array(0, "A")[1:-1, 1:-1], want[1:-1, 1:-1])
    wd.close()


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built (needs /root/reference)")
def test_2d_matches_live_reference_library():
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden_jacobi as mk
    L = mk.ref_lib()
    rng = np.random.default_rng(77)
    A0 = rng.normal(size=(61, 45))
    want, _ = mk.run_reference(L, A0, 11)
    wd = oracle_from_c_grid(A0)
    wd.step(11)
    assert np.array_equal(wd.array(0, "A")[1:-1, 1:-1], want[1:-1, 1:-1])
    wd.close()


def test_init_boundary_condition():
    wd = orc.JacobiWorld((10, 8), 4)            # 2 x 2
    wd.init()
    for r, inf in enumerate(wd.info):
        A = wd.array(r)
        top = inf["coords"][1] == wd.dims[1] - 1
        assert np.all(A[:, :-1] == 0.0)
        assert np.all(A[:, -1] == (1.0 if top else 0.0))
    wd.close()
    wd = orc.JacobiWorld((6, 5, 4), 2, dims=(1, 1, 2))
    wd.init()
    assert np.all(wd.array(1)[:, :, -1] == 1.0) and np.all(wd.array(0)[:, :, -1] == 0.0)
    wd.close()


@pytest.mark.parametrize("total", [(12, 9), (7, 6, 5)])
def test_linear_profile_is_a_fixed_point(total):
    wd = orc.JacobiWorld(total, 1)
    idx = np.meshgrid(*[np.arange(n + 2) for n in total], indexing="ij")
    lin = sum((q + 1.0) * g for q, g in enumerate(idx)) + 3.0     # small integers: every sum is exact
    wd.array(0, "A")[...] = lin
    wd.array(0, "A_new")[...] = lin
    wd.step(6)
    inner = tuple(slice(1, -1) for _ in total)
    if len(total) == 2:
        assert np.array_equal(wd.array(0)[inner], lin[inner])
    else:                                                          # 1/6 is not a power of two
        assert np.allclose(wd.array(0)[inner], lin[inner], rtol=1e-15, atol=0)
    wd.close()


@pytest.mark.parametrize("total,nprocs,dims", [((37, 23), 2, None), ((37, 23), 4, None), ((37, 23), 6, None),
                                               ((37, 23), 3, (1, 3)), ((17, 13, 11), 2, None),
                                               ((17, 13, 11), 8, None), ((17, 13, 11), 12, None),
                                               ((17, 13, 11), 3, (1, 1, 3))])
def test_decomposition_invariance_bit_exact(total, nprocs, dims):
    rng = np.random.default_rng(5)
    glob = rng.random(tuple(n + 2 for n in total))
    src = rng.random(tuple(n + 2 for n in total))

    def load(wd):
        for r, inf in enumerate(wd.info):
            sl = tuple(slice(s, s + n + 2) for s, n in zip(inf["start"], inf["n"]))
            wd.array(r, "A")[...] = glob[sl]
            wd.array(r, "A_new")[...] = glob[sl]
            wd.array(r, "f")[...] = src[sl]

    one, many = orc.JacobiWorld(total, 1), orc.JacobiWorld(total, nprocs, dims)
    load(one); load(many)
    one.step(9); many.step(9)
    assert np.array_equal(one.gather(), many.gather())
    one.close(); many.close()


def test_dims_create_2d_and_3d():
    L = orc._jac_lib()
    import ctypes as C
    for np_, nd, want in [(1, 2, (1, 1, 1)), (2, 2, (2, 1, 1)), (4, 2, (2, 2, 1)), (6, 2, (3, 2, 1)), (8, 2, (4, 2, 1)),
                          (12, 2, (4, 3, 1)), (8, 3, (2, 2, 2)), (12, 3, (3, 2, 2))]:
        d = (C.c_int * 3)()
        L.jac_dims_create(np_, nd, d)
        assert tuple(d) == want


def test_check_diff_is_max_abs_change_and_updates_previous():
    wd = orc.JacobiWorld((9, 7, 5), 2)
    wd.init()
    wd.step(3)
    a = wd.gather()
    assert wd.check_diff() == np.abs(a).max()     # A_p was the all-zero interior of init()
    assert wd.check_diff() == 0.0
    wd.close()


# ---------------- the Fortran program's text (make_golden_jacobi_fortran.py) ----------------
FGOLD = np.load(os.path.join(HERE, "golden", "ref_fortran_jacobi.npz"))


def test_fortran_jacobi_check_diff_and_init_match_the_text():
    """jacobi2d_mpi.f90:176-180 with a non-zero source term, check_diff :193-198, init :151-165 for a rank that owns the top
    boundary and one that does not"""
    wd = orc.JacobiWorld((7, 6), 1)
    wd.init()
    assert np.array_equal(wd.array(0, "A"), FGOLD["init_A_0"])
    wd.array(0, "A")[...] = FGOLD["A"]; wd.array(0, "f")[...] = FGOLD["f"]
    wd.jacobi()
    assert np.array_equal(wd.array(0, "A")[1:-1, 1:-1], FGOLD["A_new_interior"])        # the roles of A and A_new swapped
    wd.close()
    wd = orc.JacobiWorld((7, 6), 1)
    wd.init()
    wd.array(0, "A")[...] = FGOLD["A"]; wd.array(0, "A_p")[...] = FGOLD["A_p"]
    assert wd.check_diff() == FGOLD["check_diff"][0]
    wd.close()
    wd = orc.JacobiWorld((14, 12), 4, dims=(2, 2))
    wd.init()
    for r, inf in enumerate(wd.info):
        want = FGOLD["init_A_1"] if inf["coords"][1] == 1 else FGOLD["init_A_2"]
        assert np.array_equal(wd.array(r, "A"), want) and np.array_equal(wd.array(r, "A_new"), want), inf
    wd.close()


@pytest.mark.parametrize("nprocs,dims", [(1, None), (4, None), (6, (3, 2))])
def test_oracle_reproduces_the_fortran_programs_loop(nprocs, dims):
    """jacobi2d_mpi.f90:91-112 evaluated from its text on one rank (9 x 7): init, 300 iterations (two sweeps per loop body),
    check_diff every 100 -- bit for bit, also on 4 and 6 emulated ranks"""
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fortran_jacobi.npz"))
    total = tuple(int(x) for x in G["run/shape"])
    wd = orc.JacobiWorld(total, nprocs, dims)
    wd.init()
    for k, itc in enumerate((100, 200, 300)):
        wd.step(100)
        assert wd.check_diff() == G["run/errors"][k]
        assert np.array_equal(wd.gather(), G[f"run/A_{itc}"][1:-1, 1:-1]), itc
    wd.close()
