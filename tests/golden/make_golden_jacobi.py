"""Generates tests/golden/laplace2d_ref.npz from the REFERENCE's own compiled code: the unmodified
MPI/Laplace/c/laplace2d.c built as a shared object by `make -C oracle ref` (oracle/_ref/liblaplace2d_ref.so),
whose jacobi() and swap() are called on small grids.  Run in the authoring container (needs /root/reference).

Stored per case: the initial grid (C layout [x][y], boundary included), the grid after `its` iterations
of jacobi()+swap(), and the error each jacobi() returned."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def ref_lib():
    path = os.path.join(ROOT, "oracle", "_ref", "liblaplace2d_ref.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    L = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    L.jacobi.restype = C.c_double
    L.jacobi.argtypes = [dp, dp, C.c_int, C.c_int]
    L.swap.restype = None
    L.swap.argtypes = [dp, dp, C.c_int, C.c_int]
    return L


def run_reference(L, A0, its):
    """A0: C-contiguous (nx, ny) grid incl. boundary.  Returns (A after its x [jacobi, swap], errors)."""
    nx, ny = A0.shape
    A = np.ascontiguousarray(A0, dtype=np.float64).copy()
    B = A.copy()
    dp = C.POINTER(C.c_double)
    errs = []
    for _ in range(its):
        errs.append(L.jacobi(A.ctypes.data_as(dp), B.ctypes.data_as(dp), nx, ny))
        L.swap(A.ctypes.data_as(dp), B.ctypes.data_as(dp), nx, ny)
    return A, np.array(errs)


def cases():
    rng = np.random.default_rng(2021)
    # the reference's own problem (top boundary y = ny-1 set to 1, laplace2d.c:40-50), small
    a = np.zeros((19, 14)); a[:, -1] = 1.0
    yield "shipped_bc_19x14", a, 25
    yield "random_23x37", rng.random((23, 37)), 7
    yield "random_130x9", rng.uniform(-1, 1, (130, 9)), 3


if __name__ == "__main__":
    L = ref_lib()
    out = {}
    for name, A0, its in cases():
        A, errs = run_reference(L, A0, its)
        out[name + "/A0"] = A0
        out[name + "/A"] = A
        out[name + "/err"] = errs
    np.savez_compressed(os.path.join(HERE, "laplace2d_ref.npz"), **out)
    print("wrote", len(out), "arrays")
