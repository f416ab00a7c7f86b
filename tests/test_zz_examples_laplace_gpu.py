"""examples/laplace2d_driver.c on the GPU, in the role of the reference's own C program MPI/Laplace/c/laplace2d.c:
  * on small arrays its printed residuals and its final array equal, line for line and bit for bit, what the reference's own
    jacobi() / swap() (oracle/_ref/liblaplace2d_ref.so, the unmodified source) produce under the program's main() loop,
    including the early stop at the program's tolerance;
  * on the shipped 4320 x 4320 array its standard output after 1000 iterations is the reference program's
    (tests/golden/ref_laplace2d_stdout.txt, written by the compiled program itself: make_golden_laplace2d_stdout.py).
(The file sorts last on purpose: it was added after the other GPU files and a failure here must not hide their results under -x.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "liblaplace2d_ref.so")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("laplace") / "laplace2d_driver")
    lib = os.path.join(ROOT, "mglc_b200")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "laplace2d_driver.c"), "-L", lib, "-lmglc", f"-Wl,-rpath,{lib}", "-o", out])
    return out


def reference(nx, ny, itc_max):
    """the reference's jacobi() / swap() under its main() loop, laplace2d.c:41-61"""
    L = C.CDLL(REF_SO)
    dp = C.POINTER(C.c_double)
    L.jacobi.restype = C.c_double
    L.jacobi.argtypes = [dp, dp, C.c_int, C.c_int]
    L.swap.argtypes = [dp, dp, C.c_int, C.c_int]
    A, B = np.zeros((nx, ny)), np.zeros((nx, ny))
    A[:, ny - 1] = 1.0; B[:, ny - 1] = 1.0
    lines, itc, err = [], 0, 1.0
    while err > 1e-5 and itc < itc_max:
        itc += 1
        err = L.jacobi(A.ctypes.data_as(dp), B.ctypes.data_as(dp), nx, ny)
        L.swap(A.ctypes.data_as(dp), B.ctypes.data_as(dp), nx, ny)
        if itc % 100 == 0:
            lines.append("%5d, %0.6f" % (itc, err))
    return lines, A


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/liblaplace2d_ref.so not built (make -C oracle ref)")
@pytest.mark.parametrize("nx,ny,itc_max", [(64, 48, 350), (130, 257, 300), (5, 4, 100), (40, 40, 5000), (515, 301, 200)])
def test_driver_equals_the_reference_functions(exe, tmp_path, nx, ny, itc_max):
    lines, A = reference(nx, ny, itc_max)
    r = subprocess.run([exe, str(nx), str(ny), str(itc_max), "dump.bin"], capture_output=True, text=True, cwd=tmp_path, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout.splitlines() == lines
    assert np.array_equal(np.fromfile(tmp_path / "dump.bin").reshape(nx, ny), A)


def test_driver_prints_the_reference_programs_output_on_the_shipped_array(exe, tmp_path):
    want = open(os.path.join(ROOT, "tests", "golden", "ref_laplace2d_stdout.txt")).read()
    r = subprocess.run([exe], capture_output=True, text=True, cwd=tmp_path, timeout=600)
    assert r.returncode == 0, r.stderr
    assert r.stdout == want
