"""Generates tests/golden/ref_fortran_jacobi.npz -- golden vectors of the Laplace driver's Fortran program, machine-evaluated
from the REFERENCE's own source text (fortran_eval.py):

  LAP = /root/reference/MPI/Laplace/fortran/jacobi2d_mpi.f90
  init        LAP:151-165  (A = A_new = f = 0, the top halo row = 1 on the ranks that own the top boundary)
  jacobi      LAP:176-180  (A_new = 0.25*(A(i-1,j)+A(i+1,j)+A(i,j-1)+A(i,j+1)+f(i,j)); the literal 0.25 is single precision, exact)
  check_diff  LAP:193-198  (max |A_p - A| over the interior)
  run         LAP:91-112   (the program's loop on one rank: 300 iterations = 150 x two sweeps, check_diff every 100)
Only numbers are stored; run in the authoring container."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fortran_eval as fe  # noqa: E402
from make_golden_thermal2d import arr, from_full, run_full, to_full  # noqa: E402

LAP = "/root/reference/MPI/Laplace/fortran/jacobi2d_mpi.f90"


def main():
    rng = np.random.default_rng(20306)
    nx, ny = 7, 6
    out = {}
    A, f, Ap = rng.random((nx + 2, ny + 2)), 0.1 * rng.random((nx + 2, ny + 2)), rng.random((nx + 2, ny + 2))
    out["A"], out["f"], out["A_p"] = A, f, Ap
    full = ["a", "a_new", "f", "a_p", "coords", "dims"]
    sc = dict(nx=nx, ny=ny, max=max)
    ns = run_full(fe.translate(fe.read_lines(LAP, 176, 180), full_arrays=full),
                  {"a": to_full(A, (0, 0)), "a_new": fe._Arr(), "f": to_full(f, (0, 0))}, sc)
    new = np.zeros((nx + 2, ny + 2))
    new[1:-1, 1:-1] = from_full(ns["a_new__"], (nx, ny), (1, 1))
    out["A_new_interior"] = new[1:-1, 1:-1]
    ns = run_full(fe.translate(fe.read_lines(LAP, 193, 198), full_arrays=full), {"a": to_full(A, (0, 0)), "a_p": to_full(Ap, (0, 0))}, sc)
    out["check_diff"] = np.array([ns["error"]])
    for k, (co, di) in enumerate([((0, 0), (1, 1)), ((0, 1), (2, 2)), ((1, 0), (2, 2))]):      # MPI coordinates (0-based) in arrays indexed 1..2
        ns = run_full(fe.translate(fe.read_lines(LAP, 151, 165), full_arrays=full),
                      {"a": fe._Arr(), "a_new": fe._Arr(), "f": fe._Arr(), "coords": arr(co, 1), "dims": arr(di, 1)}, sc)
        out[f"init_A_{k}"] = from_full(ns["a__"], (nx + 2, ny + 2), (0, 0))
    # ---- the program's own loop (LAP:91-112) on one rank: init, then per iteration two sweeps (A -> A_new -> A, the halo
    # exchange has no neighbours), check_diff + A_p = A every 100 iterations ----
    rx, ry = 9, 7
    init = fe.translate(fe.read_lines(LAP, 151, 165), full_arrays=full)
    sweep = fe.translate(fe.read_lines(LAP, 176, 180), full_arrays=full)
    diff = fe.translate(fe.read_lines(LAP, 193, 198), full_arrays=full)
    scr = dict(nx=rx, ny=ry, max=max)
    ns = run_full(init, {"a": fe._Arr(), "a_new": fe._Arr(), "f": fe._Arr(), "coords": arr((0, 0), 1), "dims": arr((1, 1), 1)}, scr)
    A_, An_, f_ = ns["a__"], ns["a_new__"], ns["f__"]
    Ap_ = fe._Arr(A_)                                                                    # LAP:93
    itc, errs = 0, []
    while itc < 300:
        itc += 2                                                                         # LAP:95
        An_ = run_full(sweep, {"a": A_, "a_new": An_, "f": f_}, scr)["a_new__"]           # jacobi(A, A_new, f), LAP:99
        A_ = run_full(sweep, {"a": An_, "a_new": A_, "f": f_}, scr)["a_new__"]            # jacobi(A_new, A, f), LAP:103
        if itc % 100 == 0:                                                               # LAP:105-107
            errs.append(run_full(diff, {"a": A_, "a_p": Ap_}, scr)["error"])
            Ap_ = fe._Arr(A_)                                                            # LAP:200
            out[f"run/A_{itc}"] = from_full(A_, (rx + 2, ry + 2), (0, 0))
    out["run/shape"] = np.array([rx, ry])
    out["run/errors"] = np.array(errs)
    path = os.path.join(HERE, "ref_fortran_jacobi.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
