"""Generates tests/golden/ref_fortran_lid2d_fields.npz -- whole-array golden vectors of the 2-D lid driver's Fortran program
(variant "f"), machine-evaluated from the REFERENCE's own source text (fortran_eval.py) on a seeded 6 x 5 block with a one-cell
halo ring, for the single rank, the interior block and the four corner blocks of a 3 x 3 process grid:

  L2F = /root/reference/MPI/Lid_driven_cavity/fortran/2d/2d_revised/mpi_blocked
  streaming    L2F/evolution.f90:82-91
  bounceback   L2F/bounceback.f90:7-42   (left, right, bottom, then the moving top wall: it wins in the two top corners)
  check        L2F/evolution.f90:124-132 (the two rank sums)
Only numbers are stored; run in the authoring container."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fortran_eval as fe  # noqa: E402
from make_golden_thermal2d import arr, from_full, run_full, to_full  # noqa: E402

L2F = "/root/reference/MPI/Lid_driven_cavity/fortran/2d/2d_revised/mpi_blocked"
EX = [0, 1, 0, -1, 0, 1, -1, -1, 1]
EY = [0, 0, 1, 0, -1, 1, 1, -1, -1]


def main():
    rng = np.random.default_rng(20305)
    nx, ny = 6, 5
    out = {}
    fp, f0 = rng.random((9, nx + 2, ny + 2)), rng.random((9, nx, ny))
    rho = 1.0 + 0.05 * rng.uniform(-1, 1, (nx, ny))
    out["f_post"], out["f0"], out["rho"] = fp, f0, rho
    full = ["f", "f_post", "rho", "ex", "ey", "coords", "dims"]
    sc = dict(nx=nx, ny=ny, uwall=0.1)
    ns = run_full(fe.translate(fe.read_lines(L2F + "/evolution.f90", 82, 91), full_arrays=full),
                  {"f": fe._Arr(), "f_post": to_full(fp, (0, 0, 0)), "ex": arr(EX), "ey": arr(EY)}, sc)
    out["streaming_f"] = from_full(ns["f__"], (9, nx, ny), (0, 1, 1))
    cases = [((0, 0), (1, 1)), ((1, 1), (3, 3)), ((0, 0), (3, 3)), ((2, 0), (3, 3)), ((0, 2), (3, 3)), ((2, 2), (3, 3)), ((1, 2), (3, 3))]
    out["bb_cases"] = np.array([c + d for c, d in cases])
    src = fe.translate(fe.read_lines(L2F + "/bounceback.f90", 7, 42), full_arrays=full)
    for k, (co, di) in enumerate(cases):
        ns = run_full(src, {"f": to_full(f0, (0, 1, 1)), "f_post": to_full(fp, (0, 0, 0)), "rho": to_full(rho, (1, 1)),
                            "coords": arr(co), "dims": arr(di)}, sc)
        out[f"bounceback_{k}"] = from_full(ns["f__"], (9, nx, ny), (0, 1, 1))
    fl = {k: rng.uniform(-0.1, 0.1, (nx, ny)) for k in ("u", "v", "up", "vp")}
    for k, a in fl.items():
        out[f"check_{k}"] = a
    ns = run_full(fe.translate(fe.read_lines(L2F + "/evolution.f90", 124, 132), full_arrays=["u", "v", "up", "vp"]),
                  {k: to_full(a, (1, 1)) for k, a in fl.items()}, sc)
    out["check_sums"] = np.array([ns["error1"], ns["error2"]])
    path = os.path.join(HERE, "ref_fortran_lid2d_fields.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
