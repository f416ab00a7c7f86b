"""Generates tests/golden/ref_fortran_lid3d_fields.npz -- whole-array golden vectors of the 3-D lid driver's copy-type
subroutines, machine-evaluated from the REFERENCE's own source text (fortran_eval.py) on a seeded 5 x 4 x 3 block with one-cell
halos, for every position of the block in the process grid that changes which walls it owns:

  L3 = /root/reference/MPI/Lid_driven_cavity/fortran/3d/mpi_3d_blocked
  streaming    L3/streaming.f90:8-20
  bounceback   L3/bounce_back.f90:6-83   (x-, x+, y-, y+, z-, then the moving lid z+: the later wall wins on edges; the lid term
                                          uses rho of the previous macro())
  check        L3/check.f90:9-19         (the two rank sums: no w term in error1)
Only numbers are stored; run in the authoring container."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fortran_eval as fe  # noqa: E402
from make_golden_thermal2d import arr, from_full, run_full, to_full  # noqa: E402

L3 = "/root/reference/MPI/Lid_driven_cavity/fortran/3d/mpi_3d_blocked"
EX = [0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0]
EY = [0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1]
EZ = [0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1]


def main():
    rng = np.random.default_rng(20303)
    nx, ny, nz = 5, 4, 3
    out = {}
    fp = rng.random((19, nx + 2, ny + 2, nz + 2))
    f0 = rng.random((19, nx, ny, nz))
    rho = 1.0 + 0.05 * rng.uniform(-1, 1, (nx, ny, nz))
    out["f_post"], out["f0"], out["rho"] = fp, f0, rho
    full = ["f", "f_post", "rho", "ex", "ey", "ez", "coords", "dims"]
    sc = dict(nx=nx, ny=ny, nz=nz, u0=0.1)
    src = fe.translate(fe.read_lines(L3 + "/streaming.f90", 8, 20), full_arrays=full)
    ns = run_full(src, {"f": fe._Arr(), "f_post": to_full(fp, (0, 0, 0, 0)), "ex": arr(EX), "ey": arr(EY), "ez": arr(EZ)}, sc)
    out["streaming_f"] = from_full(ns["f__"], (19, nx, ny, nz), (0, 1, 1, 1))
    # the single rank, an interior block, three mixed positions and the eight corner blocks of a 3 x 3 x 3 grid
    cases = [((0, 0, 0), (1, 1, 1)), ((1, 1, 1), (3, 3, 3)), ((0, 1, 2), (3, 3, 3)), ((2, 1, 0), (3, 3, 3)), ((1, 2, 0), (3, 3, 3))] + \
            [((a, b, c), (3, 3, 3)) for a in (0, 2) for b in (0, 2) for c in (0, 2)]     # single rank, interior, mixed, the 8 corners
    out["bb_cases"] = np.array([c + d for c, d in cases])
    src = fe.translate(fe.read_lines(L3 + "/bounce_back.f90", 6, 83), full_arrays=full)
    for k, (co, di) in enumerate(cases):
        ns = run_full(src, {"f": to_full(f0, (0, 1, 1, 1)), "f_post": to_full(fp, (0, 0, 0, 0)), "rho": to_full(rho, (1, 1, 1)),
                            "coords": arr(co), "dims": arr(di)}, sc)
        out[f"bounceback_{k}"] = from_full(ns["f__"], (19, nx, ny, nz), (0, 1, 1, 1))
    fl = {k: rng.uniform(-0.1, 0.1, (nx, ny, nz)) for k in ("u", "v", "w", "up", "vp")}
    for k, a in fl.items():
        out[f"check_{k}"] = a
    src = fe.translate(fe.read_lines(L3 + "/check.f90", 9, 19), full_arrays=["u", "v", "w", "up", "vp"])
    ns = run_full(src, {k: to_full(a, (1, 1, 1)) for k, a in fl.items()}, sc)
    out["check_sums"] = np.array([ns["error1"], ns["error2"]])
    path = os.path.join(HERE, "ref_fortran_lid3d_fields.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
