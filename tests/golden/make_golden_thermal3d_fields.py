"""Generates tests/golden/ref_fortran_thermal3d_fields.npz -- whole-array golden vectors of the 3-D thermal driver's copy-type
subroutines, machine-evaluated from the REFERENCE's own source text (fortran_eval.py) on a seeded 5 x 4 x 3 block with one-cell
halos, for every position of the block in a 3 x 3 x 3 process grid (and the single rank), with both boundary macro sets:

  B3 = /root/reference/MPI/Buoyancy_driven_cavity/fortran/3d/bouyancy3d_mpi.F90
  streamingT   B3:1081-1094
  bounceback   B3:900-980    (#define noslipWalls)
  bouncebackT  B3:1106-1207  benchmarkCavity set (:16-19: back/front adiabatic, left/right constant T, plates adiabatic) and
                             RBconvection set (:9-12: back/front adiabatic, left/right adiabatic, plates constant T)
  check        B3:1242-1264  (the four rank sums)
paraA is evaluated from B3:26-47 for the cavity height of the world the test builds around the block (total_nz = 9, or 3).
Only numbers are stored; run in the authoring container."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fortran_eval as fe  # noqa: E402
from make_golden_fortran import B3, EX, EY, EZ, eval_parameters  # noqa: E402
from make_golden_thermal2d import arr, from_full, run_full, strip_cpp, to_full  # noqa: E402


def params(total_n):
    text = fe.read_lines(B3, 26, 47).replace("total_nx = 51", f"total_nx = {total_n}") + "\n" + fe.read_lines(B3, 73, 74)
    return eval_parameters(text)


def main():
    rng = np.random.default_rng(20304)
    nx, ny, nz = 5, 4, 3
    out = {}
    fp, gp = rng.random((19, nx + 2, ny + 2, nz + 2)), rng.random((7, nx + 2, ny + 2, nz + 2))
    f0, g0 = rng.random((19, nx, ny, nz)), rng.random((7, nx, ny, nz))
    out["f_post"], out["g_post"], out["f0"], out["g0"] = fp, gp, f0, g0
    full = ["f", "f_post", "g", "g_post", "ex", "ey", "ez", "coords", "dims"]
    sc = dict(nx=nx, ny=ny, nz=nz, thot=1.0, tcold=0.0)
    text = "\n".join(l for l in fe.read_lines(B3, 1081, 1094).splitlines() if not l.strip().lower().startswith("!$omp"))
    ns = run_full(fe.translate(text, full_arrays=full),
                  {"g": fe._Arr(), "g_post": to_full(gp, (0, 0, 0, 0)), "ex": arr(EX), "ey": arr(EY), "ez": arr(EZ)}, sc)
    out["streamingT_g"] = from_full(ns["g__"], (7, nx, ny, nz), (0, 1, 1, 1))
    cases = [((0, 0, 0), (1, 1, 1)), ((1, 1, 1), (3, 3, 3)), ((0, 1, 2), (3, 3, 3)), ((2, 1, 0), (3, 3, 3)), ((1, 2, 0), (3, 3, 3))] + \
            [((a, b, c), (3, 3, 3)) for a in (0, 2) for b in (0, 2) for c in (0, 2)]     # single rank, interior, mixed, the 8 corners
    out["bb_cases"] = np.array([c + d for c, d in cases])
    sets = {"cavity": {"noslipWalls", "benchmarkCavity", "BackFrontWallsAdiabatic", "LeftRightWallsConstT", "TopBottomPlatesAdiabatic"},
            "rb": {"noslipWalls", "RBconvection", "BackFrontWallsAdiabatic", "LeftRightWallsAdiabatic", "TopBottomPlatesConstT"}}
    bb = fe.translate(strip_cpp(fe.read_lines(B3, 900, 980), sets["cavity"]), full_arrays=full)
    for k, (co, di) in enumerate(cases):
        ns = run_full(bb, {"f": to_full(f0, (0, 1, 1, 1)), "f_post": to_full(fp, (0, 0, 0, 0)), "coords": arr(co), "dims": arr(di)}, sc)
        out[f"bounceback_{k}"] = from_full(ns["f__"], (19, nx, ny, nz), (0, 1, 1, 1))
        paraa = params(nz * di[2])["paraa"]
        out[f"paraA_{k}"] = np.array([paraa])
        for tag, defs in sets.items():
            src = fe.translate(strip_cpp(fe.read_lines(B3, 1106, 1207), defs), full_arrays=full)
            ns = run_full(src, {"g": to_full(g0, (0, 1, 1, 1)), "g_post": to_full(gp, (0, 0, 0, 0)), "coords": arr(co), "dims": arr(di)},
                          dict(sc, paraa=paraa))
            out[f"bouncebackT_{tag}_{k}"] = from_full(ns["g__"], (7, nx, ny, nz), (0, 1, 1, 1))
    fl = {k: rng.uniform(-0.1, 0.1, (nx, ny, nz)) for k in ("u", "v", "w", "up", "vp", "wp")}
    fl["T"], fl["Tp"] = rng.uniform(-0.2, 1.2, (nx, ny, nz)), rng.uniform(-0.2, 1.2, (nx, ny, nz))
    for k, a in fl.items():
        out[f"check_{k}"] = a
    src = fe.translate(fe.read_lines(B3, 1242, 1264), full_arrays=["u", "v", "w", "t", "up", "vp", "wp", "tp"])
    ns = run_full(src, {k.lower(): to_full(a, (1, 1, 1)) for k, a in fl.items()}, sc)
    out["check_sums"] = np.array([ns["error1"], ns["error2"], ns["error5"], ns["error6"]])
    path = os.path.join(HERE, "ref_fortran_thermal3d_fields.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
