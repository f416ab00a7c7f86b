"""Generates tests/golden/ref_fortran_particles_case1.npz: the two pieces of the reference's OTHER particle scenario
(MPI/Micro_particles/fortran/case1/mpi_complete, "P1") that the particle path offers as options -- evaluated from the
reference's Fortran TEXT (tests/golden/fortran_eval.py), never from a restatement:

  * linear-interpolated bounce-back on the particle surface, P1/particle_bounceback.F90:67-75 (`#ifdef linear`), with calQ
    (P1/particle_bounceback.F90:126-145) on links crossing a particle of the scenario's radius 25.25/2;
  * moving top / bottom walls, P1/fluid.F90:130-145 (`movingFrame`) and :151-168 (`stationaryFrame`).

    python tests/golden/make_golden_particles_case1.py      (needs /root/reference; the outputs are committed)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fortran_eval as fe  # noqa: E402
from make_golden_particles import EX, EY, OMEGA, R, arr  # noqa: E402

P1 = "/root/reference/MPI/Micro_particles/fortran/case1/mpi_complete"


def main():
    rng = np.random.default_rng(777)
    out = {}
    # ---- links crossing the surface of one particle: calQ as written there, then the `linear` branch ----
    calq_src = fe.translate(fe.read_lines(P1 + "/particle_bounceback.F90", 126, 145), full_arrays=["xcenter", "ycenter", "radius", "ex", "ey"])
    bb_src = fe.translate(fe.read_lines(P1 + "/particle_bounceback.F90", 62, 63) + "\n" + fe.read_lines(P1 + "/particle_bounceback.F90", 67, 75),
                          full_arrays=["xcenter", "ycenter", "radius", "ex", "ey", "r", "omega", "uc", "vc", "rationalomega", "f", "f_post"])
    xc, yc, rad = 100.43, 50.77, 25.25 / 2.0
    Uc, Vc, om, rhoAvg = 0.017, -0.004, 0.0031, 0.99987
    links = []
    for i in range(int(xc - rad - 2), int(xc + rad + 3)):
        for j in range(int(yc - rad - 2), int(yc + rad + 3)):
            if (i - xc) ** 2 + (j - yc) ** 2 <= rad * rad:
                continue
            for a in range(1, 9):
                if (i + EX[a] - xc) ** 2 + (j + EY[a] - yc) ** 2 <= rad * rad:
                    links.append((i, j, a))
    links = [links[q] for q in rng.permutation(len(links))[:64]]
    rec = []
    for (i, j, a) in links:
        ns = {"xcenter__": arr([xc], 1), "ycenter__": arr([yc], 1), "radius__": arr([rad], 1), "ex__": arr(EX), "ey__": arr(EY),
              "cnum": 1, "alpha": a, "i": float(i), "j": float(j), "epsradius": float(np.float32(1e-9))}
        q = fe.run(calq_src + "\nq_out__ = q\nx0_out__ = x0\ny0_out__ = y0", scalars=ns, field_out=["q_out", "x0_out", "y0_out"])
        fpatch = {}
        for b in range(9):
            for di in range(-2, 3):
                for dj in range(-2, 3):
                    fpatch[(b, i + di, j + dj)] = OMEGA[b] * (1 + 0.1 * rng.uniform(-1, 1))
        fp = fe._Arr(fpatch)
        f_bb = fe._Arr({k: v * (1 + 0.05 * rng.uniform(-1, 1)) for k, v in fpatch.items()})
        common = {"xcenter__": arr([xc], 1), "ycenter__": arr([yc], 1), "radius__": arr([rad], 1), "ex__": arr(EX), "ey__": arr(EY),
                  "r__": fe._Arr(R), "omega__": arr(OMEGA), "uc__": arr([Uc], 1), "vc__": arr([Vc], 1), "rationalomega__": arr([om], 1),
                  "cnum": 1, "alpha": a, "i": i, "j": j, "q": q["q_out"], "x0": q["x0_out"], "y0": q["y0_out"], "rhoavg": rhoAvg}
        fe.run(bb_src, scalars={**common, "f__": f_bb, "f_post__": fp})
        rec.append(dict(link=(i, j, a), q=(q["q_out"], q["x0_out"], q["y0_out"]),
                        fpost=[[fp[(b, i - s * EX[a], j - s * EY[a])] for b in range(9)] for s in range(2)], bb=f_bb[(R[a], i, j)]))
    out["link/particle"] = np.array([xc, yc, rad, Uc, Vc, om, rhoAvg])
    out["link/ija"] = np.array([r["link"] for r in rec])
    out["link/q_x0_y0"] = np.array([r["q"] for r in rec])
    out["link/fpost_0_1"] = np.array([r["fpost"] for r in rec])            # f_post(:, x - s e_alpha), s = 0, 1
    out["link/bb_linear"] = np.array([r["bb"] for r in rec])
    assert (out["link/q_x0_y0"][:, 0] < 0.5).any() and (out["link/q_x0_y0"][:, 0] >= 0.5).any()

    # ---- moving walls: one row of nodes along the bottom wall and one along the top wall ----
    nx, ny, Uwall, U0 = 23, 9, 0.1, 0.02
    for frame, lines in (("moving", ((130, 135), (140, 145))), ("stationary", ((152, 157), (162, 167)))):
        fpost = {(b, i, j): OMEGA[b] * (1 + 0.2 * rng.uniform(-1, 1)) for b in range(9) for i in range(1, nx + 1) for j in (1, ny)}
        f = fe._Arr({k: 0.0 for k in fpost})
        src = fe.translate(fe.read_lines(P1 + "/fluid.F90", *lines[0]) + "\n" + fe.read_lines(P1 + "/fluid.F90", *lines[1]),
                           full_arrays=["f", "f_post"])
        fe.run(src, scalars={"f__": f, "f_post__": fe._Arr(fpost), "nx": nx, "ny": ny, "uwall": Uwall, "u0": U0})
        out[f"walls/{frame}/f_post"] = np.array([[[fpost[(b, i, j)] for b in range(9)] for i in range(1, nx + 1)] for j in (1, ny)])
        out[f"walls/{frame}/f"] = np.array([[[f[(b, i, j)] for b in range(9)] for i in range(1, nx + 1)] for j in (1, ny)])
    out["walls/params"] = np.array([nx, ny, Uwall, U0])
    path = os.path.join(HERE, "ref_fortran_particles_case1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
