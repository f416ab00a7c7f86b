"""Generates tests/golden/ref_fortran_lid2d_incomp.npz -- golden vectors of the reference's sequential INCOMPRESSIBLE 2-D
lid-driven cavity program (variant "i") -- and tests/golden/ref_fortran_lid2d_seq.npz -- the same vectors of its sequential
COMPRESSIBLE sibling seq/lid-driven_cavity.f90 (line-aligned with the incompressible file; its subroutines are also those of
2d_old/lid-driven_cavity.f90), which variant "f" must reproduce -- machine-evaluated from the source text (fortran_eval.py)
as whole arrays:

  L2I = /root/reference/MPI/Lid_driven_cavity/fortran/2d/seq/lid-driven_cavity_incompress.f90
  initial      L2I:143-163   (u = U0 on the lid row, omega, f = omega*(...): no rho factor; rho stays 0, L2I:137)
  collision    L2I:181-236   (meq without rho factors)
  streaming    L2I:249-259
  bounceback   L2I:270-293   (lid term without rho)
  macro        L2I:304-310   (u, v undivided)
  check        L2I:322-335   (sums of dsqrt; errorU = error1/error2)
on seeded arrays of an 8 x 7 lattice, plus the program's own run: initial() and N x {collision, streaming, bounceback, macro}
(L2I:60-66) with check() after the last step.  Only numbers are stored; run in the authoring container."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fortran_eval as fe  # noqa: E402
from make_golden_thermal2d import arr, from_full, run_full, to_full  # noqa: E402

L2I = "/root/reference/MPI/Lid_driven_cavity/fortran/2d/seq/lid-driven_cavity_incompress.f90"
EX = [0, 1, 0, -1, 0, 1, -1, -1, 1]     # L2I:27-28
EY = [0, 0, 1, 0, -1, 1, 1, -1, -1]
FULL = ["f", "f_post", "rho", "u", "v", "up", "vp", "ex", "ey", "omega", "un", "uwall", "s", "m", "m_post", "meq"]


def main(L2I=L2I, name="ref_fortran_lid2d_incomp.npz", rho_init=0.0):
    rng = np.random.default_rng(20310)
    nx, ny, U0, Re = 8, 7, 0.1, 1000.0
    tau = U0 * float(nx) / Re * 3.0 + 0.5                             # L2I:11
    snu, sq = 1.0 / tau, 8.0 * (2.0 * tau - 1.0) / (8.0 * tau - 1.0)   # L2I:33
    sc = dict(nx=nx, ny=ny, u0=U0, snu=snu, sq=sq, itc=0, rho0=1.0)
    out = {"params": np.array([tau, snu, sq]), "shape": np.array([nx, ny])}
    tr = lambda a, b: fe.translate(fe.read_lines(L2I, a, b), full_arrays=FULL)
    src = {"initial": tr(143, 163), "collision": tr(181, 236), "streaming": tr(249, 259), "bounceback": tr(270, 293),
           "macro": tr(304, 310), "check": tr(322, 335)}
    tables = lambda: {"ex": arr(EX), "ey": arr(EY), "uwall": fe._Arr({1: U0, 2: 0.0}), "omega": fe._Arr(), "un": fe._Arr(),
                      "s": fe._Arr(), "m": fe._Arr(), "m_post": fe._Arr(), "meq": fe._Arr()}
    F3, H3, S2 = (0, 1, 1), (0, 0, 0), (1, 1)

    # ---------------- every subroutine on seeded arrays ----------------
    f0, fp = rng.random((9, nx, ny)), rng.random((9, nx + 2, ny + 2))
    rho = 1.0 + 0.05 * rng.uniform(-1, 1, (nx, ny))
    u, v, up, vp = (0.08 * rng.uniform(-1, 1, (nx, ny)) for _ in range(4))
    for k, a in (("f0", f0), ("f_post", fp), ("rho", rho), ("u", u), ("v", v), ("up", up), ("vp", vp)):
        out["in/" + k] = a
    ns = run_full(src["collision"], {**tables(), "f": to_full(f0, F3), "f_post": fe._Arr(), "rho": to_full(rho, S2),
                                     "u": to_full(u, S2), "v": to_full(v, S2)}, sc)
    out["collision/f_post"] = from_full(ns["f_post__"], (9, nx, ny), F3)
    ns = run_full(src["streaming"], {**tables(), "f": fe._Arr(), "f_post": to_full(fp, H3)}, sc)
    out["streaming/f"] = from_full(ns["f__"], (9, nx, ny), F3)
    ns = run_full(src["bounceback"], {**tables(), "f": to_full(f0, F3), "f_post": to_full(fp, H3), "rho": to_full(rho, S2)}, sc)
    out["bounceback/f"] = from_full(ns["f__"], (9, nx, ny), F3)
    ns = run_full(src["macro"], {**tables(), "f": to_full(f0, F3), "rho": fe._Arr(), "u": fe._Arr(), "v": fe._Arr()}, sc)
    out["macro/ruv"] = np.stack([from_full(ns[k + "__"], (nx, ny), S2) for k in ("rho", "u", "v")])
    ns = run_full(src["check"], {**tables(), **{k: to_full(a, S2) for k, a in (("u", u), ("v", v), ("up", up), ("vp", vp))}}, sc)
    out["check/e1_e2_errorU"] = np.array([ns["error1"], ns["error2"], ns["erroru"]])

    # ---------------- the program's own run ----------------
    zeros = lambda: to_full(np.zeros((nx, ny)), S2)                  # L2I:137-141: rho = u = v = up = vp = 0 (seq: rho = rho0)
    st = {**tables(), "f": fe._Arr(), "f_post": to_full(np.zeros((9, nx + 2, ny + 2)), H3), "rho": to_full(np.full((nx, ny), rho_init), S2), "u": zeros(),
          "v": zeros(), "up": zeros(), "vp": zeros()}
    names = list(st)

    def call(sub):
        ns = run_full(src[sub], st, sc)
        for k in names:
            st[k] = ns[k + "__"]
        return ns

    call("initial")
    out["run0/f"] = from_full(st["f"], (9, nx, ny), F3)
    out["run0/ruv"] = np.stack([from_full(st[k], (nx, ny), S2) for k in ("rho", "u", "v")])
    done = 0
    for n in (1, 2, 20):
        for _ in range(n - done):
            for sub in ("collision", "streaming", "bounceback", "macro"):
                call(sub)
        done = n
        out[f"run{n}/f"] = from_full(st["f"], (9, nx, ny), F3)
        out[f"run{n}/ruv"] = np.stack([from_full(st[k], (nx, ny), S2) for k in ("rho", "u", "v")])
    ns = call("check")
    out["run20/check"] = np.array([ns["error1"], ns["error2"], ns["erroru"]])
    for _ in range(5):
        for sub in ("collision", "streaming", "bounceback", "macro"):
            call(sub)
    ns = call("check")
    out["run25/check"] = np.array([ns["error1"], ns["error2"], ns["erroru"]])

    path = os.path.join(HERE, name)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
    main("/root/reference/MPI/Lid_driven_cavity/fortran/2d/seq/lid-driven_cavity.f90", "ref_fortran_lid2d_seq.npz", rho_init=1.0)
