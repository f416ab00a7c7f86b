"""Generates tests/golden/ref_fortran_lid3d_seq_run.npz -- the reference's SEQUENTIAL 3-D lid-driven cavity program run from its
own source text (fortran_eval.py, whole arrays) on a 6 x 5 x 4 lattice:

  L3S = /root/reference/MPI/Lid_driven_cavity/fortran/3d/seq/lid_driven_cavity_3d.f90
  initial      L3S:166-191  (the whole-array assignments :166-172 are applied by this script, the loops are evaluated)
  collision    L3S:212-399
  streaming    L3S:412-424
  bounceback   L3S:435-489
  macro        L3S:504-519  (after the zeroing of :500-503)
  check        L3S:531-541, :547
its loop (L3S:68-94: collision, streaming, bounceback, macro) for 1, 2 and 12 iterations with check() after 12 and 14.  The MPI
program's restatement (oracle/lid3d.c) must reproduce this run on 1 and on several emulated ranks: the reference's seq == MPI
contract, from the sequential program's own text.  Only numbers are stored; run in the authoring container."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fortran_eval as fe  # noqa: E402
from make_golden_thermal2d import arr, from_full, run_full, to_full  # noqa: E402

L3S = "/root/reference/MPI/Lid_driven_cavity/fortran/3d/seq/lid_driven_cavity_3d.f90"
EX = [0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0]      # L3S:36-44
EY = [0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1]
EZ = [0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1]
OMEGA = [1.0 / 3.0] + [1.0 / 18.0] * 6 + [1.0 / 36.0] * 12               # L3S:33-35
FULL = ["f", "f_post", "rho", "u", "v", "w", "up", "vp", "wp", "ex", "ey", "ez", "omega", "un", "s", "m", "m_post", "meq"]


def main():
    nx, ny, nz, U0, Re = 6, 5, 4, 0.1, 1000.0
    tau = U0 * float(nx) / Re * 3.0 + 0.5                               # L3S:12
    snu, sq = 1.0 / tau, 8.0 * (2.0 * tau - 1.0) / (8.0 * tau - 1.0)     # L3S:46
    sc = dict(nx=nx, ny=ny, nz=nz, u0=U0, rho0=1.0, snu=snu, sq=sq, itc=0)
    out = {"params": np.array([tau, snu, sq]), "shape": np.array([nx, ny, nz])}
    tr = lambda a, b: fe.translate(fe.read_lines(L3S, a, b), full_arrays=FULL)
    src = {"initial": tr(175, 191), "collision": tr(212, 399), "streaming": tr(412, 424), "bounceback": tr(435, 489),
           "macro": tr(504, 519), "check": tr(531, 541)}
    F4, H4, S3 = (0, 1, 1, 1), (0, 0, 0, 0), (1, 1, 1)
    field = lambda value: to_full(np.full((nx, ny, nz), value), S3)
    st = {"ex": arr(EX), "ey": arr(EY), "ez": arr(EZ), "omega": arr(OMEGA), "un": fe._Arr(), "s": fe._Arr(), "m": fe._Arr(),
          "m_post": fe._Arr(), "meq": fe._Arr(), "f": fe._Arr(), "f_post": to_full(np.zeros((19, nx + 2, ny + 2, nz + 2)), H4),
          "rho": field(1.0), "u": field(0.0), "v": field(0.0), "w": field(0.0), "up": field(0.0), "vp": field(0.0), "wp": field(0.0)}
    names = list(st)

    def call(sub, **scalars):
        ns = run_full(src[sub], st, {**sc, **scalars})
        for k in names:
            st[k] = ns[k + "__"]
        return ns

    def macro():
        for k in ("rho", "u", "v", "w"):                                # L3S:500-503
            st[k] = field(0.0)
        call("macro")

    def check():
        ns = call("check", error1=0.0, error2=0.0)                      # L3S:528-529
        for a, b in (("up", "u"), ("vp", "v"), ("wp", "w")):            # L3S:543-545
            st[a] = fe._Arr(st[b])
        return np.array([ns["error1"], ns["error2"], np.sqrt(ns["error1"]) / np.sqrt(ns["error2"])])      # L3S:547

    snap = lambda: (from_full(st["f"], (19, nx, ny, nz), F4), np.stack([from_full(st[k], (nx, ny, nz), S3) for k in ("rho", "u", "v", "w")]))
    call("initial")
    out["run0/f"], out["run0/ruvw"] = snap()
    done = 0
    for n in (1, 2, 12):
        for _ in range(n - done):
            call("collision"); call("streaming"); call("bounceback"); macro()
        done = n
        out[f"run{n}/f"], out[f"run{n}/ruvw"] = snap()
        out[f"run{n}/f_post"] = from_full(st["f_post"], (19, nx, ny, nz), F4)          # interior of f_post
    out["run12/check"] = check()
    for _ in range(2):
        call("collision"); call("streaming"); call("bounceback"); macro()
    out["run14/check"] = check()
    out["run14/f"], out["run14/ruvw"] = snap()
    path = os.path.join(HERE, "ref_fortran_lid3d_seq_run.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
