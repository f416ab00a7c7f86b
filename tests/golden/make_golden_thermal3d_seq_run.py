"""Generates tests/golden/ref_fortran_thermal3d_seq_run.npz -- the reference's SEQUENTIAL 3-D buoyancy-driven cavity program run
from its own source text (fortran_eval.py, whole arrays) on a 6 x 5 x 4 lattice with its shipped macro set (:71-85: no-slip
walls, benchmarkCavity: back/front adiabatic, left/right constant T, plates adiabatic):

  B3S = /root/reference/MPI/Buoyancy_driven_cavity/fortran/3d/seq/bouyancy3d.F90
  parameters   B3S:10-28, :54-55 (nx, ny, nz replaced by 6, 5, 4)
  initial      B3S:291-302 (weights), :316-347 (T on the constant-T walls of the macro set), :350-364 (populations); the whole-array assignments
               (:289, :307-310, :386-392) are applied by this script
  collision    B3S:415-624     streaming   B3S:639-652     bounceback   B3S:664-723
  collisionT   B3S:767-809     streamingT  B3S:824-837     bouncebackT  B3S:849-919
  macro        B3S:736-750     macroT      B3S:931-937     check        B3S:950-976
its loop (B3S:117-133: collision, streaming, bounceback, collisionT, streamingT, bouncebackT, macro, macroT) for 1, 2 and 10
iterations with check() after 10 and 12; a second file holds the same run with the program's other macro set (RBconvection:
plates at constant temperature, side walls adiabatic).  The restatement of the MPI program (oracle/thermal3d.c) must reproduce this run on 1
and on several emulated ranks.  Only numbers are stored; run in the authoring container."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fortran_eval as fe  # noqa: E402
from make_golden_fortran import EX, EY, EZ, eval_parameters  # noqa: E402
from make_golden_thermal2d import arr, from_full, run_full, strip_cpp, to_full  # noqa: E402

B3S = "/root/reference/MPI/Buoyancy_driven_cavity/fortran/3d/seq/bouyancy3d.F90"
DEFS = {"noslipWalls", "benchmarkCavity", "BackFrontWallsAdiabatic", "LeftRightWallsConstT", "TopBottomPlatesAdiabatic"}   # B3S:71-85
FULL = ["f", "f_post", "g", "g_post", "rho", "u", "v", "w", "t", "up", "vp", "wp", "tp", "fx", "fy", "fz", "ex", "ey", "ez",
        "omega", "omegat", "un", "unt", "s", "m", "m_post", "meq", "fsource", "n", "n_post", "neq", "q"]


DEFS_RB = {"noslipWalls", "RBconvection", "BackFrontWallsAdiabatic", "LeftRightWallsAdiabatic", "TopBottomPlatesConstT"}   # B3S:74-79


def main(DEFS=DEFS, name="ref_fortran_thermal3d_seq_run.npz"):
    nx, ny, nz = 6, 5, 4
    text = fe.read_lines(B3S, 10, 28).replace("nx=51, ny=nx, nz=nx", f"nx={nx}, ny={ny}, nz={nz}") + "\n" + fe.read_lines(B3S, 54, 55)
    P = eval_parameters(text)
    assert (P["nx"], P["ny"], P["nz"]) == (nx, ny, nz)
    names_p = ("tauf", "viscosity", "diffusivity", "omegaratating", "paraa", "gbeta1", "gbeta", "snu", "sq", "qd", "qnu")
    out = {"params": np.array([P[k] for k in names_p]), "shape": np.array([nx, ny, nz])}
    sc = dict(nx=nx, ny=ny, nz=nz, itc=0, **{k: P[k] for k in names_p}, thot=P["thot"], tcold=P["tcold"], tref=P["tref"])
    tr = lambda a, b: fe.translate(strip_cpp(fe.read_lines(B3S, a, b), DEFS), full_arrays=FULL)
    src = {"weights": tr(291, 302), "initT": tr(316, 347), "initial": tr(350, 364), "collision": tr(415, 624), "streaming": tr(639, 652),
           "bounceback": tr(664, 723), "macro": tr(736, 750), "collisionT": tr(767, 809), "streamingT": tr(824, 837),
           "bouncebackT": tr(849, 919), "macroT": tr(931, 937), "check": tr(950, 976)}
    F4, H4, S3 = (0, 1, 1, 1), (0, 0, 0, 0), (1, 1, 1)
    field = lambda value: to_full(np.full((nx, ny, nz), value), S3)
    st = {k: fe._Arr() for k in ("omega", "omegat", "un", "unt", "s", "m", "m_post", "meq", "fsource", "n", "n_post", "neq", "q", "f", "g")}
    st.update(ex=arr(EX), ey=arr(EY), ez=arr(EZ), rho=field(1.0),                                        # B3S:288
              f_post=to_full(np.zeros((19, nx + 2, ny + 2, nz + 2)), H4), g_post=to_full(np.zeros((7, nx + 2, ny + 2, nz + 2)), H4),   # :390-391
              **{k: field(0.0) for k in ("u", "v", "w", "t", "up", "vp", "wp", "tp", "fx", "fy", "fz")})   # :306-309, :385-388
    names = list(st)

    def call(sub):
        ns = run_full(src[sub], st, sc)
        for k in names:
            st[k] = ns[k + "__"]
        return ns

    def step():
        for sub in ("collision", "streaming", "bounceback", "collisionT", "streamingT", "bouncebackT", "macro", "macroT"):   # B3S:117-133
            call(sub)

    def snap(tag):
        out[tag + "/f"] = from_full(st["f"], (19, nx, ny, nz), F4)
        out[tag + "/g"] = from_full(st["g"], (7, nx, ny, nz), F4)
        out[tag + "/ruvwT"] = np.stack([from_full(st[k], (nx, ny, nz), S3) for k in ("rho", "u", "v", "w", "t")])

    call("weights"); call("initT"); call("initial")
    snap("run0")
    done = 0
    for n in (1, 2, 10):
        for _ in range(n - done):
            step()
        done = n
        snap(f"run{n}")
        out[f"run{n}/F"] = np.stack([from_full(st[k], (nx, ny, nz), S3) for k in ("fx", "fy", "fz")])
    ns = call("check")
    out["run10/check"] = np.array([ns["erroru"], ns["errort"]])
    step(); step()
    ns = call("check")
    out["run12/check"] = np.array([ns["erroru"], ns["errort"]])
    snap("run12")
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
    main(DEFS_RB, "ref_fortran_thermal3d_seq_run_rb.npz")       # the program's other macro set (RB convection, B3S:74-79)
