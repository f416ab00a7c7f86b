"""Generates tests/golden/ref_fortran_thermal2d_seq_run.npz -- the reference's SEQUENTIAL 2-D buoyancy-driven cavity program
seq/steady.F90 (side-heated cell, its shipped macro set :7-31) run from its own source text (fortran_eval.py, whole arrays) on a
9 x 7 lattice:

  S2 = /root/reference/MPI/Buoyancy_driven_cavity/fortran/2d/seq/steady.F90
  parameters   S2:56-62, :79, :87, :96-110, :118-120 (nx, ny replaced by 9, 7; odd sizes keep the integer divisions of :87 exact)
  initial      S2:446-457 (weights), :466-481 (wall velocities, all 0 at shearReynolds = 0), :507-516 (T linear in x),
               :545-557 (populations); the whole-array assignments (:443-444, :461-463, :597-603) are applied by this script
  collision    S2:638-709      streaming   S2:726-737      bounceback   S2:754-897 (incl. the four corners :863-897)
  collisionT   S2:958-990      streamingT  S2:1007-1019    bouncebackT  S2:1036-1131
  macro        S2:927-939      macroT      S2:1146-1152    check        S2:1168-1193
its loop (S2:189-207: collision, streaming, bounceback, collisionT, streamingT, bouncebackT, macro, macroT) for 1, 2 and 20
iterations with check() after 20 and 25.  INTEGRATION.md maps this file onto MGLC_T2D_MPI with the side-heated set: the
restatement of the MPI program (oracle/thermal2d.c) must reproduce this run on 1 and on several emulated ranks.

A second output, tests/golden/ref_fortran_thermal2d_acc_run.npz, is the same whole run of the OpenACC program
  ACC = /root/reference/MPI/Buoyancy_driven_cavity/fortran/2d/seq/bouyancy2d_acc.F90
(Rayleigh-Benard plates, vertical walls periodic for f and g, lengthUnit = dble(nx), arrays indexed f(i,j,alpha): rewritten to
(alpha,i,j) before evaluation; `!$acc` lines dropped):
  parameters ACC:55-61, :90-103   initial ACC:419-430, :511-520 (T linear in y), :532-544
  collision ACC:624-707   streaming ACC:722-743   bounceback ACC:758-822   macro ACC:836-849
  collisionT ACC:869-919  streamingT ACC:934-953  bouncebackT ACC:968-1046 macroT ACC:1060-1071   check ACC:1087-1113
which variant "acc" of the oracle with the periodic Rayleigh-Benard set must reproduce (one rank, or ranks stacked along y).
Only numbers are stored; run in the authoring container."""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fortran_eval as fe  # noqa: E402
from make_golden_fortran import eval_parameters  # noqa: E402
from make_golden_thermal2d import arr, from_full, run_full, strip_cpp, to_full  # noqa: E402

S2 = "/root/reference/MPI/Buoyancy_driven_cavity/fortran/2d/seq/steady.F90"
DEFS = {"steadyFlow", "HorizontalWallsNoslip", "VerticalWallsNoslip", "SideHeatedCell", "HorizontalWallsAdiabatic", "VerticalWallsConstT"}
EX = [0, 1, 0, -1, 0, 1, -1, -1, 1]     # S2:132-133
EY = [0, 0, 1, 0, -1, 1, 1, -1, -1]
FULL = ["f", "f_post", "g", "g_post", "rho", "u", "v", "t", "up", "vp", "tp", "fx", "fy", "ex", "ey", "omega", "omegat", "un", "s", "m",
        "m_post", "meq", "fsource", "n", "n_post", "neq", "q", "obst"]


ACC = "/root/reference/MPI/Buoyancy_driven_cavity/fortran/2d/seq/bouyancy2d_acc.F90"
ACC_DEFS = {"steadyFlow", "HorizontalWallsNoslip", "VerticalWallsPeriodicalU", "RayleighBenardCell", "HorizontalWallsConstT", "VerticalWallsPeriodicalT"}
STEADY = dict(path=S2, defs=DEFS, name="ref_fortran_thermal2d_seq_run.npz", param_lines=((56, 62), (79, 79), (87, 87), (96, 110), (118, 120)),
              size_text="nx=201, ny=201", reorder=False,
              ranges={"weights": (446, 457), "initU": (466, 481), "initT": (507, 516), "initial": (545, 557), "collision": (638, 709),
                      "streaming": (726, 737), "bounceback": (754, 897), "macro": (927, 939), "collisionT": (958, 990),
                      "streamingT": (1007, 1019), "bouncebackT": (1036, 1131), "macroT": (1146, 1152), "check": (1168, 1193)})
# the same program with its other temperature macro set switched on (S2:22-24 instead of :29-31): Rayleigh-Benard plates
STEADY_RB = dict(STEADY, name="ref_fortran_thermal2d_seq_run_rb.npz",
                 defs={"steadyFlow", "HorizontalWallsNoslip", "VerticalWallsNoslip", "RayleighBenardCell", "HorizontalWallsConstT", "VerticalWallsAdiabatic"},
                 ranges=dict(STEADY["ranges"], initT=(507, 526)))
# the sheared Rayleigh-Benard program as shipped: the same text family with walls that MOVE (shearReynolds = 100, so U0 =
# 100*viscosity/ny enters initial() :466-481 and bounceback() :755-898), Pr = 5.3, and its own corner cells in bouncebackT()
# (:1086-1106: both populations of a corner cell take the plate's constant-temperature rule; the edge loops run 2..n-1)
RB2 = "/root/reference/MPI/Buoyancy_driven_cavity/fortran/2d/seq/R_B_2d.F90"
_shift = lambda r: {k: ((a, b) if b <= 744 else (a + 1, b + 1)) for k, (a, b) in r.items()}      # one more line after :744
SHEARED_RB = dict(STEADY_RB, path=RB2, name="ref_fortran_thermal2d_seq_run_sheared_rb.npz", moving_walls=True,
                  defs={"unsteadyFlow", "HorizontalWallsNoslip", "VerticalWallsNoslip", "RayleighBenardCell", "HorizontalWallsConstT", "VerticalWallsAdiabatic"},
                  ranges=_shift(STEADY_RB["ranges"]))
# the OpenMP program of the same family as shipped (same macro set, same moving walls and corner cells; its irregular-geometry,
# tilted-cell and tracer-particle branches are compiled out / never reached in the first 25 iterations)
OMP2 = "/root/reference/MPI/Buoyancy_driven_cavity/fortran/2d/seq/bouyancy2d_omp.F90"
SHEARED_OMP = dict(SHEARED_RB, path=OMP2, name="ref_fortran_thermal2d_seq_run_sheared_omp.npz",
                   param_lines=((65, 71), (88, 88), (96, 96), (111, 125), (133, 135)),
                   ranges={"weights": (598, 609), "initU": (662, 677), "initT": (741, 760), "initial": (779, 791), "collision": (881, 969),
                           "streaming": (986, 997), "bounceback": (1089, 1232), "macro": (1262, 1274), "collisionT": (1293, 1336),
                           "streamingT": (1353, 1365), "bouncebackT": (1440, 1535), "macroT": (1550, 1556), "check": (1572, 1597)})
ACCRUN = dict(path=ACC, defs=ACC_DEFS, name="ref_fortran_thermal2d_acc_run.npz", param_lines=((55, 61), (90, 103)),
              size_text="nx=513, ny=257", reorder=True,
              ranges={"weights": (419, 430), "initT": (511, 520), "initial": (532, 544), "collision": (624, 707), "streaming": (722, 743),
                      "bounceback": (758, 822), "macro": (836, 849), "collisionT": (869, 919), "streamingT": (934, 953),
                      "bouncebackT": (968, 1046), "macroT": (1060, 1071), "check": (1087, 1113)})


def main(cfg=STEADY):
    nx, ny = 9, 7
    S2, DEFS = cfg["path"], cfg["defs"]
    text = "\n".join(fe.read_lines(S2, a, b) for a, b in cfg["param_lines"])
    assert cfg["size_text"] in text
    text = text.replace(cfg["size_text"], f"nx={nx}, ny={ny}")
    P = eval_parameters(text)
    P.setdefault("rho0", 1.0)                  # ACC:416 `rho = 1.0d0`
    assert (P["nx"], P["ny"], P["lengthunit"]) == (nx, ny, float(nx if cfg["reorder"] else ny))
    assert (P.get("u0", 0.0) != 0.0) == bool(cfg.get("moving_walls"))
    names_p = ("tauf", "viscosity", "diffusivity", "paraa", "gbeta", "snu", "sq", "qd", "qnu")
    out = {"params": np.array([P[k] for k in names_p]), "shape": np.array([nx, ny])}
    if cfg.get("moving_walls"):                 # S:118-120, in the order of mglc_t2d_desc::Uwall
        out["uwall"] = np.array([P["uwall" + k] for k in ("topleft", "topright", "bottomleft", "bottomright", "lefttop", "leftbottom", "righttop", "rightbottom")])
        out["prandtl"] = np.array(P["prandtl"])
    sc = {k: v for k, v in P.items()}
    sc.update(itc=0)
    def tr(a, b):
        t = strip_cpp(fe.read_lines(S2, a, b), DEFS)
        t = "\n".join(l for l in t.splitlines() if not l.strip().lower().startswith("!$acc"))
        if cfg["reorder"]:                      # f(i,j,alpha) -> f(alpha,i,j)
            t = re.sub(r"\b(f_post|g_post|f|g)\(([^(),]+),([^(),]+),([^(),]+)\)", r"\1(\4,\2,\3)", t)
        return fe.translate(t, full_arrays=FULL)
    src = {k: tr(*r) for k, r in cfg["ranges"].items()}
    F3, H3, S = (0, 1, 1), (0, 0, 0), (1, 1)
    field = lambda value: to_full(np.full((nx, ny), value), S)
    st = {k: fe._Arr() for k in ("omega", "omegat", "un", "s", "m", "m_post", "meq", "fsource", "n", "n_post", "neq", "q", "f", "g")}
    st.update(ex=arr(EX), ey=arr(EY), rho=field(P["rho0"]), obst=fe._Arr({(i, j): 0 for i in range(nx + 2) for j in range(ny + 2)}),   # S2:443-444
              f_post=to_full(np.zeros((9, nx + 2, ny + 2)), H3), g_post=to_full(np.zeros((5, nx + 2, ny + 2)), H3),                  # :602-603
              **{k: field(0.0) for k in ("u", "v", "t", "up", "vp", "tp", "fx", "fy")})                                             # :461-463, :597-600
    names = list(st)

    def call(sub):
        ns = run_full(src[sub], st, sc)
        for k in names:
            st[k] = ns[k + "__"]
        return ns

    def step():
        for sub in ("collision", "streaming", "bounceback", "collisionT", "streamingT", "bouncebackT", "macro", "macroT"):   # S2:189-207
            call(sub)

    def snap(tag):
        out[tag + "/f"] = from_full(st["f"], (9, nx, ny), F3)
        out[tag + "/g"] = from_full(st["g"], (5, nx, ny), F3)
        out[tag + "/ruvT"] = np.stack([from_full(st[k], (nx, ny), S) for k in ("rho", "u", "v", "t")])
        out[tag + "/F"] = np.stack([from_full(st[k], (nx, ny), S) for k in ("fx", "fy")])

    for sub in ("weights", "initU", "initT", "initial"):
        if sub in src:
            call(sub)
    snap("run0")
    done = 0
    for n in (1, 2, 20):
        for _ in range(n - done):
            step()
        done = n
        snap(f"run{n}")
    ns = call("check")
    out["run20/check"] = np.array([ns["erroru"], ns["errort"]])
    for _ in range(5):
        step()
    ns = call("check")
    out["run25/check"] = np.array([ns["erroru"], ns["errort"]])
    snap("run25")
    path = os.path.join(HERE, cfg["name"])
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    only = sys.argv[1:]
    for cfg in (STEADY, STEADY_RB, SHEARED_RB, SHEARED_OMP, ACCRUN):
        if not only or cfg["name"] in only:
            main(cfg)
