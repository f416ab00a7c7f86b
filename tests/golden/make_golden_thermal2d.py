"""Generates tests/golden/ref_fortran_thermal2d.npz -- golden vectors for the 2-D thermal (D2Q9 + D2Q5) path, from the
REFERENCE's own source text:

  B2 = /root/reference/MPI/Buoyancy_driven_cavity/fortran/2d/mpi_blocked/

Fortran + MPI cannot be built in this image, so the hot-path subroutines are machine-evaluated from the files where they
lie (fortran_eval.py: verbatim loop bodies -> Python floats = IEEE binary64, left to right, no contraction):
  params/*        module.F90:69-81 (tauf, viscosity, diffusivity, paraA, gBeta, Snu, Sq, Qd, Qnu)
  collision/*     evolution_f.F90:15-78 on seeded cells  -> f_post, Fx, Fy
  collisionT/*    evolution_g.F90:13-39                   -> g_post
  macro/*         evolution_f.F90:335-337, evolution_g.F90:171
  initial/*       initial.F90:201-212 (weights), :258/:268 (T profile), :278-286 (populations)
  field/*         whole-array subroutines on a seeded 6 x 5 block with one-cell halos: streaming :96-105, bounceback :283-321,
                  streamingT g:56-65, bouncebackT g:79-142 (both macro sets of macros.F90), check.F90:10-30 (the four rank sums)
  acc/*           the same subroutines of the OpenACC program seq/bouyancy2d_acc.F90 (the reference's only GPU code; population index
                  last): module constants :91-103, collision :631-695 (f_post(0) rounded term by term), collisionT :879-905, and
                  streaming :722-743, bounceback :758-822, streamingT :934-954, bouncebackT :968-1046 with its shipped macro
                  set (periodic vertical walls for f and g, constant-temperature plates)
Only numbers are stored; run in the authoring container."""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fortran_eval as fe  # noqa: E402

B2 = "/root/reference/MPI/Buoyancy_driven_cavity/fortran/2d/mpi_blocked"
EX = [0, 1, 0, -1, 0, 1, -1, -1, 1]
EY = [0, 0, 1, 0, -1, 1, 1, -1, -1]
W9 = [4 / 9] + [1 / 9] * 4 + [1 / 36] * 4


def arr(seq, lo=0):
    return fe._Arr({k + lo: x for k, x in enumerate(seq)})


def params(total_ny, Rayleigh=1e7, Prandtl=0.71, Mach=0.1):
    """module.F90:69-81: `real(kind=8), parameter :: a=..., b=...` lines turned into assignments and evaluated in order"""
    text = []
    for line in fe.read_lines(B2 + "/module.F90", 69, 81).splitlines():
        line = line.split("!")[0].strip()
        m = re.match(r"real\(kind=8\),\s*parameter\s*::\s*(.*)$", line)
        if not m:
            continue
        text += fe._split_args(m.group(1))
    src = fe.translate("\n".join(text))
    ns = fe.run(src + "\nout__ = dict(tauf=tauf, viscosity=viscosity, diffusivity=diffusivity, paraa=paraa, gbeta=gbeta, snu=snu, sq=sq, qd=qd, qnu=qnu)",
                scalars=dict(mach=Mach, lengthunit=float(total_ny), prandtl=Prandtl, rayleigh=Rayleigh), field_out=["out"])
    return ns["out"]


def strip_cpp(text, defined):
    """resolve the #ifdef / #endif blocks of a reference excerpt for the given macro set"""
    out, keep = [], [True]
    for line in text.splitlines():
        s = line.strip()
        if s.startswith("#ifdef"):
            keep.append(keep[-1] and s.split()[1] in defined)
        elif s.startswith("#ifndef"):
            keep.append(keep[-1] and s.split()[1] not in defined)
        elif s.startswith("#endif"):
            keep.pop()
        elif keep[-1]:
            out.append(line)
    return "\n".join(out)


def field_arrays(rng, nx, ny):
    """f_post / g_post with halos, f / g interior, all seeded; keyed the way the `full` rewrite indexes them"""
    fp = rng.random((9, nx + 2, ny + 2))
    gp = rng.random((5, nx + 2, ny + 2))
    f0 = rng.random((9, nx, ny))
    g0 = rng.random((5, nx, ny))
    return fp, gp, f0, g0


def to_full(a, lo):
    """numpy array -> _Arr keyed by Fortran index tuples with lower bounds `lo`"""
    d = fe._Arr()
    for idx in np.ndindex(a.shape):
        d[tuple(int(i + l) for i, l in zip(idx, lo))] = float(a[idx])
    return d


def from_full(d, shape, lo):
    out = np.empty(shape)
    for idx in np.ndindex(shape):
        out[idx] = d[tuple(int(i + l) for i, l in zip(idx, lo))]
    return out


def run_full(src, arrays, scalars):
    ns = {"sqrt": np.sqrt, "abs": abs, "float": float, "int": int, "range": range}
    ns.update(scalars)
    for k, v in arrays.items():
        ns[k + "__"] = v
    fe.safe_exec(src, ns, "<reference subroutine>")
    return ns


def main():
    rng = np.random.default_rng(20272)
    out = {}
    total = 201
    P = params(total)
    out["params/201"] = np.array([P[k] for k in ("tauf", "viscosity", "diffusivity", "paraa", "gbeta", "snu", "sq", "qd", "qnu")])
    P64 = params(64, Rayleigh=1e6, Prandtl=0.71, Mach=0.1)
    out["params/64_ra1e6"] = np.array([P64[k] for k in ("tauf", "viscosity", "diffusivity", "paraa", "gbeta", "snu", "sq", "qd", "qnu")])

    # ---------------- per-cell arithmetic ----------------
    ncell = 48
    cells = []
    for _ in range(ncell):
        rho, u, v, T = 1.0 + 0.05 * rng.uniform(-1, 1), 0.08 * rng.uniform(-1, 1), 0.08 * rng.uniform(-1, 1), rng.uniform(-0.2, 1.2)
        f = [rho * W9[a] * (1 + 3 * (u * EX[a] + v * EY[a]) + 4.5 * (u * EX[a] + v * EY[a]) ** 2 - 1.5 * (u * u + v * v)) *
             (1 + 0.02 * rng.uniform(-1, 1)) for a in range(9)]
        g = [T * 0.2 * (1 + 0.3 * rng.uniform(-1, 1)) for _ in range(5)]
        cells.append(dict(f=f, g=g, rho=1.0 + 0.05 * rng.uniform(-1, 1), u=0.08 * rng.uniform(-1, 1), v=0.08 * rng.uniform(-1, 1),
                          T=rng.uniform(-0.2, 1.2), Fx=0.0, Fy=1e-4 * rng.uniform(-1, 1)))
    # a few exact-zero / signed-zero cases: the `+0.5d0*Fx` and `u*Fx` terms decide the sign of zero results
    cells[0].update(u=0.0, v=0.0, T=0.0)
    cells[1].update(u=-0.0, v=0.0, T=0.5)
    cells[2]["f"] = [W9[a] for a in range(9)]
    cells[2].update(rho=1.0, u=0.0, v=0.0, T=0.0, Fy=0.0)
    out["cells/f"] = np.array([c["f"] for c in cells])
    out["cells/g"] = np.array([c["g"] for c in cells])
    out["cells/ruvT"] = np.array([[c[k] for k in ("rho", "u", "v", "T")] for c in cells])
    out["cells/FxFy"] = np.array([[c["Fx"], c["Fy"]] for c in cells])
    sc = dict(snu=P["snu"], sq=P["sq"], gbeta=P["gbeta"], tref=0.0, paraa=P["paraa"], qd=P["qd"], qnu=P["qnu"])

    la = ["s", "m", "m_post", "meq", "fsource"]
    src = fe.translate(fe.read_lines(B2 + "/evolution_f.F90", 15, 78), cell_arrays=["f", "f_post"], fields=["rho", "u", "v", "T", "Fx", "Fy"], local_arrays=la)
    res = [fe.run(src, cell_in={"f": c["f"]}, field_in={k: c[k] for k in ("rho", "u", "v", "T")}, scalars=sc, local_arrays=la,
                  cell_out=["f_post"], field_out=["fx", "fy"]) for c in cells]
    out["collision/f_post"] = np.array([r["f_post"] for r in res])
    out["collision/FxFy"] = np.array([[r["fx"], r["fy"]] for r in res])

    la = ["n", "n_post", "neq", "q"]
    src = fe.translate(fe.read_lines(B2 + "/evolution_g.F90", 13, 39), cell_arrays=["g", "g_post"], fields=["T", "u", "v"], local_arrays=la)
    res = [fe.run(src, cell_in={"g": c["g"]}, field_in={k: c[k] for k in ("u", "v", "T")}, scalars=sc, local_arrays=la, cell_out=["g_post"])
           for c in cells]
    out["collisionT/g_post"] = np.array([r["g_post"] for r in res])

    src = fe.translate(fe.read_lines(B2 + "/evolution_f.F90", 335, 337), cell_arrays=["f"], fields=["rho", "u", "v", "Fx", "Fy"])
    res = [fe.run(src, cell_in={"f": c["f"]}, field_in={"fx": c["Fx"], "fy": c["Fy"]}, field_out=["rho", "u", "v"]) for c in cells]
    out["macro/ruv"] = np.array([[r[k] for k in ("rho", "u", "v")] for r in res])
    src = fe.translate(fe.read_lines(B2 + "/evolution_g.F90", 171, 171), cell_arrays=["g"], fields=["T"])
    out["macro/T"] = np.array([fe.run(src, cell_in={"g": c["g"]}, field_out=["T"])["T"] for c in cells])

    # ---------------- initial() ----------------
    src_w = fe.translate(fe.read_lines(B2 + "/initial.F90", 201, 212), local_arrays=["omega", "omegat"])
    ns = run_full(src_w, {"omega": fe._Arr(), "omegat": fe._Arr()}, dict(paraa=P["paraa"]))
    om, omT = ns["omega__"], ns["omegat__"]
    out["initial/omega"] = np.array([om[a] for a in range(9)])
    out["initial/omegaT"] = np.array([omT[a] for a in range(5)])
    src_T = fe.translate(fe.read_lines(B2 + "/initial.F90", 258, 258), fields=["T"])
    out["initial/T_profile_201"] = np.array([fe.run(src_T, scalars=dict(i_start_global=s, i=i, total_nx=201, tcold=0.0, thot=1.0), field_out=["T"])["T"]
                                             for s, i in [(0, 1), (0, 2), (0, 101), (101, 1), (101, 100), (150, 51), (67, 33)]])
    out["initial/T_profile_args"] = np.array([(0, 1), (0, 2), (0, 101), (101, 1), (101, 100), (150, 51), (67, 33)])
    src_T = fe.translate(fe.read_lines(B2 + "/initial.F90", 268, 268), fields=["T"])
    out["initial/T_profile_y_77"] = np.array([fe.run(src_T, scalars=dict(j_start_global=s, j=j, total_ny=77, tcold=-0.5, thot=0.5), field_out=["T"])["T"]
                                              for s, j in [(0, 1), (0, 39), (39, 38), (20, 7)]])
    src = fe.translate(fe.read_lines(B2 + "/initial.F90", 278, 286), cell_arrays=["f", "g"], fields=["rho", "u", "v", "T"],
                       local_arrays=["ex", "ey", "omega", "omegat", "un"])
    res = [fe.run(src, field_in={k: c[k] for k in ("rho", "u", "v", "T")},
                  scalars={"ex__": arr(EX), "ey__": arr(EY), "omega__": om, "omegat__": omT, "paraa": P["paraa"]},
                  local_arrays=["un"], cell_out=["f", "g"]) for c in cells]
    out["initial/feq"] = np.array([r["f"] for r in res])
    out["initial/geq"] = np.array([r["g"] for r in res])

    # ---------------- whole-array subroutines on a small block ----------------
    nx, ny = 6, 5
    fp, gp, f0, g0 = field_arrays(rng, nx, ny)
    out["field/f_post"], out["field/g_post"], out["field/f0"], out["field/g0"] = fp, gp, f0, g0
    full = ["f", "f_post", "g", "g_post", "ex", "ey", "coords", "dims"]
    common = dict(nx=nx, ny=ny, paraa=P["paraa"], thot=1.0, tcold=0.0)
    idx = {"ex": arr(EX), "ey": arr(EY)}

    def stream(path, first, last, name, q, post):
        src = fe.translate(fe.read_lines(path, first, last), full_arrays=full)
        A = {name: fe._Arr(), name + "_post": to_full(post, (0, 0, 0)), **idx}
        ns = run_full(src, A, common)
        return from_full(ns[name + "__"], (q, nx, ny), (0, 1, 1))
    out["field/streaming_f"] = stream(B2 + "/evolution_f.F90", 96, 105, "f", 9, fp)
    out["field/streamingT_g"] = stream(B2 + "/evolution_g.F90", 56, 65, "g", 5, gp)

    # bounceback / bouncebackT for every position of the block in a 3 x 3 process grid collapsed to the cases that matter:
    # (coords, dims) = single rank, and the four corner ranks of a 2 x 2 grid
    cases = [((0, 0), (1, 1)), ((0, 0), (2, 2)), ((1, 0), (2, 2)), ((0, 1), (2, 2)), ((1, 1), (2, 2))]
    out["field/bb_cases"] = np.array([c + d for c, d in cases])
    bb_text = fe.read_lines(B2 + "/evolution_f.F90", 283, 321)
    bbT_text = fe.read_lines(B2 + "/evolution_g.F90", 79, 142)
    macro_sets = {"side": {"VerticalWallsNoslip", "HorizontalWallsNoslip", "HorizontalWallsAdiabatic", "VerticalWallsConstT"},
                  "rb": {"VerticalWallsNoslip", "HorizontalWallsNoslip", "HorizontalWallsConstT", "VerticalWallsAdiabatic"}}
    for k, (co, di) in enumerate(cases):
        src = fe.translate(strip_cpp(bb_text, macro_sets["side"]), full_arrays=full)
        A = {"f": to_full(f0, (0, 1, 1)), "f_post": to_full(fp, (0, 0, 0)), "coords": arr(co), "dims": arr(di)}
        ns = run_full(src, A, common)
        out[f"field/bounceback_{k}"] = from_full(ns["f__"], (9, nx, ny), (0, 1, 1))
        for tag, defs in macro_sets.items():
            src = fe.translate(strip_cpp(bbT_text, defs), full_arrays=full)
            A = {"g": to_full(g0, (0, 1, 1)), "g_post": to_full(gp, (0, 0, 0)), "coords": arr(co), "dims": arr(di)}
            # paraA of the world the test builds around this block: total_ny = ny * dims(1)   (module.F90:29,69-73)
            ns = run_full(src, A, dict(common, paraa=params(ny * di[1])["paraa"]))
            out[f"field/bouncebackT_{tag}_{k}"] = from_full(ns["g__"], (5, nx, ny), (0, 1, 1))

    # check(): the four rank sums and the up/vp/Tp update
    fl = {k: rng.uniform(-0.1, 0.1, (nx, ny)) for k in ("u", "v", "up", "vp")}
    fl["T"], fl["Tp"] = rng.uniform(-0.2, 1.2, (nx, ny)), rng.uniform(-0.2, 1.2, (nx, ny))
    for k, a in fl.items():
        out[f"field/check_{k}"] = a
    src = fe.translate(fe.read_lines(B2 + "/check.F90", 10, 30), full_arrays=["u", "v", "t", "up", "vp", "tp"])
    ns = run_full(src, {k.lower(): to_full(a, (1, 1)) for k, a in fl.items()}, dict(nx=nx, ny=ny))
    out["field/check_sums"] = np.array([ns["error1"], ns["error2"], ns["error5"], ns["error6"]])
    out["field/check_up_after"] = from_full(ns["up__"], (nx, ny), (1, 1))

    # ---------------- the OpenACC program (seq/bouyancy2d_acc.F90): the reference's only GPU code ----------------
    # same subroutines with the population index LAST, f(i,j,alpha); reorder the subscripts, then evaluate as above
    ACC = "/root/reference/MPI/Buoyancy_driven_cavity/fortran/2d/seq/bouyancy2d_acc.F90"

    def acc_text(first, last, defined=()):
        t = strip_cpp(fe.read_lines(ACC, first, last), set(defined))
        t = "\n".join(l for l in t.splitlines() if not l.strip().lower().startswith("!$acc"))
        return re.sub(r"\b(f_post|g_post|f|g)\(([^(),]+),([^(),]+),([^(),]+)\)", r"\1(\4,\2,\3)", t)
    Pa = acc_params = None
    # module constants of the OpenACC program: nx = 513, ny = 257, lengthUnit = dble(nx), Ra = 1e5 (acc:55-60,91-103)
    text = []
    for line in fe.read_lines(ACC, 91, 103).splitlines():
        m = re.match(r"\s*real\(kind=8\),\s*parameter\s*::\s*(.*)$", line.split("!")[0].strip())
        if m:
            text += fe._split_args(m.group(1))
    ns = fe.run(fe.translate("\n".join(text)) + "\nout__ = dict(tauf=tauf, viscosity=viscosity, diffusivity=diffusivity, paraa=paraa, gbeta=gbeta, snu=snu, sq=sq, qd=qd, qnu=qnu)",
                scalars=dict(mach=0.1, lengthunit=513.0, prandtl=0.71, rayleigh=1e5), field_out=["out"])
    Pa = ns["out"]
    out["acc/params"] = np.array([Pa[k] for k in ("tauf", "viscosity", "diffusivity", "paraa", "gbeta", "snu", "sq", "qd", "qnu")])
    sca = dict(snu=Pa["snu"], sq=Pa["sq"], gbeta=Pa["gbeta"], tref=0.0, paraa=Pa["paraa"], qd=Pa["qd"], qnu=Pa["qnu"])
    la = ["s", "m", "m_post", "meq", "fsource"]
    src = fe.translate(acc_text(631, 695), cell_arrays=["f", "f_post"], fields=["rho", "u", "v", "T", "Fx", "Fy"], local_arrays=la)
    res = [fe.run(src, cell_in={"f": c["f"]}, field_in={k: c[k] for k in ("rho", "u", "v", "T")}, scalars=sca, local_arrays=la,
                  cell_out=["f_post"], field_out=["fx", "fy"]) for c in cells]
    out["acc/collision_f_post"] = np.array([r["f_post"] for r in res])
    out["acc/collision_FxFy"] = np.array([[r["fx"], r["fy"]] for r in res])
    la = ["n", "n_post", "neq", "q"]
    src = fe.translate(acc_text(879, 905), cell_arrays=["g", "g_post"], fields=["T", "u", "v"], local_arrays=la)
    res = [fe.run(src, cell_in={"g": c["g"]}, field_in={k: c[k] for k in ("u", "v", "T")}, scalars=sca, local_arrays=la, cell_out=["g_post"])
           for c in cells]
    out["acc/collisionT_g_post"] = np.array([r["g_post"] for r in res])
    # whole-array: streaming, bounceback, streamingT, bouncebackT with the shipped macro set (periodic vertical walls, RB plates)
    acc_defs = {"HorizontalWallsNoslip", "VerticalWallsPeriodicalU", "RayleighBenardCell", "HorizontalWallsConstT", "VerticalWallsPeriodicalT",
                "steadyFlow"}
    full_acc = ["f", "f_post", "g", "g_post", "ex", "ey", "obst"]
    obst = fe._Arr({(i, j): 0 for i in range(0, nx + 2) for j in range(0, ny + 2)})
    cm = dict(nx=nx, ny=ny, paraa=Pa["paraa"], thot=1.0, tcold=0.0)
    src = fe.translate(acc_text(722, 743, acc_defs), full_arrays=full_acc)
    ns = run_full(src, {"f": fe._Arr(), "f_post": to_full(fp, (0, 0, 0)), "obst": obst}, cm)
    out["acc/streaming_f"] = from_full(ns["f__"], (9, nx, ny), (0, 1, 1))
    src = fe.translate(acc_text(758, 822, acc_defs), full_arrays=full_acc)
    ns = run_full(src, {"f": to_full(f0, (0, 1, 1)), "f_post": to_full(fp, (0, 0, 0)), "obst": obst, **idx}, cm)
    out["acc/bounceback_f"] = from_full(ns["f__"], (9, nx, ny), (0, 1, 1))
    src = fe.translate(acc_text(934, 954, acc_defs), full_arrays=full_acc)
    ns = run_full(src, {"g": fe._Arr(), "g_post": to_full(gp, (0, 0, 0)), "obst": obst}, cm)
    out["acc/streamingT_g"] = from_full(ns["g__"], (5, nx, ny), (0, 1, 1))
    src = fe.translate(acc_text(968, 1046, acc_defs), full_arrays=full_acc)
    ns = run_full(src, {"g": to_full(g0, (0, 1, 1)), "g_post": to_full(gp, (0, 0, 0)), "obst": obst, **idx}, cm)
    out["acc/bouncebackT_g"] = from_full(ns["g__"], (5, nx, ny), (0, 1, 1))

    path = os.path.join(HERE, "ref_fortran_thermal2d.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
