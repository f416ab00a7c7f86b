"""Generates tests/golden/ref_fortran_kernels.npz: input/output vectors of the REFERENCE's per-cell
arithmetic, obtained by machine-evaluating the reference's own Fortran source text (fortran_eval.py) on
seeded random inputs.  Run in the authoring container (reads /root/reference; nothing of it is stored here
except numbers).  The oracle (oracle/*.c) and the CUDA kernels are tested against these vectors bit for bit
(strict arithmetic) -- this is what pins the restatement to the reference for the Fortran paths, which no
compiler in this image can build.

Sources evaluated (file:lines):
  L3 = MPI/Lid_driven_cavity/fortran/3d/mpi_3d_blocked          B3 = MPI/Buoyancy_driven_cavity/fortran/3d/bouyancy3d_mpi.F90
  lid_collision   L3/collision.f90:20-189        lid_macro   L3/macro.f90:13-22      lid_feq   L3/initial.f90:66-70
  th_params       B3:26-47,73-74                 th_collision B3:656-856             th_macro  B3:995-1002
  th_collisionT   B3:1028-1062                   th_feq      B3:589-597 (+ omega tables :496-507)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fortran_eval as fe  # noqa: E402

REF = "/root/reference/MPI"
L3 = REF + "/Lid_driven_cavity/fortran/3d/mpi_3d_blocked"
B3 = REF + "/Buoyancy_driven_cavity/fortran/3d/bouyancy3d_mpi.F90"

EX = [0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0]
EY = [0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1]
EZ = [0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1]
W19 = [1 / 3] + [1 / 18] * 6 + [1 / 36] * 12


def eval_parameters(text, known=None):
    """`real(kind=8), parameter :: a=expr, b=expr` lines -> {name: value}, evaluated in order"""
    import math
    ns = {"sqrt": math.sqrt, "float": float, "int": int}
    ns.update(known or {})
    for line in fe._logical_lines(text.lower()):
        if "parameter" not in line or "::" not in line:
            continue
        rhs = fe._numbers(line.split("::", 1)[1])
        rhs = rhs.replace("dsqrt", "sqrt").replace("dble", "float")
        for item in fe._split_args(rhs):
            name, expr = item.split("=", 1)
            ns[name.strip()] = fe.safe_eval(expr, ns)
    return {k: v for k, v in ns.items() if isinstance(v, (int, float))}


def arr(seq):
    return fe._Arr({k: x for k, x in enumerate(seq)})


def random_cells(rng, n, thermal=False):
    """populations near a moving equilibrium plus noise, so that every moment is exercised"""
    cells = []
    for _ in range(n):
        rho = 1.0 + 0.05 * rng.uniform(-1, 1)
        u, v, w = (0.08 * rng.uniform(-1, 1) for _ in range(3))
        f = []
        for a in range(19):
            un = u * EX[a] + v * EY[a] + w * EZ[a]
            f.append(rho * W19[a] * (1 + 3 * un + 4.5 * un * un - 1.5 * (u * u + v * v + w * w)) * (1 + 0.02 * rng.uniform(-1, 1)))
        c = dict(f=f, rho=1.0 + 0.05 * rng.uniform(-1, 1), u=0.08 * rng.uniform(-1, 1), v=0.08 * rng.uniform(-1, 1),
                 w=0.08 * rng.uniform(-1, 1))
        if thermal:
            c["T"] = rng.uniform(0, 1)
            c["g"] = [c["T"] / 7 * (1 + 0.3 * rng.uniform(-1, 1)) for _ in range(7)]
            c["Fx"], c["Fy"], c["Fz"] = (1e-4 * rng.uniform(-1, 1) for _ in range(3))
        cells.append(c)
    return cells


def main():
    rng = np.random.default_rng(20211)
    out = {}

    # ---------------- lid-driven cavity (L3) ----------------
    tau = 0.1 * 65 / 1000.0 * 3.0 + 0.5                                  # config 1 (commondata.f90:9)
    snu, sq = 1.0 / tau, 8.0 * (2.0 * tau - 1.0) / (8.0 * tau - 1.0)     # commondata.f90:42
    src = fe.translate(fe.read_lines(L3 + "/collision.f90", 20, 189), cell_arrays=["f", "f_post"],
                       fields=["rho", "u", "v", "w"], local_arrays=["m", "m_post", "meq", "s"])
    cells = random_cells(rng, 32)
    res = [fe.run(src, cell_in={"f": c["f"]}, field_in={k: c[k] for k in ("rho", "u", "v", "w")},
                  scalars={"snu": snu, "sq": sq}, local_arrays=["m", "m_post", "meq", "s"], cell_out=["f_post"]) for c in cells]
    out["lid_collision/f"] = np.array([c["f"] for c in cells])
    out["lid_collision/ruvw"] = np.array([[c[k] for k in ("rho", "u", "v", "w")] for c in cells])
    out["lid_collision/snu_sq"] = np.array([snu, sq])
    out["lid_collision/f_post"] = np.array([r["f_post"] for r in res])

    src = fe.translate(fe.read_lines(L3 + "/macro.f90", 13, 22), cell_arrays=["f"], fields=["rho", "u", "v", "w"],
                       local_arrays=["ex", "ey", "ez"])
    res = [fe.run(src, cell_in={"f": c["f"]}, field_in={"rho": 0.0, "u": 0.0, "v": 0.0, "w": 0.0},     # macro.f90:6-9 zero-fill
                  scalars={"ex__": arr(EX), "ey__": arr(EY), "ez__": arr(EZ)}, field_out=["rho", "u", "v", "w"]) for c in cells]
    out["lid_macro/f"] = out["lid_collision/f"]
    out["lid_macro/ruvw"] = np.array([[r[k] for k in ("rho", "u", "v", "w")] for r in res])

    src = fe.translate(fe.read_lines(L3 + "/initial.f90", 66, 70), cell_arrays=["f"], fields=["rho", "u", "v", "w"],
                       local_arrays=["ex", "ey", "ez", "omega", "un"])
    om = arr([1.0 / 3.0] + [1.0 / 18.0] * 6 + [1.0 / 36.0] * 12)          # initial.f90:24-30
    res = [fe.run(src, field_in={k: c[k] for k in ("rho", "u", "v", "w")},
                  scalars={"ex__": arr(EX), "ey__": arr(EY), "ez__": arr(EZ), "omega__": om}, local_arrays=["un"], cell_out=["f"])
           for c in cells]
    out["lid_feq/ruvw"] = out["lid_collision/ruvw"]
    out["lid_feq/f"] = np.array([r["f"] for r in res])

    # ---------------- thermal (B3) ----------------
    P = eval_parameters(fe.read_lines(B3, 26, 47) + "\n" + fe.read_lines(B3, 73, 74))
    names = ["total_nx", "rayleigh", "prandtl", "mach", "tauf", "thot", "tcold", "tref", "viscosity", "diffusivity", "ekman",
             "omegaratating", "paraa", "gbeta1", "gbeta", "snu", "sq", "qd", "qnu"]
    out["th_params/names"] = np.array(names)
    out["th_params/values"] = np.array([float(P[n]) for n in names])

    cells = random_cells(rng, 32, thermal=True)
    src = fe.translate(fe.read_lines(B3, 656, 856), cell_arrays=["f", "f_post"],
                       fields=["rho", "u", "v", "w", "t", "fx", "fy", "fz"], local_arrays=["m", "m_post", "meq", "s", "fsource"])
    res = [fe.run(src, cell_in={"f": c["f"]}, field_in={"rho": c["rho"], "u": c["u"], "v": c["v"], "w": c["w"], "t": c["T"]},
                  scalars={k: P[k] for k in ("snu", "sq", "omegaratating", "gbeta", "tref")},
                  local_arrays=["m", "m_post", "meq", "s", "fsource"], cell_out=["f_post"], field_out=["fx", "fy", "fz"]) for c in cells]
    out["th_collision/f"] = np.array([c["f"] for c in cells])
    out["th_collision/ruvwT"] = np.array([[c[k] for k in ("rho", "u", "v", "w", "T")] for c in cells])
    out["th_collision/f_post"] = np.array([r["f_post"] for r in res])
    out["th_collision/F"] = np.array([[r["fx"], r["fy"], r["fz"]] for r in res])

    src = fe.translate(fe.read_lines(B3, 995, 1002), cell_arrays=["f"], fields=["rho", "u", "v", "w", "fx", "fy", "fz"])
    res = [fe.run(src, cell_in={"f": c["f"]}, field_in={"fx": c["Fx"], "fy": c["Fy"], "fz": c["Fz"]},
                  field_out=["rho", "u", "v", "w"]) for c in cells]
    out["th_macro/f"] = out["th_collision/f"]
    out["th_macro/F"] = np.array([[c["Fx"], c["Fy"], c["Fz"]] for c in cells])
    out["th_macro/ruvw"] = np.array([[r[k] for k in ("rho", "u", "v", "w")] for r in res])

    src = fe.translate(fe.read_lines(B3, 1028, 1062), cell_arrays=["g", "g_post"], fields=["u", "v", "w", "t"],
                       local_arrays=["n", "n_post", "neq", "q"])
    res = [fe.run(src, cell_in={"g": c["g"]}, field_in={"u": c["u"], "v": c["v"], "w": c["w"], "t": c["T"]},
                  scalars={k: P[k] for k in ("paraa", "qd", "qnu")}, local_arrays=["n", "n_post", "neq", "q"], cell_out=["g_post"])
           for c in cells]
    out["th_collisionT/g"] = np.array([c["g"] for c in cells])
    out["th_collisionT/uvwT"] = np.array([[c[k] for k in ("u", "v", "w", "T")] for c in cells])
    out["th_collisionT/g_post"] = np.array([r["g_post"] for r in res])

    # initial equilibria incl. the weight tables (B3:496-507 evaluated from the text too)
    wsrc = fe.translate(fe.read_lines(B3, 496, 507), local_arrays=["omega", "omegat"])
    wres = fe.run(wsrc, scalars={"paraa": P["paraa"]}, cell_out=["omega", "omegat"])
    src = fe.translate(fe.read_lines(B3, 589, 597), cell_arrays=["f", "g"], fields=["rho", "u", "v", "w", "t"],
                       local_arrays=["ex", "ey", "ez", "omega", "omegat", "un", "unt"])
    res = [fe.run(src, field_in={"rho": c["rho"], "u": c["u"], "v": c["v"], "w": c["w"], "t": c["T"]},
                  scalars={"ex__": arr(EX), "ey__": arr(EY), "ez__": arr(EZ), "omega__": arr(wres["omega"]),
                           "omegat__": arr(wres["omegat"]), "paraa": P["paraa"]},
                  local_arrays=["un", "unt"], cell_out=["f", "g"]) for c in cells]
    out["th_feq/ruvwT"] = out["th_collision/ruvwT"]
    out["th_feq/f"] = np.array([r["f"] for r in res])
    out["th_feq/g"] = np.array([r["g"] for r in res])

    np.savez_compressed(os.path.join(HERE, "ref_fortran_kernels.npz"), **out)
    print("wrote", len(out), "arrays;", "tauf =", P["tauf"], "paraA =", P["paraa"], "gBeta =", P["gbeta"])


if __name__ == "__main__":
    main()
