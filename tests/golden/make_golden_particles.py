"""Generates tests/golden/ref_fortran_particles.npz: vectors of the particle-laden D2Q9 path obtained by
machine-evaluating the REFERENCE's own Fortran text (fortran_eval.py) on seeded inputs.  Run in the authoring
container (reads /root/reference).  P4 = MPI/Micro_particles/fortran/case4/mpi_particle.

  params      P4/commondata.F90:3-60          collision   P4/fluid.F90:14-64        macro     P4/fluid.F90:174-176
  feq         P4/initial.F90:147-151          calQ        P4/particle_bounceback.F90:112-139
  bb          P4/particle_bounceback.F90:62-75 (both interpolation branches)
  force_link  P4/particle_force.F90:56-62     forces      P4/particle_force.F90:96-190 (springs, walls, weight)
  advance     P4/particle_update.F90:37-50    refill      P4/particle_update.F90:137-192
  dims        P4/mpi_starts.F90:166-178
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fortran_eval as fe  # noqa: E402

P4 = "/root/reference/MPI/Micro_particles/fortran/case4/mpi_particle"
EX = [0, 1, 0, -1, 0, 1, -1, -1, 1]
EY = [0, 0, 1, 0, -1, 1, 1, -1, -1]
R = {1: 3, 2: 4, 3: 1, 4: 2, 5: 7, 6: 8, 7: 5, 8: 6}
OMEGA = [4.0 / 9.0] + [1.0 / 9.0] * 4 + [1.0 / 36.0] * 4


def arr(seq, base=0):
    return fe._Arr({k + base: x for k, x in enumerate(seq)})


def eval_parameters(text):
    ns = {"sqrt": math.sqrt, "float": float, "int": int, "atan": math.atan, "square__": lambda x: x * x, "pow": math.pow}
    for line in fe._logical_lines(text.lower()):
        if "parameter" not in line or "::" not in line:
            continue
        rhs = fe._expr(fe._numbers(line.split("::", 1)[1]), ([], [], [], []))
        for item in fe._split_args(rhs):
            name, expr = item.split("=", 1)
            ns[name.strip()] = fe.safe_eval(expr, ns)
    return {k: v for k, v in ns.items() if isinstance(v, (int, float))}


def main():
    rng = np.random.default_rng(4242)
    out = {}
    P = eval_parameters(fe.read_lines(P4 + "/commondata.F90", 3, 60))
    names = ["pi", "total_nx", "total_ny", "l0", "t0", "radius0", "rho0", "rhosolid", "viscosity", "tauf", "snu", "sq", "gravity",
             "thresholdwall", "stiffwall", "thresholdparticle", "stiffparticle", "cnummax"]
    out["params/names"] = np.array(names)
    out["params/values"] = np.array([float(P[n]) for n in names])

    # ---- fluid cell arithmetic ----
    n = 32
    cells = []
    for _ in range(n):
        rho, u, v = 1 + 0.05 * rng.uniform(-1, 1), 0.08 * rng.uniform(-1, 1), 0.08 * rng.uniform(-1, 1)
        f = [rho * OMEGA[a] * (1 + 3 * (u * EX[a] + v * EY[a]) + 4.5 * (u * EX[a] + v * EY[a]) ** 2 - 1.5 * (u * u + v * v)) *
             (1 + 0.02 * rng.uniform(-1, 1)) for a in range(9)]
        cells.append(dict(f=f, rho=1 + 0.05 * rng.uniform(-1, 1), u=0.08 * rng.uniform(-1, 1), v=0.08 * rng.uniform(-1, 1)))
    src = fe.translate(fe.read_lines(P4 + "/fluid.F90", 14, 64), cell_arrays=["f", "f_post"], fields=["rho", "u", "v"],
                       local_arrays=["m", "m_post", "meq", "s"])
    res = [fe.run(src, cell_in={"f": c["f"]}, field_in={k: c[k] for k in ("rho", "u", "v")}, scalars={"snu": P["snu"], "sq": P["sq"]},
                  local_arrays=["m", "m_post", "meq", "s"], cell_out=["f_post"]) for c in cells]
    out["collision/f"] = np.array([c["f"] for c in cells])
    out["collision/ruv"] = np.array([[c[k] for k in ("rho", "u", "v")] for c in cells])
    out["collision/f_post"] = np.array([r["f_post"] for r in res])
    src = fe.translate(fe.read_lines(P4 + "/fluid.F90", 174, 176), cell_arrays=["f"], fields=["rho", "u", "v"])
    res = [fe.run(src, cell_in={"f": c["f"]}, field_out=["rho", "u", "v"]) for c in cells]
    out["macro/ruv"] = np.array([[r[k] for k in ("rho", "u", "v")] for r in res])
    src = fe.translate(fe.read_lines(P4 + "/initial.F90", 147, 151), cell_arrays=["f"], fields=["u", "v"],
                       local_arrays=["ex", "ey", "omega", "un"])
    res = [fe.run(src, field_in={"u": c["u"], "v": c["v"]}, scalars={"ex__": arr(EX), "ey__": arr(EY), "omega__": arr(OMEGA)},
                  local_arrays=["un"], cell_out=["f"]) for c in cells]
    out["feq/uv"] = np.array([[c["u"], c["v"]] for c in cells])
    out["feq/f"] = np.array([r["f"] for r in res])

    # ---- one particle, links crossing its surface: calQ, interpolated bounce-back, momentum exchange ----
    calq_src = fe.translate(fe.read_lines(P4 + "/particle_bounceback.F90", 112, 139), full_arrays=["xcenter", "ycenter", "radius", "ex", "ey"])
    bb_src = fe.translate(fe.read_lines(P4 + "/particle_bounceback.F90", 62, 75),
                          full_arrays=["xcenter", "ycenter", "radius", "ex", "ey", "r", "omega", "uc", "vc", "rationalomega", "f", "f_post"])
    fl_src = fe.translate(fe.read_lines(P4 + "/particle_force.F90", 56, 62),
                          full_arrays=["xcenter", "ycenter", "radius", "ex", "ey", "r", "uc", "vc", "rationalomega", "f", "f_post"])
    links = []
    xc, yc, rad = 40.37, 55.81, 10.0
    Uc, Vc, om, rhoAvg = 0.013, -0.021, 0.0017, 1.0003
    for i in range(int(xc - rad - 2), int(xc + rad + 3)):
        for j in range(int(yc - rad - 2), int(yc + rad + 3)):
            if (i - xc) ** 2 + (j - yc) ** 2 <= rad * rad:
                continue
            for a in range(1, 9):
                ip, jp = i + EX[a], j + EY[a]
                if (ip - xc) ** 2 + (jp - yc) ** 2 <= rad * rad:
                    links.append((i, j, a))
    links = [links[q] for q in rng.permutation(len(links))[:48]]
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
        ff = fe._Arr({k: v * (1 + 0.05 * rng.uniform(-1, 1)) for k, v in fpatch.items()})
        common = {"xcenter__": arr([xc], 1), "ycenter__": arr([yc], 1), "radius__": arr([rad], 1), "ex__": arr(EX), "ey__": arr(EY),
                  "r__": fe._Arr(R), "omega__": arr(OMEGA), "uc__": arr([Uc], 1), "vc__": arr([Vc], 1), "rationalomega__": arr([om], 1),
                  "cnum": 1, "alpha": a, "i": i, "j": j, "q": q["q_out"], "x0": q["x0_out"], "y0": q["y0_out"], "rhoavg": rhoAvg}
        f_bb = fe._Arr(dict(ff))
        fe.run(bb_src, scalars={**common, "f__": f_bb, "f_post__": fp})
        fo = fe.run(fl_src + "\nfx_out__ = tempforcex\nfy_out__ = tempforcey\ntq_out__ = temptorque",
                    scalars={**common, "f__": ff, "f_post__": fp}, field_out=["fx_out", "fy_out", "tq_out"])
        rec.append(dict(link=(i, j, a), q=(q["q_out"], q["x0_out"], q["y0_out"]),
                        fpost=[[fp[(b, i - s * EX[a], j - s * EY[a])] for b in range(9)] for s in range(3)],
                        f=[ff[(b, i, j)] for b in range(9)], bb=f_bb[(R[a], i, j)], force=(fo["fx_out"], fo["fy_out"], fo["tq_out"])))
    out["link/particle"] = np.array([xc, yc, rad, Uc, Vc, om, rhoAvg])
    out["link/ija"] = np.array([r["link"] for r in rec])
    out["link/q_x0_y0"] = np.array([r["q"] for r in rec])
    out["link/fpost_0_1_2"] = np.array([r["fpost"] for r in rec])          # f_post(:, x - s e_alpha), s = 0,1,2
    out["link/f"] = np.array([r["f"] for r in rec])
    out["link/bb"] = np.array([r["bb"] for r in rec])
    out["link/force"] = np.array([r["force"] for r in rec])
    assert (out["link/q_x0_y0"][:, 0] < 0.5).any() and (out["link/q_x0_y0"][:, 0] >= 0.5).any()

    # ---- per-particle forces: springs between particles, wall springs, weight ----
    fsrc = fe.translate(fe.read_lines(P4 + "/particle_force.F90", 96, 190),
                        full_arrays=["xcenter", "ycenter", "radius", "local_mask", "fxij", "fyij", "fwxij", "fwyij", "walltotalforcex",
                                     "walltotalforcey", "totaltorque"])
    N = 7
    X = [30.0, 54.5, 100.0, 186.0, 12.5, 100.0, 150.0]
    Y = [40.0, 44.0, 13.0, 300.0, 500.0, 37.5, 700.0]       # 0-1 close, 2 near bottom wall + close to 5, 3 near right wall, 4 near left wall
    rads = [10.0] * N
    hyd = [[1e-3 * rng.uniform(-1, 1) for _ in range(N)] for _ in range(3)]
    sc = {k: P[k] for k in ("pi", "rhosolid", "rho0", "gravity", "stiffparticle", "stiffwall", "thresholdparticle", "thresholdwall",
                            "radius0", "total_nx")}
    sc.update({"cnummax": N, "rhoavg": 1.0004, "itc": 1, "xcenter__": arr(X, 1), "ycenter__": arr(Y, 1), "radius__": arr(rads, 1),
               "local_mask__": arr([1] * N, 1), "walltotalforcex__": arr(hyd[0], 1), "walltotalforcey__": arr(hyd[1], 1),
               "totaltorque__": arr(hyd[2], 1), "fxij__": fe._Arr(), "fyij__": fe._Arr(), "fwxij__": fe._Arr(), "fwyij__": fe._Arr()})
    fe.run(fsrc, scalars=sc)
    out["forces/xy_rad"] = np.array([X, Y, rads])
    out["forces/hydro"] = np.array(hyd)
    out["forces/rhoAvg"] = np.array([1.0004])
    out["forces/total"] = np.array([[sc["walltotalforcex__"][c] for c in range(1, N + 1)], [sc["walltotalforcey__"][c] for c in range(1, N + 1)]])
    assert any(abs(out["forces/total"][0][c] - hyd[0][c]) > 1e-6 for c in range(N))

    # ---- explicit kinematics ----
    asrc = fe.translate(fe.read_lines(P4 + "/particle_update.F90", 37, 50),
                        full_arrays=["walltotalforcex", "walltotalforcey", "totaltorque", "radius", "uc", "vc", "ucold", "vcold", "rationalomega",
                                     "rationalomegaold", "xcenter", "ycenter", "xcenterold", "ycenterold"])
    adv_in, adv_out = [], []
    for _ in range(16):
        v = [1e-2 * rng.uniform(-1, 1), 1e-2 * rng.uniform(-1, 1), 1e-2 * rng.uniform(-1, 1), 10.0, 50 + 100 * rng.random(), 50 + 700 * rng.random(),
             0.05 * rng.uniform(-1, 1), 0.05 * rng.uniform(-1, 1), 1e-3 * rng.uniform(-1, 1)]
        ns = {k: P[k] for k in ("pi", "radius0", "rhosolid")}
        ns.update({"cnum": 1, "walltotalforcex__": arr([v[0]], 1), "walltotalforcey__": arr([v[1]], 1), "totaltorque__": arr([v[2]], 1),
                   "radius__": arr([v[3]], 1), "xcenterold__": arr([v[4]], 1), "ycenterold__": arr([v[5]], 1), "ucold__": arr([v[6]], 1),
                   "vcold__": arr([v[7]], 1), "rationalomegaold__": arr([v[8]], 1), "uc__": fe._Arr(), "vc__": fe._Arr(),
                   "rationalomega__": fe._Arr(), "xcenter__": fe._Arr(), "ycenter__": fe._Arr()})
        fe.run(asrc, scalars=ns)
        adv_in.append(v)
        adv_out.append([ns["xcenter__"][1], ns["ycenter__"][1], ns["uc__"][1], ns["vc__"][1], ns["rationalomega__"][1]])
    out["advance/in"] = np.array(adv_in)          # Fx, Fy, torque, radius, xOld, yOld, UOld, VOld, omegaOld
    out["advance/out"] = np.array(adv_out)        # x, y, U, V, omega

    # ---- refill of a newly uncovered node ----
    rsrc = fe.translate(fe.read_lines(P4 + "/particle_update.F90", 137, 192),
                        full_arrays=["xcenter", "ycenter", "ex", "ey", "uc", "vc", "rationalomega", "f", "rho", "u", "v", "m"])
    rf = []
    for _ in range(16):
        ang = rng.uniform(0, 2 * math.pi)
        xcn, ycn = 60.3 + rng.random(), 80.6 + rng.random()
        i, j = int(round(xcn + 10.4 * math.cos(ang))), int(round(ycn + 10.4 * math.sin(ang)))
        fpatch = {(b, i + di, j + dj): OMEGA[b] * (1 + 0.1 * rng.uniform(-1, 1)) for b in range(9) for di in range(-3, 4) for dj in range(-3, 4)}
        ff = fe._Arr(dict(fpatch))
        ns = {"cnum": 1, "i": i, "j": j, "i_start_global": 0, "j_start_global": 0, "rhoavg": 1.0002, "xcenter__": arr([xcn], 1), "ycenter__": arr([ycn], 1),
              "ex__": arr(EX), "ey__": arr(EY), "uc__": arr([0.02], 1), "vc__": arr([-0.03], 1), "rationalomega__": arr([0.001], 1), "f__": ff,
              "rho__": fe._Arr(), "u__": fe._Arr(), "v__": fe._Arr(), "m__": fe._Arr()}
        fe.run(rsrc, scalars=ns)
        rf.append(dict(ij=(i, j), c=(xcn, ycn), patch=[[[fpatch[(b, i + di, j + dj)] for dj in range(-3, 4)] for di in range(-3, 4)] for b in range(9)],
                       f=[ff[(b, i, j)] for b in range(9)], ruv=(ns["rho__"][(i, j)], ns["u__"][(i, j)], ns["v__"][(i, j)])))
    out["refill/ij"] = np.array([r["ij"] for r in rf])
    out["refill/center"] = np.array([r["c"] for r in rf])
    out["refill/scal"] = np.array([1.0002, 0.02, -0.03, 0.001])      # rhoAvg, Uc, Vc, omega
    out["refill/patch"] = np.array([r["patch"] for r in rf])         # f(b, i-3..i+3, j-3..j+3) before
    out["refill/f"] = np.array([r["f"] for r in rf])
    out["refill/ruv"] = np.array([r["ruv"] for r in rf])

    # ---- MPI_Dims_create_2d ----
    dsrc = fe.translate(fe.read_lines(P4 + "/mpi_starts.F90", 166, 178), full_arrays=["dims"])
    dd = []
    for np_ in range(1, 13):
        ns = {"num_process": np_, "total_nx": 201, "total_ny": 801, "dims__": fe._Arr({0: 0, 1: 0})}
        fe.run(dsrc, scalars=ns)
        dd.append([np_, ns["dims__"][0], ns["dims__"][1]])
    out["dims/np_d0_d1"] = np.array(dd)

    np.savez_compressed(os.path.join(HERE, "ref_fortran_particles.npz"), **out)
    print("wrote", len(out), "arrays; gravity =", P["gravity"], "tauf =", P["tauf"], "dims(8) =", dd[7])


if __name__ == "__main__":
    main()
