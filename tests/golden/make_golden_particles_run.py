"""Generates tests/golden/ref_fortran_particles_run.npz -- the reference's particle-laden D2Q9 program run from its own source
text (fortran_eval.py, whole arrays) on ONE rank, 56 x 72 lattice, two particles (positions are an input: the program draws its
own from a compiler-specific random_number, P4/initial.F90:52-72):

  P4 = /root/reference/MPI/Micro_particles/fortran/case4/mpi_particle
  parameters           P4/commondata.F90:3-60 (total_nx, total_ny, cNumMax replaced)
  initial              P4/initial.F90:106-116 (mask), :118-124, :133-139, :144-154, :158-188 (ghost layers)
  collision            P4/fluid.F90:10-68          streaming   P4/fluid.F90:97-108      bounceback  P4/fluid.F90:121-158
  bounceback_particle  P4/particle_bounceback.F90:15-27, :35, :40-95 with calQ :112-139 called as written
  macro                P4/fluid.F90:170-180
  calForce             P4/particle_force.F90:35-87 (link sums), :96-190 (springs, walls, weight)
  updateCenter         P4/particle_update.F90:26-52, :81-102, :105-115, :120, :126-203
  check                P4/fluid.F90:195-209, :216
Its loop (P4/main.F90:35-71: collision, send_all_fp, streaming, bounceback, bounceback_particle, macro, calForce, send_all_f,
updateCenter) for 1, 2 and 30 iterations.  What one rank turns into the identity is applied by this script: MPI_Allreduce over
one rank, the halo exchanges without neighbours, update_particle_mask() with local_mask = 1, and the whole-array assignments
(zeroing of the force sums :23-29, obst = obstNew :206).  The restatement (oracle/particles2d.c) must reproduce every
population, field, mask and particle state bit for bit.  Only numbers are stored; run in the authoring container."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fortran_eval as fe  # noqa: E402
from make_golden_particles import EX, EY, OMEGA, P4, R, arr, eval_parameters  # noqa: E402

FULL = ["f", "f_post", "rho", "u", "v", "up", "vp", "obst", "obstnew", "ex", "ey", "r", "omega", "un", "s", "m", "m_post", "meq",
        "xcenter", "ycenter", "xcenterold", "ycenterold", "uc", "vc", "ucold", "vcold", "rationalomega", "rationalomegaold", "radius",
        "walltotalforcex", "walltotalforcey", "totaltorque", "force_x", "force_y", "torque", "fxij", "fyij", "fwxij", "fwyij",
        "local_mask", "coords", "dims"]
CALQ_CALL = "call calq(cnum,dble(i + i_start_global),dble(j + j_start_global),alpha,x0,y0,q)"


def main():
    nx, ny, N = 56, 72, 2
    text = fe.read_lines(P4 + "/commondata.F90", 3, 60)
    assert "total_nx=201, total_ny=801" in text and "cNumMax=64" in text
    P = eval_parameters(text.replace("total_nx=201, total_ny=801", f"total_nx={nx}, total_ny={ny}").replace("cNumMax=64", f"cNumMax={N}"))
    names_p = ["pi", "radius0", "rho0", "rhosolid", "viscosity", "tauf", "snu", "sq", "gravity", "thresholdwall", "stiffwall",
               "thresholdparticle", "stiffparticle"]
    out = {"params": np.array([float(P[k]) for k in names_p]), "shape": np.array([nx, ny, N])}
    # 3.5 lattice units apart: the particle-particle spring is active; the node (28, 60) sits 1e-3 inside the upper particle, so
    # it is uncovered as the particle starts to sink and the refill branch of updateCenter() runs
    x0s, y0s = [28.01, 28.0], [50.001, 26.501]
    out["positions"] = np.array([x0s, y0s])

    import math
    fn = {"square__": lambda x: x * x, "pow": math.pow, "mod": lambda a, b: a % b}
    sc = dict(P, nx=nx, ny=ny, cnummax=N, i_start_global=0, j_start_global=0, itc=0, epsradius=float(np.float32(1e-9)), **fn)

    def tr(path, a, b, sub=None):
        t = fe.read_lines(P4 + "/" + path, a, b)
        if sub:
            low = "\n".join(fe._logical_lines(t.lower()))
            assert sub[0] in low, (path, a, b)
            return fe.translate(low.replace(sub[0], sub[1]), full_arrays=FULL)
        return fe.translate(t, full_arrays=FULL)

    calq_call = (CALQ_CALL, "x0, y0, q = calq__(cnum, float(i + i_start_global), float(j + j_start_global), alpha)")
    src = {
        "mask0": tr("initial.F90", 106, 116), "rho_solid": tr("initial.F90", 118, 124), "weights": tr("initial.F90", 133, 139),
        "feq": tr("initial.F90", 144, 154), "ghosts": tr("initial.F90", 158, 188),
        "collision": tr("fluid.F90", 10, 68), "streaming": tr("fluid.F90", 97, 108), "bounceback": tr("fluid.F90", 121, 158),
        "rho_sum": tr("particle_bounceback.F90", 15, 27), "bb_particle": tr("particle_bounceback.F90", 40, 95, calq_call),
        "macro": tr("fluid.F90", 170, 180), "links": tr("particle_force.F90", 35, 87, calq_call),
        "forces": tr("particle_force.F90", 96, 190), "advance": tr("particle_update.F90", 26, 52),
        "mask": tr("particle_update.F90", 81, 102), "rho_sum_new": tr("particle_update.F90", 105, 115),
        "refill": tr("particle_update.F90", 126, 203), "check": tr("fluid.F90", 195, 209),
    }
    calq_src = fe.translate(fe.read_lines(P4 + "/particle_bounceback.F90", 112, 139), full_arrays=["xcenter", "ycenter", "radius", "ex", "ey"])
    calq_code = fe.safe_compile(calq_src, "<calQ>")

    zeros = lambda keys: fe._Arr({k: 0.0 for k in keys})
    cells = [(i, j) for i in range(1, nx + 1) for j in range(1, ny + 1)]
    st = {
        "f": zeros((a, i, j) for a in range(9) for i in range(-2, nx + 4) for j in range(-2, ny + 4)),                 # P4/initial.F90:141
        "f_post": zeros((a, i, j) for a in range(9) for i in range(-1, nx + 3) for j in range(-1, ny + 3)),            # :142
        "obst": fe._Arr({(i, j): 0 for i in range(nx + 2) for j in range(ny + 2)}),                                     # :103
        "obstnew": fe._Arr({(i, j): 0 for i in range(nx + 2) for j in range(ny + 2)}),                                  # :104
        "rho": fe._Arr({c: float(P["rho0"]) for c in cells}),                                                           # :105
        "u": zeros(cells), "v": zeros(cells), "up": zeros(cells), "vp": zeros(cells),                                   # :127-131
        "ex": arr(EX), "ey": arr(EY), "r": fe._Arr(R), "omega": fe._Arr(), "un": fe._Arr(), "s": fe._Arr(), "m": fe._Arr(),
        "m_post": fe._Arr(), "meq": fe._Arr(),
        "xcenter": arr(x0s, 1), "ycenter": arr(y0s, 1), "xcenterold": arr(x0s, 1), "ycenterold": arr(y0s, 1),             # :81-82
        "radius": arr([float(P["radius0"])] * N, 1),                                                                    # :84
        **{k: arr([0.0] * N, 1) for k in ("uc", "vc", "rationalomega", "ucold", "vcold", "rationalomegaold", "walltotalforcex",
                                          "walltotalforcey", "totaltorque", "force_x", "force_y", "torque", "fxij", "fyij",
                                          "fwxij", "fwyij")},                                                           # :95-101
        "local_mask": arr([1] * N, 1), "coords": arr([0, 0]), "dims": arr([1, 1]),                                      # one rank
    }
    names = list(st)
    carry = {"rhoavg": 0.0}

    def calq(cnum, i, j, alpha):
        ns = {"sqrt": math.sqrt, "abs": abs, "float": float, "int": int, "range": range, **fn, "epsradius": sc["epsradius"],
              "xcenter__": st["xcenter"], "ycenter__": st["ycenter"], "radius__": st["radius"], "ex__": st["ex"], "ey__": st["ey"],
              "cnum": cnum, "i": i, "j": j, "alpha": alpha}
        fe.safe_exec(calq_code, ns, "<calQ>")
        return ns["x0"], ns["y0"], ns["q"]

    def call(sub, **scalars):
        ns = {"sqrt": math.sqrt, "abs": abs, "float": float, "int": int, "range": range, "calq__": calq}
        ns.update(sc)
        ns.update(carry)
        ns.update(scalars)
        for k in names:
            ns[k + "__"] = st[k]
        fe.safe_exec(src[sub], ns, "<" + sub + ">")
        for k in names:
            st[k] = ns[k + "__"]
        return ns

    def rho_avg(sub):
        ns = call(sub, rhoavg=0.0, fluidnum=0, total_rho=0.0, total_fluidnum=0)
        carry["rhoavg"] = ns["rhoavg"] / float(ns["fluidnum"])           # Allreduce over one rank, then :35 / :120

    def step():
        call("collision")                                               # send_all_fp: no neighbours on one rank
        call("streaming")
        call("bounceback")
        rho_avg("rho_sum")
        call("bb_particle")
        call("macro")
        for k in ("force_x", "force_y", "torque"):                      # P4/particle_force.F90:27-29
            st[k] = arr([0.0] * N, 1)
        call("links")
        st["walltotalforcex"], st["walltotalforcey"], st["totaltorque"] = (fe._Arr(st[k]) for k in ("force_x", "force_y", "torque"))   # :90-92
        call("forces")                                                  # send_all_f: no neighbours
        call("advance")                                                 # update_particle_mask(): identity with local_mask = 1
        call("mask")
        rho_avg("rho_sum_new")
        call("refill")
        st["obst"] = fe._Arr(st["obstnew"])                             # P4/particle_update.F90:206

    def snap(tag):
        out[tag + "/f"] = np.array([[[st["f"][(a, i, j)] for j in range(-2, ny + 4)] for i in range(-2, nx + 4)] for a in range(9)])
        out[tag + "/f_post"] = np.array([[[st["f_post"][(a, i, j)] for j in range(-1, ny + 3)] for i in range(-1, nx + 3)] for a in range(9)])
        out[tag + "/ruv"] = np.array([[[st[k][(i, j)] for j in range(1, ny + 1)] for i in range(1, nx + 1)] for k in ("rho", "u", "v")])
        out[tag + "/obst"] = np.array([[st["obst"][(i, j)] for j in range(ny + 2)] for i in range(nx + 2)], dtype=np.int32)
        keys = ("xcenter", "ycenter", "uc", "vc", "rationalomega", "walltotalforcex", "walltotalforcey", "totaltorque", "xcenterold",
                "ycenterold")
        out[tag + "/particles"] = np.array([[st[k][c] for c in range(1, N + 1)] for k in keys])
        out[tag + "/rhoAvg"] = np.array([carry["rhoavg"]])

    for sub in ("mask0", "rho_solid", "weights", "feq", "ghosts"):
        call(sub)
    snap("run0")
    done = 0
    for n in (1, 2, 30):
        for _ in range(n - done):
            step()
        done = n
        snap(f"run{n}")
    ns = call("check", error1=0.0, error2=0.0)
    out["run30/check"] = np.array([ns["error1"], ns["error2"], math.sqrt(ns["error1"]) / math.sqrt(ns["error2"])])     # P4/fluid.F90:216
    path = os.path.join(HERE, "ref_fortran_particles_run.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
