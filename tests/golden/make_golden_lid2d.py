"""Generates tests/golden/ref_lid2d.npz -- golden vectors for the 2-D D2Q9 lid-driven cavity path, from the REFERENCE itself:

  c_*   outputs of the reference's own compiled C program (MPI/Lid_driven_cavity/c/lid_driven_cavity.c, built unmodified
        into oracle/_ref/liblid2d_ref.so by `make -C oracle ref`): its collision() on seeded populations written into its
        global arrays, and the fields of its own run (initial() + N x {collision, streaming, boundary, macro}) on its
        shipped 200 x 200 grid, sampled on a few rows/columns plus checksums, and check()'s residual.
  f_*   the Fortran variant's per-cell arithmetic (MPI/Lid_driven_cavity/fortran/2d/2d_revised/mpi_blocked/evolution.f90
        collision :16-66, macro :105-107; initial.f90:62-66), machine-evaluated from the source text (fortran_eval.py).

Run in the authoring container (reads /root/reference through oracle/_ref and the evaluator); only numbers are stored."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import fortran_eval as fe  # noqa: E402
from oracle import oracle as orc  # noqa: E402

L2F = "/root/reference/MPI/Lid_driven_cavity/fortran/2d/2d_revised/mpi_blocked"
EX = [0, 1, 0, -1, 0, 1, -1, -1, 1]
EY = [0, 0, 1, 0, -1, 1, 1, -1, -1]
W9 = [4 / 9] + [1 / 9] * 4 + [1 / 36] * 4


def arr(seq):
    return fe._Arr({k: x for k, x in enumerate(seq)})


def random_cells(rng, n):
    cells = []
    for _ in range(n):
        rho, u, v = 1.0 + 0.05 * rng.uniform(-1, 1), 0.08 * rng.uniform(-1, 1), 0.08 * rng.uniform(-1, 1)
        f = [rho * W9[a] * (1 + 3 * (u * EX[a] + v * EY[a]) + 4.5 * (u * EX[a] + v * EY[a]) ** 2 - 1.5 * (u * u + v * v)) *
             (1 + 0.02 * rng.uniform(-1, 1)) for a in range(9)]
        cells.append(dict(f=f, rho=1.0 + 0.05 * rng.uniform(-1, 1), u=0.08 * rng.uniform(-1, 1), v=0.08 * rng.uniform(-1, 1)))
    return cells


def main():
    rng = np.random.default_rng(20262)
    out = {}
    cells = random_cells(rng, 48)
    out["cells/f"] = np.array([c["f"] for c in cells])
    out["cells/ruv"] = np.array([[c[k] for k in ("rho", "u", "v")] for c in cells])

    # ---------------- the compiled C reference ----------------
    cwd = os.getcwd()
    import tempfile
    os.chdir(tempfile.mkdtemp())              # its check()/output_*() write log files into the cwd
    ref = orc.RefLid2D()
    ref.lib.initial()                          # sets tau, s_nu, s_q
    out["c/params"] = np.array([ref.scalar("tau"), ref.scalar("s_nu"), ref.scalar("s_q")])
    out["c/initial_f_top"] = ref.f[:, -1, :].copy()        # lid row j = NY-1
    out["c/initial_f_bulk"] = ref.f[7, 3, :].copy()
    # (1) collision() on seeded cells placed in its arrays
    for k, c in enumerate(cells):
        ref.f[k, 0, :] = c["f"]
        ref.rho[k, 0], ref.u[k, 0], ref.v[k, 0] = c["rho"], c["u"], c["v"]
    ref.lib.collision()
    out["c/collision_f_post"] = ref.f_post[:len(cells), 0, :].copy()
    # (2) its own run from initial()
    ref.lib.initial()
    rows = {}
    done = 0
    for n in (1, 10, 100, 1000):
        ref.step(n - done)
        done = n
        for k in ("rho", "u", "v"):
            a = getattr(ref, k)
            out[f"c/run{n}/{k}_col100"] = a[100, :].copy()          # x = 100, all y
            out[f"c/run{n}/{k}_row199"] = a[:, 199].copy()          # lid row
            out[f"c/run{n}/{k}_row0"] = a[:, 0].copy()
            out[f"c/run{n}/{k}_sum"] = np.array([a.sum(), np.abs(a).sum(), (a * a).sum()])
        out[f"c/run{n}/f_corner"] = ref.f[:3, :3, :].copy()
        out[f"c/run{n}/f_topright"] = ref.f[-3:, -3:, :].copy()
    out["c/check_1000"] = np.array([ref.lib.check(1000)])           # up = vp = 0 before: 1.0
    ref.step(1000)
    out["c/check_2000"] = np.array([ref.lib.check(2000)])
    out["c/run2000/u_col100"] = ref.u[100, :].copy()
    out["c/run2000/v_row100"] = ref.v[:, 100].copy()
    ref.lib.output_binary()
    raw = open("flow_binary", "rb").read()
    out["c/output_binary_len"] = np.array([len(raw)])
    import hashlib
    out["c/output_binary_sha256"] = np.frombuffer(hashlib.sha256(raw).digest(), dtype=np.uint8)
    out["c/run2000/rho_full"] = ref.rho.copy()                       # the full fields behind that file (3 x 320 kB)
    out["c/run2000/u_full"] = ref.u.copy()
    out["c/run2000/v_full"] = ref.v.copy()
    os.chdir(cwd)

    # ---------------- the Fortran variant's source text ----------------
    tau = 0.1 * 201 / 1000.0 * 3.0 + 0.5                              # commondata.f90:9
    snu, sq = 1.0 / tau, 8.0 * (2.0 * tau - 1.0) / (8.0 * tau - 1.0)   # commondata.f90:31
    out["f/params"] = np.array([tau, snu, sq])
    la = ["s", "m", "m_post", "meq"]
    src = fe.translate(fe.read_lines(L2F + "/evolution.f90", 16, 66), cell_arrays=["f", "f_post"], fields=["rho", "u", "v"], local_arrays=la)
    res = [fe.run(src, cell_in={"f": c["f"]}, field_in={k: c[k] for k in ("rho", "u", "v")}, scalars={"snu": snu, "sq": sq},
                  local_arrays=la, cell_out=["f_post"]) for c in cells]
    out["f/collision_f_post"] = np.array([r["f_post"] for r in res])
    src = fe.translate(fe.read_lines(L2F + "/evolution.f90", 105, 107), cell_arrays=["f"], fields=["rho", "u", "v"])
    res = [fe.run(src, cell_in={"f": c["f"]}, field_out=["rho", "u", "v"]) for c in cells]
    out["f/macro_ruv"] = np.array([[r[k] for k in ("rho", "u", "v")] for r in res])
    src = fe.translate(fe.read_lines(L2F + "/initial.f90", 62, 66), cell_arrays=["f"], fields=["rho", "u", "v"],
                       local_arrays=["ex", "ey", "omega", "un"])
    om = arr([4.0 / 9.0] + [1.0 / 9.0] * 4 + [1.0 / 36.0] * 4)         # initial.f90:48-54
    res = [fe.run(src, field_in={k: c[k] for k in ("rho", "u", "v")}, scalars={"ex__": arr(EX), "ey__": arr(EY), "omega__": om},
                  local_arrays=["un"], cell_out=["f"]) for c in cells]
    out["f/feq"] = np.array([r["f"] for r in res])

    path = os.path.join(HERE, "ref_lid2d.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
