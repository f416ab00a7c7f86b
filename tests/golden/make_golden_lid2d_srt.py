"""Generates tests/golden/ref_lid2d_srt.npz -- golden vectors of the reference's C lid-driven cavity program run with its own
model switch set to SRT / BGK (MPI/Lid_driven_cavity/c/lid_driven_cavity.c:13-14, collision c:160-176), from the REFERENCE itself:
oracle/_ref/liblid2d_srt_ref.so is that program compiled by `make -C oracle ref` with the one token `model = 2` -> `model = 1`
changed in the compiler's input stream.  Stored: its collision() on the seeded cells of ref_lid2d.npz written into its global
arrays, and the fields of its own run (initial() + N x {collision, streaming, boundary, macro}) on its shipped 200 x 200 grid,
sampled on rows / columns / a stride-4 grid plus checksums, and check()'s residual.  Run in the authoring container."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402


def main():
    cells = np.load(os.path.join(HERE, "ref_lid2d.npz"))
    f, ruv = cells["cells/f"], cells["cells/ruv"]
    out = {}
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())              # its check()/output_*() write log files into the cwd
    ref = orc.RefLid2D(os.path.join(ROOT, "oracle", "_ref", "liblid2d_srt_ref.so"))
    assert orc.C.c_int.in_dll(ref.lib, "model").value == 1
    ref.lib.initial()
    out["params"] = np.array([ref.scalar("tau"), ref.scalar("s_nu"), ref.scalar("s_q")])
    for k in range(len(f)):
        ref.f[k, 0, :] = f[k]
        ref.rho[k, 0], ref.u[k, 0], ref.v[k, 0] = ruv[k]
    ref.lib.collision()
    out["collision_f_post"] = ref.f_post[:len(f), 0, :].copy()
    ref.lib.initial()
    done = 0
    for n in (1, 10, 100, 1000):
        ref.step(n - done)
        done = n
        for k in ("rho", "u", "v"):
            a = getattr(ref, k)
            out[f"run{n}/{k}_col100"] = a[100, :].copy()
            out[f"run{n}/{k}_row199"] = a[:, 199].copy()
            out[f"run{n}/{k}_row0"] = a[:, 0].copy()
            out[f"run{n}/{k}_stride4"] = a[::4, ::4].copy()
            out[f"run{n}/{k}_sum"] = np.array([a.sum(), np.abs(a).sum(), (a * a).sum()])
        out[f"run{n}/f_corner"] = ref.f[:3, :3, :].copy()
        out[f"run{n}/f_topright"] = ref.f[-3:, -3:, :].copy()
    out["check_1000"] = np.array([ref.lib.check(1000)])
    ref.step(100)
    out["check_1100"] = np.array([ref.lib.check(1100)])
    # the program's own output file after 2000 iterations (what examples/lid2d_driver.c must write with model = 1)
    ref.lib.initial()
    ref.step(2000)
    ref.lib.output_binary()
    raw = open("flow_binary", "rb").read()
    import hashlib
    out["output_binary_len"] = np.array([len(raw)])
    out["output_binary_sha256"] = np.frombuffer(hashlib.sha256(raw).digest(), dtype=np.uint8)
    out["run2000/u_sum"] = np.array([ref.u.sum(), np.abs(ref.u).sum()])
    os.chdir(cwd)
    path = os.path.join(HERE, "ref_lid2d_srt.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
