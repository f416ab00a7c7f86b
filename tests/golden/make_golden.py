"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/lid3d.c).

The reference (Fortran + MPI) cannot be built or imported in this image, so these vectors are the
oracle's own output on seeded inputs: they guard the oracle against regressions and give the GPU
parity tests a fixed target that does not depend on rebuilding the oracle.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402


def lid(total, nsteps, seed):
    wd = orc.LidWorld(total, 1)
    wd.initial()
    rng = np.random.default_rng(seed)
    rho = 1.0 + 0.01 * rng.uniform(-1, 1, total)
    u, v, w = (0.05 * rng.uniform(-1, 1, total) for _ in range(3))
    for k, a in (("rho", rho), ("u", u), ("v", v), ("w", w)):
        wd.scatter(k, np.asfortranarray(a))
    wd.scatter("f", np.asfortranarray(orc.feq(rho, u, v, w)))
    wd.step(nsteps)
    out = {k: wd.gather(k) for k in ("rho", "u", "v", "w")}
    out["errorU"] = wd.check()
    out["nsteps"], out["seed"] = nsteps, seed
    wd.close()
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    np.savez_compressed(os.path.join(here, "lid_9x8x7.npz"), **lid((9, 8, 7), 10, 1234))
    print("wrote", os.path.join(here, "lid_9x8x7.npz"))
