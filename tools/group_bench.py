"""P subdomains of the lid-driven cavity driven by ONE process on P GPUs (mglc_group_*: direct halo stores through peer
pointers, ordering by events).  Used under ncu -- which must not wrap a multi-rank launch -- to capture fast::k_fused<0,1> with
its NVLink traffic, and to time one decomposition axis at a time.
    python tools/group_bench.py --gpus 2 --dims 2,1,1 --size 768 --steps 10"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mglc_b200 as mg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--dims", default="")
    ap.add_argument("--size", type=int, default=768)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--thermal", action="store_true")
    ap.add_argument("--aa", action="store_true", help="single-lattice (AA-pattern) blocks, mglc_aa_group_*")
    a = ap.parse_args()
    dims = tuple(int(x) for x in a.dims.split(",")) if a.dims else mg.dims_create(a.gpus)
    total = tuple(a.size * d for d in dims)
    if a.aa:
        sim = mg.LidDrivenCavityAA(total, nranks=a.gpus, dims=dims, devices=list(range(a.gpus)), arith="fast") if a.gpus > 1 \
            else mg.LidDrivenCavityAA(total, arith="fast")
    else:
        Driver = mg.BuoyancyDrivenCavity if a.thermal else mg.LidDrivenCavity
        sim = Driver(total, nprocs=a.gpus, dims=dims, devices=list(range(a.gpus)), arith="fast")
    sim.initial()
    sim.step(3); sim.sync()
    ms = sim.step_timed(a.steps)
    cells = total[0] * total[1] * total[2]
    print(json.dumps({"dims": dims, "per_gpu": a.size, "gpus": a.gpus, "ms_per_step": round(ms / a.steps, 4),
                      "mlups": round(cells * a.steps / ms / 1e3, 1), "thermal": a.thermal, "aa": a.aa}), flush=True)
    sim.close()


if __name__ == "__main__":
    main()
