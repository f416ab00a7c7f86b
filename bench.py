#!/usr/bin/env python
"""bench.py -- throughput of the D3Q19 lid-driven-cavity hot path on B200 (BASELINE.json config 3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload lid|thermal|jacobi|particles|...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one lattice timestep (collision, halo exchange, streaming, bounce-back, macro; the
reference's loop body, L3/main.f90:85-97) over the whole lattice.  N=1 runs the 768^3 lattice the
metric is quoted on; N>1 runs one 768^3 block per GPU (weak scaling, 3-D Cartesian decomposition,
halos by direct NVLink stores or NCCL), `--scaling strong` splits one 768^3 lattice instead.  Rank 0 prints ONE JSON line.

  value     MLUPS, whole job, lattice resident in HBM, CUDA-event time of K steps, max over ranks
  e2e       MLUPS through the C ABI with HOST arrays: upload f,rho,u,v,w from pinned host memory ->
            K steps -> check() -> download rho,u,v,w, all inside the timed region (+ the achieved copy rates)
  roofline  fused stream+collide kernel: 304 B/cell (19 loads + 19 stores of fp64) x cells per launch
            / its mean CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the CPU oracle (restated reference, OpenMP over all host threads) on a bounded sample
  N > 1 adds  parity      strict-arithmetic decomposed lid + thermal runs against the CPU oracle for every halo transport
                          (tests/dist/parity_suite.py), checked BEFORE anything is timed; a mismatch fails the run (rc 3)
              transports  the same weak-scaling lattice timed with the other halo transports
              strong      768^3 split over the N GPUs (strong scaling), timed in the same invocation

`--impl reference` times that same CPU restatement alone (the reference's Fortran+MPI cannot be built
in this image) on the same metric/config, each step a thin-slab sample of the workload.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_CELL = 304          # D3Q19 fp64 pull scheme: 19 x 8 B read + 19 x 8 B write
BYTES_PER_CELL_THERMAL = 464  # (19 + 7) x 16 B + the carried force Fx,Fy,Fz in and out (48 B); SURVEY 8d quotes 416 without it
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
HALO_NAMES = {"direct": "fused kernel stores outgoing populations into the neighbours' halos over NVLink (CUDA IPC) + flag barrier",
              "push": "plain fused kernel, then one launch that copies the outgoing populations into the neighbours' halos over NVLink (CUDA IPC) + flag barrier",
              "overlap": "NCCL send/recv on a second stream, overlapped with the interior update",
              "blocking": "NCCL send/recv, blocking before the update"}
HALO_MODE = {"direct": 2, "push": 3, "overlap": 1, "blocking": 0}


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="mglc", choices=["mglc", "reference"])
    p.add_argument("--no-overlap", action="store_true", help="same as --halo blocking")
    p.add_argument("--halo", default="auto", choices=["auto", "direct", "push", "overlap", "blocking"],
                   help="multi-GPU halo transport of the fused step: direct = stores into the neighbours' halos over NVLink "
                        "(CUDA IPC), overlap = NCCL exchange beside the interior update, blocking = NCCL exchange, then update; "
                        "auto = direct when the mappings came up, else overlap")
    p.add_argument("--workload", default="lid", choices=["lid", "thermal", "jacobi", "particles", "lid2d", "thermal2d", "lid_aa"],
                   help="lid = BASELINE.json's metric (default); thermal / jacobi / particles = the other configs, for profiles/")
    p.add_argument("--size", type=int, default=0, help="per-GPU block edge (weak) / global edge (strong); 0 = the workload's config size")
    p.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    p.add_argument("--dims", default="", help="process grid as X,Y,Z instead of MPI_Dims_create's (experiments)")
    p.add_argument("--arith", default="fast", choices=["fast", "strict"])
    p.add_argument("--variant", default="mpi", choices=["mpi", "acc"],
                   help="thermal2d only: mpi = mpi_blocked/ (side-heated), acc = the OpenACC program seq/bouyancy2d_acc.F90 "
                        "(Rayleigh-Benard plates, periodic vertical walls; --size 0 = its shipped 513 x 257)")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--no-parity", action="store_true", help="N > 1: skip the parity runs against the oracle")
    p.add_argument("--no-extras", action="store_true", help="N > 1: skip the other transports and the strong-scaling leg")
    p.add_argument("--cpu-seconds", type=float, default=12.0)
    return p.parse_args()


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, cells_per_launch):
    """dram bytes per launch from the committed ncu capture of the same kernel and launch size
    (profiles/ncu_traffic.json), or None when no capture of that size exists."""
    try:
        for cap in json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["captures"]:
            if cap["cells_per_launch"] == cells_per_launch and kernel in cap.get("kernel", kernel):
                return cap
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            time.sleep(0.3)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for n, v in zip(names, c[4:8]):
                if v == "Active":
                    reasons.add(n)
        os.unlink(self.path)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm), "power_w_max": max(pw)}
        return out


def host_mem_available():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except Exception:
        pass
    return 0


def lattice_bytes(n, thermal=False):
    px = ((n + 17 + 15) // 16) * 16
    pops, fields = (2 * 26, 19) if thermal else (2 * 19, 7)
    return pops * px * (n + 2) * (n + 2) * 8 + fields * n ** 3 * 8 + (1 << 28)


def dims_create(nranks):
    """MPI_Dims_create(nranks, 3, dims) as the reference uses it (L3/main.f90:24): the most balanced factorisation, larger
    factors first.  Pure Python so that the reference arm does not have to load the product library."""
    best = None
    for a in range(1, nranks + 1):
        if nranks % a:
            continue
        for b in range(1, nranks // a + 1):
            if (nranks // a) % b:
                continue
            d = tuple(sorted((a, b, nranks // a // b), reverse=True))
            if best is None or (d[0] - d[2], d) < (best[0] - best[2], best):
                best = d
    return best


def lattice_for(args, world):
    """(dims, per-GPU block, global lattice) of a run: weak = one size^3 block per GPU, strong = size^3 split over the GPUs.
    The process grid has the factors MPI_Dims_create gives the reference (L3/main.f90:24), assigned z first, then y, then x:
    1x1x2, 1x2x2, 2x2x2.  Rows are contiguous along x on the device, so a z or y face is a block of whole rows while an x face
    is one double per row; with fewer than 8 subdomains the faces that have to move are the cheap ones.  --dims overrides."""
    dims = tuple(int(x) for x in args.dims.split(",")) if args.dims else tuple(sorted(dims_create(world)))
    n = args.size
    per_gpu = [n, n, n] if args.scaling == "weak" else [n // d for d in dims]
    gn = tuple(p * d for p, d in zip(per_gpu, dims))
    return dims, per_gpu, gn


def lattice_config(args, thermal, dims, per_gpu, gn):
    """the `config` object: what both arms (product and --impl reference) run, spelled identically by both"""
    cells_local = per_gpu[0] * per_gpu[1] * per_gpu[2]
    return {"workload": ("buoyancy_driven_cavity_d3q19_d3q7_mrt" if thermal else "lid_driven_cavity_d3q19_mrt") +
                        f"_{per_gpu[0]}x{per_gpu[1]}x{per_gpu[2]}_per_gpu",
            "global_lattice": list(gn), "decomposition": "x".join(map(str, dims)),
            **({"Ra": 1e6, "Pr": 0.71, "Ma": 0.1, "Ek": 1e-3} if thermal else {"Re": 1000.0, "U0": 0.1}),
            "l2": "lattice (2 x %.1f GB per GPU) far exceeds the 126 MB L2; no flush needed" % ((26 if thermal else 19) * cells_local * 8 / 1e9)}


# ------------------------------------------------------------------------------------------------------
def oracle_threads(orc, native):
    """all the host threads this process may use, whatever OMP_NUM_THREADS a launcher exported (torch.distributed.run sets 1);
    returns the count the OpenMP runtime reports -- that is what the JSON line calls `cores`"""
    return orc.set_threads(0, native=native)


def cpu_baseline(seconds, thermal=False, edge=256):
    """Oracle (restated reference, AoS, un-fused five sweeps) on all host threads; bounded sample."""
    from oracle import oracle as orc
    try:
        orc.build(native=True)
        native = True
    except Exception:
        native = False
    if thermal:
        native = False             # ThermalWorld runs the parity build
        edge = 128
    threads = oracle_threads(orc, native)
    total = (edge, edge, edge)
    wd = orc.ThermalWorld(total, 1) if thermal else orc.LidWorld(total, 1, native=native)
    wd.initial()
    t0 = time.perf_counter(); wd.step(1); t1 = time.perf_counter() - t0
    n = max(1, min(200, int(seconds / max(t1, 1e-3))))
    t0 = time.perf_counter(); wd.step(n); dt = time.perf_counter() - t0
    wd.close()
    mlups = edge ** 3 * n / dt / 1e6
    src = "oracle/thermal3d.c (-O2" if thermal else f"oracle/lid3d.c ({'-O3 -march=native' if native else '-O2'}"
    return {"value": round(mlups, 2), "unit": "MLUPS", "cores": threads, "kind": "port",
            "sample": f"{edge}^3 lattice x {n} steps ({dt:.1f} s), {src}, OpenMP, no FMA contraction)"}


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement of the reference's path, all host threads, same metric/config."""
    if rank != 0:
        return
    from oracle import oracle as orc
    thermal = args.workload == "thermal"
    try:
        if thermal:
            raise RuntimeError("thermal oracle runs the parity build")
        orc.build(native=True); native = True
    except Exception:
        native = False
    threads = oracle_threads(orc, native)
    dims, per_gpu, gn = lattice_for(args, world)
    n = per_gpu[0]
    World = (lambda t: orc.ThermalWorld(t, 1)) if thermal else (lambda t: orc.LidWorld(t, 1, native=native))
    # calibrate on a thin slab, then size the per-step sample so K+W steps take ~90 s
    cal = World((n, per_gpu[1], 8))
    cal.initial()
    t0 = time.perf_counter(); cal.step(1); t1 = time.perf_counter() - t0
    cal.close()
    plane = n * per_gpu[1]
    rate = plane * 8 / t1
    nz = int(max(8, min(per_gpu[2], rate * 90.0 / max(1, args.steps + args.warmup) / plane)))
    # ... and keep the slab inside the host memory: f + halo'd f_post + seven fields ~ 400 B per cell (thermal ~ 600)
    avail = host_mem_available()
    if avail:
        nz = int(max(8, min(nz, 0.5 * avail / ((600.0 if thermal else 400.0) * plane))))
    wd = World((n, per_gpu[1], nz))
    wd.initial()
    wd.step(args.warmup)
    t0 = time.perf_counter(); wd.step(args.steps); dt = time.perf_counter() - t0
    wd.close()
    cells = plane * nz
    mlups = cells * args.steps / dt / 1e6
    sample = (f"{n}x{per_gpu[1]}x{nz} slab of one {per_gpu[0]}x{per_gpu[1]}x{per_gpu[2]} block per step, {args.steps} steps in {dt:.1f} s, "
              f"oracle/{'thermal3d' if thermal else 'lid3d'}.c on {threads} OpenMP threads (one host; the rate does not depend on N)")
    print(json.dumps({
        "impl": "reference", "metric": "MLUPS", "value": round(mlups, 2), "unit": "MLUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": lattice_config(args, thermal, dims, per_gpu, gn),
        "detail": {"note": "reference Fortran+MPI cannot be built in this image; this is its C restatement (oracle/)",
                   "omp_threads": threads, "OMP_NUM_THREADS_inherited": os.environ.get("OMP_NUM_THREADS")},
        "cpu_baseline": {"value": round(mlups, 2), "unit": "MLUPS", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(mlups, 2), "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ------------------------------------------------------------------------------------------------------
class Dist:
    """torch.distributed plumbing of one rank (NCCL); at world == 1 everything is local"""

    def __init__(self, rank, local_rank, world):
        import torch
        self.torch, self.rank, self.local, self.world = torch, rank, local_rank, world
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; mglc_b200 has no CPU path")
        torch.cuda.set_device(local_rank)
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x, op):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX) if self.dist else x

    def min(self, x):
        return self._reduce(x, self.dist.ReduceOp.MIN) if self.dist else x

    def sum(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM) if self.dist else x

    def bcast_obj(self, b):
        box = [b]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]

    def communicator(self, mg):
        return mg.Communicator(self.world, self.rank, self.local, self.bcast_obj) if self.dist else None

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


def run_parity(D, comm, thermal_too=True):
    """strict-arithmetic decomposed runs against the CPU oracle, every halo transport (the oracle is the checker here)"""
    sys.path.insert(0, os.path.join(ROOT, "tests", "dist"))
    import parity_suite as ps
    t0 = time.perf_counter()
    out = {"lid": ps.lid(comm, D.rank, D.world, total=(41, 37, 35), nsteps=12)}
    if thermal_too:
        out["thermal"] = ps.thermal(comm, D.rank, D.world, total=(27, 25, 23), nsteps=10)
        # the same lid case on ONE lattice per block (--workload lid_aa times that path): the blocks store into each other
        out["lid_aa"] = ps.lid_aa(comm, D.rank, D.world, total=(41, 37, 35), nsteps=12)
    bad = D.max(1.0 if (D.rank == 0 and ps.failed(out)) else 0.0) > 0
    flat = {"lid_41x37x35_12_steps": out["lid"], "oracle": "oracle/lid3d.c, oracle/thermal3d.c on one emulated rank",
            "arith": "strict"}
    if thermal_too:
        flat["thermal_27x25x23_10_steps"] = out["thermal"]
        flat["lid_41x37x35_12_steps_single_lattice_blocks"] = out["lid_aa"]["single_lattice"]
    flat["seconds"] = round(time.perf_counter() - t0, 1)
    # the headline keys the driver's record is read for
    for name, key in (("direct", "direct"), ("push", "push"), ("nccl_overlap", "nccl_overlap"), ("nccl_blocking", "nccl_blocking")):
        verdicts = [v[key] for v in out.values() if key in v]
        flat[name] = "MISMATCH" if "MISMATCH" in verdicts else ("unavailable" if "unavailable" in verdicts else "bit-exact")
    return flat, bad


def time_steps(D, sim, steps, cells_total):
    """K steps between barriers, CUDA events on the library's stream, max over ranks -> (MLUPS, ms per step)"""
    D.barrier()
    ms = D.max(sim.step_timed(steps))
    D.barrier()
    return cells_total * steps / (ms * 1e-3) / 1e6, ms / steps


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    thermal = args.workload == "thermal"
    if args.size == 0:
        args.size = {"lid": 768, "thermal": 512 if world == 1 else 256, "jacobi": 512, "particles": 0, "lid2d": 8192,
                     "thermal2d": 8192 if args.variant == "mpi" else 0, "lid_aa": 896}[args.workload]
    if args.workload in ("jacobi", "particles", "lid2d", "thermal2d", "lid_aa"):
        import bench_workloads
        getattr(bench_workloads, args.workload)(args, rank, local_rank, world)
        return
    if args.impl == "reference":
        run_reference(args, rank, max(world, args.gpus))       # the config of the N-GPU product arm, with or without a launcher
        return
    bytes_per_cell = BYTES_PER_CELL_THERMAL if thermal else BYTES_PER_CELL

    import numpy as np

    import mglc_b200 as mg
    from mglc_b200 import _lib as L

    D = Dist(rank, local_rank, world)
    torch = D.torch
    lib = L.lib()
    comm = D.communicator(mg)

    # ---- parity first (N > 1): a fast wrong answer is not a result ----
    parity = None
    if world > 1 and not args.no_parity:
        parity, bad = run_parity(D, comm)
        if bad:
            if rank == 0:
                print(json.dumps({"metric": "MLUPS", "value": None, "n_gpus": world, "parity": parity,
                                  "error": "decomposed run does not match the oracle; nothing was timed"}), flush=True)
            comm.close(); D.close()
            sys.exit(3)

    # ---- lattice ----
    free, _ = torch.cuda.mem_get_info()
    dims, per_gpu, gn = lattice_for(args, world)
    reduced = False
    while lattice_bytes(max(per_gpu), thermal) > free and args.size > 64:
        args.size -= 64
        dims, per_gpu, gn = lattice_for(args, world)
        reduced = True
    Driver = mg.BuoyancyDrivenCavity if thermal else mg.LidDrivenCavity

    def make_sim(total):
        kw = dict(arith=args.arith, device=local_rank, dims=dims)
        return Driver(total, comm=comm, **kw) if comm else Driver(total, **kw)

    sim = make_sim(gn)
    sub = sim.ranks[0]
    if args.no_overlap:
        args.halo = "blocking"
    avail = C.c_int()
    L.check(lib.mglc_lbm_direct_halo(sub._h, C.byref(avail)))
    if args.halo == "auto":                 # what the library picked for this block size (push for large blocks, in-kernel stores for small ones)
        mode = C.c_int()
        L.check(lib.mglc_lbm_get_overlap(sub._h, C.byref(mode)))
        args.halo = {v: k for k, v in HALO_MODE.items()}[mode.value] if avail.value else "overlap"
    if world > 1:
        L.check(lib.mglc_lbm_set_overlap(sub._h, HALO_MODE[args.halo]))
    cells_local = int(np.prod(sub.n))
    cells_total = int(np.prod(gn))
    sim.initial()

    # ---- device-resident throughput ----
    sim.step(max(args.warmup, 3))
    sim.sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = sim.launch_count()
    D.barrier()
    t0 = time.perf_counter()
    ms = sim.step_timed(args.steps)
    D.barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    launches = D.sum(float(sim.launch_count() - launches0))
    ms_min = D.min(ms)
    ms = D.max(ms)
    value = cells_total * args.steps / (ms * 1e-3) / 1e6

    # ---- roofline of the fused kernel: per-launch CUDA events on the launching stream ----
    L.check(lib.mglc_lbm_set_profiling(sub._h, 1))
    sim.step(min(args.steps, 40) + 1)
    fms, fl = C.c_float(), C.c_longlong()
    L.check(lib.mglc_lbm_kernel_time(sub._h, C.byref(fms), C.byref(fl)))
    L.check(lib.mglc_lbm_set_profiling(sub._h, 0))
    D.barrier()
    peak, peak_src = hbm_peak()
    avg_ms = D.max(fms.value / max(1, fl.value))
    achieved = bytes_per_cell * cells_local / (avg_ms * 1e-3) / 1e9
    kname = "k_th_fused" if thermal else "k_fused"
    tr = ncu_traffic(kname, cells_local) if world == 1 else None
    roofline = {"bound": "hbm", "kernel": f"mglc::{args.arith}::{kname}" + ("<PEER>" if world > 1 and args.halo == "direct" else ""),
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": (tr or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                "bytes_per_cell": bytes_per_cell, "cells_per_launch": cells_local,
                "avg_launch_ms": round(avg_ms, 4), "launches_timed": fl.value,
                "traffic_source": (tr or {}).get("source")}
    if thermal:
        # SURVEY 8(d) quotes the thermal update at 416 B/cell = the populations alone.  The kernel also has to carry the body
        # force of the previous collision across the streaming step (macro() adds F/2, B3:998-1002): 24 B in + 24 B out.
        roofline["frac_vs_416_B_per_cell"] = round(416 * cells_local / (avg_ms * 1e-3) / 1e9 / peak, 4)
        roofline["bytes_per_cell_note"] = "464 = 416 (26 populations in and out) + 48 (Fx,Fy,Fz of the previous collision in, of this one out)"

    # ---- the other halo transports on the same lattice (N > 1) ----
    transports = None
    if world > 1 and not args.no_extras:
        transports = {args.halo: {"value": round(value, 1), "ms_per_step": round(ms / args.steps, 4)}}
        for name in ("direct", "push", "blocking", "overlap"):
            if name == args.halo or (name in ("direct", "push") and not avail.value):
                continue
            L.check(lib.mglc_lbm_set_overlap(sub._h, HALO_MODE[name]))
            sim.step(3); sim.sync()
            v, m = time_steps(D, sim, max(5, args.steps // 2), cells_total)
            transports[name] = {"value": round(v, 1), "ms_per_step": round(m, 4)}
        L.check(lib.mglc_lbm_set_overlap(sub._h, HALO_MODE[args.halo]))
        sim.step(2); sim.sync()

    # ---- end to end through the C ABI with host arrays ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, D, sim, sub, lib, L, cells_local, cells_total, thermal)

    sim.close()
    D.barrier()          # a lattice mapped by the neighbours (CUDA IPC) is only returned to the driver once they have closed it too

    # ---- strong scaling in the same invocation (N > 1, default weak run): the config-3 lattice split over the GPUs ----
    strong = None
    if world > 1 and args.scaling == "weak" and not args.no_extras:
        gs = (args.size,) * 3 if not thermal else (512,) * 3
        sim2 = make_sim(gs)
        if avail.value and args.halo in ("direct", "push"):
            mode = C.c_int()
            L.check(lib.mglc_lbm_get_overlap(sim2.ranks[0]._h, C.byref(mode)))
            strong_halo = {v: k for k, v in HALO_MODE.items()}[mode.value]
        else:
            L.check(lib.mglc_lbm_set_overlap(sim2.ranks[0]._h, HALO_MODE[args.halo]))
            strong_halo = args.halo
        sim2.initial()
        sim2.step(max(args.warmup, 3)); sim2.sync()
        cs = int(np.prod(gs))
        v, m = time_steps(D, sim2, args.steps, cs)
        strong = {"global_lattice": list(gs), "per_gpu": list(sim2.ranks[0].n), "halo": strong_halo,
                  "value": round(v, 1), "unit": "MLUPS", "ms_per_step": round(m, 4), "steps": args.steps,
                  "note": "same lattice as the N=1 line of this bench: speed-up = value / the N=1 value"}
        sim2.close()

    if comm:
        comm.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu = cpu_baseline(args.cpu_seconds, thermal)
        except Exception as ex:   # the baseline is a reported extra, never the product
            cpu = {"value": None, "unit": "MLUPS", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}

    if rank == 0:
        out = {
            "metric": "MLUPS", "value": round(value, 1), "unit": "MLUPS", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": lattice_config(args, thermal, dims, per_gpu, gn),
            "detail": {"arith": args.arith, "storage": "SoA fp64, ping-pong, 1-cell halo",
                       "halo_exchange": "none (1 subdomain)" if world == 1 else HALO_NAMES[args.halo],
                       "reduced_to_fit": reduced, "wall_ms_per_step": round(wall_ms / args.steps, 4),
                       "ms_per_step_fastest_rank": round(ms_min / args.steps, 4)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": int(launches),
        }
        if parity is not None:
            out["parity"] = parity
        if transports is not None:
            out["transports"] = transports
        if strong is not None:
            out["strong"] = strong
        print(json.dumps(out), flush=True)
    D.close()


def run_e2e(args, D, sim, sub, lib, L, cells_local, cells_total, thermal):
    """upload (pinned host -> device) + K steps + check() + download rho,u,v,w(,T), all timed."""
    world = D.world
    nup = (19 + 4 + (7 + 1 if thermal else 0)) * cells_local * 8
    ndown = ((5 if thermal else 4) * cells_local * 8 + 16)
    avail = host_mem_available()
    if avail and nup * world > 0.8 * avail:
        return {"value": None, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "note": f"skipped: host arrays need {nup * world / 1e9:.0f} GB pinned, {avail / 1e9:.0f} GB available"}
    ptrs = []

    def pinned(count):
        p = C.c_void_p()
        L.check(lib.mglc_host_alloc(C.byref(p), C.c_size_t(count * 8)))
        ptrs.append(p)
        return p

    try:
        t_pin = time.perf_counter()
        pf = pinned(19 * cells_local)
        fields = [pinned(cells_local) for _ in range(4)]
        pg = pinned(7 * cells_local) if thermal else None
        pT = pinned(cells_local) if thermal else None
        t_pin = time.perf_counter() - t_pin
        # the host arrays a driver would own: the current device state, downloaded once (untimed)
        L.check(lib.mglc_lbm_download_f(sub._h, pf))
        L.check(lib.mglc_lbm_download_macro(sub._h, *fields))
        if thermal:
            L.check(lib.mglc_lbm_download_thermal(sub._h, pg, pT, None, None, None))
        D.barrier()
        t0 = time.perf_counter()
        L.check(lib.mglc_lbm_upload(sub._h, pf, *fields))
        if thermal:
            L.check(lib.mglc_lbm_upload_thermal(sub._h, pg, pT, None, None, None))
        t_up = time.perf_counter() - t0
        sim.step(args.steps)
        err = sim.check()
        t_run = time.perf_counter() - t0 - t_up
        L.check(lib.mglc_lbm_download_macro(sub._h, *fields))
        if thermal:
            L.check(lib.mglc_lbm_download_thermal(sub._h, None, pT, None, None, None))
        t_down = time.perf_counter() - t0 - t_up - t_run
        D.barrier()
        dt = D.max(time.perf_counter() - t0)
        t_up, t_down = D.max(t_up), D.max(t_down)
    finally:
        for p in ptrs:
            lib.mglc_host_free(p)
    return {"value": round(cells_total * args.steps / dt / 1e6, 1), "unit": "MLUPS",
            "h2d_bytes_per_step": int(nup * world / args.steps), "d2h_bytes_per_step": int(ndown * world / args.steps),
            "seconds": round(dt, 3), "errorU": err if not isinstance(err, tuple) else list(err),
            "h2d_gbs_per_gpu": round(nup / t_up / 1e9, 1), "d2h_gbs_per_gpu": round(ndown / t_down / 1e9, 1),
            "seconds_upload_steps_download": [round(t_up, 3), round(t_run, 3), round(t_down, 3)], "pin_seconds": round(t_pin, 1),
            "region": f"upload f,rho,u,v,w{',g,T' if thermal else ''} from pinned host + {args.steps} steps + check() + download rho,u,v,w{',T' if thermal else ''}"}


if __name__ == "__main__":
    main()
