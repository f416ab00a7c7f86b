#!/usr/bin/env python
"""bench.py -- throughput of the D3Q19 lid-driven-cavity hot path on B200 (BASELINE.json config 3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one lattice timestep (collision, halo exchange, streaming, bounce-back, macro; the
reference's loop body, L3/main.f90:85-97) over the whole lattice.  N=1 runs the 768^3 lattice the
metric is quoted on; N>1 runs one 768^3 block per GPU (weak scaling, 3-D Cartesian decomposition,
NCCL halo exchange), `--scaling strong` splits one 768^3 lattice instead.  Rank 0 prints ONE JSON line.

  value     MLUPS, whole job, lattice resident in HBM, CUDA-event time of K steps, max over ranks
  e2e       MLUPS through the C ABI with HOST arrays: upload f,rho,u,v,w from pinned host memory ->
            K steps -> check() -> download rho,u,v,w, all inside the timed region
  roofline  fused stream+collide kernel: 304 B/cell (19 loads + 19 stores of fp64) x cells per launch
            / its mean CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the CPU oracle (restated reference, OpenMP over all host threads) on a bounded sample

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
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="mglc", choices=["mglc", "reference"])
    p.add_argument("--no-overlap", action="store_true", help="same as --halo blocking")
    p.add_argument("--halo", default="auto", choices=["auto", "direct", "overlap", "blocking"],
                   help="multi-GPU halo transport of the fused step: direct = stores into the neighbours' halos over NVLink "
                        "(CUDA IPC), overlap = NCCL exchange beside the interior update, blocking = NCCL exchange, then update; "
                        "auto = direct when the mappings came up, else overlap")
    p.add_argument("--workload", default="lid", choices=["lid", "thermal", "jacobi", "particles", "lid2d", "thermal2d", "lid_aa"],
                   help="lid = BASELINE.json's metric (default); thermal / jacobi = the other configs, for profiles/")
    p.add_argument("--size", type=int, default=0, help="per-GPU block edge (weak) / global edge (strong); 0 = the workload's config size")
    p.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    p.add_argument("--arith", default="fast", choices=["fast", "strict"])
    p.add_argument("--variant", default="mpi", choices=["mpi", "acc"],
                   help="thermal2d only: mpi = mpi_blocked/ (side-heated), acc = the OpenACC program seq/bouyancy2d_acc.F90 "
                        "(Rayleigh-Benard plates, periodic vertical walls; --size 0 = its shipped 513 x 257)")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--cpu-seconds", type=float, default=12.0)
    return p.parse_args()


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(cells_per_launch):
    """dram bytes per fused launch from the committed ncu capture of the same launch size
    (profiles/ncu_traffic.json), or None when no capture of that size exists."""
    try:
        for cap in json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["captures"]:
            if cap["cells_per_launch"] == cells_per_launch:
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


# ------------------------------------------------------------------------------------------------------
def cpu_baseline(seconds, edge=256):
    """Oracle (restated reference, AoS, un-fused five sweeps) on all host threads; bounded sample."""
    import numpy as np  # noqa: F401
    from oracle import oracle as orc
    threads = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    try:
        orc.build(native=True)
        native = True
    except Exception:
        native = False
    total = (edge, edge, edge)
    wd = orc.LidWorld(total, 1, native=native)
    wd.initial()
    t0 = time.perf_counter(); wd.step(1); t1 = time.perf_counter() - t0
    n = max(1, min(200, int(seconds / max(t1, 1e-3))))
    t0 = time.perf_counter(); wd.step(n); dt = time.perf_counter() - t0
    wd.close()
    mlups = edge ** 3 * n / dt / 1e6
    return {"value": round(mlups, 2), "unit": "MLUPS", "cores": threads, "kind": "port",
            "sample": f"{edge}^3 lattice x {n} steps ({dt:.1f} s), oracle/lid3d.c "
                      f"({'-O3 -march=native' if native else '-O2'}, OpenMP, no FMA contraction)"}


def run_reference(args, rank):
    """--impl reference: the CPU restatement of the reference's path, all host threads, same metric/config."""
    if rank != 0:
        return
    from oracle import oracle as orc
    threads = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    try:
        orc.build(native=True); native = True
    except Exception:
        native = False
    n = args.size
    # calibrate on a thin slab, then size the per-step sample so K+W steps take ~90 s
    cal = orc.LidWorld((n, n, 8), 1, native=native)
    cal.initial()
    t0 = time.perf_counter(); cal.step(1); t1 = time.perf_counter() - t0
    cal.close()
    rate = n * n * 8 / t1
    nz = int(max(8, min(n, rate * 90.0 / max(1, args.steps + args.warmup) / (n * n))))
    # ... and keep the slab inside the host memory: f + halo'd f_post + seven fields ~ 400 B per cell
    avail = host_mem_available()
    if avail:
        nz = int(max(8, min(nz, 0.5 * avail / (400.0 * n * n))))
    wd = orc.LidWorld((n, n, nz), 1, native=native)
    wd.initial()
    wd.step(args.warmup)
    t0 = time.perf_counter(); wd.step(args.steps); dt = time.perf_counter() - t0
    wd.close()
    cells = n * n * nz
    mlups = cells * args.steps / dt / 1e6
    sample = f"{n}x{n}x{nz} slab of the {n}^3 lattice per step, {args.steps} steps in {dt:.1f} s"
    print(json.dumps({
        "impl": "reference", "metric": "MLUPS", "value": round(mlups, 2), "unit": "MLUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"lid_driven_cavity_d3q19_mrt_{n}x{n}x{n}_per_gpu", "Re": 1000.0, "U0": 0.1,
                   "note": "reference Fortran+MPI cannot be built in this image; this is its C restatement (oracle/)"},
        "cpu_baseline": {"value": round(mlups, 2), "unit": "MLUPS", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(mlups, 2), "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    thermal = args.workload == "thermal"
    if args.size == 0:
        args.size = {"lid": 768, "thermal": 512 if world == 1 else 256, "jacobi": 512, "particles": 0, "lid2d": 8192, "thermal2d": 8192 if args.variant == "mpi" else 0, "lid_aa": 896}[args.workload]
    if args.workload == "jacobi":
        import bench_workloads
        bench_workloads.jacobi(args, rank, local_rank, world)
        return
    if args.workload == "particles":
        import bench_workloads
        bench_workloads.particles(args, rank, local_rank, world)
        return
    if args.workload == "lid2d":
        import bench_workloads
        bench_workloads.lid2d(args, rank, local_rank, world)
        return
    if args.workload == "thermal2d":
        import bench_workloads
        bench_workloads.thermal2d(args, rank, local_rank, world)
        return
    if args.workload == "lid_aa":
        import bench_workloads
        bench_workloads.lid_aa(args, rank, local_rank, world)
        return
    if args.impl == "reference":
        if thermal:
            raise SystemExit("bench.py: --impl reference times the headline (lid) workload only")
        run_reference(args, rank)
        return
    bytes_per_cell = 464 if thermal else BYTES_PER_CELL      # thermal: (19+7) x 16 B + 48 B of carried force

    import numpy as np
    import torch
    import torch.distributed as dist

    import mglc_b200 as mg
    from mglc_b200 import _lib as L

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; mglc_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- lattice ----
    dims = mg.dims_create(world)
    n = args.size
    free, _ = torch.cuda.mem_get_info()
    per_gpu = [n, n, n] if args.scaling == "weak" else [n // d for d in dims]
    reduced = False
    while lattice_bytes(max(per_gpu), thermal) > free and n > 64:
        n -= 64
        per_gpu = [n, n, n] if args.scaling == "weak" else [n // d for d in dims]
        reduced = True
    gn = tuple(p * d for p, d in zip(per_gpu, dims))

    comm = None
    if world > 1:
        def bcast(b):
            box = [b]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        comm = mg.Communicator(world, rank, local_rank, bcast)
    Driver = mg.BuoyancyDrivenCavity if thermal else mg.LidDrivenCavity
    sim = Driver(gn, comm=comm, arith=args.arith, device=local_rank) if comm else \
        Driver(gn, arith=args.arith, device=local_rank)
    sub = sim.ranks[0]
    if args.no_overlap:
        args.halo = "blocking"
    avail = C.c_int()
    L.check(L.lib().mglc_lbm_direct_halo(sub._h, C.byref(avail)))
    if args.halo == "auto":
        args.halo = "direct" if avail.value else "overlap"
    if world > 1:
        L.check(L.lib().mglc_lbm_set_overlap(sub._h, {"direct": 2, "overlap": 1, "blocking": 0}[args.halo]))
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
    barrier()
    t0 = time.perf_counter()
    ms = sim.step_timed(args.steps)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    launches = reduce_sum(float(sim.launch_count() - launches0))
    ms = reduce_max(ms)
    value = cells_total * args.steps / (ms * 1e-3) / 1e6

    # ---- roofline of the fused kernel: per-launch CUDA events on the launching stream ----
    lib = L.lib()
    L.check(lib.mglc_lbm_set_profiling(sub._h, 1))
    sim.step(min(args.steps, 40) + 1)
    fms, fl = C.c_float(), C.c_longlong()
    L.check(lib.mglc_lbm_kernel_time(sub._h, C.byref(fms), C.byref(fl)))
    L.check(lib.mglc_lbm_set_profiling(sub._h, 0))
    barrier()
    peak, peak_src = hbm_peak()
    avg_ms = fms.value / max(1, fl.value)
    achieved = bytes_per_cell * cells_local / (avg_ms * 1e-3) / 1e9
    tr = None if thermal else ncu_traffic(cells_local)
    kname = "k_th_fused" if thermal else "k_fused"
    roofline = {"bound": "hbm", "kernel": f"mglc::{args.arith}::{kname}",
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": (tr or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                "bytes_per_cell": bytes_per_cell, "cells_per_launch": cells_local,
                "avg_launch_ms": round(avg_ms, 4), "launches_timed": fl.value,
                "traffic_source": (tr or {}).get("source")}

    # ---- end to end through the C ABI with host arrays ----
    e2e = None
    if not args.no_e2e and not thermal:
        e2e = run_e2e(args, sim, sub, lib, L, np, barrier, reduce_max, cells_local, cells_total, world)

    sim.close()
    if comm:
        comm.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and not thermal:
        try:
            cpu = cpu_baseline(args.cpu_seconds)
        except Exception as ex:   # the baseline is a reported extra, never the product
            cpu = {"value": None, "unit": "MLUPS", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}

    if rank == 0:
        out = {
            "metric": "MLUPS", "value": round(value, 1), "unit": "MLUPS", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": ("buoyancy_driven_cavity_d3q19_d3q7_mrt" if thermal else "lid_driven_cavity_d3q19_mrt") +
                                   f"_{per_gpu[0]}x{per_gpu[1]}x{per_gpu[2]}_per_gpu",
                       "global_lattice": list(gn), "decomposition": "x".join(map(str, dims)),
                       **({"Ra": 1e6, "Pr": 0.71, "Ma": 0.1, "Ek": 1e-3} if thermal else {"Re": 1000.0, "U0": 0.1}),
                       "arith": args.arith, "storage": "SoA fp64, ping-pong, 1-cell halo",
                       "halo_exchange": "none (1 subdomain)" if world == 1 else {
                           "direct": "fused kernel stores outgoing populations into the neighbours' halos over NVLink (CUDA IPC) + flag barrier",
                           "overlap": "NCCL send/recv on a second stream, overlapped with the interior update",
                           "blocking": "NCCL send/recv, blocking before the update"}[args.halo],
                       "l2": "lattice (2 x %.1f GB) far exceeds the 126 MB L2; no flush needed" % ((26 if thermal else 19) * cells_local * 8 / 1e9),
                       "reduced_to_fit": reduced, "wall_ms_per_step": round(wall_ms / args.steps, 4)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": int(launches),
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, sim, sub, lib, L, np, barrier, reduce_max, cells_local, cells_total, world):
    """upload (pinned host -> device) + K steps + check() + download rho,u,v,w, all timed."""
    need = (19 + 4) * cells_local * 8
    avail = host_mem_available()
    if avail and need * world > 0.6 * avail:
        return {"value": None, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "note": f"skipped: host arrays need {need * world / 1e9:.0f} GB pinned, {avail / 1e9:.0f} GB available"}
    ptrs = []

    def pinned(count):
        p = C.c_void_p()
        L.check(lib.mglc_host_alloc(C.byref(p), C.c_size_t(count * 8)))
        ptrs.append(p)
        return p

    try:
        pf = pinned(19 * cells_local)
        fields = [pinned(cells_local) for _ in range(4)]
        # the host arrays a driver would own: the current device state, downloaded once (untimed)
        L.check(lib.mglc_lbm_download_f(sub._h, pf))
        L.check(lib.mglc_lbm_download_macro(sub._h, *fields))
        barrier()
        t0 = time.perf_counter()
        L.check(lib.mglc_lbm_upload(sub._h, pf, *fields))
        sim.step(args.steps)
        err = sim.check()
        L.check(lib.mglc_lbm_download_macro(sub._h, *fields))
        barrier()
        dt = reduce_max(time.perf_counter() - t0)
    finally:
        for p in ptrs:
            lib.mglc_host_free(p)
    return {"value": round(cells_total * args.steps / dt / 1e6, 1), "unit": "MLUPS",
            "h2d_bytes_per_step": int((19 + 4) * cells_local * 8 * world / args.steps),
            "d2h_bytes_per_step": int((4 * cells_local * 8 + 16) * world / args.steps),
            "seconds": round(dt, 3), "errorU": err,
            "region": f"upload f,rho,u,v,w from pinned host + {args.steps} steps + check() + download rho,u,v,w"}


if __name__ == "__main__":
    main()
