"""The secondary workloads of bench.py (`--workload jacobi | particles | lid2d | thermal2d | lid_aa`); the headline lid and the
3-D thermal workloads live in bench.py itself.  Every function prints ONE JSON line in the shape of bench.py's.

jacobi: BASELINE.json config 2 (3-D Jacobi, 512^3 per GPU block, halo exchange over NCCL at N > 1).  One step = one
exchange_message + one jacobi sweep (LAP:94-103); the metric is million cell updates per second, algorithmic traffic 16 B/cell
(one read, one write; the source term is identically zero in the reference problem and is not read)."""
import ctypes as C
import json
import os
import sys
import tempfile
import time


def jacobi(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist

    import bench as B
    import mglc_b200 as mg

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

    n = args.size
    dims = mg.dims_create_nd(world, 3)
    gn = tuple(n * d for d in dims) if args.scaling == "weak" else (n, n, n)
    comm = None
    if world > 1:
        def bcast(b):
            box = [b]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        comm = mg.Communicator(world, rank, local_rank, bcast)
    sim = mg.Jacobi(gn, comm=comm) if comm else mg.Jacobi(gn, device=local_rank)
    cells_local = int(np.prod(sim.info[0]["n"]))
    cells_total = int(np.prod(gn))
    sim.init()
    sim.step(max(args.warmup, 3))
    sim.sync()
    sampler = B.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = sim.launch_count()
    barrier()
    ms = reduce_max(sim.step_timed(args.steps))
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = sim.launch_count() - l0
    value = cells_total * args.steps / (ms * 1e-3) / 1e6
    peak, peak_src = B.hbm_peak()
    achieved = 16 * cells_local * args.steps / (ms * 1e-3) / 1e9
    # end to end: host arrays in, `steps` iterations, check_diff, host array out
    e2e = None
    if not args.no_e2e and world == 1:
        A = sim.download(0)
        barrier()
        t0 = time.perf_counter()
        sim.upload(0, A=A, A_new=A)
        sim.step(args.steps)
        err = sim.check_diff()
        A = sim.download(0)
        barrier()
        dt = time.perf_counter() - t0
        e2e = {"value": round(cells_total * args.steps / dt / 1e6, 1), "unit": "Mcells/s", "h2d_bytes_per_step": int(2 * A.nbytes / args.steps),
               "d2h_bytes_per_step": int(A.nbytes / args.steps), "seconds": round(dt, 3), "check_diff": err,
               "region": f"upload A, A_new (pageable host) + {args.steps} iterations + check_diff + download A"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as orc
        os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
        wd = orc.JacobiWorld((256, 256, 256), 1)
        wd.init()
        wd.step(2)
        t0 = time.perf_counter(); wd.step(40); dt = time.perf_counter() - t0
        cpu = {"value": round(256 ** 3 * 40 / dt / 1e6, 1), "unit": "Mcells/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"256^3 x 40 iterations ({dt:.1f} s), oracle/jacobi.c (-O2, OpenMP)"}
        wd.close()
    sim.close()
    if rank == 0:
        print(json.dumps({
            "metric": "Mcells/s", "value": round(value, 1), "unit": "Mcells/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"jacobi3d_7pt_{n}^3_per_gpu", "global_grid": list(gn), "decomposition": "x".join(map(str, dims)),
                       "l2": "two %.1f GB arrays exceed the 126 MB L2; no flush needed" % (cells_local * 8 / 1e9)},
            "roofline": {"bound": "hbm", "kernel": "k_jacobi3d", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src, "bytes_per_cell": 16,
                         "note": "whole step (exchange + sweep) timed, not the kernel alone"},
            "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": int(launches)}), flush=True)
    if comm:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


def particles(args, rank, local_rank, world):
    """bench.py --workload particles: the reference's shipped particle problem (201 x 801 nodes, 64 particles, D2Q9;
    BASELINE.json config 5 in the reference's own 2-D form).  The whole state is L2-resident (1.4 MB per population
    set), so a step is bounded by kernel-launch latency, not HBM: the line reports MLUPS and microseconds per step."""
    import numpy as np
    import torch

    import bench as B
    import mglc_b200 as mg

    if world > 1:
        if rank == 0:
            print(json.dumps({"metric": "MLUPS", "value": None, "note": "particles bench runs on one GPU (161k nodes)"}))
        return
    torch.cuda.set_device(local_rank)
    rng = np.random.default_rng(11)
    xs, ys, tx, ty = [], [], 25.0, 25.0
    for _ in range(64):
        xs.append(tx + (rng.random() - 0.5) * 20); ys.append(ty + (rng.random() - 0.5) * 20)
        tx += 50.0
        if tx > 200.0:
            tx, ty = 25.0, ty + 50.0
    sim = mg.ParticleChannel(xs, ys, device=local_rank)
    sim.initial()
    steps = max(args.steps, 200)
    sim.step(max(args.warmup, 3)); sim.sync()
    l0 = sim.launch_count()
    sampler = B.ClockSampler(local_rank); sampler.start()
    ms = sim.step_timed(steps)
    clocks = sampler.stop()
    launches = sim.launch_count() - l0
    cells = 201 * 801
    cpu = None
    if not args.no_cpu:
        from oracle import oracle as orc
        wd = orc.ParticleWorld(xs, ys, nprocs=1)
        wd.initial(); wd.step(2)
        t0 = time.perf_counter(); wd.step(40); dt = time.perf_counter() - t0
        cpu = {"value": round(cells * 40 / dt / 1e6, 2), "unit": "MLUPS", "cores": 1, "kind": "port",
               "sample": f"201x801, 64 particles, 40 steps ({dt:.1f} s), oracle/particles2d.c"}
        wd.close()
    flags = sim.error_flags()
    sim.close()
    print(json.dumps({
        "metric": "MLUPS", "value": round(cells * steps / (ms * 1e-3) / 1e6, 1), "unit": "MLUPS", "n_gpus": 1, "steps": steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms / steps, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": {"workload": "micro_particles_d2q9_201x801_64_particles", "error_flags": flags,
                                                         "l2": "state is L2-resident by design of the reference problem (161k nodes)"},
        "roofline": {"bound": "launch latency", "achieved": None, "peak": None, "unit": None, "frac": None, "traffic": None,
                     "launches_per_step": round(launches / steps, 2), "us_per_step": round(ms / steps * 1e3, 2)},
        "cpu_baseline": cpu, "e2e": None, "clocks": clocks, "gpu_launches": int(launches)}), flush=True)


def lid2d(args, rank, local_rank, world):
    """bench.py --workload lid2d: the reference's 2-D D2Q9 lid-driven cavity on one GPU, lattice far larger than
    L2 (default 8192 x 8192: 4.8 GB per population set).  Roofline: 144 B/cell = 9 loads + 9 stores of fp64 per fused launch.
    cpu_baseline: the reference's OWN compiled C program (oracle/_ref/liblid2d_ref.so, 200 x 200 as shipped, one thread)."""
    import torch

    import bench as B
    import mglc_b200 as mg

    if world > 1:
        if rank == 0:
            print(json.dumps({"metric": "MLUPS", "value": None, "note": "lid2d bench runs on one GPU"}))
        return
    torch.cuda.set_device(local_rank)
    n = args.size or 8192
    # MGLC_BENCH_L2D_VARIANT = c | f | i | s picks another of the reference's programs (default: the Fortran + MPI one)
    l2d_variant = os.environ.get("MGLC_BENCH_L2D_VARIANT", "f")
    sim = mg.LidDrivenCavity2D((n, n), variant=l2d_variant, strict=args.arith == "strict", device=local_rank)
    sim.initial()
    sim.step(max(args.warmup, 3)); sim.sync()
    if n * n <= (1 << 20):
        # an L2-resident lattice replays its fused launches from CUDA graphs of 64 kernels, one per ping-pong index: have both
        # instantiated before the timed region, and time enough steps for a stable figure
        sim.step(66); sim.step(66); sim.sync()
        args.steps = max(args.steps, 4000)
    l0 = sim.launch_count()
    sampler = B.ClockSampler(local_rank); sampler.start()
    ms = sim.step_timed(args.steps)
    clocks = sampler.stop()
    launches = sim.launch_count() - l0
    err = sim.check()
    sim.close()
    cells = n * n
    peak, peak_src = B.hbm_peak()
    # a step(K) call is collision + (K-1) fused launches + stream/macro: all three move 144-176 B/cell; report the whole step
    achieved = 144.0 * cells * args.steps / (ms * 1e-3) / 1e9
    cpu = None
    ref_so = os.path.join(B.ROOT, "oracle", "_ref", "liblid2d_ref.so")
    if not args.no_cpu:
        from oracle import oracle as orc
        if os.path.exists(ref_so):
            cwd = os.getcwd()
            tmp = tempfile.mkdtemp()
            os.chdir(tmp)                  # the reference program writes its output files into the working directory
            sys.stdout.flush()
            saved, null = os.dup(1), os.open(os.devnull, os.O_WRONLY)
            os.dup2(null, 1)               # ... and prints a banner: keep this process's stdout to the one JSON line
            try:
                ref = orc.RefLid2D(ref_so)
                ref.lib.initial(); ref.step(20)
                t0 = time.perf_counter(); ref.step(2000); dt = time.perf_counter() - t0
                C.CDLL(None).fflush(None)
            finally:
                os.dup2(saved, 1); os.close(saved); os.close(null)
                os.chdir(cwd)
            cpu = {"value": round(200 * 200 * 2000 / dt / 1e6, 2), "unit": "MLUPS", "cores": 1, "kind": "reference",
                   "sample": f"MPI/Lid_driven_cavity/c/lid_driven_cavity.c as shipped (200x200), 2000 steps ({dt:.1f} s), gcc -O2"}
        else:
            wd = orc.Lid2DWorld((1024, 1024), 1)
            wd.initial(); wd.step(2)
            t0 = time.perf_counter(); wd.step(20); dt = time.perf_counter() - t0
            wd.close()
            cpu = {"value": round(1024 * 1024 * 20 / dt / 1e6, 2), "unit": "MLUPS", "cores": 1, "kind": "port",
                   "sample": f"1024x1024, 20 steps ({dt:.1f} s), oracle/lid2d.c"}
    print(json.dumps({
        "metric": "MLUPS", "value": round(cells * args.steps / (ms * 1e-3) / 1e6, 1), "unit": "MLUPS", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 5), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"lid_driven_cavity_d2q9_mrt_{n}x{n}" + ("" if l2d_variant == "f" else f"_variant_{l2d_variant}"), "Re": 1000.0, "U0": 0.1, "arith": args.arith, "errorU": err,
                   "l2": ("lattice (2 x %.1f GB) far exceeds the 126 MB L2; no flush needed" % (9 * cells * 8 / 1e9)) if cells > (1 << 22) else
                         ("lattice (2 x %.1f MB) is L2-resident: the rate is not an HBM figure; fused launches replayed from CUDA graphs" % (9 * cells * 8 / 1e6))},
        "roofline": {"bound": "hbm" if cells > (1 << 22) else "launch latency / L2", "kernel": f"mglc::{args.arith}::k_l2_fused", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src, "bytes_per_cell": 144,
                     "cells_per_launch": cells, "note": "whole step(K) call timed: collision + (K-1) fused + stream/macro launches"},
        "cpu_baseline": cpu, "e2e": None, "clocks": clocks, "gpu_launches": int(launches)}), flush=True)


def thermal2d(args, rank, local_rank, world):
    """bench.py --workload thermal2d: the reference's 2-D thermal D2Q9 + D2Q5 driver (Buoyancy_driven_cavity/fortran/2d, SURVEY 8f
    row 4) on one GPU, lattice far larger than L2 (default 8192 x 8192).  Roofline: 240 B/cell = (9 + 5) loads + (9 + 5) stores of
    fp64 + the carried force Fy (8 B in, 8 B out) per fused launch.  cpu_baseline: oracle/thermal2d.c on all host threads."""
    import torch

    import bench as B
    import mglc_b200 as mg

    if world > 1:
        if rank == 0:
            print(json.dumps({"metric": "MLUPS", "value": None, "note": "thermal2d bench runs on one GPU"}))
        return
    torch.cuda.set_device(local_rank)
    acc = getattr(args, "variant", "mpi") == "acc"
    if acc:
        # the OpenACC program (the reference's only GPU code): shipped 513 x 257 at Ra = 1e5, or 2n+1 x n+1 at a Rayleigh number that
        # keeps paraA inside (-4, 1) with lengthUnit = nx
        n = args.size
        nx, ny = (513, 257) if not n else (2 * n + 1, n + 1)
        Ra = 1e5 if nx <= 2049 else 1e9
        sim = mg.BuoyancyDrivenCavity2D((nx, ny), variant="acc", strict=args.arith == "strict", device=local_rank, Rayleigh=Ra)
    else:
        n = args.size or 8192
        nx = ny = n
        # Ra = 1e7 (shipped) keeps paraA inside (-4, 1) only up to ~6600 cells per side at Ma = 0.1 (initial.F90:30 stops otherwise);
        # the large lattices run Ra = 1e9
        Ra = 1e7 if n <= 4096 else 1e9
        sim = mg.BuoyancyDrivenCavity2D((n, n), strict=args.arith == "strict", device=local_rank, Rayleigh=Ra)
    sim.initial()
    sim.step(max(args.warmup, 3)); sim.sync()
    if nx * ny <= (1 << 20):
        # an L2-resident lattice replays its fused launches from CUDA graphs of 64 kernels, one per ping-pong index: have both
        # instantiated before the timed region, and time enough steps for a stable figure
        sim.step(66); sim.step(66); sim.sync()
        args.steps = max(args.steps, 4000)
    l0 = sim.launch_count()
    sampler = B.ClockSampler(local_rank); sampler.start()
    ms = sim.step_timed(args.steps)
    clocks = sampler.stop()
    launches = sim.launch_count() - l0
    eu, et = sim.check()
    nure = sim.calNuRe()
    sim.close()
    cells = nx * ny
    peak, peak_src = B.hbm_peak()
    # a step(K) call is 2 collision launches + (K-1) fused launches + stream/macro; report the whole call against 240 B/cell
    achieved = 240.0 * cells * args.steps / (ms * 1e-3) / 1e9
    cpu = None
    if not args.no_cpu:
        from oracle import oracle as orc
        threads = os.cpu_count() or 1
        os.environ.setdefault("OMP_NUM_THREADS", str(threads))
        wd = orc.Thermal2DWorld((2048, 2048), 1)
        wd.initial(); wd.step(1)
        t0 = time.perf_counter(); wd.step(2); t1 = (time.perf_counter() - t0) / 2
        k = max(2, min(200, int(args.cpu_seconds / max(t1, 1e-3))))
        t0 = time.perf_counter(); wd.step(k); dt = time.perf_counter() - t0
        wd.close()
        cpu = {"value": round(2048 * 2048 * k / dt / 1e6, 2), "unit": "MLUPS", "cores": threads, "kind": "port",
               "sample": f"2048x2048, {k} steps ({dt:.1f} s), oracle/thermal2d.c (-O2, OpenMP over rows, no FMA contraction)"}
    print(json.dumps({
        "metric": "MLUPS", "value": round(cells * args.steps / (ms * 1e-3) / 1e6, 1), "unit": "MLUPS", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 5), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"buoyancy_driven_cavity_d2q9_d2q5_mrt_{nx}x{ny}" + ("_openacc_program" if acc else ""), "Ra": Ra, "Pr": 0.71,
                   "Ma": 0.1, "bc": "Rayleigh-Benard plates, periodic vertical walls (seq/bouyancy2d_acc.F90)" if acc else "side-heated",
                   "arith": args.arith, "errorU": eu, "errorT": et, "NuVolAvg": nure[1], "ReVolAvg": nure[2],
                   "l2": ("lattice (2 x %.1f GB) far exceeds the 126 MB L2; no flush needed" % (14 * cells * 8 / 1e9)) if cells >= 4_000_000 else
                         ("lattice (2 x %.1f MB) is L2-resident at the program's shipped size: the rate is not an HBM figure" % (14 * cells * 8 / 1e6))},
        "roofline": {"bound": "hbm" if cells >= 4_000_000 else "launch latency / L2", "kernel": f"mglc::{args.arith}::k_t2_fused", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src, "bytes_per_cell": 240,
                     "cells_per_launch": cells, "note": "whole step(K) call timed: 2 collision + (K-1) fused + stream/macro launches"},
        "cpu_baseline": cpu, "e2e": None, "clocks": clocks, "gpu_launches": int(launches)}), flush=True)


def lid_aa(args, rank, local_rank, world):
    """bench.py --workload lid_aa: the D3Q19 lid-driven cavity on ONE lattice (AA-pattern storage) at a size the
    two-lattice path cannot hold on one B200 (default 896^3 = 1.59x the cells of 768^3; 114 GB of lattice + 23 GB of fields).
    Roofline: 304 B/cell per launch, as for the ping-pong kernel."""
    import torch

    import bench as B
    import mglc_b200 as mg

    if world > 1:
        if rank == 0:
            print(json.dumps({"metric": "MLUPS", "value": None, "note": "lid_aa bench runs on one GPU (the AA path is single-subdomain)"}))
        return
    torch.cuda.set_device(local_rank)
    n = args.size or 896
    free, _ = torch.cuda.mem_get_info()
    reduced = False
    need = lambda e: 19 * (((e + 17 + 15) // 16) * 16) * (e + 2) ** 2 * 8 + 4 * e ** 3 * 8 + (1 << 28)
    while need(n) > free and n > 64:
        n -= 64; reduced = True
    sim = mg.LidDrivenCavityAA((n, n, n), arith=args.arith, device=local_rank)
    sim.initial()
    sim.step(max(args.warmup, 3)); sim.sync()
    l0 = sim.launch_count()
    sampler = B.ClockSampler(local_rank); sampler.start()
    ms = sim.step_timed(args.steps)
    clocks = sampler.stop()
    launches = sim.launch_count() - l0
    nbytes = sim.device_bytes()
    m = sim.download_macro()
    mass = float(m["rho"].sum()) / n ** 3
    umax = float(abs(m["u"]).max())
    sim.close()
    cells = n ** 3
    peak, peak_src = B.hbm_peak()
    achieved = 304.0 * cells * args.steps / (ms * 1e-3) / 1e9
    print(json.dumps({
        "metric": "MLUPS", "value": round(cells * args.steps / (ms * 1e-3) / 1e6, 1), "unit": "MLUPS", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"lid_driven_cavity_d3q19_mrt_{n}x{n}x{n}_single_lattice", "Re": 1000.0, "U0": 0.1, "arith": args.arith,
                   "storage": "SoA fp64, ONE lattice updated in place (AA pattern)", "device_bytes": nbytes, "reduced_to_fit": reduced,
                   "mean_rho": mass, "max_u": umax,
                   "l2": "lattice (%.1f GB) far exceeds the 126 MB L2; no flush needed" % (19 * cells * 8 / 1e9)},
        "roofline": {"bound": "hbm", "kernel": f"mglc::{args.arith}::k_aa_odd / k_aa_even", "achieved": round(achieved, 1), "peak": peak,
                     "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src, "bytes_per_cell": 304,
                     "cells_per_launch": cells,
                     "note": "whole step(K) call timed: the launches alternate k_aa_odd / k_aa_even, plus collision and macro at its ends"},
        "cpu_baseline": None, "e2e": None, "clocks": clocks, "gpu_launches": int(launches)}), flush=True)
