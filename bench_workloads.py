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


def jacobi_config(n, dims, gn, cells_local):
    return {"workload": f"jacobi3d_7pt_{n}^3_per_gpu", "global_grid": list(gn), "decomposition": "x".join(map(str, dims)),
            "l2": "two %.1f GB arrays per GPU exceed the 126 MB L2; no flush needed" % (cells_local * 8 / 1e9)}


def jacobi_reference(args, rank, world):
    """--impl reference --workload jacobi: the CPU restatement of jacobi2d_mpi.f90's loop in 3-D (oracle/jacobi.c, OpenMP) on a
    bounded slab of one block, all host threads"""
    if rank != 0:
        return
    import bench as B
    from oracle import oracle as orc
    threads = orc.set_threads(0)
    n = args.size
    dims = tuple(sorted(B.dims_create(world)))           # z first, as bench.lattice_for
    gn = tuple(n * d for d in dims) if args.scaling == "weak" else (n, n, n)
    per = [n, n, n] if args.scaling == "weak" else [n // d for d in dims]
    nz = max(8, min(per[2], 64))
    wd = orc.JacobiWorld((per[0], per[1], nz), 1)
    wd.init(); wd.step(max(1, args.warmup))
    t0 = time.perf_counter(); wd.step(args.steps); dt = time.perf_counter() - t0
    wd.close()
    v = per[0] * per[1] * nz * args.steps / dt / 1e6
    sample = f"{per[0]}x{per[1]}x{nz} slab of one block per iteration, {args.steps} iterations in {dt:.1f} s, oracle/jacobi.c on {threads} OpenMP threads"
    print(json.dumps({
        "impl": "reference", "metric": "Mcells/s", "value": round(v, 1), "unit": "Mcells/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": jacobi_config(n, dims, gn, per[0] * per[1] * per[2]),
        "cpu_baseline": {"value": round(v, 1), "unit": "Mcells/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(v, 1), "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def jacobi(args, rank, local_rank, world):
    if args.impl == "reference":
        return jacobi_reference(args, rank, max(world, args.gpus))
    import numpy as np

    import bench as B
    import mglc_b200 as mg

    D = B.Dist(rank, local_rank, world)
    n = args.size
    dims = tuple(sorted(mg.dims_create_nd(world, 3)))    # MPI_Dims_create's factors, assigned z first (see bench.lattice_for)
    gn = tuple(n * d for d in dims) if args.scaling == "weak" else (n, n, n)
    comm = D.communicator(mg)
    parity = None
    if world > 1 and not args.no_parity:
        sys.path.insert(0, os.path.join(B.ROOT, "tests", "dist"))
        import parity_suite as ps
        parity = ps.jacobi(comm, rank, world)
        if D.max(1.0 if (rank == 0 and ps.failed(parity)) else 0.0) > 0:
            if rank == 0:
                print(json.dumps({"metric": "Mcells/s", "value": None, "n_gpus": world, "parity": parity,
                                  "error": "decomposed run does not match the oracle; nothing was timed"}), flush=True)
            comm.close(); D.close()
            sys.exit(3)
    sim = mg.Jacobi(gn, comm=comm, dims=dims) if comm else mg.Jacobi(gn, device=local_rank)
    cells_local = int(np.prod(sim.info[0]["n"]))
    cells_total = int(np.prod(gn))
    sim.init()
    sim.step(max(args.warmup, 3))
    sim.sync()
    sampler = B.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = sim.launch_count()
    D.barrier()
    ms = D.max(sim.step_timed(args.steps))
    D.barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = D.sum(float(sim.launch_count() - l0))
    value = cells_total * args.steps / (ms * 1e-3) / 1e6
    peak, peak_src = B.hbm_peak()
    achieved = 16 * cells_local * args.steps / (ms * 1e-3) / 1e9
    kernel = "k_jacobi3d_tma" if os.environ.get("MGLC_JACOBI_KERNEL") == "tma" else "k_jacobi3d<0,4>"
    tr = B.ncu_traffic(kernel, cells_local)
    # end to end: host arrays in, `steps` iterations, check_diff, host array out
    e2e = None
    if not args.no_e2e:
        A = sim.download(0)
        D.barrier()
        t0 = time.perf_counter()
        sim.upload(0, A=A, A_new=A)
        sim.step(args.steps)
        err = sim.check_diff()
        A = sim.download(0)
        D.barrier()
        dt = D.max(time.perf_counter() - t0)
        e2e = {"value": round(cells_total * args.steps / dt / 1e6, 1), "unit": "Mcells/s", "h2d_bytes_per_step": int(2 * A.nbytes * world / args.steps),
               "d2h_bytes_per_step": int(A.nbytes * world / args.steps), "seconds": round(dt, 3), "check_diff": err,
               "region": f"upload A, A_new (pageable host) + {args.steps} iterations + check_diff + download A"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as orc
        threads = orc.set_threads(0)
        wd = orc.JacobiWorld((256, 256, 256), 1)
        wd.init()
        wd.step(2)
        t0 = time.perf_counter(); wd.step(40); dt = time.perf_counter() - t0
        cpu = {"value": round(256 ** 3 * 40 / dt / 1e6, 1), "unit": "Mcells/s", "cores": threads, "kind": "port",
               "sample": f"256^3 x 40 iterations ({dt:.1f} s), oracle/jacobi.c (-O2, OpenMP)"}
        wd.close()
    halo = sim.halo_mode() if hasattr(sim, "halo_mode") else None
    sim.close()
    if rank == 0:
        out = {
            "metric": "Mcells/s", "value": round(value, 1), "unit": "Mcells/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": jacobi_config(n, dims, gn, cells_local),
            "detail": {"kernel": kernel, "halo_exchange": halo},
            "roofline": {"bound": "hbm", "kernel": kernel, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": (tr or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                         "bytes_per_cell": 16, "cells_per_launch": cells_local, "traffic_source": (tr or {}).get("source"),
                         "note": "whole step (exchange + sweep) timed, not the kernel alone"},
            "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": int(launches)}
        if parity is not None:
            out["parity"] = parity
        print(json.dumps(out), flush=True)
    if comm:
        comm.close()
    D.close()


def particle_raster(nx, ny, seed=11):
    """the reference's seeding (P4/initial.F90:49-75: a raster with 50 nodes between centres, each centre moved by up to +-10),
    extended over a lattice of any size; the jitter comes from our own seeded generator (random_number is compiler-specific)"""
    import numpy as np
    rng = np.random.default_rng(seed)
    gx, gy = np.meshgrid(np.arange(25.0, nx - 20, 50.0), np.arange(25.0, ny - 20, 50.0), indexing="ij")
    xs = (gx + (rng.random(gx.shape) - 0.5) * 20).ravel(order="F")
    ys = (gy + (rng.random(gy.shape) - 0.5) * 20).ravel(order="F")
    return xs, ys


def particles_config(edge, world, total, nparticles):
    return {"workload": f"micro_particles_d2q9_{edge}x{edge}_per_gpu" if edge else "micro_particles_d2q9_201x801_64_particles",
            "global_lattice": list(total), "decomposition": f"1x{world}" if edge else "1x1", "particles": int(nparticles),
            "radius": 10.0, "seeding": "50-node raster, +-10 jitter (P4/initial.F90:49-75)",
            "l2": ("lattice (2 x %.1f GB per GPU) far exceeds the 126 MB L2; no flush needed" % (9 * edge * edge * 8 / 1e9)) if edge else
                  "state is L2-resident by design of the reference problem (161k nodes)"}


def particles_reference(args, rank, world):
    """--impl reference --workload particles: oracle/particles2d.c (the restated P4 loop body) on a bounded lattice with the same
    particle density; one thread for the particle loops, OpenMP for the node sweeps"""
    if rank != 0:
        return
    from oracle import oracle as orc
    threads = orc.set_threads(0)
    edge = args.size
    total = (edge, edge * world) if edge else (201, 801)
    nx, ny = (601, 1201) if edge else (201, 801)
    xs, ys = particle_raster(nx, ny)
    if not edge:
        xs, ys = xs[:64], ys[:64]
    wd = orc.ParticleWorld(xs, ys, nprocs=1, total_nx=nx, total_ny=ny)
    wd.initial(); wd.step(max(1, args.warmup))
    t0 = time.perf_counter(); wd.step(args.steps); dt = time.perf_counter() - t0
    wd.close()
    v = nx * ny * args.steps / dt / 1e6
    npart = len(particle_raster(*total)[0]) if edge else 64
    sample = f"{nx}x{ny} nodes, {len(xs)} particles (same density), {args.steps} steps in {dt:.1f} s, oracle/particles2d.c on {threads} OpenMP threads"
    print(json.dumps({
        "impl": "reference", "metric": "MLUPS", "value": round(v, 2), "unit": "MLUPS", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": particles_config(edge, world, total, npart),
        "cpu_baseline": {"value": round(v, 2), "unit": "MLUPS", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(v, 2), "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def particles(args, rank, local_rank, world):
    """bench.py --workload particles: BASELINE.json config 5 in the reference's own 2-D form (P4/main.F90:35-73).
    --size 0 (one GPU): the shipped problem, 201 x 801 nodes and 64 particles -- L2-resident, bounded by launch latency.
    --size E: an E x E block per GPU (default 8192: 67 M nodes, 26 732 particles per GPU on the reference's raster), blocks
    stacked along y (1 x N, what MPI_Dims_create_2d picks for this shape), particles crossing the block boundaries, halos and
    the three per-step reductions over NCCL.  Roofline: 144 B/node of populations + 8 B of mask = 152 B/node/step."""
    if args.impl == "reference":
        return particles_reference(args, rank, max(world, args.gpus))
    import numpy as np

    import bench as B
    import mglc_b200 as mg

    D = B.Dist(rank, local_rank, world)
    comm = D.communicator(mg)
    edge = args.size if (args.size or world == 1) else 8192
    parity = None
    if world > 1 and not args.no_parity:
        sys.path.insert(0, os.path.join(B.ROOT, "tests", "dist"))
        import parity_suite as ps
        parity = ps.particles(comm, rank, world)
        if D.max(1.0 if (rank == 0 and ps.failed(parity)) else 0.0) > 0:
            if rank == 0:
                print(json.dumps({"metric": "MLUPS", "value": None, "n_gpus": world, "parity": parity,
                                  "error": "decomposed run does not match the oracle; nothing was timed"}), flush=True)
            comm.close(); D.close()
            sys.exit(3)
    if edge:
        total = (edge, edge * world)
        xs, ys = particle_raster(*total)
    else:
        total = (201, 801)
        xs, ys = particle_raster(*total)
        xs, ys = xs[:64], ys[:64]
    kw = dict(total_nx=total[0], total_ny=total[1])
    sim = mg.ParticleChannel(xs, ys, comm=comm, dims=(1, world), **kw) if comm else mg.ParticleChannel(xs, ys, device=local_rank, **kw)
    sim.initial()
    steps = args.steps if edge else max(args.steps, 200)
    sim.step(max(args.warmup, 3)); sim.sync()
    l0 = sim.launch_count()
    sampler = B.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    D.barrier()
    ms = D.max(sim.step_timed(steps))
    D.barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = sim.launch_count() - l0
    cells = total[0] * total[1]
    cells_local = int(np.prod(sim.info[0]["n"]))
    flags = int(D.max(float(sim.error_flags())))
    # end to end: the host arrays of a restart (f, f_post, rho, u, v, obst) in, K steps, fields and particle state out
    e2e = None
    if not args.no_e2e and edge:
        st = sim.download(0)
        D.barrier()
        t0 = time.perf_counter()
        sim.upload(0, **st)
        sim.step(steps)
        out = sim.download(0, ("rho", "u", "v"))
        pstate = sim.particles()
        D.barrier()
        dt = D.max(time.perf_counter() - t0)
        up = sum(a.nbytes for a in st.values()); down = sum(a.nbytes for a in out.values()) + 8 * 8 * len(xs)
        e2e = {"value": round(cells * steps / dt / 1e6, 1), "unit": "MLUPS", "h2d_bytes_per_step": int(up * world / steps),
               "d2h_bytes_per_step": int(down * world / steps), "seconds": round(dt, 3), "mean_yCenter": float(pstate["yCenter"].mean()),
               "region": f"upload f,f_post,rho,u,v,obst (pageable host) + {steps} steps + download rho,u,v + particle state"}
    p = sim.particles()
    sim.close()
    peak, peak_src = B.hbm_peak()
    achieved = 152.0 * cells_local * steps / (ms * 1e-3) / 1e9
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as orc
        threads = orc.set_threads(0)
        cn = (601, 1201) if edge else (201, 801)
        cx, cy = particle_raster(*cn)
        if not edge:
            cx, cy = cx[:64], cy[:64]
        wd = orc.ParticleWorld(cx, cy, nprocs=1, total_nx=cn[0], total_ny=cn[1])
        wd.initial(); wd.step(2)
        t0 = time.perf_counter(); wd.step(20); dt = time.perf_counter() - t0
        cpu = {"value": round(cn[0] * cn[1] * 20 / dt / 1e6, 2), "unit": "MLUPS", "cores": threads, "kind": "port",
               "sample": f"{cn[0]}x{cn[1]} nodes, {len(cx)} particles (same density), 20 steps ({dt:.1f} s), oracle/particles2d.c"}
        wd.close()
    if rank == 0:
        out = {
            "metric": "MLUPS", "value": round(cells * steps / (ms * 1e-3) / 1e6, 1), "unit": "MLUPS", "n_gpus": world, "steps": steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms / steps, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": particles_config(edge, world, total, len(xs)),
            "detail": {"error_flags": flags, "launches_per_step": round(launches / steps, 2), "mean_settling_velocity": float(p["Vc"].mean()),
                       "collectives_per_step": 0 if world == 1 else 3,
                       "collectives": "rhoAvg sums (2 doubles) -> link force sums (3 x N doubles) -> rhoAvg sums (2 doubles): each needs the previous one's result"},
            "roofline": ({"bound": "hbm", "kernel": "k_p_collision_sum + k_p_update + k_p_links + k_p_mask_sum + k_p_refill", "achieved": round(achieved, 1),
                          "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src, "bytes_per_cell": 152,
                          "cells_per_launch": cells_local, "note": "whole step timed; the reference's two persistent lattices f and f_post are each read and written once per step by two node sweeps"}
                         if edge else
                         {"bound": "launch latency", "achieved": None, "peak": None, "unit": None, "frac": None, "traffic": None,
                          "launches_per_step": round(launches / steps, 2), "us_per_step": round(ms / steps * 1e3, 2)}),
            "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": int(launches)}
        if parity is not None:
            out["parity"] = parity
        print(json.dumps(out), flush=True)
    if comm:
        comm.close()
    D.close()


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
    """bench.py --workload lid_aa: the D3Q19 lid-driven cavity on ONE lattice per block (AA-pattern storage) at a
    size the two-lattice path cannot hold on a B200 (default 896^3 per GPU = 1.59x the cells of 768^3; 114 GB of lattice +
    23 GB of fields).  N > 1: one block per GPU (weak), the blocks store into each other's lattices (mglc_aa_create_comm), parity
    against the oracle in the same run.  Roofline: 304 B/cell per launch, as for the two-lattice kernel."""
    import torch

    import bench as B
    import mglc_b200 as mg

    D = B.Dist(rank, local_rank, world)
    comm = D.communicator(mg)
    n = args.size or 896
    free = D.min(float(torch.cuda.mem_get_info()[0]))
    reduced = False
    need = lambda e: 19 * (((e + 17 + 15) // 16) * 16) * (e + 2) ** 2 * 8 + 4 * e ** 3 * 8 + (1 << 28)
    while need(n) > free and n > 64:
        n -= 64; reduced = True
    dims = tuple(int(x) for x in args.dims.split(",")) if args.dims else tuple(sorted(B.dims_create(world)))
    gn = tuple(n * d for d in dims)
    parity = None
    if comm is not None and not args.no_parity:
        sys.path.insert(0, os.path.join(B.ROOT, "tests", "dist"))
        import parity_suite as ps
        parity = ps.lid_aa(comm, rank, world)
        if D.max(1.0 if (rank == 0 and ps.failed(parity)) else 0.0) > 0:
            if rank == 0:
                print(json.dumps({"metric": "MLUPS", "value": None, "parity": parity, "error": "parity mismatch against the oracle"}), flush=True)
            comm.close(); D.close()
            sys.exit(3)
        parity = dict(parity, case="lid 41x37x35, 12 steps, strict, vs oracle/lid3d.c on one emulated rank")
    sim = mg.LidDrivenCavityAA(gn, arith=args.arith, comm=comm, dims=dims) if comm else mg.LidDrivenCavityAA(gn, arith=args.arith, device=local_rank)
    sim.initial()
    sim.step(max(args.warmup, 3)); sim.sync()
    l0 = sim.launch_count()
    sampler = B.ClockSampler(local_rank); sampler.start()
    D.barrier()
    ms = D.max(sim.step_timed(args.steps))
    D.barrier()
    clocks = sampler.stop()
    launches = D.sum(float(sim.launch_count() - l0))
    nbytes = sim.device_bytes()
    m = sim.download_macro()
    cells_local = n ** 3
    mass = D.sum(float(m["rho"].sum())) / (cells_local * world)
    umax = D.max(float(abs(m["u"]).max()))
    sim.close()
    D.barrier()
    if comm:
        comm.close()
    cells = cells_local * world
    peak, peak_src = B.hbm_peak()
    achieved = 304.0 * cells_local * args.steps / (ms * 1e-3) / 1e9
    if rank == 0:
        print(json.dumps({
            "metric": "MLUPS", "value": round(cells * args.steps / (ms * 1e-3) / 1e6, 1), "unit": "MLUPS", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"lid_driven_cavity_d3q19_mrt_{n}x{n}x{n}_per_gpu_single_lattice", "global_lattice": list(gn),
                       "decomposition": "x".join(map(str, dims)), "Re": 1000.0, "U0": 0.1, "arith": args.arith,
                       "storage": "SoA fp64, ONE lattice per block updated in place (AA pattern)", "device_bytes_per_gpu": nbytes,
                       "reduced_to_fit": reduced, "mean_rho": mass, "max_u": umax,
                       "halo": None if world == 1 else "none packed: every launch stores into the neighbours' lattices (CUDA IPC), one flag barrier per launch",
                       "l2": "lattice (%.1f GB per GPU) far exceeds the 126 MB L2; no flush needed" % (19 * cells_local * 8 / 1e9)},
            "parity": parity,
            "roofline": {"bound": "hbm", "kernel": f"mglc::{args.arith}::k_aa_odd / k_aa_even", "achieved": round(achieved, 1), "peak": peak,
                         "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src, "bytes_per_cell": 304,
                         "cells_per_launch": cells_local,
                         "note": "per GPU; whole step(K) call timed: the launches alternate k_aa_odd / k_aa_even, plus collision and macro at its ends"},
            "cpu_baseline": None, "e2e": None, "clocks": clocks, "gpu_launches": int(launches)}), flush=True)
    D.close()
