"""bench.py --workload jacobi: BASELINE.json config 2 (3-D Jacobi, 512^3 per GPU block, halo exchange over
NCCL at N > 1).  One step = one exchange_message + one jacobi sweep (LAP:94-103).  Prints ONE JSON line in the
shape of bench.py's; the metric is million cell updates per second, algorithmic traffic 16 B/cell (one read,
one write; the source term is identically zero in the reference problem and is not read)."""
import json
import os
import time


def main(args, rank, local_rank, world):
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
