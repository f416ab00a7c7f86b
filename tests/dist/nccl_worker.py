"""One rank of the multi-GPU parity run (launched by tests/test_multigpu.py through torch.distributed.run, one
process per GPU, halos over NCCL).  Every rank steps its own subdomain through libmglc.so; rank 0 gathers the fields
and compares them with the single-rank CPU oracle.  Prints 'MULTIGPU OK' on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mglc_b200 as mg  # noqa: E402
from mglc_b200 import _lib as L  # noqa: E402


def gather_blocks(block, rank, world):
    """all ranks' (start, array) pairs on rank 0"""
    out = [None] * world
    dist.gather_object(block, out if rank == 0 else None, dst=0)
    return out


def assemble(blocks, total, lead=()):
    glob = np.empty(lead + tuple(total), order="F")
    for start, arr in blocks:
        sl = tuple(slice(s, s + n) for s, n in zip(start, arr.shape[len(lead):]))
        glob[(slice(None),) * len(lead) + sl] = arr
    return glob


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def bcast(b):
        box = [b]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    comm = mg.Communicator(world, rank, local, bcast)
    if rank == 0:
        from oracle import oracle as orc
    ok = True

    # ---- lid-driven cavity, uneven blocks, strict arithmetic: bit-exact in every halo mode:
    #      2 = direct stores into the neighbours' halos (CUDA IPC mappings), 1 = overlapped NCCL exchange, 0 = blocking ----
    total, nsteps = (41, 37, 35), 12
    import ctypes as C
    for overlap in (2, 1, 0):
        sim = mg.LidDrivenCavity(total, comm=comm, arith="strict")
        if overlap == 2:
            avail = C.c_int()
            L.check(L.lib().mglc_lbm_direct_halo(sim.ranks[0]._h, C.byref(avail)))
            if rank == 0:
                print(f"direct halo mappings: {'established' if avail.value else 'UNAVAILABLE (NCCL transport)'}")
            if not avail.value:
                ok = ok and bool(os.environ.get("MGLC_NO_DIRECT"))     # on an NVLink box the mappings must come up
                sim.close()
                continue
        L.check(L.lib().mglc_lbm_set_overlap(sim.ranks[0]._h, overlap))
        sim.initial()
        sim.step(5); sim.step(nsteps - 5)            # two calls: the rotated state carries the in-flight exchange across
        S = sim.ranks[0]
        m = S.download_macro()
        blocks = {k: gather_blocks((S.start, m[k]), rank, world) for k in ("rho", "u", "v", "w")}
        fb = gather_blocks((S.start, S.download_f()), rank, world)
        err = sim.check()
        if rank == 0:
            wd = orc.LidWorld(total, 1)
            wd.initial(); wd.step(nsteps)
            for k in blocks:
                same = np.array_equal(assemble(blocks[k], total), wd.gather(k))
                ok &= same
                print(f"lid overlap={overlap} {k}: {'bit-exact' if same else 'MISMATCH'}")
            same = np.array_equal(assemble(fb, total, (19,)), wd.gather("f"))
            ok &= same
            print(f"lid overlap={overlap} f: {'bit-exact' if same else 'MISMATCH'}; errorU {err} vs {wd.check()}")
            ok &= bool(np.isclose(err, wd.errorU if hasattr(wd, 'errorU') else err, rtol=1e-12))
            wd.close()
        sim.close()

    # ---- thermal cavity ----
    total, nsteps = (27, 25, 23), 10
    for overlap in (2, 1, 0):
        sim = mg.BuoyancyDrivenCavity(total, comm=comm, arith="strict")
        if overlap == 2:
            avail = C.c_int()
            L.check(L.lib().mglc_lbm_direct_halo(sim.ranks[0]._h, C.byref(avail)))
            if not avail.value:
                sim.close()
                continue
        L.check(L.lib().mglc_lbm_set_overlap(sim.ranks[0]._h, overlap))
        sim.initial()
        sim.step(4); sim.step(nsteps - 4)
        S = sim.ranks[0]
        m = S.download_macro(); th = S.download_thermal(with_g=False); m["T"] = th["T"]
        blocks = {k: gather_blocks((S.start, m[k]), rank, world) for k in ("rho", "u", "v", "w", "T")}
        if rank == 0:
            wd = orc.ThermalWorld(total, 1)
            wd.initial(); wd.step(nsteps)
            for k in blocks:
                same = np.array_equal(assemble(blocks[k], total), wd.gather(k))
                ok &= same
                print(f"thermal overlap={overlap} {k}: {'bit-exact' if same else 'MISMATCH'}")
            wd.close()
        sim.close()

    # ---- Jacobi 3-D and 2-D ----
    for total in ((37, 29, 23), (61, 45)):
        sim = mg.Jacobi(total, comm=comm)
        sim.init(); sim.step(15)
        diff = sim.check_diff()
        inf = sim.info[0]
        inner = tuple(slice(1, n + 1) for n in inf["n"])
        blocks = gather_blocks((inf["start"], sim.download(0)[inner]), rank, world)
        if rank == 0:
            wd = orc.JacobiWorld(total, 1)
            wd.init(); wd.step(15)
            same = np.array_equal(assemble(blocks, total), wd.gather()) and diff == wd.check_diff()
            ok &= same
            print(f"jacobi {len(total)}-D: {'bit-exact' if same else 'MISMATCH'}")
            wd.close()
        sim.close()

    # ---- particles crossing the subdomain boundary ----
    px, py, params = [20.3, 41.2], [60.0, 33.7], dict(total_nx=61, total_ny=90)
    sim = mg.ParticleChannel(px, py, comm=comm, **params)
    sim.initial(); sim.step(100)
    inf = sim.info[0]
    st = sim.download(0, ("rho", "u", "v"))
    blocks = {k: gather_blocks((inf["start"], st[k]), rank, world) for k in ("rho", "u", "v")}
    p = sim.particles()
    flags = sim.error_flags()
    if rank == 0:
        wd = orc.ParticleWorld(px, py, nprocs=1, **params)
        wd.initial(); wd.step(100)
        for k in blocks:
            a, b = assemble(blocks[k], wd.total), wd.gather(k)
            good = np.abs(a - b).max() <= 1e-10 and np.linalg.norm((a - b).ravel()) <= 1e-12 * max(np.linalg.norm(b.ravel()), 0.02 * np.sqrt(a.size))
            ok &= bool(good)
            print(f"particles {k}: max|diff| {np.abs(a - b).max():.2e} {'ok' if good else 'MISMATCH'}")
        for k in ("xCenter", "yCenter", "Uc", "Vc", "rationalOmega"):
            good = np.allclose(p[k], getattr(wd, k), rtol=1e-12, atol=1e-12)
            ok &= bool(good)
        ok &= flags == 0
        wd.close()
    sim.close()

    # ---- the 2-D drivers (2-D Cartesian blocks; faces + corner messages over NCCL): strict arithmetic, bit-exact ----
    total, nsteps = (45, 38), 25
    for name in ("lid2d", "thermal2d"):
        sim = mg.LidDrivenCavity2D(total, comm=comm, variant="f", strict=True) if name == "lid2d" else \
            mg.BuoyancyDrivenCavity2D(total, comm=comm, strict=True, Rayleigh=1e6)
        keys = ("rho", "u", "v") + (("T",) if name == "thermal2d" else ())
        sim.initial()
        sim.step(7); sim.step(nsteps - 7)
        inf = sim.info[0]
        blocks = {k: gather_blocks((inf["start"], sim.download(0, k)), rank, world) for k in keys}
        err = sim.check()
        if rank == 0:
            wd = orc.Lid2DWorld(total, 1, variant="f") if name == "lid2d" else orc.Thermal2DWorld(total, 1, Rayleigh=1e6)
            wd.initial(); wd.step(nsteps)
            for k in keys:
                same = np.array_equal(assemble(blocks[k], total), wd.gather(k))
                ok &= same
                print(f"{name} {k}: {'bit-exact' if same else 'MISMATCH'}")
            ok &= bool(np.allclose(err, wd.check(), rtol=1e-12))
            wd.close()
        sim.close()

    comm.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTIGPU OK" if ok else "MULTIGPU FAILED")
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
