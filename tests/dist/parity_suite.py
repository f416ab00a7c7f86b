"""Multi-process parity checks shared by tests/dist/nccl_worker.py (pytest -m multigpu) and bench.py's N > 1 runs (so that
the driver's scaling records carry a correctness verdict next to every throughput figure).

Every rank steps its own subdomain through libmglc.so in STRICT arithmetic; rank 0 gathers the blocks and compares them with
the single-rank CPU oracle (test infrastructure: the oracle is the checker here, never the thing measured).  The verdict of a
case is "bit-exact", "MISMATCH" or "unavailable" (direct halo stores need the CUDA IPC mappings of an NVLink box).

Transports of the fused step (mglc_lbm_set_overlap): 2 = direct stores into the neighbours' halos from inside the update
kernel, 3 = the same stores from one small launch after it ("push"), 1 = overlapped NCCL exchange
(lid3_mpi_nonblock.f90:1108-1230), 0 = blocking NCCL exchange (L3/ex_sendrecv.f90:12-123)."""
import ctypes as C
import os

import numpy as np
import torch.distributed as dist

import mglc_b200 as mg
from mglc_b200 import _lib as L

TRANSPORTS = {"direct": 2, "push": 3, "nccl_overlap": 1, "nccl_blocking": 0}


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


def _oracle():
    from oracle import oracle as orc
    return orc


def _direct_available(sim):
    avail = C.c_int()
    L.check(L.lib().mglc_lbm_direct_halo(sim.ranks[0]._h, C.byref(avail)))
    return bool(avail.value)


def lid(comm, rank, world, total=(41, 37, 35), nsteps=12, transports=("direct", "push", "nccl_overlap", "nccl_blocking"),
        reinit=True, log=None):
    """3-D lid-driven cavity, uneven blocks.  Per transport: step(5); step(n-5) (the rotated state carries the in-flight halos
    across the two calls) and, with reinit, step(3); initial(); step(n) (stale halo state after re-initialising)."""
    out = {}
    ref = None
    if rank == 0:
        orc = _oracle()
        wd = orc.LidWorld(total, 1)
        wd.initial(); wd.step(nsteps)
        ref = {k: wd.gather(k) for k in ("rho", "u", "v", "w", "f")}
        ref["errorU"] = wd.check()
        wd.close()
    for name in transports:
        sim = mg.LidDrivenCavity(total, comm=comm, arith="strict")
        if name in ("direct", "push") and not _direct_available(sim):
            out[name] = "unavailable"
            sim.close()
            continue
        L.check(L.lib().mglc_lbm_set_overlap(sim.ranks[0]._h, TRANSPORTS[name]))
        same = True
        for scenario in (("split",) + (("reinit",) if reinit else ())):
            sim.initial()
            if scenario == "split":
                sim.step(5); sim.step(nsteps - 5)
            else:
                sim.step(3); sim.initial(); sim.step(nsteps)
            S = sim.ranks[0]
            m = S.download_macro()
            blocks = {k: gather_blocks((S.start, m[k]), rank, world) for k in ("rho", "u", "v", "w")}
            fb = gather_blocks((S.start, S.download_f()), rank, world)
            err = sim.check()
            if rank == 0:
                for k in blocks:
                    ok = np.array_equal(assemble(blocks[k], total), ref[k])
                    same &= ok
                    if log:
                        log(f"lid {name} {scenario} {k}: {'bit-exact' if ok else 'MISMATCH'}")
                ok = np.array_equal(assemble(fb, total, (19,)), ref["f"])
                same &= ok
                if scenario == "split":
                    same &= bool(np.isclose(err, ref["errorU"], rtol=1e-12))
                if log:
                    log(f"lid {name} {scenario} f: {'bit-exact' if ok else 'MISMATCH'}; errorU {err} vs {ref['errorU']}")
        out[name] = "bit-exact" if same else "MISMATCH"
        sim.close()
    return out


def lid_aa(comm, rank, world, total=(41, 37, 35), nsteps=12, log=None):
    """the same cavity on ONE lattice per block (AA-pattern storage, mglc_aa_create_comm): the blocks store into each other's
    lattices through the CUDA IPC mappings, one neighbour barrier per launch.  step(5); step(n-5) crosses both layouts; then
    initial() again in the middle of a run, as for the two-lattice transports"""
    ref = None
    if rank == 0:
        orc = _oracle()
        wd = orc.LidWorld(total, 1)
        wd.initial(); wd.step(nsteps)
        ref = {k: wd.gather(k) for k in ("rho", "u", "v", "w", "f")}
        ref["errorU"] = wd.check()
        wd.close()
    try:
        sim = mg.LidDrivenCavityAA(total, comm=comm, arith="strict")
    except L.MglcError as e:           # collective verdict inside the library: every rank raises or none does
        if log:
            log(f"lid_aa: unavailable ({e})")
        return {"single_lattice": "unavailable"}
    same = True
    for scenario in ("split", "reinit"):
        sim.initial()
        if scenario == "split":
            sim.step(5); sim.step(nsteps - 5)
        else:
            sim.step(3); sim.initial(); sim.step(nsteps)
        m = sim.download_macro()
        blocks = {k: gather_blocks((sim.start, m[k]), rank, world) for k in ("rho", "u", "v", "w")}
        fb = gather_blocks((sim.start, sim.download_f()), rank, world)
        err = sim.check()
        if rank == 0:
            for k in blocks:
                ok = np.array_equal(assemble(blocks[k], total), ref[k])
                same &= ok
                if log:
                    log(f"lid_aa {scenario} {k}: {'bit-exact' if ok else 'MISMATCH'}")
            ok = np.array_equal(assemble(fb, total, (19,)), ref["f"])
            same &= ok
            if scenario == "split":
                same &= bool(np.isclose(err, ref["errorU"], rtol=1e-12))
            if log:
                log(f"lid_aa {scenario} f: {'bit-exact' if ok else 'MISMATCH'}; errorU {err} vs {ref['errorU']}")
    sim.close()
    return {"single_lattice": "bit-exact" if same else "MISMATCH"}


def thermal(comm, rank, world, total=(27, 25, 23), nsteps=10, transports=("direct", "push", "nccl_overlap", "nccl_blocking"), log=None):
    """3-D thermal cavity (f + g exchange), uneven blocks"""
    out = {}
    ref = None
    if rank == 0:
        orc = _oracle()
        wd = orc.ThermalWorld(total, 1)
        wd.initial(); wd.step(nsteps)
        ref = {k: wd.gather(k) for k in ("rho", "u", "v", "w", "T")}
        wd.close()
    for name in transports:
        sim = mg.BuoyancyDrivenCavity(total, comm=comm, arith="strict")
        if name in ("direct", "push") and not _direct_available(sim):
            out[name] = "unavailable"
            sim.close()
            continue
        L.check(L.lib().mglc_lbm_set_overlap(sim.ranks[0]._h, TRANSPORTS[name]))
        sim.initial()
        sim.step(4); sim.step(nsteps - 4)
        S = sim.ranks[0]
        m = S.download_macro(); th = S.download_thermal(with_g=False); m["T"] = th["T"]
        blocks = {k: gather_blocks((S.start, m[k]), rank, world) for k in ("rho", "u", "v", "w", "T")}
        same = True
        if rank == 0:
            for k in blocks:
                ok = np.array_equal(assemble(blocks[k], total), ref[k])
                same &= ok
                if log:
                    log(f"thermal {name} {k}: {'bit-exact' if ok else 'MISMATCH'}")
        out[name] = "bit-exact" if same else "MISMATCH"
        sim.close()
    return out


def jacobi(comm, rank, world, totals=((37, 29, 23), (61, 45)), nsteps=15, log=None):
    """Jacobi halo-exchange path, 3-D and 2-D, with both halo transports of the fused step: direct stores into the neighbours'
    ghost layers, and exchange_message over NCCL followed by the sweep (LAP:94-103 as written)"""
    out = {}
    for total in totals:
        for halo in ("direct", "exchange"):
            sim = mg.Jacobi(total, comm=comm)
            key = f"{len(total)}d_{halo}"
            if halo == "direct" and not sim.direct_halo_available():
                out[key] = "unavailable"
                sim.close()
                continue
            sim.set_halo(halo)
            sim.init(); sim.step(7); sim.step(nsteps - 7)
            diff = sim.check_diff()
            inf = sim.info[0]
            inner = tuple(slice(1, n + 1) for n in inf["n"])
            blocks = gather_blocks((inf["start"], sim.download(0)[inner]), rank, world)
            same = True
            if rank == 0:
                orc = _oracle()
                wd = orc.JacobiWorld(total, 1)
                wd.init(); wd.step(nsteps)
                same = bool(np.array_equal(assemble(blocks, total), wd.gather()) and diff == wd.check_diff())
                wd.close()
                if log:
                    log(f"jacobi {len(total)}-D {halo}: {'bit-exact' if same else 'MISMATCH'}")
            out[key] = "bit-exact" if same else "MISMATCH"
            sim.sync()
            sim.close()
    return out


def particles(comm, rank, world, nsteps=100, log=None):
    """two particles settling through the subdomain boundary (the reference's DKT pair, scaled down); fields within the
    north_star tolerance, trajectories to rounding"""
    px, py, params = [20.3, 41.2], [60.0, 33.7], dict(total_nx=61, total_ny=90)
    sim = mg.ParticleChannel(px, py, comm=comm, **params)
    sim.initial(); sim.step(nsteps)
    inf = sim.info[0]
    st = sim.download(0, ("rho", "u", "v"))
    blocks = {k: gather_blocks((inf["start"], st[k]), rank, world) for k in ("rho", "u", "v")}
    p = sim.particles()
    flags = sim.error_flags()
    ok = True
    if rank == 0:
        orc = _oracle()
        wd = orc.ParticleWorld(px, py, nprocs=1, **params)
        wd.initial(); wd.step(nsteps)
        for k in blocks:
            a, b = assemble(blocks[k], wd.total), wd.gather(k)
            good = np.abs(a - b).max() <= 1e-10 and np.linalg.norm((a - b).ravel()) <= 1e-12 * max(np.linalg.norm(b.ravel()), 0.02 * np.sqrt(a.size))
            ok &= bool(good)
            if log:
                log(f"particles {k}: max|diff| {np.abs(a - b).max():.2e} {'ok' if good else 'MISMATCH'}")
        for k in ("xCenter", "yCenter", "Uc", "Vc", "rationalOmega"):
            ok &= bool(np.allclose(p[k], getattr(wd, k), rtol=1e-12, atol=1e-12))
        ok &= flags == 0
        wd.close()
    sim.close()
    return {"fields<=1e-12,trajectories<=1e-12": "ok" if ok else "MISMATCH"}


def drivers_2d(comm, rank, world, total=(45, 38), nsteps=25, log=None):
    """the 2-D drivers (2-D Cartesian blocks; faces + corner messages over NCCL): strict arithmetic, bit-exact"""
    out = {}
    for name in ("lid2d", "thermal2d"):
        sim = mg.LidDrivenCavity2D(total, comm=comm, variant="f", strict=True) if name == "lid2d" else \
            mg.BuoyancyDrivenCavity2D(total, comm=comm, strict=True, Rayleigh=1e6)
        keys = ("rho", "u", "v") + (("T",) if name == "thermal2d" else ())
        sim.initial()
        sim.step(7); sim.step(nsteps - 7)
        inf = sim.info[0]
        blocks = {k: gather_blocks((inf["start"], sim.download(0, k)), rank, world) for k in keys}
        err = sim.check()
        same = True
        if rank == 0:
            orc = _oracle()
            wd = orc.Lid2DWorld(total, 1, variant="f") if name == "lid2d" else orc.Thermal2DWorld(total, 1, Rayleigh=1e6)
            wd.initial(); wd.step(nsteps)
            for k in keys:
                ok = np.array_equal(assemble(blocks[k], total), wd.gather(k))
                same &= ok
                if log:
                    log(f"{name} {k}: {'bit-exact' if ok else 'MISMATCH'}")
            same &= bool(np.allclose(err, wd.check(), rtol=1e-12))
            wd.close()
        out[name] = "bit-exact" if same else "MISMATCH"
        sim.close()
    return out


def failed(verdicts):
    """True if any verdict in a (nested) dict of verdicts is a mismatch"""
    for v in verdicts.values():
        if isinstance(v, dict):
            if failed(v):
                return True
        elif v == "MISMATCH":
            return True
    return False
