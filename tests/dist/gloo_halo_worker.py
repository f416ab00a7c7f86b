"""One rank of the CPU (gloo) multi-process test of the halo plumbing: every process owns ONE subdomain of the
lid-driven cavity, packs its outgoing populations exactly as include/mglc.h's mglc_halo_plan prescribes
(message order, neighbours, counts, buffer layout [slot][t2][t1]), moves them with torch.distributed send/recv
the way the library moves them with ncclSend/ncclRecv, and unpacks.  The result must equal the reference
semantics: the in-process P-rank oracle's message_passing_sendrecv(), bit for bit.  Also checks the NCCL-id
broadcast plumbing of mg.Communicator up to the point where a GPU is needed."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mglc_b200 as mg  # noqa: E402
from mglc_b200 import _lib as L  # noqa: E402
from oracle import oracle as orc  # noqa: E402

EX, EY, EZ = orc.EX, orc.EY, orc.EZ
FACE_POPS = [[1, 7, 9, 11, 13], [2, 8, 10, 12, 14], [3, 7, 8, 15, 17], [4, 9, 10, 16, 18], [5, 11, 12, 15, 16], [6, 13, 14, 17, 18]]


def region(n, d, recv):
    """index tuple into f_post (19, nx+2, ny+2, nz+2) of message `d`'s slab: source layer or destination halo"""
    nx, ny, nz = n
    if d < 6:
        axis, plus = d >> 1, not (d & 1)
        nfix = n[axis]
        fix = (0 if plus else nfix + 1) if recv else (nfix if plus else 1)
        sl = [slice(1, nx + 1), slice(1, ny + 1), slice(1, nz + 1)]
        sl[axis] = fix
        return FACE_POPS[d], tuple(sl)
    e = (EX[d], EY[d], EZ[d])
    sl = []
    for q in range(3):
        if e[q] == 0:
            sl.append(slice(1, n[q] + 1))
        elif e[q] > 0:
            sl.append(0 if recv else n[q])
        else:
            sl.append(n[q] + 1 if recv else 1)
    return [d], tuple(sl)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    total = (11, 9, 7)
    dims = None if len(sys.argv) < 2 else tuple(int(x) for x in sys.argv[1].split("x"))
    desc = mg.make_desc(total, world, rank, dims)
    plan = mg.halo_plan(desc)
    n = tuple(desc.ln)

    # the same seeded global state on every process; mine is the block the reference decomposition gives me
    wd = orc.LidWorld(total, world, dims=dims)
    rng = np.random.default_rng(77)
    for R in wd.ranks:
        R.f_post[...] = rng.random(R.f_post.shape)
    mine = wd.ranks[rank].f_post.copy(order="F")
    assert wd.ranks[rank].n == n and wd.ranks[rank].start == tuple(desc.start)

    # pack -> send/recv -> unpack, message by message in plan order (both sides walk the same list, so pairs match up)
    reqs, recvs = [], []
    for m in plan:
        if m["send_count"]:
            pops, sl = region(n, m["dir"], recv=False)
            buf = np.stack([np.asarray(mine[(a,) + sl]).T.ravel() if m["dir"] < 6 else np.asarray(mine[(a,) + sl]).ravel() for a in pops])
            assert buf.size == m["send_count"]
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(buf.ravel())), dst=m["send_to"], tag=m["dir"]))
        if m["recv_count"]:
            t = torch.empty(m["recv_count"], dtype=torch.float64)
            reqs.append(dist.irecv(t, src=m["recv_from"], tag=m["dir"]))
            recvs.append((m, t))
    for r in reqs:
        r.wait()
    for m, t in recvs:
        pops, sl = region(n, m["dir"], recv=True)
        per = m["recv_count"] // len(pops)
        for q, a in enumerate(pops):
            chunk = t.numpy()[q * per:(q + 1) * per]
            tgt = mine[(a,) + sl]
            mine[(a,) + sl] = chunk.reshape(tgt.shape[::-1]).T if m["dir"] < 6 else chunk.reshape(tgt.shape)

    wd.message_passing_sendrecv()                       # the reference semantics, all ranks in one process
    ok = np.array_equal(mine, wd.ranks[rank].f_post)

    # NCCL id plumbing: rank 0 creates the id, everyone receives the same 128 bytes; without a GPU init must refuse
    lib = L.lib()
    buf = C.create_string_buffer(128)
    if rank == 0:
        L.check(lib.mglc_comm_unique_id(buf))
    box = [buf.raw if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ids = [None] * world
    dist.all_gather_object(ids, box[0])
    ok &= all(i == ids[0] and len(i) == 128 for i in ids)
    if not torch.cuda.is_available():
        h = C.c_void_p()
        ok &= lib.mglc_comm_init_rank(C.byref(h), box[0], world, rank, 0) == L.E_NOGPU

    flags = [None] * world
    dist.all_gather_object(flags, bool(ok))
    dist.destroy_process_group()
    if rank == 0:
        print("GLOO HALO OK" if all(flags) else f"GLOO HALO FAILED {flags}")
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
