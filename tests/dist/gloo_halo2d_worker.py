"""One rank of the CPU (gloo) multi-process test of the 2-D drivers' halo plumbing: every process owns ONE block of the 2-D
Cartesian decomposition, packs its outgoing populations exactly as include/mglc.h's mglc_halo_plan_2d prescribes (the table the
CUDA drivers cross-check their own against at create time), moves them with torch.distributed send/recv the way the library
moves them with ncclSend/ncclRecv, and unpacks.  The result must equal the in-process P-rank oracles' message_passing_sendrecv()
(2-D lid), message_passing_f() and message_passing_g() (2-D thermal), bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mglc_b200 as mg  # noqa: E402
from oracle import oracle as orc  # noqa: E402

EX9 = [0, 1, 0, -1, 0, 1, -1, -1, 1]
EY9 = [0, 0, 1, 0, -1, 1, 1, -1, -1]


def region(n, d, recv):
    """index tuple into a halo'd lattice (q, nx+2, ny+2) of 2-D message d (0..3 faces, 4..7 corners): source or destination"""
    nx, ny = n
    if d < 4:
        axis, plus = d >> 1, not (d & 1)
        nfix = n[axis]
        fix = (0 if plus else nfix + 1) if recv else (nfix if plus else 1)
        return (fix, slice(1, ny + 1)) if axis == 0 else (slice(1, nx + 1), fix)
    a = d + 1
    px, py = EX9[a] > 0, EY9[a] > 0
    if recv:
        return (0 if px else nx + 1, 0 if py else ny + 1)
    return (nx if px else 1, ny if py else 1)


def exchange(plan, lattices, n):
    """lattices = {'f': f_post, 'g': g_post or None}; messages 0..7 move f, 8..11 move g"""
    reqs, recvs = [], []
    for m in plan:
        lat = lattices["g" if m["dir"] >= 8 else "f"]
        if lat is None:
            continue
        d = m["dir"] - 8 if m["dir"] >= 8 else m["dir"]
        if m["send_count"]:
            sl = region(n, d, recv=False)
            buf = np.stack([np.atleast_1d(np.asarray(lat[(a,) + sl])).ravel() for a in m["pops"]])       # [slot][t]
            assert buf.size == m["send_count"], (m, buf.shape)
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(buf.ravel())), dst=m["send_to"], tag=m["dir"]))
        if m["recv_count"]:
            t = torch.empty(m["recv_count"], dtype=torch.float64)
            reqs.append(dist.irecv(t, src=m["recv_from"], tag=m["dir"]))
            recvs.append((m, d, lat, t))
    for r in reqs:
        r.wait()
    for m, d, lat, t in recvs:
        sl = region(n, d, recv=True)
        per = m["recv_count"] // len(m["pops"])
        for s, a in enumerate(m["pops"]):
            lat[(a,) + sl] = t.numpy()[s * per:(s + 1) * per] if d < 4 else t.numpy()[s]


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    total = (13, 10)
    dims = None if len(sys.argv) < 2 else tuple(int(x) for x in sys.argv[1].split("x"))
    ok = True
    for kind in ("lid2d", "thermal2d"):
        wd = orc.Lid2DWorld(total, world, dims=dims) if kind == "lid2d" else orc.Thermal2DWorld(total, world, dims=dims)
        plan = mg.halo_plan_2d(total, wd.dims, rank)
        R = wd.ranks[rank]
        # my neighbours in the plan are the oracle's (right, left, top, bottom, then the corners populations 5..8 travel to)
        assert tuple(m["send_to"] for m in plan[:8]) == R.nbr + R.cnr, (plan, R.nbr, R.cnr)
        rng = np.random.default_rng(99)
        for Q in wd.ranks:                      # the same seeded global state on every process
            Q.f_post[...] = rng.random(Q.f_post.shape)
            if kind == "thermal2d":
                Q.g_post[...] = rng.random(Q.g_post.shape)
        mine_f = R.f_post.copy(order="F")
        mine_g = R.g_post.copy(order="F") if kind == "thermal2d" else None
        exchange(plan, {"f": mine_f, "g": mine_g}, R.n)
        if kind == "lid2d":
            wd.message_passing_sendrecv()
        else:
            wd.message_passing_f(); wd.message_passing_g()
        same = np.array_equal(mine_f, R.f_post) and (mine_g is None or np.array_equal(mine_g, R.g_post))
        # the volume is the reference's: 3 doubles per face cell, 1 per corner, 1 g double per face cell
        sent = sum(m["send_count"] for m in plan if kind == "thermal2d" or m["dir"] < 8)
        nf = sum(1 for k in R.nbr[:2] if k >= 0) * R.n[1] + sum(1 for k in R.nbr[2:] if k >= 0) * R.n[0]
        want = 3 * nf + sum(1 for k in R.cnr if k >= 0) + (nf if kind == "thermal2d" else 0)
        ok = ok and same and sent == want
        if not same or sent != want:
            print(f"rank {rank} {kind}: exchange {'ok' if same else 'MISMATCH'}, volume {sent} vs {want}", flush=True)
        wd.close()
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("GLOO HALO 2D OK" if int(flag.item()) else "GLOO HALO 2D FAILED", flush=True)
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
