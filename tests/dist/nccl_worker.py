"""One rank of the multi-GPU parity run (launched by tests/test_multigpu.py through torch.distributed.run, one
process per GPU, halos over NCCL / direct NVLink stores).  Every rank steps its own subdomain through libmglc.so; rank 0
gathers the fields and compares them with the single-rank CPU oracle (tests/dist/parity_suite.py).  Prints
'MULTIGPU OK' on success."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mglc_b200 as mg  # noqa: E402
import parity_suite as ps  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def bcast(b):
        box = [b]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def log(msg):
        if rank == 0:
            print(msg, flush=True)

    comm = mg.Communicator(world, rank, local, bcast)
    only = [c for c in os.environ.get("MGLC_PARITY_ONLY", "").split(",") if c]       # e.g. lid_aa,jacobi while developing
    verdicts = {name: getattr(ps, name)(comm, rank, world, log=log)
                for name in ("lid", "lid_aa", "thermal", "jacobi", "particles", "drivers_2d") if not only or name in only}
    ok = not ps.failed(verdicts)
    # on an NVLink box the CUDA IPC mappings of the direct path must come up (unless they were switched off on purpose)
    if rank == 0 and not os.environ.get("MGLC_NO_DIRECT") and \
            (verdicts.get("lid", {}).get("direct") == "unavailable" or verdicts.get("lid_aa", {}).get("single_lattice") == "unavailable"):
        ok = False
    comm.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(verdicts))
        print("MULTIGPU OK" if ok else "MULTIGPU FAILED")
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
