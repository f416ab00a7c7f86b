"""Host-side facts of a GPU box that bound the e2e figure of bench.py: GPU <-> NUMA topology and the pinned-memory
host<->device copy rate per NUMA placement.  Run under gpurun; prints JSON lines.
    python tools/probe_host.py [--gpus N] [--gib 4]"""
import argparse
import ctypes
import glob
import json
import os
import subprocess
import time

import torch

SYS_set_mempolicy = 238        # x86_64
MPOL_DEFAULT, MPOL_PREFERRED, MPOL_BIND = 0, 1, 2
libc = ctypes.CDLL(None, use_errno=True)


def set_mempolicy(mode, node=None):
    if node is None:
        return libc.syscall(SYS_set_mempolicy, MPOL_DEFAULT, None, 0)
    mask = ctypes.c_ulong(1 << node)
    return libc.syscall(SYS_set_mempolicy, mode, ctypes.byref(mask), 64)


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=60).stdout
    except Exception as ex:
        return f"failed: {ex}"


def copy_rate(dev, nbytes, node):
    set_mempolicy(MPOL_BIND if node is not None else MPOL_DEFAULT, node)
    t0 = time.perf_counter()
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h.fill_(1)
    pin_s = time.perf_counter() - t0
    set_mempolicy(MPOL_DEFAULT)
    d = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{dev}")
    out = {}
    for name, (dst, src) in {"h2d": (d, h), "d2h": (h, d)}.items():
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.device(dev):
            e0.record()
            for _ in range(3):
                dst.copy_(src, non_blocking=True)
            e1.record(); e1.synchronize()
        out[name + "_gbs"] = round(3 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9, 2)
    out["pin_gbs"] = round(nbytes / pin_s / 1e9, 2)
    del h, d
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--gib", type=float, default=4.0)
    a = ap.parse_args()
    nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
    facts = {"nodes": nodes, "cpus": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)),
             "node_cpulist": {n: open(f"/sys/devices/system/node/node{n}/cpulist").read().strip() for n in nodes},
             "meminfo": [l.strip() for l in open("/proc/meminfo") if l.startswith(("MemTotal", "MemAvailable"))],
             "gpu_numa": {}}
    for i in range(torch.cuda.device_count()):
        busid = sh(f"nvidia-smi -i {i} --query-gpu=pci.bus_id --format=csv,noheader").strip().lower()
        busid = busid[4:] if len(busid) > 12 else busid          # 00000000:1b:00.0 -> 0000:1b:00.0
        try:
            facts["gpu_numa"][i] = int(open(f"/sys/bus/pci/devices/{busid}/numa_node").read())
        except Exception as ex:
            facts["gpu_numa"][i] = f"{busid}: {ex}"
    print(json.dumps(facts))
    print(sh("nvidia-smi topo -m"))
    print(sh("lscpu | head -25"))
    nbytes = int(a.gib * (1 << 30))
    for dev in range(min(a.gpus, torch.cuda.device_count())):
        for node in [None] + nodes:
            print(json.dumps({"gpu": dev, "host_node": node, **copy_rate(dev, nbytes, node)}), flush=True)


if __name__ == "__main__":
    main()
