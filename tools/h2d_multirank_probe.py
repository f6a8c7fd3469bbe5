#!/usr/bin/env python
"""Concurrent pinned-host -> device copy bandwidth of this box, per GPU, against the number of GPUs copying at once and
against the NUMA node the host buffer lives on.  No kernels run: this is the ceiling the host-buffer path (`e2e`) can reach.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/h2d_multirank_probe.py

Phase 0 (rank 0): topology as the VM shows it (lscpu, NUMA nodes, GPU <-> node affinity from sysfs and nvidia-smi topo).
Phase 1: every rank ALONE, buffer bound to each NUMA node in turn (mbind + first touch + cudaHostRegister) -> best node per GPU.
Phase 2: k = 1, 2, 4, .. N ranks copying simultaneously, (a) torch's default pinned allocation, (b) buffers on each rank's
         best node, (c) two copy streams per GPU.
Writes gpurun_out/h2d_multirank_probe.json (rank 0)."""
import ctypes
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

REPO = Path(__file__).resolve().parent.parent
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

BYTES = 1 << 30                      # 1 GiB per buffer
libc = ctypes.CDLL("libc.so.6", use_errno=True)
SYS_MBIND, MPOL_BIND, MPOL_DEFAULT = 237, 2, 0
cudart = torch.cuda.cudart()


def barrier():
    if world > 1:
        dist.barrier(device_ids=[local])
    torch.cuda.synchronize()


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:
        return f"<{e}>"


def numa_nodes():
    base = Path("/sys/devices/system/node")
    return sorted(int(p.name[4:]) for p in base.glob("node[0-9]*")) if base.exists() else [0]


def bound_buffer(node):
    """1 GiB of anonymous memory bound to `node` (mbind), first-touched, then page-locked with cudaHostRegister."""
    mm = np.empty(BYTES + 4096, dtype=np.uint8)
    addr = (mm.ctypes.data + 4095) & ~4095
    if node is not None:
        mask = ctypes.c_ulong(1 << node)
        r = libc.syscall(SYS_MBIND, ctypes.c_void_p(addr), ctypes.c_ulong(BYTES), MPOL_BIND, ctypes.byref(mask), ctypes.c_ulong(64), 0)
        if r != 0:
            return None, f"mbind errno {ctypes.get_errno()}"
    view = mm[addr - mm.ctypes.data: addr - mm.ctypes.data + BYTES]
    view[::4096] = 1                                         # first touch under the policy
    err = cudart.cudaHostRegister(addr, BYTES, 0)
    if int(err) != 0:
        return None, f"cudaHostRegister {err}"
    t = torch.from_numpy(view)
    t._keep = mm
    return t, None


def copy_gbs(src, seconds=0.6, streams=1):
    """device-timed H2D rate of `src` (CPU uint8 tensor, page-locked) copied repeatedly for ~`seconds`."""
    dst = [torch.empty(BYTES // streams, dtype=torch.uint8, device=dev) for _ in range(streams)]
    ss = [torch.cuda.Stream() for _ in range(streams)]
    parts = [src[i * (BYTES // streams):(i + 1) * (BYTES // streams)] for i in range(streams)]
    for s, d, p in zip(ss, dst, parts):
        with torch.cuda.stream(s):
            d.copy_(p, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps, t0 = 0, time.perf_counter()
    e0.record()
    for s in ss:
        s.wait_event(e0)
    while time.perf_counter() - t0 < seconds:
        for s, d, p in zip(ss, dst, parts):
            with torch.cuda.stream(s):
                d.copy_(p, non_blocking=True)
        reps += 1
        if reps % 4 == 0:
            torch.cuda.synchronize()
    for s in ss:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return BYTES * reps / (e0.elapsed_time(e1) / 1e3) / 1e9


out = {"world": world}
nodes = numa_nodes()
if rank == 0:
    out["topology"] = {
        "lscpu": sh("lscpu | egrep 'Model name|Socket|NUMA|^CPU\\(s\\)|Thread|Core'"),
        "numa_nodes": nodes,
        "node_cpulists": {n: sh(f"cat /sys/devices/system/node/node{n}/cpulist") for n in nodes},
        "node_meminfo": {n: sh(f"grep MemTotal /sys/devices/system/node/node{n}/meminfo") for n in nodes},
        "nvidia_smi_topo": sh("nvidia-smi topo -m"),
        "affinity": sorted(os.sched_getaffinity(0)),
        "hugepages": sh("grep -i huge /proc/meminfo"),
    }
_pr = torch.cuda.get_device_properties(local)
try:
    bus = f"{int(getattr(_pr, 'pci_domain_id', 0)):04x}:{int(_pr.pci_bus_id):02x}:{int(getattr(_pr, 'pci_device_id', 0)):02x}.0"
except Exception:
    bus = ""
try:
    sysfs = Path(f"/sys/bus/pci/devices/{bus}/numa_node").read_text().strip()
except Exception as e:
    sysfs = f"<{type(e).__name__}>"

# ---- phase 1: alone, per NUMA node
per_node = {}
default_buf = torch.empty(BYTES, dtype=torch.uint8, pin_memory=True)
default_buf[::4096] = 1
bufs = {}
for n in nodes:
    t, err = bound_buffer(n)
    bufs[n] = t
    if t is None:
        per_node[n] = err
for r in range(world):
    barrier()
    if r == rank:
        per_node["default"] = copy_gbs(default_buf, 0.4)
        for n in nodes:
            if bufs[n] is not None:
                per_node[n] = copy_gbs(bufs[n], 0.4)
barrier()
best = max((n for n in nodes if bufs[n] is not None), key=lambda n: per_node[n], default=None)

# ---- phase 2: concurrency sweep
sweep = {}
k = 1
levels = []
while k <= world:
    levels.append(k)
    k *= 2
for k in levels:
    for tag, src, streams in (("default", default_buf, 1), ("best_node", bufs.get(best) if best is not None else None, 1),
                              ("default_2streams", default_buf, 2)):
        barrier()
        val = copy_gbs(src, 0.8, streams) if (rank < k and src is not None) else 0.0
        barrier()
        sweep[f"k{k}_{tag}"] = val

mine = {"rank": rank, "pci": bus, "sysfs_numa_node": sysfs, "alone_by_node": {str(a): b for a, b in per_node.items()},
        "best_node": best, "sweep": sweep}
if world > 1:
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
else:
    gathered = [mine]
if rank == 0:
    out["ranks"] = gathered
    summary = {}
    for key in sweep:
        vals = [g["sweep"][key] for g in gathered if g["sweep"][key] > 0]
        summary[key] = {"per_gpu_min": round(min(vals), 1), "per_gpu_mean": round(sum(vals) / len(vals), 1), "total": round(sum(vals), 1)} if vals else None
    out["summary_gbs"] = summary
    (REPO / "gpurun_out").mkdir(exist_ok=True)
    (REPO / "gpurun_out" / "h2d_multirank_probe.json").write_text(json.dumps(out, indent=1))
    print(json.dumps({"summary_gbs": summary, "best_nodes": [g["best_node"] for g in gathered],
                      "alone": [g["alone_by_node"] for g in gathered]}, indent=1))
    print(out["topology"]["lscpu"])
    print(out["topology"]["nvidia_smi_topo"])
if world > 1:
    dist.destroy_process_group()
