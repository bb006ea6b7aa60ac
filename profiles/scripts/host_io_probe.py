"""Where do the host-buffer (end-to-end) calls of N ranks on one box lose time?  Run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 profiles/scripts/host_io_probe.py

Prints (rank 0) the box topology -- GPU -> NUMA node, CPU lists, nvidia-smi topo -- and for every rank the rate of
CONCURRENT pinned host<->device copies of the sizes one e2e step moves (71 MB of ranges in, 16.8 MB of map out),
first with the pinned buffers wherever the launcher put the process, then after binding the process to the CPUs of its
GPU's NUMA node and re-allocating them (dist.bind_to_gpu_numa).  Then a coarse timeline of ShardedMappingP2P.update_scans."""
import math
import os
import subprocess
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
os.environ["B2S_NUMA_BIND"] = "0"          # first measurement: as launched
import b2slam.dist as bdist
from b2slam import synth

rank, local, world = bdist.init()
dist = torch.distributed


def everyone(value):
    t = torch.tensor([float(value)], dtype=torch.float64, device="cuda")
    parts = [torch.empty_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(parts, t)
    else:
        parts = [t]
    return [float(p.item()) for p in parts]


def say(*a):
    if rank == 0:
        print(*a, flush=True)


if rank == 0:
    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"], ["numactl", "-H"]):
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout
            if cmd[0] == "lscpu":
                out = "\n".join(l for l in out.splitlines() if any(k in l for k in ("NUMA", "Model name", "Socket", "CPU(s):", "Thread")))
            print("$ " + " ".join(cmd) + "\n" + out, flush=True)
        except Exception as e:
            print("$ %s: %s" % (" ".join(cmd), e), flush=True)
nodes = everyone(-1 if bdist.gpu_numa_node(local) is None else bdist.gpu_numa_node(local))
ncpu = everyone(len(os.sched_getaffinity(0)))
say("GPU -> NUMA node per rank:", [int(n) for n in nodes], " CPUs allowed per rank:", [int(c) for c in ncpu])

IN_BYTES, OUT_BYTES, REPS = 71303168, 16777216, 20


def copy_rates(tag):
    h_in = torch.empty(IN_BYTES, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(OUT_BYTES, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    d_in = torch.empty(IN_BYTES, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(OUT_BYTES, dtype=torch.uint8, device="cuda")
    s2 = torch.cuda.Stream()
    res = []
    for what in ("h2d", "d2h", "both"):
        for _ in range(3):
            d_in.copy_(h_in, non_blocking=True)
        torch.cuda.synchronize()
        bdist.barrier()
        t0 = time.perf_counter()
        for _ in range(REPS):
            if what in ("h2d", "both"):
                d_in.copy_(h_in, non_blocking=True)
            if what in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        nbytes = (IN_BYTES if what != "d2h" else 0) + (OUT_BYTES if what != "h2d" else 0)
        res.append(everyone(nbytes * REPS / dt / 1e9))
        bdist.barrier()
    say("== concurrent pinned copies, %s (GB/s per rank; sum)" % tag)
    for what, r in zip(("h2d 71 MB", "d2h 16.8 MB", "both"), res):
        say("   %-12s" % what + " ".join("%6.1f" % v for v in r) + "   sum %.0f" % sum(r))


def e2e(tag):
    K, N = 16384, 1080
    ranges, poses = synth.grid_scan_ranges(12001 + rank, K, N)
    keep = torch.from_numpy(ranges).pin_memory()
    sm = bdist.ShardedMappingP2P(4096, 4096, 0.05)
    for _ in range(4):
        sm.update_scans(keep, poses, -math.pi, math.pi)
    bdist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        sm.update_scans(keep, poses, -math.pi, math.pi)
    dt = (time.perf_counter() - t0) / 10
    r = everyone(dt * 1e3)
    say("== ShardedMappingP2P.update_scans, %s: ms per call per rank " % tag + " ".join("%.2f" % v for v in r) +
        "  -> %.2f G beams/s" % (world * K * N / max(r) / 1e6))
    sm.close()


copy_rates("as launched")
e2e("as launched")
os.environ["B2S_NUMA_BIND"] = "1"
node = bdist.bind_to_gpu_numa(local)
nodes = everyone(-1 if node is None else node)
ncpu = everyone(len(os.sched_getaffinity(0)))
say("bound to NUMA node per rank:", [int(n) for n in nodes], " CPUs allowed per rank:", [int(c) for c in ncpu])
copy_rates("bound to the GPU's NUMA node, buffers re-allocated")
e2e("bound")
if world > 1:
    dist.destroy_process_group()
