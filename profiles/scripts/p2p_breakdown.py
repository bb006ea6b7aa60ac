"""Per-stage device times of the multi-GPU grid step, per rank (run under torchrun on >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 profiles/scripts/p2p_breakdown.py

Stages of the flag-synchronised step (dist.ShardedMappingP2P, fence="flags"): clear the dirty tiles of the delta
planes, ray-cast (+ fold), publish (dirty map + ready flag pushed to every rank), merge (waits for every ready flag
inside the kernel, then reduce-scatter + finalize + all-gather over peer memory, raises the done flags), wait (for every
rank's done flag).  Printed for every variant: mean per stage on every rank, then max and spread over the ranks.
The NCCL-fenced form of round 1 and the plain ncclAllReduce step are timed beside it."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import b2slam.dist as bdist
from b2slam import _lib, devapi, synth

rank, local, world = bdist.init()
G, K, N = 4096, int(os.environ.get("B2S_SCANS", 16384)), 1080
REPS, WARM = int(os.environ.get("B2S_REPS", 60)), 5
host = synth.grid_scans(12001 + rank, K, N)
ox, oy, cx, cy = (torch.from_numpy(a).cuda() for a in host)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
L = _lib.lib()


def gather(rows):
    t = torch.tensor(rows, dtype=torch.float64, device="cuda")
    parts = [torch.empty_like(t) for _ in range(world)]
    torch.distributed.all_gather(parts, t)
    return torch.stack(parts).cpu().numpy()     # [rank][stage]


def report(title, names, per_rank):
    if rank != 0:
        return
    print("== %s (%d GPUs, %d scans per GPU, mean of %d steps, ms)" % (title, world, K, REPS))
    print("   %-6s" % "rank" + "".join("%12s" % n for n in names))
    for r in range(world):
        print("   %-6d" % r + "".join("%12.3f" % v for v in per_rank[r]))
    print("   %-6s" % "max" + "".join("%12.3f" % v for v in per_rank.max(0)))
    print("   %-6s" % "spread" + "".join("%12.3f" % v for v in (per_rank.max(0) - per_rank.min(0))))
    print(json.dumps({"variant": title, "n_gpus": world, "stages": names, "max_ms": per_rank.max(0).tolist(),
                      "min_ms": per_rank.min(0).tolist()}))


def run_flags(fence):
    sm = bdist.ShardedMappingP2P(G, G, 0.05, fence=fence)
    S, Hx, Hy = sm.scale
    stream = torch.cuda.current_stream().cuda_stream
    w = sm.weights
    names = ["clear", "raycast", "publish", "merge", "wait", "step"] if fence == "flags" else \
            ["clear", "raycast", "gather", "merge", "fence", "step"]
    acc = np.zeros(len(names))
    for it in range(WARM + REPS):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        ev[0].record()
        sm._begin()
        ev[1].record()
        devapi.grid_raycast(sm.d_hit, sm.d_miss, S, Hx, Hy, ox, oy, cx, cy, counters=sm._counters(), workspace=sm.workspace)
        ev[2].record()
        if fence == "flags":
            sm.epoch += 1
            ws = sm._ct.c_void_p(sm.workspace.data_ptr())
            _lib.check(L.b2s_p2p_publish(ws, sm.counters.data_ptr(), sm._dirty_ptrs, sm._flag_ptrs, world, rank, G, G,
                                         sm.epoch, stream))
            ev[3].record()
            _lib.check(L.b2s_grid_merge_p2p_tiles_sync(sm._hit_ptrs, sm._miss_ptrs, sm._pmap_ptrs, sm.all_dirty.data_ptr(),
                                                       sm._flag_ptrs, world, rank, sm.epoch, G, G, sm.tile_lo, sm.tile_hi,
                                                       sm.g_hit.data_ptr(), sm.g_miss.data_ptr(), w[0], w[1], w[2], stream))
            ev[4].record()
            _lib.check(L.b2s_p2p_wait_done(sm.flags.data_ptr(), world, sm.epoch, stream))
        else:
            torch.distributed.all_gather_into_tensor(sm.all_dirty, sm.dirty)
            ev[3].record()
            _lib.check(L.b2s_grid_merge_p2p_tiles(sm._hit_ptrs, sm._miss_ptrs, sm._pmap_ptrs, sm.all_dirty.data_ptr(), world,
                                                  G, G, sm.tile_lo, sm.tile_hi, sm.g_hit.data_ptr(), sm.g_miss.data_ptr(),
                                                  w[0], w[1], w[2], stream))
            ev[4].record()
            sm._fence()
        ev[5].record()
        flush.zero_()
        torch.cuda.synchronize()
        if it >= WARM:
            acc += [ev[i].elapsed_time(ev[i + 1]) for i in range(5)] + [ev[0].elapsed_time(ev[5])]
    sm.check()
    report("peer-memory merge, fence = %s" % fence, names, gather((acc / REPS).tolist()))
    sm.close()


def run_allreduce():
    S, Hx, Hy = devapi.grid_scale(G, G, 0.05)
    hit, miss = devapi.new_planes(G, G)
    pmap = torch.empty((G, G), dtype=torch.int8, device="cuda")
    ws = devapi.new_workspace(G, G)
    names = ["zero", "raycast", "allreduce", "finalize", "step"]
    acc = np.zeros(len(names))
    for it in range(WARM + REPS // 2):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        hit.zero_(); miss.zero_()
        ev[1].record()
        devapi.grid_raycast(hit, miss, S, Hx, Hy, ox, oy, cx, cy, workspace=ws)
        ev[2].record()
        bdist.allreduce_counts(hit, miss)
        ev[3].record()
        devapi.grid_finalize(hit, miss, pmap=pmap)
        ev[4].record()
        flush.zero_()
        torch.cuda.synchronize()
        if it >= WARM:
            acc += [ev[i].elapsed_time(ev[i + 1]) for i in range(4)] + [ev[0].elapsed_time(ev[4])]
    report("ncclAllReduce of both planes + finalize", names, gather((acc / (REPS // 2)).tolist()))


run_flags("flags")
run_flags("nccl")
run_allreduce()
torch.distributed.destroy_process_group()
