"""Does a running kernel slow pinned H2D copies?  Chunked copies on a side stream, idle GPU vs beside the ICP kernel
vs beside the grid ray-cast (1 GPU)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from b2slam import devapi, synth

side = torch.cuda.Stream()
for mb in (8, 32):
    n = mb << 20
    src = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(8)]
    dst = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(8)]

    def copies(tag):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(9)]
        with torch.cuda.stream(side):
            evs[0].record(side)
            for k in range(8):
                dst[k].copy_(src[k], non_blocking=True)
                evs[k + 1].record(side)
        return evs

    def report(tag, evs):
        torch.cuda.synchronize()
        ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(8)]
        print("%-34s %2d MB copies: %s ms -> %.1f GB/s" % (tag, mb, " ".join("%.3f" % m for m in ms), 8 * n / sum(ms) / 1e6), flush=True)

    for _ in range(2):
        report("idle GPU", copies("idle"))
    xy, _ = synth.room_sequence(9001, 10000, 360)
    tar, s2 = torch.from_numpy(np.ascontiguousarray(xy[:-1])).cuda(), torch.from_numpy(np.ascontiguousarray(xy[1:])).cuda()
    T = torch.empty((9999, 3, 3), dtype=torch.float64, device="cuda"); it = torch.empty(9999, dtype=torch.int32, device="cuda")
    devapi.icp_batch(tar, s2, 30, 1e-3, T, it); torch.cuda.synchronize()
    for _ in range(2):
        for _ in range(6):
            devapi.icp_batch(tar, s2, 30, 1e-3, T, it)
        report("beside icp_batch (6 launches)", copies("icp"))
    G, K, N = 4096, 16384, 1080
    S, Hx, Hy = devapi.grid_scale(G, G, 0.05)
    ox, oy, cx, cy = (torch.from_numpy(a).cuda() for a in synth.grid_scans(12001, K, N))
    hit, miss = devapi.new_planes(G, G); ws = devapi.new_workspace(G, G)
    devapi.grid_raycast(hit, miss, S, Hx, Hy, ox, oy, cx, cy, workspace=ws); torch.cuda.synchronize()
    for _ in range(2):
        for _ in range(3):
            devapi.grid_raycast(hit, miss, S, Hx, Hy, ox, oy, cx, cy, workspace=ws)
        report("beside grid_raycast (3 launches)", copies("grid"))
    for _ in range(2):
        for k in range(24):
            lo, hi = (k % 8) * 2048, (k % 8 + 1) * 2048
            devapi.grid_raycast(hit, miss, S, Hx, Hy, ox[lo:hi], oy[lo:hi], cx[lo:hi], cy[lo:hi], workspace=ws)
        report("beside 24 chunk ray-casts", copies("gridc"))
    del src, dst
