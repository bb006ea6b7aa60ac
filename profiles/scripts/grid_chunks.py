"""Device-side pieces of the host-buffer grid step: the ray-cast in one launch vs 8 chunk launches, plain and
fused-ingestion form, alone and with the H2D copy of the next chunk running beside it (1 GPU)."""
import math, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from b2slam import devapi, synth, scan

G, K, N = 4096, 16384, 1080
S, Hx, Hy = devapi.grid_scale(G, G, 0.05)
host = synth.grid_scans(12001, K, N)
pin = [torch.from_numpy(a).pin_memory() for a in host]
ox, oy, cx, cy = (p.cuda() for p in pin)
ranges, poses = synth.grid_scan_ranges(12001, K, N)
t0 = time.perf_counter(); pose4_h = scan.pose_table(poses); t1 = time.perf_counter()
print("host pose_table(16384 poses)      %.3f ms" % ((t1 - t0) * 1e3))
d_ranges = torch.from_numpy(ranges).cuda()
pose4 = torch.from_numpy(pose4_h).cuda()
beam_cs = torch.from_numpy(scan.beam_table(-math.pi, math.pi, N)).cuda()
hit, miss = devapi.new_planes(G, G)
ws = devapi.new_workspace(G, G)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=5):
    out = []
    for _ in range(reps + 1):
        hit.zero_(); miss.zero_(); flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return sum(out[1:]) / reps


def plain(lo, hi):
    devapi.grid_raycast(hit, miss, S, Hx, Hy, ox[lo:hi], oy[lo:hi], cx[lo:hi], cy[lo:hi], workspace=ws)


def fused(lo, hi):
    devapi.grid_raycast_ranges(hit, miss, S, Hx, Hy, d_ranges[lo:hi], pose4[lo:hi], beam_cs, workspace=ws)


for name, fn in (("plain", plain), ("fused", fused)):
    print("%s ray-cast, one launch            %.3f ms" % (name, timed(lambda: fn(0, K))))
    for c in (2, 4, 8, 16):
        print("%s ray-cast, %2d chunk launches     %.3f ms" % (name, c, timed(lambda: [fn(K * k // c, K * (k + 1) // c) for k in range(c)])))

side = torch.cuda.Stream()
dst = [torch.empty_like(p, device="cuda") for p in pin]


def with_copy():
    with torch.cuda.stream(side):
        for p, q in zip(pin, dst):
            q.copy_(p, non_blocking=True)
    plain(0, K)


print("plain ray-cast beside a 141 MB H2D  %.3f ms" % timed(with_copy))
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(side):
    a.record(side)
    for p, q in zip(pin, dst):
        q.copy_(p, non_blocking=True)
    b.record(side)
for _ in range(2):
    plain(0, K)
torch.cuda.synchronize()
print("141 MB H2D beside the ray-cast      %.3f ms" % a.elapsed_time(b))
