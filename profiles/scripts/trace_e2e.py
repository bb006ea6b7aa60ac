"""Timelines (B2S_TRACE) of the host-buffer calls at the bench sizes: grid update_batch / update_scans, ICP sequence."""
import math, os, sys
os.environ["B2S_TRACE"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import b2slam
from b2slam import synth


def pinned(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


G, K, N = 4096, 16384, 1080
keep = [pinned(a) for a in synth.grid_scans(12001, K, N)]
h = [k[1] for k in keep]
m = b2slam.Mapping(G, G, 0.05)
for rep in range(3):
    m.reset()
    print("== update_batch rep %d" % rep, file=sys.stderr, flush=True)
    m.update_batch(*h)
ranges, poses = synth.grid_scan_ranges(12001, K, N)
kr, hr = pinned(ranges)
for rep in range(2):
    m.reset()
    print("== update_scans rep %d" % rep, file=sys.stderr, flush=True)
    m.update_scans(hr, poses, -math.pi, math.pi)
xy, _ = synth.room_sequence(9001, 10000, 360)
kq, hq = pinned(xy)
icp = b2slam.ICP()
for rep in range(3):
    print("== process_sequence rep %d" % rep, file=sys.stderr, flush=True)
    icp.process_sequence(hq)
