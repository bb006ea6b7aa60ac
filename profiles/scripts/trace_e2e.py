"""Timelines (B2S_TRACE) of the host-buffer calls at the bench sizes: grid update_batch / update_scans, ICP sequence.
Each call is first repeated untraced for 80 ms (after an idle period the PCIe link and the clocks need some ms to come
back, see pinned_probe.py), then traced once."""
import math, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import b2slam
from b2slam import synth


def pinned(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


def traced(name, fn):
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.08:
        fn()
    print("== %s" % name, file=sys.stderr, flush=True)
    os.environ["B2S_TRACE"] = "1"
    try:
        fn()
    finally:
        del os.environ["B2S_TRACE"]


G, K, N = 4096, 16384, 1080
keep = [pinned(a) for a in synth.grid_scans(12001, K, N)]
h = [k[1] for k in keep]
m = b2slam.Mapping(G, G, 0.05)
traced("Mapping.update_batch, 16384 scans x 1080 beams (endpoints, 141.7 MB in, 16.8 MB out)",
       lambda: (m.reset(), m.update_batch(*h)))
ranges, poses = synth.grid_scan_ranges(12001, K, N)
kr, hr = pinned(ranges)
traced("Mapping.update_scans, the same scans as raw ranges + poses (71.3 MB in)",
       lambda: (m.reset(), m.update_scans(hr, poses, -math.pi, math.pi)))
xy, _ = synth.room_sequence(9001, 10000, 360)
kq, hq = pinned(xy)
icp = b2slam.ICP()
traced("ICP.process_sequence, 10000 scans x 360 beams (28.8 MB in)", lambda: icp.process_sequence(hq))
krr, hrr = pinned(np.hypot(xy[:, 0], xy[:, 1]).astype(np.float32))
traced("ICP.process_scans, the same stream as raw ranges (14.4 MB in)", lambda: icp.process_scans(hrr, -math.pi, math.pi))
