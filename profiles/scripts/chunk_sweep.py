"""Pipeline depth of the host-buffer grid calls on one box: update_batch (endpoints) and update_scans (raw scans)."""
import math, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import b2slam
from b2slam import synth, _lib

G, K, N = 4096, 16384, 1080
pin = [torch.from_numpy(a).pin_memory() for a in synth.grid_scans(12001, K, N)]
h = [p.numpy() for p in pin]
ranges, poses = synth.grid_scan_ranges(12001, K, N)
pr = torch.from_numpy(ranges).pin_memory()
hr = pr.numpy()
m = b2slam.Mapping(G, G, 0.05)


def timeit(fn, reps=8):
    fn(); fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        best = min(best, (time.perf_counter() - t0) / reps * 1e3)
    return best


for rnd in range(2):
    for c in (0, 4, 8, 12, 16):
        _lib.check(_lib.lib().b2s_tune(b"h2d_chunks", c))
        a = timeit(lambda: (m.reset(), m.update_batch(*h)))
        b = timeit(lambda: (m.reset(), m.update_scans(hr, poses, -math.pi, math.pi)))
        print("chunks %2d (0 = automatic): update_batch %.3f ms   update_scans %.3f ms" % (c, a, b), flush=True)
