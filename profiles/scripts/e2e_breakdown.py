"""Where the host-buffer grid step spends its time (1 GPU)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import b2slam
from b2slam import synth

G, K, N = 4096, 16384, 1080
host = synth.grid_scans(12001, K, N)
pin = [torch.from_numpy(a).pin_memory() for a in host]
h = [p.numpy() for p in pin]
m = b2slam.Mapping(G, G, 0.05)
m.update_batch(*h)


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


print("reset                         %.3f ms" % timeit(m.reset))
print("update_batch no map           %.3f ms" % timeit(lambda: m.update_batch(*h, want_pmap=False)))
print("update_batch + map            %.3f ms" % timeit(lambda: m.update_batch(*h, want_pmap=True)))
print("occupancy() read-back only    %.3f ms" % timeit(m.occupancy))
d = [p.cuda() for p in pin]
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    for p, q in zip(pin, d):
        q.copy_(p, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print("plain H2D of the inputs       %.3f ms  (%.1f GB/s)" % (dt * 1e3, sum(p.numel() * 4 for p in pin) / dt / 1e9))
for k in (1024, 2048, 4096, 8192):
    sub = [a[:k] for a in h]
    print("update_batch %5d scans no map %.3f ms" % (k, timeit(lambda: m.update_batch(*sub, want_pmap=False))))
