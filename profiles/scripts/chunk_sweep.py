"""Pipeline depth sweep of the host-buffer calls (1 GPU): ICP cfg 2 batch and grid cfg 3 batch."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import b2slam
from b2slam import synth, _lib

xy, _ = synth.room_sequence(9001, 10000, 360)
pt = torch.from_numpy(np.ascontiguousarray(xy[:-1])).pin_memory(); ps = torch.from_numpy(np.ascontiguousarray(xy[1:])).pin_memory()
ht, hs = pt.numpy(), ps.numpy()
icp = b2slam.ICP()
host = synth.grid_scans(12001, 16384, 1080)
pin = [torch.from_numpy(a).pin_memory() for a in host]
h = [p.numpy() for p in pin]
m = b2slam.Mapping(4096, 4096, 0.05)


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for c in (0, 1, 2, 3, 4, 6, 8):
    _lib.check(_lib.lib().b2s_tune(b"h2d_chunks", c))
    print("chunks %d  icp %.3f ms  grid(+map) %.3f ms" % (c, timeit(lambda: icp.process_batch(ht, hs)),
                                                          timeit(lambda: m.update_batch(*h))))
