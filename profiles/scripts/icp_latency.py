"""Single-pair / small-batch ICP latency against points per thread (one CTA per pair: fewer points per thread =
more warps per pair = shorter dependent chains).  usage: icp_latency.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import b2slam
from b2slam import _lib, devapi, synth
tune = _lib.lib().b2s_tune

for beams in (360, 1080):
    tar, src, _ = synth.icp_pairs(7001, 256, beams)
    t3, s3 = synth.homogeneous(tar[0].astype(np.float64)), synth.homogeneous(src[0].astype(np.float64))
    icp = b2slam.ICP(max_iter=10, tolerance=0.0)
    ref = None
    for r in (0, 2, 3, 4):
        _lib.check(tune(b"icp_src_per_thread", r))
        for _ in range(10):
            T = icp.process(t3, s3)
        t0 = time.perf_counter()
        for _ in range(100):
            T = icp.process(t3, s3)
        one = (time.perf_counter() - t0) / 100 * 1e3
        dt, ds = torch.from_numpy(tar).cuda(), torch.from_numpy(src).cuda()
        To = torch.empty((256, 3, 3), dtype=torch.float64, device="cuda"); it = torch.empty(256, dtype=torch.int32, device="cuda")
        res = []
        for P in (1, 32, 148, 256):
            for _ in range(3):
                devapi.icp_batch(dt[:P], ds[:P], 10, 0.0, To[:P], it[:P])
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                devapi.icp_batch(dt[:P], ds[:P], 10, 0.0, To[:P], it[:P])
            b.record(); torch.cuda.synchronize()
            res.append("%d pairs %.1f us" % (P, a.elapsed_time(b) / 20 * 1e3))
        if ref is None:
            ref = T
        print("beams %d points/thread %d: ICP.process %.4f ms | kernel: %s | max |dT| vs auto %.1e"
              % (beams, r, one, ", ".join(res), float(np.abs(T - ref).max())), flush=True)
    tune(b"icp_src_per_thread", 0)
