"""ICP kernel variants side by side (1 GPU): search mode x pruning block size, cfg 2 and cfg 4 shapes."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import b2slam
from b2slam import _lib, devapi, synth

tune = _lib.lib().b2s_tune


def run(name, tar, src):
    P = tar.shape[0]
    T = torch.empty((P, 3, 3), dtype=torch.float64, device="cuda")
    it = torch.empty(P, dtype=torch.int32, device="cuda")
    ref = None
    for prune, block in ((0, 0), (2, 8), (2, 16), (4, 8), (4, 16)):
        _lib.check(tune(b"icp_prune", prune)); _lib.check(tune(b"icp_block", block))
        for r in ((0,) if prune == 0 else (0, 3)):
            _lib.check(tune(b"icp_src_per_thread", r))
            for _ in range(2):
                devapi.icp_batch(tar, src, 30, 1e-3, T, it)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5 if prune else 2
            a.record()
            for _ in range(reps):
                devapi.icp_batch(tar, src, 30, 1e-3, T, it)
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
            if ref is None:
                ref = (T.clone(), it.clone())
            same = bool(torch.equal(T, ref[0]) and torch.equal(it, ref[1]))
            print("%s prune %d block %2d src/thread %d: %8.3f ms  %10.3e pairs/s  bit-identical to brute force: %s"
                  % (name, prune, block, r, ms, P / ms * 1e3, same), flush=True)
    tune(b"icp_prune", 4); tune(b"icp_block", 0); tune(b"icp_src_per_thread", 0)


xy, _ = synth.room_sequence(9001, 10000, 360)
run("cfg2 360 beams", torch.from_numpy(np.ascontiguousarray(xy[:-1])).cuda(), torch.from_numpy(np.ascontiguousarray(xy[1:])).cuda())
tar, src, _ = synth.icp_pairs(4001, 16384, 1080)
run("cfg4 1080 beams", torch.from_numpy(tar).cuda(), torch.from_numpy(src).cuda())
