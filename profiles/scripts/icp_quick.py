"""Default ICP kernel timing, cfg 2 (9 999 x 360-beam pairs) and cfg-4 shape (16 384 x 1080-beam pairs): mean of 10."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from b2slam import devapi, synth


def run(name, tar, src):
    P = tar.shape[0]
    T = torch.empty((P, 3, 3), dtype=torch.float64, device="cuda")
    it = torch.empty(P, dtype=torch.int32, device="cuda")
    for _ in range(3):
        devapi.icp_batch(tar, src, 30, 1e-3, T, it)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        devapi.icp_batch(tar, src, 30, 1e-3, T, it)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print("%s: %.4f ms  %.3e pairs/s  (checksum %.12e, iterations %d)" % (
        name, ms, P / ms * 1e3, float(T.sum().item()), int(it.sum().item())), flush=True)


xy, _ = synth.room_sequence(9001, 10000, 360)
run("cfg2 360 beams", torch.from_numpy(np.ascontiguousarray(xy[:-1])).cuda(), torch.from_numpy(np.ascontiguousarray(xy[1:])).cuda())
tar, src, _ = synth.icp_pairs(4001, 16384, 1080)
run("cfg4 1080 beams", torch.from_numpy(tar).cuda(), torch.from_numpy(src).cuda())
