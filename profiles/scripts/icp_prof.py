"""Minimal ICP launcher for ncu: cfg 2 (360 beams) or cfg 4 shape (1080 beams), 3 launches.
usage: icp_prof.py [360|1080] [icp_prune] [icp_block]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from b2slam import _lib, devapi, synth

if len(sys.argv) > 2:
    _lib.check(_lib.lib().b2s_tune(b"icp_prune", int(sys.argv[2])))
if len(sys.argv) > 3:
    _lib.check(_lib.lib().b2s_tune(b"icp_block", int(sys.argv[3])))

if len(sys.argv) > 1 and sys.argv[1] == "1080":
    tar, src, _ = synth.icp_pairs(4001, 8192, 1080)
else:
    xy, _ = synth.room_sequence(9001, 10000, 360)
    tar, src = np.ascontiguousarray(xy[:-1]), np.ascontiguousarray(xy[1:])
tar, src = torch.from_numpy(tar).cuda(), torch.from_numpy(src).cuda()
P = tar.shape[0]
T = torch.empty((P, 3, 3), dtype=torch.float64, device="cuda")
it = torch.empty(P, dtype=torch.int32, device="cuda")
for _ in range(3):
    devapi.icp_batch(tar, src, 30, 1e-3, T, it)
torch.cuda.synchronize()
print("ok", float(it.float().mean()))
