"""Small pass over every kernel for compute-sanitizer (memcheck / racecheck): tiny sizes, checked against the oracle."""
import math, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import b2slam
from b2slam import _lib, scan, synth
from b2slam import bresenham as drawing
from oracle import corc

tar, src, _ = synth.icp_pairs(7001, 6, 120)
icp = b2slam.ICP()
T, it = icp.process_batch(tar, src)
wT, wit = corc.icp_batch(tar, src, 30, 1e-3)
assert np.array_equal(it, wit) and np.abs(T - wT).max() < 1e-9
T2, it2 = icp.process_batch(tar[:, :, :97].copy(), src[:, :, :77].copy())     # unaligned: no bulk copy
xy, _ = synth.room_sequence(9001, 12, 120)
Ts, its = icp.process_sequence(xy)                                             # consecutive pairs in place
traj, To, ito = icp.odometry(xy, state=(0.1, 0.2, 0.3))                        # + pose chain on the device
assert np.array_equal(Ts, To)
rng = np.hypot(xy[:, 0], xy[:, 1]).astype(np.float32)
Tr, itr = icp.process_scans(rng, -math.pi, math.pi)                            # laserToNumpy inside the kernel
for prune in (0, 1, 3, 2):                                                     # every search mode
    _lib.check(_lib.lib().b2s_tune(b"icp_prune", prune))
    Tp, itp = icp.process_sequence(xy)
    assert np.array_equal(Tp, Ts) and np.array_equal(itp, its)
d, i = icp.findNearest(src[0].T.astype(np.float64), tar[0].T.astype(np.float64))
icp.getTransform(src[0].T.astype(np.float64), tar[0].T.astype(np.float64))
for v in (1, 2, 3, 4):
    _lib.check(_lib.lib().b2s_tune(b"grid_variant", v))
    ox, oy, cx, cy = synth.grid_scans(3, 6, 200, half_extent_m=5.0)
    m = b2slam.Mapping(256, 192, 0.05)
    pm = m.update_batch(ox, oy, cx, cy).copy()
    oh = np.zeros((256, 192), np.int32); om = np.zeros((256, 192), np.int32)
    corc.grid_raycast(oh, om, 20.0, 6.4, 4.8, ox, oy, cx, cy)
    h, mm = m.counts()
    assert np.array_equal(h, oh) and np.array_equal(mm, om), v
    assert np.array_equal(pm, corc.grid_finalize(oh, om)[1])
r, poses = synth.grid_scan_ranges(4, 5, 200, half_extent_m=5.0)
m = b2slam.Mapping(256, 256, 0.05)
m.update_scans(r, poses, -math.pi, math.pi)
try:
    bad = r.copy(); bad[2, 3] = np.nan
    m.update_scans(bad, poses, -math.pi, math.pi)
except ValueError:
    pass
drawing.bresenham([0, 0], [37, -11]).path
scan.compose_odometry_gpu((0, 0, 0), T)
scan.virtual_scan(np.random.rand(2, 500) * 10 - 5, (0.1, 0.2, 0.3), -math.pi, 2 * math.pi / 119, 120)
import torch
from b2slam import devapi
pm8 = torch.from_numpy(pm).cuda()
devapi.grid_pack_ros(pm8)
torch.cuda.synchronize()
print("sanitize smoke ok")
