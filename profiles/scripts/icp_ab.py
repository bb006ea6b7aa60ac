"""A/B of library builds (build/ab/lib_*.so) on the two ICP shapes; each build runs in its own process."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from b2slam import _lib, devapi, synth
def run(name, tar, src, prune, block):
    _lib.check(_lib.lib().b2s_tune(b"icp_prune", prune)); _lib.check(_lib.lib().b2s_tune(b"icp_block", block))
    P = tar.shape[0]
    T = torch.empty((P, 3, 3), dtype=torch.float64, device="cuda"); it = torch.empty(P, dtype=torch.int32, device="cuda")
    for _ in range(3): devapi.icp_batch(tar, src, 30, 1e-3, T, it)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): devapi.icp_batch(tar, src, 30, 1e-3, T, it)
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 5)
    print("   %%s prune %%d block %%2d: %%.3f ms" %% (name, prune, block, best), flush=True)
xy, _ = synth.room_sequence(9001, 10000, 360)
t360, s360 = torch.from_numpy(np.ascontiguousarray(xy[:-1])).cuda(), torch.from_numpy(np.ascontiguousarray(xy[1:])).cuda()
tar, src, _ = synth.icp_pairs(4001, 16384, 1080)
t1080, s1080 = torch.from_numpy(tar).cuda(), torch.from_numpy(src).cuda()
for prune, block in ((2, 8), (2, 16), (3, 8)):
    run("360", t360, s360, prune, block)
for prune, block in ((2, 16), (3, 8), (3, 16)):
    run("1080", t1080, s1080, prune, block)
''' % ROOT
import shutil
PKG = os.path.join(ROOT, "a-2d-lidar-based-slam-system-for-wheeled-mobile-robots_b200", "libb2slam.so")
for lib in sorted(glob.glob(os.path.join(ROOT, "build", "ab", "lib_*.so"))):
    print(os.path.basename(lib), flush=True)
    shutil.copyfile(lib, PKG)  # the package loads the in-tree library; each build runs in its own process
    subprocess.run([sys.executable, "-c", CHILD, lib])
