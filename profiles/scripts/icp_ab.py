"""ICP A/B on one GPU: search mode x block size x point layout, cfg 2 and cfg-4 shape; run once per library build
(B2S_LIB=path swaps in an alternative libb2slam.so by copying it over the packaged one before import)."""
import os, shutil, sys
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, root)
alt = os.environ.get("B2S_LIB")
if alt:
    import glob
    dst = glob.glob(os.path.join(root, "a-2d-*_b200"))[0] + "/libb2slam.so"
    shutil.copyfile(dst, dst + ".orig")
    shutil.copyfile(alt, dst)
import numpy as np, torch
from b2slam import _lib, devapi, synth
tune = _lib.lib().b2s_tune


def run(name, tar, src, combos):
    P = tar.shape[0]
    T = torch.empty((P, 3, 3), dtype=torch.float64, device="cuda")
    it = torch.empty(P, dtype=torch.int32, device="cuda")
    for prune, block, layout in combos:
        _lib.check(tune(b"icp_prune", prune)); _lib.check(tune(b"icp_block", block)); _lib.check(tune(b"icp_layout", layout))
        for _ in range(3):
            devapi.icp_batch(tar, src, 30, 1e-3, T, it)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                devapi.icp_batch(tar, src, 30, 1e-3, T, it)
            b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / 5)
        print("%s %s prune %d block %2d layout %d: %8.3f ms  %10.3e pairs/s  checksum %.12g iters %d"
              % (os.path.basename(alt or "default"), name, prune, block, layout, best, P / best * 1e3,
                 float(T.sum()), int(it.sum())), flush=True)


combos = [(p, b, l) for p in (2, 4) for b in (8, 16) for l in ((0, 1) if os.environ.get('B2S_AB_LAYOUTS') else (2,))]
xy, _ = synth.room_sequence(9001, 10000, 360)
run("cfg2/360 ", torch.from_numpy(np.ascontiguousarray(xy[:-1])).cuda(), torch.from_numpy(np.ascontiguousarray(xy[1:])).cuda(), combos)
tar, src, _ = synth.icp_pairs(4001, 16384, 1080)
run("cfg4/1080", torch.from_numpy(tar).cuda(), torch.from_numpy(src).cuda(), combos)
if alt:
    shutil.move(dst + ".orig", dst)
