"""Per-stage device times of the multi-GPU grid step (run under torchrun on >= 2 GPUs)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import b2slam.dist as bdist
from b2slam import devapi, synth, _lib

rank, local, world = bdist.init()
G, K, N = 4096, 16384, 1080
host = synth.grid_scans(12001 + rank, K, N)
ox, oy, cx, cy = (torch.from_numpy(a).cuda() for a in host)
sm = bdist.ShardedMappingP2P(G, G, 0.05)
S, Hx, Hy = sm.scale
ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
acc = np.zeros(6)
for it in range(13):
    ev[0].record()
    sm.d_hit.zero_(); sm.d_miss.zero_()
    ev[1].record()
    devapi.grid_raycast(sm.d_hit, sm.d_miss, S, Hx, Hy, ox, oy, cx, cy, workspace=sm.workspace)
    ev[2].record()
    sm._fence()
    ev[3].record()
    w = sm.weights
    _lib.check(_lib.lib().b2s_grid_merge_p2p(sm._hit_ptrs, sm._miss_ptrs, sm._pmap_ptrs, world, sm.cell_lo, sm.cell_hi,
                                             sm.g_hit.data_ptr(), sm.g_miss.data_ptr(), w[0], w[1], w[2],
                                             torch.cuda.current_stream().cuda_stream))
    ev[4].record()
    sm._fence()
    ev[5].record()
    torch.cuda.synchronize()
    if it >= 3:
        acc += [ev[i].elapsed_time(ev[i + 1]) for i in range(5)] + [ev[0].elapsed_time(ev[5])]
acc /= 10
if rank == 0:
    print("zero %.3f  raycast+fold %.3f  fence %.3f  merge %.3f  fence %.3f  total %.3f ms" % tuple(acc))
# NCCL path for comparison
hit, miss = devapi.new_planes(G, G)
pmap = torch.empty((G, G), dtype=torch.int8, device="cuda")
ws = devapi.new_workspace(G, G)
acc = np.zeros(5)
for it in range(13):
    ev[0].record()
    hit.zero_(); miss.zero_()
    ev[1].record()
    devapi.grid_raycast(hit, miss, S, Hx, Hy, ox, oy, cx, cy, workspace=ws)
    ev[2].record()
    bdist.allreduce_counts(hit, miss)
    ev[3].record()
    devapi.grid_finalize(hit, miss, pmap=pmap)
    ev[4].record()
    torch.cuda.synchronize()
    if it >= 3:
        acc += [ev[i].elapsed_time(ev[i + 1]) for i in range(4)] + [ev[0].elapsed_time(ev[4])]
acc /= 10
if rank == 0:
    print("zero %.3f  raycast+fold %.3f  allreduce %.3f  finalize %.3f  total %.3f ms" % tuple(acc))
sm.close()
torch.distributed.destroy_process_group()
