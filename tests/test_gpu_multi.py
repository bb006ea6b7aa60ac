"""Multi-GPU path on real devices (world sizes 2, 4, 8; skipped where the box has fewer GPUs): torchrun, NCCL all-reduce
of the int32 count deltas, the peer-memory merge in all its forms (flag-synchronised, NCCL-fenced, dense), rank-sharded
ICP pairs.  Every merged grid must be bit-identical to ONE pass over all streams (the CPU oracle).  The single-GPU emulation of the same logic is in
test_gpu_grid.py::test_cfg3_full_size_properties; the gloo version in test_dist_gloo.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

WORKER = r'''
import os, sys
import numpy as np, torch
sys.path.insert(0, os.environ["B2S_ROOT"])
import b2slam.dist as bdist
from b2slam import devapi, synth
rank, local, world = bdist.init()
G = 2048
sm = bdist.ShardedMapping(G, G, 0.05)
streams = 2 * world
lo, hi = bdist.shard_bounds(streams, rank, world)
for rnd in range(2):
    parts = [synth.grid_scans(5001 + s + 100 * rnd, 24, 1080, half_extent_m=40.0) for s in range(lo, hi)]
    ox, oy, cx, cy = (np.concatenate([p[k] for p in parts]) for k in range(4))
    pm = sm.update_batch(ox, oy, cx, cy)
hit, miss = sm.counts()
res = {}
ROUNDS = 2
# tile-sparse merge synchronised by flag words in peer memory (default), the same kernel fenced by NCCL, the dense kernel
for tag, sparse, fence in (("2", True, "flags"), ("3", False, "nccl"), ("4", True, "nccl")):
    p2p = bdist.ShardedMappingP2P(G, G, 0.05, sparse=sparse, fence=fence)
    assert p2p.flags_mode == (fence == "flags")
    for rnd in range(ROUNDS):
        parts = [synth.grid_scans(5001 + s + 100 * rnd, 24, 1080, half_extent_m=40.0) for s in range(lo, hi)]
        ox, oy, cx, cy = (np.concatenate([p[k] for p in parts]) for k in range(4))
        res["pm" + tag] = p2p.update_batch(ox, oy, cx, cy).copy()
    res["hit" + tag], res["miss" + tag] = p2p.counts()
    if fence == "flags":
        # many short steps back to back, device-resident inputs: the epochs must keep the ranks in lock step without
        # any host-side collective (a rank that ran ahead would clear planes a peer still reads)
        small = [torch.from_numpy(a).cuda() for a in synth.grid_scans(900 + rank, 8, 1080, half_extent_m=40.0)]
        for k in range(40):
            p2p.update_device(*small)
        p2p.check()
        res["hit5"], res["miss5"] = p2p.counts()
        torch.cuda.synchronize()
        res["pm5"] = p2p.pmap_dev.cpu().numpy()
        # a NaN on ONE rank must raise on EVERY rank (the dropped-beam counts travel with the ready flags)
        bad = [t.clone() for t in small]
        if rank == world - 1:
            bad[1][3, 5] = float("nan")
        p2p.update_device(*bad)
        try:
            p2p.check()
            res["raised"] = np.array(0)
        except ValueError:
            res["raised"] = np.array(1)
    p2p.close()
# the library's own NCCL plumbing (dlopen'ed libnccl): communicator from a broadcast unique id, in-place all-reduce
import ctypes
from b2slam import _lib
L = _lib.lib()
uid = ctypes.create_string_buffer(128)
if rank == 0:
    _lib.check(L.b2s_nccl_unique_id(uid))
box = [uid.raw]
torch.distributed.broadcast_object_list(box, src=0)
comm = ctypes.c_void_p()
_lib.check(L.b2s_nccl_comm_init(ctypes.byref(comm), world, rank, ctypes.create_string_buffer(box[0], 128)))
ah = torch.full((64, 64), rank + 1, dtype=torch.int32, device="cuda")
am = torch.full((64, 64), 10 * (rank + 1), dtype=torch.int32, device="cuda")
_lib.check(L.b2s_grid_allreduce(ah.data_ptr(), am.data_ptr(), ah.numel(), comm, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
res["nccl_hit"] = ah.cpu().numpy()
res["nccl_miss"] = am.cpu().numpy()
_lib.check(L.b2s_nccl_comm_destroy(comm))
xy, _ = synth.room_sequence(9001, 41, 360)
plo, phi = bdist.sequence_pair_bounds(41, rank, world)
tar = torch.from_numpy(np.ascontiguousarray(xy[plo:phi])).cuda()
src = torch.from_numpy(np.ascontiguousarray(xy[plo + 1:phi + 1])).cuda()
T, it = devapi.icp_batch(tar, src)
counts = [b - a for a, b in (bdist.sequence_pair_bounds(41, q, world) for q in range(world))]
allT = bdist.gather_transforms(T, counts)
torch.cuda.synchronize()
np.savez(os.path.join(os.environ["B2S_OUT"], "rank%d.npz" % rank), hit=hit, miss=miss, pm=pm, T=allT.cpu().numpy(),
         **res)
bdist.barrier()
torch.distributed.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_grid_merge_and_icp_sharding(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs, the box has %d" % (world, torch.cuda.device_count()))
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, B2S_ROOT=ROOT, B2S_OUT=str(tmp_path))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29733 + world), str(script)]
    subprocess.run(cmd, check=True, env=env, timeout=600)
    from oracle import corc
    import b2slam.synth as synth
    G = 2048
    S, Hx, Hy = 20.0, G * 0.05 / 2.0, G * 0.05 / 2.0
    oh = np.zeros((G, G), dtype=np.int32)
    om = np.zeros((G, G), dtype=np.int32)
    for rnd in range(2):
        for s in range(2 * world):
            ox, oy, cx, cy = synth.grid_scans(5001 + s + 100 * rnd, 24, 1080, half_extent_m=40.0)
            corc.grid_raycast(oh, om, S, Hx, Hy, ox, oy, cx, cy)
    xy, _ = synth.room_sequence(9001, 41, 360)
    want_T, _ = corc.icp_batch(xy[:-1], xy[1:], 30, 1e-3)
    for r in range(world):
        z = np.load(tmp_path / ("rank%d.npz" % r))
        assert np.array_equal(z["hit"], oh) and np.array_equal(z["miss"], om)   # bit-identical to one pass
        assert np.array_equal(z["pm"], corc.grid_finalize(oh, om)[1])
        # fused peer-memory merge: same counts (sharded across ranks), same map on every rank
        for tag in ("2", "3", "4"):
            assert np.array_equal(z["hit" + tag], oh) and np.array_equal(z["miss" + tag], om), tag
            assert np.array_equal(z["pm" + tag], corc.grid_finalize(oh, om)[1]), tag
        assert int(z["raised"]) == 1, "rank %d did not see the NaN of the last rank" % r
    # the 40 back-to-back flag-synchronised steps: every rank's 8 scans, 40 times, on top of the two rounds
    o5h, o5m = oh.copy(), om.copy()
    for q in range(world):
        th = np.zeros((G, G), dtype=np.int32)
        tm = np.zeros((G, G), dtype=np.int32)
        corc.grid_raycast(th, tm, S, Hx, Hy, *synth.grid_scans(900 + q, 8, 1080, half_extent_m=40.0))
        o5h += 40 * th
        o5m += 40 * tm
    for r in range(world):
        z = np.load(tmp_path / ("rank%d.npz" % r))
        assert np.array_equal(z["hit5"], o5h) and np.array_equal(z["miss5"], o5m)
        assert np.array_equal(z["pm5"], corc.grid_finalize(o5h, o5m)[1])
    for r in range(world):
        z = np.load(tmp_path / ("rank%d.npz" % r))
        np.testing.assert_allclose(z["T"], want_T, rtol=0, atol=1e-9)
        assert (z["nccl_hit"] == sum(range(1, world + 1))).all() and (z["nccl_miss"] == 10 * sum(range(1, world + 1))).all()


def test_p2p_merge_single_rank_degenerates_to_finalize():
    """world = 1: the fused kernel is just accumulate + finalize (no peers), checked against the oracle."""
    import b2slam.dist as bdist
    import b2slam.synth as synth
    from oracle import corc
    for (xw, yw, sparse) in ((1024, 1024, True), (1024, 1024, False), (1000, 780, True)):
        sm = bdist.ShardedMappingP2P(xw, yw, 0.05, sparse=sparse)
        S, Hx, Hy = 20.0, xw * 0.05 / 2.0, yw * 0.05 / 2.0
        oh = np.zeros((xw, yw), dtype=np.int32)
        om = np.zeros((xw, yw), dtype=np.int32)
        for seed in (1, 2, 3):
            ox, oy, cx, cy = synth.grid_scans(seed, 40, 360, half_extent_m=15.0)
            pm = sm.update_batch(ox, oy, cx, cy).copy()
            corc.grid_raycast(oh, om, S, Hx, Hy, ox, oy, cx, cy)
        h, m = sm.counts()
        assert np.array_equal(h, oh) and np.array_equal(m, om), (xw, yw, sparse)
        assert np.array_equal(pm, corc.grid_finalize(oh, om)[1]), (xw, yw, sparse)
        # the delta planes hold only the last call's scans, and the dirty map covers every touched cell
        if sparse:
            dh, dm = sm.d_hit.cpu().numpy(), sm.d_miss.cpu().numpy()
            lh = np.zeros((xw, yw), dtype=np.int32)
            lm = np.zeros((xw, yw), dtype=np.int32)
            corc.grid_raycast(lh, lm, S, Hx, Hy, ox, oy, cx, cy)
            assert np.array_equal(dh, lh) and np.array_equal(dm, lm)
            dirty = sm.dirty.cpu().numpy().reshape(sm.tiles_x, sm.tiles_y).astype(bool)
            touched = (lh != 0) | (lm != 0)
            tx, ty = np.nonzero(touched)
            assert dirty[tx // 64, ty // 64].all()
        # fused ingestion and a forced pipeline depth on the same object: raw scans + poses, chunked copies
        import math
        from b2slam import scan
        sm.h2d_chunks = 3
        ranges, poses = synth.grid_scan_ranges(77, 40, 360, half_extent_m=15.0)
        pm = sm.update_scans(ranges, poses, -math.pi, math.pi).copy()
        corc.grid_raycast_ranges(oh, om, S, Hx, Hy, ranges, scan.pose_table(poses), scan.beam_table(-math.pi, math.pi, 360), 30.0)
        ox, oy, cx, cy = synth.grid_scans(4, 40, 360, half_extent_m=15.0)
        pm = sm.update_batch(ox, oy, cx, cy).copy()
        corc.grid_raycast(oh, om, S, Hx, Hy, ox, oy, cx, cy)
        h, m = sm.counts()
        assert np.array_equal(h, oh) and np.array_equal(m, om), (xw, yw, sparse, "pipelined")
        assert np.array_equal(pm, corc.grid_finalize(oh, om)[1]), (xw, yw, sparse, "pipelined")
        sm.close()


def test_streamed_submit_wait_equals_blocking_calls():
    """submit_scans / submit_batch + Ticket.wait with two steps in flight: every step's map equals the oracle's map of
    the scans up to and including that step (the read-back of step k overlaps step k + 1 but must never see it), the
    counts at the end are the sum of all steps, and an error on a step surfaces at ITS ticket."""
    import math
    import b2slam.dist as bdist
    import b2slam.synth as synth
    from b2slam import scan
    from oracle import corc
    G = 1024
    S, Hx, Hy = 20.0, G * 0.05 / 2.0, G * 0.05 / 2.0
    sm = bdist.ShardedMappingP2P(G, G, 0.05)
    oh = np.zeros((G, G), dtype=np.int32)
    om = np.zeros((G, G), dtype=np.int32)
    beams = scan.beam_table(-math.pi, math.pi, 360)
    want, tickets = [], []
    for k in range(7):
        if k % 2:
            ox, oy, cx, cy = synth.grid_scans(300 + k, 64, 360, half_extent_m=20.0)
            tickets.append(sm.submit_batch(ox, oy, cx, cy))
            corc.grid_raycast(oh, om, S, Hx, Hy, ox, oy, cx, cy)
        else:
            ranges, poses = synth.grid_scan_ranges(300 + k, 64, 360, half_extent_m=20.0)
            tickets.append(sm.submit_scans(ranges, poses, -math.pi, math.pi))
            corc.grid_raycast_ranges(oh, om, S, Hx, Hy, ranges, scan.pose_table(poses), beams, 30.0)
        want.append(corc.grid_finalize(oh, om)[1].copy())
        if k >= 1:                                   # two in flight: wait for the previous one only now
            got = tickets[k - 1].wait()
            assert np.array_equal(got, want[k - 1]), "step %d" % (k - 1)
    assert np.array_equal(tickets[-1].wait(), want[-1])
    h, m = sm.counts()
    assert np.array_equal(h, oh) and np.array_equal(m, om)
    # a bad step raises at its own ticket, the following step does not
    ox, oy, cx, cy = synth.grid_scans(999, 8, 360, half_extent_m=20.0)
    bad = oy.copy()
    bad[3, 5] = np.nan
    t_bad = sm.submit_batch(ox, bad, cx, cy)
    t_ok = sm.submit_batch(ox, oy, cx, cy)
    with pytest.raises(ValueError):
        t_bad.wait()
    t_ok.wait()
    sm.close()
