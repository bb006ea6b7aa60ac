"""cfg 4 and cfg 5 of BASELINE.json at full size (run alone or under torchrun):
   cfg 4: 1 M independent 1080-beam ICP pairs, sharded across the ranks (strong scaling)
   cfg 5: global 16384 x 16384 grid from 8 scan streams, streams split by rank, merged over peer memory
Prints one JSON line per config on rank 0.  Not part of bench.py's contract (those configs are parity / scaling
cases); the numbers are quoted in DESIGN.md."""
import argparse, json, math, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import b2slam.dist as bdist
from b2slam import devapi, synth


def device_pairs(seed, pairs, beams, chunk=65536):
    """The synth.icp_pairs formulae evaluated with torch on the GPU (float32 result)."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    tar = torch.empty((pairs, 2, beams), dtype=torch.float32, device="cuda")
    src = torch.empty_like(tar)
    phi = torch.linspace(-math.pi, math.pi, beams, dtype=torch.float64, device="cuda")[None, :]
    for s in range(0, pairs, chunk):
        e = min(pairs, s + chunk)
        n = e - s
        u = lambda lo, hi, shape: lo + (hi - lo) * torch.rand(shape, generator=g, dtype=torch.float64, device="cuda")
        r0, amp, psi = u(3, 8, (n, 1)), u(0.5, 2, (n, 1)), u(0, 2 * math.pi, (n, 1))
        k = torch.randint(2, 6, (n, 1), generator=g, device="cuda").double()
        base = r0 + amp * torch.sin(k * phi + psi)
        noise = lambda: torch.randn((n, beams), generator=g, dtype=torch.float64, device="cuda") * 0.01
        rt = (base + noise()).clamp(0.10, 30.0)
        rs = (base + noise()).clamp(0.10, 30.0)
        tx, ty, th = u(-0.15, 0.15, (n, 1)), u(-0.15, 0.15, (n, 1)), u(-0.08, 0.08, (n, 1))
        tar[s:e, 0], tar[s:e, 1] = (rt * torch.cos(phi)).float(), (rt * torch.sin(phi)).float()
        sx, sy = rs * torch.cos(phi), rs * torch.sin(phi)
        c, sn = torch.cos(th), torch.sin(th)
        src[s:e, 0], src[s:e, 1] = (c * sx - sn * sy + tx).float(), (sn * sx + c * sy + ty).float()
    return tar, src


def cfg4(args, rank, world):
    lo, hi = bdist.shard_bounds(args.pairs, rank, world)
    tar, src = device_pairs(4001 + rank, hi - lo, 1080)
    T = torch.empty((hi - lo, 3, 3), dtype=torch.float64, device="cuda")
    it = torch.empty(hi - lo, dtype=torch.int32, device="cuda")
    devapi.icp_batch(tar[:4096], src[:4096], 30, 1e-3)
    torch.cuda.synchronize()
    bdist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    devapi.icp_batch(tar, src, 30, 1e-3, T, it)
    b.record()
    torch.cuda.synchronize()
    ms = bdist.max_over_ranks(a.elapsed_time(b))
    if rank == 0:
        print(json.dumps({"config": "cfg4: %d independent 1080-beam ICP pairs, strong scaling" % args.pairs, "n_gpus": world,
                          "pairs_per_s": args.pairs / (ms * 1e-3), "ms": ms, "mean_iterations": float(it.float().mean().item())}))


def cfg5(args, rank, world):
    G, streams, K, N = 16384, 8, args.scans, 1080
    lo, hi = bdist.shard_bounds(streams, rank, world)
    parts = [synth.grid_scans(5001 + s, K, N, half_extent_m=380.0) for s in range(lo, hi)]
    dev = [torch.from_numpy(np.concatenate([p[k] for p in parts])).cuda() for k in range(4)]
    sm = bdist.ShardedMappingP2P(G, G, 0.05)
    sm.update_device(*dev)
    torch.cuda.synchronize()
    bdist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    a.record()
    for _ in range(reps):
        sm.update_device(*dev)
    b.record()
    torch.cuda.synchronize()
    ms = bdist.max_over_ranks(a.elapsed_time(b)) / reps
    occ = int((sm.pmap_dev == 100).sum().item())
    if rank == 0:
        print(json.dumps({"config": "cfg5: global 16384^2 grid from 8 scan streams x %d scans, streams split by rank, "
                                    "peer-memory merge" % K, "n_gpus": world, "beams_per_s": streams * K * N / (ms * 1e-3),
                          "ms_per_step": ms, "occupied_cells": occ}))
    sm.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1000000)
    ap.add_argument("--scans", type=int, default=4096)
    ap.add_argument("--only", default="both", choices=["both", "cfg4", "cfg5"])
    args = ap.parse_args()
    rank, local, world = bdist.init()
    if args.only in ("both", "cfg4"):
        cfg4(args, rank, world)
    if args.only in ("both", "cfg5"):
        cfg5(args, rank, world)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
