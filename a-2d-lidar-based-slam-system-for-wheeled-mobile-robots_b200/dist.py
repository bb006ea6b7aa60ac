"""Multi-GPU sharding of the hot path: one process per GPU.

ICP pairs are independent: contiguous blocks per rank, no data-path collective.  Grid streams
are split by rank into private int32 count-delta planes that are summed with one all-reduce
(integer sums commute, so the merged grid is bit-identical to a single-GPU pass).  The
functions work on CPU tensors with the gloo backend as well, which is how the host logic is
tested without GPUs.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment, (0, 0, 1) without it."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def init(backend=None):
    """Bind this process to its GPU and join the process group (no-op for world size 1)."""
    rank, local_rank, world = env_world()
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def shard_bounds(total, rank, world):
    """Contiguous block [lo, hi) of `total` units owned by `rank`; sizes differ by at most 1."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sequence_pair_bounds(scans, rank, world):
    """cfg 2: consecutive pairs (k-1 -> k), k = 1..scans-1, split in contiguous blocks.

    Returns (first_pair, last_pair_exclusive); the rank needs scans [first_pair, last_pair]
    inclusive, i.e. one scan before its first source scan.
    """
    return shard_bounds(max(int(scans) - 1, 0), rank, world)


def allreduce_counts(hit, miss):
    """Sum the per-rank count deltas in place (ncclAllReduce int32 sum over NVLink on GPUs)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(hit, op=dist.ReduceOp.SUM)
        dist.all_reduce(miss, op=dist.ReduceOp.SUM)
    return hit, miss


def gather_transforms(T_local, counts):
    """All-gather per-rank (p_r,3,3) transform blocks into the full (P,3,3) stack."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return T_local
    world = dist.get_world_size()
    biggest = max(counts)
    pad = torch.zeros((biggest, 3, 3), dtype=T_local.dtype, device=T_local.device)
    pad[:T_local.shape[0]] = T_local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def max_over_ranks(value, device=None):
    """Max of a Python float over ranks (timing reduction)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64,
                     device=device or ("cuda" if torch.cuda.is_available() else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


class ShardedMapping(object):
    """One rank's share of a global occupancy grid (cfg 5): host scans in, merged map out.

    Every call ray-casts this rank's scans into zeroed int32 delta planes, sums the deltas of
    all ranks (all-reduce), folds them into the rank's copy of the global counts and refreshes
    the occupancy map, so after each call every rank holds the same global grid -- bit-identical
    to one GPU processing all streams.
    """

    def __init__(self, xw, yw, xyreso, hit_weight=20.0, miss_weight=0.01, occ_threshold=10.0):
        from b2slam import devapi
        self._dev = devapi
        self.xw, self.yw = int(xw), int(yw)
        self.scale = devapi.grid_scale(self.xw, self.yw, float(xyreso))
        self.weights = (float(hit_weight), float(miss_weight), float(occ_threshold))
        self.hit, self.miss = devapi.new_planes(self.xw, self.yw)        # global counts
        self.d_hit, self.d_miss = devapi.new_planes(self.xw, self.yw)    # this call's deltas
        self.workspace = devapi.new_workspace(self.xw, self.yw)
        self.pmap_dev = torch.empty((self.xw, self.yw), dtype=torch.int8, device="cuda")
        self.pmap_host = torch.empty((self.xw, self.yw), dtype=torch.int8).pin_memory()
        self._in = None

    def update_batch(self, ox, oy, cx, cy):
        """ox, oy (K,N), cx, cy (K,) float32 host arrays (pinned for full copy speed)."""
        host = [torch.from_numpy(a) if not isinstance(a, torch.Tensor) else a for a in (ox, oy, cx, cy)]
        if self._in is None or self._in[0].shape != host[0].shape:
            self._in = [torch.empty(h.shape, dtype=torch.float32, device="cuda") for h in host]
        for d, h in zip(self._in, host):
            d.copy_(h, non_blocking=True)
        self.d_hit.zero_()
        self.d_miss.zero_()
        S, Hx, Hy = self.scale
        self._dev.grid_raycast(self.d_hit, self.d_miss, S, Hx, Hy, *self._in, workspace=self.workspace)
        allreduce_counts(self.d_hit, self.d_miss)
        self.hit += self.d_hit
        self.miss += self.d_miss
        w_hit, w_miss, thr = self.weights
        self._dev.grid_finalize(self.hit, self.miss, w_hit, w_miss, thr, pmap=self.pmap_dev)
        self.pmap_host.copy_(self.pmap_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.pmap_host.numpy()

    def counts(self):
        return self.hit.cpu().numpy(), self.miss.cpu().numpy()
