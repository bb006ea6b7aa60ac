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


def gpu_numa_node(index):
    """NUMA node the GPU hangs off (sysfs numa_node of its PCI function), or None when the platform does not say."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(int(index))).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:        # NVML prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as fh:
            node = int(fh.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def _cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    return cpus


def bind_to_gpu_numa(index):
    """Pin this process to the CPUs of its GPU's NUMA node, BEFORE it allocates page-locked staging memory: pages are
    placed on the node of the thread that first touches them, and a host-to-device copy that has to cross the
    socket interconnect runs at a fraction of the PCIe rate once several ranks do it at the same time.  Only ever
    narrows the affinity the launcher gave us; a no-op on single-node hosts, or when sysfs / NVML do not answer,
    or with B2S_NUMA_BIND=0.  Returns the node, or None."""
    if os.environ.get("B2S_NUMA_BIND", "1") == "0":
        return None
    node = gpu_numa_node(index)
    if node is None:
        return None
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
            cpus = _cpulist(fh.read())
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def init(backend=None):
    """Bind this process to its GPU (and to the CPUs of that GPU's NUMA node) and join the process group (no-op for
    world size 1)."""
    rank, local_rank, world = env_world()
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
        if world > 1:
            bind_to_gpu_numa(local_rank)
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


class _Slot(object):
    """Buffers of one in-flight host-buffer step of ShardedMappingP2P (there are two)."""

    def __init__(self, xw, yw, flag_words):
        self.stage = torch.empty((xw, yw), dtype=torch.int8, device="cuda")       # private copy of the merged map
        self.pmap_host = torch.empty((xw, yw), dtype=torch.int8).pin_memory()
        self.flags_dev = torch.zeros(max(flag_words, 1), dtype=torch.int32, device="cuda")
        self.flags_host = torch.zeros(max(flag_words, 1), dtype=torch.int32).pin_memory()
        self.pose4_host = None
        self._dev = {}
        self.inputs_free = None
        self.done = None
        self.ticket = None

    def inputs(self, kind, shapes):
        """Device input buffers of this slot for the given call form and shapes (allocated on first use / on a change)."""
        have = self._dev.get(kind)
        if have is None or [tuple(t.shape) for t in have] != [tuple(sh) for sh in shapes]:
            dt = [torch.float32] * len(shapes) if kind == "xy" else [torch.float32, torch.float64]
            have = [torch.empty(sh, dtype=d, device="cuda") for sh, d in zip(shapes, dt)]
            self._dev[kind] = have
            if kind == "ranges":
                self.pose4_host = torch.empty(shapes[1], dtype=torch.float64).pin_memory()
        return have


class Ticket(object):
    """Handle of a submitted step (ShardedMappingP2P.submit_batch / submit_scans)."""

    def __init__(self, owner, slot):
        self._owner, self._slot = owner, slot

    def wait(self):
        """Block until this step's merged map is in host memory; raises what check() raises for this step.  The returned
        int8 (xw, yw) array is a view of a page-locked buffer that the second-next submit overwrites."""
        slot, owner = self._slot, self._owner
        if slot.ticket is self:
            slot.done.synchronize()
            slot.ticket = None
            owner.pmap_host = slot.pmap_host
            if owner.flags_mode:
                owner._raise_for(slot.flags_host.tolist())
        return slot.pmap_host.numpy()


class ShardedMappingP2P(object):
    """ShardedMapping with the merge fused over NVLink peer memory (b2s_grid_merge_p2p).

    Every rank ray-casts its scans into private int32 delta planes that its peers can read (CUDA
    IPC).  One kernel per rank then sums its shard of all ranks' deltas, adds the sums into its
    shard of the global counts, finalizes that shard and stores the int8 occupancy into every rank's
    map -- reduce-scatter, finalize and all-gather in one pass.  The global counts stay sharded
    (1/world of the grid per rank); every rank holds the full occupancy map.

    fence="flags" (default with sparse=True): no collective in the step.  The ranks push epoch words and their dirty
    maps into each other's CUDA-IPC-mapped buffers and the merge kernel itself waits for them (b2s_p2p_publish,
    b2s_grid_merge_p2p_tiles_sync, b2s_p2p_wait_done).  fence="nccl": the round-1 form, an all-gather of the dirty
    maps before the kernel and a one-element all-reduce after it, kept for comparison and for the dense kernel.
    """

    def __init__(self, xw, yw, xyreso, hit_weight=20.0, miss_weight=0.01, occ_threshold=10.0, sparse=True,
                 fence="flags"):
        import ctypes
        from b2slam import _lib, devapi
        self._lib, self._dev, self._ct = _lib, devapi, ctypes
        # sparse: tile-level dirty tracking -- zeroing and merging cost follows the touched area
        self.sparse = bool(sparse) and int(yw) % 4 == 0
        self.xw, self.yw = int(xw), int(yw)
        self.cells = self.xw * self.yw
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.scale = devapi.grid_scale(self.xw, self.yw, float(xyreso))
        self.weights = (float(hit_weight), float(miss_weight), float(occ_threshold))
        L = _lib.lib()
        if self.sparse:
            tx, ty = ctypes.c_int(0), ctypes.c_int(0)
            _lib.check(L.b2s_grid_tile_count(self.xw, self.yw, ctypes.byref(tx), ctypes.byref(ty)))
            self.tiles_x, self.tiles_y = tx.value, ty.value
            self.ntiles = self.tiles_x * self.tiles_y
            self.tile_lo, self.tile_hi = shard_bounds(self.ntiles, self.rank, self.world)
            self.cell_lo, self.cell_hi = 0, (self.tile_hi - self.tile_lo) * 4096   # tile-major shard storage
        else:
            blocks = self.cells // 4096
            if blocks * 4096 != self.cells:
                raise ValueError("the dense peer-memory merge needs xw*yw to be a multiple of 4096 cells")
            lo, hi = shard_bounds(blocks, self.rank, self.world)
            self.cell_lo, self.cell_hi = lo * 4096, hi * 4096
        self.flags_mode = self.sparse and fence == "flags"
        self.epoch = 0
        self._own = []
        ptrs = []
        sizes = [self.cells * 4, self.cells * 4, self.cells]
        if self.flags_mode:
            self.dirty_stride = int(L.b2s_p2p_dirty_stride(self.xw, self.yw))
            self.flag_words = int(L.b2s_p2p_flag_bytes(self.world)) // 4
            sizes += [self.world * self.dirty_stride, self.flag_words * 4]
        for nbytes in sizes:
            p = ctypes.c_void_p()
            _lib.check(L.b2s_device_alloc(ctypes.byref(p), nbytes))
            self._own.append(p.value)
            ptrs.append(p.value)
        self.d_hit = devapi.tensor_from_ptr(ptrs[0], (self.xw, self.yw), torch.int32)
        self.d_miss = devapi.tensor_from_ptr(ptrs[1], (self.xw, self.yw), torch.int32)
        self.pmap_dev = devapi.tensor_from_ptr(ptrs[2], (self.xw, self.yw), torch.int8)
        self.d_hit.zero_()
        self.d_miss.zero_()
        self.pmap_dev.fill_(50)
        if self.flags_mode:
            self.all_dirty = devapi.tensor_from_ptr(ptrs[3], (self.world * self.dirty_stride,), torch.uint8)
            self.flags = devapi.tensor_from_ptr(ptrs[4], (self.flag_words,), torch.int32)
            self.all_dirty.zero_()
            self.flags.zero_()
            self.counters = torch.zeros(_lib.CNT_WORDS, dtype=torch.int32, device="cuda")
        handles = []
        for p in ptrs:
            buf = ctypes.create_string_buffer(64)
            _lib.check(L.b2s_ipc_export(ctypes.c_void_p(p), buf))
            handles.append(buf.raw)
        everyone = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(everyone, handles)
        else:
            everyone[0] = handles
        self._opened = []
        nbuf = len(ptrs)
        table = [[0] * self.world for _ in range(nbuf)]
        for r in range(self.world):
            for k in range(nbuf):
                if r == self.rank:
                    table[k][r] = ptrs[k]
                else:
                    q = ctypes.c_void_p()
                    _lib.check(L.b2s_ipc_open(ctypes.create_string_buffer(everyone[r][k], 64), ctypes.byref(q)))
                    self._opened.append(q.value)
                    table[k][r] = q.value
        arr = ctypes.c_void_p * self.world
        self._hit_ptrs, self._miss_ptrs, self._pmap_ptrs = (arr(*table[k]) for k in range(3))
        if self.flags_mode:
            self._dirty_ptrs, self._flag_ptrs = arr(*table[3]), arr(*table[4])
        n = self.cell_hi - self.cell_lo
        self.g_hit = torch.zeros(max(n, 4096), dtype=torch.int32, device="cuda")
        self.g_miss = torch.zeros(max(n, 4096), dtype=torch.int32, device="cuda")
        self.workspace = devapi.new_workspace(self.xw, self.yw)
        if self.sparse:
            dptr = L.b2s_grid_workspace_dirty(ctypes.c_void_p(self.workspace.data_ptr()))
            self.dirty = devapi.tensor_from_ptr(dptr, (self.ntiles,), torch.uint8)
            if not self.flags_mode:
                self.all_dirty = torch.zeros(self.world * self.ntiles, dtype=torch.uint8, device="cuda")
        self._token = torch.zeros(1, dtype=torch.int32, device="cuda")
        # two slots of everything a host-buffer call owns (device input buffers, a private device copy of the merged
        # map, page-locked result buffers), so that call k + 1 can be submitted while call k is still being read back
        self._slots = [_Slot(self.xw, self.yw, self.flag_words if self.flags_mode else 0) for _ in range(2)]
        self._calls = 0
        self.pmap_host = self._slots[0].pmap_host          # the most recent synchronous result
        self._beam_key = None
        self._copy_stream = None
        self._d2h_stream = None
        self.h2d_chunks = 0   # 0: automatic pipeline depth of the host-buffer calls; > 0 forces it
        torch.cuda.synchronize()
        barrier()

    def _fence(self):
        if self.world > 1:
            dist.all_reduce(self._token)  # stream-ordered: completes only when every rank got here

    def _begin(self):
        """Bring the private delta planes back to zero (sparse: only the tiles the previous call dirtied)."""
        L = self._lib.lib()
        stream = torch.cuda.current_stream().cuda_stream
        if self.sparse:
            self._lib.check(L.b2s_grid_clear_dirty(self.d_hit.data_ptr(), self.d_miss.data_ptr(), self.xw, self.yw,
                                                   self.workspace.data_ptr(), stream))
            if self.flags_mode:
                self.counters.zero_()
        else:
            self.d_hit.zero_()
            self.d_miss.zero_()

    def _merge(self):
        """ONE peer-memory kernel per rank: reduce-scatter + accumulate + finalize + all-gather of the map."""
        L = self._lib.lib()
        stream = torch.cuda.current_stream().cuda_stream
        w_hit, w_miss, thr = self.weights
        if self.flags_mode:
            # no collective: publish (dirty map + ready flag to every rank) -> merge (waits for the ready flags itself,
            # its last CTA raises the done flags) -> wait for every rank's done flag
            self.epoch += 1
            ws = self._ct.c_void_p(self.workspace.data_ptr())
            self._lib.check(L.b2s_p2p_publish(ws, self.counters.data_ptr(), self._dirty_ptrs, self._flag_ptrs, self.world,
                                              self.rank, self.xw, self.yw, self.epoch, stream))
            self._lib.check(L.b2s_grid_merge_p2p_tiles_sync(
                self._hit_ptrs, self._miss_ptrs, self._pmap_ptrs, self.all_dirty.data_ptr(), self._flag_ptrs, self.world,
                self.rank, self.epoch, self.xw, self.yw, self.tile_lo, self.tile_hi, self.g_hit.data_ptr(),
                self.g_miss.data_ptr(), w_hit, w_miss, thr, stream))
            self._lib.check(L.b2s_p2p_wait_done(self.flags.data_ptr(), self.world, self.epoch, stream))
            return
        if self.sparse:
            if self.world > 1:   # the gather of the dirty maps is also the fence after every rank's ray-cast
                dist.all_gather_into_tensor(self.all_dirty, self.dirty)
            else:
                self.all_dirty.copy_(self.dirty)
            self._lib.check(L.b2s_grid_merge_p2p_tiles(
                self._hit_ptrs, self._miss_ptrs, self._pmap_ptrs, self.all_dirty.data_ptr(), self.world, self.xw,
                self.yw, self.tile_lo, self.tile_hi, self.g_hit.data_ptr(), self.g_miss.data_ptr(), w_hit, w_miss,
                thr, stream))
        else:
            self._fence()
            self._lib.check(L.b2s_grid_merge_p2p(
                self._hit_ptrs, self._miss_ptrs, self._pmap_ptrs, self.world, self.cell_lo, self.cell_hi,
                self.g_hit.data_ptr(), self.g_miss.data_ptr(), w_hit, w_miss, thr, stream))
        self._fence()

    def _counters(self):
        return self.counters if self.flags_mode else None

    def _raise_for(self, words):
        """words: the flag block (host ints) as it stood when a step's merge had finished."""
        R = self.world
        if words[3 * R + 1]:
            raise RuntimeError("a peer did not reach the merge within the time limit (rank %d)" % self.rank)
        bad = int(sum(int(w) & 0xffffffff for w in words[2 * R:3 * R]))
        if bad:
            raise ValueError("%d beam(s) with a NaN / inf coordinate or longer than the limit were dropped by some rank; "
                             "the merged grid of this step is incomplete" % bad)

    def check(self):
        """Raise on EVERY rank if any rank dropped beams in the last step (NaN / inf coordinates or over-long beams:
        what Mapping.update_batch raises ValueError / OverflowError for) or if a flag wait timed out.  The counts of
        all ranks travel with the ready flags, so this is a local read; it synchronizes the stream.  Unlike
        Mapping.update_batch the step is not rolled back: after an error the object must be rebuilt."""
        if not self.flags_mode:
            return
        t, bad = self._ct.c_int(0), self._ct.c_int64(0)
        self._lib.check(self._lib.lib().b2s_p2p_status(self.flags.data_ptr(), self.world, self._ct.byref(t),
                                                       self._ct.byref(bad), torch.cuda.current_stream().cuda_stream))
        if t.value:
            raise RuntimeError("a peer did not reach the merge within the time limit (rank %d)" % self.rank)
        if bad.value:
            raise ValueError("%d beam(s) with a NaN / inf coordinate or longer than the limit were dropped by some rank; "
                             "the merged grid of this step is incomplete" % bad.value)

    def update_device(self, ox, oy, cx, cy, events=None):
        """Scans already on the device (float32 CUDA tensors).  Leaves the merged map in pmap_dev.
        `events`: optional (start, stop) torch.cuda.Event pair recorded around the ray-cast."""
        self._begin()
        S, Hx, Hy = self.scale
        if events:
            events[0].record()
        self._dev.grid_raycast(self.d_hit, self.d_miss, S, Hx, Hy, ox, oy, cx, cy, counters=self._counters(),
                               workspace=self.workspace)
        if events:
            events[1].record()
        self._merge()

    def _submit(self, kind, host, shapes, scans, beams, make_raycast, prepare=None):
        """One host-buffer step, enqueued and NOT waited for: chunked H2D on a side stream overlapped with the ray-cast
        of the previous chunk, the merge, a device-side copy of the merged map into the slot's private buffer, and its
        read-back on a third stream.  Returns the slot's Ticket.  The next submit may follow immediately: its uploads
        and ray-casts run while this step's map is still crossing PCIe in the other direction.
        prepare(lo, hi) builds chunk [lo, hi) of a derived host table right before it is copied."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
            self._d2h_stream = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        slot = self._slots[self._calls % 2]
        self._calls += 1
        if slot.ticket is not None:          # its buffers are about to be reused
            slot.ticket.wait()
        dev = slot.inputs(kind, shapes)
        if slot.inputs_free is not None:     # the ray-casts that last read these device buffers (two submits ago)
            self._copy_stream.wait_event(slot.inputs_free)
        raycast = make_raycast(dev)
        # a step submitted while the previous one is still in flight has its whole upload hidden under that step's
        # ray-cast, so it goes as ONE launch; a step that starts on an idle device is cut into chunks (~2M beams) so that
        # the ray-cast can start after the first of them
        streaming = any(sl.ticket is not None for sl in self._slots if sl is not slot)
        nchunk = self.h2d_chunks or (1 if streaming else max(1, min(8, (scans * beams + (1 << 21) - 1) >> 21)))
        nchunk = max(1, min(nchunk, scans))
        bounds = [scans * k // nchunk for k in range(nchunk + 1)]
        self._begin()
        for k in range(nchunk):
            lo, hi = bounds[k], bounds[k + 1]
            if prepare is not None:
                prepare(slot, lo, hi)
            with torch.cuda.stream(self._copy_stream):
                for d, h in zip(dev, host(slot)):
                    d[lo:hi].copy_(h[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            main.wait_event(ev)
            raycast(lo, hi)
        slot.inputs_free = torch.cuda.Event()
        slot.inputs_free.record(main)
        self._merge()
        # the merged map is complete here (wait_done); peers overwrite pmap_dev only after this rank's NEXT ready flag,
        # which is behind this copy in stream order -- so the read-back below never sees a torn map
        slot.stage.copy_(self.pmap_dev, non_blocking=True)
        if self.flags_mode:
            slot.flags_dev.copy_(self.flags, non_blocking=True)
        merged = torch.cuda.Event()
        merged.record(main)
        with torch.cuda.stream(self._d2h_stream):
            self._d2h_stream.wait_event(merged)
            slot.pmap_host.copy_(slot.stage, non_blocking=True)
            if self.flags_mode:
                slot.flags_host.copy_(slot.flags_dev, non_blocking=True)
            slot.done = torch.cuda.Event()
            slot.done.record(self._d2h_stream)
        slot.ticket = Ticket(self, slot)
        return slot.ticket

    def submit_batch(self, ox, oy, cx, cy):
        """update_batch without the wait: returns a Ticket whose wait() gives the merged int8 occupancy of this step.
        Up to two steps may be in flight; the input arrays must stay untouched until the ticket has been waited for."""
        host_arrays = [torch.from_numpy(a) if not isinstance(a, torch.Tensor) else a for a in (ox, oy, cx, cy)]
        K, N = host_arrays[0].shape
        S, Hx, Hy = self.scale

        def make_raycast(d):
            def raycast(lo, hi):
                self._dev.grid_raycast(self.d_hit, self.d_miss, S, Hx, Hy, d[0][lo:hi], d[1][lo:hi], d[2][lo:hi],
                                       d[3][lo:hi], counters=self._counters(), workspace=self.workspace)
            return raycast
        return self._submit("xy", lambda slot: host_arrays, [tuple(h.shape) for h in host_arrays], K, N, make_raycast)

    def update_batch(self, ox, oy, cx, cy):
        """ox, oy (K,N), cx, cy (K,) float32 host arrays -> merged int8 occupancy (host, pinned)."""
        return self.submit_batch(ox, oy, cx, cy).wait()

    def submit_scans(self, ranges, poses, angle_min, angle_max, clamp_inf_to=30.0):
        """update_scans without the wait (see submit_batch)."""
        import numpy as np
        from b2slam import scan
        ranges = torch.from_numpy(np.ascontiguousarray(ranges, dtype=np.float32)) if not isinstance(ranges, torch.Tensor) else ranges
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
        K, N = ranges.shape
        if poses.shape[0] != K:
            raise ValueError("need one pose per scan, got %d poses for %d scans" % (poses.shape[0], K))
        key = (float(angle_min), float(angle_max), N)
        if self._beam_key != key:
            self._beam_cs = torch.from_numpy(scan.beam_table(angle_min, angle_max, N)).cuda()
            self._beam_key = key
        S, Hx, Hy = self.scale

        def prepare(slot, lo, hi):
            # u2T's cos / sin of the yaw (slam_ekf.py:130-137) for this chunk, while the previous one is in flight
            table = slot.pose4_host.numpy()
            table[lo:hi, 0:2] = poses[lo:hi, 0:2]
            np.cos(poses[lo:hi, 2], out=table[lo:hi, 2])
            np.sin(poses[lo:hi, 2], out=table[lo:hi, 3])

        def make_raycast(d):
            def raycast(lo, hi):
                self._dev.grid_raycast_ranges(self.d_hit, self.d_miss, S, Hx, Hy, d[0][lo:hi], d[1][lo:hi], self._beam_cs,
                                              clamp_inf_to, counters=self._counters(), workspace=self.workspace)
            return raycast
        return self._submit("ranges", lambda slot: [ranges, slot.pose4_host], [(K, N), (K, 4)], K, N, make_raycast, prepare)

    def update_scans(self, ranges, poses, angle_min, angle_max, clamp_inf_to=30.0):
        """Fused ingestion (Mapping.update_scans) for this rank's stream: ranges (K,N) float32 + poses (K,3) instead
        of world-frame endpoints, half the bytes over PCIe.  Returns the merged int8 occupancy (host, pinned)."""
        return self.submit_scans(ranges, poses, angle_min, angle_max, clamp_inf_to).wait()

    def counts(self):
        """Full (hit, miss) planes assembled from the ranks' shards (for checks; not a hot path)."""
        n = self.cell_hi - self.cell_lo
        parts_h = [None] * self.world
        parts_m = [None] * self.world
        mine = (self.g_hit[:n].cpu().numpy(), self.g_miss[:n].cpu().numpy())
        if self.world > 1:
            dist.all_gather_object(parts_h, mine[0])
            dist.all_gather_object(parts_m, mine[1])
        else:
            parts_h[0], parts_m[0] = mine
        import numpy as np
        if not self.sparse:
            return (np.concatenate(parts_h).reshape(self.xw, self.yw),
                    np.concatenate(parts_m).reshape(self.xw, self.yw))
        out = []
        for parts in (parts_h, parts_m):   # tile-major shards -> [x][y] plane
            tiles = np.concatenate(parts).reshape(self.tiles_x, self.tiles_y, 64, 64)
            full = tiles.transpose(0, 2, 1, 3).reshape(self.tiles_x * 64, self.tiles_y * 64)
            out.append(np.ascontiguousarray(full[:self.xw, :self.yw]))
        return out[0], out[1]

    def close(self):
        for slot in self._slots:
            if slot.ticket is not None:
                try:
                    slot.ticket.wait()
                except Exception:
                    pass
        torch.cuda.synchronize()
        barrier()
        L = self._lib.lib()
        for p in self._opened:
            L.b2s_ipc_close(self._ct.c_void_p(p))
        self._opened = []
        barrier()
        self.d_hit = self.d_miss = self.pmap_dev = None
        for p in self._own:
            L.b2s_device_free(self._ct.c_void_p(p))
        self._own = []
