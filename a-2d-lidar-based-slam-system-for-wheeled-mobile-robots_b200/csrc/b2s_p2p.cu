// Multi-GPU merge of the per-rank count deltas over NVLink peer memory, fused with the finalize.
//
// No reference counterpart (the reference is single-process); SURVEY.md section 8e.  Each rank owns a
// contiguous shard of the grid cells.  ONE kernel per rank
//   reduce-scatter : loads its shard of every rank's int32 hit/miss delta planes (16-byte loads; the
//                    peers' planes are mapped through CUDA IPC, so these are NVLink reads)
//   accumulate     : adds the sums into the rank's shard of the global counts
//   finalize       : evidence rule of [MAP]:42-50 on the updated counts
//   all-gather     : stores the int8 occupancy of its shard into EVERY rank's map (NVLink writes)
// so the int32 sums never travel twice and the gathered payload is 1 byte per cell instead of 8.
// Integer sums commute: the result is bit-identical to one GPU processing all streams.
#include "b2s_common.cuh"

namespace b2s {

constexpr int MAX_RANKS = 16;

struct PeerPlanes {
    const int32_t *hit[MAX_RANKS];
    const int32_t *miss[MAX_RANKS];
    int8_t *pmap[MAX_RANKS];
};

__device__ __forceinline__ int4 ld_stream(const int32_t *p)
{
    int4 v;  // read-once data, possibly remote: do not allocate in L1
    asm volatile("ld.global.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

__device__ __forceinline__ uint32_t occupancy4(const int4 &h, const int4 &m, double w_hit, double w_miss, double thresh)
{
    const int hh[4] = {h.x, h.y, h.z, h.w};
    const int mm[4] = {m.x, m.y, m.z, m.w};
    uint32_t packed = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double v = __dadd_rn(__dmul_rn(w_miss, (double)mm[j]), __dmul_rn(w_hit, (double)hh[j]));
        const uint32_t pm = (hh[j] == 0 && mm[j] == 0) ? 50u : (v > thresh ? 100u : 0u);
        packed |= pm << (8 * j);
    }
    return packed;
}

__device__ __forceinline__ void add4(int4 &a, const int4 &b)
{
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}

// A CTA walks tiles of 4096 cells; in quarter j of a tile thread t owns the 4 cells at
// tile*4096 + j*1024 + 4*t, so every warp instruction reads 512 contiguous bytes (full sectors, one
// pass over each peer's plane) and writes 128 contiguous bytes of occupancy.  All loads of a tile
// (2 planes x NRANKS ranks x 4 quarters) are issued before the first add, which is what it takes to
// keep an NVLink read stream busy.
template <int NRANKS>
__global__ void __launch_bounds__(256)
grid_merge_p2p_kernel(PeerPlanes pp, int nranks_rt, long long cell_lo, long long tiles, int32_t *__restrict__ g_hit,
                      int32_t *__restrict__ g_miss, double w_hit, double w_miss, double thresh)
{
    const int nranks = NRANKS > 0 ? NRANKS : nranks_rt;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long loc0 = tile * 4096 + 4 * threadIdx.x;  // inside this rank's shard
        const long long cell0 = cell_lo + loc0;                // global cell index
        int4 h[4], m[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            h[j] = *reinterpret_cast<const int4 *>(g_hit + loc0 + 1024 * j);
            m[j] = *reinterpret_cast<const int4 *>(g_miss + loc0 + 1024 * j);
        }
        if (NRANKS > 0) {
            int4 dh[NRANKS > 0 ? NRANKS : 1][4], dm[NRANKS > 0 ? NRANKS : 1][4];
#pragma unroll
            for (int r = 0; r < NRANKS; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dh[r][j] = ld_stream(pp.hit[r] + cell0 + 1024 * j);
                    dm[r][j] = ld_stream(pp.miss[r] + cell0 + 1024 * j);
                }
#pragma unroll
            for (int r = 0; r < NRANKS; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    add4(h[j], dh[r][j]);
                    add4(m[j], dm[r][j]);
                }
        } else {
            for (int r = 0; r < nranks; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    add4(h[j], ld_stream(pp.hit[r] + cell0 + 1024 * j));
                    add4(m[j], ld_stream(pp.miss[r] + cell0 + 1024 * j));
                }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            *reinterpret_cast<int4 *>(g_hit + loc0 + 1024 * j) = h[j];
            *reinterpret_cast<int4 *>(g_miss + loc0 + 1024 * j) = m[j];
            const uint32_t occ = occupancy4(h[j], m[j], w_hit, w_miss, thresh);
            for (int r = 0; r < nranks; ++r) *reinterpret_cast<uint32_t *>(pp.pmap[r] + cell0 + 1024 * j) = occ;
        }
    }
}

// ------------------------------------------------------------------ flag words over peer memory
//
// The step needs two rendezvous between the ranks: "every rank's ray-cast is done and its dirty map is here"
// before the merge reads peer planes, and "every rank's merge is done" before the map is consumed and the delta
// planes are cleared.  Two NCCL micro-collectives per step cost more than the merge itself at 8 GPUs, so the
// rendezvous are plain words in CUDA-IPC-mapped memory instead: a rank PUSHES its epoch number into a slot of every
// peer's flag block (st.release.sys after a system-scope fence), and waits by spinning on its OWN block
// (ld.acquire.sys on local memory, no NVLink traffic while waiting).  Flag block of a rank, uint32 words:
//   [0, R)     ready[src]  epoch of the last ray-cast + dirty map published by rank src
//   [R, 2R)    done[src]   epoch of the last merge finished by rank src
//   [2R, 3R)   bad[src]    beams rank src dropped in that step (NaN / inf / over-long), 0 in a healthy run
//   [3R]       arrival counter of this rank's merge CTAs;  [3R + 1]  set to 1 if a wait timed out
// Epochs only grow, so a flag never has to be reset.

struct PeerSync {
    uint32_t *flags[MAX_RANKS];
    uint8_t *all_dirty[MAX_RANKS];  // every rank's [nranks][dirty_stride] table of gathered dirty maps
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// the gathered dirty maps are written by the peers while this GPU runs: never through the non-coherent path
__device__ __forceinline__ unsigned ld_u8_sys(const uint8_t *p)
{
    unsigned v;
    asm volatile("ld.relaxed.sys.global.u8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr unsigned long long SPIN_LIMIT_NS = 20ull * 1000 * 1000 * 1000;  // a dead peer must not hang the GPU

// Spin until *flag has reached `epoch` (wrap-safe).  On a timeout the error word is set and the wait gives up: the
// results of this step are then garbage, and the host sees the word (b2s_p2p_status).
__device__ __forceinline__ void wait_epoch(const uint32_t *flag, uint32_t epoch, uint32_t *err)
{
    if ((int32_t)(ld_acquire_sys(flag) - epoch) >= 0) return;
    const unsigned long long t0 = global_ns();
    while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
        __nanosleep(40);
        if (global_ns() - t0 > SPIN_LIMIT_NS) {
            *err = 1;
            return;
        }
    }
}

// After the ray-cast: CTA r copies this rank's dirty map into rank r's table and then raises ready[rank] there.
__global__ void __launch_bounds__(256)
p2p_publish_kernel(PeerSync ps, const uint8_t *__restrict__ dirty, const int32_t *__restrict__ counters, int nranks,
                   int rank, int dirty_stride, uint32_t epoch)
{
    const int r = blockIdx.x;
    const uint4 *src = reinterpret_cast<const uint4 *>(dirty);
    uint4 *dst = reinterpret_cast<uint4 *>(ps.all_dirty[r] + (size_t)rank * dirty_stride);
    for (int k = threadIdx.x; k < dirty_stride / 16; k += blockDim.x) dst[k] = src[k];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t bad = 0;
        if (counters) bad = (uint32_t)(counters[B2S_CNT_NONFINITE] + counters[B2S_CNT_OVERFLOW] + counters[B2S_CNT_TOO_LONG]);
        ps.flags[r][2 * nranks + rank] = bad;
        __threadfence_system();
        st_release_sys(ps.flags[r] + rank, epoch);
    }
}

// End of the step: wait until every rank has finished its merge (so this rank's map is complete and nobody reads
// its delta planes any more).
__global__ void __launch_bounds__(32)
p2p_wait_kernel(uint32_t *flags, int first, int count, uint32_t epoch, int err_word)
{
    if ((int)threadIdx.x < count) wait_epoch(flags + first + threadIdx.x, epoch, flags + err_word);
}

// Tile-sparse form of the merge.  The ray-cast marks the 64 x 64-cell tiles it touches in a dirty map
// (one byte per tile, inside its workspace); the maps of all ranks are gathered first.  A CTA takes one tile of the
// rank's shard of tiles: if no rank dirtied it, nothing is read or written -- counts and occupancy are
// unchanged; otherwise only the ranks that did are read.  Cost follows the touched area, not the grid.
// The rank's shard of the global counts is stored tile-major: [tile - tile_lo][64][64].
//
// NR > 0: the rank loop is unrolled and ALL peer loads of a pass (2 planes x NR ranks, 16 bytes each) are issued
// before the first add, which is what keeps an NVLink read stream busy; NR = 0 is the generic loop.
// SYNC: the kernel itself waits for every rank's ready flag (first thing every CTA does; the flags are raised by the
// peers' publish kernels, which depend on nothing this kernel does), and the last CTA to finish raises done[rank]
// in every rank's block -- no host-side collective brackets the launch.
template <int NR, bool SYNC>
__global__ void __launch_bounds__(256)
grid_merge_tiles_kernel(PeerPlanes pp, PeerSync ps, const uint8_t *all_dirty, int dirty_stride, int nranks_rt,
                        int rank, uint32_t epoch, int tile_lo, int tile_count, int xw, int yw,
                        int32_t *__restrict__ g_hit, int32_t *__restrict__ g_miss, double w_hit, double w_miss,
                        double thresh)
{
    const int nranks = NR > 0 ? NR : nranks_rt;
    if (SYNC) {
        uint32_t *mine = ps.flags[rank];
        if ((int)threadIdx.x < nranks) wait_epoch(mine + threadIdx.x, epoch, mine + 3 * nranks + 1);
        __syncthreads();
    }
    const int tiles_y = grid_tiles(yw);
    for (int t = blockIdx.x; t < tile_count; t += gridDim.x) {
        const int tile = tile_lo + t;
        unsigned who = 0;
        if (NR > 0) {
#pragma unroll
            for (int r = 0; r < NR; ++r) who |= (ld_u8_sys(all_dirty + (size_t)r * dirty_stride + tile) ? 1u : 0u) << r;
        } else {
            for (int r = 0; r < nranks; ++r) who |= (ld_u8_sys(all_dirty + (size_t)r * dirty_stride + tile) ? 1u : 0u) << r;
        }
        if (!who) continue;
        const int x0 = (tile / tiles_y) * GRID_TILE, y0 = (tile % tiles_y) * GRID_TILE;
        int32_t *gh = g_hit + (size_t)t * GRID_TILE * GRID_TILE, *gm = g_miss + (size_t)t * GRID_TILE * GRID_TILE;
        // 256 threads = 16 rows x 16 quads of 4 cells per pass; 4 passes cover the 64 rows
        const int qx = threadIdx.x >> 4, qy = (threadIdx.x & 15) * 4;
#pragma unroll
        for (int pass = 0; pass < 4; ++pass) {
            const int rx = qx + 16 * pass;
            const int x = x0 + rx, y = y0 + qy;
            if (x >= xw || y >= yw) continue;  // yw is a multiple of 4 here (checked by the launcher)
            const size_t cell = (size_t)x * yw + y;
            int4 h = *reinterpret_cast<const int4 *>(gh + rx * GRID_TILE + qy);
            int4 m = *reinterpret_cast<const int4 *>(gm + rx * GRID_TILE + qy);
            if (NR > 0) {
                int4 dh[NR > 0 ? NR : 1], dm[NR > 0 ? NR : 1];
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    dh[r] = dm[r] = make_int4(0, 0, 0, 0);
                    if (who & (1u << r)) {
                        dh[r] = ld_stream(pp.hit[r] + cell);
                        dm[r] = ld_stream(pp.miss[r] + cell);
                    }
                }
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    add4(h, dh[r]);
                    add4(m, dm[r]);
                }
            } else {
                for (int r = 0; r < nranks; ++r)
                    if (who & (1u << r)) {
                        add4(h, ld_stream(pp.hit[r] + cell));
                        add4(m, ld_stream(pp.miss[r] + cell));
                    }
            }
            *reinterpret_cast<int4 *>(gh + rx * GRID_TILE + qy) = h;
            *reinterpret_cast<int4 *>(gm + rx * GRID_TILE + qy) = m;
            const uint32_t occ = occupancy4(h, m, w_hit, w_miss, thresh);
            for (int r = 0; r < nranks; ++r) *reinterpret_cast<uint32_t *>(pp.pmap[r] + cell) = occ;
        }
    }
    if (SYNC) {
        // the map stores above must be visible system-wide before done[rank] is: fence, count the CTA in, and let
        // the last one raise the flags
        __shared__ int last;
        __threadfence_system();
        __syncthreads();
        uint32_t *mine = ps.flags[rank];
        if (threadIdx.x == 0) last = (atomicAdd(mine + 3 * nranks, 1u) == gridDim.x - 1);
        __syncthreads();
        if (last) {
            if (threadIdx.x == 0) mine[3 * nranks] = 0;  // re-armed for the next launch
            __threadfence_system();
            if ((int)threadIdx.x < nranks) st_release_sys(ps.flags[threadIdx.x] + nranks + rank, epoch);
        }
    }
}

}  // namespace b2s

using namespace b2s;

namespace b2s {
static int merge_tiles_launch(const int32_t *const *delta_hit, const int32_t *const *delta_miss, int8_t *const *pmap,
                              const uint8_t *all_dirty, int dirty_stride, uint32_t *const *flags, int nranks, int rank,
                              uint32_t epoch, int xw, int yw, int tile_lo, int tile_hi, int32_t *global_hit_shard,
                              int32_t *global_miss_shard, double w_hit, double w_miss, double thresh, void *stream)
{
    B2S_REQUIRE(delta_hit && delta_miss && pmap && all_dirty && global_hit_shard && global_miss_shard,
                "b2s_grid_merge_p2p_tiles: null pointer");
    B2S_REQUIRE(nranks >= 1 && nranks <= MAX_RANKS, "b2s_grid_merge_p2p_tiles: 1..16 ranks");
    B2S_REQUIRE(xw > 0 && yw > 0 && yw % 4 == 0, "b2s_grid_merge_p2p_tiles: yw must be a multiple of 4");
    const int ntiles = grid_tiles(xw) * grid_tiles(yw);
    B2S_REQUIRE(dirty_stride >= ntiles, "b2s_grid_merge_p2p_tiles: dirty stride smaller than the tile count");
    B2S_REQUIRE(tile_lo >= 0 && tile_lo <= tile_hi && tile_hi <= ntiles, "b2s_grid_merge_p2p_tiles: tile range");
    B2S_REQUIRE(!flags || (rank >= 0 && rank < nranks), "b2s_grid_merge_p2p_tiles: rank out of range");
    PeerPlanes pp;
    PeerSync ps;
    memset(&ps, 0, sizeof(ps));
    for (int r = 0; r < nranks; ++r) {
        B2S_REQUIRE(delta_hit[r] && delta_miss[r] && pmap[r], "b2s_grid_merge_p2p_tiles: null plane");
        B2S_REQUIRE((uintptr_t)delta_hit[r] % 16 == 0 && (uintptr_t)delta_miss[r] % 16 == 0 && (uintptr_t)pmap[r] % 4 == 0,
                    "b2s_grid_merge_p2p_tiles: planes must be 16-byte aligned");
        pp.hit[r] = delta_hit[r];
        pp.miss[r] = delta_miss[r];
        pp.pmap[r] = pmap[r];
        if (flags) {
            B2S_REQUIRE(flags[r], "b2s_grid_merge_p2p_tiles: null flag block");
            ps.flags[r] = flags[r];
        }
    }
    const int count = tile_hi - tile_lo;
    // with flags the kernel is also this rank's "merge done" signal: it runs even when the shard is empty
    if (count == 0 && !flags) return B2S_OK;
    int blocks = count > 0 ? count : 1;
    const int cap = sm_count() * 8;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
#define B2S_TILES(NR, SY)                                                                                          \
    grid_merge_tiles_kernel<NR, SY><<<blocks, 256, 0, st>>>(pp, ps, all_dirty, dirty_stride, nranks, rank, epoch,  \
                                                            tile_lo, count, xw, yw, global_hit_shard,             \
                                                            global_miss_shard, w_hit, w_miss, thresh)
    if (flags) {
        switch (nranks) {
        case 2: B2S_TILES(2, true); break;
        case 4: B2S_TILES(4, true); break;
        case 8: B2S_TILES(8, true); break;
        default: B2S_TILES(0, true); break;
        }
    } else {
        switch (nranks) {
        case 2: B2S_TILES(2, false); break;
        case 4: B2S_TILES(4, false); break;
        case 8: B2S_TILES(8, false); break;
        default: B2S_TILES(0, false); break;
        }
    }
#undef B2S_TILES
    B2S_CUDA(cudaGetLastError());
    return B2S_OK;
}
}  // namespace b2s

extern "C" int b2s_grid_merge_p2p_tiles(const int32_t *const *delta_hit, const int32_t *const *delta_miss,
                                        int8_t *const *pmap, const uint8_t *all_dirty, int nranks, int xw, int yw,
                                        int tile_lo, int tile_hi, int32_t *global_hit_shard,
                                        int32_t *global_miss_shard, double w_hit, double w_miss, double thresh,
                                        void *stream)
{
    const int ntiles = (xw > 0 && yw > 0) ? grid_tiles(xw) * grid_tiles(yw) : 0;
    return merge_tiles_launch(delta_hit, delta_miss, pmap, all_dirty, ntiles, nullptr, nranks, 0, 0, xw, yw, tile_lo,
                              tile_hi, global_hit_shard, global_miss_shard, w_hit, w_miss, thresh, stream);
}

extern "C" size_t b2s_p2p_flag_bytes(int nranks) { return nranks > 0 ? (size_t)(3 * nranks + 2) * sizeof(uint32_t) : 0; }

extern "C" size_t b2s_p2p_dirty_stride(int xw, int yw) { return (xw > 0 && yw > 0) ? grid_dirty_bytes(xw, yw) : 0; }

extern "C" int b2s_p2p_publish(const void *workspace, const int32_t *counters, uint8_t *const *all_dirty,
                               uint32_t *const *flags, int nranks, int rank, int xw, int yw, uint32_t epoch, void *stream)
{
    B2S_REQUIRE(workspace && all_dirty && flags, "b2s_p2p_publish: null pointer");
    B2S_REQUIRE(nranks >= 1 && nranks <= MAX_RANKS && rank >= 0 && rank < nranks, "b2s_p2p_publish: bad rank");
    B2S_REQUIRE(xw > 0 && yw > 0, "b2s_p2p_publish: grid size");
    PeerSync ps;
    memset(&ps, 0, sizeof(ps));
    for (int r = 0; r < nranks; ++r) {
        B2S_REQUIRE(all_dirty[r] && flags[r] && (uintptr_t)all_dirty[r] % 16 == 0, "b2s_p2p_publish: null / misaligned peer buffer");
        ps.all_dirty[r] = all_dirty[r];
        ps.flags[r] = flags[r];
    }
    p2p_publish_kernel<<<nranks, 256, 0, (cudaStream_t)stream>>>(ps, (const uint8_t *)workspace + GRID_WS_HEADER, counters,
                                                                 nranks, rank, (int)grid_dirty_bytes(xw, yw), epoch);
    B2S_CUDA(cudaGetLastError());
    return B2S_OK;
}

extern "C" int b2s_grid_merge_p2p_tiles_sync(const int32_t *const *delta_hit, const int32_t *const *delta_miss,
                                             int8_t *const *pmap, const uint8_t *all_dirty, uint32_t *const *flags,
                                             int nranks, int rank, uint32_t epoch, int xw, int yw, int tile_lo,
                                             int tile_hi, int32_t *global_hit_shard, int32_t *global_miss_shard,
                                             double w_hit, double w_miss, double thresh, void *stream)
{
    B2S_REQUIRE(flags, "b2s_grid_merge_p2p_tiles_sync: null flag table");
    const int stride = (xw > 0 && yw > 0) ? (int)grid_dirty_bytes(xw, yw) : 0;
    return merge_tiles_launch(delta_hit, delta_miss, pmap, all_dirty, stride, flags, nranks, rank, epoch, xw, yw, tile_lo,
                              tile_hi, global_hit_shard, global_miss_shard, w_hit, w_miss, thresh, stream);
}

extern "C" int b2s_p2p_status(const uint32_t *my_flags, int nranks, int *timed_out, int64_t *dropped_beams, void *stream)
{
    B2S_REQUIRE(my_flags && nranks >= 1 && nranks <= MAX_RANKS, "b2s_p2p_status: bad arguments");
    uint32_t h[3 * MAX_RANKS + 2];
    B2S_CUDA(cudaMemcpyAsync(h, my_flags, (size_t)(3 * nranks + 2) * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                             (cudaStream_t)stream));
    B2S_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (timed_out) *timed_out = (int)h[3 * nranks + 1];
    if (dropped_beams) {
        int64_t total = 0;
        for (int r = 0; r < nranks; ++r) total += h[2 * nranks + r];
        *dropped_beams = total;
    }
    return B2S_OK;
}

extern "C" int b2s_p2p_wait_done(uint32_t *my_flags, int nranks, uint32_t epoch, void *stream)
{
    B2S_REQUIRE(my_flags && nranks >= 1 && nranks <= MAX_RANKS, "b2s_p2p_wait_done: bad arguments");
    p2p_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(my_flags, nranks, nranks, epoch, 3 * nranks + 1);
    B2S_CUDA(cudaGetLastError());
    return B2S_OK;
}

extern "C" int b2s_grid_merge_p2p(const int32_t *const *delta_hit, const int32_t *const *delta_miss,
                                  int8_t *const *pmap, int nranks, size_t cell_lo, size_t cell_hi,
                                  int32_t *global_hit_shard, int32_t *global_miss_shard, double w_hit,
                                  double w_miss, double thresh, void *stream)
{
    B2S_REQUIRE(delta_hit && delta_miss && pmap && global_hit_shard && global_miss_shard, "b2s_grid_merge_p2p: null pointer");
    B2S_REQUIRE(nranks >= 1 && nranks <= MAX_RANKS, "b2s_grid_merge_p2p: 1..16 ranks");
    B2S_REQUIRE(cell_lo <= cell_hi && cell_lo % 4096 == 0 && cell_hi % 4096 == 0,
                "b2s_grid_merge_p2p: shard bounds must be multiples of 4096 cells");
    PeerPlanes pp;
    for (int r = 0; r < nranks; ++r) {
        B2S_REQUIRE(delta_hit[r] && delta_miss[r] && pmap[r], "b2s_grid_merge_p2p: null plane");
        B2S_REQUIRE((uintptr_t)delta_hit[r] % 16 == 0 && (uintptr_t)delta_miss[r] % 16 == 0 && (uintptr_t)pmap[r] % 16 == 0,
                    "b2s_grid_merge_p2p: planes must be 16-byte aligned");
        pp.hit[r] = delta_hit[r];
        pp.miss[r] = delta_miss[r];
        pp.pmap[r] = pmap[r];
    }
    B2S_REQUIRE((uintptr_t)global_hit_shard % 16 == 0 && (uintptr_t)global_miss_shard % 16 == 0,
                "b2s_grid_merge_p2p: shard accumulators must be 16-byte aligned");
    const long long tiles = (long long)((cell_hi - cell_lo) / 4096);
    if (tiles == 0) return B2S_OK;
    long long blocks = tiles;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
#define B2S_MERGE(NR)                                                                                      \
    grid_merge_p2p_kernel<NR><<<(unsigned)blocks, 256, 0, st>>>(pp, nranks, (long long)cell_lo, tiles,     \
                                                                 global_hit_shard, global_miss_shard, w_hit, \
                                                                 w_miss, thresh)
    switch (nranks) {
    case 1: B2S_MERGE(1); break;
    case 2: B2S_MERGE(2); break;
    case 4: B2S_MERGE(4); break;
    default: B2S_MERGE(0); break;
    }
#undef B2S_MERGE
    B2S_CUDA(cudaGetLastError());
    return B2S_OK;
}

// ------------------------------------------------------------------ device memory + CUDA IPC plumbing

extern "C" int b2s_device_alloc(void **out, size_t bytes)
{
    B2S_REQUIRE(out, "b2s_device_alloc: null pointer");
    *out = nullptr;
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        *out = nullptr;
        cuda_fail(e, "cudaMalloc");
        return e == cudaErrorMemoryAllocation ? B2S_ERR_NOMEM : B2S_ERR_CUDA;
    }
    return B2S_OK;
}

extern "C" int b2s_device_free(void *p)
{
    if (p) B2S_CUDA(cudaFree(p));
    return B2S_OK;
}

extern "C" int b2s_ipc_export(const void *dev_ptr, void *handle64)
{
    B2S_REQUIRE(dev_ptr && handle64, "b2s_ipc_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    cudaIpcMemHandle_t h;
    B2S_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)));
    memcpy(handle64, &h, sizeof(h));
    return B2S_OK;
}

extern "C" int b2s_ipc_open(const void *handle64, void **dev_ptr_out)
{
    B2S_REQUIRE(handle64 && dev_ptr_out, "b2s_ipc_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    *dev_ptr_out = nullptr;
    B2S_CUDA(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return B2S_OK;
}

extern "C" int b2s_ipc_close(void *dev_ptr)
{
    if (dev_ptr) B2S_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return B2S_OK;
}
