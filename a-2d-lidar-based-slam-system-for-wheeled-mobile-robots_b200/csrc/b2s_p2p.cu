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

// Tile-sparse form of the merge.  The ray-cast marks the 64 x 64-cell tiles it touches in a dirty map
// (one byte per tile, inside its workspace); the maps of all ranks are gathered first (that gather is
// also the fence that orders this kernel after every rank's ray-cast).  A CTA takes one tile of the
// rank's shard of tiles: if no rank dirtied it, nothing is read or written -- counts and occupancy are
// unchanged; otherwise only the ranks that did are read.  Cost follows the touched area, not the grid.
// The rank's shard of the global counts is stored tile-major: [tile - tile_lo][64][64].
__global__ void __launch_bounds__(256)
grid_merge_tiles_kernel(PeerPlanes pp, const uint8_t *__restrict__ all_dirty, int nranks, int ntiles, int tile_lo,
                        int tile_count, int xw, int yw, int32_t *__restrict__ g_hit, int32_t *__restrict__ g_miss,
                        double w_hit, double w_miss, double thresh)
{
    const int tiles_y = grid_tiles(yw);
    for (int t = blockIdx.x; t < tile_count; t += gridDim.x) {
        const int tile = tile_lo + t;
        unsigned who = 0;
        for (int r = 0; r < nranks; ++r) who |= (all_dirty[(size_t)r * ntiles + tile] ? 1u : 0u) << r;
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
            for (int r = 0; r < nranks; ++r)
                if (who & (1u << r)) {
                    add4(h, ld_stream(pp.hit[r] + cell));
                    add4(m, ld_stream(pp.miss[r] + cell));
                }
            *reinterpret_cast<int4 *>(gh + rx * GRID_TILE + qy) = h;
            *reinterpret_cast<int4 *>(gm + rx * GRID_TILE + qy) = m;
            const uint32_t occ = occupancy4(h, m, w_hit, w_miss, thresh);
            for (int r = 0; r < nranks; ++r) *reinterpret_cast<uint32_t *>(pp.pmap[r] + cell) = occ;
        }
    }
}

}  // namespace b2s

using namespace b2s;

extern "C" int b2s_grid_merge_p2p_tiles(const int32_t *const *delta_hit, const int32_t *const *delta_miss,
                                        int8_t *const *pmap, const uint8_t *all_dirty, int nranks, int xw, int yw,
                                        int tile_lo, int tile_hi, int32_t *global_hit_shard,
                                        int32_t *global_miss_shard, double w_hit, double w_miss, double thresh,
                                        void *stream)
{
    B2S_REQUIRE(delta_hit && delta_miss && pmap && all_dirty && global_hit_shard && global_miss_shard,
                "b2s_grid_merge_p2p_tiles: null pointer");
    B2S_REQUIRE(nranks >= 1 && nranks <= MAX_RANKS, "b2s_grid_merge_p2p_tiles: 1..16 ranks");
    B2S_REQUIRE(xw > 0 && yw > 0 && yw % 4 == 0, "b2s_grid_merge_p2p_tiles: yw must be a multiple of 4");
    const int ntiles = grid_tiles(xw) * grid_tiles(yw);
    B2S_REQUIRE(tile_lo >= 0 && tile_lo <= tile_hi && tile_hi <= ntiles, "b2s_grid_merge_p2p_tiles: tile range");
    PeerPlanes pp;
    for (int r = 0; r < nranks; ++r) {
        B2S_REQUIRE(delta_hit[r] && delta_miss[r] && pmap[r], "b2s_grid_merge_p2p_tiles: null plane");
        B2S_REQUIRE((uintptr_t)delta_hit[r] % 16 == 0 && (uintptr_t)delta_miss[r] % 16 == 0 && (uintptr_t)pmap[r] % 4 == 0,
                    "b2s_grid_merge_p2p_tiles: planes must be 16-byte aligned");
        pp.hit[r] = delta_hit[r];
        pp.miss[r] = delta_miss[r];
        pp.pmap[r] = pmap[r];
    }
    const int count = tile_hi - tile_lo;
    if (count == 0) return B2S_OK;
    int blocks = count;
    const int cap = sm_count() * 16;
    if (blocks > cap) blocks = cap;
    grid_merge_tiles_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pp, all_dirty, nranks, ntiles, tile_lo, count, xw,
                                                                     yw, global_hit_shard, global_miss_shard, w_hit,
                                                                     w_miss, thresh);
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
