// Standalone ICP ops (findNearest [ICP]:90-114, getTransform [ICP]:149-179) and the tuning switches of the
// batched kernel, which lives in b2s_icp_kernel.cuh (instantiated in b2s_icp_f32.cu / b2s_icp_f64.cu).
#include "b2s_icp_kernel.cuh"

namespace b2s {

// ICP.findNearest as a standalone op: src [n][2], tar [m][2] (point rows).  Targets are tiled
// through shared memory; each thread owns one source point.
constexpr int NN_TILE = 1024;

__global__ void __launch_bounds__(128)
nearest_kernel(const double *__restrict__ src_xy, int n, const double *__restrict__ tar_xy, int m,
               double *__restrict__ dist_out, int64_t *__restrict__ idx_out)
{
    __shared__ double2 tile[NN_TILE];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double sx = 0.0, sy = 0.0;
    if (i < n) {
        sx = src_xy[2 * i];
        sy = src_xy[2 * i + 1];
    }
    double best = INFINITY;
    long long arg = 0;
    for (int base = 0; base < m; base += NN_TILE) {
        const int cnt = min(NN_TILE, m - base);
        __syncthreads();
        for (int j = threadIdx.x; j < cnt; j += blockDim.x)
            tile[j] = make_double2(tar_xy[2 * (size_t)(base + j)], tar_xy[2 * (size_t)(base + j) + 1]);
        __syncthreads();
        for (int j = 0; j < cnt; ++j) {
            // the reference's own expression ([ICP]:102: np.linalg.norm of the difference), square root included:
            // sqrt merges neighbouring doubles, and the ties it creates go to the lower index
            const double d = ref_dist(sx - tile[j].x, sy - tile[j].y);
            if (d < best) {
                best = d;
                arg = base + j;
            }
        }
    }
    if (i < n) {
        dist_out[i] = (best == INFINITY) ? 0.0 : best;
        idx_out[i] = arg;
    }
}

// ICP.getTransform as a standalone op: one CTA, grid-stride over n row-matched points.
__global__ void __launch_bounds__(256)
rigid_fit_kernel(const double *__restrict__ src_xy, const double *__restrict__ tar_xy, int n,
                 double *__restrict__ T_out)
{
    __shared__ double scratch[4 * 32];
    double s4[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        s4[0] += src_xy[2 * i];
        s4[1] += src_xy[2 * i + 1];
        s4[2] += tar_xy[2 * i];
        s4[3] += tar_xy[2 * i + 1];
    }
    block_sum<4>(s4, scratch);
    const double inv = 1.0 / (double)n;
    const double cax = s4[0] * inv, cay = s4[1] * inv, cbx = s4[2] * inv, cby = s4[3] * inv;
    double w4[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double px = src_xy[2 * i] - cax, py = src_xy[2 * i + 1] - cay;
        const double qx = tar_xy[2 * i] - cbx, qy = tar_xy[2 * i + 1] - cby;
        w4[0] = fma(qx, px, w4[0]);
        w4[1] = fma(qx, py, w4[1]);
        w4[2] = fma(qy, px, w4[2]);
        w4[3] = fma(qy, py, w4[3]);
    }
    block_sum<4>(w4, scratch);
    if (threadIdx.x == 0) {
        double T[6];
        rotation_from_w(w4[0], w4[1], w4[2], w4[3], cax, cay, cbx, cby, T);
        T_out[0] = T[0]; T_out[1] = T[1]; T_out[2] = T[2];
        T_out[3] = T[3]; T_out[4] = T[4]; T_out[5] = T[5];
        T_out[6] = 0.0;  T_out[7] = 0.0;  T_out[8] = 1.0;
    }
}

int g_icp_src_per_thread = 0;  // 0: choose per problem size; 2..4 force (tuning hook)
int g_icp_layout = 2;          // point groups per warp: 2 balanced when that takes no extra group (default); 1 balanced; 0 strided (tuning hook)
int g_icp_prune = 4;           // exact search: 4 queued (default: both pruning tests, the surviving point x block pairs spread over the lanes);
                               // 2 collective, warp-level + per-lane pruning; 3 warp-level only; 1 per-lane only; 0 plain brute force
int g_icp_block = 0;           // targets per pruning block: 0 chooses per mode and scan size (8 up to 600 targets, 16 above); 8 / 16 / 32 force

}  // namespace b2s

using namespace b2s;

extern "C" int b2s_nearest_f64(const double *src_xy, int n, const double *tar_xy, int m,
                               double *dist_out, int64_t *idx_out, void *stream)
{
    B2S_REQUIRE(n >= 0 && m >= 0, "b2s_nearest: negative size");
    if (n == 0) return B2S_OK;
    B2S_REQUIRE(src_xy && dist_out && idx_out && (tar_xy || m == 0), "b2s_nearest: null pointer");
    nearest_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(src_xy, n, tar_xy, m, dist_out, idx_out);
    B2S_CUDA(cudaGetLastError());
    return B2S_OK;
}

extern "C" int b2s_rigid_fit_f64(const double *src_xy, const double *tar_xy, int n, double *T_out,
                                 void *stream)
{
    B2S_REQUIRE(n > 0, "b2s_rigid_fit: need at least one point");
    B2S_REQUIRE(src_xy && tar_xy && T_out, "b2s_rigid_fit: null pointer");
    rigid_fit_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(src_xy, tar_xy, n, T_out);
    B2S_CUDA(cudaGetLastError());
    return B2S_OK;
}
