// Adjacent steps of the hot path: odometry pose chain, virtual scan.
//
//   pose chain    [ICP]:185-190 / W9 localization.py:79-83: x += cos(th) T02 - sin(th) T12, y += ..., th += atan2(T10, T00)
//                 over a stream of per-pair transforms -> a prefix sum of the yaw increments followed by a
//                 prefix sum of the rotated translations (one CTA, blocked scan).
//   virtual scan  W9 localization.py:128-150 (laserEstimation): min range per bearing bin over the static-map
//                 obstacle cells -> one thread per obstacle, atomicMin on the float64 bit pattern.
#include "b2s_common.cuh"

#include <math.h>

namespace b2s {

constexpr int CHAIN_THREADS = 1024;

// Exclusive prefix over the CTA of one double per thread; returns this thread's offset, total in *total.
__device__ __forceinline__ double cta_exclusive_scan(double v, double *scratch, double *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += up;
    }
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        double w = scratch[lane];  // CHAIN_THREADS / 32 == 32 warps
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += up;
        }
        scratch[32 + lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const double before_warp = warp ? scratch[32 + warp - 1] : 0.0;
    *total = scratch[63];
    __syncthreads();
    return before_warp + inc - v;
}

__global__ void __launch_bounds__(CHAIN_THREADS)
pose_chain_kernel(const double *__restrict__ T, int pairs, double x0, double y0, double th0,
                  double *__restrict__ traj)
{
    __shared__ double scratch[64];
    const int tid = threadIdx.x;
    const int per = (pairs + CHAIN_THREADS - 1) / CHAIN_THREADS;
    const int lo = min(pairs, tid * per), hi = min(pairs, lo + per);
    // pass 1: yaw.  th_k = th0 + sum_{i<k} atan2(T10_i, T00_i)
    double part = 0.0;
    for (int k = lo; k < hi; ++k) part += atan2(T[9 * k + 3], T[9 * k + 0]);
    double total;
    double th = th0 + cta_exclusive_scan(part, scratch, &total);
    if (tid == 0) {
        traj[0] = x0;
        traj[1] = y0;
        traj[2] = th0;
    }
    // pass 2: position, each increment rotated by the yaw BEFORE the step ([ICP]:187-190)
    double px = 0.0, py = 0.0;
    double t2 = th;
    for (int k = lo; k < hi; ++k) {
        const double c = cos(t2), s = sin(t2);
        px += c * T[9 * k + 2] - s * T[9 * k + 5];
        py += s * T[9 * k + 2] + c * T[9 * k + 5];
        t2 += atan2(T[9 * k + 3], T[9 * k + 0]);
    }
    const double bx = x0 + cta_exclusive_scan(px, scratch, &total);
    const double by = y0 + cta_exclusive_scan(py, scratch, &total);
    double x = bx, y = by;
    for (int k = lo; k < hi; ++k) {
        const double c = cos(th), s = sin(th);
        x += c * T[9 * k + 2] - s * T[9 * k + 5];
        y += s * T[9 * k + 2] + c * T[9 * k + 5];
        th += atan2(T[9 * k + 3], T[9 * k + 0]);
        traj[3 * (k + 1)] = x;
        traj[3 * (k + 1) + 1] = y;
        traj[3 * (k + 1) + 2] = th;
    }
}

__global__ void __launch_bounds__(256)
virtual_scan_init_kernel(double *__restrict__ ranges, int beams, double far_range)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < beams) ranges[i] = far_range;
}

__global__ void __launch_bounds__(256)
virtual_scan_kernel(const double *__restrict__ obs_x, const double *__restrict__ obs_y, int count, double x,
                    double y, double yaw, double angle_min, double angle_increment, int beams,
                    double *__restrict__ ranges)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double ox = obs_x[i], oy = obs_y[i];
    const double dist = hypot(x - ox, y - oy);                                        // localization.py:138
    const double bin = (atan2(oy - y, ox - x) - angle_min - yaw) / angle_increment;    // localization.py:139
    if (!(fabs(bin) < 2.0e9) || !(dist == dist)) return;                               // int() would raise / overflow
    int index = (int)bin;                                                              // truncation toward zero
    index %= beams;                                                                    // the two while loops, :141-144
    if (index < 0) index += beams;
    // distances are >= 0, so the IEEE bit patterns order like the values: atomicMin on the bits
    atomicMin(reinterpret_cast<unsigned long long *>(ranges) + index, (unsigned long long)__double_as_longlong(dist));
}

}  // namespace b2s

using namespace b2s;

extern "C" int b2s_pose_chain(const double *T, int pairs, double x0, double y0, double th0, double *traj,
                              void *stream)
{
    B2S_REQUIRE(pairs >= 0 && traj && (T || pairs == 0), "b2s_pose_chain: bad arguments");
    pose_chain_kernel<<<1, CHAIN_THREADS, 0, (cudaStream_t)stream>>>(T, pairs, x0, y0, th0, traj);
    B2S_CUDA(cudaGetLastError());
    return B2S_OK;
}

extern "C" int b2s_virtual_scan(const double *obs_x, const double *obs_y, int count, double x, double y,
                                double yaw, double angle_min, double angle_increment, int beams,
                                double far_range, double *ranges, void *stream)
{
    B2S_REQUIRE(count >= 0 && beams > 0 && ranges && (count == 0 || (obs_x && obs_y)), "b2s_virtual_scan: bad arguments");
    B2S_REQUIRE(angle_increment != 0.0 && angle_increment == angle_increment, "b2s_virtual_scan: angle_increment");
    cudaStream_t st = (cudaStream_t)stream;
    virtual_scan_init_kernel<<<(beams + 255) / 256, 256, 0, st>>>(ranges, beams, far_range);
    B2S_CUDA(cudaGetLastError());
    if (count > 0) {
        virtual_scan_kernel<<<(count + 255) / 256, 256, 0, st>>>(obs_x, obs_y, count, x, y, yaw, angle_min,
                                                                 angle_increment, beams, ranges);
        B2S_CUDA(cudaGetLastError());
    }
    return B2S_OK;
}

// Host-buffer forms (the calls scan.py makes).
extern "C" int b2s_pose_chain_host(const double *T, int pairs, double x0, double y0, double th0, double *traj)
{
    B2S_REQUIRE(pairs >= 0 && traj && (T || pairs == 0), "b2s_pose_chain_host: bad arguments");
    ScratchPool *sp = scratch_pool();
    if (!sp) return cuda_fail(cudaErrorUnknown, "b2s_pose_chain_host: no device");
    int rc;
    if ((rc = sp->a.reserve((size_t)(pairs > 0 ? pairs : 1) * 72))) return rc;
    if ((rc = sp->b.reserve((size_t)(pairs + 1) * 24))) return rc;
    if (pairs > 0) B2S_CUDA(cudaMemcpyAsync(sp->a.p, T, (size_t)pairs * 72, cudaMemcpyHostToDevice, cudaStreamPerThread));
    if ((rc = b2s_pose_chain((const double *)sp->a.p, pairs, x0, y0, th0, (double *)sp->b.p, cudaStreamPerThread))) return rc;
    B2S_CUDA(cudaMemcpyAsync(traj, sp->b.p, (size_t)(pairs + 1) * 24, cudaMemcpyDeviceToHost, cudaStreamPerThread));
    B2S_CUDA(cudaStreamSynchronize(cudaStreamPerThread));
    return B2S_OK;
}

extern "C" int b2s_virtual_scan_host(const double *obs_x, const double *obs_y, int count, double x, double y,
                                     double yaw, double angle_min, double angle_increment, int beams,
                                     double far_range, double *ranges)
{
    B2S_REQUIRE(count >= 0 && beams > 0 && ranges && (count == 0 || (obs_x && obs_y)), "b2s_virtual_scan_host: bad arguments");
    ScratchPool *sp = scratch_pool();
    if (!sp) return cuda_fail(cudaErrorUnknown, "b2s_virtual_scan_host: no device");
    const size_t ob = (size_t)(count > 0 ? count : 1) * 8;
    int rc;
    if ((rc = sp->a.reserve(ob))) return rc;
    if ((rc = sp->b.reserve(ob))) return rc;
    if ((rc = sp->c.reserve((size_t)beams * 8))) return rc;
    if (count > 0) {
        B2S_CUDA(cudaMemcpyAsync(sp->a.p, obs_x, (size_t)count * 8, cudaMemcpyHostToDevice, cudaStreamPerThread));
        B2S_CUDA(cudaMemcpyAsync(sp->b.p, obs_y, (size_t)count * 8, cudaMemcpyHostToDevice, cudaStreamPerThread));
    }
    if ((rc = b2s_virtual_scan((const double *)sp->a.p, (const double *)sp->b.p, count, x, y, yaw, angle_min, angle_increment,
                               beams, far_range, (double *)sp->c.p, cudaStreamPerThread)))
        return rc;
    B2S_CUDA(cudaMemcpyAsync(ranges, sp->c.p, (size_t)beams * 8, cudaMemcpyDeviceToHost, cudaStreamPerThread));
    B2S_CUDA(cudaStreamSynchronize(cudaStreamPerThread));
    return B2S_OK;
}

// ---------------------------------------------------------------------------------------------
// Same-run FP64 issue peak (bench.py quotes the ICP kernel against it): 8 independent DFMA chains per thread.
namespace b2s {
__global__ void __launch_bounds__(256)
fp64_peak_kernel(double *out, int iters, double a, double b)
{
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (double)(threadIdx.x + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fma(v[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
    if (s == 123.456) out[0] = s;  // never true; keeps the chains alive
}
}  // namespace b2s

extern "C" int b2s_measure_fp64_peak(double *tflops_out)
{
    B2S_REQUIRE(tflops_out, "b2s_measure_fp64_peak: null pointer");
    double *d = nullptr;
    B2S_CUDA(cudaMalloc((void **)&d, 8));
    cudaEvent_t e0, e1;
    B2S_CUDA(cudaEventCreate(&e0));
    B2S_CUDA(cudaEventCreate(&e1));
    const int blocks = sm_count() * 8, iters = 4096;
    fp64_peak_kernel<<<blocks, 256>>>(d, iters, 0.999999, 1e-9);  // warm-up
    B2S_CUDA(cudaEventRecord(e0));
    fp64_peak_kernel<<<blocks, 256>>>(d, iters, 0.999999, 1e-9);
    B2S_CUDA(cudaEventRecord(e1));
    B2S_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    B2S_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *tflops_out = 2.0 * 8.0 * iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return B2S_OK;
}
