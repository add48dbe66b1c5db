// bl_batch.cu -- batched versions of the reference's scalar kernels, used by the host-side mirrors of
// physics.toi_ball_ball (physics.py:10-76), physics.elastic_collision (:255-282),
// Disk/InfiniteWall.detect_collision and .resolve_collision (obstacles.py:70-76, :112-144).
// One thread per problem; inputs are AoS rows so a caller can pass numpy-shaped data unchanged.
#include "bl_arith.cuh"
#include "bl_internal.h"

namespace {

template <bool FMA>
__global__ void __launch_bounds__(256) toi_batch_kernel(const double *__restrict__ in, const double *__restrict__ rsum2,
                                                         const double t_eps, double *__restrict__ out, const long long m) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const double *r = in + 10 * q;
    out[q] = bl_toi_ball_ball<FMA>(r[0], r[1], r[2], r[3], r[5], r[6], r[7], r[8], rsum2[q], t_eps);
}

template <bool FMA>
__global__ void __launch_bounds__(256) elastic_batch_kernel(const double *__restrict__ in, double *__restrict__ out,
                                                             int32_t *__restrict__ status, const long long m) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const double *r = in + 10 * q;
    double w1x = 0, w1y = 0, w2x = 0, w2y = 0;
    const int st = bl_elastic<FMA>(r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], w1x, w1y, w2x, w2y);
    status[q] = st;
    out[4 * q] = w1x; out[4 * q + 1] = w1y; out[4 * q + 2] = w2x; out[4 * q + 3] = w2y;
}

template <bool FMA>
__global__ void __launch_bounds__(256) obstacle_detect_kernel(const bl_obstacle ob, const double *__restrict__ in,
                                                               const double *__restrict__ rsum2,
                                                               double *__restrict__ out_t, double *__restrict__ out_arg,
                                                               const long long m) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const double *r = in + 5 * q;
    double arg;
    const double rs = (ob.kind != BL_OBS_WALL && rsum2) ? rsum2[q] : 0.0;
    out_t[q] = bl_obstacle_toi<FMA>(ob, r[0], r[1], r[2], r[3], r[4], rs, arg);
    out_arg[q] = arg;
}

template <bool FMA>
__global__ void __launch_bounds__(256) obstacle_resolve_kernel(const bl_obstacle ob, const double *__restrict__ in,
                                                                const double *__restrict__ arg, double *__restrict__ out,
                                                                int32_t *__restrict__ status, const long long m) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const double *r = in + 5 * q;
    double wx = 0, wy = 0;
    status[q] = bl_obstacle_bounce<FMA>(ob, r[0], r[1], r[2], r[3], arg ? arg[q] : 0.0, wx, wy);
    out[2 * q] = wx; out[2 * q + 1] = wy;
}

template <bool FMA>
__global__ void __launch_bounds__(256) point_batch_kernel(const double *__restrict__ in, const double *__restrict__ r2,
                                                           const double t_eps, double *__restrict__ out, const long long m) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const double *r = in + 6 * q;
    out[q] = bl_toi_ball_point<FMA>(r[0], r[1], r[2], r[3], r2[q], r[4], r[5], t_eps);
}

template <bool FMA>
__global__ void __launch_bounds__(256) segment_param_kernel(const bl_obstacle sg, const double *__restrict__ in,
                                                             const double t_eps, double *__restrict__ out_t,
                                                             double *__restrict__ out_u, int32_t *__restrict__ out_code,
                                                             const long long m) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const double *r = in + 5 * q;
    double u;
    int code;
    out_t[q] = bl_toi_param_ball_segment<FMA>(sg, r[0], r[1], r[2], r[3], r[4], t_eps, u, code);
    out_u[q] = u;
    out_code[q] = code;
}

inline unsigned blocks(long long m) { return (unsigned)((m + 255) / 256); }

}  // namespace

int bl_launch_toi_batch(bl_handle *h, const double *d_in, const double *d_rsum2, double t_eps, double *d_out,
                        long long m, cudaStream_t st) {
    if (m <= 0) return BL_OK;
    if (h->dot_fma) toi_batch_kernel<true><<<blocks(m), 256, 0, st>>>(d_in, d_rsum2, t_eps, d_out, m);
    else toi_batch_kernel<false><<<blocks(m), 256, 0, st>>>(d_in, d_rsum2, t_eps, d_out, m);
    h->launches++;
    return bl_check_cuda(h, cudaGetLastError(), "launch toi_batch_kernel");
}

int bl_launch_elastic_batch(bl_handle *h, const double *d_in, double *d_out, int32_t *d_status, long long m,
                            cudaStream_t st) {
    if (m <= 0) return BL_OK;
    if (h->dot_fma) elastic_batch_kernel<true><<<blocks(m), 256, 0, st>>>(d_in, d_out, d_status, m);
    else elastic_batch_kernel<false><<<blocks(m), 256, 0, st>>>(d_in, d_out, d_status, m);
    h->launches++;
    return bl_check_cuda(h, cudaGetLastError(), "launch elastic_batch_kernel");
}

int bl_launch_obstacle_detect(bl_handle *h, bl_obstacle ob, const double *d_in, const double *d_rsum2, double *d_t,
                              double *d_arg, long long m, cudaStream_t st) {
    if (m <= 0) return BL_OK;
    if (h->dot_fma) obstacle_detect_kernel<true><<<blocks(m), 256, 0, st>>>(ob, d_in, d_rsum2, d_t, d_arg, m);
    else obstacle_detect_kernel<false><<<blocks(m), 256, 0, st>>>(ob, d_in, d_rsum2, d_t, d_arg, m);
    h->launches++;
    return bl_check_cuda(h, cudaGetLastError(), "launch obstacle_detect_kernel");
}

int bl_launch_obstacle_resolve(bl_handle *h, bl_obstacle ob, const double *d_in, const double *d_arg, double *d_out,
                               int32_t *d_status, long long m, cudaStream_t st) {
    if (m <= 0) return BL_OK;
    if (h->dot_fma) obstacle_resolve_kernel<true><<<blocks(m), 256, 0, st>>>(ob, d_in, d_arg, d_out, d_status, m);
    else obstacle_resolve_kernel<false><<<blocks(m), 256, 0, st>>>(ob, d_in, d_arg, d_out, d_status, m);
    h->launches++;
    return bl_check_cuda(h, cudaGetLastError(), "launch obstacle_resolve_kernel");
}

int bl_launch_point_batch(bl_handle *h, const double *d_in, const double *d_r2, double t_eps, double *d_out, long long m,
                          cudaStream_t st) {
    if (m <= 0) return BL_OK;
    if (h->dot_fma) point_batch_kernel<true><<<blocks(m), 256, 0, st>>>(d_in, d_r2, t_eps, d_out, m);
    else point_batch_kernel<false><<<blocks(m), 256, 0, st>>>(d_in, d_r2, t_eps, d_out, m);
    h->launches++;
    return bl_check_cuda(h, cudaGetLastError(), "launch point_batch_kernel");
}

int bl_launch_segment_param(bl_handle *h, bl_obstacle sg, const double *d_in, double t_eps, double *d_t, double *d_u,
                            int32_t *d_code, long long m, cudaStream_t st) {
    if (m <= 0) return BL_OK;
    if (h->dot_fma) segment_param_kernel<true><<<blocks(m), 256, 0, st>>>(sg, d_in, t_eps, d_t, d_u, d_code, m);
    else segment_param_kernel<false><<<blocks(m), 256, 0, st>>>(sg, d_in, t_eps, d_t, d_u, d_code, m);
    h->launches++;
    return bl_check_cuda(h, cudaGetLastError(), "launch segment_param_kernel");
}

// ---- measurement: the FP64 pipe's FMA peak, the denominator of the ensemble kernel's roofline -------------------
namespace {
__global__ void __launch_bounds__(256) bl_fp64_peak_kernel(double *out, const double y, const int iters) {
    double a[8];
#pragma unroll
    for (int u = 0; u < 8; u++) a[u] = (double)(threadIdx.x + u);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) a[u] = __fma_rn(a[u], y, y);
    }
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < 8; u++) s += a[u];
    if (s == 12345.678) out[0] = s;  // keeps the chain alive, never true
}
}  // namespace

int bl_launch_fp64_peak(bl_handle *h, double *tflops, cudaStream_t st) {
    double *d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_out, 8);
    if (e != cudaSuccess) return bl_check_cuda(h, e, "cudaMalloc");
    const int iters = 1 << 15, ctas = h->sm_count * 8, threads = 256;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {  // first one warms up
        cudaEventRecord(h->ev0, st);
        bl_fp64_peak_kernel<<<ctas, threads, 0, st>>>(d_out, 0.999999, iters);
        cudaEventRecord(h->ev1, st);
        h->launches++;
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev0, h->ev1);
        if (rep && ms < best) best = ms;
    }
    cudaFree(d_out);
    if (e != cudaSuccess) return bl_check_cuda(h, e, "bl_fp64_peak_kernel");
    *tflops = 2.0 * 8.0 * iters * (double)ctas * threads / (best * 1e-3) / 1e12;
    return BL_OK;
}
