// bl_table.cu -- time-of-impact table maintenance outside the event loop.
//
//   bl_recompute_kernel      add_ball's new row (simulation.py:176-209) and recompute_toi (:266-281): for each listed
//                            ball, every pair with it is re-solved at the current time and stored in both triangles of
//                            the symmetric table; its next-obstacle entry is refreshed.
//   bl_symmetrize_kernel     the event kernel rewrites only the row of a colliding ball; this pass copies every
//                            current cell (row of the ball with the larger sequence number) onto its mirror.
//   bl_rebuild_cover_kernel  per-ball candidate minima for the event kernel (full-row minimum, one warp per row).
//   bl_row_minima_kernel     the reference-shaped _balls_toi/_balls_idx (:85-86, :184-191): minimum of the lower
//                            triangle row, first index on ties, column 0 for an all-inf row, (inf,-1) for row 0.
//   bl_positions_kernel      _move (:556-560) as a read-only projection: pos = p0 + v*(time - t0).
#include "bl_arith.cuh"
#include "bl_internal.h"

namespace {

// idx == nullptr: the listed balls are range_first, range_first + 1, ... -- a pair of two listed balls is then solved by
// both of them (the solve is bit-symmetric in the two balls), so each stores only the cell of its own row: no
// strided stores into the other triangle when a whole table is built
template <bool FMA>
__global__ void __launch_bounds__(256) bl_recompute_kernel(const bl_system S, const double time,
                                                            const int32_t *__restrict__ idx, const int range_first,
                                                            const int range_count) {
    const int i = idx ? idx[blockIdx.y] : range_first + (int)blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= S.n) return;
    const size_t pitch = (size_t)S.pitch;
    const double ivx = S.vx[i], ivy = S.vy[i], it0 = S.t0[i];
    const double ipx = bl_advance(S.p0x[i], ivx, time, it0), ipy = bl_advance(S.p0y[i], ivy, time, it0);
    const int irc = S.rclass[i];
    if (j == i) {
        S.toi[(size_t)i * pitch + i] = BL_INF;
        double ot, oarg;
        int oi;
        bl_next_obstacle<FMA>(S.obstacles, S.n_obstacles, S.rsum2_obs, S.n_classes, irc, ipx, ipy, ivx, ivy,
                              S.radius[i], time, ot, oi, oarg);
        S.obs_t[i] = ot; S.obs_idx[i] = oi; S.obs_arg[i] = oarg;
        return;
    }
    const double jvx = S.vx[j], jvy = S.vy[j], jt0 = S.t0[j];
    const double jpx = bl_advance(S.p0x[j], jvx, time, jt0), jpy = bl_advance(S.p0y[j], jvy, time, jt0);
    const double rs = S.rsum2[(size_t)irc * S.n_classes + S.rclass[j]];
    // _detect_ball_collision: time + toi_ball_ball (bit-symmetric in the two balls)
    const double v = bl_add(time, bl_toi_ball_ball<FMA>(ipx, ipy, ivx, ivy, jpx, jpy, jvx, jvy, rs));
    S.toi[(size_t)i * pitch + j] = v;
    if (idx || j < range_first || j >= range_first + range_count) S.toi[(size_t)j * pitch + i] = v;
}

__device__ __forceinline__ void warp_min_key(uint64_t &ts, uint64_t &cd) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint64_t t2 = __shfl_xor_sync(0xffffffffu, ts, o), c2 = __shfl_xor_sync(0xffffffffu, cd, o);
        if (bl_key_less(t2, c2, ts, cd)) { ts = t2; cd = c2; }
    }
}

// one 32x32 tile pair (bi <= bj) per block: tile A = rows of block bi x columns of block bj, tile B its mirror
__global__ void __launch_bounds__(256) bl_symmetrize_kernel(const bl_system S) {
    __shared__ double A[32][33], B[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bi > bj) return;
    const int tx = threadIdx.x, ty = threadIdx.y, n = S.n;
    const size_t pitch = (size_t)S.pitch;
    const bool diag = bi == bj;
    for (int r = ty; r < 32; r += 8) {
        const int i = bi * 32 + r, j = bj * 32 + tx;
        A[r][tx] = (i < n && j < n) ? S.toi[(size_t)i * pitch + j] : BL_INF;
        if (!diag) {
            const int i2 = bj * 32 + r, j2 = bi * 32 + tx;
            B[r][tx] = (i2 < n && j2 < n) ? S.toi[(size_t)i2 * pitch + j2] : BL_INF;
        }
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int i = bi * 32 + r, j = bj * 32 + tx;
        if (i >= n || j >= n || (diag && r >= tx)) continue;
        const long long si = S.seq[i], sj = S.seq[j];
        if (diag) {
            if (si > sj) A[tx][r] = A[r][tx];
            else if (sj > si) A[r][tx] = A[tx][r];
        } else {
            if (si > sj) B[tx][r] = A[r][tx];
            else if (sj > si) A[r][tx] = B[tx][r];
        }
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int i = bi * 32 + r, j = bj * 32 + tx;
        if (i < n && j < n) S.toi[(size_t)i * pitch + j] = A[r][tx];
        if (!diag) {
            const int i2 = bj * 32 + r, j2 = bi * 32 + tx;
            if (i2 < n && j2 < n) S.toi[(size_t)i2 * pitch + j2] = B[r][tx];
        }
    }
}

__global__ void __launch_bounds__(256) bl_rebuild_cover_kernel(const bl_system S) {
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= S.n) return;
    const double *row = S.toi + (size_t)k * S.pitch;
    uint64_t ts = ~0ull, cd = BL_CODE_END;
    for (int x = lane; x < S.n; x += 32) {
        const double v = row[x];
        if (x != k && v < BL_INF) {
            const uint64_t s2 = bl_sortable(v), c2 = bl_pair_code(k, x);
            if (bl_key_less(s2, c2, ts, cd)) { ts = s2; cd = c2; }
        }
    }
    warp_min_key(ts, cd);
    if (lane == 0) {
        if (cd == BL_CODE_END) { S.cover_t[k] = BL_INF; S.cover_code[k] = BL_CODE_NONE; }
        else { S.cover_t[k] = bl_unsortable(ts); S.cover_code[k] = cd; }
    }
}

__global__ void __launch_bounds__(256) bl_row_minima_kernel(const bl_system S, double *__restrict__ out_t,
                                                             int64_t *__restrict__ out_idx) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= S.n) return;
    if (i == 0) {
        if (lane == 0) { out_t[0] = BL_INF; out_idx[0] = -1; }
        return;
    }
    const double *row = S.toi + (size_t)i * S.pitch;
    uint64_t ts = ~0ull, cd = BL_CODE_END;
    for (int x = lane; x < i; x += 32) {
        const uint64_t s2 = bl_sortable(row[x]), c2 = (uint64_t)x;
        if (bl_key_less(s2, c2, ts, cd)) { ts = s2; cd = c2; }
    }
    warp_min_key(ts, cd);
    if (lane == 0) { out_t[i] = bl_unsortable(ts); out_idx[i] = (int64_t)cd; }
}

__global__ void __launch_bounds__(256) bl_positions_kernel(const bl_system S, const double time,
                                                            double *__restrict__ pos) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= S.n) return;
    const double t0 = S.t0[k];
    pos[2 * k] = bl_advance(S.p0x[k], S.vx[k], time, t0);
    pos[2 * k + 1] = bl_advance(S.p0y[k], S.vy[k], time, t0);
}

// everything the host mirrors of a Billiard in one pass: out = [4][n][2] = balls_position (at `time`), balls_velocity,
// balls_initial_position, balls_initial_time (both columns) -- one launch and one copy per frame of an animation loop
// d_time != NULL: the time is read from device memory (the clock an evolve launch earlier on the stream has left)
// out2 != NULL: a second copy, meant for page-locked host memory written over the bus (no copy operation behind the
// kernel); res_dst != NULL: the result block of the evolve launch goes the same way
__global__ void __launch_bounds__(256) bl_snapshot_kernel(const bl_system S, const double time_arg,
                                                          const double *__restrict__ d_time, double *__restrict__ out,
                                                          double *__restrict__ out2, const bl_result *__restrict__ res_src,
                                                          bl_result *__restrict__ res_dst) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = S.n;
    if (res_dst && blockIdx.x == 0) {
        static_assert(sizeof(bl_result) % 8 == 0, "bl_result is copied in 8-byte words");
        for (int w = threadIdx.x; w < (int)(sizeof(bl_result) / 8); w += blockDim.x)
            reinterpret_cast<unsigned long long *>(res_dst)[w] = reinterpret_cast<const unsigned long long *>(res_src)[w];
    }
    if (k >= n) return;
    const double time = d_time ? *d_time : time_arg;
    const double t0 = S.t0[k], p0x = S.p0x[k], p0y = S.p0y[k], vx = S.vx[k], vy = S.vy[k];
    double2 *o = reinterpret_cast<double2 *>(out);
    const double2 pos = make_double2(bl_advance(p0x, vx, time, t0), bl_advance(p0y, vy, time, t0));
    o[k] = pos;
    o[n + k] = make_double2(vx, vy);
    o[2 * n + k] = make_double2(p0x, p0y);
    o[3 * n + k] = make_double2(t0, t0);
    if (out2) {
        double2 *o2 = reinterpret_cast<double2 *>(out2);
        o2[k] = pos;
        o2[n + k] = make_double2(vx, vy);
        o2[2 * n + k] = make_double2(p0x, p0y);
        o2[3 * n + k] = make_double2(t0, t0);
    }
}

// recompute_toi's edit detection, simulation.py:255-260
__global__ void __launch_bounds__(256) bl_apply_edits_kernel(const bl_system S, const double time,
                                                              const double *__restrict__ pos,
                                                              const double *__restrict__ vel) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= S.n) return;
    const double vx = vel[2 * k], vy = vel[2 * k + 1], t0 = S.t0[k];
    const double ox = bl_advance(S.p0x[k], vx, time, t0), oy = bl_advance(S.p0y[k], vy, time, t0);
    const double px = pos[2 * k], py = pos[2 * k + 1];
    S.vx[k] = vx; S.vy[k] = vy;
    if (px != ox || py != oy) { S.t0[k] = time; S.p0x[k] = px; S.p0y[k] = py; }
}

}  // namespace

int bl_launch_apply_edits(bl_handle *h, const bl_system *s, double time, const double *d_pos, const double *d_vel,
                          cudaStream_t st) {
    if (s->n <= 0) return BL_OK;
    bl_apply_edits_kernel<<<(unsigned)((s->n + 255) / 256), 256, 0, st>>>(*s, time, d_pos, d_vel);
    h->launches++;
    return bl_check_cuda(h, cudaGetLastError(), "launch bl_apply_edits_kernel");
}

int bl_launch_recompute(bl_handle *h, const bl_system *s, double time, const int32_t *d_idx, int first, int k,
                        cudaStream_t st) {
    if (k <= 0 || s->n <= 0) return BL_OK;
    const int chunk = 32768;  // gridDim.y limit is 65535
    for (int off = 0; off < k; off += chunk) {
        const int kk = k - off < chunk ? k - off : chunk;
        dim3 grid((unsigned)((s->n + 255) / 256), (unsigned)kk, 1);
        // (a chunk of a range still skips the mirror stores for the whole range: every ball of it is solved)
        const int32_t *ix = d_idx ? d_idx + off : nullptr;
        if (h->dot_fma) bl_recompute_kernel<true><<<grid, 256, 0, st>>>(*s, time, ix, first + off, k - off);
        else bl_recompute_kernel<false><<<grid, 256, 0, st>>>(*s, time, ix, first + off, k - off);
        h->launches++;
    }
    return bl_check_cuda(h, cudaGetLastError(), "launch bl_recompute_kernel");
}

int bl_launch_symmetrize(bl_handle *h, const bl_system *s, cudaStream_t st) {
    if (s->n <= 0) return BL_OK;
    const unsigned nb = (unsigned)((s->n + 31) / 32);
    bl_symmetrize_kernel<<<dim3(nb, nb, 1), dim3(32, 8, 1), 0, st>>>(*s);
    h->launches++;
    return bl_check_cuda(h, cudaGetLastError(), "launch bl_symmetrize_kernel");
}

int bl_launch_rebuild_cover(bl_handle *h, const bl_system *s, cudaStream_t st) {
    if (s->n <= 0) return BL_OK;
    bl_rebuild_cover_kernel<<<(unsigned)((s->n + 7) / 8), 256, 0, st>>>(*s);
    h->launches++;
    return bl_check_cuda(h, cudaGetLastError(), "launch bl_rebuild_cover_kernel");
}

int bl_launch_row_minima(bl_handle *h, const bl_system *s, double *d_toi, int64_t *d_idx, cudaStream_t st) {
    if (s->n <= 0) return BL_OK;
    bl_row_minima_kernel<<<(unsigned)((s->n + 7) / 8), 256, 0, st>>>(*s, d_toi, d_idx);
    h->launches++;
    return bl_check_cuda(h, cudaGetLastError(), "launch bl_row_minima_kernel");
}

// the result block of an evolve / query launch into page-locked host memory, behind the launch on its stream (a copy
// operation in its place costs several microseconds more before the host's wait returns)
__global__ void bl_result_to_host_kernel(const bl_result *__restrict__ src, bl_result *__restrict__ dst) {
    for (int w = threadIdx.x; w < (int)(sizeof(bl_result) / 8); w += blockDim.x)
        reinterpret_cast<unsigned long long *>(dst)[w] = reinterpret_cast<const unsigned long long *>(src)[w];
}

int bl_launch_result_to_host(bl_handle *h, const bl_result *src, bl_result *dst_dev, cudaStream_t st) {
    bl_result_to_host_kernel<<<1, 64, 0, st>>>(src, dst_dev);
    return bl_check_cuda(h, cudaGetLastError(), "launch bl_result_to_host_kernel");
}

int bl_launch_snapshot(bl_handle *h, const bl_system *s, double time, const double *d_time, double *d_out, cudaStream_t st,
                       double *out2, const bl_result *res_src, bl_result *res_dst) {
    if (s->n <= 0) return BL_OK;
    bl_snapshot_kernel<<<(unsigned)((s->n + 255) / 256), 256, 0, st>>>(*s, time, d_time, d_out, out2, res_src, res_dst);
    h->launches++;
    return bl_check_cuda(h, cudaGetLastError(), "launch bl_snapshot_kernel");
}

int bl_launch_positions(bl_handle *h, const bl_system *s, double time, double *d_pos, cudaStream_t st) {
    if (s->n <= 0) return BL_OK;
    bl_positions_kernel<<<(unsigned)((s->n + 255) / 256), 256, 0, st>>>(*s, time, d_pos);
    h->launches++;
    return bl_check_cuda(h, cudaGetLastError(), "launch bl_positions_kernel");
}
