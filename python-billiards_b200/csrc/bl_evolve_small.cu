// bl_evolve_small.cu -- the event loop for systems of at most 32 balls (Galperin's two balls, the pool table, Newton's
// cradle: the reference's own examples) as ONE WARP.
//
// Such systems are strictly serial -- consecutive collisions share balls, a batch would hold one collision -- so the
// grid-wide round machinery of bl_evolve.cu buys nothing and costs a few thousand cycles per collision.  Here lane i
// owns ball i (state of record in registers), the whole time-of-impact table sits in shared memory (both triangles
// current), and one collision is: row minima -> warp argmin -> every lane resolves the collision redundantly from
// shuffled states -> every other lane re-solves its two pairs -> the colliding lanes refresh their obstacle entry.
// No barrier but __syncwarp.  Same arithmetic, same keys, same stop rules as the big kernel; same reference lines:
// simulation.py:354-438, :440-502, :504-554, :556-560, :562-636.
#include "bl_arith.cuh"
#include "bl_internal.h"

namespace {

enum { KIND_BOTH = 0, KIND_BB = 1, KIND_BO = 2 };

struct SmallParams {
    bl_system s;
    double time, end_time;
    long long max_events;
    int first_kind;
    int query_only;
    bl_log log;
    bl_result *result;
};

__device__ __forceinline__ int warp_argmin_s(uint64_t ts, uint64_t code) {
    const unsigned FULL = 0xffffffffu;
    unsigned x = (unsigned)(ts >> 32);
    unsigned m = __reduce_min_sync(FULL, x);
    bool c = x == m;
    x = c ? (unsigned)ts : 0xffffffffu;
    m = __reduce_min_sync(FULL, x);
    c = c && x == m;
    x = c ? (unsigned)(code >> 32) : 0xffffffffu;
    m = __reduce_min_sync(FULL, x);
    c = c && x == m;
    x = c ? (unsigned)code : 0xffffffffu;
    m = __reduce_min_sync(FULL, x);
    c = c && x == m;
    return __ffs(__ballot_sync(FULL, c)) - 1;
}
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ uint64_t shfl_u(uint64_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }

template <bool FMA>
__global__ void __launch_bounds__(32, 1) bl_evolve_small_kernel(const __grid_constant__ SmallParams P) {
    __shared__ double T[32][33];
    __shared__ bl_obstacle s_obs[BL_SMALL_MAX_OBS];
    const bl_system &S = P.s;
    const int n = S.n, K = S.n_classes, n_obs = S.n_obstacles;
    const int lane = threadIdx.x;
    const bool own = lane < n;
    for (int q = lane; q < n_obs; q += 32) s_obs[q] = S.obstacles[q];
    double p0x = 0, p0y = 0, vx = 0, vy = 0, t0 = 0, mass = 0, radius = 0, oarg = 0;
    uint64_t ots = bl_sortable(BL_INF);
    int rc = 0, oidx = -1;
    if (own) {
        p0x = S.p0x[lane]; p0y = S.p0y[lane]; vx = S.vx[lane]; vy = S.vy[lane]; t0 = S.t0[lane];
        mass = S.mass[lane]; radius = S.radius[lane]; rc = S.rclass[lane];
        ots = bl_sortable(S.obs_t[lane]); oarg = S.obs_arg[lane]; oidx = S.obs_idx[lane];
        for (int j = 0; j < n; j++) T[lane][j] = S.toi[(size_t)lane * S.pitch + j];
    }
    __syncwarp();

    const uint64_t ts_inf = bl_sortable(BL_INF);
    uint64_t end_excl = bl_sortable(P.end_time);
    end_excl = end_excl >= ts_inf ? ts_inf : end_excl + 1;
    const bool want_log = P.log.cap > 0;
    bl_result *const res = P.result;
    long long done = 0, n_bb = 0;
    double now = P.time;
    int kindmask = P.first_kind == BL_BALL_BALL ? KIND_BB : (P.first_kind == BL_BALL_OBSTACLE ? KIND_BO : KIND_BOTH);
    int final_stage = P.query_only ? 1 : 0;
    if (final_stage == 1) kindmask = KIND_BB;

    for (;;) {
        // ---- my key: minimum of my table row (first index on ties), or my next obstacle ----
        uint64_t kts = ~0ull, kcode = BL_CODE_END;
        if (own) {
            if (kindmask != KIND_BO) {
                kts = ts_inf; kcode = BL_CODE_NONE;
                for (int j = 0; j < n; j++) {
                    const double v = T[lane][j];
                    if (j != lane && v < BL_INF) {
                        const uint64_t s2 = bl_sortable(v), c2 = bl_pair_code(lane, j);
                        if (bl_key_less(s2, c2, kts, kcode)) { kts = s2; kcode = c2; }
                    }
                }
            }
            if (kindmask != KIND_BB) {
                const uint64_t oc = BL_CODE_OBS | (uint64_t)lane;
                if (kindmask == KIND_BO || bl_key_less(ots, oc, kts, kcode)) { kts = ots; kcode = oc; }
            }
        }
        const int w = warp_argmin_s(kts, kcode);
        const uint64_t ts = shfl_u(kts, w), cd = shfl_u(kcode, w);
        const double t = bl_unsortable(ts);

        // ---- stop rules (same order as the analysis of bl_evolve.cu) ----
        int stop_kind = -1, st_status = BL_OK;
        const bool max_reached = P.max_events >= 0 && done >= P.max_events;
        if (final_stage == 0 && max_reached) stop_kind = BL_STOP_EVENTS;
        else if (final_stage != 0) stop_kind = BL_STOP_TIME;
        else if (ts == BL_TS_NAN) { stop_kind = BL_STOP_ERROR; st_status = BL_ENAN; }
        else if (ts >= end_excl) stop_kind = BL_STOP_TIME;
        else if (want_log && done >= P.log.cap) stop_kind = BL_STOP_LOG;

        int a = -1, b = -1;
        if (cd != BL_CODE_NONE && cd != BL_CODE_END) {
            if (cd & BL_CODE_OBS) a = (int)(cd & 0x7fffffffull);
            else { a = bl_code_lo(cd); b = bl_code_hi(cd); }
        }
        double ax = 0, ay = 0, avx = 0, avy = 0, bx = 0, by = 0, bvx = 0, bvy = 0;
        double a_vx = 0, a_vy = 0, b_vx = 0, b_vy = 0, qarg = 0;
        int oi = -1, rca = 0, rcb = 0;
        if (stop_kind < 0) {
            // ---- resolve, redundantly in every lane (states come by shuffle) ----
            const int sa = a, sb = b >= 0 ? b : a;
            const double a0x = shfl_d(p0x, sa), a0y = shfl_d(p0y, sa), a_t0 = shfl_d(t0, sa), am = shfl_d(mass, sa);
            a_vx = shfl_d(vx, sa); a_vy = shfl_d(vy, sa);
            const double b0x = shfl_d(p0x, sb), b0y = shfl_d(p0y, sb), b_t0 = shfl_d(t0, sb), bm = shfl_d(mass, sb);
            b_vx = shfl_d(vx, sb); b_vy = shfl_d(vy, sb);
            qarg = shfl_d(oarg, sa);
            oi = __shfl_sync(0xffffffffu, oidx, sa);
            rca = __shfl_sync(0xffffffffu, rc, sa); rcb = __shfl_sync(0xffffffffu, rc, sb);
            ax = bl_advance(a0x, a_vx, t, a_t0); ay = bl_advance(a0y, a_vy, t, a_t0);
            int st;
            if (b < 0) {
                avx = a_vx; avy = a_vy;
                st = oi >= 0 ? bl_obstacle_bounce<FMA>(s_obs[oi], ax, ay, a_vx, a_vy, qarg, avx, avy) : BL_EASSERT;
            } else {
                bx = bl_advance(b0x, b_vx, t, b_t0); by = bl_advance(b0y, b_vy, t, b_t0);
                double m1 = am, m2 = bm;
                bl_remap_masses(m1, m2);
                avx = a_vx; avy = a_vy; bvx = b_vx; bvy = b_vy;
                st = bl_elastic<FMA>(ax, ay, a_vx, a_vy, m1, bx, by, b_vx, b_vy, m2, avx, avy, bvx, bvy);
            }
            if (st != BL_OK) {
                // the reference has moved to t and raised; the collision is not applied
                if (lane == 0) {
                    res->err_i = a; res->err_j = b; res->time = t; res->stop = BL_STOP_ERROR; res->status = st;
                    res->n_ball_ball = n_bb; res->n_ball_obstacle = done - n_bb; res->n_logged = want_log ? done : 0;
                }
                now = t;
                final_stage = 1; kindmask = KIND_BB;
                continue;
            }
        }
        if (stop_kind >= 0) {
            if (final_stage == 0) {
                if (lane == 0) {
                    double tnow = now;
                    res->err_i = -1; res->err_j = -1;
                    if (st_status != BL_OK) { res->err_i = a; res->err_j = b; tnow = t; }
                    if (stop_kind == BL_STOP_TIME && P.end_time < BL_INF) tnow = P.end_time;  // _move(end_time), :436
                    res->time = tnow; res->stop = stop_kind; res->status = st_status;
                    res->n_ball_ball = n_bb; res->n_ball_obstacle = done - n_bb;
                    res->n_logged = want_log ? done : 0;
                }
                final_stage = 1; kindmask = KIND_BB;
                continue;
            }
            if (final_stage == 1) {
                if (lane == 0) {
                    if (P.query_only) { res->time = P.time; res->err_i = -1; res->err_j = -1; }
                    const bool some = cd != BL_CODE_NONE && cd != BL_CODE_END && t < BL_INF;
                    res->bb_t = some ? t : BL_INF; res->bb_i = some ? bl_code_lo(cd) : -1; res->bb_j = some ? bl_code_hi(cd) : 0;
                }
                final_stage = 2; kindmask = KIND_BO;
                continue;
            }
            // numpy argmin over _obstacles_toi: first ball attaining the minimum, even when it is +inf
            const bool some = cd != BL_CODE_END;
            const int ball = some ? (int)(cd & 0x7fffffffull) : 0;
            const int q_oi = __shfl_sync(0xffffffffu, oidx, ball);
            const double q_arg = shfl_d(oarg, ball);
            if (lane == 0) {
                res->bo_t = some ? t : BL_INF; res->bo_ball = some ? ball : -1;
                res->bo_obs = some ? q_oi : -1; res->bo_arg = some ? q_arg : 0.0;
            }
            break;
        }

        // ---- log, simulation.py:590-594 / :626-631 payloads ----
        if (want_log && lane == 0) {
            const long long e = done;
            P.log.t[e] = t; P.log.i[e] = a; P.log.j[e] = b >= 0 ? (long long)b : -1 - (long long)oi;
            if (P.log.payload) {
                double *o = P.log.payload + 12 * e;
                o[0] = ax; o[1] = ay; o[2] = a_vx; o[3] = a_vy; o[4] = avx; o[5] = avy;
                if (b >= 0) { o[6] = bx; o[7] = by; o[8] = b_vx; o[9] = b_vy; o[10] = bvx; o[11] = bvy; }
                else { o[6] = qarg; o[7] = 0; o[8] = 0; o[9] = 0; o[10] = 0; o[11] = 0; }
            }
        }
        // ---- every other ball re-solves its pair with the colliding ball(s); simulation.py:461-471, :522-526 ----
        if (own && lane != a && lane != b) {
            const double px = bl_advance(p0x, vx, t, t0), py = bl_advance(p0y, vy, t, t0);
            const double rsa = S.rsum2[(size_t)rca * K + rc];
            const double va = bl_add(t, bl_toi_ball_ball<FMA>(ax, ay, avx, avy, px, py, vx, vy, rsa));
            T[lane][a] = va; T[a][lane] = va;
            if (b >= 0) {
                const double rsb = S.rsum2[(size_t)rcb * K + rc];
                const double vb = bl_add(t, bl_toi_ball_ball<FMA>(bx, by, bvx, bvy, px, py, vx, vy, rsb));
                T[lane][b] = vb; T[b][lane] = vb;
            }
        }
        // ---- the colliding balls: new state of record, pair entry = inf (:458), next obstacle (:490-495, :545-547) ----
        if (lane == a || lane == b) {
            if (lane == a) { p0x = ax; p0y = ay; vx = avx; vy = avy; }
            else { p0x = bx; p0y = by; vx = bvx; vy = bvy; }
            t0 = t;
            if (b >= 0) T[lane][lane == a ? b : a] = BL_INF;
            double ot, na;
            int ni;
            bl_next_obstacle<FMA>(s_obs, n_obs, S.rsum2_obs, K, rc, p0x, p0y, vx, vy, radius, t, ot, ni, na);
            ots = bl_sortable(ot); oarg = na; oidx = ni;
        }
        __syncwarp();
        done++;
        n_bb += b >= 0 ? 1 : 0;
        now = t;
        kindmask = KIND_BOTH;  // only the first event's kind was forced
    }

    // ---- write the state, the table and the caches back ----
    if (own) {
        S.p0x[lane] = p0x; S.p0y[lane] = p0y; S.vx[lane] = vx; S.vy[lane] = vy; S.t0[lane] = t0;
        S.obs_t[lane] = bl_unsortable(ots); S.obs_arg[lane] = oarg; S.obs_idx[lane] = oidx;
        uint64_t mts = ts_inf, mcode = BL_CODE_NONE;
        for (int j = 0; j < n; j++) {
            const double v = T[lane][j];
            S.toi[(size_t)lane * S.pitch + j] = v;
            if (j != lane && v < BL_INF) {
                const uint64_t s2 = bl_sortable(v), c2 = bl_pair_code(lane, j);
                if (bl_key_less(s2, c2, mts, mcode)) { mts = s2; mcode = c2; }
            }
        }
        S.cover_t[lane] = bl_unsortable(mts); S.cover_code[lane] = mcode;
    }
    if (lane == 0) {
        for (int q = 0; q < 16; q++) res->prof[q] = 0;
        res->prof[5] = done;
        res->window = 0.0;
        *S.seq_counter += done;
    }
}

}  // namespace

int bl_launch_evolve_small(bl_handle *h, const bl_system *s, double time, double end_time, long long max_events,
                           int first_kind, const bl_log *log, int query_only, cudaStream_t st) {
    SmallParams P;
    P.s = *s;
    P.time = time; P.end_time = end_time; P.max_events = max_events; P.first_kind = first_kind;
    P.query_only = query_only;
    if (log) P.log = *log; else { P.log.cap = 0; P.log.t = nullptr; P.log.i = nullptr; P.log.j = nullptr; P.log.payload = nullptr; }
    P.result = h->d_result;
    cudaError_t e = cudaMemsetAsync(h->d_result, 0, sizeof(bl_result), st);
    if (e != cudaSuccess) return bl_check_cuda(h, e, "memset result");
    h->launches++;
    if (h->dot_fma) bl_evolve_small_kernel<true><<<1, 32, 0, st>>>(P);
    else bl_evolve_small_kernel<false><<<1, 32, 0, st>>>(P);
    return bl_check_cuda(h, cudaGetLastError(), "launch bl_evolve_small_kernel");
}
