// bl_evolve_small.cu -- the event loop for systems of at most 32 balls, ONE GROUP OF LANES PER SYSTEM.
//
// Galperin's two balls, the pool table, Newton's cradle (the reference's own examples) are strictly serial --
// consecutive collisions share balls, a batch would hold one collision -- so the grid-wide round machinery of
// bl_evolve.cu buys nothing there.  Here a system of n balls is served by W = 2, 4, 8, 16 or 32 lanes of one warp
// (the smallest power of two >= n): lane i owns ball i (state of record in registers), the system's whole
// time-of-impact table sits in shared memory (both triangles current), and one collision is: row minima (compare
// tree) -> group argmin (xor butterfly over the W lanes) -> every lane resolves the collision redundantly from
// shuffled states -> every other lane re-solves its pair(s) with the colliding ball(s) -> the (ball, obstacle) solves
// of the colliding balls are spread over the lanes and combined by a second butterfly.  No barrier but __syncwarp on
// the group's lanes.
//
// The same kernel serves
//   * a single Billiard (bl_evolve / bl_next on a bl_system with n <= 32): SINGLE = true, one group, with the
//     event log, the forced first kind of bounce_ball_ball / bounce_ball_obstacle and the next-collision queries;
//   * ENSEMBLES of independent multi-ball billiards (bl_multi_run): SINGLE = false, 32 / W systems per warp, four
//     warps per CTA, systems sharded over CTAs (and over GPUs by the caller) -- north_star's "one system per warp or
//     CTA, ball state staged in shared memory".  Every system is exactly "a Billiard with these balls": same
//     arithmetic, same keys, same stop rules, so an ensemble of one reproduces the single path bit for bit.
//
// Reference lines restated: simulation.py:354-438 (evolve), :440-502 (bounce_ball_ball), :504-554
// (bounce_ball_obstacle), :556-560 (_move), :562-636 (resolve), :176-209 (add_ball's row, for table builds).
#include <cstring>

#include "bl_small.cuh"

using namespace bl_small;

namespace {

// lexicographic (time, code): numpy's first-index argmin (FP compare, so -0.0 == 0.0 like there) with the pair /
// obstacle codes of bl_arith.cuh breaking ties
__device__ __forceinline__ bool key_less(double ta, uint64_t ca, double tb, uint64_t cb) {
    return ta < tb || (ta == tb && ca < cb);
}

// (measured: the same reduction as three REDUX.MIN over the order-preserving integer image of the time and a packed code
// is slower -- pool shot 0.73 vs 0.71 ms, 2^18 pool tables 1.64 vs 1.83 G collisions/s)
template <int W>
__device__ __forceinline__ void group_min(unsigned gmask, double &t, uint64_t &c) {
#pragma unroll
    for (int o = 1; o < W; o <<= 1) {
        const double t2 = __shfl_xor_sync(gmask, t, o, W);
        const uint64_t c2 = __shfl_xor_sync(gmask, c, o, W);
        const bool take = key_less(t2, c2, t, c);
        t = take ? t2 : t;
        c = take ? c2 : c;
    }
}

// minimum of a table row of W entries (entries past the last ball hold +inf) and its first index: chunks of eight
// loads reduced by a compare tree -- a dependent chain of 3 + W/8 compares instead of W
template <int W>
__device__ __forceinline__ void row_min(const double *Trow, double &best, int &bj) {
    best = BL_INF; bj = -1;
    constexpr int C = W < 8 ? W : 8;
#pragma unroll
    for (int base = 0; base < W; base += C) {
        double v[C];
        int ix[C];
#pragma unroll
        for (int q = 0; q < C; q++) { v[q] = Trow[base + q]; ix[q] = base + q; }
#pragma unroll
        for (int step = 1; step < C; step <<= 1)
#pragma unroll
            for (int q = 0; q + step < C; q += 2 * step) {
                const bool lt = v[q + step] < v[q];  // strict: the lower index survives a tie (numpy's argmin)
                v[q] = lt ? v[q + step] : v[q];
                ix[q] = lt ? ix[q + step] : ix[q];
            }
        const bool lt = v[0] < best;
        best = lt ? v[0] : best;
        bj = lt ? ix[0] : bj;
    }
}

// Ensemble launches: six CTAs of four warps per SM (80 registers, a few spilled words outside the hot path) -- the loop
// is a chain of long-latency steps, and 24 resident warps hide it better than the 16 that the compiler's own 127
// registers allow: 1.79 -> 1.99 G collisions/s on 2^16 pool tables (seven and eight CTAs per SM measure the same).
template <bool FMA, int W, bool SINGLE>
__global__ void __launch_bounds__(SINGLE ? 32 : 32 * BL_SMALL_WARPS, SINGLE ? 1 : 6)
    bl_evolve_small_kernel(const __grid_constant__ SmallParams P) {
    // one table row per lane: row (warp, lane) = T[ball li of that lane's system][0..W), padded to W + 1 (odd pitch:
    // the 32 lanes of a warp hit distinct banks whatever W is)
    extern __shared__ double smem[];
    __shared__ bl_obstacle s_obs[BL_SMALL_MAX_OBS];
    constexpr int RP = W + 1;
    const int n = P.n, K = P.K, n_obs = P.n_obs;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int li = lane & (W - 1);                      // my ball
    const int gbase = lane & ~(W - 1);                  // first lane of my group
    // Every shuffle, vote and __syncwarp names the whole warp: the groups of a warp run the collision loop in lockstep
    // (a group that is done idles until the last one is), so the warp is converged wherever they are issued.  With a
    // per-group mask -- a run-time value that differs between the lanes of a warp -- the compiler wraps every one of
    // them in MATCH.ANY + REDUX.OR + VOTEU sequences (7 of each per iteration in the first version of this kernel, 6 %
    // of its stall samples on the MATCH alone).  Threads that have exited do not count as participants.
    constexpr unsigned gmask = 0xffffffffu;
    (void)gbase;
    double *const Tw = smem + (size_t)warp * 32 * RP;   // this warp's rows
    double *const Trow = Tw + (size_t)lane * RP;        // my row
    double *const Tg = Tw + (size_t)gbase * RP;         // row 0 of my group: row of ball a = Tg + a * RP
    for (int q = threadIdx.x; q < n_obs; q += blockDim.x) s_obs[q] = P.obstacles[q];
    __syncthreads();
    const long long sys = SINGLE ? (lane / W) : ((long long)blockIdx.x * BL_SMALL_WARPS + warp) * (32 / W) + (lane / W);
    if (sys >= P.n_sys) return;  // whole groups leave together: every later sync names the group's lanes only
    const bool own = li < n;
    const long long sb = sys * n;  // first ball of my system
    // my system's composition (shared by all systems unless the strides say otherwise)
    const double *const c_mass = P.mass + sys * P.ball_stride, *const c_radius = P.radius + sys * P.ball_stride;
    const int32_t *const c_rclass = P.rclass + sys * P.ball_stride;
    const double *const c_rsum2 = P.rsum2 + sys * P.rs_stride, *const c_rsum2_obs = P.rsum2_obs + sys * P.rso_stride;

    double p0x = 0, p0y = 0, vx = 0, vy = 0, t0 = 0, mass = 0, radius = 0, oarg = 0, ot = BL_INF;
    int rc = 0, oidx = -1;
    double now = SINGLE ? P.time : P.sys_time[sys];
    if (own) {
        p0x = P.p0x[sb + li]; p0y = P.p0y[sb + li]; vx = P.vx[sb + li]; vy = P.vy[sb + li]; t0 = P.t0[sb + li];
        mass = c_mass[li]; radius = c_radius[li]; rc = c_rclass[li];
    }
    const double *const rs_row = c_rsum2 + (size_t)rc * K;  // (r_me + r_x)**2 by the class of x
    const double rs_one = c_rsum2[0];
    if (!SINGLE && P.build_tables) {
        // add_ball, row by row at the current time (simulation.py:176-209): every pair solved from the positions of
        // `now`; the solve is bit-symmetric in the two balls, so lane i fills its own row
        const double px = bl_advance(p0x, vx, now, t0), py = bl_advance(p0y, vy, now, t0);
        for (int j = 0; j < W; j++) {
            const double jx = __shfl_sync(gmask, px, j, W), jy = __shfl_sync(gmask, py, j, W);
            const double jvx = __shfl_sync(gmask, vx, j, W), jvy = __shfl_sync(gmask, vy, j, W);
            const int jrc = __shfl_sync(gmask, rc, j, W);
            double v = BL_INF;
            if (own && j < n && j != li)
                v = bl_add(now, bl_toi_ball_ball<FMA>(px, py, vx, vy, jx, jy, jvx, jvy, K == 1 ? rs_one : rs_row[jrc]));
            Trow[j] = v;
        }
        if (own) {
            bl_next_obstacle<FMA>(s_obs, n_obs, c_rsum2_obs, K, rc, px, py, vx, vy, radius, now, ot, oidx, oarg);
        }
    } else {
        if (own) {
            ot = P.obs_t[sb + li]; oarg = P.obs_arg[sb + li]; oidx = P.obs_idx[sb + li];
            const double *src = P.toi + sys * P.toi_sys_stride + (size_t)li * P.toi_pitch;
            for (int j = 0; j < n; j++) Trow[j] = src[j];
            for (int j = n; j < W; j++) Trow[j] = BL_INF;  // row_min scans all W entries
            Trow[li] = BL_INF;
        } else {
            for (int j = 0; j < W; j++) Trow[j] = BL_INF;
        }
    }
    __syncwarp(gmask);

    bool want_log = SINGLE && P.log.cap > 0;
    bool k_one = K == 1;  // one radius class: (r_a + r_b)**2 is one number
    int n_obs_r = n_obs;
    {
        // loop invariants as opaque register values, not constant-bank reloads + compares inside the loop
        int wl = want_log ? 1 : 0, ko = k_one ? 1 : 0;
        asm volatile("" : "+r"(wl), "+r"(ko), "+r"(n_obs_r));
        want_log = wl != 0; k_one = ko != 0;
    }
    // the limits as opaque register values (re-reading kernel parameters from the constant bank inside the loop costs a
    // load, a compare and a branch each time): collisions this launch may handle, room in the log, last admissible time
    long long max_ev = P.max_events >= 0 ? P.max_events : 0x7fffffffffffffffLL;
    long long log_cap = want_log ? P.log.cap : 0x7fffffffffffffffLL;
    double end_t = P.end_time;
    asm volatile("" : "+l"(max_ev), "+l"(log_cap), "+d"(end_t));
    bl_result *const res = P.result;
    long long done = 0, n_bb = 0;
    int kindmask = KIND_BOTH, final_stage = 0;
    if (SINGLE) {
        kindmask = P.first_kind == BL_BALL_BALL ? KIND_BB : (P.first_kind == BL_BALL_OBSTACLE ? KIND_BO : KIND_BOTH);
        if (P.query_only) { final_stage = 1; kindmask = KIND_BB; }
    }
    int end_stop = BL_STOP_TIME, end_status = BL_OK;
    bool fin = false;  // ensembles: my system has stopped (its group keeps pace with the others of the warp)

    for (;;) {
        // ---- my key: minimum of my table row (first index on ties), or my next obstacle ----
        double kt = BL_INF;
        uint64_t kc = BL_CODE_END;
        if (own && !fin) {
            if (kindmask != KIND_BO) {
                int bj;
                row_min<W>(Trow, kt, bj);
                kc = bj >= 0 ? bl_pair_code(li, bj) : BL_CODE_NONE;
            }
            if (kindmask != KIND_BB) {
                // a ball-ball collision wins a tie (simulation.py:421), a NaN obstacle time never wins
                if (kindmask == KIND_BO || ot < kt) { kt = ot; kc = BL_CODE_OBS | (uint64_t)li; }
            }
        }
        group_min<W>(gmask, kt, kc);
        const double t = kt;
        const uint64_t cd = kc;

        // ---- stop rules ----
        int stop_kind = -1, st_status = BL_OK;
        if (SINGLE && final_stage != 0) stop_kind = BL_STOP_TIME;
        else if (done >= max_ev) stop_kind = BL_STOP_EVENTS;
        else if (t != t) { stop_kind = BL_STOP_ERROR; st_status = BL_ENAN; }
        else if (!(t <= end_t) || !(t < BL_INF) || cd == BL_CODE_NONE || cd == BL_CODE_END) stop_kind = BL_STOP_TIME;
        else if (done >= log_cap) stop_kind = BL_STOP_LOG;

        int a = -1, b = -1;
        if (cd != BL_CODE_NONE && cd != BL_CODE_END) {
            if (cd & BL_CODE_OBS) a = (int)(cd & 0x7fffffffull);
            else { a = bl_code_lo(cd); b = bl_code_hi(cd); }
        }
        double ax = 0, ay = 0, avx = 0, avy = 0, bx = 0, by = 0, bvx = 0, bvy = 0;
        double a_vx = 0, a_vy = 0, b_vx = 0, b_vy = 0, qarg = 0;
        int oi = -1, rca = 0, rcb = 0;
        // the states of the two balls come by shuffle (issued by every lane of the warp, whatever its group decides)
        const int sa = a >= 0 ? a : 0, sb2 = b >= 0 ? b : sa;
        const double a0x = __shfl_sync(gmask, p0x, sa, W), a0y = __shfl_sync(gmask, p0y, sa, W);
        const double a_t0 = __shfl_sync(gmask, t0, sa, W), am = __shfl_sync(gmask, mass, sa, W);
        a_vx = __shfl_sync(gmask, vx, sa, W); a_vy = __shfl_sync(gmask, vy, sa, W);
        const double b0x = __shfl_sync(gmask, p0x, sb2, W), b0y = __shfl_sync(gmask, p0y, sb2, W);
        const double b_t0 = __shfl_sync(gmask, t0, sb2, W), bm = __shfl_sync(gmask, mass, sb2, W);
        b_vx = __shfl_sync(gmask, vx, sb2, W); b_vy = __shfl_sync(gmask, vy, sb2, W);
        qarg = __shfl_sync(gmask, oarg, sa, W);
        oi = __shfl_sync(gmask, oidx, sa, W);
        rca = __shfl_sync(gmask, rc, sa, W); rcb = __shfl_sync(gmask, rc, sb2, W);
        if (stop_kind < 0) {
            // ---- resolve, redundantly in every lane of the group ----
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
                stop_kind = BL_STOP_ERROR; st_status = st;
                now = t;
            }
        }
        if (!SINGLE) {
            if (stop_kind >= 0 && !fin) { end_stop = stop_kind; end_status = st_status; fin = true; }
            if (__all_sync(gmask, fin)) break;
        }
        const bool act = stop_kind < 0;  // my group handles a collision in this iteration
        if (SINGLE && stop_kind >= 0) {
            if (final_stage == 0) {
                if (li == 0) {
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
                if (li == 0) {
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
            const int q_oi = __shfl_sync(gmask, oidx, ball, W);
            const double q_arg = __shfl_sync(gmask, oarg, ball, W);
            if (li == 0) {
                res->bo_t = some ? t : BL_INF; res->bo_ball = some ? ball : -1;
                res->bo_obs = some ? q_oi : -1; res->bo_arg = some ? q_arg : 0.0;
            }
            break;
        }

        // ---- log, simulation.py:590-594 / :626-631 payloads ----
        if (want_log && li == 0) {
            const long long e = done;
            P.log.t[e] = t; P.log.i[e] = a; P.log.j[e] = b >= 0 ? (long long)b : -1 - (long long)oi;
            if (P.log.payload) {
                double *o = P.log.payload + 12 * e;
                o[0] = ax; o[1] = ay; o[2] = a_vx; o[3] = a_vy; o[4] = avx; o[5] = avy;
                if (b >= 0) { o[6] = bx; o[7] = by; o[8] = b_vx; o[9] = b_vy; o[10] = bvx; o[11] = bvy; }
                else { o[6] = qarg; o[7] = 0; o[8] = 0; o[9] = 0; o[10] = 0; o[11] = 0; }
            }
        }
        if (act && own && li != a && li != b) {
            // ---- every other ball re-solves its pair with the colliding ball(s); simulation.py:461-471, :522-526 ----
            const double px = bl_advance(p0x, vx, t, t0), py = bl_advance(p0y, vy, t, t0);
            // both quadratics of physics.py:33-54 side by side (two independent chains), then the square root and the
            // division only for a pair that gets past the two early exits
            const double rsa = k_one ? rs_one : rs_row[rca];
            const double dax = bl_sub(px, ax), day = bl_sub(py, ay), dvax = bl_sub(vx, avx), dvay = bl_sub(vy, avy);
            const double qb_a = bl_dot2<FMA>(dax, day, dvax, dvay);
            const double qc_a = bl_sub(bl_dot2<FMA>(dax, day, dax, day), rsa);
            const double qd_a = bl_sub(bl_mul(qb_a, qb_a), bl_mul(bl_dot2<FMA>(dvax, dvay, dvax, dvay), qc_a));
            double va = BL_INF, vb = BL_INF;
            double qb_b = 0.0, qc_b = 0.0, qd_b = 0.0;
            if (b >= 0) {
                const double rsb = k_one ? rs_one : rs_row[rcb];
                const double dbx = bl_sub(px, bx), dby = bl_sub(py, by), dvbx = bl_sub(vx, bvx), dvby = bl_sub(vy, bvy);
                qb_b = bl_dot2<FMA>(dbx, dby, dvbx, dvby);
                qc_b = bl_sub(bl_dot2<FMA>(dbx, dby, dbx, dby), rsb);
                qd_b = bl_sub(bl_mul(qb_b, qb_b), bl_mul(bl_dot2<FMA>(dvbx, dvby, dvbx, dvby), qc_b));
            }
            if (!(qb_a >= 0.0) && !(qd_a <= 0.0)) {
                const double t1 = bl_div(qc_a, bl_add(-qb_a, bl_sqrt(qd_a)));
                va = t1 >= BL_T_EPS ? t1 : BL_INF;
            }
            va = bl_add(t, va);
            Trow[a] = va; Tg[a * RP + li] = va;
            if (b >= 0) {
                if (!(qb_b >= 0.0) && !(qd_b <= 0.0)) {
                    const double t1 = bl_div(qc_b, bl_add(-qb_b, bl_sqrt(qd_b)));
                    vb = t1 >= BL_T_EPS ? t1 : BL_INF;
                }
                vb = bl_add(t, vb);
                Trow[b] = vb; Tg[b * RP + li] = vb;
            }
        } else if (act && own) {
            // ---- the colliding balls: new state of record, pair entry = inf (:458) ----
            if (li == a) { p0x = ax; p0y = ay; vx = avx; vy = avy; }
            else { p0x = bx; p0y = by; vx = bvx; vy = bvy; }
            t0 = t;
            if (b >= 0) Trow[li == a ? b : a] = BL_INF;
        }
        // ---- next obstacle of the colliding balls (:490-495, :545-547), spread over the lanes of the group: lane l
        // takes the obstacles (l >> 1), (l >> 1) + W/2, ... for ball a (even lanes) or b (odd lanes); the partial minima
        // are combined by a butterfly over the lanes of equal parity -- smaller time, then smaller list index, which is
        // the reference's strict < in list order ----
        {
            const int which = li & 1;
            const int tball = act ? (which ? b : a) : -1;
            const double tx = which ? bx : ax, ty = which ? by : ay, tvx = which ? bvx : avx, tvy = which ? bvy : avy;
            const int src = tball >= 0 ? tball : 0;
            const double tradius = __shfl_sync(gmask, radius, src, W);
            const int trc = which ? rcb : rca;
            double omin = BL_INF, oam = 0.0;
            int oq = -1;
            if (tball >= 0) {
                for (int q = li >> 1; q < n_obs_r; q += (W > 2 ? W / 2 : 1)) {
                    double arg;
                    const double rs = s_obs[q].kind != BL_OBS_WALL ? c_rsum2_obs[(size_t)q * K + trc] : 0.0;
                    const double tq = bl_obstacle_toi<FMA>(s_obs[q], tx, ty, tvx, tvy, tradius, rs, arg);
                    if (tq < omin) { omin = tq; oq = q; oam = arg; }
                }
            }
#pragma unroll
            for (int o = 2; o < W; o <<= 1) {
                const double t2 = __shfl_xor_sync(gmask, omin, o, W), a2 = __shfl_xor_sync(gmask, oam, o, W);
                const int q2 = __shfl_xor_sync(gmask, oq, o, W);
                const bool take = t2 < omin || (t2 == omin && q2 >= 0 && (oq < 0 || q2 < oq));
                omin = take ? t2 : omin; oam = take ? a2 : oam; oq = take ? q2 : oq;
            }
            const double t_a = __shfl_sync(gmask, omin, 0, W), g_a = __shfl_sync(gmask, oam, 0, W);
            const int q_a = __shfl_sync(gmask, oq, 0, W);
            const double t_b = __shfl_sync(gmask, omin, 1, W), g_b = __shfl_sync(gmask, oam, 1, W);
            const int q_b = __shfl_sync(gmask, oq, 1, W);
            if (act && li == a) { ot = bl_add(t, t_a); oidx = q_a; oarg = g_a; }
            if (act && li == b) { ot = bl_add(t, t_b); oidx = q_b; oarg = g_b; }
        }
        __syncwarp(gmask);
        if (act) {
            done++;
            n_bb += b >= 0 ? 1 : 0;
            now = t;
        }
        kindmask = KIND_BOTH;  // only the first event's kind was forced
    }

    // ---- write the state, the table and the caches back ----
    if (own) {
        P.p0x[sb + li] = p0x; P.p0y[sb + li] = p0y; P.vx[sb + li] = vx; P.vy[sb + li] = vy; P.t0[sb + li] = t0;
        P.obs_t[sb + li] = ot; P.obs_arg[sb + li] = oarg; P.obs_idx[sb + li] = oidx;
        double *dst = P.toi + sys * P.toi_sys_stride + (size_t)li * P.toi_pitch;
        double mt = BL_INF;
        int mj = -1;
        for (int j = 0; j < n; j++) {
            const double v = Trow[j];
            dst[j] = v;
            const bool lt = v < mt;
            mt = lt ? v : mt;
            mj = lt ? j : mj;
        }
        if (SINGLE) { P.cover_t[li] = mt; P.cover_code[li] = mj >= 0 ? bl_pair_code(li, mj) : BL_CODE_NONE; }
    }
    if (li == 0) {
        if (SINGLE) {
            for (int q = 0; q < 16; q++) res->prof[q] = 0;
            res->prof[5] = done;
            res->window = 0.0;
            *P.seq_counter += done;
        } else {
            // _move(end_time) when the limit was a time (simulation.py:436); otherwise the clock of the last collision
            P.sys_time[sys] = (end_stop == BL_STOP_TIME && P.end_time < BL_INF) ? P.end_time : now;
            P.n_bb[sys] += n_bb; P.n_bo[sys] += done - n_bb;
            P.status[sys] = end_status;
        }
    }
}

// positions of every ball of every system at `time` (or at its own system's clock): _move as a projection
__global__ void __launch_bounds__(256) bl_multi_positions_kernel(const long long total, const int n,
                                                                  const double *__restrict__ p0x,
                                                                  const double *__restrict__ p0y,
                                                                  const double *__restrict__ vx,
                                                                  const double *__restrict__ vy,
                                                                  const double *__restrict__ t0,
                                                                  const double *__restrict__ sys_time, const double time,
                                                                  const int at_system_time, double *__restrict__ pos) {
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const double t = at_system_time ? sys_time[k / n] : time;
        pos[2 * k] = bl_advance(p0x[k], vx[k], t, t0[k]);
        pos[2 * k + 1] = bl_advance(p0y[k], vy[k], t, t0[k]);
    }
}

template <bool FMA, bool SINGLE>
cudaError_t launch_w(const SmallParams &P, int sm_count, cudaStream_t st) {
    int W = 2;
    while (W < P.n) W <<= 1;
#define BL_SMALL_LAUNCH(WW)                                                                                            \
    if (W == WW) {                                                                                                     \
        const int warps = SINGLE ? 1 : BL_SMALL_WARPS;                                                                 \
        const size_t smem = (size_t)warps * 32 * (WW + 1) * sizeof(double);                                            \
        const long long per_cta = (long long)warps * (32 / WW);                                                        \
        const unsigned grid = SINGLE ? 1u : (unsigned)((P.n_sys + per_cta - 1) / per_cta);                             \
        bl_evolve_small_kernel<FMA, WW, SINGLE><<<grid, 32 * warps, smem, st>>>(P);                                    \
        return cudaGetLastError();                                                                                     \
    }
    BL_SMALL_LAUNCH(2) BL_SMALL_LAUNCH(4) BL_SMALL_LAUNCH(8) BL_SMALL_LAUNCH(16) BL_SMALL_LAUNCH(32)
#undef BL_SMALL_LAUNCH
    (void)sm_count;
    return cudaErrorInvalidValue;
}

}  // namespace

int bl_launch_evolve_small(bl_handle *h, const bl_system *s, double time, double end_time, long long max_events,
                           int first_kind, const bl_log *log, int query_only, cudaStream_t st) {
    SmallParams P;
    memset(&P, 0, sizeof(P));
    P.n_sys = 1; P.n = s->n; P.K = s->n_classes; P.n_obs = s->n_obstacles;
    P.toi_pitch = s->pitch; P.toi_sys_stride = 0;
    P.p0x = s->p0x; P.p0y = s->p0y; P.vx = s->vx; P.vy = s->vy; P.t0 = s->t0;
    P.radius = s->radius; P.mass = s->mass; P.rclass = s->rclass;
    P.rsum2 = s->rsum2; P.rsum2_obs = s->rsum2_obs; P.obstacles = s->obstacles;
    P.toi = s->toi; P.obs_t = s->obs_t; P.obs_arg = s->obs_arg; P.obs_idx = s->obs_idx;
    P.end_time = end_time; P.max_events = max_events;
    P.time = time; P.first_kind = first_kind; P.query_only = query_only;
    if (log) P.log = *log;
    P.result = h->d_result;
    P.cover_t = s->cover_t; P.cover_code = s->cover_code; P.seq_counter = s->seq_counter;
    cudaError_t e = cudaMemsetAsync(h->d_result, 0, sizeof(bl_result), st);
    if (e != cudaSuccess) return bl_check_cuda(h, e, "memset result");
    h->launches++;
    if (P.n > BL_SMALL_MAX_BALLS)
        return bl_check_cuda(h, bl_launch_cta(P, h->dot_fma != 0, true, st), "launch bl_evolve_cta_kernel");
    if (P.n <= BL_TINY_MAX_BALLS && !h->small_force_lanes)
        return bl_check_cuda(h, bl_launch_tiny(P, h->dot_fma != 0, true, h->sm_count, st), "launch bl_evolve_tiny_kernel");
    e = h->dot_fma ? launch_w<true, true>(P, h->sm_count, st) : launch_w<false, true>(P, h->sm_count, st);
    return bl_check_cuda(h, e, "launch bl_evolve_small_kernel");
}

int bl_launch_multi(bl_handle *h, const bl_multi *m, int build_tables, double end_time, long long max_events,
                    cudaStream_t st) {
    if (m->n_sys <= 0) return BL_OK;
    SmallParams P;
    memset(&P, 0, sizeof(P));
    P.n_sys = m->n_sys; P.n = m->n; P.K = m->n_classes; P.n_obs = m->n_obstacles;
    P.toi_pitch = m->n; P.toi_sys_stride = (long long)m->n * m->n;
    P.p0x = m->p0x; P.p0y = m->p0y; P.vx = m->vx; P.vy = m->vy; P.t0 = m->t0;
    P.radius = m->radius; P.mass = m->mass; P.rclass = m->rclass;
    P.rsum2 = m->rsum2; P.rsum2_obs = m->rsum2_obs; P.obstacles = m->obstacles;
    P.toi = m->toi; P.obs_t = m->obs_t; P.obs_arg = m->obs_arg; P.obs_idx = m->obs_idx;
    P.end_time = end_time; P.max_events = max_events;
    P.build_tables = build_tables;
    P.sys_time = m->time; P.n_bb = m->n_ball_ball; P.n_bo = m->n_ball_obstacle; P.status = m->status;
    if (m->per_system) {
        P.ball_stride = m->n; P.rs_stride = (long long)m->n_classes * m->n_classes;
        P.rso_stride = (long long)m->n_obstacles * m->n_classes;
    }
    h->launches++;
    if (P.n > BL_SMALL_MAX_BALLS)
        return bl_check_cuda(h, bl_launch_cta(P, h->dot_fma != 0, false, st), "launch bl_evolve_cta_kernel (ensemble)");
    if (P.n <= BL_TINY_MAX_BALLS && !h->small_force_lanes)
        return bl_check_cuda(h, bl_launch_tiny(P, h->dot_fma != 0, false, h->sm_count, st),
                             "launch bl_evolve_tiny_kernel (ensemble)");
    cudaError_t e = h->dot_fma ? launch_w<true, false>(P, h->sm_count, st) : launch_w<false, false>(P, h->sm_count, st);
    return bl_check_cuda(h, e, "launch bl_evolve_small_kernel (ensemble)");
}

int bl_launch_multi_positions(bl_handle *h, const bl_multi *m, double time, int at_system_time, double *d_pos,
                              cudaStream_t st) {
    const long long total = m->n_sys * m->n;
    if (total <= 0) return BL_OK;
    const long long want = (total + 255) / 256;
    const unsigned grid = (unsigned)(want < 8LL * h->sm_count ? want : 8LL * h->sm_count);
    bl_multi_positions_kernel<<<grid, 256, 0, st>>>(total, m->n, m->p0x, m->p0y, m->vx, m->vy, m->t0, m->time, time,
                                                    at_system_time, d_pos);
    h->launches++;
    return bl_check_cuda(h, cudaGetLastError(), "launch bl_multi_positions_kernel");
}
