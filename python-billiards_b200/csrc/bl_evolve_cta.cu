// bl_evolve_cta.cu -- the event loop for systems of 33 ... 224 balls, ONE CTA PER SYSTEM, one collision per iteration.
//
// The reference's own example sizes (collapse and maze: 200 balls) sit between the lane-group kernel (<= 32 balls, one
// warp) and the sizes where the grid-wide batches of bl_evolve.cu pay off: at 100 ... 250 balls a synchronisation
// round of that kernel holds 4 ... 6 collisions and costs the same ~15 k cycles as one of 30.  Here thread i owns ball
// i of the system (state of record in registers, mirrored in shared memory for the others to read), the system's
// time-of-impact table sits in shared memory as ONE TRIANGLE (entry (i, j), i > j, at i (i - 1) / 2 + j: 224 balls =
// 200 KB), every thread caches the minimum of its table row, and one collision is:
//   block argmin of the per-ball keys (butterfly per warp, partial minima through shared memory)
//   -> every thread resolves the collision redundantly from the mirrored states
//   -> every other thread re-solves its pair(s) with the colliding ball(s) and updates its cached row minimum; a thread
//      whose minimum was a colliding ball, and the colliding balls themselves, are put on a list
//   -> one warp solves the (ball, obstacle) pairs of the colliding balls spread over its lanes (ensembles: warp 0; a
//      single system: an extra warp without balls, beside the pair solves of the others)
//   -> the rows on the list are scanned again, one warp per row (the colliding balls' own cached minima simply start
//      empty: all their pairs compete in the other balls' minima -- the "cover" of bl_evolve.cu).
// Two __syncthreads per collision, a third when a row is listed (about every second collision).  Same arithmetic,
// keys, stop rules, log and result block as the lane-group kernel (bl_evolve_small.cu), which it follows line by line
// where the two coincide; the same two front ends: a single Billiard (SINGLE) and an ensemble of independent systems,
// one CTA each (bl_multi_run).
//
// Reference lines restated: simulation.py:354-438 (evolve), :440-502 (bounce_ball_ball), :504-554
// (bounce_ball_obstacle), :556-560 (_move), :562-636 (resolve), :176-209 (add_ball's row, for table builds).
#include <cstring>

#include "bl_small.cuh"

using namespace bl_small;

namespace {

__device__ __forceinline__ bool key_less(double ta, uint64_t ca, double tb, uint64_t cb) {
    return ta < tb || (ta == tb && ca < cb);
}

// entry of the pair (i, j), i != j, in the triangle
__device__ __forceinline__ int tri_at(int i, int j) {
    const int hi = i > j ? i : j, lo = i > j ? j : i;
    return ((hi * (hi - 1)) >> 1) + lo;
}

// shared-memory map, in 8-byte words after the triangle
struct CtaLayout {
    int tri, p0x, p0y, vx, vy, t0, oarg, ct, wt, wc, obr, oidx, cj, list, misc, words;
};
__host__ __device__ inline CtaLayout cta_layout(int n) {
    CtaLayout L;
    int w = 0;
    L.tri = w; w += (n * (n - 1)) / 2 + 1;
    L.p0x = w; w += n; L.p0y = w; w += n; L.vx = w; w += n; L.vy = w; w += n; L.t0 = w; w += n;
    L.oarg = w; w += n;
    L.ct = w; w += n;            // re-scanned row minima
    L.wt = w; w += 2 * 8;        // partial minima of the warps: time, two buffers
    L.wc = w; w += 2 * 8;        // ... and code
    L.obr = w; w += 4;           // next-obstacle time and argument of the two colliding balls
    L.oidx = w; w += (n + 1) / 2;   // ints from here
    L.cj = w; w += (n + 1) / 2;
    L.list = w; w += (n + 1) / 2;
    L.misc = w; w += 2;          // ints: two list lengths, obstacle indices of the two colliding balls
    L.words = w;
    return L;
}

// MAXW: warps per CTA the build is for.  Ensembles of the smaller systems are bound by how many CTAs an SM holds (each
// one a chain of ~1000-cycle steps), so the 2- and 4-warp builds trade registers for residency.
template <bool FMA, bool SINGLE, int MAXW>
__global__ void __launch_bounds__(32 * (MAXW + (SINGLE ? 1 : 0)), (SINGLE || MAXW > 4) ? 1 : (MAXW == 2 ? 10 : 5))
    bl_evolve_cta_kernel(const __grid_constant__ SmallParams P) {
    extern __shared__ double smem[];
    __shared__ bl_obstacle s_obs[BL_SMALL_MAX_OBS];
    const int n = P.n, K = P.K, n_obs = P.n_obs;
    const int li = threadIdx.x, lane = li & 31, warp = li >> 5, nw = blockDim.x >> 5;
    const CtaLayout L = cta_layout(n);
    double *const tri = smem + L.tri;
    double *const s_p0x = smem + L.p0x, *const s_p0y = smem + L.p0y, *const s_vx = smem + L.vx, *const s_vy = smem + L.vy;
    double *const s_t0 = smem + L.t0, *const s_oarg = smem + L.oarg, *const s_ct = smem + L.ct;
    double *const s_wt = smem + L.wt, *const s_obr = smem + L.obr;
    uint64_t *const s_wc = reinterpret_cast<uint64_t *>(smem + L.wc);
    int *const s_oidx = reinterpret_cast<int *>(smem + L.oidx), *const s_cj = reinterpret_cast<int *>(smem + L.cj);
    int *const s_list = reinterpret_cast<int *>(smem + L.list), *const s_misc = reinterpret_cast<int *>(smem + L.misc);
    const unsigned FULL = 0xffffffffu;
    // the warp that solves the (ball, obstacle) pairs of the colliding balls: for a single system an extra warp without
    // balls, so that these solves run beside the pair solves of the others instead of after warp 0's (1.2 k of ~5.5 k
    // cycles per collision); in ensembles warp 0 (the SM is kept busy by other CTAs, an extra warp would only redo the resolve)
    const int ow = SINGLE ? nw - 1 : 0;

    for (int q = li; q < n_obs; q += blockDim.x) s_obs[q] = P.obstacles[q];
    const long long sys = SINGLE ? 0 : (long long)blockIdx.x;
    const bool own = li < n;
    const long long sb = sys * n;
    // my system's composition (shared by all systems unless the strides say otherwise)
    const double *const c_mass = P.mass + sys * P.ball_stride, *const c_radius = P.radius + sys * P.ball_stride;
    const int32_t *const c_rclass = P.rclass + sys * P.ball_stride;
    const double *const c_rsum2 = P.rsum2 + sys * P.rs_stride, *const c_rsum2_obs = P.rsum2_obs + sys * P.rso_stride;

    double p0x = 0, p0y = 0, vx = 0, vy = 0, t0 = 0, radius = 0, oarg = 0, ot = BL_INF;
    int rc = 0, oidx = -1;
    double now = SINGLE ? P.time : P.sys_time[sys];
    if (own) {
        p0x = P.p0x[sb + li]; p0y = P.p0y[sb + li]; vx = P.vx[sb + li]; vy = P.vy[sb + li]; t0 = P.t0[sb + li];
        radius = c_radius[li]; rc = c_rclass[li];
        s_p0x[li] = p0x; s_p0y[li] = p0y; s_vx[li] = vx; s_vy[li] = vy; s_t0[li] = t0;
    }
    if (li == 0) { s_misc[0] = 0; s_misc[1] = 0; s_misc[2] = -1; s_misc[3] = -1; }
    const double *const rs_row = c_rsum2 + (size_t)rc * K;  // (r_me + r_x)**2 by the class of x
    const double rs_one = c_rsum2[0];
    __syncthreads();
    if (!SINGLE && P.build_tables) {
        // add_ball, row by row at the current time (simulation.py:176-209): every pair solved from the positions of
        // `now`; thread i fills the entries (i, j < i)
        if (own) {
            const double px = bl_advance(p0x, vx, now, t0), py = bl_advance(p0y, vy, now, t0);
            for (int j = 0; j < li; j++) {
                const double jx = bl_advance(s_p0x[j], s_vx[j], now, s_t0[j]), jy = bl_advance(s_p0y[j], s_vy[j], now, s_t0[j]);
                const double rs = K == 1 ? rs_one : rs_row[c_rclass[j]];
                tri[tri_at(li, j)] = bl_add(now, bl_toi_ball_ball<FMA>(px, py, vx, vy, jx, jy, s_vx[j], s_vy[j], rs));
            }
            bl_next_obstacle<FMA>(s_obs, n_obs, c_rsum2_obs, K, rc, px, py, vx, vy, radius, now, ot, oidx, oarg);
        }
    } else {
        if (own) { ot = P.obs_t[sb + li]; oarg = P.obs_arg[sb + li]; oidx = P.obs_idx[sb + li]; }
        // rows of the table in global memory (both triangles current there), one warp per row, lanes along the row
        const double *const src = P.toi + sys * P.toi_sys_stride;
        for (int i = warp + 1; i < n; i += nw)
            for (int j = lane; j < i; j += 32) tri[((i * (i - 1)) >> 1) + j] = src[(size_t)i * P.toi_pitch + j];
    }
    if (own) { s_oarg[li] = oarg; s_oidx[li] = oidx; }
    __syncthreads();

    // minimum of row k of the table and its first index; `step` lanes share the row (x = first, first + step, ...)
    auto row_scan = [&](const int k, const int first, const int step, double &best, int &bj) {
        best = BL_INF; bj = -1;
        for (int x = first; x < n; x += step) {
            if (x == k) continue;
            const double v = tri[tri_at(k, x)];
            const bool lt = v < best;  // strict: the lower index survives a tie (numpy's argmin)
            best = lt ? v : best;
            bj = lt ? x : bj;
        }
    };
    double ct = BL_INF;    // minimum of my row (of its cover, see below) ...
    int cj = -1;           // ... and the first ball it is attained with (-1: nothing finite)
    bool partial = false;  // the cached minimum covers only a part of my row
    if (own) row_scan(li, 0, 1, ct, cj);

    bool want_log = SINGLE && P.log.cap > 0;
    long long max_ev = P.max_events >= 0 ? P.max_events : 0x7fffffffffffffffLL;
    long long log_cap = want_log ? P.log.cap : 0x7fffffffffffffffLL;
    double end_t = P.end_time;
    asm volatile("" : "+l"(max_ev), "+l"(log_cap), "+d"(end_t));
    bl_result *const res = P.result;
    long long done = 0, n_bb = 0, n_rows = 0;
    int kindmask = KIND_BOTH, final_stage = 0;
    if (SINGLE) {
        kindmask = P.first_kind == BL_BALL_BALL ? KIND_BB : (P.first_kind == BL_BALL_OBSTACLE ? KIND_BO : KIND_BOTH);
        if (P.query_only) { final_stage = 1; kindmask = KIND_BB; }
    }
    int end_stop = BL_STOP_TIME, end_status = BL_OK;
    int par = 0, lp = 0;
    // single systems: phase clocks (argmin, resolve, pair solves, obstacle solves, barrier, re-scan), reported by
    // thread 0 in result.prof[8..13] (tools/cta_once.py); compiled out of the ensemble builds
    long long pc[6] = {0, 0, 0, 0, 0, 0};
    long long c0 = SINGLE ? clock64() : 0;

    for (;;) {
        // ---- my key: the cached minimum of my table row (first index on ties), or my next obstacle ----
        double kt = BL_INF;
        uint64_t kc = BL_CODE_END;
        if (own) {
            if (kindmask != KIND_BO) { kt = ct; kc = cj >= 0 ? bl_pair_code(li, cj) : BL_CODE_NONE; }
            if (kindmask != KIND_BB) {
                // a ball-ball collision wins a tie (simulation.py:421), a NaN obstacle time never wins
                if (kindmask == KIND_BO || ot < kt) { kt = ot; kc = BL_CODE_OBS | (uint64_t)li; }
            }
        }
        // ---- block argmin: butterfly per warp, the partial minima through shared memory ----
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t2 = __shfl_xor_sync(FULL, kt, o);
            const uint64_t c2 = __shfl_xor_sync(FULL, kc, o);
            const bool take = key_less(t2, c2, kt, kc);
            kt = take ? t2 : kt;
            kc = take ? c2 : kc;
        }
        if (lane == 0) { s_wt[par * 8 + warp] = kt; s_wc[par * 8 + warp] = kc; }
        __syncthreads();
        double t = s_wt[par * 8];
        uint64_t cd = s_wc[par * 8];
        for (int w = 1; w < nw; w++) {
            const double t2 = s_wt[par * 8 + w];
            const uint64_t c2 = s_wc[par * 8 + w];
            const bool take = key_less(t2, c2, t, cd);
            t = take ? t2 : t;
            cd = take ? c2 : cd;
        }
        par ^= 1;
        if (SINGLE) { const long long c1 = clock64(); pc[0] += c1 - c0; c0 = c1; }

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
        if (stop_kind < 0) {
            // ---- resolve, redundantly in every thread, from the mirrored states ----
            const double a0x = s_p0x[a], a0y = s_p0y[a], a_t0 = s_t0[a];
            a_vx = s_vx[a]; a_vy = s_vy[a];
            rca = c_rclass[a];
            ax = bl_advance(a0x, a_vx, t, a_t0); ay = bl_advance(a0y, a_vy, t, a_t0);
            int st;
            if (b < 0) {
                qarg = s_oarg[a]; oi = s_oidx[a];
                avx = a_vx; avy = a_vy;
                st = oi >= 0 ? bl_obstacle_bounce<FMA>(s_obs[oi], ax, ay, a_vx, a_vy, qarg, avx, avy) : BL_EASSERT;
            } else {
                const double b0x = s_p0x[b], b0y = s_p0y[b], b_t0 = s_t0[b];
                b_vx = s_vx[b]; b_vy = s_vy[b];
                rcb = c_rclass[b];
                bx = bl_advance(b0x, b_vx, t, b_t0); by = bl_advance(b0y, b_vy, t, b_t0);
                double m1 = c_mass[a], m2 = c_mass[b];
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
        if (!SINGLE && stop_kind >= 0) { end_stop = stop_kind; end_status = st_status; break; }
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
            if (li == 0) {
                const bool some = cd != BL_CODE_END;
                const int ball = some ? (int)(cd & 0x7fffffffull) : 0;
                res->bo_t = some ? t : BL_INF; res->bo_ball = some ? ball : -1;
                res->bo_obs = some ? s_oidx[ball] : -1; res->bo_arg = some ? s_oarg[ball] : 0.0;
            }
            break;
        }

        if (SINGLE) { const long long c1 = clock64(); pc[1] += c1 - c0; c0 = c1; }
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
        bool rescan = false;
        if (own && li != a && li != b) {
            // ---- every other ball re-solves its pair with the colliding ball(s); simulation.py:461-471, :522-526 ----
            const double px = bl_advance(p0x, vx, t, t0), py = bl_advance(p0y, vy, t, t0);
            // both quadratics of physics.py:33-54 side by side (two independent chains), then the square root and the
            // division only for a pair that gets past the two early exits
            const double rsa = K == 1 ? rs_one : rs_row[rca];
            const double dax = bl_sub(px, ax), day = bl_sub(py, ay), dvax = bl_sub(vx, avx), dvay = bl_sub(vy, avy);
            const double qb_a = bl_dot2<FMA>(dax, day, dvax, dvay);
            const double qc_a = bl_sub(bl_dot2<FMA>(dax, day, dax, day), rsa);
            const double qd_a = bl_sub(bl_mul(qb_a, qb_a), bl_mul(bl_dot2<FMA>(dvax, dvay, dvax, dvay), qc_a));
            double va = BL_INF, vb = BL_INF;
            double qb_b = 0.0, qc_b = 0.0, qd_b = 0.0;
            if (b >= 0) {
                const double rsb = K == 1 ? rs_one : rs_row[rcb];
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
            tri[tri_at(li, a)] = va;
            if (b >= 0) {
                if (!(qb_b >= 0.0) && !(qd_b <= 0.0)) {
                    const double t1 = bl_div(qc_b, bl_add(-qb_b, bl_sqrt(qd_b)));
                    vb = t1 >= BL_T_EPS ? t1 : BL_INF;
                }
                vb = bl_add(t, vb);
                tri[tri_at(li, b)] = vb;
            }
            // my cached row minimum: a colliding ball was it -> scan the row again; otherwise the new values compete
            // (smaller value, or the same value at a lower index -- the first-index rule of the full scan)
            if (cj == a || (b >= 0 && cj == b)) { rescan = true; partial = false; }
            else {
                if (va < ct || (va == ct && a < cj)) { ct = va; cj = a; }
                if (b >= 0 && (vb < ct || (vb == ct && b < cj))) { ct = vb; cj = b; }
            }
        } else if (own) {
            // ---- the colliding balls: new state of record, pair entry = inf (:458) ----
            if (li == a) { p0x = ax; p0y = ay; vx = avx; vy = avy; }
            else { p0x = bx; p0y = by; vx = bvx; vy = bvy; }
            t0 = t;  // (the mirror in shared memory follows after the barrier: slower warps are still resolving from it)
            if (b >= 0 && li == a) tri[tri_at(a, b)] = BL_INF;
            // Every pair of mine has just been re-solved by its other ball and competes in THAT ball's cached minimum,
            // so mine starts empty instead of being scanned again: the cached minimum of a row is the minimum over a
            // subset of the row that contains every pair last written at a collision of the OTHER ball (the "cover" of
            // bl_evolve.cu) -- every pair is in the cover of one of its balls, every cached value is a current table
            // entry, so the block minimum of the cached keys is the minimum of the table, ties included.
            ct = BL_INF; cj = -1; partial = true;
        }
        if (rescan) s_list[atomicAdd(&s_misc[lp], 1)] = li;
        if (SINGLE) { const long long c1 = clock64(); pc[2] += c1 - c0; c0 = c1; }
        // ---- next obstacle of the colliding balls (:490-495, :545-547), spread over the lanes of warp 0: lane l takes
        // the obstacles (l >> 1), (l >> 1) + 16, ... for ball a (even lanes) or b (odd lanes); the partial minima are
        // combined by a butterfly over the lanes of equal parity -- smaller time, then smaller list index, which is the
        // reference's strict < in list order ----
        if (warp == ow) {
            const int which = lane & 1;
            const int tball = which ? b : a;
            const double tx = which ? bx : ax, ty = which ? by : ay, tvx = which ? bvx : avx, tvy = which ? bvy : avy;
            const int trc = which ? rcb : rca;
            double omin = BL_INF, oam = 0.0;
            int oq = -1;
            if (tball >= 0) {
                const double tradius = c_radius[tball];
                for (int q = lane >> 1; q < n_obs; q += 16) {
                    double arg;
                    const double rs = s_obs[q].kind != BL_OBS_WALL ? c_rsum2_obs[(size_t)q * K + trc] : 0.0;
                    const double tq = bl_obstacle_toi<FMA>(s_obs[q], tx, ty, tvx, tvy, tradius, rs, arg);
                    if (tq < omin) { omin = tq; oq = q; oam = arg; }
                }
            }
#pragma unroll
            for (int o = 2; o < 32; o <<= 1) {
                const double t2 = __shfl_xor_sync(FULL, omin, o), a2 = __shfl_xor_sync(FULL, oam, o);
                const int q2 = __shfl_xor_sync(FULL, oq, o);
                const bool take = t2 < omin || (t2 == omin && q2 >= 0 && (oq < 0 || q2 < oq));
                omin = take ? t2 : omin; oam = take ? a2 : oam; oq = take ? q2 : oq;
            }
            if (lane < 2) { s_obr[2 * lane] = omin; s_obr[2 * lane + 1] = oam; s_misc[2 + lane] = oq; }
        }
        if (SINGLE) { const long long c1 = clock64(); pc[3] += c1 - c0; c0 = c1; }
        __syncthreads();  // table entries, mirrored states, the list and the obstacle results are in place
        if (SINGLE) { const long long c1 = clock64(); pc[4] += c1 - c0; c0 = c1; }
        // ---- the listed rows again, one warp per row ----
        int nl;
        {
            nl = s_misc[lp];
            n_rows += nl;
            for (int q = warp; q < nl; q += nw) {
                const int k = s_list[q];
                double best;
                int bj;
                row_scan(k, lane, 32, best, bj);
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double v2 = __shfl_xor_sync(FULL, best, o);
                    const int j2 = __shfl_xor_sync(FULL, bj, o);
                    const bool take = v2 < best || (v2 == best && j2 >= 0 && (bj < 0 || j2 < bj));
                    best = take ? v2 : best;
                    bj = take ? j2 : bj;
                }
                if (lane == 0) { s_ct[k] = best; s_cj[k] = bj; }
            }
        }
        if (own && (li == a || li == b)) {
            const int w = li == a ? 0 : 1;
            ot = bl_add(t, s_obr[2 * w]); oarg = s_obr[2 * w + 1]; oidx = s_misc[2 + w];
            s_oarg[li] = oarg; s_oidx[li] = oidx;
            s_p0x[li] = p0x; s_p0y[li] = p0y; s_vx[li] = vx; s_vy[li] = vy; s_t0[li] = t0;
        }
        if (li == 0) s_misc[lp ^ 1] = 0;  // the other list: everybody is past its last use, the next append is a barrier away
        lp ^= 1;
        if (nl > 0) {
            __syncthreads();  // re-scanned minima in place
            if (rescan) { ct = s_ct[li]; cj = s_cj[li]; }
        }
        if (SINGLE) { const long long c1 = clock64(); pc[5] += c1 - c0; c0 = c1; }
        done++;
        n_bb += b >= 0 ? 1 : 0;
        now = t;
        kindmask = KIND_BOTH;  // only the first event's kind was forced
    }

    // ---- write the state, the table and the caches back ----
    __syncthreads();
    if (own && partial) row_scan(li, 0, 1, ct, cj);  // the caches leave as full row minima (what the other kernels expect)
    if (own) {
        P.p0x[sb + li] = p0x; P.p0y[sb + li] = p0y; P.vx[sb + li] = vx; P.vy[sb + li] = vy; P.t0[sb + li] = t0;
        P.obs_t[sb + li] = ot; P.obs_arg[sb + li] = oarg; P.obs_idx[sb + li] = oidx;
        if (SINGLE) { P.cover_t[li] = ct; P.cover_code[li] = cj >= 0 ? bl_pair_code(li, cj) : BL_CODE_NONE; }
    }
    {
        // both triangles of the table in global memory, one warp per row, lanes along the row
        double *const dst = P.toi + sys * P.toi_sys_stride;
        for (int i = warp; i < n; i += nw)
            for (int j = lane; j < n; j += 32) dst[(size_t)i * P.toi_pitch + j] = i == j ? BL_INF : tri[tri_at(i, j)];
    }
    if (li == 0) {
        if (SINGLE) {
            for (int q = 0; q < 16; q++) res->prof[q] = 0;
            res->prof[5] = done;
            res->prof[6] = n_rows;  // table rows scanned again
            for (int q = 0; q < 6; q++) res->prof[8 + q] = pc[q];
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

template <bool FMA, bool SINGLE, int MAXW>
cudaError_t launch_cta_w(const SmallParams &P, cudaStream_t st) {
    const size_t smem = (size_t)cta_layout(P.n).words * sizeof(double);
    auto kern = bl_evolve_cta_kernel<FMA, SINGLE, MAXW>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned threads = (unsigned)(((P.n + 31) / 32) * 32) + (SINGLE ? 32u : 0u);  // (single: + the obstacle warp)
    kern<<<SINGLE ? 1u : (unsigned)P.n_sys, threads, smem, st>>>(P);
    return cudaGetLastError();
}

template <bool FMA, bool SINGLE>
cudaError_t launch_cta(const SmallParams &P, cudaStream_t st) {
    if (!SINGLE && P.n <= 64) return launch_cta_w<FMA, SINGLE, 2>(P, st);
    if (!SINGLE && P.n <= 128) return launch_cta_w<FMA, SINGLE, 4>(P, st);
    return launch_cta_w<FMA, SINGLE, BL_CTA_MAX_WARPS>(P, st);
}

}  // namespace

size_t bl_cta_smem_bytes(int n) { return (size_t)cta_layout(n).words * sizeof(double); }

cudaError_t bl_launch_cta(const SmallParams &P, bool fma, bool single, cudaStream_t st) {
    if (single) return fma ? launch_cta<true, true>(P, st) : launch_cta<false, true>(P, st);
    return fma ? launch_cta<true, false>(P, st) : launch_cta<false, false>(P, st);
}
