// bl_evolve_tiny.cu -- the event loop for systems of at most FOUR balls, ONE THREAD PER SYSTEM.
//
// Galperin's pi-with-pool (two balls and a wall, 314159 strictly serial collisions; README.md:58-63 of the reference)
// is a pure latency problem: every collision depends on the one before, and there is nothing to spread over lanes --
// after a ball-ball collision of two balls there is no other pair to re-solve at all.  With one thread per system the
// whole state (5 doubles per ball, <= 6 pair times, <= 4 obstacle entries) lives in registers, the argmin is a short
// compare / select tree, and there are no shuffles, no shared-memory table, no warp synchronisation and no divergence:
// a collision costs the dependent FP64 chain of resolve -> re-solve and little else.
//
// Same contract as bl_evolve_small.cu (and therefore as bl_evolve.cu): same arithmetic (bl_arith.cuh), same keys
// (time, then ball-ball before ball-obstacle, then smallest row / column / ball), same stop rules, same result block.
// Serves a single Billiard (SINGLE) and ensembles of tiny systems (bl_multi_run with n <= 4: 128 systems per CTA).
// Reference lines: simulation.py:354-438, :440-502, :504-554, :556-560, :562-636, :176-209.
#include <cstring>
#include <type_traits>

#include "bl_small.cuh"

using namespace bl_small;

namespace {

template <int N>
__device__ __forceinline__ double pick(const double (&x)[N], int k) {
    double r = x[0];
#pragma unroll
    for (int q = 1; q < N; q++) r = k == q ? x[q] : r;
    return r;
}
template <int N>
__device__ __forceinline__ int pick(const int (&x)[N], int k) {
    int r = x[0];
#pragma unroll
    for (int q = 1; q < N; q++) r = k == q ? x[q] : r;
    return r;
}

// next obstacle of ball k and (if kb >= 0) of ball kb.  Lists of walls only (Galperin, pool, the gases) take a loop
// that handles both balls per wall without early exits, so that the two dependent chains (dot, dot, divide) overlap;
// anything else goes through bl_next_obstacle, one ball after the other.
template <bool FMA>
__device__ __forceinline__ void next_obstacle_pair(const bl_obstacle *s_obs, int n_obs, bool walls_only,
                                                   const double *rsum2_obs, int K, double t, bool two,  //
                                                   int rc1, double p1x, double p1y, double v1x, double v1y, double r1,
                                                   double &ot1, int &oi1, double &oa1,                   //
                                                   int rc2, double p2x, double p2y, double v2x, double v2y, double r2,
                                                   double &ot2, int &oi2, double &oa2) {
    if (walls_only) {
        double m1 = BL_INF, m2 = BL_INF, a1 = 0.0, a2 = 0.0;
        int q1 = -1, q2 = -1;
        for (int q = 0; q < n_obs; q++) {
            const double nx = s_obs[q].c, ny = s_obs[q].d, sx = s_obs[q].a, sy = s_obs[q].b;
            // obstacles.py:112-133: headway = -(v.n); gap = (p - start).n - radius; t = gap / headway
            const double h1 = -bl_dot2<FMA>(v1x, v1y, nx, ny), h2 = -bl_dot2<FMA>(v2x, v2y, nx, ny);
            const double g1 = bl_sub(bl_dot2<FMA>(bl_sub(p1x, sx), bl_sub(p1y, sy), nx, ny), r1);
            const double g2 = bl_sub(bl_dot2<FMA>(bl_sub(p2x, sx), bl_sub(p2y, sy), nx, ny), r2);
            const double u1 = bl_div(g1, h1);
            const bool ok1 = !(h1 <= 0.0) && !(u1 < BL_T_EPS);
            if (ok1 && u1 < m1) { m1 = u1; q1 = q; a1 = h1; }
            if (two) {
                const double u2 = bl_div(g2, h2);
                const bool ok2 = !(h2 <= 0.0) && !(u2 < BL_T_EPS);
                if (ok2 && u2 < m2) { m2 = u2; q2 = q; a2 = h2; }
            }
        }
        ot1 = bl_add(t, m1); oi1 = q1; oa1 = a1;
        if (two) { ot2 = bl_add(t, m2); oi2 = q2; oa2 = a2; }
        return;
    }
    bl_next_obstacle<FMA>(s_obs, n_obs, rsum2_obs, K, rc1, p1x, p1y, v1x, v1y, r1, t, ot1, oi1, oa1);
    if (two) bl_next_obstacle<FMA>(s_obs, n_obs, rsum2_obs, K, rc2, p2x, p2y, v2x, v2y, r2, t, ot2, oi2, oa2);
}

template <bool FMA, int N, bool SINGLE>
__global__ void __launch_bounds__(SINGLE ? 32 : 128, 1) bl_evolve_tiny_kernel(const __grid_constant__ SmallParams P) {
    __shared__ bl_obstacle s_obs[BL_SMALL_MAX_OBS];
    __shared__ int s_walls_only;
    const int n = P.n, K = P.K, n_obs = P.n_obs;
    if (threadIdx.x == 0) s_walls_only = 1;
    __syncthreads();
    for (int q = threadIdx.x; q < n_obs; q += blockDim.x) {
        s_obs[q] = P.obstacles[q];
        if (s_obs[q].kind != BL_OBS_WALL) s_walls_only = 0;
    }
    __syncthreads();
    const bool walls_only = s_walls_only != 0;
    const long long sys = SINGLE ? (long long)threadIdx.x : (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (sys >= P.n_sys) return;
    const long long sb = sys * n;
    // my system's composition (shared by all systems unless the strides say otherwise)
    const double *const c_mass = P.mass + sys * P.ball_stride, *const c_radius = P.radius + sys * P.ball_stride;
    const int32_t *const c_rclass = P.rclass + sys * P.ball_stride;
    const double *const c_rsum2 = P.rsum2 + sys * P.rs_stride, *const c_rsum2_obs = P.rsum2_obs + sys * P.rso_stride;

    double p0x[N], p0y[N], vx[N], vy[N], t0[N], mass[N], radius[N], ot[N], oarg[N];
    int rc[N], oidx[N];
    double T[N][N];   // T[lo][hi], lo < hi: the pair's absolute impact time (toi_table[hi][lo] of the reference)
    double rs[N][N];  // (r_lo + r_hi)**2 from the host's table
    double now = SINGLE ? P.time : P.sys_time[sys];
#pragma unroll
    for (int k = 0; k < N; k++) {
        const bool own = k < n;
        p0x[k] = own ? P.p0x[sb + k] : 0.0; p0y[k] = own ? P.p0y[sb + k] : 0.0;
        vx[k] = own ? P.vx[sb + k] : 0.0; vy[k] = own ? P.vy[sb + k] : 0.0; t0[k] = own ? P.t0[sb + k] : 0.0;
        mass[k] = own ? c_mass[k] : 1.0; radius[k] = own ? c_radius[k] : 0.0; rc[k] = own ? c_rclass[k] : 0;
        ot[k] = BL_INF; oarg[k] = 0.0; oidx[k] = -1;
    }
#pragma unroll
    for (int hi = 1; hi < N; hi++)
#pragma unroll
        for (int lo = 0; lo < hi; lo++) {
            rs[lo][hi] = c_rsum2[(size_t)rc[lo] * K + rc[hi]];
            T[lo][hi] = BL_INF;
        }
    if (!SINGLE && P.build_tables) {
        // add_ball, row by row at the current time (simulation.py:176-209)
        double px[N], py[N];
#pragma unroll
        for (int k = 0; k < N; k++) { px[k] = bl_advance(p0x[k], vx[k], now, t0[k]); py[k] = bl_advance(p0y[k], vy[k], now, t0[k]); }
#pragma unroll
        for (int hi = 1; hi < N; hi++)
#pragma unroll
            for (int lo = 0; lo < hi; lo++)
                if (hi < n)
                    T[lo][hi] = bl_add(now, bl_toi_ball_ball<FMA>(px[lo], py[lo], vx[lo], vy[lo], px[hi], py[hi], vx[hi],
                                                                  vy[hi], rs[lo][hi]));
#pragma unroll
        for (int k = 0; k < N; k++)
            if (k < n)
                bl_next_obstacle<FMA>(s_obs, n_obs, c_rsum2_obs, K, rc[k], px[k], py[k], vx[k], vy[k], radius[k], now,
                                      ot[k], oidx[k], oarg[k]);
    } else {
        const double *tab = P.toi + sys * P.toi_sys_stride;
#pragma unroll
        for (int hi = 1; hi < N; hi++)
#pragma unroll
            for (int lo = 0; lo < hi; lo++)
                if (hi < n) T[lo][hi] = tab[(size_t)hi * P.toi_pitch + lo];
#pragma unroll
        for (int k = 0; k < N; k++)
            if (k < n) { ot[k] = P.obs_t[sb + k]; oarg[k] = P.obs_arg[sb + k]; oidx[k] = P.obs_idx[sb + k]; }
    }

    const bool want_log = SINGLE && P.log.cap > 0;
    long long done = 0, n_bb = 0;
    int kindmask = KIND_BOTH;
    if (SINGLE) kindmask = P.first_kind == BL_BALL_BALL ? KIND_BB : (P.first_kind == BL_BALL_OBSTACLE ? KIND_BO : KIND_BOTH);
    int stop_kind = BL_STOP_TIME, st_status = BL_OK, err_i = -1, err_j = -1;
    double bb_t, bo_t, t_fail = 0.0;
    int bb_lo, bb_hi, bo_k;

    for (;;) {
        // ---- the two minima the reference caches (simulation.py:193-198, :205-209): first index on ties ----
        bb_t = BL_INF; bb_lo = -1; bb_hi = 0;
#pragma unroll
        for (int hi = 1; hi < N; hi++)
#pragma unroll
            for (int lo = 0; lo < hi; lo++) {
                const bool lt = T[lo][hi] < bb_t;
                bb_t = lt ? T[lo][hi] : bb_t; bb_lo = lt ? lo : bb_lo; bb_hi = lt ? hi : bb_hi;
            }
        bo_t = ot[0]; bo_k = 0;
#pragma unroll
        for (int k = 1; k < N; k++) {
            const bool lt = ot[k] < bo_t;
            bo_t = lt ? ot[k] : bo_t; bo_k = lt ? k : bo_k;
        }
        if (SINGLE && P.query_only) break;
        // a ball-ball collision wins a tie (simulation.py:421)
        const bool take_bb = kindmask == KIND_BB || (kindmask == KIND_BOTH && bb_lo >= 0 && !(bo_t < bb_t));
        const double t = take_bb ? bb_t : bo_t;
        int a = take_bb ? bb_lo : bo_k, b = take_bb ? bb_hi : -1;
        // Opaque to the optimiser on purpose.  nvcc 12.9 folded pick(x, b) to x[0] in the <N = 2, ensemble> instance of
        // this kernel (the PTX used ball 0 for both partners of a ball-ball collision: mass, position and velocity;
        // seen on the GPU as BL_EDIVZERO in the multi-ball ensemble parity test of tests/test_gpu_ensembles.py, case [2-1]).
        // With the indices hidden behind an empty asm every instance selects by the run-time value.  (Systems of one
        // or two balls have since moved to bl_evolve_pair_kernel below; the guard stays.)
        asm volatile("" : "+r"(a), "+r"(b));

        // ---- stop rules (same order as bl_evolve_small.cu) ----
        if (P.max_events >= 0 && done >= P.max_events) { stop_kind = BL_STOP_EVENTS; break; }
        if (t != t) { stop_kind = BL_STOP_ERROR; st_status = BL_ENAN; err_i = a; err_j = b; t_fail = t; break; }
        if (!(t <= P.end_time) || !(t < BL_INF) || a < 0) { stop_kind = BL_STOP_TIME; break; }
        if (want_log && done >= P.log.cap) { stop_kind = BL_STOP_LOG; break; }

        // ---- _move(t) for every ball (simulation.py:556-560), then resolve ----
        double px[N], py[N];
#pragma unroll
        for (int k = 0; k < N; k++) { px[k] = bl_advance(p0x[k], vx[k], t, t0[k]); py[k] = bl_advance(p0y[k], vy[k], t, t0[k]); }
        const double ax = pick(px, a), ay = pick(py, a), a_vx = pick(vx, a), a_vy = pick(vy, a);
        double avx = a_vx, avy = a_vy, bx = 0, by = 0, b_vx = 0, b_vy = 0, bvx = 0, bvy = 0, qarg = 0;
        int oi = -1, st;
        if (b < 0) {
            oi = pick(oidx, a); qarg = pick(oarg, a);
            st = oi >= 0 ? bl_obstacle_bounce<FMA>(s_obs[oi], ax, ay, a_vx, a_vy, qarg, avx, avy) : BL_EASSERT;
        } else {
            bx = pick(px, b); by = pick(py, b); b_vx = pick(vx, b); b_vy = pick(vy, b);
            double m1 = pick(mass, a), m2 = pick(mass, b);
            bl_remap_masses(m1, m2);
            bvx = b_vx; bvy = b_vy;
            st = bl_elastic<FMA>(ax, ay, a_vx, a_vy, m1, bx, by, b_vx, b_vy, m2, avx, avy, bvx, bvy);
        }
        if (st != BL_OK) {
            // the reference has moved to t and raised; the collision is not applied
            stop_kind = BL_STOP_ERROR; st_status = st; err_i = a; err_j = b; t_fail = t;
            break;
        }
        // ---- log, simulation.py:590-594 / :626-631 payloads ----
        if (want_log) {
            const long long e = done;
            P.log.t[e] = t; P.log.i[e] = a; P.log.j[e] = b >= 0 ? (long long)b : -1 - (long long)oi;
            if (P.log.payload) {
                double *o = P.log.payload + 12 * e;
                o[0] = ax; o[1] = ay; o[2] = a_vx; o[3] = a_vy; o[4] = avx; o[5] = avy;
                if (b >= 0) { o[6] = bx; o[7] = by; o[8] = b_vx; o[9] = b_vy; o[10] = bvx; o[11] = bvy; }
                else { o[6] = qarg; o[7] = 0; o[8] = 0; o[9] = 0; o[10] = 0; o[11] = 0; }
            }
        }
        // ---- new state of record of the colliding balls (:597-602, :634-635) ----
#pragma unroll
        for (int k = 0; k < N; k++) {
            if (k == a) { p0x[k] = ax; p0y[k] = ay; vx[k] = avx; vy[k] = avy; t0[k] = t; }
            if (k == b) { p0x[k] = bx; p0y[k] = by; vx[k] = bvx; vy[k] = bvy; t0[k] = t; }
        }
        // ---- re-solve every pair with a colliding ball (:461-471, :522-526); the colliding pair itself = inf (:458) ----
#pragma unroll
        for (int hi = 1; hi < N; hi++)
#pragma unroll
            for (int lo = 0; lo < hi; lo++) {
                const bool inv = lo == a || hi == a || lo == b || hi == b;
                if (inv && hi < n) {
                    if (lo == a && hi == b) T[lo][hi] = BL_INF;
                    else
                        T[lo][hi] = bl_add(t, bl_toi_ball_ball<FMA>(px[lo], py[lo], vx[lo], vy[lo], px[hi], py[hi], vx[hi],
                                                                   vy[hi], rs[lo][hi]));
                }
            }
        // ---- next obstacle of the colliding balls (:490-495, :545-547) ----
        {
            const int kb = b >= 0 ? b : a;
            double o1, o2, g1, g2;
            int i1, i2;
            next_obstacle_pair<FMA>(s_obs, n_obs, walls_only, c_rsum2_obs, K, t, b >= 0,  //
                                    pick(rc, a), ax, ay, avx, avy, pick(radius, a), o1, i1, g1,  //
                                    pick(rc, kb), bx, by, bvx, bvy, pick(radius, kb), o2, i2, g2);
#pragma unroll
            for (int k = 0; k < N; k++) {
                if (k == a) { ot[k] = o1; oidx[k] = i1; oarg[k] = g1; }
                if (k == b) { ot[k] = o2; oidx[k] = i2; oarg[k] = g2; }
            }
        }
        done++;
        n_bb += b >= 0 ? 1 : 0;
        now = t;
        kindmask = KIND_BOTH;  // only the first event's kind was forced
    }

    // ---- write the state, the table and the caches back ----
#pragma unroll
    for (int k = 0; k < N; k++)
        if (k < n) {
            P.p0x[sb + k] = p0x[k]; P.p0y[sb + k] = p0y[k]; P.vx[sb + k] = vx[k]; P.vy[sb + k] = vy[k];
            P.t0[sb + k] = t0[k];
            P.obs_t[sb + k] = ot[k]; P.obs_arg[sb + k] = oarg[k]; P.obs_idx[sb + k] = oidx[k];
        }
    {
        double *tab = P.toi + sys * P.toi_sys_stride;
#pragma unroll
        for (int i = 0; i < N; i++) {
            if (i >= n) continue;
            double mt = BL_INF;
            int mj = -1;
#pragma unroll
            for (int j = 0; j < N; j++) {
                if (j >= n) continue;
                const double v = i == j ? BL_INF : (i < j ? T[i][j] : T[j][i]);
                tab[(size_t)i * P.toi_pitch + j] = v;
                const bool lt = v < mt;
                mt = lt ? v : mt; mj = lt ? j : mj;
            }
            if (SINGLE) { P.cover_t[i] = mt; P.cover_code[i] = mj >= 0 ? bl_pair_code(i, mj) : BL_CODE_NONE; }
        }
    }
    if (SINGLE) {
        bl_result *const res = P.result;
        double tnow = now;
        if (st_status != BL_OK) tnow = t_fail;
        if (stop_kind == BL_STOP_TIME && P.end_time < BL_INF) tnow = P.end_time;  // _move(end_time), :436
        if (P.query_only) { tnow = P.time; stop_kind = BL_STOP_TIME; }
        res->time = tnow; res->stop = stop_kind; res->status = st_status;
        res->err_i = err_i; res->err_j = err_j;
        res->n_ball_ball = n_bb; res->n_ball_obstacle = done - n_bb; res->n_logged = want_log ? done : 0;
        const bool some = bb_lo >= 0 && bb_t < BL_INF;
        res->bb_t = some ? bb_t : BL_INF; res->bb_i = some ? bb_lo : -1; res->bb_j = some ? bb_hi : 0;
        // numpy argmin over _obstacles_toi: first ball attaining the minimum, even when it is +inf
        res->bo_t = bo_t; res->bo_ball = bo_k; res->bo_obs = pick(oidx, bo_k); res->bo_arg = pick(oarg, bo_k);
        for (int q = 0; q < 16; q++) res->prof[q] = 0;
        res->prof[5] = done;
        res->window = 0.0;
        *P.seq_counter += done;
    } else {
        P.sys_time[sys] = st_status != BL_OK ? t_fail
                          : ((stop_kind == BL_STOP_TIME && P.end_time < BL_INF) ? P.end_time : now);
        P.n_bb[sys] += n_bb; P.n_bo[sys] += done - n_bb;
        P.status[sys] = st_status;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Systems of ONE or TWO balls (Galperin): the same loop with every ball index resolved by control flow instead of
// select chains -- a ball-ball collision can only be (0, 1), an obstacle collision is "ball `me` against the list,
// then the pair with `other`" with the two roles passed by reference.  One thread per system, 314159 strictly serial
// collisions: what counts is the length of the dependent instruction chain per collision (ncu on the generic
// instance above: 462 instructions and 1720 cycles per collision, ~150 of them FP64).  Hence
//   * one compare each for the two stop rules (a count limit and a time limit prepared before the loop, held in
//     registers: re-reading kernel parameters from the constant bank costs ~25 cycles of latency per read);
//   * the first two walls of a walls-only list live in registers (no shared-memory address arithmetic per collision);
//   * the two quotients of the impulse share one reciprocal, an exactly-zero numerator (one-dimensional motion: every
//     y-component) skips the division, and after a ball-ball collision the two balls' wall divisions are interleaved.
struct Ball {
    double p0x, p0y, vx, vy, t0, mass, radius, ot, oarg;
    int rc, oidx;
};
struct Wall {
    double sx, sy, nx, ny;
};

// physics.py:273-282 like bl_elastic; the quotient of an exactly-zero numerator over a finite non-zero denominator is
// +-0 with the sign of the operands (what __ddiv_rn returns, but through its slow path)
template <bool FMA>
__device__ __forceinline__ int elastic_pair(double p1x, double p1y, double v1x, double v1y, double m1, double p2x,
                                            double p2y, double v2x, double v2y, double m2, double &w1x, double &w1y,
                                            double &w2x, double &w2y) {
    const double pdx = bl_sub(p2x, p1x), pdy = bl_sub(p2y, p1y);
    const double vdx = bl_sub(v2x, v1x), vdy = bl_sub(v2y, v1y);
    const double pdv = bl_dot2<FMA>(pdx, pdy, vdx, vdy);
    if (pdv > BL_SEP_EPS) return BL_ESEPARATING;
    const double den = bl_mul(bl_add(m1, m2), bl_dot2<FMA>(pdx, pdy, pdx, pdy));
    if (den == 0.0) return BL_EDIVZERO;
    const double nx = bl_mul(2.0, bl_mul(pdv, pdx)), ny = bl_mul(2.0, bl_mul(pdv, pdy));
    const double r = bl_div_rcp(den);
    const double ix = bl_div_by(nx, den, r);
    double iy;
    if (ny == 0.0 && den == den) iy = __hiloint2double((__double2hiint(ny) ^ __double2hiint(den)) & 0x80000000, 0);
    else iy = bl_div_by(ny, den, r);
    w1x = bl_add(v1x, bl_mul(m2, ix));
    w1y = bl_add(v1y, bl_mul(m2, iy));
    w2x = bl_sub(v2x, bl_mul(m1, ix));
    w2y = bl_sub(v2y, bl_mul(m1, iy));
    return BL_OK;
}

// one wall against one ball (obstacles.py:112-133), folded into the running strict minimum
template <bool FMA>
__device__ __forceinline__ void wall_fold(const Wall &w, int q, const Ball &B, double &m, int &qm, double &arg) {
    const double h = -bl_dot2<FMA>(B.vx, B.vy, w.nx, w.ny);
    if (h <= 0.0) return;  // moving away from this wall (:117-118)
    const double g = bl_sub(bl_dot2<FMA>(bl_sub(B.p0x, w.sx), bl_sub(B.p0y, w.sy), w.nx, w.ny), B.radius);
    const double u = bl_div(g, h);
    if (!(u < BL_T_EPS) && u < m) { m = u; qm = q; arg = h; }
}
// the same for two balls at once: both divisions issued side by side (a ball that moves away divides by 1 and is
// ignored), so that their dependent chains overlap
template <bool FMA>
__device__ __forceinline__ void wall_fold2(const Wall &w, int q, const Ball &A, const Ball &B, double &ma, int &qa,
                                           double &arga, double &mb, int &qb, double &argb) {
    const double ha = -bl_dot2<FMA>(A.vx, A.vy, w.nx, w.ny), hb = -bl_dot2<FMA>(B.vx, B.vy, w.nx, w.ny);
    const bool offa = ha <= 0.0, offb = hb <= 0.0;
    if (offa && offb) return;
    const double ga = bl_sub(bl_dot2<FMA>(bl_sub(A.p0x, w.sx), bl_sub(A.p0y, w.sy), w.nx, w.ny), A.radius);
    const double gb = bl_sub(bl_dot2<FMA>(bl_sub(B.p0x, w.sx), bl_sub(B.p0y, w.sy), w.nx, w.ny), B.radius);
    const double da = offa ? 1.0 : ha, db = offb ? 1.0 : hb;
    const double ra = bl_div_rcp(da), rb = bl_div_rcp(db);
    const double ua = bl_div_by(ga, da, ra), ub = bl_div_by(gb, db, rb);
    if (!offa && !(ua < BL_T_EPS) && ua < ma) { ma = ua; qa = q; arga = ha; }
    if (!offb && !(ub < BL_T_EPS) && ub < mb) { mb = ub; qb = q; argb = hb; }
}

struct Obstacles {
    const bl_obstacle *s_obs;  // the list in shared memory
    const double *rsum2_obs;
    Wall w0, w1;               // the first two walls of a walls-only list
    int n_obs, K;
    bool walls_only;
};

__device__ __forceinline__ Wall wall_at(const Obstacles &O, int q) {
    Wall w;
    w.sx = O.s_obs[q].a; w.sy = O.s_obs[q].b; w.nx = O.s_obs[q].c; w.ny = O.s_obs[q].d;
    return w;
}

// next obstacle of one ball whose state of record was just re-based at time t (simulation.py:332-352)
template <bool FMA>
__device__ __forceinline__ void next_obstacle_ball(const Obstacles &O, double t, Ball &B) {
    if (O.walls_only) {
        double m = BL_INF, arg = 0.0;
        int qm = -1;
        if (O.n_obs > 0) {
            wall_fold<FMA>(O.w0, 0, B, m, qm, arg);
            if (O.n_obs > 1) {
                wall_fold<FMA>(O.w1, 1, B, m, qm, arg);
#pragma unroll 1
                for (int q = 2; q < O.n_obs; q++) wall_fold<FMA>(wall_at(O, q), q, B, m, qm, arg);
            }
        }
        B.ot = bl_add(t, m); B.oidx = qm; B.oarg = arg;
        return;
    }
    bl_next_obstacle<FMA>(O.s_obs, O.n_obs, O.rsum2_obs, O.K, B.rc, B.p0x, B.p0y, B.vx, B.vy, B.radius, t, B.ot, B.oidx,
                          B.oarg);
}
template <bool FMA>
__device__ __forceinline__ void next_obstacle_both(const Obstacles &O, double t, Ball &A, Ball &B) {
    if (O.walls_only) {
        double ma = BL_INF, mb = BL_INF, arga = 0.0, argb = 0.0;
        int qa = -1, qb = -1;
        if (O.n_obs > 0) {
            wall_fold2<FMA>(O.w0, 0, A, B, ma, qa, arga, mb, qb, argb);
            if (O.n_obs > 1) {
                wall_fold2<FMA>(O.w1, 1, A, B, ma, qa, arga, mb, qb, argb);
#pragma unroll 1
                for (int q = 2; q < O.n_obs; q++) wall_fold2<FMA>(wall_at(O, q), q, A, B, ma, qa, arga, mb, qb, argb);
            }
        }
        A.ot = bl_add(t, ma); A.oidx = qa; A.oarg = arga;
        B.ot = bl_add(t, mb); B.oidx = qb; B.oarg = argb;
        return;
    }
    next_obstacle_ball<FMA>(O, t, A);
    next_obstacle_ball<FMA>(O, t, B);
}

// one obstacle collision of `me` at time t (simulation.py:504-554 with N <= 2); returns a bl_status
template <bool FMA>
__device__ __forceinline__ int obstacle_event(const Obstacles &O, bool two, double rs01, double t, Ball &me,
                                              const Ball &other, double &T01, double *payload) {
    const double ax = bl_advance(me.p0x, me.vx, t, me.t0), ay = bl_advance(me.p0y, me.vy, t, me.t0);
    double wx = me.vx, wy = me.vy;
    const int oi = me.oidx;
    if (oi < 0) return BL_EASSERT;
    if (O.walls_only) {  // InfiniteWall.resolve_collision, obstacles.py:135-144
        const double nx = oi == 0 ? O.w0.nx : (oi == 1 ? O.w1.nx : O.s_obs[oi].c);
        const double ny = oi == 0 ? O.w0.ny : (oi == 1 ? O.w1.ny : O.s_obs[oi].d);
        wx = bl_add(me.vx, bl_mul(2.0, bl_mul(me.oarg, nx)));
        wy = bl_add(me.vy, bl_mul(2.0, bl_mul(me.oarg, ny)));
    } else {
        const int st = bl_obstacle_bounce<FMA>(O.s_obs[oi], ax, ay, me.vx, me.vy, me.oarg, wx, wy);
        if (st != BL_OK) return st;
    }
    if (payload) {
        payload[0] = ax; payload[1] = ay; payload[2] = me.vx; payload[3] = me.vy; payload[4] = wx; payload[5] = wy;
        payload[6] = me.oarg; payload[7] = 0; payload[8] = 0; payload[9] = 0; payload[10] = 0; payload[11] = 0;
    }
    me.p0x = ax; me.p0y = ay; me.vx = wx; me.vy = wy; me.t0 = t;
    if (two) {  // the one pair of the system is re-solved (:522-526)
        const double ox = bl_advance(other.p0x, other.vx, t, other.t0), oy = bl_advance(other.p0y, other.vy, t, other.t0);
        T01 = bl_add(t, bl_toi_ball_ball<FMA>(ax, ay, wx, wy, ox, oy, other.vx, other.vy, rs01));
    }
    next_obstacle_ball<FMA>(O, t, me);
    return BL_OK;
}

// LOG: the launch writes the device event log (a compile-time flag: three fewer branches per collision otherwise)
template <bool FMA, bool SINGLE, bool LOG>
__global__ void __launch_bounds__(SINGLE ? 32 : 128, 1) bl_evolve_pair_kernel(const __grid_constant__ SmallParams P) {
    __shared__ bl_obstacle s_obs[BL_SMALL_MAX_OBS];
    __shared__ int s_walls_only;
    const int n = P.n, K = P.K, n_obs = P.n_obs;
    if (threadIdx.x == 0) s_walls_only = 1;
    __syncthreads();
    for (int q = threadIdx.x; q < n_obs; q += blockDim.x) {
        s_obs[q] = P.obstacles[q];
        if (s_obs[q].kind != BL_OBS_WALL) s_walls_only = 0;
    }
    __syncthreads();
    const long long sys = SINGLE ? (long long)threadIdx.x : (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (sys >= P.n_sys) return;
    const long long sb = sys * n;
    // my system's composition (shared by all systems unless the strides say otherwise)
    const double *const c_mass = P.mass + sys * P.ball_stride, *const c_radius = P.radius + sys * P.ball_stride;
    const int32_t *const c_rclass = P.rclass + sys * P.ball_stride;
    const double *const c_rsum2 = P.rsum2 + sys * P.rs_stride, *const c_rsum2_obs = P.rsum2_obs + sys * P.rso_stride;
    bool two = n == 2;
    {
        int tw = two ? 1 : 0;
        asm volatile("" : "+r"(tw));
        two = tw != 0;
    }
    Obstacles O;
    O.s_obs = s_obs; O.rsum2_obs = c_rsum2_obs; O.n_obs = n_obs; O.K = K; O.walls_only = s_walls_only != 0;
    {
        // loop-invariant flags as opaque register values: left to itself the compiler re-derives them from the
        // constant bank inside the loop (load + compare + branch, ~20 cycles of latency each time, several per collision)
        int wo = O.walls_only ? 1 : 0, no = n_obs;
        asm volatile("" : "+r"(wo), "+r"(no));
        O.walls_only = wo != 0; O.n_obs = no;
    }
    O.w0.sx = O.w0.sy = O.w0.nx = O.w0.ny = 0.0;
    O.w1 = O.w0;
    if (O.walls_only && n_obs > 0) O.w0 = wall_at(O, 0);
    if (O.walls_only && n_obs > 1) O.w1 = wall_at(O, 1);

    Ball b0, b1;
    b0.p0x = P.p0x[sb]; b0.p0y = P.p0y[sb]; b0.vx = P.vx[sb]; b0.vy = P.vy[sb]; b0.t0 = P.t0[sb];
    b0.mass = c_mass[0]; b0.radius = c_radius[0]; b0.rc = c_rclass[0];
    b0.ot = BL_INF; b0.oarg = 0.0; b0.oidx = -1;
    b1 = b0;
    b1.p0x = 0.0; b1.p0y = 0.0; b1.vx = 0.0; b1.vy = 0.0; b1.t0 = 0.0; b1.mass = 1.0; b1.radius = 0.0; b1.rc = 0;
    if (two) {
        b1.p0x = P.p0x[sb + 1]; b1.p0y = P.p0y[sb + 1]; b1.vx = P.vx[sb + 1]; b1.vy = P.vy[sb + 1]; b1.t0 = P.t0[sb + 1];
        b1.mass = c_mass[1]; b1.radius = c_radius[1]; b1.rc = c_rclass[1];
    }
    const double rs01 = c_rsum2[(size_t)b0.rc * K + b1.rc];
    double m1r = b0.mass, m2r = b1.mass;  // simulation.py:579-582, the same for every collision of the pair
    bl_remap_masses(m1r, m2r);
    double T01 = BL_INF;
    double now = SINGLE ? P.time : P.sys_time[sys];
    double *const tab = P.toi + sys * P.toi_sys_stride;
    if (!SINGLE && P.build_tables) {
        // add_ball, row by row at the current time (simulation.py:176-209)
        const double x0 = bl_advance(b0.p0x, b0.vx, now, b0.t0), y0 = bl_advance(b0.p0y, b0.vy, now, b0.t0);
        bl_next_obstacle<FMA>(s_obs, n_obs, c_rsum2_obs, K, b0.rc, x0, y0, b0.vx, b0.vy, b0.radius, now, b0.ot, b0.oidx, b0.oarg);
        if (two) {
            const double x1 = bl_advance(b1.p0x, b1.vx, now, b1.t0), y1 = bl_advance(b1.p0y, b1.vy, now, b1.t0);
            T01 = bl_add(now, bl_toi_ball_ball<FMA>(x0, y0, b0.vx, b0.vy, x1, y1, b1.vx, b1.vy, rs01));
            bl_next_obstacle<FMA>(s_obs, n_obs, c_rsum2_obs, K, b1.rc, x1, y1, b1.vx, b1.vy, b1.radius, now, b1.ot, b1.oidx, b1.oarg);
        }
    } else {
        b0.ot = P.obs_t[sb]; b0.oarg = P.obs_arg[sb]; b0.oidx = P.obs_idx[sb];
        if (two) {
            b1.ot = P.obs_t[sb + 1]; b1.oarg = P.obs_arg[sb + 1]; b1.oidx = P.obs_idx[sb + 1];
            T01 = tab[P.toi_pitch];
        }
    }

    constexpr bool want_log = LOG;
    // the two limits, in registers: collisions this launch may handle (max_events and the log's capacity) and the
    // last admissible time (finite: an impact time of +inf or NaN fails `t <= t_last` too)
    long long limit = P.max_events >= 0 ? P.max_events : 0x7fffffffffffffffLL;
    if (want_log && P.log.cap < limit) limit = P.log.cap;
    double t_last = P.end_time < 1.7976931348623157e308 ? P.end_time : 1.7976931348623157e308;
    asm volatile("" : "+l"(limit), "+d"(t_last));
    long long done = 0, n_bb = 0;
    int kind = KIND_BOTH;
    if (SINGLE) kind = P.first_kind == BL_BALL_BALL ? KIND_BB : (P.first_kind == BL_BALL_OBSTACLE ? KIND_BO : KIND_BOTH);
    // why the loop ended: 0 a limit (which one is sorted out below), 4 collision not resolvable (status in st)
    int why = 0, st = BL_OK, a = -1, b = -1;
    double t = now;
    bool o1 = false, take_bb = false;
    if (!(SINGLE && P.query_only)) {
        // The loop exists twice: once with "walls only" and "two balls" known at compile time (Galperin's set-up: ~6
        // fewer branches per collision, each worth 15-25 cycles to a lone warp), once with both read from registers.
        auto event_loop = [&](auto fast_tag) {
            constexpr int FAST = decltype(fast_tag)::value;  // 0 generic, 1 walls only + two balls, 2 ... + exactly one wall
            Obstacles OO = O;
            if (FAST) OO.walls_only = true;
            if (FAST == 2) OO.n_obs = 1;
            const bool two_ = FAST ? true : two;
            for (;;) {
                // next_ball_ball_collision / next_ball_obstacle_collision: first ball on ties, ball-ball wins a tie
                o1 = b1.ot < b0.ot;
                const double bo_t = o1 ? b1.ot : b0.ot;
                take_bb = kind == KIND_BOTH ? !(bo_t < T01) : kind == KIND_BB;
                t = take_bb ? T01 : bo_t;
                if (done >= limit || !(t <= t_last)) break;
                double *payload = nullptr;
                if (want_log) {
                    P.log.t[done] = t;
                    if (P.log.payload) payload = P.log.payload + 12 * done;
                }
                if (take_bb) {
                    // bounce_ball_ball, simulation.py:440-502 with N = 2: nothing else to re-solve
                    const double ax = bl_advance(b0.p0x, b0.vx, t, b0.t0), ay = bl_advance(b0.p0y, b0.vy, t, b0.t0);
                    const double bx = bl_advance(b1.p0x, b1.vx, t, b1.t0), by = bl_advance(b1.p0y, b1.vy, t, b1.t0);
                    const double m1 = m1r, m2 = m2r;
                    double avx = b0.vx, avy = b0.vy, bvx = b1.vx, bvy = b1.vy;
                    st = elastic_pair<FMA>(ax, ay, b0.vx, b0.vy, m1, bx, by, b1.vx, b1.vy, m2, avx, avy, bvx, bvy);
                    if (st != BL_OK) { why = 4; a = 0; b = 1; break; }
                    if (want_log) {
                        P.log.i[done] = 0; P.log.j[done] = 1;
                        if (payload) {
                            payload[0] = ax; payload[1] = ay; payload[2] = b0.vx; payload[3] = b0.vy; payload[4] = avx; payload[5] = avy;
                            payload[6] = bx; payload[7] = by; payload[8] = b1.vx; payload[9] = b1.vy; payload[10] = bvx; payload[11] = bvy;
                        }
                    }
                    b0.p0x = ax; b0.p0y = ay; b0.vx = avx; b0.vy = avy; b0.t0 = t;
                    b1.p0x = bx; b1.p0y = by; b1.vx = bvx; b1.vy = bvy; b1.t0 = t;
                    T01 = BL_INF;  // :458
                    next_obstacle_both<FMA>(OO, t, b0, b1);
                    n_bb++;
                } else {
                    if (want_log) { P.log.i[done] = (int)o1; P.log.j[done] = -1 - (long long)(o1 ? b1.oidx : b0.oidx); }
                    st = o1 ? obstacle_event<FMA>(OO, two_, rs01, t, b1, b0, T01, payload)
                            : obstacle_event<FMA>(OO, two_, rs01, t, b0, b1, T01, payload);
                    if (st != BL_OK) { why = 4; a = (int)o1; b = -1; break; }
                }
                done++;
                now = t;
                kind = KIND_BOTH;  // only the first event's kind was forced
            }
        };
        if (O.walls_only && two && O.n_obs == 1) event_loop(std::integral_constant<int, 2>{});
        else if (O.walls_only && two) event_loop(std::integral_constant<int, 1>{});
        else event_loop(std::integral_constant<int, 0>{});
    }
    // which limit ended the loop, in the order of the other kernels' stop rules: count, NaN, time, log
    if (why == 0 && !(SINGLE && P.query_only)) {
        if (P.max_events >= 0 && done >= P.max_events) why = 1;
        else if (t != t) { why = 3; a = take_bb ? 0 : (int)o1; b = take_bb ? 1 : -1; }
        else if (!(t <= t_last)) why = 0;
        else why = 2;
    }
    o1 = b1.ot < b0.ot;

    // ---- write the state, the table and the caches back ----
    P.p0x[sb] = b0.p0x; P.p0y[sb] = b0.p0y; P.vx[sb] = b0.vx; P.vy[sb] = b0.vy; P.t0[sb] = b0.t0;
    P.obs_t[sb] = b0.ot; P.obs_arg[sb] = b0.oarg; P.obs_idx[sb] = b0.oidx;
    tab[0] = BL_INF;
    if (two) {
        P.p0x[sb + 1] = b1.p0x; P.p0y[sb + 1] = b1.p0y; P.vx[sb + 1] = b1.vx; P.vy[sb + 1] = b1.vy; P.t0[sb + 1] = b1.t0;
        P.obs_t[sb + 1] = b1.ot; P.obs_arg[sb + 1] = b1.oarg; P.obs_idx[sb + 1] = b1.oidx;
        tab[1] = T01; tab[P.toi_pitch] = T01; tab[P.toi_pitch + 1] = BL_INF;
    }
    const bool failed = why >= 3;
    const int status = why == 3 ? (int)BL_ENAN : (why == 4 ? st : (int)BL_OK);
    if (SINGLE) {
        const bool some = T01 < BL_INF;
        P.cover_t[0] = some ? T01 : BL_INF; P.cover_code[0] = some ? bl_pair_code(0, 1) : BL_CODE_NONE;
        if (two) { P.cover_t[1] = some ? T01 : BL_INF; P.cover_code[1] = some ? bl_pair_code(0, 1) : BL_CODE_NONE; }
        bl_result *const res = P.result;
        double tnow = failed ? t : now;
        if (why == 0 && P.end_time < BL_INF) tnow = P.end_time;  // _move(end_time), :436
        if (P.query_only) tnow = P.time;
        res->time = tnow;
        res->stop = failed ? BL_STOP_ERROR : (why == 1 ? BL_STOP_EVENTS : (why == 2 ? BL_STOP_LOG : BL_STOP_TIME));
        res->status = status;
        res->err_i = failed ? a : -1; res->err_j = failed ? b : -1;
        res->n_ball_ball = n_bb; res->n_ball_obstacle = done - n_bb; res->n_logged = want_log ? done : 0;
        res->bb_t = some ? T01 : BL_INF; res->bb_i = some ? 0 : -1; res->bb_j = some ? 1 : 0;
        // numpy argmin over _obstacles_toi: first ball attaining the minimum, even when it is +inf
        res->bo_t = o1 ? b1.ot : b0.ot; res->bo_ball = (int)o1;
        res->bo_obs = o1 ? b1.oidx : b0.oidx; res->bo_arg = o1 ? b1.oarg : b0.oarg;
        for (int q = 0; q < 16; q++) res->prof[q] = 0;
        res->prof[5] = done;
        res->window = 0.0;
        *P.seq_counter += done;
    } else {
        P.sys_time[sys] = failed ? t : ((why == 0 && P.end_time < BL_INF) ? P.end_time : now);
        P.n_bb[sys] += n_bb; P.n_bo[sys] += done - n_bb;
        P.status[sys] = status;
    }
}

template <bool FMA, bool SINGLE>
cudaError_t launch_n(const SmallParams &P, cudaStream_t st) {
    const unsigned grid = SINGLE ? 1u : (unsigned)((P.n_sys + 127) / 128);
    const unsigned block = SINGLE ? 32u : 128u;
    if (P.n <= 2) {
        if (SINGLE && P.log.cap > 0) bl_evolve_pair_kernel<FMA, SINGLE, SINGLE><<<grid, block, 0, st>>>(P);
        else bl_evolve_pair_kernel<FMA, SINGLE, false><<<grid, block, 0, st>>>(P);
    }
    else bl_evolve_tiny_kernel<FMA, 4, SINGLE><<<grid, block, 0, st>>>(P);
    return cudaGetLastError();
}

}  // namespace

cudaError_t bl_launch_tiny(const SmallParams &P, bool fma, bool single, int sm_count, cudaStream_t st) {
    (void)sm_count;
    if (P.n < 1 || P.n > BL_TINY_MAX_BALLS) return cudaErrorInvalidValue;
    if (fma) return single ? launch_n<true, true>(P, st) : launch_n<true, false>(P, st);
    return single ? launch_n<false, true>(P, st) : launch_n<false, false>(P, st);
}
