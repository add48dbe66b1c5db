// bl_evolve.cu -- the event loop of Billiard.evolve as ONE persistent kernel.
//
// Replaces simulation.py:354-438 (evolve), :440-502 (bounce_ball_ball), :504-554 (bounce_ball_obstacle),
// :556-560 (_move), :562-636 (_resolve_*), with physics.py:10-76/:255-282 and obstacles.py:70-76/:112-144
// inlined from bl_arith.cuh.
//
// Shape of the computation: events are strictly serial (each needs a global argmin and changes state the
// next one reads), so the kernel is latency-bound; it runs as ONE thread-block cluster (1..16 CTAs on
// neighbouring SMs) whose barrier costs a few hundred cycles instead of a grid-wide sync:
//   - every ball is owned by one thread; its state of record (p0, v, t0, mass, radius), its candidate minimum
//     and its next-obstacle entry stay resident in that CTA's shared memory for the whole launch;
//   - the all-pairs table of absolute impact times stays in HBM/L2 (symmetric, n x pitch doubles); per event the
//     owners of all other balls solve their pair with the colliding ball(s) (2(N-2) solves) and store the row
//     entry (coalesced) and the column entry (strided);
//   - the argmin is a warp REDUX -> CTA -> cluster reduction over (time, code) keys whose lexicographic order
//     reproduces numpy's first-index tie-breaking; CTA winners are exchanged through distributed shared memory.
//
// "Cover" instead of row minima: ball k's candidate m[k] is the minimum over a SUBSET of row k that
// always contains every pair whose value was last written at an event of the OTHER ball.  After an event of ball
// a every pair (a,k) is re-solved and folded into m[k] by k's owner, so m[a] can simply be emptied -- no
// cross-thread reduction for the rows of the colliding balls.  If k's candidate was its pair with a and the new
// value is larger, m[k] keeps the old time as a lower bound ("stale"); stale rows are re-read from the table only
// when one of them wins the argmin, all stale rows of the cluster at once (bandwidth-bound, amortised).
// The global minimum over {m[k]} equals the minimum over the whole table, so the committed event order is the
// reference's, bit for bit.
#include <cstdio>

#include "bl_arith.cuh"
#include "bl_internal.h"

namespace {

// ---- per-ball fields resident in shared memory: fd[F_*][cap] doubles, then u64 / int arrays ----------
enum {
    F_P0X = 0, F_P0Y, F_VX, F_VY, F_T0, F_MASS,          // state of record of the ball (payload words 2..7)
    F_QP0X, F_QP0Y, F_QVX, F_QVY, F_QT0, F_QMASS,        // cached state of the candidate's partner (words 8..13)
    F_OARG,                                              // obstacle arg (word 16)
    F_MT, F_OT, F_RADIUS,
    NFD
};
enum { I_RC = 0, I_QRC, I_OIDX, NFI };

// payload exchanged between CTAs (8-byte words)
enum {
    W_TS = 0, W_CODE, W_K0 = 2, W_Q0 = 8, W_INT0 = 14, W_INT1 = 15, W_OARG = 16, NW = 17, NW_PAD = 18
};

enum { EV_STOP = 0, EV_BB = 1, EV_BO = 2, EV_RESCAN = 3 };
enum { MODE_BOTH = 0, MODE_BB = 1, MODE_BO = 2 };

struct Bcast {
    double t;
    int kind, a, b, rca, rcb, status, stop, obs;
    double ax, ay, avx, avy, amass;  // ball a after the event (position = new p0, t0 = t)
    double bx, by, bvx, bvy, bmass;
    uint64_t ts, code;
    double q_arg;
    int k, pad;
};

struct Params {
    bl_system s;
    double time, end_time;
    long long max_events;
    int first_kind;
    int query_only;
    bl_log log;
    bl_result *result;
    int cap;   // balls per CTA
    int bpt;   // balls per thread
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_u64(uint32_t addr, uint64_t v) {
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// lane holding the lexicographically smallest (ts, code); 4 REDUX + 1 ballot
__device__ __forceinline__ int warp_argmin(uint64_t ts, uint64_t code) {
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
    unsigned b = __ballot_sync(FULL, c);
    return __ffs(b) - 1;
}

struct Smem {
    double *fd;        // [NFD][cap]
    uint64_t *mcode;   // [cap]
    int *fi;           // [NFI][cap]
    int *dirty;        // [cap]
    bl_obstacle *obs;  // [n_obs]
    uint64_t *slots;   // [32][4]  (ts, code, local ball, pad)
    uint64_t *inbox;   // [2][16][NW_PAD]
    Bcast *bc;
    int *n_dirty;
    int cap;
    __device__ __forceinline__ double &d(int f, int k) const { return fd[f * cap + k]; }
    __device__ __forceinline__ int &i(int f, int k) const { return fi[f * cap + k]; }
};

__host__ __device__ inline size_t smem_bytes(int cap, int n_obs) {
    size_t b = 0;
    b += (size_t)NFD * cap * 8;
    b += (size_t)cap * 8;
    b += (size_t)NFI * cap * 4;
    b += (size_t)cap * 4;
    b = (b + 15) & ~(size_t)15;
    b += (size_t)(n_obs > 0 ? n_obs : 1) * sizeof(bl_obstacle);
    b += 32 * 4 * 8;
    b += 2 * 16 * NW_PAD * 8;
    b += sizeof(Bcast) + 16;
    b += 16;
    return b;
}

__device__ __forceinline__ Smem carve(unsigned char *raw, int cap, int n_obs) {
    Smem m;
    m.cap = cap;
    unsigned char *p = raw;
    m.fd = (double *)p; p += (size_t)NFD * cap * 8;
    m.mcode = (uint64_t *)p; p += (size_t)cap * 8;
    m.fi = (int *)p; p += (size_t)NFI * cap * 4;
    m.dirty = (int *)p; p += (size_t)cap * 4;
    p = (unsigned char *)(((uintptr_t)p + 15) & ~(uintptr_t)15);
    m.obs = (bl_obstacle *)p; p += (size_t)(n_obs > 0 ? n_obs : 1) * sizeof(bl_obstacle);
    m.slots = (uint64_t *)p; p += 32 * 4 * 8;
    m.inbox = (uint64_t *)p; p += 2 * 16 * NW_PAD * 8;
    m.bc = (Bcast *)p; p += (sizeof(Bcast) + 15) & ~(size_t)15;
    m.n_dirty = (int *)p;
    return m;
}

// ---------------------------------------------------------------------------------------------
template <bool FMA>
__global__ void __launch_bounds__(BL_MAX_THREADS, 1) bl_evolve_kernel(const Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const bl_system &S = P.s;
    const int n = S.n, K = S.n_classes, n_obs = S.n_obstacles;
    const size_t pitch = (size_t)S.pitch;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x, nwarps = nthreads >> 5;
    const int rank = (int)cluster_ctarank(), C = (int)cluster_nctarank();
    const int cap = P.cap, bpt = P.bpt;
    const Smem M = carve(smem_raw, cap, n_obs);
    Bcast &bc = *M.bc;

    // ball owned by (rank, tid, slot u): local index kl = u*nthreads + tid, global index k = rank*cap + kl
    const int kbase = rank * cap;

    // ---- load the resident state --------------------------------------------------------------
    for (int q = tid; q < n_obs; q += nthreads) M.obs[q] = S.obstacles[q];
    for (int u = 0; u < bpt; u++) {
        const int kl = u * nthreads + tid, k = kbase + kl;
        if (k < n) {
            M.d(F_P0X, kl) = S.p0x[k]; M.d(F_P0Y, kl) = S.p0y[k];
            M.d(F_VX, kl) = S.vx[k];   M.d(F_VY, kl) = S.vy[k];
            M.d(F_T0, kl) = S.t0[k];   M.d(F_MASS, kl) = S.mass[k];
            M.d(F_RADIUS, kl) = S.radius[k];
            M.i(I_RC, kl) = S.rclass[k];
            M.d(F_OT, kl) = S.obs_t[k]; M.i(I_OIDX, kl) = S.obs_idx[k]; M.d(F_OARG, kl) = S.obs_arg[k];
            const double mt = S.cover_t[k];
            const uint64_t mc = S.cover_code[k];
            M.d(F_MT, kl) = mt; M.mcode[kl] = mc;
            int q = -1;
            if (mc != BL_CODE_STALE && mc != BL_CODE_NONE) q = bl_code_hi(mc) == k ? bl_code_lo(mc) : bl_code_hi(mc);
            if (q >= 0) {
                M.d(F_QP0X, kl) = S.p0x[q]; M.d(F_QP0Y, kl) = S.p0y[q];
                M.d(F_QVX, kl) = S.vx[q];   M.d(F_QVY, kl) = S.vy[q];
                M.d(F_QT0, kl) = S.t0[q];   M.d(F_QMASS, kl) = S.mass[q];
                M.i(I_QRC, kl) = S.rclass[q];
            } else {
                M.d(F_QP0X, kl) = 0; M.d(F_QP0Y, kl) = 0; M.d(F_QVX, kl) = 0; M.d(F_QVY, kl) = 0;
                M.d(F_QT0, kl) = 0; M.d(F_QMASS, kl) = 0; M.i(I_QRC, kl) = 0;
            }
        }
    }
    if (tid == 0) *M.n_dirty = 0;
    __syncthreads();
    cluster_sync_all();  // every CTA's shared memory is live before anybody writes into an inbox

    long long n_bb = 0, n_bo = 0, n_logged = 0, n_rescans = 0;
    double now = P.time;
    int mode = P.first_kind == BL_BALL_BALL ? MODE_BB : (P.first_kind == BL_BALL_OBSTACLE ? MODE_BO : MODE_BOTH);
    int parity = 0;
    int final_stage = P.query_only ? 1 : 0;  // 0: event loop, 1: report next ball-ball, 2: report next ball-obstacle
    int stop = BL_STOP_TIME, status = BL_OK;
    long long err_i = -1, err_j = -1;
    double bb_t = BL_INF, bo_t = BL_INF, bo_arg = 0.0;
    long long bb_i = -1, bb_j = 0, bo_ball = -1, bo_obs = -1;
    if (final_stage == 1) mode = MODE_BB;

    long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (;;) {
        const long long c0 = clock64();
        prof[5]++;
        // ---- argmin, thread level: best candidate among the owned balls ----------------------------
        uint64_t best_ts = ~0ull, best_code = BL_CODE_END;
        int best_kl = 0;
        for (int u = 0; u < bpt; u++) {
            const int kl = u * nthreads + tid, k = kbase + kl;
            if (k >= n) continue;
            if (mode != MODE_BO) {
                const uint64_t ts = bl_sortable(M.d(F_MT, kl)), cd = M.mcode[kl];
                if (bl_key_less(ts, cd, best_ts, best_code)) { best_ts = ts; best_code = cd; best_kl = kl; }
            }
            if (mode != MODE_BB) {
                const uint64_t ts = bl_sortable(M.d(F_OT, kl)), cd = BL_CODE_OBS | (uint64_t)k;
                if (bl_key_less(ts, cd, best_ts, best_code)) { best_ts = ts; best_code = cd; best_kl = kl; }
            }
        }
        // ---- warp level ---------------------------------------------------------------------------
        {
            const int w = warp_argmin(best_ts, best_code);
            if (lane == w) {
                M.slots[warp * 4 + 0] = best_ts; M.slots[warp * 4 + 1] = best_code;
                M.slots[warp * 4 + 2] = (uint64_t)best_kl;
            }
        }
        __syncthreads();
        // ---- CTA level: warp 0 picks the CTA winner and posts its payload into every CTA's inbox -------
        if (warp == 0) {
            uint64_t ts = ~0ull, cd = BL_CODE_END;
            if (lane < nwarps) { ts = M.slots[lane * 4 + 0]; cd = M.slots[lane * 4 + 1]; }
            const int w = warp_argmin(ts, cd);
            const int kl = (int)M.slots[w * 4 + 2];
            uint64_t word = 0;
            if (lane == W_TS) word = M.slots[w * 4 + 0];
            else if (lane == W_CODE) word = M.slots[w * 4 + 1];
            else if (lane < W_INT0) word = (uint64_t)__double_as_longlong(M.fd[(lane - W_K0) * cap + kl]);
            else if (lane == W_INT0) word = ((uint64_t)(unsigned)(kbase + kl) << 32) | (unsigned)M.i(I_RC, kl);
            else if (lane == W_INT1) word = ((uint64_t)(unsigned)M.i(I_QRC, kl) << 32) | (unsigned)M.i(I_OIDX, kl);
            else if (lane == W_OARG) word = (uint64_t)__double_as_longlong(M.d(F_OARG, kl));
            if (lane < NW) {
                const uint32_t local = smem_u32(&M.inbox[(parity * 16 + rank) * NW_PAD + lane]);
                for (int c = 0; c < C; c++) st_cluster_u64(mapa(local, (uint32_t)c), word);
            }
        }
        cluster_sync_all();
        const long long c1 = clock64();
        prof[0] += c1 - c0;
        // ---- cluster level + resolve: warp 0 of every CTA redundantly --------------------------------
        if (warp == 0) {
            const uint64_t *box = &M.inbox[parity * 16 * NW_PAD];
            uint64_t ts = ~0ull, cd = BL_CODE_END;
            if (lane < C) { ts = box[lane * NW_PAD + W_TS]; cd = box[lane * NW_PAD + W_CODE]; }
            const int cw = warp_argmin(ts, cd);
            const uint64_t *w = &box[cw * NW_PAD];
            ts = w[W_TS]; cd = w[W_CODE];
            const double t = bl_unsortable(ts);
            int kind = EV_STOP, st_local = BL_OK, stop_local = BL_STOP_TIME;
            const int k = (int)(w[W_INT0] >> 32), krc = (int)(w[W_INT0] & 0xffffffffu);
            const int qrc = (int)(w[W_INT1] >> 32), oidx = (int)(w[W_INT1] & 0xffffffffu);
            const long long events = n_bb + n_bo;
            if (cd == BL_CODE_STALE) {
                kind = EV_RESCAN;
            } else if (final_stage) {
                kind = EV_STOP;
            } else if (cd == BL_CODE_NONE || cd == BL_CODE_END) {
                kind = EV_STOP;
                stop_local = (P.max_events >= 0 && events >= P.max_events) ? BL_STOP_EVENTS : BL_STOP_TIME;
            } else if (ts == BL_TS_NAN) {
                kind = EV_STOP; st_local = BL_ENAN; stop_local = BL_STOP_ERROR;
            } else if (P.max_events >= 0 && events >= P.max_events) {
                kind = EV_STOP; stop_local = BL_STOP_EVENTS;  // checked first: the clock stays at the last collision
            } else if (!(t <= P.end_time) || t == BL_INF) {
                kind = EV_STOP; stop_local = BL_STOP_TIME;
            } else if (P.log.cap > 0 && n_logged >= P.log.cap) {
                kind = EV_STOP; stop_local = BL_STOP_LOG;
            } else if (cd & BL_CODE_OBS) {
                // ---- ball-obstacle, simulation.py:504-519 + :604-636 ----
                kind = EV_BO;
                const double p0x = __longlong_as_double(w[W_K0 + 0]), p0y = __longlong_as_double(w[W_K0 + 1]);
                const double vx = __longlong_as_double(w[W_K0 + 2]), vy = __longlong_as_double(w[W_K0 + 3]);
                const double t0 = __longlong_as_double(w[W_K0 + 4]);
                const double arg = __longlong_as_double(w[W_OARG]);
                const double px = bl_advance(p0x, vx, t, t0), py = bl_advance(p0y, vy, t, t0);
                double wx = 0, wy = 0;
                st_local = bl_obstacle_bounce<FMA>(M.obs[oidx], px, py, vx, vy, arg, wx, wy);
                if (st_local != BL_OK) { kind = EV_STOP; stop_local = BL_STOP_ERROR; }
                if (lane == 0) {
                    bc.a = k; bc.b = -1; bc.rca = krc; bc.rcb = 0; bc.obs = oidx;
                    bc.ax = px; bc.ay = py; bc.avx = wx; bc.avy = wy; bc.amass = __longlong_as_double(w[W_K0 + 5]);
                    bc.q_arg = arg;
                    if (kind == EV_BO && rank == 0 && P.log.cap > 0) {
                        const long long e = n_logged;
                        P.log.t[e] = t; P.log.i[e] = k; P.log.j[e] = -1 - (long long)oidx;
                        if (P.log.payload) {
                            double *pl = P.log.payload + 12 * e;
                            pl[0] = px; pl[1] = py; pl[2] = vx; pl[3] = vy; pl[4] = wx; pl[5] = wy; pl[6] = arg;
                            pl[7] = 0; pl[8] = 0; pl[9] = 0; pl[10] = 0; pl[11] = 0;
                        }
                    }
                }
            } else {
                // ---- ball-ball, simulation.py:446-455 + :562-602 ----
                kind = EV_BB;
                const int hi = bl_code_hi(cd), lo = bl_code_lo(cd);
                // the publisher is ball k, its cached partner is the other one; a = smaller index (idx1 < idx2)
                const bool k_is_a = (k == lo);
                const uint64_t *wa = k_is_a ? &w[W_K0] : &w[W_Q0];
                const uint64_t *wb = k_is_a ? &w[W_Q0] : &w[W_K0];
                const double a_p0x = __longlong_as_double(wa[0]), a_p0y = __longlong_as_double(wa[1]);
                const double a_vx = __longlong_as_double(wa[2]), a_vy = __longlong_as_double(wa[3]);
                const double a_t0 = __longlong_as_double(wa[4]), a_m = __longlong_as_double(wa[5]);
                const double b_p0x = __longlong_as_double(wb[0]), b_p0y = __longlong_as_double(wb[1]);
                const double b_vx = __longlong_as_double(wb[2]), b_vy = __longlong_as_double(wb[3]);
                const double b_t0 = __longlong_as_double(wb[4]), b_m = __longlong_as_double(wb[5]);
                const double ax = bl_advance(a_p0x, a_vx, t, a_t0), ay = bl_advance(a_p0y, a_vy, t, a_t0);
                const double bx = bl_advance(b_p0x, b_vx, t, b_t0), by = bl_advance(b_p0y, b_vy, t, b_t0);
                double m1 = a_m, m2 = b_m;
                bl_remap_masses(m1, m2);
                double w1x = 0, w1y = 0, w2x = 0, w2y = 0;
                st_local = bl_elastic<FMA>(ax, ay, a_vx, a_vy, m1, bx, by, b_vx, b_vy, m2, w1x, w1y, w2x, w2y);
                if (st_local != BL_OK) { kind = EV_STOP; stop_local = BL_STOP_ERROR; }
                if (lane == 0) {
                    bc.a = lo; bc.b = hi;
                    bc.rca = k_is_a ? krc : qrc; bc.rcb = k_is_a ? qrc : krc; bc.obs = -1;
                    bc.ax = ax; bc.ay = ay; bc.avx = w1x; bc.avy = w1y; bc.amass = a_m;
                    bc.bx = bx; bc.by = by; bc.bvx = w2x; bc.bvy = w2y; bc.bmass = b_m;
                    if (kind == EV_BB && rank == 0 && P.log.cap > 0) {
                        const long long e = n_logged;
                        P.log.t[e] = t; P.log.i[e] = lo; P.log.j[e] = hi;
                        if (P.log.payload) {
                            double *pl = P.log.payload + 12 * e;
                            pl[0] = ax; pl[1] = ay; pl[2] = a_vx; pl[3] = a_vy; pl[4] = w1x; pl[5] = w1y;
                            pl[6] = bx; pl[7] = by; pl[8] = b_vx; pl[9] = b_vy; pl[10] = w2x; pl[11] = w2y;
                        }
                    }
                }
            }
            if (lane == 0) {
                bc.t = t; bc.kind = kind; bc.status = st_local; bc.stop = stop_local; bc.ts = ts; bc.code = cd;
                bc.k = k;
                if (kind == EV_STOP && (cd & BL_CODE_OBS) && cd != BL_CODE_END) {
                    bc.obs = oidx; bc.q_arg = __longlong_as_double(w[W_OARG]);
                }
            }
        }
        parity ^= 1;
        __syncthreads();
        const long long c2 = clock64();
        prof[1] += c2 - c1;

        const int kind = bc.kind;
        if (kind == EV_RESCAN) {
            // ---- repair every stale candidate of this CTA from the table (one warp per row) -------------
            for (int u = 0; u < bpt; u++) {
                const int kl = u * nthreads + tid;
                if (kbase + kl < n && M.mcode[kl] == BL_CODE_STALE) M.dirty[atomicAdd(M.n_dirty, 1)] = kl;
            }
            __syncthreads();
            const int nd = *M.n_dirty;
            for (int d = warp; d < nd; d += nwarps) {
                const int kl = M.dirty[d], k = kbase + kl;
                const double *row = S.toi + (size_t)k * pitch;
                uint64_t ts = ~0ull, cd = BL_CODE_END;
                for (int x = lane; x < n; x += 32) {
                    const double v = __ldcg(row + x);
                    if (x != k && v < BL_INF) {
                        const uint64_t s2 = bl_sortable(v), c2 = bl_pair_code(k, x);
                        if (bl_key_less(s2, c2, ts, cd)) { ts = s2; cd = c2; }
                    }
                }
                const int w = warp_argmin(ts, cd);
                ts = __shfl_sync(0xffffffffu, ts, w);
                cd = __shfl_sync(0xffffffffu, cd, w);
                if (lane == 0) {
                    if (cd == BL_CODE_END) {
                        M.d(F_MT, kl) = BL_INF; M.mcode[kl] = BL_CODE_NONE;
                    } else {
                        const int q = bl_code_hi(cd) == k ? bl_code_lo(cd) : bl_code_hi(cd);
                        M.d(F_MT, kl) = bl_unsortable(ts); M.mcode[kl] = cd;
                        M.d(F_QP0X, kl) = __ldcg(S.p0x + q); M.d(F_QP0Y, kl) = __ldcg(S.p0y + q);
                        M.d(F_QVX, kl) = __ldcg(S.vx + q);   M.d(F_QVY, kl) = __ldcg(S.vy + q);
                        M.d(F_QT0, kl) = __ldcg(S.t0 + q);   M.d(F_QMASS, kl) = S.mass[q];
                        M.i(I_QRC, kl) = S.rclass[q];
                    }
                }
            }
            n_rescans += nd;  // per-CTA count; summed over CTAs at the end
            __syncthreads();
            if (tid == 0) *M.n_dirty = 0;
            prof[3] += clock64() - c2;
            prof[4]++;
            continue;
        }
        if (kind == EV_STOP) {
            if (final_stage == 0) {
                stop = bc.stop; status = bc.status;
                if (status != BL_OK) { err_i = bc.a; err_j = bc.b; now = bc.t; }  // the reference has moved to t
                if (stop == BL_STOP_TIME && P.end_time < BL_INF) now = P.end_time;  // _move(end_time), simulation.py:436
                final_stage = 1; mode = MODE_BB;
                continue;
            }
            if (final_stage == 1) {
                const uint64_t cd = bc.code;
                if (cd != BL_CODE_NONE && cd != BL_CODE_END && bc.t < BL_INF) {
                    bb_t = bc.t; bb_i = bl_code_lo(cd); bb_j = bl_code_hi(cd);
                }
                final_stage = 2; mode = MODE_BO;
                continue;
            }
            {
                const uint64_t cd = bc.code;
                if (cd != BL_CODE_END) {
                    // numpy argmin over _obstacles_toi: first ball attaining the minimum, even when it is +inf
                    bo_t = bc.t; bo_ball = (long long)(cd & ~BL_CODE_OBS); bo_obs = bc.obs; bo_arg = bc.q_arg;
                }
            }
            break;
        }

        // ---- apply the event: every owner re-solves its pair(s) with the colliding ball(s) ---------------
        const double t = bc.t;
        const int a = bc.a, b = bc.b;  // b == -1 for an obstacle event
        const double ax = bc.ax, ay = bc.ay, avx = bc.avx, avy = bc.avy;
        const double bx = bc.bx, by = bc.by, bvx = bc.bvx, bvy = bc.bvy;
        const int rca = bc.rca, rcb = bc.rcb;
        for (int u = 0; u < bpt; u++) {
            const int kl = u * nthreads + tid, k = kbase + kl;
            if (k >= n) continue;
            if (k == a || k == b) {
                const bool is_a = (k == a);
                const double px = is_a ? ax : bx, py = is_a ? ay : by, wx = is_a ? avx : bvx, wy = is_a ? avy : bvy;
                M.d(F_P0X, kl) = px; M.d(F_P0Y, kl) = py; M.d(F_VX, kl) = wx; M.d(F_VY, kl) = wy; M.d(F_T0, kl) = t;
                M.d(F_MT, kl) = BL_INF; M.mcode[kl] = BL_CODE_NONE;
                // keep the global mirror current: stale-row repairs read partner states from it
                __stcg(S.p0x + k, px); __stcg(S.p0y + k, py); __stcg(S.vx + k, wx); __stcg(S.vy + k, wy);
                __stcg(S.t0 + k, t);
                if (is_a && b >= 0) {  // toi_table[idx2][idx1] = INF, simulation.py:458
                    __stcg(S.toi + (size_t)a * pitch + b, BL_INF);
                    __stcg(S.toi + (size_t)b * pitch + a, BL_INF);
                }
                double ot, oarg; int oi;
                bl_next_obstacle<FMA>(M.obs, n_obs, S.rsum2_obs, K, M.i(I_RC, kl), px, py, wx, wy, M.d(F_RADIUS, kl), t,
                                      ot, oi, oarg);
                M.d(F_OT, kl) = ot; M.i(I_OIDX, kl) = oi; M.d(F_OARG, kl) = oarg;
                continue;
            }
            const double vx = M.d(F_VX, kl), vy = M.d(F_VY, kl), t0 = M.d(F_T0, kl);
            const double px = bl_advance(M.d(F_P0X, kl), vx, t, t0), py = bl_advance(M.d(F_P0Y, kl), vy, t, t0);
            const int rck = M.i(I_RC, kl);
            const double na = bl_add(t, bl_toi_ball_ball<FMA>(ax, ay, avx, avy, px, py, vx, vy,
                                                              __ldg(S.rsum2 + (size_t)rca * K + rck)));
            __stcg(S.toi + (size_t)a * pitch + k, na);
            __stcg(S.toi + (size_t)k * pitch + a, na);
            uint64_t c_ts = bl_sortable(na), c_code = na < BL_INF ? bl_pair_code(k, a) : BL_CODE_NONE;
            bool c_is_a = true;
            if (b >= 0) {
                const double nb = bl_add(t, bl_toi_ball_ball<FMA>(bx, by, bvx, bvy, px, py, vx, vy,
                                                                  __ldg(S.rsum2 + (size_t)rcb * K + rck)));
                __stcg(S.toi + (size_t)b * pitch + k, nb);
                __stcg(S.toi + (size_t)k * pitch + b, nb);
                const uint64_t b_ts = bl_sortable(nb), b_code = nb < BL_INF ? bl_pair_code(k, b) : BL_CODE_NONE;
                if (bl_key_less(b_ts, b_code, c_ts, c_code)) { c_ts = b_ts; c_code = b_code; c_is_a = false; }
            }
            // fold the new value(s) into the candidate of ball k
            const uint64_t o_code = M.mcode[kl], o_ts = bl_sortable(M.d(F_MT, kl));
            bool take = false;
            if (o_code != BL_CODE_STALE && o_code != BL_CODE_NONE) {
                const int q = bl_code_hi(o_code) == k ? bl_code_lo(o_code) : bl_code_hi(o_code);
                if (q == a || q == b) {
                    // the entry that was the candidate has just been overwritten
                    if (!bl_key_less(o_ts, o_code, c_ts, c_code)) take = true;  // new <= old: still the minimum
                    else M.mcode[kl] = BL_CODE_STALE;                               // old time stays as a lower bound
                } else {
                    take = bl_key_less(c_ts, c_code, o_ts, o_code);
                }
            } else {
                take = bl_key_less(c_ts, c_code, o_ts, o_code);  // stale: only a strictly earlier time is decisive
            }
            if (take && c_code != BL_CODE_NONE) {
                M.d(F_MT, kl) = bl_unsortable(c_ts); M.mcode[kl] = c_code;
                M.d(F_QP0X, kl) = c_is_a ? ax : bx; M.d(F_QP0Y, kl) = c_is_a ? ay : by;
                M.d(F_QVX, kl) = c_is_a ? avx : bvx; M.d(F_QVY, kl) = c_is_a ? avy : bvy;
                M.d(F_QT0, kl) = t; M.d(F_QMASS, kl) = c_is_a ? bc.amass : bc.bmass;
                M.i(I_QRC, kl) = c_is_a ? rca : rcb;
            }
        }
        now = t;
        if (kind == EV_BB) n_bb++; else n_bo++;
        if (P.log.cap > 0) n_logged++;
        mode = MODE_BOTH;
        prof[2] += clock64() - c2;
        // no barrier needed here: the next writer of `bc` is warp 0 after the next __syncthreads pair
    }

    // ---- write back the resident cache ------------------------------------------------------------------
    for (int u = 0; u < bpt; u++) {
        const int kl = u * nthreads + tid, k = kbase + kl;
        if (k < n) {
            S.cover_t[k] = M.d(F_MT, kl); S.cover_code[k] = M.mcode[kl];
            S.obs_t[k] = M.d(F_OT, kl); S.obs_idx[k] = M.i(I_OIDX, kl); S.obs_arg[k] = M.d(F_OARG, kl);
        }
    }
    if (tid == 0 && n_rescans) atomicAdd((unsigned long long *)&P.result->n_rescans, (unsigned long long)n_rescans);
    if (rank == 0 && tid == 0) {
        bl_result *r = P.result;
        r->time = now; r->n_ball_ball = n_bb; r->n_ball_obstacle = n_bo;
        r->bb_t = bb_t; r->bb_i = bb_i; r->bb_j = bb_j;
        r->bo_t = bo_t; r->bo_ball = bo_ball; r->bo_obs = bo_obs; r->bo_arg = bo_arg;
        r->status = status; r->stop = stop; r->n_logged = n_logged; r->err_i = err_i; r->err_j = err_j;
        for (int q = 0; q < 8; q++) r->prof[q] = prof[q];
    }
    cluster_sync_all();  // nobody exits while a peer may still post into its inbox
}

}  // namespace

// ---------------------------------------------------------------------------------------------
int bl_evolve_pick_config(const bl_handle *h, int n, int *cluster, int *threads, int *bpt) {
    if (n <= 0) { *cluster = 1; *threads = 32; *bpt = 1; return BL_OK; }
    int c = 1, t = ((n + 31) / 32) * 32, b = 1;
    if (t > BL_MAX_THREADS) {
        t = BL_MAX_THREADS;
        c = (n + BL_MAX_THREADS - 1) / BL_MAX_THREADS;
        int p = 1;
        while (p < c) p <<= 1;
        c = p;
        if (c > h->max_cluster) {
            c = h->max_cluster;
            b = (n + c * t - 1) / (c * t);
        }
    }
    *cluster = c; *threads = t; *bpt = b;
    size_t need = smem_bytes(t * b, 8);
    if (need > h->max_smem_optin) return BL_EUNSUPPORTED;
    return BL_OK;
}

template <bool FMA>
static int launch_t(bl_handle *h, const Params &P, int cluster, int threads, size_t smem, cudaStream_t st) {
    auto kern = bl_evolve_kernel<FMA>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return bl_check_cuda(h, e, "cudaFuncSetAttribute(smem)");
    if (cluster > 8) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return bl_check_cuda(h, e, "cudaFuncSetAttribute(non-portable cluster)");
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)cluster, 1, 1);
    cfg.blockDim = dim3((unsigned)threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, P);
    return bl_check_cuda(h, e, "launch bl_evolve_kernel");
}

int bl_launch_evolve(bl_handle *h, const bl_system *s, double time, double end_time, long long max_events,
                     int first_kind, const bl_log *log, int query_only, cudaStream_t st) {
    int cluster, threads, bpt;
    int rc = bl_evolve_pick_config(h, s->n, &cluster, &threads, &bpt);
    if (rc != BL_OK)
        return bl_set_error(h, rc, "event kernel keeps ball state resident in shared memory: n=%d does not fit "
                            "%d CTAs x %d threads", s->n, cluster, threads);
    Params P;
    P.s = *s;
    P.time = time; P.end_time = end_time; P.max_events = max_events; P.first_kind = first_kind;
    P.query_only = query_only;
    if (log) P.log = *log; else { P.log.cap = 0; P.log.t = nullptr; P.log.i = nullptr; P.log.j = nullptr; P.log.payload = nullptr; }
    P.result = h->d_result;
    P.cap = threads * bpt;
    P.bpt = bpt;
    size_t smem = smem_bytes(P.cap, s->n_obstacles);
    if (smem > h->max_smem_optin)
        return bl_set_error(h, BL_EUNSUPPORTED, "event kernel needs %zu B of shared memory per CTA (limit %zu)", smem,
                            h->max_smem_optin);
    cudaError_t e = cudaMemsetAsync(h->d_result, 0, sizeof(bl_result), st);
    if (e != cudaSuccess) return bl_check_cuda(h, e, "memset result");
    h->launches++;
    return h->dot_fma ? launch_t<true>(h, P, cluster, threads, smem, st) : launch_t<false>(h, P, cluster, threads, smem, st);
}
