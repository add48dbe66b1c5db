// bl_evolve.cu -- the event loop of Billiard.evolve as ONE persistent, grid-wide kernel that commits BATCHES of
// causally independent collisions per synchronisation round, in exactly the reference's order and with exactly
// the reference's bits.
//
// Replaces simulation.py:354-438 (evolve), :440-502 (bounce_ball_ball), :504-554 (bounce_ball_obstacle),
// :556-560 (_move), :562-636 (_resolve_*), with physics.py:10-76/:255-282 and obstacles.py:70-76/:112-144
// inlined from bl_arith.cuh.
//
// Why batches.  The reference handles one collision per iteration: global argmin -> resolve -> 2(N-2) pair
// solves -> argmin ...  One iteration is a chain of dependent steps (reduction, broadcast, solve), so a kernel
// that follows it literally is bound by synchronisation latency, microseconds per event whatever the bandwidth.
// But in a system of N balls the next few dozen collisions almost always involve different balls and do not
// influence each other: collision e2 is "the next one after e1" as long as nothing e1 produces (the new impact
// times of its two balls) comes earlier than e2.  That condition can be CHECKED exactly, so a round works on a
// whole window of pending collisions at once:
//
//   1. post      every ball whose candidate key (time, code) lies below the window end `theta` is posted, together
//                with the violation word of the previous batch (step 5); ONE grid barrier later every CTA holds the
//                same complete list of keys < theta;
//   2. validate  the previous batch was committed optimistically; if its violation word says that nothing it
//                produced came too early (the normal case) only its table rows and log records are left to write,
//                otherwise the commit is rolled back, redone for the valid prefix, and the keys are posted again;
//   3. analyse   all threads rank the keys pairwise, warp 0 sorts them into its lanes and cuts the list at the
//                first entry that shares a ball with an earlier one (or is only a lower bound); one lane per
//                remaining collision resolves it (elastic_collision / obstacle reflection);
//   4. solve     every ball is solved against every ball of the batch -- with its state as of just before that
//                collision -- and the values are parked in shared memory; any new key that would come earlier
//                than a collision of the batch is a violation;
//   5. commit    optimistically, as if nothing violated: candidates folded, states / mirrors / sequence numbers
//                updated, so that the next keys can be posted right away.
//
// Collision 1 of a round is the true global minimum, so every round commits at least what the reference's
// iteration would; the committed sequence, every table entry and every state bit are the reference's.
//
// Data placement: a CTA owns Tb consecutive balls; their state of record (p0, v, t0), candidate keys and
// next-obstacle entries live in its shared memory for the whole launch (mirrored to global memory when they
// change, so that resolvers can read any ball).  H threads serve each ball (thread = helper h of ball b): the
// per-ball bookkeeping is done by helper 0, the pair solves, table stores and repair loads of a round are split
// between the helpers -- with one warp per scheduler the FP64 dependency chains would otherwise sit exposed.
// The table of absolute impact times stays in HBM and a collision only WRITES it: the whole row of each
// colliding ball, coalesced.  The mirrored cells (column of the colliding ball) are left outdated -- a per-ball
// sequence number says which of toi[i][j] / toi[j][i] is current (the row of the ball whose collision came
// later) and bl_symmetrize makes both current again before anything on the host side reads the table.  The
// kernel reads the table only to repair stale candidates.
// Grid: ceil(N/Tb) CTAs, one per SM, cooperative launch with a global-memory barrier (64 balls x 8 threads up to 9472
// balls, 128 x 4 above; launching <= 16 CTAs as one thread-block cluster with the hardware barrier is kept behind
// BL_CLUSTER_MAX -- the 64 x 8 grid shape is faster at every size).
//
// "Cover" instead of row minima: ball k's candidate m[k] is the minimum over a SUBSET of row k that
// always contains every pair whose value was last written at an event of the OTHER ball.  After an event of ball
// a every pair (a,k) is re-solved and folded into m[k], so m[a] is simply emptied.  If k's candidate was its
// pair with a and the new value is larger, m[k] keeps the old time as a lower bound ("stale"); stale candidates
// are made exact again only when they come near the front of the queue.  The global minimum over {m[k]} equals
// the minimum over the whole table.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "bl_arith.cuh"
#include "bl_internal.h"

namespace {

enum { MODE_WINDOW = 0, MODE_ARGMIN = 1 };
enum { KIND_BOTH = 0, KIND_BB = 1, KIND_BO = 2 };
enum { SYNC_CTA = 0, SYNC_CLUSTER = 1, SYNC_GRID = 2 };
enum { F_DUP = 1, F_CONFLICT = 2 };
enum { ACT_BATCH = 0, ACT_OVERFLOW, ACT_EMPTY, ACT_REPAIR, ACT_STOP, ACT_ERROR };
// a resolved key, by collection slot: QS_D doubles (the log payload: position, old and new velocity of both balls, time;
// an odd stride keeps the lanes of a warp on different banks) and QS_I ints (status, obstacle, is ball-ball, radius classes)
enum { QS_D = 13, QS_I = 8 };

struct Params {
    bl_system s;
    double time, end_time;
    long long max_events;
    int first_kind;
    int query_only;
    bl_log log;
    bl_result *result;
    int Tb;         // balls per CTA
    int H;          // threads per ball
    int ncta;
    int m_max;      // collisions per round the stash has room for
    int target;     // collisions per round the window aims at
    int sync_mode;
    double window;  // initial window length (0: unknown)
    unsigned *bar;  // [0] grid barrier counter, [4] stale-list length, [8] readers' gate (zero at launch)
    uint64_t *post; // [2][BL_MAX_CTAS][BL_POST_WORDS]
    uint64_t *stale_list; // [BL_MAX_CTAS * BL_MAX_THREADS][2] (ball, seq) of the candidates under repair
    uint64_t *partial;    // [2][BL_MAX_CTAS][BL_RCHUNK][2] per-CTA minima of a repair chunk
    double repair_horizon; // stale candidates within this many windows of `now` are repaired together
    double wc[5];          // window control: factor after a violation; growth factor, below this fraction of the target; shrink factor, above this fraction
};

// ---- shared memory map (8-byte words) -----------------------------------------------------------------
struct Layout {
    int stash;                        // [2*m_max][Tb] doubles (repair scratch while no batch is in flight)
    int bx, by, bvx, bvy, bt;         // [2*m_max] batch balls: state right after their collision
    int bid, brc;                     // [2*m_max] ints: ball index (-1: none), radius class
    int e_ts, e_code;                 // [BL_EV_CAP] collected keys (unsorted)
    int e_ab;                         // [BL_EV_CAP] their balls, two 32-bit words (key_words)
    int s_ts, s_code;                 // [BL_SORT_CAP] sorted keys
    int s_flag, a_rank, a_flag;       // [BL_SORT_CAP] ints: flags in sorted order; rank and flags by collection slot
    int r_pay;                        // [m_max][13] log payload per rank
    int r_info;                       // [m_max][4] ints: status, obstacle index, is ball-ball, pad
    int pl_ts, pl_code;               // [BL_R_POST] this CTA's posted keys
    int w_ts, w_code;                 // [32] warp slots of the exact argmin
    int p0x, p0y, vx, vy, t0;         // [Tb] state of record of my balls
    int mts, mcode, ots;              // [Tb] candidate key, next-obstacle time (sortable)
    int sts, scode, lts, lcode;       // [Tb] second candidate; key that bounds the rest of the cover from below
    int seq;                          // [Tb] sequence number of the ball's last collision
    int radius, oarg;                 // [Tb] doubles
    int oidx, brank, rclass;          // [Tb] ints (brank: [2][Tb])
    int undo;                         // [15][Tb] state before the optimistic commit of the last batch
    int bts;                          // [2][m_max] sortable collision times of the batch, by rank
    int obsp;                         // [H][3][Tb] next-obstacle partial minima of the helpers (t, arg, index)
    int hl;                           // ints: [Tb] count, [Tb] partner-hit flag, [Tb][BL_HL_CAP] finite new values of a ball
    int dirty;                        // [Tb] ints
    int scratch;                      // [3*Tb] repair results (ts, partner) / new obstacle entry (ts, arg, idx)
    int wq;                           // [T/32][4*BL_WQ_CAP] per-warp queue of pairs that reach the sqrt; between the solves:
                                      // [BL_SORT_CAP][QS_D] doubles + [BL_SORT_CAP][QS_I] ints, the collected keys resolved
    int obs;                          // BL_OBS_WORDS * n_obs
    int misc;                         // 32 words
    int prof;                         // 16 words: phase clocks of this CTA (kept out of the registers)
    int words;
};
__host__ __device__ inline Layout make_layout(int Tb, int H, int m_max, int n_obs) {
    Layout L;
    int w = 0;
    L.stash = w; w += 2 * m_max * Tb;
    // batch tables: two copies, the previous round's batch is finalised while the next one is being resolved
    L.bx = w; w += 4 * m_max; L.by = w; w += 4 * m_max; L.bvx = w; w += 4 * m_max; L.bvy = w; w += 4 * m_max;
    L.bt = w; w += 4 * m_max;
    L.bid = w; w += 2 * m_max; L.brc = w; w += 2 * m_max;
    L.e_ts = w; w += BL_EV_CAP; L.e_code = w; w += BL_EV_CAP; L.e_ab = w; w += BL_EV_CAP;
    L.s_ts = w; w += BL_SORT_CAP; L.s_code = w; w += BL_SORT_CAP;
    L.s_flag = w; w += BL_SORT_CAP / 2; L.a_rank = w; w += BL_SORT_CAP / 2; L.a_flag = w; w += BL_SORT_CAP / 2;
    L.r_pay = w; w += 2 * 13 * m_max;
    L.r_info = w; w += 2 * 2 * m_max;
    L.pl_ts = w; w += BL_R_POST; L.pl_code = w; w += BL_R_POST;
    L.w_ts = w; w += 32; L.w_code = w; w += 32;
    L.p0x = w; w += Tb; L.p0y = w; w += Tb; L.vx = w; w += Tb; L.vy = w; w += Tb; L.t0 = w; w += Tb;
    L.mts = w; w += Tb; L.mcode = w; w += Tb; L.ots = w; w += Tb;
    L.sts = w; w += Tb; L.scode = w; w += Tb; L.lts = w; w += Tb; L.lcode = w; w += Tb;
    L.seq = w; w += Tb;
    L.radius = w; w += Tb; L.oarg = w; w += Tb;
    L.oidx = w; w += (Tb + 1) / 2; L.brank = w; w += Tb; L.rclass = w; w += (Tb + 1) / 2;
    L.undo = w; w += 15 * Tb;
    L.bts = w; w += 2 * m_max;
    L.obsp = w; w += 3 * H * Tb;
    L.hl = w; w += (Tb * (BL_HL_CAP + 2) + 1) / 2;
    L.dirty = w; w += (Tb + 1) / 2;
    L.scratch = w; w += 3 * Tb;
    {
        const int q_solve = ((Tb * H) / 32) * 4 * BL_WQ_CAP, q_slots = BL_SORT_CAP * QS_D + BL_SORT_CAP * QS_I / 2;
        L.wq = w; w += q_solve > q_slots ? q_solve : q_slots;
    }
    L.obs = w; w += BL_OBS_WORDS * (n_obs > 0 ? n_obs : 1);
    L.misc = w; w += 32;
    L.prof = w; w += 16;
    L.words = w;
    return L;
}
// ints of `misc`; M_VIOL is a 64-bit word index
enum { M_PCOUNT = 0, M_ECOUNT, M_NDIRTY, M_SBASE, M_ACTION, M_NF, M_NE, M_STOPKIND, M_STATUS };
enum { M_VIOL = 8, M_GMIN = 9 };

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// all CTAs of the launch; global-memory writes made before it are visible to ld.cg after it
__device__ __forceinline__ void sync_all(const Params &P, unsigned &epoch) {
    if (P.sync_mode == SYNC_CTA) { __syncthreads(); return; }
    if (P.sync_mode == SYNC_CLUSTER) { cluster_sync_all(); return; }
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += (unsigned)P.ncta;
        __threadfence();
        atomicAdd(P.bar, 1u);
        while ((int)(ld_acquire_u32(P.bar) - epoch) < 0) {}
    }
    __syncthreads();
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
// minimum of a 64-bit value over the warp (two REDUX)
__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
    const unsigned FULL = 0xffffffffu;
    const unsigned hi = __reduce_min_sync(FULL, (unsigned)(v >> 32));
    const unsigned lo = __reduce_min_sync(FULL, (unsigned)(v >> 32) == hi ? (unsigned)v : 0xffffffffu);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }

__device__ __forceinline__ double as_d(uint64_t u) { return __longlong_as_double((long long)u); }
__device__ __forceinline__ uint64_t as_u(double d) { return (uint64_t)__double_as_longlong(d); }
__device__ __forceinline__ uint64_t pack2(int lo, int hi) { return (uint64_t)(unsigned)lo | ((uint64_t)(unsigned)hi << 32); }
__device__ __forceinline__ int lo32(uint64_t u) { return (int)(u & 0xffffffffu); }
__device__ __forceinline__ int hi32(uint64_t u) { return (int)(u >> 32); }

// balls of a key: pair -> (lo, hi); obstacle hit -> (ball, -1); stale / none -> (-1, -1)
__device__ __forceinline__ void key_balls(uint64_t code, int &a, int &b) {
    if (code == BL_CODE_STALE || code == BL_CODE_NONE || code == BL_CODE_END) { a = -1; b = -1; }
    else if (code & BL_CODE_OBS) { a = (int)(code & 0x7fffffffull); b = -1; }
    else { a = bl_code_lo(code); b = bl_code_hi(code); }
}

// the balls of a key as two 32-bit words for equality tests; absent balls (obstacle hit: the second, stale / none: both)
// become a word no ball index equals
#define BL_NO_BALL 0xffffffffu
__device__ __forceinline__ uint64_t key_words(uint64_t code) {
    int a, b;
    key_balls(code, a, b);
    return (uint64_t)(a >= 0 ? (unsigned)a : BL_NO_BALL) | ((uint64_t)(b >= 0 ? (unsigned)b : BL_NO_BALL) << 32);
}

// all ones if (tj, cj, j) < (ti, ci, i) lexicographically, else 0: the borrow of a 160-bit subtraction, branch-free
__device__ __forceinline__ unsigned key_before_mask(uint64_t tj, uint64_t cj, unsigned j, uint64_t ti, uint64_t ci, unsigned i) {
    unsigned r;
    asm("{\n\t.reg .u32 t;\n\t"
        "sub.cc.u32 t, %1, %2;\n\t"
        "subc.cc.u32 t, %3, %4;\n\t"
        "subc.cc.u32 t, %5, %6;\n\t"
        "subc.cc.u32 t, %7, %8;\n\t"
        "subc.cc.u32 t, %9, %10;\n\t"
        "subc.u32 %0, 0, 0;\n\t}"
        : "=r"(r)
        : "r"(j), "r"(i), "r"((unsigned)cj), "r"((unsigned)ci), "r"((unsigned)(cj >> 32)), "r"((unsigned)(ci >> 32)),
          "r"((unsigned)tj), "r"((unsigned)ti), "r"((unsigned)(tj >> 32)), "r"((unsigned)(ti >> 32)));
    return r;
}

// collisions per round the parked values have room for: 2 values of 8 bytes per collision and ball
__host__ __device__ constexpr int pick_m_max(int tb) {
    return (BL_STASH_BYTES / 16) / tb > BL_M_MAX ? BL_M_MAX : ((BL_STASH_BYTES / 16) / tb < 1 ? 1 : (BL_STASH_BYTES / 16) / tb);
}

// ---------------------------------------------------------------------------------------------
// TB / HH > 0: balls per CTA and threads per ball fixed at compile time (the shape of the large systems), which turns
// every shared-memory offset into an immediate; 0: taken from the launch parameters
// KONE: all balls have the same radius (one radius class): (r1 + r2)**2 is one constant
template <bool FMA, int MAXT, int TB, int HH, bool KONE = false>
__global__ void __launch_bounds__(MAXT, 1) bl_evolve_kernel(const __grid_constant__ Params P) {
    extern __shared__ __align__(16) uint64_t sm[];
    const bl_system &S = P.s;
    const int n = S.n, K = S.n_classes, n_obs = S.n_obstacles;
    const int Tb = TB ? TB : P.Tb, H = HH ? HH : P.H, T = Tb * H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int h = tid / Tb, bl = tid - h * Tb;  // helper h of local ball bl
    const bool primary = h == 0;
    const int cta = blockIdx.x, ncta = P.ncta, m_max = TB ? pick_m_max(TB) : P.m_max;
    const Layout L = make_layout(Tb, H, m_max, n_obs);
    double *stash = reinterpret_cast<double *>(sm + L.stash);
    double *const bx_b = reinterpret_cast<double *>(sm + L.bx), *const by_b = reinterpret_cast<double *>(sm + L.by);
    double *const bvx_b = reinterpret_cast<double *>(sm + L.bvx), *const bvy_b = reinterpret_cast<double *>(sm + L.bvy);
    double *const bt_b = reinterpret_cast<double *>(sm + L.bt);
    int *const bid_b = reinterpret_cast<int *>(sm + L.bid), *const brc_b = reinterpret_cast<int *>(sm + L.brc);
    uint64_t *e_ts = sm + L.e_ts, *e_code = sm + L.e_code, *s_ts = sm + L.s_ts, *s_code = sm + L.s_code;
    uint64_t *e_ab = sm + L.e_ab;
    int *s_flag = reinterpret_cast<int *>(sm + L.s_flag), *a_rank = reinterpret_cast<int *>(sm + L.a_rank);
    int *a_flag = reinterpret_cast<int *>(sm + L.a_flag);
    double *const r_pay_b = reinterpret_cast<double *>(sm + L.r_pay);
    int *const r_info_b = reinterpret_cast<int *>(sm + L.r_info);
    uint64_t *pl_ts = sm + L.pl_ts, *pl_code = sm + L.pl_code, *w_ts = sm + L.w_ts, *w_code = sm + L.w_code;
    double *st_p0x = reinterpret_cast<double *>(sm + L.p0x), *st_p0y = reinterpret_cast<double *>(sm + L.p0y);
    double *st_vx = reinterpret_cast<double *>(sm + L.vx), *st_vy = reinterpret_cast<double *>(sm + L.vy);
    double *st_t0 = reinterpret_cast<double *>(sm + L.t0);
    uint64_t *c_mts = sm + L.mts, *c_mcode = sm + L.mcode, *c_ots = sm + L.ots;
    uint64_t *c_sts = sm + L.sts, *c_scode = sm + L.scode, *c_lts = sm + L.lts, *c_lcode = sm + L.lcode;
    long long *c_seq = reinterpret_cast<long long *>(sm + L.seq);
    double *c_radius = reinterpret_cast<double *>(sm + L.radius);
    double *c_oarg = reinterpret_cast<double *>(sm + L.oarg);
    int *c_oidx = reinterpret_cast<int *>(sm + L.oidx), *const brank_b = reinterpret_cast<int *>(sm + L.brank);
    uint64_t *const undo = sm + L.undo;  // [f][Tb], f = 0 mts, 1 mcode, 2..6 p0x p0y vx vy t0, 7 ots, 8 oarg, 9 oidx, 10 seq, 11..14 second candidate and bound
    int *c_rclass = reinterpret_cast<int *>(sm + L.rclass);
    int *dirty = reinterpret_cast<int *>(sm + L.dirty);
    uint64_t *scratch = sm + L.scratch;
    double *wq = reinterpret_cast<double *>(sm + L.wq);
    const bl_obstacle *obs = reinterpret_cast<const bl_obstacle *>(sm + L.obs);
    int *misc = reinterpret_cast<int *>(sm + L.misc);
    uint64_t *const bts_b = sm + L.bts;
    uint64_t *const obsp = sm + L.obsp;
    int *const hl_cnt = reinterpret_cast<int *>(sm + L.hl), *const ph = hl_cnt + Tb, *const hl = ph + Tb;
    unsigned long long *m_viol = reinterpret_cast<unsigned long long *>(sm + L.misc + M_VIOL);
    unsigned long long *m_gmin = reinterpret_cast<unsigned long long *>(sm + L.misc + M_GMIN);

    const int base = cta * Tb;
    const int k = base + bl;  // the ball this thread serves
    const bool own = k < n;

    // ---- load the resident state --------------------------------------------------------------
    const uint64_t ts_inf = bl_sortable(BL_INF);
    for (int q = tid; q < BL_OBS_WORDS * n_obs; q += T) sm[L.obs + q] = reinterpret_cast<const uint64_t *>(S.obstacles)[q];
    const long long seq_base = *S.seq_counter;
    if (primary) {
        if (own) {
            st_p0x[bl] = S.p0x[k]; st_p0y[bl] = S.p0y[k]; st_vx[bl] = S.vx[k]; st_vy[bl] = S.vy[k]; st_t0[bl] = S.t0[k];
            c_seq[bl] = S.seq[k];
            c_rclass[bl] = S.rclass[k];
            c_radius[bl] = S.radius[k];
            c_oarg[bl] = S.obs_arg[k];
            c_oidx[bl] = S.obs_idx[k];
            c_ots[bl] = bl_sortable(S.obs_t[k]);
            // what the table kernels / the previous launch left: the row minimum (everything else is later), a lower
            // bound, or nothing
            const uint64_t ct = bl_sortable(S.cover_t[k]), cc = S.cover_code[k];
            c_mts[bl] = ct; c_mcode[bl] = cc;
            c_sts[bl] = ts_inf; c_scode[bl] = BL_CODE_NONE;
            c_lts[bl] = cc == BL_CODE_NONE ? ts_inf : ct;
            c_lcode[bl] = cc == BL_CODE_NONE ? BL_CODE_NONE : (cc == BL_CODE_STALE ? 0ull : cc + 1);
        } else {
            st_p0x[bl] = 0; st_p0y[bl] = 0; st_vx[bl] = 0; st_vy[bl] = 0; st_t0[bl] = 0;
            c_seq[bl] = 0; c_rclass[bl] = 0; c_radius[bl] = 0; c_oarg[bl] = 0; c_oidx[bl] = -1;
            c_ots[bl] = ts_inf; c_mts[bl] = ts_inf; c_mcode[bl] = BL_CODE_NONE;
            c_sts[bl] = ts_inf; c_scode[bl] = BL_CODE_NONE; c_lts[bl] = ts_inf; c_lcode[bl] = BL_CODE_NONE;
        }
        brank_b[bl] = -1; brank_b[Tb + bl] = -1;
        hl_cnt[bl] = 0; ph[bl] = 0;
    }
    const double rs11 = (K == 1 && n > 0) ? S.rsum2[0] : 0.0;
    double *const myrow = S.toi + (size_t)(own ? k : 0) * (size_t)S.pitch;
    if (tid < 32) sm[L.misc + tid] = 0;
    if (tid < BL_R_POST) { pl_ts[tid] = ~0ull; pl_code[tid] = BL_CODE_END; }
    if (tid < BL_SORT_CAP) { a_rank[tid] = 0; a_flag[tid] = 0; }
    __syncthreads();
    if (tid == 0) { *m_viol = ~0ull; *m_gmin = ~0ull; }
    const int myrc = c_rclass[bl];

    const bool prof_thread = warp == 0;  // the whole warp reads the clock: a single-lane branch would leave warp 0 diverged
    // phase clocks: the whole of warp 0 reads the clock, lane 0 adds to the counters in shared memory
    long long *const prof = reinterpret_cast<long long *>(sm + L.prof);
    const bool prof_lane = tid == 0;
    if (tid < 16) prof[tid] = 0;
    long long done = 0, n_bb = 0, n_rescans = 0;
    double now = P.time;   // Billiard.time: time of the last committed collision
    double prev_now = P.time;
    double win = P.window;
    // target 1: always the exact argmin, one collision per round -- the reference's schedule
    int mode = (P.first_kind != BL_ANY || P.query_only || !(win > 0.0) || P.target <= 1) ? MODE_ARGMIN : MODE_WINDOW;
    int kindmask = P.first_kind == BL_BALL_BALL ? KIND_BB : (P.first_kind == BL_BALL_OBSTACLE ? KIND_BO : KIND_BOTH);
    int final_stage = P.query_only ? 1 : 0;  // 0: event loop, 1: report next ball-ball, 2: report next ball-obstacle
    if (final_stage == 1) kindmask = KIND_BB;
    int shrinks = 0;
    unsigned par = 0, epoch = 0, reads_epoch = 0;
    const bool want_log = P.log.cap > 0;
    bl_result *const res = P.result;
    // exclusive upper bound of the keys that may be processed: t <= end_time and t < inf
    uint64_t end_excl = bl_sortable(P.end_time);
    end_excl = end_excl >= ts_inf ? ts_inf : end_excl + 1;
    const unsigned lt_mask = (1u << lane) - 1u;
    __syncthreads();
    if (P.sync_mode == SYNC_CLUSTER) cluster_sync_all();

    // ---- pieces of a commit; shared by the optimistic commit, the deferred row stores and the redo ----
    // Row stores of the first nc_ of the nf_ collisions of a batch (ball list Bbid, ranks of my CTA's balls in it Brank):
    // the parked values of my CTA's columns go to the rows of the colliding balls.  Which thread stores a cell does
    // not matter: warp sw of nsw keeps one 32-column tile and walks down the parked rows with several in flight.
    auto row_stores = [&](const int *Bbid, const int *Brank, const int nf_, const int nc_, const int sw, const int nsw) {
        const int segs = Tb >> 5;
        int seg0, segstep, j0, jstep;
        if (nsw >= segs) { jstep = nsw / segs; if (sw >= jstep * segs) return; seg0 = sw % segs; segstep = segs; j0 = sw / segs; }
        else { seg0 = sw; segstep = nsw; j0 = 0; jstep = 1; }
        for (int seg = seg0; seg < segs; seg += segstep) {
            const int c = (seg << 5) + lane, kc = base + c;
            if (kc >= n) continue;
            const int pr = Brank[c];
            const int rk = pr >= 0 && pr < nf_ ? pr : -1;  // rank of ball kc in that batch
            const int rk_c = rk < nc_ ? rk : -1;           // ... if its collision is committed
            double *const col = S.toi + kc;
            const double *const sv = stash + c;
#pragma unroll 4
            for (int j = j0; j < 2 * nc_; j += jstep) {
                const int x = Bbid[j];
                const int r = j >> 1;
                // r == rk: toi_table[idx2][idx1] = INF, simulation.py:458, the ball's cell in its partner's row (nothing
                // on the diagonal); r < rk_c: that pair is re-solved at kc's own, later collision, from x's side;
                // otherwise the cell in the row of the ball whose collision is the later one of the pair
                const double v = r == rk ? BL_INF : sv[j * Tb];
                if (x >= 0 && !(r == rk && x == kc) && !(r < rk_c)) __stcg(col + (size_t)x * (size_t)S.pitch, v);
            }
        }
    };
    // event log of the first nc_ collisions of a batch, simulation.py:590-594 / :626-631 payloads
    auto write_log = [&](const double *Br_pay, const int *Br_info, const int *Bbid, const long long first_event,
                         const int nc_, const int first, const int stride) {
        for (int r = first; r < nc_; r += stride) {
            const long long e = first_event + r;
            const double *pl = Br_pay + 13 * r;
            const bool bb = Br_info[4 * r + 2] != 0;
            P.log.t[e] = pl[12]; P.log.i[e] = Bbid[2 * r];
            P.log.j[e] = bb ? (long long)Bbid[2 * r + 1] : -1 - (long long)Br_info[4 * r + 1];
            if (P.log.payload) {
                double *o = P.log.payload + 12 * e;
#pragma unroll
                for (int f = 0; f < 12; f++) o[f] = pl[f];
            }
        }
    };
    // The cover of my ball after the first nc_ collisions of a batch.  Kept: the two smallest entries E1 < E2 (exact,
    // current table values) and a key L that bounds every other entry from below; E1 < L makes E1 THE minimum.  An
    // entry whose partner collided is dropped (its pair has a new value), new finite values below L are inserted,
    // the third smallest becomes the new L.  No exact entry left: "stale" -- L is all that is known until a repair.
    //   from_list: the finite new values are listed in hl (optimistic commit) / scan all parked values (redo)
    auto cover_fold = [&](const int *Bbid, const int nc_, const int myrank_, const bool mine_, const bool from_list,
                          int dropped /* bit 0: E1's partner collided, bit 1: E2's */) {
        // the three smallest keys seen, in registers (sentinel = larger than any key)
        uint64_t k1t = ~0ull, k1c = ~0ull, k2t = ~0ull, k2c = ~0ull, k3t = ~0ull, k3c = ~0ull;
        auto insert = [&](uint64_t t_, uint64_t c_) {
            if (!bl_key_less(t_, c_, k3t, k3c)) return;
            k3t = t_; k3c = c_;
            if (bl_key_less(k3t, k3c, k2t, k2c)) { const uint64_t tt = k3t, cc = k3c; k3t = k2t; k3c = k2c; k2t = tt; k2c = cc; }
            if (bl_key_less(k2t, k2c, k1t, k1c)) { const uint64_t tt = k2t, cc = k2c; k2t = k1t; k2c = k1c; k1t = tt; k1c = cc; }
        };
        uint64_t lt = c_lts[bl], lc = c_lcode[bl];
        const uint64_t e1c = c_mcode[bl], e2c = c_scode[bl];
        int q1 = -1, q2 = -1;
        if (mine_) {
            lt = ts_inf; lc = BL_CODE_NONE;  // my cover is emptied: nothing unknown is left
        } else {
            const bool has1 = e1c != BL_CODE_STALE && e1c != BL_CODE_NONE, has2 = has1 && e2c != BL_CODE_NONE;
            if (has1) q1 = bl_code_hi(e1c) == k ? bl_code_lo(e1c) : bl_code_hi(e1c);
            if (has2) q2 = bl_code_hi(e2c) == k ? bl_code_lo(e2c) : bl_code_hi(e2c);
            if (!from_list) {
                dropped = 0;
                for (int j = 0; j < 2 * nc_; j++) { const int x = Bbid[j]; if (x >= 0) dropped |= (x == q1 ? 1 : 0) | (x == q2 ? 2 : 0); }
            }
            if (has1 && !(dropped & 1)) insert(c_mts[bl], e1c);
            if (has2 && !(dropped & 2)) insert(c_sts[bl], e2c);
        }
        const int cnt = from_list ? hl_cnt[bl] : BL_HL_CAP + 1;
        const int nscan = cnt <= BL_HL_CAP ? cnt : 2 * nc_;
        for (int i = 0; i < nscan; i++) {
            const int j = cnt <= BL_HL_CAP ? hl[bl * BL_HL_CAP + i] : i;
            const int r = j >> 1, x = Bbid[j];
            if (r >= nc_ || x < 0 || r == myrank_ || (mine_ && r < myrank_)) continue;
            const double v = stash[j * Tb + bl];
            if (v < BL_INF) {
                const uint64_t vs = bl_sortable(v), vc = bl_pair_code(k, x);
                if (bl_key_less(vs, vc, lt, lc)) insert(vs, vc);
            }
        }
        if (k3c != ~0ull) { lt = k3t; lc = k3c; }
        c_lts[bl] = lt; c_lcode[bl] = lc;
        c_sts[bl] = k2c != ~0ull ? k2t : ts_inf; c_scode[bl] = k2c != ~0ull ? k2c : BL_CODE_NONE;
        if (k1c != ~0ull) { c_mts[bl] = k1t; c_mcode[bl] = k1c; }
        else if (lc == BL_CODE_NONE) { c_mts[bl] = ts_inf; c_mcode[bl] = BL_CODE_NONE; }
        else { c_mts[bl] = lt; c_mcode[bl] = BL_CODE_STALE; }  // the bound's time is what is posted
    };
    // state of record, obstacle entry, sequence number and global mirror of a ball whose collision is committed
    auto primary_apply = [&](const double *Bx, const double *By, const double *Bvx, const double *Bvy, const double *Bt,
                             const int *Bbid, const int myrank_, const bool mine_, const long long done_base,
                             const bool save_undo) {
        if (save_undo) {
            undo[bl] = c_mts[bl]; undo[Tb + bl] = c_mcode[bl];
            undo[11 * Tb + bl] = c_sts[bl]; undo[12 * Tb + bl] = c_scode[bl]; undo[13 * Tb + bl] = c_lts[bl];
            undo[14 * Tb + bl] = c_lcode[bl];
        }
        if (mine_) {
            if (save_undo) {
                undo[2 * Tb + bl] = as_u(st_p0x[bl]); undo[3 * Tb + bl] = as_u(st_p0y[bl]); undo[4 * Tb + bl] = as_u(st_vx[bl]);
                undo[5 * Tb + bl] = as_u(st_vy[bl]); undo[6 * Tb + bl] = as_u(st_t0[bl]); undo[7 * Tb + bl] = c_ots[bl];
                undo[8 * Tb + bl] = as_u(c_oarg[bl]); undo[9 * Tb + bl] = (uint64_t)(unsigned)c_oidx[bl];
                undo[10 * Tb + bl] = (uint64_t)c_seq[bl];
            }
            const int myj = Bbid[2 * myrank_] == k ? 2 * myrank_ : 2 * myrank_ + 1;
            const double nx = Bx[myj], ny = By[myj], nvx = Bvx[myj], nvy = Bvy[myj], nt0 = Bt[myj];
            st_p0x[bl] = nx; st_p0y[bl] = ny; st_vx[bl] = nvx; st_vy[bl] = nvy; st_t0[bl] = nt0;
            c_ots[bl] = scratch[3 * bl];
            const double oarg = as_d(scratch[3 * bl + 1]);
            const int oi = (int)scratch[3 * bl + 2];
            c_oarg[bl] = oarg; c_oidx[bl] = oi;
            const long long sq = seq_base + done_base + myrank_ + 1;
            c_seq[bl] = sq;
            // keep the global mirror current: resolvers read ball states from it
            __stcg(S.p0x + k, nx); __stcg(S.p0y + k, ny); __stcg(S.vx + k, nvx); __stcg(S.vy + k, nvy);
            __stcg(S.t0 + k, nt0);
            __stcg(S.obs_arg + k, oarg); __stcg(S.obs_idx + k, oi);
            __stcg(S.seq + k, sq);
        }
    };
    // clock, counters and window after nc_ of the nf_ collisions of a batch have been committed
    auto advance_clock = [&](const double *Bt, const int *Br_info, const int nc_, const int nf_, const int nE_) {
        n_bb += Br_info[4 * (nc_ - 1) + 3];
        done += nc_;
        prev_now = now;
        now = Bt[2 * (nc_ - 1)];
        if (mode == MODE_ARGMIN) {
            kindmask = KIND_BOTH;  // only the first event's kind was forced
            if (!(win > 0.0)) {
                const double gap = bl_sub(now, prev_now);
                if (done >= 2 && gap > 0.0) win = bl_mul(gap, (double)P.target);
            }
            if (win > 0.0 && P.target > 1) mode = MODE_WINDOW;
        } else {
            if (nc_ < nf_) win = bl_mul(win, P.wc[0]);
            else if ((double)nE_ < P.wc[2] * P.target) win = bl_mul(win, P.wc[1]);
            else if ((double)nE_ > P.wc[4] * P.target) win = bl_mul(win, P.wc[3]);
        }
    };
    // the last batch: committed optimistically, validated (and its rows stored) one barrier later
    bool pending = false;
    int nf_prev = 0, pb = 0, u_nE = 0, u_mode = mode, u_kind = kindmask;
    long long u_done = 0, u_nbb = 0;
    double u_now = now, u_prev_now = now, u_win = win;

    for (;;) {
        long long c0 = 0;
        if (prof_thread) { c0 = clock64(); if (prof_lane) prof[5]++; }
        // batch tables of this round, and of the previous batch
        double *const bx = bx_b + pb * 2 * m_max, *const by = by_b + pb * 2 * m_max, *const bvx = bvx_b + pb * 2 * m_max;
        double *const bvy = bvy_b + pb * 2 * m_max, *const bt = bt_b + pb * 2 * m_max;
        int *const bid = bid_b + pb * 2 * m_max, *const brc = brc_b + pb * 2 * m_max;
        double *const r_pay = r_pay_b + pb * 13 * m_max;
        int *const r_info = r_info_b + pb * 4 * m_max, *const brank = brank_b + pb * Tb;
        uint64_t *const bts = bts_b + pb * m_max;
        const int qb = pb ^ 1;
        const uint64_t *const pbts = bts_b + qb * m_max;
        const double *const pbt = bt_b + qb * 2 * m_max;
        const int *const pbid = bid_b + qb * 2 * m_max;
        int *const pbrank = brank_b + qb * Tb;
        // =========================== 1. post =================================================================
        uint64_t theta = 0;
        if (mode == MODE_WINDOW) {
            theta = bl_sortable(bl_add(now, win));
            if (theta > end_excl) theta = end_excl;
        }
        if (tid == 0) *m_gmin = ~0ull;
        if (warp < (Tb >> 5)) {  // the primaries' warps
            uint64_t kts = ~0ull, kcode = BL_CODE_END;
            if (own) {
                if (kindmask != KIND_BO) { kts = c_mts[bl]; kcode = c_mcode[bl]; }
                if (kindmask != KIND_BB) {
                    const uint64_t ots = c_ots[bl], oc = BL_CODE_OBS | (uint64_t)k;
                    if (kindmask == KIND_BO || bl_key_less(ots, oc, kts, kcode)) { kts = ots; kcode = oc; }
                }
            }
            if (mode == MODE_WINDOW) {
                if (own && kts < theta) {
                    const int s = atomicAdd(&misc[M_PCOUNT], 1);
                    if (s < BL_R_POST) { pl_ts[s] = kts; pl_code[s] = kcode; }
                }
            } else {
                const int w = warp_argmin(kts, kcode);
                if (lane == w) { w_ts[warp] = kts; w_code[warp] = kcode; }
            }
        }
        __syncthreads();
        if (mode == MODE_ARGMIN) {
            if (warp == 0) {
                uint64_t ts = ~0ull, cd = BL_CODE_END;
                if (lane < (Tb >> 5)) { ts = w_ts[lane]; cd = w_code[lane]; }
                const int w2 = Tb > 32 ? warp_argmin(ts, cd) : 0;
                if (lane == w2) { pl_ts[0] = ts; pl_code[0] = cd; misc[M_PCOUNT] = cd != BL_CODE_END ? 1 : 0; }
            }
            __syncthreads();
        }
        if (tid <= 2 * BL_R_POST) {
            uint64_t *const post = P.post + ((size_t)par * BL_MAX_CTAS + cta) * BL_POST_WORDS;
            const int pc = misc[M_PCOUNT];
            uint64_t word;
            if (tid == 0) word = pack2(pc < BL_R_POST ? pc : BL_R_POST, pc > BL_R_POST ? 1 : 0);
            else word = (tid & 1) ? pl_ts[(tid - 1) >> 1] : pl_code[(tid - 1) >> 1];
            __stcg(post + (tid ? tid + 1 : 0), word);  // header in word 0, key e in words 2+2e, 3+2e (16-byte aligned)
            // word 1: the earliest violating new key of the batch this CTA has just committed optimistically
            if (tid == 0) __stcg(post + 1, pending ? (uint64_t)*m_viol : ~0ull);
        }
        long long cb = 0;
        if (prof_thread) cb = clock64();
        sync_all(P, epoch);
        if (tid == 0) *m_viol = ~0ull;
        if (prof_thread) { const long long ce = clock64(); if (prof_lane) { prof[8] += cb - c0; prof[9] += ce - cb; } c0 = ce; }
        // =========================== 2. collect ==============================================================
        {
            const uint64_t *const box = P.post + (size_t)par * BL_MAX_CTAS * BL_POST_WORDS;
            for (int c = tid; c < ncta; c += T) {
                const uint64_t *rec = box + (size_t)c * BL_POST_WORDS;
                // one trip: header and the first two keys (a CTA rarely posts more)
                const uint64_t hd = __ldcg(rec), vw = __ldcg(rec + 1);
                const ulonglong2 k0 = __ldcg(reinterpret_cast<const ulonglong2 *>(rec + 2));
                const ulonglong2 k1 = __ldcg(reinterpret_cast<const ulonglong2 *>(rec + 4));
                const int cnt = lo32(hd);
                if (hi32(hd)) misc[M_ACTION] = ACT_OVERFLOW;
                if (vw != ~0ull) atomicMin(m_gmin, (unsigned long long)vw);
                if (cnt > 0) {
                    const int slot = atomicAdd(&misc[M_ECOUNT], cnt);
                    if (slot < BL_EV_CAP) { e_ts[slot] = k0.x; e_code[slot] = k0.y; e_ab[slot] = key_words(k0.y); }
                    if (cnt > 1 && slot + 1 < BL_EV_CAP) { e_ts[slot + 1] = k1.x; e_code[slot + 1] = k1.y; e_ab[slot + 1] = key_words(k1.y); }
                    for (int e = 2; e < cnt; e++)
                        if (slot + e < BL_EV_CAP) {
                            const uint64_t ce = __ldcg(rec + 3 + 2 * e);
                            e_ts[slot + e] = __ldcg(rec + 2 + 2 * e); e_code[slot + e] = ce; e_ab[slot + e] = key_words(ce);
                        }
                }
            }
        }
        par ^= 1;
        __syncthreads();
        if (prof_thread) { const long long cc = clock64(); if (prof_lane) prof[11] += cc - c0; c0 = cc; }

        // =========================== 2b. validate the previous batch ===========================================
        // It was committed optimistically (candidates, states, mirrors) before this round's keys were posted; its
        // violation minimum came with the keys.  Normally nothing violates: only the row stores and the log are
        // left to do, and they run beside warp 0's serial section further down.  Otherwise the commit is rolled back
        // and redone for the valid prefix, and the keys -- posted from a state that never existed -- are thrown away.
        int nc_prev = 0;  // collisions of the previous batch whose rows are still to be stored
        if (pending) {
            const uint64_t gmin = (uint64_t)*m_gmin;
            // committed: the first collision always, the others while they come before every violation (the times
            // are sorted, so that is a count)
            int nc = __popc(__ballot_sync(0xffffffffu, lane < nf_prev && pbts[lane < nf_prev ? lane : 0] < gmin)) +
                     __popc(__ballot_sync(0xffffffffu, lane + 32 < nf_prev && pbts[lane + 32 < nf_prev ? lane + 32 : 0] < gmin));
            if (nc < 1) nc = 1;
            pending = false;
            if (nc < nf_prev) {
                const int prank = pbrank[bl];
                if (primary && own) {
                    c_mts[bl] = undo[bl]; c_mcode[bl] = undo[Tb + bl];
                    c_sts[bl] = undo[11 * Tb + bl]; c_scode[bl] = undo[12 * Tb + bl]; c_lts[bl] = undo[13 * Tb + bl];
                    c_lcode[bl] = undo[14 * Tb + bl];
                    if (prank >= 0 && prank < nf_prev) {
                        const double ox = as_d(undo[2 * Tb + bl]), oy = as_d(undo[3 * Tb + bl]), ovx = as_d(undo[4 * Tb + bl]);
                        const double ovy = as_d(undo[5 * Tb + bl]), ot0 = as_d(undo[6 * Tb + bl]), oarg = as_d(undo[8 * Tb + bl]);
                        const int oi = (int)undo[9 * Tb + bl];
                        const long long sq = (long long)undo[10 * Tb + bl];
                        st_p0x[bl] = ox; st_p0y[bl] = oy; st_vx[bl] = ovx; st_vy[bl] = ovy; st_t0[bl] = ot0;
                        c_ots[bl] = undo[7 * Tb + bl]; c_oarg[bl] = oarg; c_oidx[bl] = oi; c_seq[bl] = sq;
                        __stcg(S.p0x + k, ox); __stcg(S.p0y + k, oy); __stcg(S.vx + k, ovx); __stcg(S.vy + k, ovy);
                        __stcg(S.t0 + k, ot0); __stcg(S.obs_arg + k, oarg); __stcg(S.obs_idx + k, oi); __stcg(S.seq + k, sq);
                    }
                }
                done = u_done; n_bb = u_nbb; now = u_now; prev_now = u_prev_now; win = u_win; mode = u_mode; kindmask = u_kind;
                __syncthreads();
                const int myrank_p = prank >= 0 && prank < nf_prev ? prank : -1;
                const bool mine_p = myrank_p >= 0 && myrank_p < nc;
                row_stores(pbid, pbrank, nf_prev, nc, warp, nwarps);
                if (want_log && cta == 0) write_log(r_pay_b + qb * 13 * m_max, r_info_b + qb * 4 * m_max, pbid, u_done, nc, tid, T);
                if (primary && own) {
                    cover_fold(pbid, nc, myrank_p, mine_p, false, 0);
                    primary_apply(bx_b + qb * 2 * m_max, by_b + qb * 2 * m_max, bvx_b + qb * 2 * m_max, bvy_b + qb * 2 * m_max,
                                  pbt, pbid, myrank_p, mine_p, u_done, false);
                }
                advance_clock(pbt, r_info_b + qb * 4 * m_max, nc, nf_prev, u_nE);
                if (prof_lane) prof[7]++;
                // start over from the corrected state: drop the collected keys
                __syncthreads();
                if (primary) pbrank[bl] = -1;
                if (tid == 0) { misc[M_ECOUNT] = 0; misc[M_ACTION] = ACT_BATCH; misc[M_PCOUNT] = 0; }
                if (tid < BL_R_POST) { pl_ts[tid] = ~0ull; pl_code[tid] = BL_CODE_END; }
                __syncthreads();
                continue;
            }
            nc_prev = nc;
        }

        // =========================== 3. resolve + rank (all warps) =============================================
        // Every collected key is resolved speculatively, one thread per key spread over the warps (elastic_collision /
        // obstacle reflection from the global mirrors), into a record by collection slot; which of them form the batch
        // is decided afterwards.  While the mirror loads are in flight the warps rank the keys pairwise: entry j comes
        // before entry i -> rank[i]++; and if they share a ball, i is a duplicate (same pair, posted by both balls) or
        // a conflict.
        const int nE_all = misc[M_ECOUNT];
        // Roles of the warps until the batch is known: the first quarter resolves (waiting on L2 most of the time), the
        // others rank.  (Handing the row stores to a quarter of the warps through both sections was slower.)
        // (measured at N = 16384, 16 warps: 8 resolver warps 4.24 k cycles for this section, 4: 3.71 k, 2: 3.87 k, 1: 5.27 k)
        const int rw = nwarps >= 8 ? nwarps >> 2 : (nwarps >= 2 ? nwarps >> 1 : 1);
        double *const qs_d = wq;
        int *const qs_i = reinterpret_cast<int *>(wq + BL_SORT_CAP * QS_D);
        if (mode == MODE_ARGMIN) {
            // only the smallest of the per-CTA minima is known to be next: it moves to slot 0
            if (warp == 0) {
                uint64_t ts0 = ~0ull, cd0 = BL_CODE_END;
                for (int i = lane; i < nE_all && i < BL_EV_CAP; i += 32)
                    if (bl_key_less(e_ts[i], e_code[i], ts0, cd0)) { ts0 = e_ts[i]; cd0 = e_code[i]; }
                const int w = warp_argmin(ts0, cd0);
                ts0 = shfl_u64(ts0, w); cd0 = shfl_u64(cd0, w);
                if (lane == 0) { e_ts[0] = ts0; e_code[0] = cd0; }
            }
            __syncthreads();
        }
        {
            const int nres = final_stage != 0 ? 0 : (mode == MODE_ARGMIN ? (nE_all > 0 ? 1 : 0) : (nE_all <= BL_SORT_CAP ? nE_all : 0));
            if (warp < rw) {
                // (consecutive lanes write consecutive records: the odd record stride keeps them on different banks)
                const int per = (BL_SORT_CAP + rw - 1) / rw;
                for (int e = warp * per + lane; e < nres && e < (warp + 1) * per; e += 32) {
                    const uint64_t ts = e_ts[e], cd = e_code[e];
                    int a, b;
                    key_balls(cd, a, b);
                    if (a < 0) continue;
                    double b0x = 0, b0y = 0, b_vx = 0, b_vy = 0, b_t0 = 0;
                    double m1 /* obstacle: argument of the hit */, m2 = 0;
                    int rcb /* obstacle: its index */;
                    const double a0x = __ldcg(S.p0x + a), a0y = __ldcg(S.p0y + a), a_vx = __ldcg(S.vx + a), a_vy = __ldcg(S.vy + a);
                    const double a_t0 = __ldcg(S.t0 + a);
                    const int rca = S.rclass[a];
                    if (b >= 0) {
                        b0x = __ldcg(S.p0x + b); b0y = __ldcg(S.p0y + b); b_vx = __ldcg(S.vx + b); b_vy = __ldcg(S.vy + b);
                        b_t0 = __ldcg(S.t0 + b);
                        m1 = S.mass[a]; m2 = S.mass[b];
                        rcb = S.rclass[b];
                    } else {
                        m1 = __ldcg(S.obs_arg + a);
                        rcb = __ldcg(S.obs_idx + a);
                    }
                    const double t = bl_unsortable(ts);
                    double *const q = qs_d + QS_D * e;
                    int *const qi = qs_i + QS_I * e;
                    const double ax = bl_advance(a0x, a_vx, t, a_t0), ay = bl_advance(a0y, a_vy, t, a_t0);
                    if (b < 0) {
                        // ball-obstacle, simulation.py:504-519 + :604-636
                        const int oi = rcb;
                        double nvx = a_vx, nvy = a_vy;
                        const int st = oi >= 0 ? bl_obstacle_bounce<FMA>(obs[oi], ax, ay, a_vx, a_vy, m1, nvx, nvy) : BL_EASSERT;
                        q[0] = ax; q[1] = ay; q[2] = a_vx; q[3] = a_vy; q[4] = nvx; q[5] = nvy; q[6] = m1;
                        q[7] = 0; q[8] = 0; q[9] = 0; q[10] = 0; q[11] = 0; q[12] = t;
                        qi[0] = st; qi[1] = oi; qi[2] = 0; qi[3] = rca; qi[4] = 0;
                    } else {
                        // ball-ball, simulation.py:446-455 + :562-602 (a = idx1 < b = idx2)
                        const double bxx = bl_advance(b0x, b_vx, t, b_t0), byy = bl_advance(b0y, b_vy, t, b_t0);
                        bl_remap_masses(m1, m2);
                        double avx = a_vx, avy = a_vy, nbx = b_vx, nby = b_vy;
                        const int st = bl_elastic<FMA>(ax, ay, a_vx, a_vy, m1, bxx, byy, b_vx, b_vy, m2, avx, avy, nbx, nby);
                        q[0] = ax; q[1] = ay; q[2] = a_vx; q[3] = a_vy; q[4] = avx; q[5] = avy;
                        q[6] = bxx; q[7] = byy; q[8] = b_vx; q[9] = b_vy; q[10] = nbx; q[11] = nby; q[12] = t;
                        qi[0] = st; qi[1] = -1; qi[2] = 1; qi[3] = rca; qi[4] = rcb;
                    }
                }
                if (prof_thread) { const long long cq = clock64(); if (prof_lane) prof[13] += cq - c0; }
            }
            const long long cr0 = (nwarps >= 2 && warp == rw) ? clock64() : 0;
            if ((nwarps < 2 || warp >= rw) && mode == MODE_WINDOW && nE_all > 1 && nE_all <= BL_SORT_CAP) {
                // G neighbouring lanes share entry i; each compares it with every G-th entry j: j comes before i
                // -> rank[i]++; and if they share a ball, i is a duplicate (same pair) or a conflict
                const int nr = nE_all;
                const int pt = nwarps < 2 ? tid : tid - 32 * rw, PT = nwarps < 2 ? T : T - 32 * rw;  // ranking threads
                // as many lanes per key as one pass over the keys allows (PT / G keys per pass); G is a compile-time
                // constant of the loop
                auto rank_pass = [&](auto g_tag) {
                    constexpr int G = decltype(g_tag)::value;
                    const int c = pt & (G - 1);
                    const unsigned gmask = G == 32 ? 0xffffffffu : ((1u << G) - 1u) << (lane & ~(G - 1));
                    for (int i0 = (pt - lane) / G; i0 < nr; i0 += PT / G) {  // first entry of my warp: one trip count for its lanes
                        const int i = i0 + lane / G;
                        const bool act = i < nr;
                        const uint64_t ti = act ? e_ts[i] : 0, ci = act ? e_code[i] : 0, wi = act ? e_ab[i] : ~0ull;
                        // my side's absent balls must not equal the other side's
                        const unsigned Ai = (unsigned)wi == BL_NO_BALL ? BL_NO_BALL - 1 : (unsigned)wi;
                        const unsigned Bi = (unsigned)(wi >> 32) == BL_NO_BALL ? BL_NO_BALL - 1 : (unsigned)(wi >> 32);
                        int rk = 0, fl = 0;
#pragma unroll 4
                        for (int j = c; j < nr; j += G) {
                            const uint64_t tj = e_ts[j], cj = e_code[j], wj = e_ab[j];
                            const unsigned before = key_before_mask(tj, cj, (unsigned)j, ti, ci, (unsigned)i);
                            const unsigned Aj = (unsigned)wj, Bj = (unsigned)(wj >> 32);
                            const bool share = Aj == Ai || Aj == Bi || Bj == Ai || Bj == Bi;
                            rk -= (int)before;
                            fl |= (int)(before & (unsigned)(cj == ci ? F_DUP : (share ? F_CONFLICT : 0)));
                        }
                        if (Ai == BL_NO_BALL - 1) fl = 0;  // a key without balls is neither
                        if (G > 1) {
                            rk = __reduce_add_sync(gmask, rk);
                            fl = __reduce_or_sync(gmask, fl);
                        }
                        if (act && c == 0) { a_rank[i] = rk; a_flag[i] = fl; }
                    }
                };
                if (PT >= 256) {
                    if (nr * 8 <= PT) rank_pass(std::integral_constant<int, 8>());
                    else rank_pass(std::integral_constant<int, 4>());
                } else if (PT >= 128) rank_pass(std::integral_constant<int, 2>());
                else rank_pass(std::integral_constant<int, 1>());
            }
            if (nwarps >= 2 && warp == rw) { const long long cq = clock64(); if (lane == 0) prof[12] += cq - cr0; }  // the first ranking warp's clock
        }
        if (prof_thread) { const long long cq = clock64(); if (prof_lane) prof[1] += cq - c0; }
        __syncthreads();  // the resolved records and the ranks are in place
        if (prof_thread) { const long long cv = clock64(); if (prof_lane) prof[2] += cv - c0; c0 = cv; }

        // =========================== 3b. the batch (warp 0) beside the previous batch's row stores (the others) ==
        if (warp == 0) {
            int nE = nE_all;
            int action = misc[M_ACTION] == ACT_OVERFLOW ? ACT_OVERFLOW : ACT_BATCH;
            int stop_kind = 0, st_status = BL_OK, nf = 0;
            // sorted entries `lane` and `lane + 32`: key, flags, collection slot
            uint64_t ts0 = ~0ull, cd0 = BL_CODE_END, ts1 = ~0ull, cd1 = BL_CODE_END;
            int fl0 = 0, fl1 = 0, sl0 = 0, sl1 = 0;
            if (mode == MODE_ARGMIN) {
                nE = nE > 0 ? 1 : 0;
                if (lane == 0 && nE) { ts0 = e_ts[0]; cd0 = e_code[0]; }
            } else if (nE > BL_SORT_CAP) {
                action = ACT_OVERFLOW;
            } else if (nE == 1) {
                if (lane == 0) { ts0 = e_ts[0]; cd0 = e_code[0]; }
            } else if (nE > 1) {
                // into sorted order (ranks and flags come from the pair pass)
                if (lane < nE) { const int rk = a_rank[lane]; s_ts[rk] = e_ts[lane]; s_code[rk] = e_code[lane]; s_flag[rk] = a_flag[lane] | (lane << 8); }
                if (lane + 32 < nE) { const int rk = a_rank[lane + 32]; s_ts[rk] = e_ts[lane + 32]; s_code[rk] = e_code[lane + 32]; s_flag[rk] = a_flag[lane + 32] | ((lane + 32) << 8); }
                __syncwarp();
                if (lane < nE) { ts0 = s_ts[lane]; cd0 = s_code[lane]; fl0 = s_flag[lane]; }
                if (lane + 32 < nE) { ts1 = s_ts[lane + 32]; cd1 = s_code[lane + 32]; fl1 = s_flag[lane + 32]; }
                sl0 = fl0 >> 8; fl0 &= 0xff;
                sl1 = fl1 >> 8; fl1 &= 0xff;
            }
            const bool max_reached = P.max_events >= 0 && done >= P.max_events;
            if (action == ACT_OVERFLOW) {
            } else if (nE == 0) {
                if (final_stage == 0 && !max_reached && mode == MODE_WINDOW && theta < end_excl) action = ACT_EMPTY;
                else { action = ACT_STOP; stop_kind = max_reached ? BL_STOP_EVENTS : BL_STOP_TIME; }
                if (lane == 0) { s_ts[0] = ts_inf; s_code[0] = BL_CODE_END; }
            } else {
                if (prof_thread) { const long long cq = clock64(); if (prof_lane) prof[14] += cq - c0; }
                // the batch ends before the first entry that is only a lower bound or shares a ball with an earlier one
                const unsigned cm0 = __ballot_sync(0xffffffffu, lane < nE && (cd0 == BL_CODE_STALE || (fl0 & F_CONFLICT)));
                const unsigned cm1 = __ballot_sync(0xffffffffu, lane + 32 < nE && (cd1 == BL_CODE_STALE || (fl1 & F_CONFLICT)));
                const int cut = cm0 ? __ffs(cm0) - 1 : (cm1 ? 31 + __ffs(cm1) : nE);
                const uint64_t tsf = shfl_u64(ts0, 0);
                if (final_stage == 0 && max_reached) {
                    action = ACT_STOP; stop_kind = BL_STOP_EVENTS;  // checked first: the clock stays at the last collision
                } else if (cut == 0) {
                    action = ACT_REPAIR;  // the front entry is only a lower bound (it cannot be a conflict)
                } else if (final_stage != 0) {
                    action = ACT_STOP; stop_kind = BL_STOP_TIME;  // query rounds only look at the front entry
                } else if (tsf == BL_TS_NAN) {
                    action = ACT_STOP; stop_kind = BL_STOP_ERROR; st_status = BL_ENAN;
                } else if (tsf >= end_excl) {
                    action = ACT_STOP; stop_kind = BL_STOP_TIME;
                } else if (want_log && done >= P.log.cap) {
                    action = ACT_STOP; stop_kind = BL_STOP_LOG;
                } else {
                    long long allowed = mode == MODE_ARGMIN ? 1 : m_max;
                    if (P.max_events >= 0 && P.max_events - done < allowed) allowed = P.max_events - done;
                    if (want_log && P.log.cap - done < allowed) allowed = P.log.cap - done;
                    const bool valid0 = lane < cut && !(fl0 & F_DUP) && ts0 < end_excl;
                    const bool valid1 = lane + 32 < cut && !(fl1 & F_DUP) && ts1 < end_excl;
                    const unsigned vm0 = __ballot_sync(0xffffffffu, valid0), vm1 = __ballot_sync(0xffffffffu, valid1);
                    const int r0 = __popc(vm0 & lt_mask), r1 = __popc(vm0) + __popc(vm1 & lt_mask);
                    nf = __popc(vm0) + __popc(vm1);
                    if (nf > allowed) nf = (int)allowed;
                    // ---- the batch tables, by rank r: one lane per collision copies its resolved record ----
                    int er_mine = 0x7fffffff;
                    const int halves = nf > __popc(vm0) ? 2 : 1;
#pragma unroll 1
                    for (int hf = 0; hf < halves; hf++) {
                        const uint64_t ts = hf ? ts1 : ts0, cd = hf ? cd1 : cd0;
                        const int r = hf ? r1 : r0;
                        if (!((hf ? valid1 : valid0) && r < nf)) continue;
                        int a, b;
                        key_balls(cd, a, b);
                        const double *const q = qs_d + QS_D * (hf ? sl1 : sl0);
                        const int *const qi = qs_i + QS_I * (hf ? sl1 : sl0);
                        const int st = qi[0], bb = qi[2];
                        const double t = q[12];
                        bx[2 * r] = q[0]; by[2 * r] = q[1]; bvx[2 * r] = q[4]; bvy[2 * r] = q[5]; bt[2 * r] = t; bts[r] = ts;
                        bid[2 * r] = a; brc[2 * r] = qi[3];
                        bt[2 * r + 1] = t;
                        if (bb) {
                            bx[2 * r + 1] = q[6]; by[2 * r + 1] = q[7]; bvx[2 * r + 1] = q[10]; bvy[2 * r + 1] = q[11];
                            bid[2 * r + 1] = b; brc[2 * r + 1] = qi[4];
                        } else {
                            bid[2 * r + 1] = -1; brc[2 * r + 1] = 0;
                        }
                        if (want_log) {
                            double *pl = r_pay + 13 * r;
#pragma unroll
                            for (int f = 0; f < 13; f++) pl[f] = q[f];
                        }
                        r_info[4 * r] = st; r_info[4 * r + 1] = qi[1]; r_info[4 * r + 2] = bb;
                        if (a >= base && a < base + Tb) brank[a - base] = r;
                        if (b >= base && b < base + Tb) brank[b - base] = r;
                        if (st != BL_OK && r < er_mine) er_mine = r;
                    }
                    if (prof_thread) { const long long cq = clock64(); if (prof_lane) prof[15] += cq - c0; }
                    // a collision that cannot be resolved (physics.py:277-279 / zero masses): commit what comes
                    // before it; if it is the front one, stop with the reference's error
                    const int er = __reduce_min_sync(0xffffffffu, er_mine);
                    if (er < nf) nf = er;
                    if (nf == 0) action = ACT_ERROR;
                    // running count of ball-ball collisions by rank
                    __syncwarp();
                    {
                        const unsigned b0m = __ballot_sync(0xffffffffu, lane < nf && r_info[4 * (lane < nf ? lane : 0) + 2] != 0);
                        const unsigned b1m = __ballot_sync(0xffffffffu, lane + 32 < nf && r_info[4 * (lane + 32 < nf ? lane + 32 : 0) + 2] != 0);
                        const unsigned le_mask = lt_mask | (1u << lane);
                        if (lane < nf) r_info[4 * lane + 3] = __popc(b0m & le_mask);
                        if (lane + 32 < nf) r_info[4 * (lane + 32) + 3] = __popc(b0m) + __popc(b1m & le_mask);
                    }
                }
                if (lane == 0) { s_ts[0] = ts0; s_code[0] = cd0; }  // the front entry, for the stop stages
            }
            if (lane == 0) {
                misc[M_ACTION] = action; misc[M_NF] = nf; misc[M_NE] = nE; misc[M_STOPKIND] = stop_kind;
                misc[M_STATUS] = st_status; misc[M_PCOUNT] = 0; misc[M_ECOUNT] = 0;
            }
            if (lane < BL_R_POST) { pl_ts[lane] = ~0ull; pl_code[lane] = BL_CODE_END; }
            __syncwarp();
            a_rank[lane] = 0; a_flag[lane] = 0; a_rank[lane + 32] = 0; a_flag[lane + 32] = 0;
        }
        if (ncta > 1 && tid == T - 1) {
            // Readers' gate.  The commit further down overwrites the global mirrors of the batch's balls, and a slower
            // CTA may still be resolving from them: every CTA says here that its mirror loads are done (they were
            // consumed before the barrier above), and nobody leaves this section -- so nobody commits -- before all
            // have said so.  A second grid barrier in effect, but it is waited for beside warp 0's serial section
            // and the row stores instead of on the critical path.
            reads_epoch += (unsigned)ncta;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(P.bar + 8) : "memory");
            while ((int)(ld_acquire_u32(P.bar + 8) - reads_epoch) < 0) {}
        }
        if (nc_prev > 0 && (warp != 0 || nwarps == 1)) {
            // the previous batch stands: its rows and log records, by every warp but the one that is busy above
            const int sw = nwarps == 1 ? 0 : warp - 1, nsw = nwarps == 1 ? 1 : nwarps - 1;
            row_stores(pbid, pbrank, nf_prev, nc_prev, sw, nsw);
            if (want_log && cta == 0)
                write_log(r_pay_b + qb * 13 * m_max, r_info_b + qb * 4 * m_max, pbid, u_done, nc_prev, sw * 32 + lane, nsw * 32);
        }
        __syncthreads();
        const int action = misc[M_ACTION];
        int nf = misc[M_NF];
        if (primary) pbrank[bl] = -1;  // the row stores have read the ranks of the previous batch by now
        const int nE = misc[M_NE];
        if (prof_thread) { const long long c1 = clock64(); if (prof_lane) prof[0] += c1 - c0; c0 = c1; }

        if (action == ACT_OVERFLOW) {  // too many keys below theta: halve the window and post again
            win = bl_mul(win, 0.5);
            if (++shrinks >= 6) mode = MODE_ARGMIN;
            if (prof_lane) prof[7]++;
            __syncthreads();
            if (tid == 0) misc[M_ACTION] = ACT_BATCH;
            continue;
        }
        shrinks = 0;
        if (action == ACT_EMPTY) {  // empty window: find the true minimum
            mode = MODE_ARGMIN;
            win = bl_mul(win, 2.0);
            if (prof_lane) prof[7]++;
            continue;
        }

        // =========================== repair: exact minima for the stale candidates near the front =============
        // The pairs of a stale ball k that matter are those written at a LATER collision of the other ball x
        // (seq[x] >= seq[k]); their current values sit in row x, column k.  Every ball x reads its own row at the
        // columns of all balls under repair (its rows: a handful of pages), the minima are reduced per CTA in
        // shared memory, across CTAs through global memory.
        if (action == ACT_REPAIR) {
            uint64_t hz = bl_sortable(bl_add(now, bl_mul(win, P.repair_horizon)));
            if (!(hz >= s_ts[0])) hz = s_ts[0];
            const bool sel = primary && own && c_mcode[bl] == BL_CODE_STALE && c_mts[bl] <= hz;
            int li = -1;
            if (sel) { li = atomicAdd(&misc[M_NDIRTY], 1); dirty[li] = bl; }
            __syncthreads();
            const int nd = misc[M_NDIRTY];
            if (tid == 0) misc[M_SBASE] = (ncta > 1 && nd) ? (int)atomicAdd(P.bar + 4, (unsigned)nd) : 0;
            __syncthreads();
            const int sbase = misc[M_SBASE];
            if (sel) {
                __stcg(P.stale_list + 2 * (size_t)(sbase + li), (uint64_t)k);
                __stcg(P.stale_list + 2 * (size_t)(sbase + li) + 1, (uint64_t)c_seq[bl]);
            }
            sync_all(P, epoch);
            const int total = ncta > 1 ? (int)__ldcg(P.bar + 4) : nd;
            long long *q_seq = reinterpret_cast<long long *>(sm + L.stash);
            unsigned long long *smin_ts = reinterpret_cast<unsigned long long *>(sm + L.stash + BL_RCHUNK);
            int *q_k = reinterpret_cast<int *>(sm + L.stash + 2 * BL_RCHUNK);
            int *smin_x = q_k + BL_RCHUNK;
            const long long myseq = c_seq[bl];
            int chunk = 0;
            for (int q0 = 0; q0 < total; q0 += BL_RCHUNK, chunk++) {
                const int cn = total - q0 < BL_RCHUNK ? total - q0 : BL_RCHUNK;
                for (int e = tid; e < cn; e += T) {
                    q_k[e] = (int)__ldcg(P.stale_list + 2 * (size_t)(q0 + e));
                    q_seq[e] = (long long)__ldcg(P.stale_list + 2 * (size_t)(q0 + e) + 1);
                    smin_ts[e] = ~0ull; smin_x[e] = 0x7fffffff;
                }
                __syncthreads();
                for (int g = 0; g < cn; g += 8 * H) {
                    const int e0 = g + 8 * h;  // helper h takes the h-th group of eight
                    double v[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int e = e0 + u;
                        const bool ok = own && e < cn && q_k[e] != k && myseq >= q_seq[e];
                        v[u] = ok ? __ldcg(myrow + q_k[e]) : BL_INF;
                    }
                    unsigned hit = 0;
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        if (v[u] < BL_INF) {
                            const unsigned long long ts = bl_sortable(v[u]);
                            if (ts <= smin_ts[e0 + u]) { atomicMin(&smin_ts[e0 + u], ts); hit |= 1u << u; }
                        }
                    }
                    __syncthreads();
                    // ties: the smallest pair code = the smallest partner index
#pragma unroll
                    for (int u = 0; u < 8; u++)
                        if ((hit >> u) & 1u)
                            if (bl_sortable(v[u]) == smin_ts[e0 + u]) atomicMin(&smin_x[e0 + u], k);
                }
                __syncthreads();
                if (ncta > 1) {
                    uint64_t *const part = P.partial + ((size_t)(chunk & 1) * BL_MAX_CTAS + cta) * BL_RCHUNK * 2;
                    for (int e = tid; e < cn; e += T) { __stcg(part + 2 * e, (uint64_t)smin_ts[e]); __stcg(part + 2 * e + 1, (uint64_t)(unsigned)smin_x[e]); }
                    sync_all(P, epoch);
                }
                for (int d = warp; d < nd; d += nwarps) {
                    const int e = sbase + d - q0;
                    if (e < 0 || e >= cn) continue;
                    uint64_t ts = ~0ull, xx = 0x7fffffffull;
                    if (ncta > 1) {
                        const uint64_t *const pbase = P.partial + (size_t)(chunk & 1) * BL_MAX_CTAS * BL_RCHUNK * 2;
                        for (int c = lane; c < ncta; c += 32) {
                            const uint64_t t2 = __ldcg(pbase + ((size_t)c * BL_RCHUNK + e) * 2);
                            const uint64_t x2 = __ldcg(pbase + ((size_t)c * BL_RCHUNK + e) * 2 + 1);
                            if (bl_key_less(t2, x2, ts, xx)) { ts = t2; xx = x2; }
                        }
                        const int w = warp_argmin(ts, xx);
                        ts = shfl_u64(ts, w);
                        xx = shfl_u64(xx, w);
                    } else {
                        ts = smin_ts[e]; xx = (uint64_t)(unsigned)smin_x[e];
                    }
                    if (lane == 0) { scratch[3 * dirty[d]] = ts; scratch[3 * dirty[d] + 1] = xx; }
                }
                __syncthreads();
            }
            if (sel) {
                const uint64_t ts = scratch[3 * bl];
                // the exact minimum of everything current in my row: the rest comes later
                c_sts[bl] = ts_inf; c_scode[bl] = BL_CODE_NONE;
                if (ts == ~0ull) { c_mts[bl] = ts_inf; c_mcode[bl] = BL_CODE_NONE; c_lts[bl] = ts_inf; c_lcode[bl] = BL_CODE_NONE; }
                else {
                    const uint64_t cc = bl_pair_code(k, (int)scratch[3 * bl + 1]);
                    c_mts[bl] = ts; c_mcode[bl] = cc; c_lts[bl] = ts; c_lcode[bl] = cc + 1;
                }
            }
            if (tid == 0) { misc[M_NDIRTY] = 0; if (cta == 0 && ncta > 1) *(P.bar + 4) = 0; }
            if (prof_thread) { const long long cq = clock64(); if (prof_lane) { prof[3] += cq - c0; prof[4]++; } }
            n_rescans += nd;
            __syncthreads();
            continue;
        }

        // =========================== stop / query stages / error ==============================================
        if (action == ACT_STOP || action == ACT_ERROR) {
            const bool writer = cta == 0 && tid == 0;
            const uint64_t cd = s_code[0];
            const double t = bl_unsortable(s_ts[0]);
            if (action == ACT_ERROR) {
                // rank 0 of the batch could not be resolved: the reference has moved to t and raised
                if (writer) {
                    res->err_i = bid[0]; res->err_j = bid[1];
                    res->time = bt[0];
                    res->stop = BL_STOP_ERROR; res->status = r_info[0];
                    res->n_ball_ball = n_bb; res->n_ball_obstacle = done - n_bb;
                    res->n_logged = want_log ? done : 0;
                }
                now = bt[0];
                final_stage = 1; kindmask = KIND_BB; mode = MODE_ARGMIN;
                __syncthreads();
                if (primary) brank[bl] = -1;
                __syncthreads();
                continue;
            }
            const int stop_kind = misc[M_STOPKIND], st_status = misc[M_STATUS];
            if (final_stage == 0) {
                if (writer) {
                    double tnow = now;
                    res->err_i = -1; res->err_j = -1;
                    if (st_status != BL_OK) { int a, b; key_balls(cd, a, b); res->err_i = a; res->err_j = b; tnow = t; }
                    if (stop_kind == BL_STOP_TIME && P.end_time < BL_INF) tnow = P.end_time;  // _move(end_time), :436
                    res->time = tnow; res->stop = stop_kind; res->status = st_status;
                    res->n_ball_ball = n_bb; res->n_ball_obstacle = done - n_bb;
                    res->n_logged = want_log ? done : 0;
                }
                final_stage = 1; kindmask = KIND_BB; mode = MODE_ARGMIN;
                __syncthreads();
                continue;
            }
            if (final_stage == 1) {
                if (writer) {
                    if (P.query_only) { res->time = P.time; res->err_i = -1; res->err_j = -1; }
                    const bool some = cd != BL_CODE_NONE && cd != BL_CODE_END && t < BL_INF;
                    res->bb_t = some ? t : BL_INF; res->bb_i = some ? bl_code_lo(cd) : -1; res->bb_j = some ? bl_code_hi(cd) : 0;
                }
                final_stage = 2; kindmask = KIND_BO;
                __syncthreads();
                continue;
            }
            if (writer) {
                // numpy argmin over _obstacles_toi: first ball attaining the minimum, even when it is +inf
                const bool some = cd != BL_CODE_END;
                const long long ball = some ? (long long)(cd & ~BL_CODE_OBS) : -1;
                res->bo_t = some ? t : BL_INF; res->bo_ball = ball;
                res->bo_obs = some ? (long long)__ldcg(S.obs_idx + ball) : -1;
                res->bo_arg = some ? __ldcg(S.obs_arg + ball) : 0.0;
            }
            break;
        }
        if (prof_thread) c0 = clock64();

        // =========================== 4. solve: my ball against the balls of the batch =========================
        // helper h takes the collisions r = h, h + H, ...
        int myrank = brank[bl];
        if (myrank >= nf) myrank = -1;
        const uint64_t vbound = bts[nf - 1];  // time of the last collision of the batch
        uint64_t vmin = ~0ull;
        const double p0x = st_p0x[bl], p0y = st_p0y[bl], vx = st_vx[bl], vy = st_vy[bl], t0 = st_t0[bl];
        double nx = 0, ny = 0, nvx = 0, nvy = 0, nt0 = 0;   // my state after my own collision (batch balls only)
        int qpart = -1, qpart2 = -1;  // partners of my two candidates (a ball of the batch overwrites those entries)
        if (own && myrank >= 0) {
            const int myj = bid[2 * myrank] == k ? 2 * myrank : 2 * myrank + 1;
            nx = bx[myj]; ny = by[myj]; nvx = bvx[myj]; nvy = bvy[myj]; nt0 = bt[myj];
            // my next obstacle after the collision, simulation.py:490-495 / :545-547: helper h looks at the obstacles
            // h, h + H, ... (list order, strict <); the primary combines the partial minima
            double t_min = BL_INF, arg_min = 0.0;
            int q_min = -1;
            const double radius = c_radius[bl];
            for (int q = h; q < n_obs; q += H) {
                double a;
                const double rs = obs[q].kind != BL_OBS_WALL ? S.rsum2_obs[(size_t)q * K + myrc] : 0.0;
                const double t = bl_obstacle_toi<FMA>(obs[q], nx, ny, nvx, nvy, radius, rs, a);
                if (t < t_min) { t_min = t; q_min = q; arg_min = a; }
            }
            obsp[(h * 3 + 0) * Tb + bl] = as_u(t_min); obsp[(h * 3 + 1) * Tb + bl] = as_u(arg_min);
            obsp[(h * 3 + 2) * Tb + bl] = (uint64_t)(unsigned)q_min;
        } else if (own) {
            const uint64_t mcode = c_mcode[bl];
            if (mcode != BL_CODE_STALE && mcode != BL_CODE_NONE) {
                qpart = bl_code_hi(mcode) == k ? bl_code_lo(mcode) : bl_code_hi(mcode);
                const uint64_t scode = c_scode[bl];
                if (scode != BL_CODE_NONE) qpart2 = bl_code_hi(scode) == k ? bl_code_lo(scode) : bl_code_hi(scode);
            }
        }
        // Pass 1, branch-free: advance my ball to the collision time once, then form the quadratic of
        // physics.py:33-54 against both balls of the collision.  Only ~5 % of the pairs get past the two early
        // exits; those are queued per warp and finished densely (sqrt + division) by all 32 lanes.
        {
            double *const q_b = wq + warp * (4 * BL_WQ_CAP), *const q_c = q_b + BL_WQ_CAP, *const q_d = q_c + BL_WQ_CAP;
            int *const q_dst = reinterpret_cast<int *>(q_d + BL_WQ_CAP);
            int qn = 0;
            // my state as of just before the collision at hand: the state of record, after my own collision the new one
            double sx = p0x, sy = p0y, svx = vx, svy = vy, st0 = t0;
            int switch_after = myrank >= 0 ? myrank : 0x7fffffff;
            for (int r = h;; r += H) {
                unsigned m0 = 0, m1 = 0;
                bool hit0 = false, hit1 = false;
                double b0 = 0, c0q = 0, d0 = 0, b1 = 0, c1q = 0, d1 = 0;
                const bool last = r >= nf;
                if (!last) {
                    if (__any_sync(0xffffffffu, r > switch_after)) {  // rare: a ball of the batch passes its own collision
                        if (r > switch_after) { sx = nx; sy = ny; svx = nvx; svy = nvy; st0 = nt0; switch_after = 0x7fffffff; }
                    }
                    const double t = bt[2 * r];
                    const bool valid = own && r != myrank;  // not me / my partner in the same collision
                    const double px = bl_advance(sx, svx, t, st0), py = bl_advance(sy, svy, t, st0);
                    const int x1 = bid[2 * r + 1];
                    if (qpart >= 0) {
                        const int x0 = bid[2 * r];
                        const int fl = ((x0 == qpart || x1 == qpart) ? 1 : 0) | ((qpart2 >= 0 && (x0 == qpart2 || x1 == qpart2)) ? 2 : 0);
                        if (fl) atomicOr(&ph[bl], fl);
                    }
                    {
                        const int j = 2 * r;
                        const double rs = (KONE || K == 1) ? rs11 : __ldg(S.rsum2 + (size_t)brc[j] * K + myrc);
                        const double dpx = bl_sub(px, bx[j]), dpy = bl_sub(py, by[j]);
                        const double dvx = bl_sub(svx, bvx[j]), dvy = bl_sub(svy, bvy[j]);
                        b0 = bl_dot2<FMA>(dpx, dpy, dvx, dvy);
                        c0q = bl_sub(bl_dot2<FMA>(dpx, dpy, dpx, dpy), rs);
                        d0 = bl_sub(bl_mul(b0, b0), bl_mul(bl_dot2<FMA>(dvx, dvy, dvx, dvy), c0q));
                        hit0 = valid && !(b0 >= 0.0) && !(d0 <= 0.0);
                        if (valid) stash[j * Tb + bl] = BL_INF;
                    }
                    if (x1 >= 0) {
                        const int j = 2 * r + 1;
                        const double rs = (KONE || K == 1) ? rs11 : __ldg(S.rsum2 + (size_t)brc[j] * K + myrc);
                        const double dpx = bl_sub(px, bx[j]), dpy = bl_sub(py, by[j]);
                        const double dvx = bl_sub(svx, bvx[j]), dvy = bl_sub(svy, bvy[j]);
                        b1 = bl_dot2<FMA>(dpx, dpy, dvx, dvy);
                        c1q = bl_sub(bl_dot2<FMA>(dpx, dpy, dpx, dpy), rs);
                        d1 = bl_sub(bl_mul(b1, b1), bl_mul(bl_dot2<FMA>(dvx, dvy, dvx, dvy), c1q));
                        hit1 = valid && !(b1 >= 0.0) && !(d1 <= 0.0);
                        if (valid) stash[j * Tb + bl] = BL_INF;
                    }
                    m0 = __ballot_sync(0xffffffffu, hit0);
                    m1 = __ballot_sync(0xffffffffu, hit1);
                }
                const int cnt = __popc(m0) + __popc(m1);
                if (last || qn + cnt > BL_WQ_CAP) {
                    // Pass 2, dense: t1 = c / (-b + sqrt(disc)); entry = time + (t1 if t1 >= t_eps else inf)
                    __syncwarp();
                    for (int i = lane; i < qn; i += 32) {
                        const int dst = q_dst[i], j = dst >> 16;
                        const double t1 = bl_div(q_c[i], bl_add(-q_b[i], bl_sqrt(q_d[i])));
                        const double v = bl_add(bt[j], t1 >= BL_T_EPS ? t1 : BL_INF);
                        stash[j * Tb + (dst & 0xffff)] = v;
                        if (v < BL_INF) {  // the few finite values of a ball: all its optimistic fold has to look at
                            const int ob = dst & 0xffff;
                            const int slot = atomicAdd(&hl_cnt[ob], 1);
                            if (slot < BL_HL_CAP) hl[ob * BL_HL_CAP + slot] = j;
                        }
                        const uint64_t vs = bl_sortable(v);
                        if (vs <= vbound && (j >> 1) < nf - 1) vmin = vs < vmin ? vs : vmin;
                    }
                    __syncwarp();
                    qn = 0;
                }
                if (last) break;
                if (hit0) { const int i = qn + __popc(m0 & lt_mask); q_b[i] = b0; q_c[i] = c0q; q_d[i] = d0; q_dst[i] = ((2 * r) << 16) | bl; }
                qn += __popc(m0);
                if (hit1) { const int i = qn + __popc(m1 & lt_mask); q_b[i] = b1; q_c[i] = c1q; q_d[i] = d1; q_dst[i] = ((2 * r + 1) << 16) | bl; }
                qn += __popc(m1);
            }
        }
        if (prof_thread) { const long long cs = clock64(); if (prof_lane) prof[10] += cs - c0; c0 = cs; }
        {
            const uint64_t wmin = warp_min_u64(vmin);
            if (lane == 0 && wmin != ~0ull) atomicMin(m_viol, (unsigned long long)wmin);
        }
        __syncthreads();  // all parked values and the CTA's violation minimum are in place

        // =========================== 5. optimistic commit =====================================================
        // as if nothing violated (nc = nf): candidates folded, states / mirrors / sequence numbers updated, so that
        // the next round's keys can be posted together with this round's violation word -- ONE grid barrier per
        // round.  The row stores and the log wait for the validation at the start of the next round.
        if (primary && own) {
            const bool mine = myrank >= 0;
            if (mine) {
                // combine the helpers' next-obstacle minima (first obstacle in list order among equal times)
                double t_min = BL_INF, arg_min = 0.0;
                int q_min = -1;
                for (int g = 0; g < H && g < n_obs; g++) {  // (helper g >= n_obs had no obstacle to look at)
                    const double t = as_d(obsp[(g * 3 + 0) * Tb + bl]);
                    const int qg = (int)obsp[(g * 3 + 2) * Tb + bl];
                    if (qg >= 0 && (t < t_min || (t == t_min && qg < q_min))) { t_min = t; q_min = qg; arg_min = as_d(obsp[(g * 3 + 1) * Tb + bl]); }
                }
                const uint64_t n_ots = bl_sortable(bl_add(bt[2 * myrank], t_min));
                scratch[3 * bl] = n_ots; scratch[3 * bl + 1] = as_u(arg_min); scratch[3 * bl + 2] = (uint64_t)(unsigned)q_min;
                if (n_ots <= vbound && myrank < nf - 1) atomicMin(m_viol, (unsigned long long)n_ots);
            }
            primary_apply(bx, by, bvx, bvy, bt, bid, myrank, mine, done, true);   // saves the undo copy first
            // fold the finite new values of my ball (listed by the solve pass; all of them if the list overflowed)
            cover_fold(bid, nf, myrank, mine, true, ph[bl]);
            hl_cnt[bl] = 0; ph[bl] = 0;
        }
        u_done = done; u_nbb = n_bb; u_now = now; u_prev_now = prev_now; u_win = win; u_mode = mode; u_kind = kindmask;
        u_nE = nE;
        advance_clock(bt, r_info, nf, nf, nE);
        pending = true;
        nf_prev = nf;
        pb ^= 1;
        if (prof_thread) { const long long cq = clock64(); if (prof_lane) prof[6] += cq - c0; }
        // the barrier of the next post orders this round's mirror stores before anybody reads them
    }

    // ---- write back the resident cache ------------------------------------------------------------------
    __syncthreads();
    if (primary && own) {
        S.cover_t[k] = bl_unsortable(c_mts[bl]); S.cover_code[k] = c_mcode[bl];
        S.obs_t[k] = bl_unsortable(c_ots[bl]);
        S.obs_idx[k] = c_oidx[bl];
        S.obs_arg[k] = c_oarg[bl];
    }
    if (tid == 0 && n_rescans) atomicAdd((unsigned long long *)&res->n_rescans, (unsigned long long)n_rescans);
    if (cta == 0 && tid == 0) {
#pragma unroll
        for (int q = 0; q < 16; q++) res->prof[q] = prof[q];
        res->window = win;
        *S.seq_counter = seq_base + done;
    }
    if (P.sync_mode == SYNC_CLUSTER) cluster_sync_all();
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// CTA shape: Tb balls per CTA, H threads per ball ("Grid").
// `m_max` (collisions per round the shared-memory layout has room for) follows from Tb, except for the largest systems
// (more than 148 x 224 balls), where it is what is left of the shared memory after the per-ball state.
static int pick_shape(const bl_handle *h, int n, int n_obs, int *ncta, int *tb, int *hh, int *sync_mode, int *m_max_out) {
    *hh = 1;
    *m_max_out = 0;  // 0: pick_m_max(tb)
    if (n <= 0) { *ncta = 1; *tb = 32; *sync_mode = SYNC_CTA; return BL_OK; }
    if (n <= 64) {
        *ncta = 1; *sync_mode = SYNC_CTA;
        *tb = 64; *hh = 8;
        return BL_OK;
    }
    if (n <= h->cluster_max) {
        int c = (n + 127) / 128;
        if (c < 2) c = 2;
        int p = 1;
        while (p < c) p <<= 1;  // cluster sizes: powers of two
        c = p;
        int t = (n + c - 1) / c;
        t = ((t + 31) / 32) * 32;
        *ncta = c; *tb = t; *hh = 4; *sync_mode = SYNC_CLUSTER;
        return BL_OK;
    }
    const int sms = h->sm_count > 0 ? h->sm_count : 148;
    // smaller systems: 32 balls per CTA with 16 threads each (measured against 64 x 8: N=100 +14 %, 256 +9 %,
    // 1024 +3 %, 2025 +3 %, 3600 +1 %, 4096 +-0; profiles/r1_cta_shape_scan.log)
    if (n <= h->tb32_max && (n + 31) / 32 <= sms) {
        *ncta = (n + 31) / 32; *tb = 32; *hh = 16; *sync_mode = SYNC_GRID;
        return BL_OK;
    }
    // medium systems: 64 balls per CTA with 8 threads each -- twice the CTAs, half the pair solves per thread
    if ((n + 63) / 64 <= sms) {
        *ncta = (n + 63) / 64; *tb = 64; *hh = 8; *sync_mode = SYNC_GRID;
        return BL_OK;
    }
    // large systems: the fewest balls per CTA that one CTA per SM can hold, as many threads per ball and collisions per
    // round as its state, candidates, undo copy and parked solutions leave room for in shared memory
    const size_t limit = h->max_smem_optin;
    for (int t = 128; t <= BL_MAX_THREADS; t += 32) {
        if ((n + t - 1) / t > sms) continue;
        for (int hq = t <= 128 ? 4 : (t <= 256 ? 2 : 1); hq >= 1; hq >>= 1) {
            for (int mm = pick_m_max(t); mm >= BL_M_MIN_LARGE; mm--) {
                if ((size_t)make_layout(t, hq, mm, n_obs).words * 8 > limit) continue;
                *ncta = (n + t - 1) / t; *tb = t; *hh = hq; *sync_mode = SYNC_GRID;
                *m_max_out = mm == pick_m_max(t) ? 0 : mm;
                return BL_OK;
            }
        }
        break;  // more balls per CTA need more, not less
    }
    return BL_EUNSUPPORTED;
}

int bl_evolve_pick_config(const bl_handle *h, int n, int *ncta, int *threads, int *sync_mode) {
    int tb, hh, mm;
    const int rc = pick_shape(h, n, 4, ncta, &tb, &hh, sync_mode, &mm);
    *threads = tb * hh;
    return rc;
}

template <bool FMA, int MAXT, int TB, int HH, bool KONE = false>
static int launch_t(bl_handle *h, Params &P, size_t smem, cudaStream_t st) {
    auto kern = bl_evolve_kernel<FMA, MAXT, TB, HH, KONE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return bl_check_cuda(h, e, "cudaFuncSetAttribute(smem)");
    const unsigned threads = (unsigned)(P.Tb * P.H);
    if (P.sync_mode == SYNC_GRID) {
        void *args[1] = {(void *)&P};
        e = cudaLaunchCooperativeKernel((const void *)kern, dim3((unsigned)P.ncta), dim3(threads), args, smem, st);
        return bl_check_cuda(h, e, "cooperative launch of bl_evolve_kernel");
    }
    if (P.ncta > 8) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return bl_check_cuda(h, e, "cudaFuncSetAttribute(non-portable cluster)");
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)P.ncta, 1, 1);
    cfg.blockDim = dim3(threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)P.ncta;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, P);
    return bl_check_cuda(h, e, "launch bl_evolve_kernel");
}

int bl_launch_evolve(bl_handle *h, const bl_system *s, double time, double end_time, long long max_events,
                     int first_kind, const bl_log *log, int query_only, cudaStream_t st) {
    if (s->n_obstacles <= BL_SMALL_MAX_OBS && !h->small_kernel_off &&
        (s->n <= BL_SMALL_MAX_BALLS || s->n <= h->cta_single_max))
        return bl_launch_evolve_small(h, s, time, end_time, max_events, first_kind, log, query_only, st);
    int ncta, tb, hh, sync_mode, mm_fit;
    int rc = pick_shape(h, s->n, s->n_obstacles, &ncta, &tb, &hh, &sync_mode, &mm_fit);
    if (rc != BL_OK)
        return bl_set_error(h, rc, "the event kernel keeps ball state resident on chip: the shared memory of %d SMs "
                            "(%zu B each) does not hold n=%d balls with %d obstacles (limit: about %d balls)", h->sm_count,
                            h->max_smem_optin, s->n, s->n_obstacles, h->sm_count * 448);
    Params P;
    P.s = *s;
    P.time = time; P.end_time = end_time; P.max_events = max_events; P.first_kind = first_kind;
    P.query_only = query_only;
    if (log) P.log = *log; else { P.log.cap = 0; P.log.t = nullptr; P.log.i = nullptr; P.log.j = nullptr; P.log.payload = nullptr; }
    P.result = h->d_result;
    P.Tb = tb; P.H = hh; P.ncta = ncta; P.sync_mode = sync_mode;
    P.m_max = mm_fit ? mm_fit : pick_m_max(tb);
    // Keys per round the window aims at.  The chance that a round is cut short grows like m^2 / n, the synchronisation
    // cost per collision falls like 1 / m; measured optimum on B200 (profiles/r1_target_scan.log): N = 256: 12-14,
    // 1024: 22-28, 4096: 36-40, >= 16384: 56-60, i.e. ~ 2.46 n^0.316.
    int target = (int)(2.46 * pow((double)(s->n > 1 ? s->n : 1), 0.316) + 0.5);
    if (target > BL_SORT_CAP - 8) target = BL_SORT_CAP - 8;
    if (target > 2 * P.m_max) target = 2 * P.m_max;
    if (target < 1) target = 1;
    if (h->target_override > 0) target = h->target_override < BL_SORT_CAP - 4 ? h->target_override : BL_SORT_CAP - 4;
    P.target = target;
    // the window length found by the previous launch on this system is a good first guess
    P.window = (h->window_n == s->n && h->window_sys == (const void *)s->toi) ? h->window : 0.0;
    P.bar = h->d_bar; P.post = h->d_post;
    P.stale_list = h->d_stale; P.partial = h->d_partial;
    P.repair_horizon = h->repair_horizon;
    for (int q = 0; q < 5; q++) P.wc[q] = h->winctl[q];
    size_t smem = (size_t)make_layout(P.Tb, P.H, P.m_max, s->n_obstacles).words * 8;
    if (smem > h->max_smem_optin)
        return bl_set_error(h, BL_EUNSUPPORTED, "event kernel needs %zu B of shared memory per CTA (limit %zu)", smem,
                            h->max_smem_optin);
    cudaError_t e = cudaMemsetAsync(h->d_result, 0, sizeof(bl_result), st);
    if (e != cudaSuccess) return bl_check_cuda(h, e, "memset result");
    e = cudaMemsetAsync(h->d_bar, 0, 64, st);
    if (e != cudaSuccess) return bl_check_cuda(h, e, "memset barrier");
    h->launches++;
    const bool kone = s->n_classes == 1;
    if (tb == 128 && hh == 4 && !mm_fit) {
        if (kone) return h->dot_fma ? launch_t<true, 512, 128, 4, true>(h, P, smem, st) : launch_t<false, 512, 128, 4, true>(h, P, smem, st);
        return h->dot_fma ? launch_t<true, 512, 128, 4>(h, P, smem, st) : launch_t<false, 512, 128, 4>(h, P, smem, st);
    }
    if (tb == 32 && hh == 16 && !mm_fit) {
        if (kone) return h->dot_fma ? launch_t<true, 512, 32, 16, true>(h, P, smem, st) : launch_t<false, 512, 32, 16, true>(h, P, smem, st);
        return h->dot_fma ? launch_t<true, 512, 32, 16>(h, P, smem, st) : launch_t<false, 512, 32, 16>(h, P, smem, st);
    }
    if (tb == 64 && hh == 8 && !mm_fit) {
        if (kone) return h->dot_fma ? launch_t<true, 512, 64, 8, true>(h, P, smem, st) : launch_t<false, 512, 64, 8, true>(h, P, smem, st);
        return h->dot_fma ? launch_t<true, 512, 64, 8>(h, P, smem, st) : launch_t<false, 512, 64, 8>(h, P, smem, st);
    }
    if (tb * hh <= 512)
        return h->dot_fma ? launch_t<true, 512, 0, 0>(h, P, smem, st) : launch_t<false, 512, 0, 0>(h, P, smem, st);
    return h->dot_fma ? launch_t<true, BL_MAX_THREADS, 0, 0>(h, P, smem, st)
                      : launch_t<false, BL_MAX_THREADS, 0, 0>(h, P, smem, st);
}
