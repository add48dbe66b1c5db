// bl_internal.h -- launcher prototypes shared by the translation units of libbilliards_b200.so
#pragma once
#include <cuda_runtime.h>
#include <pthread.h>

#include "../../include/billiards_b200.h"

struct bl_handle {
    pthread_mutex_t mutex;    // one library call at a time per handle (the handle owns barrier counters, exchange
                              // buffers and ONE result block: two concurrent launches would race on them)
    int device;
    int dot_fma;
    int sm_count;
    size_t max_smem_optin;
    bl_result *d_result;      // device result block
    bl_result *h_result;      // pinned host mirror
    cudaEvent_t ev0, ev1;
    float last_ms;
    long long launches;
    // event kernel exchange buffers (global memory; see bl_evolve.cu)
    unsigned *d_bar;          // grid barrier counter
    uint64_t *d_post;         // [2][BL_MAX_CTAS][BL_POST_WORDS] keys posted by each CTA
    uint64_t *d_stale;        // [BL_MAX_CTAS * BL_MAX_THREADS][2] candidates under repair
    uint64_t *d_partial;      // [2][BL_MAX_CTAS][BL_RCHUNK][2] per-CTA minima of a repair chunk
    double repair_horizon;    // in windows (BL_REPAIR_HORIZON, default 100)
    double winctl[5];         // window control (BL_WINCTL): factor after a violation; growth factor, below this fraction
                              // of the target; shrink factor, above this fraction
    int cluster_max;          // systems up to this many balls launch as one thread-block cluster (BL_CLUSTER_MAX, default 0: never)
    int tb32_max;             // systems of 65 .. this many balls run 32 balls x 16 threads per CTA (BL_TB32_MAX, default 3600: measured break-even near 4096)
    // window length the last launch ended with, and the system it belongs to
    double window;
    int window_n;
    const void *window_sys;
    int target_override;      // > 0: collisions per round to aim at (bl_set_batch_target)
    int small_kernel_off;     // bl_set_small_kernel(h, 0): every system takes the grid-wide event kernel (default: <= 32 balls lane groups, <= 80 one CTA)
    int cta_single_max;       // single systems of 33 .. this many balls take the one-CTA kernel (default BL_CTA_SINGLE_MAX: above, the grid-wide kernel is faster; bl_set_small_kernel(h, 3): all sizes up to 224)
    int small_force_lanes;    // bl_set_small_kernel(h, 2): systems of <= 4 balls take the lane-group kernel, not the one-thread kernel
    int ensemble_no_box;      // bl_set_ensemble_generic(h, 2): shaped kernel, but without the axis-aligned-box shortcut
    int ensemble_generic;     // BL_ENSEMBLE_GENERIC=1: always use the generic (list in shared memory) ensemble kernel
    char err[512];
};

#define BL_MAX_THREADS 1024
#define BL_OBS_WORDS ((int)(sizeof(bl_obstacle) / 8))
#define BL_MAX_CTAS 160       // >= SM count
#define BL_POST_WORDS 16      // header, violation word, BL_R_POST keys of 2 words
#define BL_R_POST 6           // keys one CTA may post per round
#define BL_EV_CAP 256         // keys collected per round (>= number of CTAs: the exact argmin posts one per CTA)
#define BL_SORT_CAP 64        // keys a window round sorts (two per lane of warp 0); more -> the window is halved
#define BL_M_MAX 48           // collisions committed per round, at most
#define BL_M_MIN_LARGE 4      // the largest systems trade collisions per round for balls per CTA, down to this many
#define BL_WQ_CAP 64          // per-warp queue of pair solves that need the square root (two ballots may add 64)
#define BL_SMALL_MAX_BALLS 32  // systems up to this size run in the one-warp kernel (bl_evolve_small.cu)
#define BL_SMALL_MAX_OBS 32
#define BL_CTA_MAX_WARPS 7     // systems of 33 .. 224 balls run one CTA per system, thread = ball (bl_evolve_cta.cu)
#define BL_CTA_MAX_BALLS (32 * BL_CTA_MAX_WARPS)
#define BL_CTA_SINGLE_MAX 80   // measured on B200 (tools/mid_bench.py): N = 36 x1.57, 64 x1.44, 100 x0.85 against the grid kernel
#define BL_TINY_MAX_BALLS 4    // systems up to this size run one thread per system (bl_evolve_tiny.cu)
#define BL_HL_CAP 6           // finite new values of one ball listed per round (more -> that ball scans all of them)
#define BL_RCHUNK 256         // stale candidates repaired per pass
#define BL_STASH_BYTES 98304  // shared memory for the parked pair solutions of a round

// Every extern "C" entry that touches the device starts with `bl_entry_guard g(h);`: it serialises the calls on one
// handle and makes the handle's device current for the duration of the call (a kernel launch on a stream of a
// non-current device fails, and cudaFuncSetAttribute would land on the wrong device), restoring the caller's device
// on the way out.
struct bl_entry_guard {
    bl_handle *h;
    int prev;
    bool switched;
    explicit bl_entry_guard(const bl_handle *handle) : h(const_cast<bl_handle *>(handle)), prev(-1), switched(false) {
        if (!h) return;
        pthread_mutex_lock(&h->mutex);
        if (cudaGetDevice(&prev) == cudaSuccess && prev != h->device) switched = cudaSetDevice(h->device) == cudaSuccess;
    }
    ~bl_entry_guard() {
        if (!h) return;
        if (switched) cudaSetDevice(prev);
        pthread_mutex_unlock(&h->mutex);
    }
    bl_entry_guard(const bl_entry_guard &) = delete;
    bl_entry_guard &operator=(const bl_entry_guard &) = delete;
};

int bl_set_error(bl_handle *h, int code, const char *fmt, ...);
int bl_check_cuda(bl_handle *h, cudaError_t e, const char *what);

// bl_evolve.cu
int bl_launch_evolve(bl_handle *h, const bl_system *s, double time, double end_time, long long max_events,
                     int first_kind, const bl_log *log, int query_only, cudaStream_t st);
int bl_launch_evolve_small(bl_handle *h, const bl_system *s, double time, double end_time, long long max_events,
                           int first_kind, const bl_log *log, int query_only, cudaStream_t st);
int bl_launch_multi(bl_handle *h, const bl_multi *m, int build_tables, double end_time, long long max_events,
                    cudaStream_t st);
int bl_launch_multi_positions(bl_handle *h, const bl_multi *m, double time, int at_system_time, double *d_pos,
                              cudaStream_t st);
int bl_evolve_pick_config(const bl_handle *h, int n, int *ncta, int *threads, int *sync_mode);
// bl_table.cu
int bl_launch_recompute(bl_handle *h, const bl_system *s, double time, const int32_t *d_idx, int first, int k,
                        cudaStream_t st);  // d_idx == nullptr: balls first .. first + k - 1
int bl_launch_rebuild_cover(bl_handle *h, const bl_system *s, cudaStream_t st);
int bl_launch_symmetrize(bl_handle *h, const bl_system *s, cudaStream_t st);
int bl_launch_row_minima(bl_handle *h, const bl_system *s, double *d_toi, int64_t *d_idx, cudaStream_t st);
int bl_launch_apply_edits(bl_handle *h, const bl_system *s, double time, const double *d_pos, const double *d_vel,
                          cudaStream_t st);
int bl_launch_result_to_host(bl_handle *h, const bl_result *src, bl_result *dst_dev, cudaStream_t st);
int bl_launch_snapshot(bl_handle *h, const bl_system *s, double time, const double *d_time, double *d_out, cudaStream_t st,
                       double *out2 = nullptr, const bl_result *res_src = nullptr, bl_result *res_dst = nullptr);
int bl_launch_positions(bl_handle *h, const bl_system *s, double time, double *d_pos, cudaStream_t st);
// bl_batch.cu
int bl_launch_toi_batch(bl_handle *h, const double *d_in, const double *d_rsum2, double t_eps, double *d_out,
                        long long m, cudaStream_t st);
int bl_launch_elastic_batch(bl_handle *h, const double *d_in, double *d_out, int32_t *d_status, long long m,
                            cudaStream_t st);
int bl_launch_obstacle_detect(bl_handle *h, bl_obstacle ob, const double *d_in, const double *d_rsum2, double *d_t,
                              double *d_arg, long long m, cudaStream_t st);
int bl_launch_obstacle_resolve(bl_handle *h, bl_obstacle ob, const double *d_in, const double *d_arg, double *d_out,
                               int32_t *d_status, long long m, cudaStream_t st);
int bl_launch_point_batch(bl_handle *h, const double *d_in, const double *d_r2, double t_eps, double *d_out, long long m,
                          cudaStream_t st);
int bl_launch_segment_param(bl_handle *h, bl_obstacle sg, const double *d_in, double t_eps, double *d_t, double *d_u,
                            int32_t *d_code, long long m, cudaStream_t st);
int bl_launch_fp64_peak(bl_handle *h, double *tflops, cudaStream_t st);
// bl_ensemble.cu
int bl_launch_ensemble(bl_handle *h, const bl_obstacle *obs, int n_obs, const double *rsum2_obs, double radius,
                       long long n_sys, double *d_state, double *d_time, long long n_coll, double end_time,
                       int64_t *d_counts, int32_t *d_hits, int32_t *d_status, cudaStream_t st);
int bl_launch_selftest_div2(bl_handle *h, uint64_t seed, long long n, long long *mismatches, double *first_bad,
                            cudaStream_t st);
int bl_launch_ensemble_stats(bl_handle *h, const int64_t *d_counts, const int32_t *d_hits, int n_obs, long long n_sys,
                             int64_t *d_out, cudaStream_t st);
