// bl_small.cuh -- launch parameters shared by the kernels for small and mid-size systems (bl_evolve_cta.cu: 33 ... 224 balls) and, at most 32 balls:
// bl_evolve_small.cu (one group of W lanes per system) and bl_evolve_tiny.cu (one thread per system, n <= 4).
#pragma once
#include "bl_arith.cuh"
#include "bl_internal.h"

namespace bl_small {

enum { KIND_BOTH = 0, KIND_BB = 1, KIND_BO = 2 };

#define BL_SMALL_WARPS 4  // warps per CTA in ensemble launches

struct SmallParams {
    long long n_sys;
    int n, K, n_obs;
    int toi_pitch;             // row pitch of one system's table in doubles
    long long toi_sys_stride;  // doubles between the tables of consecutive systems
    double *p0x, *p0y, *vx, *vy, *t0;  // ball i of system s at [s*n + i]
    const double *radius, *mass;       // [n]
    const int32_t *rclass;             // [n]
    const double *rsum2, *rsum2_obs;
    const bl_obstacle *obstacles;
    double *toi;
    double *obs_t, *obs_arg;
    int32_t *obs_idx;
    double end_time;
    long long max_events;
    // single system only
    double time;
    int first_kind, query_only;
    bl_log log;
    bl_result *result;
    double *cover_t;
    uint64_t *cover_code;
    int64_t *seq_counter;
    // ensembles only
    // per-system composition (bl_multi.per_system): radius / mass / rclass of system s start at [s * ball_stride],
    // its (r_a + r_b)**2 tables at rsum2 + s * rs_stride and rsum2_obs + s * rso_stride; 0 = one set shared by all
    long long ball_stride, rs_stride, rso_stride;
    int build_tables;          // solve every pair at the system's current time first (what add_ball does row by row)
    double *sys_time;
    int64_t *n_bb, *n_bo;
    int32_t *status;
};


}  // namespace bl_small

// bl_evolve_cta.cu: one CTA per system, 33 .. BL_CTA_MAX_BALLS balls
cudaError_t bl_launch_cta(const bl_small::SmallParams &P, bool fma, bool single, cudaStream_t st);
size_t bl_cta_smem_bytes(int n);

// bl_evolve_tiny.cu
cudaError_t bl_launch_tiny(const bl_small::SmallParams &P, bool fma, bool single, int sm_count, cudaStream_t st);
