/*
 * billiards_oracle.c -- CPU restatement of python-billiards' event-driven hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (python-billiards_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file bit-for-bit
 * against event logs and end states produced by the live reference (imported from
 * /root/reference/src by tests/golden/make_golden.py) and against the known-answer
 * values of the reference's own tests and doctests.
 *
 * What it restates (reference file:line, relative to /root/reference/src/billiards):
 *   toi_ball_ball            physics.py:33-76
 *   elastic_collision        physics.py:273-282
 *   InfiniteWall.detect/resolve  obstacles.py:112-144
 *   LineSegment.detect/resolve   obstacles.py:172-198, physics.toi_ball_point :79-144,
 *                                physics.toi_and_param_ball_segment :147-252
 *   Disk.detect/resolve      obstacles.py:70-76
 *   Billiard.add_ball        simulation.py:141-224
 *   Billiard.recompute_toi   simulation.py:226-309
 *   Billiard._detect_*       simulation.py:311-352
 *   Billiard.evolve          simulation.py:354-438
 *   Billiard.bounce_ball_ball / bounce_ball_obstacle   simulation.py:440-554
 *   Billiard._move / _resolve_*   simulation.py:556-636
 *
 * Arithmetic contract: every + - * / is one IEEE binary64
 * operation (compile with -ffp-contract=off); the 2-vector dot product is
 * fma(a1,b1, a0*b0) when dot_fma != 0 (what OpenBLAS ddot does on FMA hosts) and
 * (a0*b0)+(a1*b1) otherwise; (r1+r2)**2 goes through libm pow() exactly like
 * CPython's float.__pow__ (compile with -fno-builtin so gcc does not fold it to s*s).
 *
 * The algorithm is deliberately the reference's: a lower-triangular table of absolute
 * impact times, a full re-argmin of every row >= the smaller colliding index after each
 * event, positions of ALL balls advanced at every event.  (ORC_FAST_ROWMIN selects an
 * incremental row-minimum repair with identical results, for long parity runs.)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_INF (INFINITY)
#define ORC_OBS_WALL 0
#define ORC_OBS_DISK 1
#define ORC_OBS_SEGMENT 2

#define ORC_OK 0
#define ORC_ESEPARATING (-3) /* physics.py:277-279 ValueError */
#define ORC_EDIVZERO (-4)    /* 0-mass vs 0-mass: division by zero in the impulse */
#define ORC_EINVAL (-2)

#define ORC_FAST_ROWMIN 1

typedef struct {
    int kind;
    double a, b, c, d; /* wall: start x,y, normal x,y ; disk: centre x,y, radius, unused ; segment: start, covector */
    double e, f, g, h; /* segment: unit normal x,y, end point x,y (obstacles.py:150-170) */
} orc_obstacle;

typedef struct {
    int n, cap;
    int dot_fma;
    int flags;
    double time;
    double *t0, *p0, *v, *pos, *r, *m; /* p0, v, pos: [n][2] */
    double **row;                      /* toi_table: row[i] has i entries */
    double *balls_toi;                 /* _balls_toi */
    int64_t *balls_idx;                /* _balls_idx */
    double *obs_toi;                   /* _obstacles_toi */
    int *obs_obs;                      /* index into obstacles, -1 = None */
    double *obs_arg;                   /* headway for walls */
    int nobs, obs_cap;
    orc_obstacle *obs;
    double bb_t; int64_t bb_i, bb_j;   /* _next_ball_ball_collision = (t, i, j), i<j */
    double bo_t; int64_t bo_ball; int bo_obs; double bo_arg;
    /* optional event log */
    int64_t log_cap, log_len;
    double *log_t; int64_t *log_i, *log_j;
    /* optional per-event payload log (for callback parity): pos/v_before/v_after of both */
    double *log_payload; /* [log_cap][12] */
    long long solves;    /* number of toi_ball_ball calls, for flop accounting */
} orc_billiard;

/* ---------------------------------------------------------------- arithmetic */

static inline double dot2(const orc_billiard *o, double a0, double a1, double b0, double b1) {
    double p = a0 * b0;
    if (o->dot_fma) return fma(a1, b1, p);
    return p + a1 * b1;
}

/* CPython float_pow(x, 2.0) ends in libm pow(x, 2.0) for finite non-special x */
static double (*volatile orc_pow)(double, double) = pow;
static inline double pow2(double s) { return orc_pow(s, 2.0); }

/* physics.py:33-76 */
static double toi_ball_ball(orc_billiard *o, double p1x, double p1y, double v1x, double v1y, double r1,
                            double p2x, double p2y, double v2x, double v2y, double r2) {
    o->solves++;
    double dpx = p2x - p1x, dpy = p2y - p1y;
    double dvx = v2x - v1x, dvy = v2y - v1y;
    double b = dot2(o, dpx, dpy, dvx, dvy);
    if (b >= 0) return ORC_INF;
    double dist = dot2(o, dpx, dpy, dpx, dpy);
    double speed = dot2(o, dvx, dvy, dvx, dvy);
    double c = dist - pow2(r1 + r2);
    double disc = b * b - speed * c;
    if (disc <= 0) return ORC_INF;
    double t1 = c / (-b + sqrt(disc));
    return t1 >= -1e-10 ? t1 : ORC_INF;
}

/* physics.py:273-282; returns status, writes new velocities */
static int elastic_collision(orc_billiard *o, const double *p1, const double *v1, double m1,
                             const double *p2, const double *v2, double m2, double *w1, double *w2) {
    double pdx = p2[0] - p1[0], pdy = p2[1] - p1[1];
    double vdx = v2[0] - v1[0], vdy = v2[1] - v1[1];
    double pdv = dot2(o, pdx, pdy, vdx, vdy);
    if (pdv > 1e-15) return ORC_ESEPARATING;
    double den = (m1 + m2) * dot2(o, pdx, pdy, pdx, pdy);
    if (den == 0.0) return ORC_EDIVZERO;
    double ix = (2 * (pdv * pdx)) / den;
    double iy = (2 * (pdv * pdy)) / den;
    w1[0] = v1[0] + m2 * ix; w1[1] = v1[1] + m2 * iy;
    w2[0] = v2[0] - m1 * ix; w2[1] = v2[1] - m1 * iy;
    return ORC_OK;
}

/* physics.py:100-144: a moving ball against a static point */
static double toi_ball_point(orc_billiard *o, const double *pos, const double *vel, double radius, double qx, double qy,
                             double t_eps) {
    double dpx = pos[0] - qx, dpy = pos[1] - qy;
    double b = dot2(o, dpx, dpy, vel[0], vel[1]);
    if (b >= 0) return ORC_INF;
    double dist = dot2(o, dpx, dpy, dpx, dpy);
    double speed = dot2(o, vel[0], vel[1], vel[0], vel[1]);
    double c = dist - pow2(radius);
    double disc = b * b - speed * c;
    if (disc <= 0) return ORC_INF;
    double t1 = c / (-b + sqrt(disc));
    return t1 >= t_eps ? t1 : ORC_INF;
}

/* physics.py:213-252: a moving ball against an open segment; *code = 0: *u is the line parameter of the hit,
 * 1: "try the start point", 2: "try the end point", 3: None */
static double toi_param_ball_segment(orc_billiard *o, const double *pos, const double *vel, double radius,
                                     const orc_obstacle *sg, double t_eps, double *u, int *code) {
    double dpx = (pos[0] - sg->a) + t_eps * vel[0], dpy = (pos[1] - sg->b) + t_eps * vel[1];
    double dl = dot2(o, sg->c, sg->d, dpx, dpy);
    double dn = dot2(o, sg->e, sg->f, dpx, dpy);
    *u = 0.0;
    if (fabs(dn) <= radius) {
        *code = dl < 0 ? 1 : (dl > 1 ? 2 : 3);
        return ORC_INF;
    }
    double vn = dot2(o, sg->e, sg->f, vel[0], vel[1]);
    if (vn == 0) { *code = 3; return ORC_INF; }
    double t = -(dn + (dn > 0 ? -radius : radius)) / vn;
    if (t < 0) { *code = 3; return ORC_INF; }
    double uu = dl + t * dot2(o, sg->c, sg->d, vel[0], vel[1]);
    if (0 <= uu && uu <= 1) { *u = uu; *code = 0; return t + t_eps; }
    *code = uu < 0 ? 1 : 2;
    return ORC_INF;
}

/* obstacles.py:112-133 (wall), :70-72 (disk), :172-183 (segment); returns relative time, *arg = headway / u */
static double obstacle_detect(orc_billiard *o, const orc_obstacle *ob, const double *pos, const double *vel,
                              double radius, double *arg) {
    *arg = 0.0;
    if (ob->kind == ORC_OBS_WALL) {
        double headway = -dot2(o, vel[0], vel[1], ob->c, ob->d);
        if (headway <= 0) return ORC_INF;
        double gap = dot2(o, pos[0] - ob->a, pos[1] - ob->b, ob->c, ob->d) - radius;
        double t = gap / headway;
        if (t < -1e-10) return ORC_INF;
        *arg = headway;
        return t;
    }
    if (ob->kind == ORC_OBS_SEGMENT) {
        double u;
        int code;
        double t = toi_param_ball_segment(o, pos, vel, radius, ob, -1e-10, &u, &code);
        if (isinf(t)) {
            if (code == 1) { *arg = 0.0; return toi_ball_point(o, pos, vel, radius, ob->a, ob->b, -1e-10); }
            if (code == 2) { *arg = 1.0; return toi_ball_point(o, pos, vel, radius, ob->g, ob->h, -1e-10); }
            return t;
        }
        *arg = u;
        return t;
    }
    /* Disk: toi_ball_ball(center, (0,0), R, pos, vel, radius) */
    return toi_ball_ball(o, ob->a, ob->b, 0.0, 0.0, ob->c, pos[0], pos[1], vel[0], vel[1], radius);
}

/* obstacles.py:135-144 (wall) and :74-76 (disk) */
static int obstacle_resolve(orc_billiard *o, const orc_obstacle *ob, const double *pos, const double *vel,
                            double arg, double *w) {
    if (ob->kind == ORC_OBS_WALL) {
        w[0] = vel[0] + 2 * (arg * ob->c);
        w[1] = vel[1] + 2 * (arg * ob->d);
        return ORC_OK;
    }
    double c[2] = {ob->a, ob->b}, z[2] = {0.0, 0.0}, w1[2];
    if (ob->kind == ORC_OBS_SEGMENT) {
        /* obstacles.py:185-198: u == 0 / u == 1 -> elastic collision with that end point, else mirror at the line */
        if (arg == 0.0) return elastic_collision(o, c, z, 1.0, pos, vel, 0.0, w1, w);
        if (arg == 1.0) { c[0] = ob->g; c[1] = ob->h; return elastic_collision(o, c, z, 1.0, pos, vel, 0.0, w1, w); }
        double d2 = 2 * dot2(o, ob->e, ob->f, vel[0], vel[1]);
        w[0] = vel[0] - d2 * ob->e;
        w[1] = vel[1] - d2 * ob->f;
        return ORC_OK;
    }
    return elastic_collision(o, c, z, 1.0, pos, vel, 0.0, w1, w);
}

/* ---------------------------------------------------------------- container */

orc_billiard *orc_create(int dot_fma, int flags) {
    orc_billiard *o = (orc_billiard *)calloc(1, sizeof(orc_billiard));
    o->dot_fma = dot_fma;
    o->flags = flags;
    o->bb_t = ORC_INF; o->bb_i = -1; o->bb_j = 0;
    o->bo_t = ORC_INF; o->bo_ball = -1; o->bo_obs = -1;
    return o;
}

void orc_destroy(orc_billiard *o) {
    if (!o) return;
    for (int i = 0; i < o->n; i++) free(o->row[i]);
    free(o->row); free(o->t0); free(o->p0); free(o->v); free(o->pos); free(o->r); free(o->m);
    free(o->balls_toi); free(o->balls_idx); free(o->obs_toi); free(o->obs_obs); free(o->obs_arg);
    free(o->obs); free(o->log_t); free(o->log_i); free(o->log_j); free(o->log_payload);
    free(o);
}

static int add_obstacle8(orc_billiard *o, orc_obstacle ob) {
    if (o->n > 0) return ORC_EINVAL; /* obstacles are fixed at construction, simulation.py:96-103 */
    if (o->nobs == o->obs_cap) {
        o->obs_cap = o->obs_cap ? 2 * o->obs_cap : 8;
        o->obs = (orc_obstacle *)realloc(o->obs, sizeof(orc_obstacle) * o->obs_cap);
    }
    o->obs[o->nobs] = ob;
    return o->nobs++;
}
int orc_add_segment(orc_billiard *o, const double *sg /* start, covector, normal, end */) {
    orc_obstacle ob = {ORC_OBS_SEGMENT, sg[0], sg[1], sg[2], sg[3], sg[4], sg[5], sg[6], sg[7]};
    return add_obstacle8(o, ob);
}
static int add_obstacle(orc_billiard *o, int kind, double a, double b, double c, double d) {
    if (o->n > 0) return ORC_EINVAL; /* obstacles are fixed at construction, simulation.py:96-103 */
    if (o->nobs == o->obs_cap) {
        o->obs_cap = o->obs_cap ? 2 * o->obs_cap : 8;
        o->obs = (orc_obstacle *)realloc(o->obs, sizeof(orc_obstacle) * o->obs_cap);
    }
    orc_obstacle ob = {kind, a, b, c, d, 0.0, 0.0, 0.0, 0.0};
    o->obs[o->nobs] = ob;
    return o->nobs++;
}
int orc_add_wall(orc_billiard *o, double sx, double sy, double nx, double ny) {
    return add_obstacle(o, ORC_OBS_WALL, sx, sy, nx, ny);
}
int orc_add_disk(orc_billiard *o, double cx, double cy, double radius) {
    return add_obstacle(o, ORC_OBS_DISK, cx, cy, radius, 0.0);
}

void orc_enable_log(orc_billiard *o, long long cap, int payload) {
    free(o->log_t); free(o->log_i); free(o->log_j); free(o->log_payload);
    o->log_cap = cap; o->log_len = 0;
    o->log_t = (double *)malloc(sizeof(double) * (cap ? cap : 1));
    o->log_i = (int64_t *)malloc(sizeof(int64_t) * (cap ? cap : 1));
    o->log_j = (int64_t *)malloc(sizeof(int64_t) * (cap ? cap : 1));
    o->log_payload = payload ? (double *)calloc((size_t)(cap ? cap : 1) * 12, sizeof(double)) : NULL;
}

/* simulation.py:556-560 */
static void move_all(orc_billiard *o, double time) {
    for (int k = 0; k < o->n; k++) {
        double dt = time - o->t0[k];
        o->pos[2 * k] = o->p0[2 * k] + o->v[2 * k] * dt;
        o->pos[2 * k + 1] = o->p0[2 * k + 1] + o->v[2 * k + 1] * dt;
    }
    o->time = time;
}

/* simulation.py:311-330 */
static double detect_ball(orc_billiard *o, int i, int j) {
    return o->time + toi_ball_ball(o, o->pos[2 * i], o->pos[2 * i + 1], o->v[2 * i], o->v[2 * i + 1], o->r[i],
                                   o->pos[2 * j], o->pos[2 * j + 1], o->v[2 * j], o->v[2 * j + 1], o->r[j]);
}

/* simulation.py:332-352: obstacles in list order, strict < keeps the first */
static void detect_next_obstacle(orc_billiard *o, int idx) {
    double t_min = ORC_INF, arg_min = 0.0;
    int obs_min = -1;
    for (int q = 0; q < o->nobs; q++) {
        double arg;
        double t = obstacle_detect(o, &o->obs[q], &o->pos[2 * idx], &o->v[2 * idx], o->r[idx], &arg);
        if (t < t_min) { t_min = t; obs_min = q; arg_min = arg; }
    }
    o->obs_toi[idx] = o->time + t_min;
    o->obs_obs[idx] = obs_min;
    o->obs_arg[idx] = arg_min;
}

/* numpy argmin: first index attaining the minimum (all-INF row -> 0) */
static int64_t argmin_first(const double *x, int64_t n) {
    int64_t best = 0;
    for (int64_t k = 1; k < n; k++)
        if (x[k] < x[best]) best = k;
    return best;
}

static void row_min(orc_billiard *o, int i) {
    if (i == 0) { o->balls_toi[0] = ORC_INF; o->balls_idx[0] = -1; return; }
    int64_t k = argmin_first(o->row[i], i);
    o->balls_toi[i] = o->row[i][k];
    o->balls_idx[i] = k;
}

static void global_minima(orc_billiard *o) {
    if (o->n == 0) return;
    int64_t nx = argmin_first(o->balls_toi, o->n);
    o->bb_t = o->balls_toi[nx]; o->bb_i = o->balls_idx[nx]; o->bb_j = nx;
    int64_t bx = argmin_first(o->obs_toi, o->n);
    o->bo_t = o->obs_toi[bx]; o->bo_ball = bx; o->bo_obs = o->obs_obs[bx]; o->bo_arg = o->obs_arg[bx];
}

static void grow(orc_billiard *o) {
    if (o->n < o->cap) return;
    int cap = o->cap ? 2 * o->cap : 16;
    o->t0 = (double *)realloc(o->t0, sizeof(double) * cap);
    o->p0 = (double *)realloc(o->p0, sizeof(double) * 2 * cap);
    o->v = (double *)realloc(o->v, sizeof(double) * 2 * cap);
    o->pos = (double *)realloc(o->pos, sizeof(double) * 2 * cap);
    o->r = (double *)realloc(o->r, sizeof(double) * cap);
    o->m = (double *)realloc(o->m, sizeof(double) * cap);
    o->row = (double **)realloc(o->row, sizeof(double *) * cap);
    o->balls_toi = (double *)realloc(o->balls_toi, sizeof(double) * cap);
    o->balls_idx = (int64_t *)realloc(o->balls_idx, sizeof(int64_t) * cap);
    o->obs_toi = (double *)realloc(o->obs_toi, sizeof(double) * cap);
    o->obs_obs = (int *)realloc(o->obs_obs, sizeof(int) * cap);
    o->obs_arg = (double *)realloc(o->obs_arg, sizeof(double) * cap);
    o->cap = cap;
}

/* simulation.py:141-224 */
int orc_add_ball(orc_billiard *o, double px, double py, double vx, double vy, double radius, double mass) {
    grow(o);
    int idx = o->n;
    o->t0[idx] = o->time;
    o->p0[2 * idx] = px; o->p0[2 * idx + 1] = py;
    o->v[2 * idx] = vx; o->v[2 * idx + 1] = vy;
    o->r[idx] = radius; o->m[idx] = mass;
    o->row[idx] = (double *)malloc(sizeof(double) * (idx ? idx : 1));
    o->n = idx + 1;
    move_all(o, o->time);
    for (int j = 0; j < idx; j++) o->row[idx][j] = detect_ball(o, j, idx);
    row_min(o, idx);
    detect_next_obstacle(o, idx);
    global_minima(o);
    return idx;
}

/* after rows/columns of ball a (and b) changed: refresh row minima (simulation.py:474-478) */
static void refresh_rows_from(orc_billiard *o, int first, int a, int b, const double *old_a, const double *old_b) {
    if (!(o->flags & ORC_FAST_ROWMIN) || old_a == NULL) {
        for (int i = first > 0 ? first : 1; i < o->n; i++) row_min(o, i);
        return;
    }
    /* incremental repair with identical results: row a (and b) fully rescanned; a row i > a only
       had entries [i][a] (and [i][b]) rewritten */
    if (a > 0) row_min(o, a);
    if (b > 0) row_min(o, b);
    for (int i = a + 1; i < o->n; i++) {
        if (i == b) continue;
        int touched_min = (o->balls_idx[i] == a) || (b >= 0 && b < i && o->balls_idx[i] == b);
        if (touched_min) { row_min(o, i); continue; }
        double na = o->row[i][a];
        if (na < o->balls_toi[i] || (na == o->balls_toi[i] && a < o->balls_idx[i])) {
            o->balls_toi[i] = na; o->balls_idx[i] = a;
        }
        if (b >= 0 && b < i) {
            double nb = o->row[i][b];
            if (nb < o->balls_toi[i] || (nb == o->balls_toi[i] && b < o->balls_idx[i])) {
                o->balls_toi[i] = nb; o->balls_idx[i] = b;
            }
        }
    }
    (void)old_b;
}

static void log_event(orc_billiard *o, double t, int64_t i, int64_t j, const double *payload) {
    if (o->log_len < o->log_cap) {
        o->log_t[o->log_len] = t; o->log_i[o->log_len] = i; o->log_j[o->log_len] = j;
        if (o->log_payload && payload) memcpy(&o->log_payload[12 * o->log_len], payload, 12 * sizeof(double));
    }
    if (o->log_cap) o->log_len++;
}

/* simulation.py:440-502 */
int orc_bounce_ball_ball(orc_billiard *o) {
    double t = o->bb_t;
    int a = (int)o->bb_i, b = (int)o->bb_j;
    if (a < 0 || !(a < b)) return ORC_EINVAL;
    move_all(o, t);
    /* _resolve_ball_collision, simulation.py:562-602 */
    double m1 = o->m[a], m2 = o->m[b];
    if (isinf(m1)) { m1 = 1; m2 = 0; }
    else if (isinf(m2)) { m1 = 0; m2 = 1; }
    double w1[2], w2[2];
    int st = elastic_collision(o, &o->pos[2 * a], &o->v[2 * a], m1, &o->pos[2 * b], &o->v[2 * b], m2, w1, w2);
    if (st != ORC_OK) return st;
    double payload[12] = {o->pos[2 * a], o->pos[2 * a + 1], o->v[2 * a], o->v[2 * a + 1], w1[0], w1[1],
                          o->pos[2 * b], o->pos[2 * b + 1], o->v[2 * b], o->v[2 * b + 1], w2[0], w2[1]};
    log_event(o, t, a, b, payload);
    o->t0[a] = t; o->t0[b] = t;
    o->p0[2 * a] = o->pos[2 * a]; o->p0[2 * a + 1] = o->pos[2 * a + 1];
    o->p0[2 * b] = o->pos[2 * b]; o->p0[2 * b + 1] = o->pos[2 * b + 1];
    o->v[2 * a] = w1[0]; o->v[2 * a + 1] = w1[1];
    o->v[2 * b] = w2[0]; o->v[2 * b + 1] = w2[1];

    o->row[b][a] = ORC_INF;
    for (int j = 0; j < a; j++) {
        o->row[a][j] = detect_ball(o, a, j);
        o->row[b][j] = detect_ball(o, b, j);
    }
    for (int i = a + 1; i < b; i++) {
        o->row[i][a] = detect_ball(o, i, a);
        o->row[b][i] = detect_ball(o, b, i);
    }
    for (int i = b + 1; i < o->n; i++) {
        o->row[i][a] = detect_ball(o, i, a);
        o->row[i][b] = detect_ball(o, i, b);
    }
    refresh_rows_from(o, a, a, b, o->balls_toi, o->balls_toi);
    detect_next_obstacle(o, a);
    detect_next_obstacle(o, b);
    global_minima(o);
    return ORC_OK;
}

/* simulation.py:504-554 */
int orc_bounce_ball_obstacle(orc_billiard *o) {
    double t = o->bo_t;
    int idx = (int)o->bo_ball, q = o->bo_obs;
    if (idx < 0 || q < 0) return ORC_EINVAL;
    move_all(o, t);
    double w[2];
    int st = obstacle_resolve(o, &o->obs[q], &o->pos[2 * idx], &o->v[2 * idx], o->bo_arg, w);
    if (st != ORC_OK) return st;
    double payload[12] = {o->pos[2 * idx], o->pos[2 * idx + 1], o->v[2 * idx], o->v[2 * idx + 1], w[0], w[1],
                          o->bo_arg, 0, 0, 0, 0, 0};
    log_event(o, t, idx, -1 - (int64_t)q, payload);
    o->t0[idx] = t;
    o->p0[2 * idx] = o->pos[2 * idx]; o->p0[2 * idx + 1] = o->pos[2 * idx + 1];
    o->v[2 * idx] = w[0]; o->v[2 * idx + 1] = w[1];

    for (int j = 0; j < idx; j++) o->row[idx][j] = detect_ball(o, idx, j);
    for (int i = idx + 1; i < o->n; i++) o->row[i][idx] = detect_ball(o, i, idx);
    refresh_rows_from(o, idx, idx, -1, o->balls_toi, NULL);
    detect_next_obstacle(o, idx);
    global_minima(o);
    return ORC_OK;
}

/* simulation.py:354-438; max_events < 0 = unlimited (count-based stop is an extension used
   by the ensemble workloads: "repeated bounce_*()" in the reference's terms) */
int orc_evolve(orc_billiard *o, double end_time, long long max_events, long long counts[2]) {
    counts[0] = counts[1] = 0;
    for (;;) {
        double tn = o->bb_t <= o->bo_t ? o->bb_t : o->bo_t;
        if (!(tn <= end_time)) break;
        if (max_events >= 0 && counts[0] + counts[1] >= max_events) return ORC_OK; /* no final move */
        int st;
        if (o->bb_t <= o->bo_t) { st = orc_bounce_ball_ball(o); if (st) return st; counts[0]++; }
        else { st = orc_bounce_ball_obstacle(o); if (st) return st; counts[1]++; }
    }
    move_all(o, end_time);
    return ORC_OK;
}

/* In-place edits of balls_position / balls_velocity / balls_radius / balls_mass (the user-facing
   arrays of the reference); they take effect at the next orc_recompute_toi, like in the reference. */
int orc_set_ball(orc_billiard *o, int idx, const double *pos, const double *vel, const double *radius,
                 const double *mass) {
    if (idx < 0 || idx >= o->n) return ORC_EINVAL;
    if (pos) { o->pos[2 * idx] = pos[0]; o->pos[2 * idx + 1] = pos[1]; }
    if (vel) { o->v[2 * idx] = vel[0]; o->v[2 * idx + 1] = vel[1]; }
    if (radius) o->r[idx] = *radius;
    if (mass) o->m[idx] = *mass;
    return ORC_OK;
}

/* simulation.py:226-309 */
int orc_recompute_toi(orc_billiard *o, const int *indices, int k) {
    int min_idx = o->n;
    /* :255-260 -- a ball whose current position no longer equals p0 + v*(time - t0) (because its position
       or its velocity was edited) is re-based at the current time */
    for (int q = 0; q < o->n; q++) {
        double dt = o->time - o->t0[q];
        double ox = o->p0[2 * q] + o->v[2 * q] * dt, oy = o->p0[2 * q + 1] + o->v[2 * q + 1] * dt;
        if (o->pos[2 * q] != ox || o->pos[2 * q + 1] != oy) {
            o->t0[q] = o->time;
            o->p0[2 * q] = o->pos[2 * q]; o->p0[2 * q + 1] = o->pos[2 * q + 1];
        }
    }
    /* recompute_pairs set: an entry is recomputed once even if both balls are listed */
    unsigned char *done = (unsigned char *)calloc((size_t)(o->n ? o->n : 1), 1);
    for (int q = 0; q < k; q++) {
        int idx = indices[q];
        if (idx < 0 || idx >= o->n) { free(done); return ORC_EINVAL; }
        for (int j = 0; j < idx; j++)
            if (!done[j]) o->row[idx][j] = detect_ball(o, idx, j);
        for (int i = idx + 1; i < o->n; i++)
            if (!done[i]) o->row[i][idx] = detect_ball(o, i, idx);
        done[idx] = 1;
        detect_next_obstacle(o, idx);
        if (idx < min_idx) min_idx = idx;
    }
    free(done);
    for (int i = min_idx > 0 ? min_idx : 1; i < o->n; i++) row_min(o, i);
    global_minima(o);
    return ORC_OK;
}

void orc_set_time(orc_billiard *o, double t) { o->time = t; }

/* ---------------------------------------------------------------- read-out */

int orc_count(const orc_billiard *o) { return o->n; }
double orc_time(const orc_billiard *o) { return o->time; }
long long orc_solves(const orc_billiard *o) { return o->solves; }
long long orc_log_len(const orc_billiard *o) { return o->log_len; }
void orc_get_log(const orc_billiard *o, double *t, int64_t *i, int64_t *j, double *payload) {
    long long n = o->log_len < o->log_cap ? o->log_len : o->log_cap;
    memcpy(t, o->log_t, sizeof(double) * n);
    memcpy(i, o->log_i, sizeof(int64_t) * n);
    memcpy(j, o->log_j, sizeof(int64_t) * n);
    if (payload && o->log_payload) memcpy(payload, o->log_payload, sizeof(double) * 12 * n);
}
void orc_clear_log(orc_billiard *o) { o->log_len = 0; }

void orc_get_state(const orc_billiard *o, double *t0, double *p0, double *v, double *pos, double *r, double *m) {
    int n = o->n;
    if (t0) memcpy(t0, o->t0, sizeof(double) * n);
    if (p0) memcpy(p0, o->p0, sizeof(double) * 2 * n);
    if (v) memcpy(v, o->v, sizeof(double) * 2 * n);
    if (pos) memcpy(pos, o->pos, sizeof(double) * 2 * n);
    if (r) memcpy(r, o->r, sizeof(double) * n);
    if (m) memcpy(m, o->m, sizeof(double) * n);
}

/* flat lower-triangular table: row i occupies [i(i-1)/2, i(i+1)/2) */
void orc_get_table(const orc_billiard *o, double *tri) {
    size_t off = 0;
    for (int i = 0; i < o->n; i++) { memcpy(tri + off, o->row[i], sizeof(double) * i); off += i; }
}
void orc_get_minima(const orc_billiard *o, double *balls_toi, int64_t *balls_idx, double *obs_toi, int *obs_obs,
                    double *obs_arg) {
    int n = o->n;
    memcpy(balls_toi, o->balls_toi, sizeof(double) * n);
    memcpy(balls_idx, o->balls_idx, sizeof(int64_t) * n);
    memcpy(obs_toi, o->obs_toi, sizeof(double) * n);
    memcpy(obs_obs, o->obs_obs, sizeof(int) * n);
    memcpy(obs_arg, o->obs_arg, sizeof(double) * n);
}
void orc_get_next(const orc_billiard *o, double *tt /*[2]*/, int64_t *idx /*[4]*/, double *arg) {
    tt[0] = o->bb_t; idx[0] = o->bb_i; idx[1] = o->bb_j;
    tt[1] = o->bo_t; idx[2] = o->bo_ball; idx[3] = o->bo_obs;
    *arg = o->bo_arg;
}

/* ---------------------------------------------------------------- scalar kernels, for unit parity */

double orc_toi_ball_ball(int dot_fma, const double *in /* p1x,p1y,v1x,v1y,r1,p2x,p2y,v2x,v2y,r2 */) {
    orc_billiard o; memset(&o, 0, sizeof(o)); o.dot_fma = dot_fma;
    return toi_ball_ball(&o, in[0], in[1], in[2], in[3], in[4], in[5], in[6], in[7], in[8], in[9]);
}
int orc_elastic_collision(int dot_fma, const double *p1, const double *v1, double m1, const double *p2,
                          const double *v2, double m2, double *w1, double *w2) {
    orc_billiard o; memset(&o, 0, sizeof(o)); o.dot_fma = dot_fma;
    return elastic_collision(&o, p1, v1, m1, p2, v2, m2, w1, w2);
}
double orc_wall_detect(int dot_fma, const double *wall /* sx,sy,nx,ny */, const double *pos, const double *vel,
                       double radius, double *headway) {
    orc_billiard o; memset(&o, 0, sizeof(o)); o.dot_fma = dot_fma;
    orc_obstacle ob = {ORC_OBS_WALL, wall[0], wall[1], wall[2], wall[3], 0.0, 0.0, 0.0, 0.0};
    return obstacle_detect(&o, &ob, pos, vel, radius, headway);
}
double orc_disk_detect(int dot_fma, const double *disk /* cx,cy,R */, const double *pos, const double *vel,
                       double radius) {
    orc_billiard o; memset(&o, 0, sizeof(o)); o.dot_fma = dot_fma;
    orc_obstacle ob = {ORC_OBS_DISK, disk[0], disk[1], disk[2], 0.0, 0.0, 0.0, 0.0, 0.0};
    double arg;
    return obstacle_detect(&o, &ob, pos, vel, radius, &arg);
}

double orc_toi_ball_point(int dot_fma, const double *pos, const double *vel, double radius, const double *point,
                          double t_eps) {
    orc_billiard o; memset(&o, 0, sizeof(o)); o.dot_fma = dot_fma;
    return toi_ball_point(&o, pos, vel, radius, point[0], point[1], t_eps);
}
double orc_segment_param(int dot_fma, const double *sg /* start, covector, normal, end */, const double *pos,
                         const double *vel, double radius, double t_eps, double *u, int *code) {
    orc_billiard o; memset(&o, 0, sizeof(o)); o.dot_fma = dot_fma;
    orc_obstacle ob = {ORC_OBS_SEGMENT, sg[0], sg[1], sg[2], sg[3], sg[4], sg[5], sg[6], sg[7]};
    return toi_param_ball_segment(&o, pos, vel, radius, &ob, t_eps, u, code);
}
double orc_segment_detect(int dot_fma, const double *sg, const double *pos, const double *vel, double radius,
                          double *arg) {
    orc_billiard o; memset(&o, 0, sizeof(o)); o.dot_fma = dot_fma;
    orc_obstacle ob = {ORC_OBS_SEGMENT, sg[0], sg[1], sg[2], sg[3], sg[4], sg[5], sg[6], sg[7]};
    return obstacle_detect(&o, &ob, pos, vel, radius, arg);
}
int orc_segment_resolve(int dot_fma, const double *sg, const double *pos, const double *vel, double arg, double *w) {
    orc_billiard o; memset(&o, 0, sizeof(o)); o.dot_fma = dot_fma;
    orc_obstacle ob = {ORC_OBS_SEGMENT, sg[0], sg[1], sg[2], sg[3], sg[4], sg[5], sg[6], sg[7]};
    return obstacle_resolve(&o, &ob, pos, vel, arg, w);
}

/* ---------------------------------------------------------------- single-particle ensemble (config S)
 * One ball per system, obstacles shared; each system runs n_coll x bounce_ball_obstacle()
 * (simulation.py:504-554 with N = 1) starting from time 0.  state: [n_sys][4] = px,py,vx,vy in,
 * p0x,p0y,vx,vy out (position at the last collision); t_out[n_sys] = time of the last collision;
 * hits[n_sys][nobs] optional per-obstacle hit counters. */
/* status != NULL: a system whose collision cannot be resolved (ORC_ESEPARATING / ORC_EDIVZERO: the reference would
 * raise inside bounce_ball_obstacle, after _move) stops there -- clock at the failed collision, state of record
 * untouched -- and the other systems carry on; status == NULL: the first such error is returned. */
int orc_ensemble_run2(int dot_fma, const orc_obstacle *obs, int nobs, long long n_sys, double radius,
                      double *state, double *t_out, long long n_coll, long long *done, int32_t *hits,
                      int32_t *status) {
    orc_billiard o; memset(&o, 0, sizeof(o)); o.dot_fma = dot_fma;
    for (long long s = 0; s < n_sys; s++) {
        double p0[2] = {state[4 * s], state[4 * s + 1]}, v[2] = {state[4 * s + 2], state[4 * s + 3]};
        double t0 = 0.0, time = 0.0, pos[2] = {p0[0], p0[1]};
        long long c = 0;
        for (; c < n_coll; c++) {
            double t_min = ORC_INF, arg_min = 0.0; int q_min = -1;
            for (int q = 0; q < nobs; q++) {
                double arg, t = obstacle_detect(&o, &obs[q], pos, v, radius, &arg);
                if (t < t_min) { t_min = t; q_min = q; arg_min = arg; }
            }
            double tn = time + t_min;
            if (q_min < 0) break;
            double dt = tn - t0;
            pos[0] = p0[0] + v[0] * dt; pos[1] = p0[1] + v[1] * dt;
            time = tn;
            double w[2];
            int st = obstacle_resolve(&o, &obs[q_min], pos, v, arg_min, w);
            if (st != ORC_OK) {
                if (!status) return st;
                status[s] = st;
                break;
            }
            t0 = time; p0[0] = pos[0]; p0[1] = pos[1]; v[0] = w[0]; v[1] = w[1];
            if (hits) hits[s * nobs + q_min]++;
        }
        state[4 * s] = p0[0]; state[4 * s + 1] = p0[1]; state[4 * s + 2] = v[0]; state[4 * s + 3] = v[1];
        t_out[s] = time;
        if (done) done[s] = c;
    }
    return ORC_OK;
}

/* libm's fma, for the dot-mode probe of oracle.py */
double orc_fma(double a, double b, double c) { return fma(a, b, c); }

int orc_ensemble_run(int dot_fma, const orc_obstacle *obs, int nobs, long long n_sys, double radius,
                     double *state, double *t_out, long long n_coll, long long *done, int32_t *hits) {
    return orc_ensemble_run2(dot_fma, obs, nobs, n_sys, radius, state, t_out, n_coll, done, hits, NULL);
}
