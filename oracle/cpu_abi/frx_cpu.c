/*
 * frx_cpu.c -- the C ABI of include/frx.h implemented on the HOST CORES (C + OpenMP) -> oracle/_build/libfrx_cpu_omp*.so
 *
 * TEST / BASELINE INFRASTRUCTURE, not part of the product: SURVEY.md 2.2 / 8b ask for "one C++/OpenMP CPU implementation
 * of the same C ABI (serves as the timed CPU baseline, because frenetix itself is not installable here)".  It wraps
 * orc_plan of oracle/c/frx_oracle.c (the port of the reference's Python path, lazy collision walk like
 * planner.py:329-392) behind the entry points a host binds for the B200 library, so that bench.py's reference arm times
 * the CPU path THROUGH THE SAME CALLS it times the GPU path through.  The product library (libfrx_b200.so) contains no CPU
 * path and the package never loads this file; device-only entry points return FRX_ERR_UNSUPPORTED here.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "frx.h"

/* ---- the port (oracle/c/frx_oracle.c), linked into this library */
typedef struct {
    double dt;
    int32_t N, low, draw, debug;
    double a_max, v_switch, delta_max, wheelbase, wb_rear, length, width, x0_orientation, v_des;
    int32_t n_costs;
    int32_t cost_ids[10];
    double w[10];
    int32_t check_all_collisions;
    int32_t collision_check;
    int32_t kd_from_v_delta;
    int32_t vo_norm;
    double v_delta_max;
} orc_params;
typedef struct {
    int64_t argmin;
    double min_cost;
    int64_t n_in_list, n_feasible, n_candidates, collision_counter;
    int64_t reason_counts[11];
} orc_result;
int orc_plan(const orc_params* P, int64_t n, const double* sampling, int M, const double* ref_pos, const double* ref_theta,
             const double* ref_curv, const double* ref_curv_d, const double* ref_x, const double* ref_y, int nT,
             const double* Tvals, const int32_t* Tlen, const double* tpow, int O, int T, const double* pos, const double* cov,
             const double* otheta, const double* ohl, const double* ohw, const int32_t* olen, int n_obs_pos,
             const double* obs_pos, int B, const double* sobb5, double* states, double* costs, double* total, uint32_t* flags,
             int32_t* traj_len, double* margins, orc_result* res, int nthreads);

struct frx_ctx {
    char err[256];
    frx_params prm;
    int have_params, have_ref, have_tables;
    int M; double* ref[6];
    int nT; double* Tvals; int32_t* Tlen; double* tpow;
    int O, T; double *pos, *cov, *theta, *hl, *hw; int32_t* olen;
    int n_obs_pos; double* obs_pos;
    int B; double* sobb;
    int64_t N, cap; int K, Nt;
    double *states, *costs, *total; uint32_t* flags; int32_t* traj_len;
    double* grid_rows; int64_t grid_cap;
    int64_t winner;          /* local row of the selected candidate, -1 none */
    int64_t row_base;
    int nthreads;
};

static int fail(frx_ctx* c, int code, const char* msg) { snprintf(c->err, sizeof(c->err), "%s", msg); return code; }
static void* dup_mem(const void* p, size_t n) { void* q = malloc(n ? n : 1); if (q && p) memcpy(q, p, n); return q; }

int frx_abi_version(void) { return FRX_ABI_VERSION; }

int frx_create(int device_ordinal, frx_ctx** out) {
    (void)device_ordinal;
    if (!out) return FRX_ERR_INVALID;
    frx_ctx* c = (frx_ctx*)calloc(1, sizeof(frx_ctx));
    if (!c) return FRX_ERR_NOMEM;
    c->winner = -1;
    const char* e = getenv("FRX_CPU_THREADS");
    c->nthreads = e ? atoi(e) : 0;
    *out = c;
    return FRX_OK;
}

static void free_preds(frx_ctx* c) { free(c->pos); free(c->cov); free(c->theta); free(c->hl); free(c->hw); free(c->olen);
    c->pos = c->cov = c->theta = c->hl = c->hw = NULL; c->olen = NULL; c->O = 0; }

int frx_destroy(frx_ctx* c) {
    if (!c) return FRX_OK;
    for (int k = 0; k < 6; k++) free(c->ref[k]);
    free(c->Tvals); free(c->Tlen); free(c->tpow); free_preds(c); free(c->obs_pos); free(c->sobb);
    free(c->states); free(c->costs); free(c->total); free(c->flags); free(c->traj_len); free(c->grid_rows);
    free(c);
    return FRX_OK;
}

const char* frx_last_error(const frx_ctx* c) { return c ? c->err : "null context"; }

int frx_set_reference(frx_ctx* c, int32_t M, const double* p, const double* th, const double* k, const double* kd, const double* x,
                      const double* y) {
    if (!c) return FRX_ERR_INVALID;
    if (M < 2 || !p || !th || !k || !kd || !x || !y) return fail(c, FRX_ERR_INVALID, "frx_set_reference: bad arguments");
    const double* src[6] = {p, th, k, kd, x, y};
    for (int q = 0; q < 6; q++) { free(c->ref[q]); c->ref[q] = (double*)dup_mem(src[q], sizeof(double) * M); }
    c->M = M; c->have_ref = 1;
    return FRX_OK;
}

int frx_set_params(frx_ctx* c, const frx_params* p) {
    if (!c) return FRX_ERR_INVALID;
    if (!p || p->N < 1 || p->N > 63 || p->dt <= 0 || p->n_costs < 0 || p->n_costs > FRX_MAX_COSTS)
        return fail(c, FRX_ERR_INVALID, "frx_set_params: bad arguments");
    if (p->prediction_cost_mode != 0) return fail(c, FRX_ERR_UNSUPPORTED, "frx_set_params: the CPU port has no collision-probability cost");
    if (c->have_params && c->prm.N != p->N) c->have_tables = 0;
    c->prm = *p; c->have_params = 1;
    return FRX_OK;
}

int frx_set_time_tables(frx_ctx* c, int32_t nT, const double* Tv, const int32_t* tl, const double* tpow) {
    if (!c) return FRX_ERR_INVALID;
    if (!c->have_params) return fail(c, FRX_ERR_INVALID, "frx_set_time_tables: call frx_set_params first");
    if (nT < 1 || !Tv || !tl || !tpow) return fail(c, FRX_ERR_INVALID, "frx_set_time_tables: bad arguments");
    const int Nt = c->prm.N + 1;
    free(c->Tvals); free(c->Tlen); free(c->tpow);
    c->Tvals = (double*)dup_mem(Tv, sizeof(double) * nT); c->Tlen = (int32_t*)dup_mem(tl, sizeof(int32_t) * nT);
    c->tpow = (double*)dup_mem(tpow, sizeof(double) * (size_t)nT * 5 * Nt);
    c->nT = nT; c->have_tables = 1;
    return FRX_OK;
}

int frx_set_predictions(frx_ctx* c, int32_t O, int32_t T, const double* pos, const double* cov, const double* theta,
                        const double* hl, const double* hw, const int32_t* len_valid) {
    if (!c) return FRX_ERR_INVALID;
    free_preds(c);
    if (O <= 0) return FRX_OK;
    if (T < 1 || !pos || !cov || !theta || !hl || !hw || !len_valid) return fail(c, FRX_ERR_INVALID, "frx_set_predictions: bad arguments");
    const size_t n = (size_t)O * T;
    c->pos = (double*)dup_mem(pos, sizeof(double) * n * 2); c->cov = (double*)dup_mem(cov, sizeof(double) * n * 4);
    c->theta = (double*)dup_mem(theta, sizeof(double) * n); c->hl = (double*)dup_mem(hl, sizeof(double) * O);
    c->hw = (double*)dup_mem(hw, sizeof(double) * O); c->olen = (int32_t*)dup_mem(len_valid, sizeof(int32_t) * O);
    c->O = O; c->T = T;
    return FRX_OK;
}

int frx_set_obstacle_positions(frx_ctx* c, int32_t n, const double* xy) {
    if (!c) return FRX_ERR_INVALID;
    free(c->obs_pos); c->obs_pos = NULL; c->n_obs_pos = 0;
    if (n > 0 && xy) { c->obs_pos = (double*)dup_mem(xy, sizeof(double) * 2 * n); c->n_obs_pos = n; }
    return FRX_OK;
}

int frx_set_static_obbs(frx_ctx* c, int32_t B, const double* obb) {
    if (!c) return FRX_ERR_INVALID;
    free(c->sobb); c->sobb = NULL; c->B = 0;
    if (B > 0 && obb) { c->sobb = (double*)dup_mem(obb, sizeof(double) * 5 * B); c->B = B; }
    return FRX_OK;
}

static int ensure(frx_ctx* c, int64_t N, int K, int Nt) {
    if (N > c->cap || K != c->K || Nt != c->Nt) {
        free(c->states); free(c->costs); free(c->total); free(c->flags); free(c->traj_len);
        c->states = (double*)malloc(sizeof(double) * 14 * (size_t)N * Nt);
        c->costs = (double*)malloc(sizeof(double) * (size_t)N * (K > 0 ? K : 1));
        c->total = (double*)malloc(sizeof(double) * (size_t)N);
        c->flags = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)N);
        c->traj_len = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
        if (!c->states || !c->costs || !c->total || !c->flags || !c->traj_len) return 0;
        c->cap = N; c->K = K; c->Nt = Nt;
    }
    return 1;
}

int frx_plan(frx_ctx* c, int64_t N, const double* sampling, int64_t row_index_base, frx_result* out) {
    if (!c) return FRX_ERR_INVALID;
    if (!c->have_params || !c->have_ref || !c->have_tables)
        return fail(c, FRX_ERR_INVALID, "frx_plan: frx_set_params, frx_set_reference and frx_set_time_tables must be called first");
    if (N < 1 || !sampling || !out) return fail(c, FRX_ERR_INVALID, "frx_plan: empty sampling matrix");
    const frx_params* p = &c->prm;
    const int Nt = p->N + 1, K = p->n_costs;
    if (!ensure(c, N, K, Nt)) return fail(c, FRX_ERR_NOMEM, "frx_plan: out of memory");
    orc_params P;
    memset(&P, 0, sizeof(P));
    P.dt = p->dt; P.N = p->N; P.low = p->low_vel_mode; P.draw = p->draw_traj_set; P.debug = p->kinematic_debug;
    P.a_max = p->a_max; P.v_switch = p->v_switch; P.delta_max = p->delta_max; P.wheelbase = p->wheelbase; P.wb_rear = p->wb_rear_axle;
    P.length = p->length; P.width = p->width; P.x0_orientation = p->x0_orientation; P.v_des = p->desired_velocity;
    P.n_costs = K;
    for (int k = 0; k < K; k++) { P.cost_ids[k] = p->cost_ids[k]; P.w[k] = p->cost_weights[k]; }
    P.check_all_collisions = 0;                       /* the reference walks the sorted list lazily (planner.py:329-392) */
    P.collision_check = p->check_collisions;
    P.kd_from_v_delta = p->curvature_rate_from_v_delta; P.vo_norm = p->velocity_offset_norm; P.v_delta_max = p->v_delta_max;
    orc_result r;
    int rc = orc_plan(&P, N, sampling, c->M, c->ref[0], c->ref[1], c->ref[2], c->ref[3], c->ref[4], c->ref[5], c->nT, c->Tvals,
                      c->Tlen, c->tpow, p->check_collisions ? c->O : 0, c->T, c->pos, c->cov, c->theta, c->hl, c->hw, c->olen,
                      c->n_obs_pos, c->obs_pos, p->check_collisions ? c->B : 0, c->sobb, c->states, c->costs, c->total, c->flags,
                      c->traj_len, NULL, &r, c->nthreads);
    if (rc != 0) return fail(c, FRX_ERR_INVALID, "frx_plan: the port rejected the plan (horizon or cost count out of range)");
    memset(out, 0, sizeof(*out));
    c->N = N; c->winner = r.argmin; c->row_base = row_index_base;
    out->argmin = r.argmin >= 0 ? r.argmin + row_index_base : -1;
    out->min_cost = r.argmin >= 0 ? r.min_cost : INFINITY;
    out->n_rows = N; out->n_in_list = r.n_in_list; out->n_feasible = r.n_feasible; out->n_candidates = r.n_candidates;
    out->collision_counter = r.collision_counter;
    for (int q = 0; q < 11; q++) out->reason_counts[q] = r.reason_counts[q];
    for (int64_t q = 0; q < N; q++) {
        if ((c->flags[q] & FRX_FLAG_CANDIDATE) && (c->flags[q] & FRX_FLAG_COLLIDE)) out->n_collide++;
        if ((c->flags[q] & FRX_FLAG_CANDIDATE) && (c->flags[q] & FRX_FLAG_BOUNDARY)) out->n_boundary++;
    }
    return FRX_OK;
}

int frx_plan_grid(frx_ctx* c, int32_t nt, const double* t1, int32_t nv, const double* ss1, int32_t nd, const double* d1,
                  const double* x_cl, int64_t row_first, int64_t row_count, frx_result* out) {
    if (!c) return FRX_ERR_INVALID;
    const int64_t total = (int64_t)nt * nv * nd;
    if (nt < 1 || nv < 1 || nd < 1 || !t1 || !ss1 || !d1 || !x_cl || row_first < 0 || row_count < 1 || row_first + row_count > total)
        return fail(c, FRX_ERR_INVALID, "frx_plan_grid: bad arguments");
    if (row_count > c->grid_cap) { free(c->grid_rows); c->grid_rows = (double*)malloc(sizeof(double) * 13 * (size_t)row_count); c->grid_cap = row_count; }
    if (!c->grid_rows) return fail(c, FRX_ERR_NOMEM, "frx_plan_grid: out of memory");
    for (int64_t q = 0; q < row_count; q++) {         /* generate_sampling_matrix order: t1 slowest, then ss1, then d1 */
        const int64_t g = row_first + q;
        const int it = (int)(g / ((int64_t)nv * nd)), rem = (int)(g - (int64_t)it * nv * nd), iv = rem / nd, id = rem - iv * nd;
        double* r = c->grid_rows + 13 * q;
        r[0] = 0.0; r[1] = t1[it]; r[2] = x_cl[0]; r[3] = x_cl[1]; r[4] = x_cl[2]; r[5] = ss1[iv]; r[6] = 0.0;
        r[7] = x_cl[3]; r[8] = x_cl[4]; r[9] = x_cl[5]; r[10] = d1[id]; r[11] = 0.0; r[12] = 0.0;
    }
    return frx_plan(c, row_count, c->grid_rows, row_first, out);
}

int32_t frx_state_pitch(const frx_ctx* c) { return c ? ((c->Nt + 3) & ~3) : 0; }
int32_t frx_last_launches(const frx_ctx* c) { (void)c; return 0; }

static int copy_row(const frx_ctx* c, int64_t row, uint32_t mask, int64_t n_idx, int64_t slot, double* out) {
    const int pitch = (c->Nt + 3) & ~3;
    int fo = 0;
    for (int f = 0; f < FRX_NUM_FIELDS; f++) {
        if (!(mask & (1u << f))) continue;
        double* dst = out + ((size_t)fo * n_idx + slot) * pitch;
        memcpy(dst, c->states + ((size_t)f * c->N + row) * c->Nt, sizeof(double) * c->Nt);
        for (int i = c->Nt; i < pitch; i++) dst[i] = 0.0;
        fo++;
    }
    return fo;
}

int frx_get_states(frx_ctx* c, int64_t n_idx, const int64_t* idx, uint32_t mask, double* out) {
    if (!c || c->N <= 0 || n_idx < 1 || !idx || !out || !mask) return FRX_ERR_INVALID;
    for (int64_t k = 0; k < n_idx; k++) {
        if (idx[k] < 0 || idx[k] >= c->N) return fail(c, FRX_ERR_INVALID, "frx_get_states: row index out of range");
        copy_row(c, idx[k], mask, n_idx, k, out);
    }
    return FRX_OK;
}

int frx_get_states_range(frx_ctx* c, int64_t first, int64_t count, uint32_t mask, double* out) {
    if (!c || c->N <= 0 || first < 0 || count < 1 || first + count > c->N || !out || !mask) return FRX_ERR_INVALID;
    for (int64_t k = 0; k < count; k++) copy_row(c, first + k, mask, count, k, out);
    return FRX_OK;
}

int frx_get_costs(frx_ctx* c, int64_t first, int64_t count, double* costs, double* total) {
    if (!c || c->N <= 0 || first < 0 || count < 1 || first + count > c->N) return FRX_ERR_INVALID;
    if (costs && c->K > 0) memcpy(costs, c->costs + (size_t)first * c->K, sizeof(double) * (size_t)count * c->K);
    if (total) memcpy(total, c->total + first, sizeof(double) * (size_t)count);
    return FRX_OK;
}

int frx_get_flags(frx_ctx* c, int64_t first, int64_t count, uint32_t* flags, int32_t* tl) {
    if (!c || c->N <= 0 || first < 0 || count < 1 || first + count > c->N) return FRX_ERR_INVALID;
    if (flags) memcpy(flags, c->flags + first, sizeof(uint32_t) * (size_t)count);
    if (tl) memcpy(tl, c->traj_len + first, sizeof(int32_t) * (size_t)count);
    return FRX_OK;
}

int frx_winner_states(frx_ctx* c, uint32_t mask, double* out) {
    if (!c || c->N <= 0 || !out || !mask) return FRX_ERR_INVALID;
    if (c->winner < 0) return fail(c, FRX_ERR_INVALID, "frx_winner_states: the last plan selected no candidate");
    copy_row(c, c->winner, mask, 1, 0, out);
    return FRX_OK;
}

int frx_winner_record(frx_ctx* c, uint32_t* flags, int32_t* tl, double* total, double* costs) {
    if (!c || c->N <= 0) return FRX_ERR_INVALID;
    if (c->winner < 0) return fail(c, FRX_ERR_INVALID, "frx_winner_record: the last plan selected no candidate");
    if (flags) *flags = c->flags[c->winner];
    if (tl) *tl = c->traj_len[c->winner];
    if (total) *total = c->total[c->winner];
    if (costs) memcpy(costs, c->costs + (size_t)c->winner * c->K, sizeof(double) * c->K);
    return FRX_OK;
}

int frx_synchronize(frx_ctx* c) { return c ? FRX_OK : FRX_ERR_INVALID; }
int frx_set_stream(frx_ctx* c, void* s) { (void)s; return c ? FRX_OK : FRX_ERR_INVALID; }

/* ---- device-only entry points */
#define UNSUPPORTED(c, what) (c ? fail(c, FRX_ERR_UNSUPPORTED, what ": not available in the CPU build of the ABI") : FRX_ERR_INVALID)
int frx_plan_device(frx_ctx* c, int64_t N, const void* d, int64_t b, frx_result* o) { (void)N; (void)d; (void)b; (void)o; return UNSUPPORTED(c, "frx_plan_device"); }
int frx_plan_device_async(frx_ctx* c, int64_t N, const void* d, int64_t b) { (void)N; (void)d; (void)b; return UNSUPPORTED(c, "frx_plan_device_async"); }
int frx_plan_wait(frx_ctx* c, frx_result* o) { (void)o; return UNSUPPORTED(c, "frx_plan_wait"); }
int frx_plan_batched(int32_t n, frx_ctx** cs, const int64_t* rows, const double* const* S, frx_result* res) {
    if (n < 1 || !cs || !rows || !S || !res) return FRX_ERR_INVALID;
    for (int a = 0; a < n; a++) { int rc = frx_plan(cs[a], rows[a], S[a], 0, &res[a]); if (rc != FRX_OK) return rc; }
    return FRX_OK;                                     /* (sequential: what AgentBatch._step_agents does) */
}
int frx_device_pointers(frx_ctx* c, void** a, void** b, void** d, void** e) { (void)a; (void)b; (void)d; (void)e; return UNSUPPORTED(c, "frx_device_pointers"); }
int frx_winner_device_pointer(frx_ctx* c, void** w) { (void)w; return UNSUPPORTED(c, "frx_winner_device_pointer"); }
int frx_selftest_fdiv(frx_ctx* c, int64_t n, const double* a, const double* b, double* q1, double* q2) { (void)n; (void)a; (void)b; (void)q1; (void)q2; return UNSUPPORTED(c, "frx_selftest_fdiv"); }
int frx_selftest_divc(frx_ctx* c, int64_t n, const double* a, double b, double* q1, double* q2) { (void)n; (void)a; (void)b; (void)q1; (void)q2; return UNSUPPORTED(c, "frx_selftest_divc"); }
int frx_selftest_fp64_peak(frx_ctx* c, double* t) { (void)t; return UNSUPPORTED(c, "frx_selftest_fp64_peak"); }
int frx_set_exchange(frx_ctx* c, void* p, int32_t r, int32_t w) { (void)p; (void)r; (void)w; return UNSUPPORTED(c, "frx_set_exchange"); }
int frx_exchange_wait(frx_ctx* c, int64_t t, double* a, int64_t* b, int32_t* o) { (void)t; (void)a; (void)b; (void)o; return UNSUPPORTED(c, "frx_exchange_wait"); }
int frx_set_reference_polyline(frx_ctx* c, int32_t M, const double* xy) { (void)M; (void)xy; return UNSUPPORTED(c, "frx_set_reference_polyline"); }
int frx_get_reference(frx_ctx* c, int32_t M, double* out) { (void)M; (void)out; return UNSUPPORTED(c, "frx_get_reference"); }
int frx_initial_state(frx_ctx* c, const double* x0, int32_t low, double wb, double* x_cl) { (void)x0; (void)low; (void)wb; (void)x_cl; return UNSUPPORTED(c, "frx_initial_state"); }
