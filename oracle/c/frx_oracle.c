/*
 * frx_oracle.c -- CPU ORACLE (plain C, OpenMP over candidates).  TEST INFRASTRUCTURE ONLY.
 *
 * A second, independent restatement of the reference's Python path (use_cpp=False) so that the
 * CUDA library can be checked at sizes the numpy oracle (oracle/frenet_oracle.py) cannot reach,
 * and so that bench.py has a CPU baseline to time.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this file's library; the product never
 * does.
 *
 * One thread walks one candidate with the reference's own loop structure:
 *   coefficients ......... frenetix_motion_planner/polynomial_trajectory.py:293-343,452-488
 *                          (closed forms instead of LAPACK gesv: results agree to a few ulp, see
 *                          the margin protocol in oracle/frenet_oracle.py)
 *   Frenet samples ....... frenetix_motion_planner/reactive_planner.py:296-355
 *   per-step projection .. reactive_planner.py:389-478,  gates :483-533,  x/y :536-547
 *   costs ................ cost_functions/partial_cost_functions.py:24-64,120-196,341-356 with
 *                          numpy's pairwise summation order reproduced (np_sum below),
 *                          cost_functions/cost_function.py:78-91
 *   Mahalanobis .......... risk_assessment/collision_probability.py:264-299
 *   collision ............ frenetix_motion_planner/planner.py:329-392 (lazy walk over the
 *                          cost-sorted list or all candidates), collision_check.py:110-200
 * Third-party stand-ins (CCosy point conversion, make_valid_orientation, obb-sum hull, OBB
 * overlap) use the definitions documented in oracle/frenet_oracle.py -- PARITY-UNPINNED there,
 * identical here.
 *
 * Pinning: tests/test_c_oracle.py checks this file against the numpy oracle and against the golden
 * vectors produced by the reference's own code (tests/golden).
 * Build: oracle/build.py  (gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
static int orc_threads_used = 1;

#define EPS 1e-5
#define NF 14
enum { F_X, F_Y, F_THETA, F_V, F_A, F_KAPPA, F_KAPPA_DOT, F_S, F_D, F_THETA_CL, F_S_DOT, F_S_DDOT, F_D_DOT, F_D_DDOT };
#define FLAG_VALID (1u << 0)
#define FLAG_FEASIBLE (1u << 1)
#define FLAG_REASON(r) (1u << (1 + (r)))
#define FLAG_COLLIDE (1u << 12)
#define FLAG_BOUNDARY (1u << 13)
#define FLAG_STORED (1u << 14)
#define FLAG_IN_LIST (1u << 15)
#define FLAG_COSTED (1u << 16)
#define FLAG_CANDIDATE (1u << 17)
enum { C_ACCELERATION, C_DISTANCE_TO_OBSTACLES, C_DISTANCE_TO_REFERENCE_PATH, C_JERK, C_LATERAL_JERK,
       C_LONGITUDINAL_JERK, C_ORIENTATION_OFFSET, C_PATH_LENGTH, C_PREDICTION, C_VELOCITY_OFFSET };
#define MAXNT 64

typedef struct {
    double dt;
    int32_t N, low, draw, debug;
    double a_max, v_switch, delta_max, wheelbase, wb_rear, length, width, x0_orientation, v_des;
    int32_t n_costs;
    int32_t cost_ids[10];
    double w[10];
    int32_t check_all_collisions; /* 1: collision test for every candidate; 0: lazy like planner.py:329-392 */
    int32_t collision_check;      /* 0: skip the prediction collision test (selection = first of the sorted list) */
    int32_t kd_from_v_delta;      /* cpp flavour: kappa_dot_max = v_delta_max / (wheelbase cos^2(atan(wheelbase kappa))) */
    int32_t vo_norm;              /* 2: squared velocity offsets */
    double v_delta_max;
} orc_params;

typedef struct {
    int64_t argmin;
    double min_cost;
    int64_t n_in_list, n_feasible, n_candidates, collision_counter;
    int64_t reason_counts[11];
} orc_result;

/* numpy's pairwise summation (numpy/core/src/umath/loops_utils.h.src, n <= 128 branch) */
static double np_sum(const double* a, int n) {
    if (n < 8) {
        double res = 0.;
        for (int i = 0; i < n; i++) res += a[i];
        return res;
    }
    double r[8], res;
    int i;
    for (i = 0; i < 8; i++) r[i] = a[i];
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; j++) r[j] += a[i + j];
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
}

static double make_valid_orientation(double angle) {
    const double two_pi = 2.0 * M_PI;
    double m = fmod(angle, two_pi); /* python float %: result takes the sign of the divisor */
    if (m != 0.0) { if (m < 0.0) m += two_pi; } else { m = 0.0; }
    if (M_PI <= m && m <= two_pi) m = m - two_pi;
    return m;
}

static int first_greater(const double* p, int M, double s) { /* np.argmax(p > s) */
    for (int j = 0; j < M; j++)
        if (p[j] > s) return j;
    return 0;
}

static double ppos(const double* c, double t, double t2, double t3, double t4, double t5) {
    return c[0] + c[1] * t + c[2] * t2 + c[3] * t3 + c[4] * t4 + c[5] * t5;
}
static double pvel(const double* c, double t, double t2, double t3, double t4) {
    return c[1] + 2. * c[2] * t + 3. * c[3] * t2 + 4. * c[4] * t3 + 5. * c[5] * t4;
}
static double pacc(const double* c, double t, double t2, double t3) {
    return 2 * c[2] + 6 * c[3] * t + 12 * c[4] * t2 + 20 * c[5] * t3;
}
static double sq_jerk_integral(const double* c, double t) {
    double t2 = t * t, t3 = t2 * t, t4 = t3 * t, t5 = t4 * t;
    return (36 * c[3] * c[3] * t + 144 * c[3] * c[4] * t2 + 240 * c[3] * c[5] * t3 + 192 * c[4] * c[4] * t3 +
            720 * c[4] * c[5] * t4 + 720 * c[5] * c[5] * t5);
}

typedef struct { double cx, cy, ux, uy, ha, hb; } hull_t;

static hull_t obb_sum_hull(double c0x, double c0y, double th0, double c1x, double c1y, double th1, double hl, double hw) {
    double ux = cos(th0), uy = sin(th0), u1x = cos(th1), u1y = sin(th1);
    double dx = c1x - c0x, dy = c1y - c0y;
    double du = dx * ux + dy * uy;
    double dv = dy * ux - dx * uy;
    double c = fabs(ux * u1x + uy * u1y);
    double sn = fabs(ux * u1y - uy * u1x);
    double eu = hl * c + hw * sn, ev = hl * sn + hw * c;
    double lo_u = fmin(-hl, du - eu), hi_u = fmax(hl, du + eu);
    double lo_v = fmin(-hw, dv - ev), hi_v = fmax(hw, dv + ev);
    double mu = 0.5 * (lo_u + hi_u), mv = 0.5 * (lo_v + hi_v);
    hull_t h;
    h.ha = 0.5 * (hi_u - lo_u); h.hb = 0.5 * (hi_v - lo_v);
    h.cx = c0x + (mu * ux - mv * uy); h.cy = c0y + (mu * uy + mv * ux);
    h.ux = ux; h.uy = uy;
    return h;
}

static int obb_overlap(const hull_t* e, const hull_t* o) {
    double dx = o->cx - e->cx, dy = o->cy - e->cy;
    double c = fabs(e->ux * o->ux + e->uy * o->uy);
    double sn = fabs(e->ux * o->uy - e->uy * o->ux);
    if (fabs(dx * e->ux + dy * e->uy) > e->ha + (o->ha * c + o->hb * sn)) return 0;
    if (fabs(dy * e->ux - dx * e->uy) > e->hb + (o->ha * sn + o->hb * c)) return 0;
    if (fabs(dx * o->ux + dy * o->uy) > o->ha + (e->ha * c + e->hb * sn)) return 0;
    if (fabs(dy * o->ux - dx * o->uy) > o->hb + (e->ha * sn + e->hb * c)) return 0;
    return 1;
}

typedef struct {
    const orc_params* p;
    int M; const double *rp, *rth, *rc, *rcd, *rx, *ry;
    int nT; const double* Tvals; const int32_t* Tlen; const double* tpow;
    int O, T; const double *pos, *icov, *otheta, *ohl, *ohw; const int32_t* olen; const hull_t* ohull;
    int n_obs_pos; const double* obs_pos;
    int B; const hull_t* sobb;
    double kappa_max;
} env_t;

#define UPD(m) do { double m__ = fabs(m); if (m__ != m__) m__ = 0.0; if (m__ < margin) margin = m__; } while (0)

/* one candidate: returns flags; fills st[14][Nt] (caller scratch), coefficients, traj_len, margin */
static uint32_t eval_candidate(const env_t* E, const double* row, double* st, double* c_lon, double* c_lat,
                               int* traj_len_out, double* margin_out) {
    const orc_params* P = E->p;
    const int Nt = P->N + 1;
    const double dT = P->dt;
    const int low = P->low, draw = P->draw, debug = P->debug;
    const int brk = !draw && !debug;
    double margin = INFINITY;
    double T = row[1], s0 = row[2], ss0 = row[3], sss0 = row[4], ss1 = row[5];
    double d0 = row[7], dd0 = row[8], ddd0 = row[9], d1 = row[10], dd1 = row[11], ddd1 = row[12];

    int tix = -1;
    for (int k = 0; k < E->nT; k++) if (E->Tvals[k] == T) { tix = k; break; }
    if (tix < 0) { *traj_len_out = 0; *margin_out = 0; return 0; }
    const int traj_len = E->Tlen[tix];
    const double* tp = E->tpow + (size_t)tix * 5 * Nt;
    *traj_len_out = traj_len;

    {   /* quartic, end acceleration hard-wired to 0 */
        double T2 = T * T, T3 = T2 * T;
        double b0 = (ss1 - ss0) - sss0 * T, b1 = -sss0;
        c_lon[0] = s0; c_lon[1] = ss0; c_lon[2] = sss0 / 2.0;
        c_lon[3] = (3 * b0 - T * b1) / (3 * T2);
        c_lon[4] = (T * b1 - 2 * b0) / (4 * T3);
        c_lon[5] = 0.0;
    }
    {
        double tau = T;
        if (low) {
            double t2 = T * T, t3 = t2 * T, t4 = t2 * t2, t5 = t3 * t2;
            double goal = ppos(c_lon, T, t2, t3, t4, t5) - s0;
            UPD(goal);
            tau = (goal <= 0) ? T : goal;
        }
        double u2 = tau * tau, u3 = u2 * tau, u4 = u2 * u2, u5 = u4 * tau;
        double b0 = ((d1 - d0) - dd0 * tau) - (.5 * ddd0) * u2;
        double b1 = (dd1 - dd0) - ddd0 * tau;
        double b2 = ddd1 - ddd0;
        c_lat[0] = d0; c_lat[1] = dd0; c_lat[2] = .5 * ddd0;
        c_lat[3] = ((10 * b0 - (4 * b1) * tau) + (0.5 * b2) * u2) / u3;
        c_lat[4] = ((-15 * b0 + (7 * b1) * tau) - b2 * u2) / u4;
        c_lat[5] = ((6 * b0 - (3 * b1) * tau) + (0.5 * b2) * u2) / u5;
    }

    double *x = st + F_X * Nt, *y = st + F_Y * Nt, *thg = st + F_THETA * Nt, *v = st + F_V * Nt, *a = st + F_A * Nt;
    double *kap = st + F_KAPPA * Nt, *kapd = st + F_KAPPA_DOT * Nt, *s = st + F_S * Nt, *d = st + F_D * Nt;
    double *thc = st + F_THETA_CL * Nt, *sd = st + F_S_DOT * Nt, *sdd = st + F_S_DDOT * Nt, *dd = st + F_D_DOT * Nt,
           *ddd = st + F_D_DDOT * Nt;
    memset(st, 0, sizeof(double) * NF * Nt);

    for (int i = 0; i < traj_len; i++) {
        s[i] = ppos(c_lon, tp[i], tp[Nt + i], tp[2 * Nt + i], tp[3 * Nt + i], tp[4 * Nt + i]);
        sd[i] = pvel(c_lon, tp[i], tp[Nt + i], tp[2 * Nt + i], tp[3 * Nt + i]);
        sdd[i] = pacc(c_lon, tp[i], tp[Nt + i], tp[2 * Nt + i]);
    }
    for (int i = traj_len; i < Nt; i++) {
        s[i] = s[i - 1] + dT * sd[traj_len - 1];
        sd[i] = sd[traj_len - 1];
        sdd[i] = 0.0;
    }
    if (!low) {
        for (int i = 0; i < traj_len; i++) {
            d[i] = ppos(c_lat, tp[i], tp[Nt + i], tp[2 * Nt + i], tp[3 * Nt + i], tp[4 * Nt + i]);
            dd[i] = pvel(c_lat, tp[i], tp[Nt + i], tp[2 * Nt + i], tp[3 * Nt + i]);
            ddd[i] = pacc(c_lat, tp[i], tp[Nt + i], tp[2 * Nt + i]);
        }
    } else {
        for (int i = 0; i < traj_len; i++) {
            double q1 = s[i] - s[0], q2 = q1 * q1, q3 = q2 * q1, q4 = q2 * q2, q5 = q4 * q1;
            d[i] = ppos(c_lat, q1, q2, q3, q4, q5);
            dd[i] = pvel(c_lat, q1, q2, q3, q4);
            ddd[i] = pacc(c_lat, q1, q2, q3);
        }
    }
    for (int i = traj_len; i < Nt; i++) { d[i] = d[traj_len - 1]; dd[i] = 0.0; ddd[i] = 0.0; }

    int valid = 1, feasible = 1, in_list = 1, stored = 1;
    uint32_t reasons = 0;
    int any_neg = 0, any_acc = 0;
    for (int i = 0; i < Nt; i++) {
        UPD(sd[i] + EPS);
        UPD(fabs(sd[i]) - EPS);
        if (!draw) UPD(fabs(sdd[i]) - P->a_max);
        if (sd[i] < -EPS) any_neg = 1;
        if (fabs(sdd[i]) > P->a_max) any_acc = 1;
    }
    if (any_neg) {
        valid = 0; reasons |= FLAG_REASON(10);
        if (brk) { in_list = 0; stored = 0; }
    }
    for (int i = 0; i < Nt; i++) if (fabs(sd[i]) < EPS) sd[i] = 0.0;
    if (in_list && !draw) {
        if (any_acc) { feasible = 0; reasons |= FLAG_REASON(1); stored = 0; }
        else if (any_neg) { feasible = 0; reasons |= FLAG_REASON(2); stored = 0; }
    }

    if (in_list && stored) {
        const double* rp = E->rp;
        const int M = E->M;
        uint32_t g_or = 0;
        for (int i = 0; i < Nt; i++) {
            double dp, dpp;
            if (!low) {
                UPD(sd[i] - 0.001);
                dp = (sd[i] > 0.001) ? dd[i] / sd[i] : 0.;
                double ddot = ddd[i] - dp * sdd[i];
                dpp = (sd[i] > 0.001) ? ddot / (sd[i] * sd[i]) : 0.;
            } else { dp = dd[i]; dpp = ddd[i]; }
            int j = first_greater(rp, M, s[i]);
            int ia = (j == 0) ? M - 1 : j - 1;
            double pa = rp[ia], pb = rp[j];
            double lam = (s[i] - pa) / (pb - pa);
            double interp = make_valid_orientation((E->rth[j] - E->rth[ia]) * (s[i] - pa) / (pb - pa) + E->rth[ia]);
            if (sd[i] > 0.001 || low) {
                thc[i] = atan2(dp, 1.0);
                thg[i] = thc[i] + interp;
            } else {
                thg[i] = (i == 0) ? P->x0_orientation : thg[i - 1];
                thc[i] = thg[i] - interp;
            }
            double k_r = (E->rc[j] - E->rc[ia]) * lam + E->rc[ia];
            double k_r_d = (E->rcd[j] - E->rcd[ia]) * lam + E->rcd[ia];
            double oneKrD = 1 - k_r * d[i];
            double cosT = cos(thc[i]), tanT = tan(thc[i]);
            double cq = cosT / oneKrD;
            kap[i] = (dpp + (k_r * dp + k_r_d * d[i]) * tanT) * cosT * (cq * cq) + cq * k_r;
            double qc = oneKrD / cosT;
            v[i] = sd[i] * qc;
            a[i] = sdd[i] * qc + ((sd[i] * sd[i]) / cosT) * (oneKrD * tanT * (kap[i] * qc - k_r) - (k_r_d * d[i] + k_r * dp));

            uint32_t g = 0;
            UPD(v[i] + EPS);
            if (v[i] < -EPS) g |= 1u;
            UPD(fabs(kap[i]) - E->kappa_max);
            if (fabs(kap[i]) > E->kappa_max) g |= 2u;
            double yaw_rate = (i > 0) ? (thg[i] - thg[i - 1]) / dT : 0.;
            double ry = fabs(rint(yaw_rate * 100000.0) / 100000.0), tmax = E->kappa_max * v[i];
            if (!(ry == 0 && tmax == 0)) UPD(ry - tmax);
            if (ry > tmax) g |= 4u;
            double kdot = (i > 0) ? (kap[i] - kap[i - 1]) / dT : 0.;
            double kd_max = 0.4;
            if (P->kd_from_v_delta) {       /* the algebraic form the device uses: cos^2(atan(x)) = 1 / (1 + x^2) */
                double wk = P->wheelbase * kap[i];
                kd_max = (P->v_delta_max / P->wheelbase) * (1.0 + wk * wk);
            }
            UPD(fabs(kdot) - kd_max);
            if (fabs(kdot) > kd_max) g |= 8u;
            double a_hi = (v[i] > P->v_switch) ? P->a_max * P->v_switch / v[i] : P->a_max;
            UPD(a[i] - a_hi); UPD(a[i] + P->a_max);
            if (!(-P->a_max <= a[i] && a[i] <= a_hi)) g |= 16u;
            if (g) {
                if (brk) { g_or = g & (~g + 1u); break; }
                g_or |= g;
            }
        }
        if (g_or) {
            feasible = 0;
            if (g_or & 1u) reasons |= FLAG_REASON(4);
            if (g_or & 2u) reasons |= FLAG_REASON(5);
            if (g_or & 4u) reasons |= FLAG_REASON(6);
            if (g_or & 8u) reasons |= FLAG_REASON(7);
            if (g_or & 16u) reasons |= FLAG_REASON(8);
        }
        for (int i = 0; i < Nt; i++) { UPD(s[i] - rp[0]); UPD(s[i] - rp[M - 1]); }
        stored = feasible || draw;
        in_list = stored;
        if (stored) {
            for (int i = 0; i < Nt; i++) {
                if (!(s[i] >= rp[0]) || !(s[i] < rp[M - 1])) { valid = 0; reasons |= FLAG_REASON(9); break; }
                int j = first_greater(rp, M, s[i]);
                int ia = j - 1;
                double lam = (s[i] - rp[ia]) / (rp[j] - rp[ia]);
                double px = (1.0 - lam) * E->rx[ia] + lam * E->rx[j];
                double py = (1.0 - lam) * E->ry[ia] + lam * E->ry[j];
                double th = E->rth[ia] + lam * (E->rth[j] - E->rth[ia]);
                x[i] = px - d[i] * sin(th);
                y[i] = py + d[i] * cos(th);
            }
            kapd[0] = 0.0;
            for (int i = 1; i < Nt; i++) kapd[i] = kap[i] - kap[i - 1];
        }
    }
    uint32_t fl = reasons;
    if (valid) fl |= FLAG_VALID;
    if (feasible) fl |= FLAG_FEASIBLE;
    if (stored) fl |= FLAG_STORED;
    if (in_list) fl |= FLAG_IN_LIST;
    *margin_out = margin;
    return fl;
}

static double simps(const double* yv, int n, double dx) { /* scipy 1.13 simps(y, dx=dx) */
    double tmp[MAXNT];
    if (n == 1) return 0.0;
    if (n == 2) return 0.5 * dx * (yv[0] + yv[1]);
    int nb = (n & 1) ? n : n - 1, m = 0;
    for (int k = 0; k + 2 < nb; k += 2) tmp[m++] = yv[k] + 4.0 * yv[k + 1] + yv[k + 2];
    double result = dx / 3.0 * np_sum(tmp, m);
    if (!(n & 1)) {
        double alpha = (2 * (dx * dx) + 3 * dx * dx) / (6 * (dx + dx));
        double beta = ((dx * dx) + 3.0 * dx * dx) / (6 * dx);
        double eta = (1 * (dx * dx * dx)) / (6 * dx * (dx + dx));
        result += alpha * yv[n - 1] + beta * yv[n - 2] - eta * yv[n - 3];
    }
    return result;
}

static void eval_costs(const env_t* E, const double* st, const double* c_lon, const double* c_lat, double* costs,
                       double* total) {
    const orc_params* P = E->p;
    const int Nt = P->N + 1;
    const double dT = P->dt;
    const double *x = st + F_X * Nt, *y = st + F_Y * Nt, *v = st + F_V * Nt, *a = st + F_A * Nt, *d = st + F_D * Nt,
                 *thc = st + F_THETA_CL * Nt;
    double tmp[MAXNT], wsum[10];
    for (int k = 0; k < P->n_costs; k++) {
        double c = 0.0;
        switch (P->cost_ids[k]) {
        case C_LATERAL_JERK: c = sq_jerk_integral(c_lat, dT); break;
        case C_LONGITUDINAL_JERK: c = sq_jerk_integral(c_lon, dT); break;
        case C_VELOCITY_OFFSET: {
            int half = Nt / 2, m = 0;
            for (int i = half; i < Nt - 1; i++) { double dv = v[i] - P->v_des; tmp[m++] = (P->vo_norm == 2) ? dv * dv : fabs(dv); }
            c = np_sum(tmp, m);
            double dv = v[Nt - 1] - P->v_des;
            c += fabs(dv * dv);
        } break;
        case C_DISTANCE_TO_REFERENCE_PATH: {
            for (int i = 0; i < Nt; i++) tmp[i] = fabs(d[i]);
            c = (np_sum(tmp, Nt) + fabs(d[Nt - 1]) * 5) / (double)Nt;
        } break;
        case C_PREDICTION: {
            for (int o = 0; o < E->O; o++) {
                const int len = E->olen[o];
                for (int i = 1; i < Nt; i++) {
                    if (i < len) {
                        const double* mu = E->pos + ((size_t)o * E->T + (i - 1)) * 2;
                        const double* iv = E->icov + ((size_t)o * E->T + (i - 1)) * 4;
                        double ex = x[i] - mu[0], ey = y[i] - mu[1];
                        double t0 = ex * iv[0] + ey * iv[2];
                        double t1 = ex * iv[1] + ey * iv[3];
                        double m = t0 * ex + t1 * ey;
                        tmp[i - 1] = 1.0 / (m * m);
                    } else tmp[i - 1] = 0.0;
                }
                c += np_sum(tmp, Nt - 1);
            }
        } break;
        case C_DISTANCE_TO_OBSTACLES: {
            for (int o = 0; o < E->n_obs_pos; o++) {
                for (int i = 0; i < Nt; i++) {
                    double ex = x[i] - E->obs_pos[2 * o], ey = y[i] - E->obs_pos[2 * o + 1];
                    double dist = sqrt(ex * ex + ey * ey);
                    tmp[i] = 1.0 / (dist * dist);
                }
                c += np_sum(tmp, Nt);
            }
        } break;
        case C_ACCELERATION: {
            for (int i = 0; i < Nt; i++) tmp[i] = a[i] * a[i];
            c = simps(tmp, Nt, dT);
        } break;
        case C_JERK: {
            double q[MAXNT];
            for (int i = 0; i < Nt - 1; i++) { double j = (a[i + 1] - a[i]) / dT; q[i] = j * j; }
            c = simps(q, Nt - 1, dT);
        } break;
        case C_ORIENTATION_OFFSET: {
            double q[MAXNT];
            for (int i = 0; i < Nt - 1; i++) { double j = (thc[i + 1] - thc[i]) / dT; q[i] = j * j; }
            c = simps(q, Nt - 1, dT);
        } break;
        case C_PATH_LENGTH: c = simps(v, Nt, dT); break;
        default: break;
        }
        costs[k] = c;
        wsum[k] = P->w[k] * c;
    }
    *total = np_sum(wsum, P->n_costs);
}

static void ego_hulls(const env_t* E, const double* st, hull_t* eh) {
    const orc_params* P = E->p;
    const int Nt = P->N + 1;
    const double *x = st + F_X * Nt, *y = st + F_Y * Nt, *th = st + F_THETA * Nt;
    double cx[MAXNT], cy[MAXNT];
    for (int i = 0; i < Nt; i++) { cx[i] = x[i] + P->wb_rear * cos(th[i]); cy[i] = y[i] + P->wb_rear * sin(th[i]); }
    for (int k = 0; k < Nt - 1; k++)
        eh[k] = obb_sum_hull(cx[k], cy[k], th[k], cx[k + 1], cy[k + 1], th[k + 1], P->length / 2, P->width / 2);
}

/* index k of the first ego hull that meets a predicted obstacle's hull k - 1, -1 if none (the time index pycrcc reports) */
static int collides_predictions(const env_t* E, const hull_t* eh) {
    const int Nt = E->p->N + 1;
    int first = -1;
    for (int o = 0; o < E->O; o++) {
        int L = E->olen[o] < Nt ? E->olen[o] : Nt;
        if (L <= 2) continue;
        int kmax = (Nt - 2 < L - 1) ? Nt - 2 : L - 1;
        if (first >= 0 && kmax >= first) kmax = first - 1;
        for (int k = 1; k <= kmax; k++)
            if (obb_overlap(&eh[k], &E->ohull[(size_t)o * E->T + (k - 1)])) { first = k; break; }
    }
    return first;
}

static int collides_static(const env_t* E, const hull_t* eh) {
    const int Nt = E->p->N + 1;
    for (int k = 0; k < Nt - 1; k++)
        for (int b = 0; b < E->B; b++)
            if (obb_overlap(&eh[k], &E->sobb[b])) return k;
    return -1;
}
#define COLLIDE_BITS(k) (FLAG_COLLIDE | ((uint32_t)(k) << 18))
#define BOUNDARY_BITS(k) (FLAG_BOUNDARY | ((uint32_t)(k) << 24))

typedef struct { double cost; int64_t idx; } key_t2;
static int key_cmp(const void* a, const void* b) {
    const key_t2 *p = (const key_t2*)a, *q = (const key_t2*)b;
    if (p->cost < q->cost) return -1;
    if (p->cost > q->cost) return 1;
    return (p->idx > q->idx) - (p->idx < q->idx);
}

/*
 * states: [14][n][Nt] or NULL; costs [n][K]; total [n]; flags [n]; traj_len [n]; margins [n] or NULL
 */
int orc_plan(const orc_params* P, int64_t n, const double* sampling, int M, const double* ref_pos,
             const double* ref_theta, const double* ref_curv, const double* ref_curv_d, const double* ref_x,
             const double* ref_y, int nT, const double* Tvals, const int32_t* Tlen, const double* tpow, int O, int T,
             const double* pos, const double* cov, const double* otheta, const double* ohl, const double* ohw,
             const int32_t* olen, int n_obs_pos, const double* obs_pos, int B, const double* sobb5, double* states,
             double* costs, double* total, uint32_t* flags, int32_t* traj_len, double* margins, orc_result* res,
             int nthreads) {
    const int Nt = P->N + 1, K = P->n_costs;
    if (Nt > MAXNT || K > 10) return -1;
    env_t E;
    memset(&E, 0, sizeof(E));
    E.p = P; E.M = M; E.rp = ref_pos; E.rth = ref_theta; E.rc = ref_curv; E.rcd = ref_curv_d; E.rx = ref_x; E.ry = ref_y;
    E.nT = nT; E.Tvals = Tvals; E.Tlen = Tlen; E.tpow = tpow;
    E.O = O; E.T = T; E.pos = pos; E.otheta = otheta; E.ohl = ohl; E.ohw = ohw; E.olen = olen;
    E.n_obs_pos = n_obs_pos; E.obs_pos = obs_pos; E.B = B;
    E.kappa_max = tan(P->delta_max) / P->wheelbase;
    double* icov = NULL; hull_t* ohull = NULL; hull_t* sh = NULL;
    if (O > 0) {
        icov = (double*)malloc(sizeof(double) * (size_t)O * T * 4);
        ohull = (hull_t*)calloc((size_t)O * T, sizeof(hull_t));
        for (size_t q = 0; q < (size_t)O * T; q++) {   /* np.linalg.inv of each 2x2 (closed form) */
            double a = cov[q * 4], b = cov[q * 4 + 1], c = cov[q * 4 + 2], d = cov[q * 4 + 3];
            double det = a * d - b * c;
            icov[q * 4] = d / det; icov[q * 4 + 1] = -b / det; icov[q * 4 + 2] = -c / det; icov[q * 4 + 3] = a / det;
        }
        for (int o = 0; o < O; o++)
            for (int t = 0; t + 1 < T; t++) {
                size_t q = (size_t)o * T + t;
                ohull[q] = obb_sum_hull(pos[q * 2], pos[q * 2 + 1], otheta[q], pos[(q + 1) * 2], pos[(q + 1) * 2 + 1],
                                        otheta[q + 1], ohl[o], ohw[o]);
            }
    }
    if (B > 0) {
        sh = (hull_t*)malloc(sizeof(hull_t) * B);
        for (int b = 0; b < B; b++) {
            sh[b].cx = sobb5[b * 5]; sh[b].cy = sobb5[b * 5 + 1];
            sh[b].ux = cos(sobb5[b * 5 + 2]); sh[b].uy = sin(sobb5[b * 5 + 2]);
            sh[b].ha = sobb5[b * 5 + 3]; sh[b].hb = sobb5[b * 5 + 4];
        }
    }
    E.icov = icov; E.ohull = ohull; E.sobb = sh;
    const int do_pred = P->collision_check && O > 0;
    const int check_all = P->check_all_collisions && (do_pred || B > 0);
    const int keep_xyth = !check_all && (do_pred || B > 0) && states == NULL;
    double* xyth = keep_xyth ? (double*)malloc(sizeof(double) * (size_t)n * 3 * Nt) : NULL;

#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        double st[NF * MAXNT], cl[6], ct[6];
        hull_t eh[MAXNT];
#ifdef _OPENMP
#pragma omp single nowait
        orc_threads_used = omp_get_num_threads();
#endif
#pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < n; r++) {
            int tl = 0; double mg = 0;
            uint32_t fl = eval_candidate(&E, sampling + r * 13, st, cl, ct, &tl, &mg);
            const int valid = fl & FLAG_VALID, feas = (fl & FLAG_FEASIBLE) != 0, in_list = (fl & FLAG_IN_LIST) != 0,
                      stored = (fl & FLAG_STORED) != 0;
            const int costed = P->draw ? in_list : (in_list && valid && feas && stored);
            const int cand = P->draw ? (in_list && feas) : costed;
            double tot = 0.0;
            for (int k = 0; k < K; k++) costs[r * K + k] = 0.0;
            if (costed) { eval_costs(&E, st, cl, ct, costs + r * K, &tot); fl |= FLAG_COSTED; }
            if (cand) {
                fl |= FLAG_CANDIDATE;
                if (check_all) {
                    ego_hulls(&E, st, eh);
                    const int hk = do_pred ? collides_predictions(&E, eh) : -1, bk = (B > 0) ? collides_static(&E, eh) : -1;
                    if (hk >= 0) fl |= COLLIDE_BITS(hk);
                    if (bk >= 0) fl |= BOUNDARY_BITS(bk);
                }
            }
            total[r] = tot; flags[r] = fl; traj_len[r] = tl;
            if (margins) margins[r] = mg;
            if (states)
                for (int f = 0; f < NF; f++) memcpy(states + ((size_t)f * n + r) * Nt, st + f * Nt, sizeof(double) * Nt);
            if (xyth) {
                memcpy(xyth + (size_t)r * 3 * Nt, st + F_X * Nt, sizeof(double) * Nt);
                memcpy(xyth + (size_t)r * 3 * Nt + Nt, st + F_Y * Nt, sizeof(double) * Nt);
                memcpy(xyth + (size_t)r * 3 * Nt + 2 * Nt, st + F_THETA * Nt, sizeof(double) * Nt);
            }
        }
    }

    /* statistics (reactive_planner.py:229-235) */
    memset(res, 0, sizeof(*res));
    res->argmin = -1; res->min_cost = INFINITY;
    int64_t n_cand = 0;
    for (int64_t r = 0; r < n; r++) {
        uint32_t fl = flags[r];
        if (fl & FLAG_IN_LIST) {
            res->n_in_list++;
            if ((fl & FLAG_VALID) && (fl & FLAG_FEASIBLE)) res->n_feasible++; else res->reason_counts[0]++;
        }
        for (int q = 1; q <= 10; q++) if (fl & FLAG_REASON(q)) res->reason_counts[q]++;
        if (fl & FLAG_CANDIDATE) n_cand++;
    }
    res->n_candidates = n_cand;

    /* selection: walk the cost-sorted candidate list (trajectories.py:524-561, planner.py:329-392) */
    key_t2* keys = (key_t2*)malloc(sizeof(key_t2) * (size_t)(n_cand > 0 ? n_cand : 1));
    int64_t m = 0;
    for (int64_t r = 0; r < n; r++)
        if (flags[r] & FLAG_CANDIDATE) { keys[m].cost = total[r]; keys[m].idx = r; m++; }
    qsort(keys, (size_t)m, sizeof(key_t2), key_cmp);
    hull_t eh[MAXNT];
    double stx[NF * MAXNT];
    for (int64_t q = 0; q < m; q++) {
        int64_t r = keys[q].idx;
        int hit, off;
        if (check_all || !(do_pred || B > 0)) {
            hit = (flags[r] & FLAG_COLLIDE) != 0; off = (flags[r] & FLAG_BOUNDARY) != 0;
        } else {
            memset(stx, 0, sizeof(stx));
            if (states) {
                memcpy(stx + F_X * Nt, states + ((size_t)F_X * n + r) * Nt, sizeof(double) * Nt);
                memcpy(stx + F_Y * Nt, states + ((size_t)F_Y * n + r) * Nt, sizeof(double) * Nt);
                memcpy(stx + F_THETA * Nt, states + ((size_t)F_THETA * n + r) * Nt, sizeof(double) * Nt);
            } else {
                memcpy(stx + F_X * Nt, xyth + (size_t)r * 3 * Nt, sizeof(double) * Nt);
                memcpy(stx + F_Y * Nt, xyth + (size_t)r * 3 * Nt + Nt, sizeof(double) * Nt);
                memcpy(stx + F_THETA * Nt, xyth + (size_t)r * 3 * Nt + 2 * Nt, sizeof(double) * Nt);
            }
            ego_hulls(&E, stx, eh);
            const int hk = do_pred ? collides_predictions(&E, eh) : -1, bk = (B > 0) ? collides_static(&E, eh) : -1;
            hit = hk >= 0; off = bk >= 0;
            if (hit) flags[r] |= COLLIDE_BITS(hk);
            if (off) flags[r] |= BOUNDARY_BITS(bk);
        }
        if (hit) res->collision_counter++;
        if (!hit && !off) { res->argmin = r; res->min_cost = total[r]; break; }
    }
    free(keys); free(icov); free(ohull); free(sh); free(xyth);
    return 0;
}

/* exported for tests: the claim "np_sum is numpy's summation order" is checked bit-for-bit */
double orc_np_sum(const double* a, int n) { return np_sum(a, n); }
double orc_simps(const double* y, int n, double dx) { return simps(y, n, dx); }

/* threads the last orc_plan actually ran on (bench.py asserts the CPU arm got every host thread) */
int orc_last_threads(void) { return orc_threads_used; }

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
