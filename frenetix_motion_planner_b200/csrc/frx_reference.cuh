// frx_reference.cuh -- the step in front of the hot path, on the device (included by frx_kernels.cu):
//   * frx_reference_tables_kernel: the six reference tables from a polyline -- CoordinateSystem.__init__
//     (cr_scenario_handler/utils/utils_coordinate_system.py:203-207: ref_pos, ref_curv, ref_theta = np.unwrap(...),
//     ref_curv_d = np.gradient(ref_curv, ref_pos)) with the polyline helpers of SURVEY.md A.0 (cumulative chord length,
//     heading of the outgoing segment with the last one repeated, index-based np.gradient curvature);
//   * frx_initial_state_kernel: Cartesian ego state -> Frenet state x_cl = ([s, s', s''], [d, d', d'']),
//     Planner._compute_initial_states (frenetix_motion_planner/planner.py:567-635), including the projection
//     (x, y) -> (s, d) the reference asks CCosy for: the exact inverse of the Frenet -> Cartesian map of the eval kernel
//     (closest segment in parallel, then Newton on the tangential miss -- coordinate_system.py has the same iteration).
// One CTA each: the work is O(M) for M <= a few thousand vertices, once per reference path / once per planning step; what
// matters is that the tables never leave the device and that a host without numpy/scipy/CCosy can feed the planner a
// polyline and a Cartesian state.  Operation order follows numpy so that the tables agree with the host ones to the last
// few ulp (atan2 / pow come from different math libraries).
#pragma once

#define FRX_REF_THREADS 256

// numpy's float mod (sign of the divisor)
__device__ __forceinline__ double frx_np_mod(double a, double b) {
    double m = fmod(a, b);
    if (m != 0.0 && ((m < 0.0) != (b < 0.0))) m += b;
    return m;
}

// np.gradient(f) with unit spacing at index i
__device__ __forceinline__ double frx_grad1(const double* f, int i, int M) {
    if (i == 0) return f[1] - f[0];
    if (i == M - 1) return f[M - 1] - f[M - 2];
    return (f[i + 1] - f[i - 1]) / 2.0;
}

// tables: [6][Mpad] = pos, theta, curv, curv_d, x, y ; scratch: [4][M] (x', y', curv is written to tables directly)
__global__ void __launch_bounds__(FRX_REF_THREADS)
frx_reference_tables_kernel(int M, int Mpad, const double* __restrict__ xy, double* __restrict__ tab, double* __restrict__ scratch) {
    double* pos = tab; double* theta = tab + Mpad; double* curv = tab + 2 * Mpad; double* curv_d = tab + 3 * Mpad;
    double* x = tab + 4 * Mpad; double* y = tab + 5 * Mpad;
    double* xd = scratch; double* yd = scratch + M; double* seg = scratch + 2 * M; double* raw = scratch + 3 * M;
    const int t = threadIdx.x;
    for (int i = t; i < Mpad; i += FRX_REF_THREADS) {
        const bool in = i < M;
        x[i] = in ? xy[2 * i] : 0.0; y[i] = in ? xy[2 * i + 1] : 0.0;
        if (!in) { pos[i] = theta[i] = curv[i] = curv_d[i] = 0.0; }
    }
    __syncthreads();
    for (int i = t; i < M - 1; i += FRX_REF_THREADS) {
        const double dx = x[i + 1] - x[i], dy = y[i + 1] - y[i];
        seg[i] = sqrt(dx * dx + dy * dy);              // np.sqrt(np.sum(np.diff(p) ** 2, axis=1))
        raw[i] = atan2(dy, dx);                         // heading of the outgoing segment
    }
    for (int i = t; i < M; i += FRX_REF_THREADS) { xd[i] = frx_grad1(x, i, M); yd[i] = frx_grad1(y, i, M); }
    __syncthreads();
    if (t == 0) {
        // np.cumsum and np.unwrap are sequential by definition (same association as numpy)
        double acc = 0.0;
        pos[0] = 0.0;
        for (int i = 0; i < M - 1; ++i) { acc += seg[i]; pos[i + 1] = acc; }
        raw[M - 1] = raw[M - 2];                        // the last vertex repeats the previous heading
        const double pi = 3.141592653589793, two_pi = 6.283185307179586;
        double corr = 0.0;
        theta[0] = raw[0];
        for (int i = 1; i < M; ++i) {
            const double dd = raw[i] - raw[i - 1];
            double ddmod = frx_np_mod(dd + pi, two_pi) - pi;
            if (ddmod == -pi && dd > 0) ddmod = pi;
            double ph = ddmod - dd;
            if (fabs(dd) < pi) ph = 0.0;
            corr += ph;
            theta[i] = raw[i] + corr;
        }
    }
    for (int i = t; i < M; i += FRX_REF_THREADS) {
        const double xdd = frx_grad1(xd, i, M), ydd = frx_grad1(yd, i, M);
        const double q = xd[i] * xd[i] + yd[i] * yd[i];
        curv[i] = (xd[i] * ydd - xdd * yd[i]) / pow(q, 1.5);
    }
    __syncthreads();
    for (int i = t; i < M; i += FRX_REF_THREADS) {       // np.gradient(curv, pos): second-order, non-uniform spacing
        double g;
        if (i == 0) g = (curv[1] - curv[0]) / (pos[1] - pos[0]);
        else if (i == M - 1) g = (curv[M - 1] - curv[M - 2]) / (pos[M - 1] - pos[M - 2]);
        else {
            const double hs = pos[i] - pos[i - 1], hd = pos[i + 1] - pos[i];
            const double a = -(hd) / (hs * (hd + hs)), b = (hd - hs) / (hd * hs), c = hs / (hd * (hd + hs));
            g = a * curv[i - 1] + b * curv[i] + c * curv[i + 1];
        }
        curv_d[i] = g;
    }
}

// in: x, y, orientation, velocity, acceleration, steering angle ; out: s, s', s'', d, d', d'', status (0 ok, 1 s' < 0)
__global__ void __launch_bounds__(FRX_REF_THREADS)
frx_initial_state_kernel(int M, int Mpad, const double* __restrict__ tab, const double* __restrict__ in, double wheelbase,
                         int low_vel_mode, double* __restrict__ out) {
    const double* pos = tab; const double* theta = tab + Mpad; const double* curv = tab + 2 * Mpad;
    const double* curv_d = tab + 3 * Mpad; const double* rx = tab + 4 * Mpad; const double* ry = tab + 5 * Mpad;
    const double X = in[0], Y = in[1];
    __shared__ double s_d2[FRX_REF_THREADS];
    __shared__ int s_i[FRX_REF_THREADS];
    // ---- closest segment (orthogonal projection clipped to the segment), lowest index on ties like np.argmin
    double best = __longlong_as_double(0x7ff0000000000000LL);
    int bi = 0;
    for (int i = threadIdx.x; i < M - 1; i += FRX_REF_THREADS) {
        const double abx = rx[i + 1] - rx[i], aby = ry[i + 1] - ry[i];
        double lam = ((X - rx[i]) * abx + (Y - ry[i]) * aby) / (abx * abx + aby * aby);
        lam = fmin(fmax(lam, 0.0), 1.0);
        const double qx = rx[i] + lam * abx, qy = ry[i] + lam * aby;
        const double d2 = (X - qx) * (X - qx) + (Y - qy) * (Y - qy);
        if (d2 < best) { best = d2; bi = i; }
    }
    s_d2[threadIdx.x] = best; s_i[threadIdx.x] = bi;
    __syncthreads();
    for (int off = FRX_REF_THREADS / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            const double o = s_d2[threadIdx.x + off];
            const int oi = s_i[threadIdx.x + off];
            if (o < s_d2[threadIdx.x] || (o == s_d2[threadIdx.x] && oi < s_i[threadIdx.x])) { s_d2[threadIdx.x] = o; s_i[threadIdx.x] = oi; }
        }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    int i = s_i[0];
    double lam;
    {
        const double abx = rx[i + 1] - rx[i], aby = ry[i + 1] - ry[i];
        lam = fmin(fmax(((X - rx[i]) * abx + (Y - ry[i]) * aby) / (abx * abx + aby * aby), 0.0), 1.0);
    }
    // ---- Newton on f(lam) = (X - P(lam)) . t(theta(lam)): foot point of the INTERPOLATED-heading normal
    for (int it = 0; it < 50; ++it) {
        const double abx = rx[i + 1] - rx[i], aby = ry[i + 1] - ry[i], dth = theta[i + 1] - theta[i];
        const double th = theta[i] + lam * dth;
        double sn, cs;
        sincos(th, &sn, &cs);
        const double rxv = X - (rx[i] + lam * abx), ryv = Y - (ry[i] + lam * aby);
        const double f = rxv * cs + ryv * sn;
        const double df = -(abx * cs + aby * sn) + (ryv * cs - rxv * sn) * dth;
        const double step = (df != 0.0) ? f / df : 0.0;
        double ln = lam - step;
        if (ln < 0.0 && i > 0) { i -= 1; lam = 1.0 - 1e-12; continue; }
        if (ln >= 1.0 && i < M - 2) { i += 1; lam = 0.0; continue; }
        ln = fmin(fmax(ln, 0.0), 1.0);
        const bool done = fabs(ln - lam) <= 1e-15;
        lam = ln;
        if (done) break;
    }
    const double th_ref_raw = theta[i] + lam * (theta[i + 1] - theta[i]);
    double sn, cs;
    sincos(th_ref_raw, &sn, &cs);
    const double s = pos[i] + lam * (pos[i + 1] - pos[i]);
    const double d = (Y - (ry[i] + lam * (ry[i + 1] - ry[i]))) * cs - (X - (rx[i] + lam * (rx[i + 1] - rx[i]))) * sn;
    // ---- planner.py:580-630 (segment of s by the reference's own search: first table entry above s, minus one)
    int j = first_greater(pos, M, s, pos[0], (double)(M - 1) / (pos[M - 1] - pos[0]));
    const int ia = (j == 0) ? (M - 1) : (j - 1);
    const double s_lambda = (s - pos[ia]) / (pos[j] - pos[ia]);
    const double theta_ref = make_valid_orientation((theta[j] - theta[ia]) * (s - pos[ia]) / (pos[j] - pos[ia]) + theta[ia]);
    const double theta_cl = in[2] - theta_ref;
    const double kr = (curv[j] - curv[ia]) * s_lambda + curv[ia];
    const double kr_d = (curv_d[j] - curv_d[ia]) * s_lambda + curv_d[ia];
    const double kappa_0 = tan(in[5]) / wheelbase;
    const double tn = tan(theta_cl), c = cos(theta_cl);
    const double one = 1 - kr * d;
    const double d_p = one * tn;
    const double d_pp = -(kr_d * d + kr * d_p) * tn + (one / (c * c)) * (kappa_0 * one / c - kr);
    const double s_velocity = in[3] * c / one;
    double s_acc = in[4];
    s_acc -= (s_velocity * s_velocity / c) * (one * tn * (kappa_0 * one / c - kr) - (kr_d * d + kr * d_p));
    s_acc /= (one / c);
    double d_velocity, d_acc;
    if (low_vel_mode) { d_velocity = d_p; d_acc = d_pp; }
    else { d_velocity = in[3] * sin(theta_cl); d_acc = s_acc * d_p + s_velocity * s_velocity * d_pp; }
    out[0] = s; out[1] = s_velocity; out[2] = s_acc; out[3] = d; out[4] = d_velocity; out[5] = d_acc;
    out[6] = (s_velocity < 0) ? 1.0 : 0.0;
}

void frx_launch_reference_tables(int M, int Mpad, const double* xy, double* tab, double* scratch, cudaStream_t st) {
    frx_reference_tables_kernel<<<1, FRX_REF_THREADS, 0, st>>>(M, Mpad, xy, tab, scratch);
}
void frx_launch_initial_state(int M, int Mpad, const double* tab, const double* in, double wheelbase, int low, double* out,
                              cudaStream_t st) {
    frx_initial_state_kernel<<<1, FRX_REF_THREADS, 0, st>>>(M, Mpad, tab, in, wheelbase, low, out);
}
