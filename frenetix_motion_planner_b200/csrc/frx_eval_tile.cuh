// frx_eval_tile.cuh -- the eval kernel body: ONE THREAD PER CANDIDATE TRAJECTORY, one warp per tile of 32
// consecutive sampling rows (included by frx_kernels.cu, which provides the arithmetic helpers).
//
//   * SEG lanes per candidate (SEG = 1, 2 or 4, picked from the row count): a warp works on a tile of C = 32 / SEG
//     consecutive rows, lane = (segment, candidate), each lane walks its own contiguous segment of the time steps.
//     With few rows (50,000 rows are only 1,563 warps of 32) the extra warps are what hides the fp64 dependency
//     chains; with many rows SEG = 1 does the least redundant work.  Segment boundaries: the yaw-rate / curvature-
//     rate gates and kappa diff of a segment's first step need the previous step, which another lane computes --
//     they are evaluated after the loop from shuffled values; the stand-still heading carry is resolved in the
//     prologue from the memo (which steps are stand-still depends on s(t) only);
//   * lane = candidate: the polynomial coefficient solve, the per-candidate bookkeeping and the cost assembly are
//     per-thread scalar code (useful work in all 32 lanes), the time-step loop runs sequentially in each thread --
//     time-coupled quantities (yaw rate, curvature rate, stand-still heading carry) are plain registers, no
//     shuffles, ballots or warp-synchronous code inside the loop, so lanes may diverge freely;
//   * the state tensor is laid out in blocks of 32 candidates, [block][step][field][32]: at every step the 32 lanes of a
//     warp write 32 consecutive candidates of one field = one fully coalesced 256-byte store, and the 14 fields of the
//     step are one contiguous 3.5 KB span (base pointer + immediate offsets);
//   * everything that depends on the longitudinal motion s(t) only (quartic, samples, reference segment, lambda,
//     interpolated heading/curvature, foot point, normal, time-power row) is shared by all candidates with the same
//     (t1, s0, ss0, sss0, ss1): the warp computes it cooperatively ONCE per key (lane = time step, the reference
//     tables in shared memory staged by a TMA bulk copy) into a per-warp shared-memory memo of two slots; rows of a
//     sampling matrix come as a cartesian product with d1 fastest, so a tile spans one or two keys (tiles with more
//     are processed in several passes of two keys each -- correct for ANY matrix, fast for cartesian ones);
//   * obstacle data (prediction cost, collision sweep) is indexed by the step only -> warp-uniform loads;
//     the second pass over the steps runs only for the lanes that need it and re-reads x, y, theta of the candidate
//     from the state tensor (coalesced, L2).
//
// Arithmetic follows reactive_planner.py:274-577 of the reference op for op (comments next to each block).
#pragma once
// FRX_TRACE (tuning builds only): every warp records globaltimer stamps of its first tile into A.trace[warp][8]
#ifndef FRX_TRACE
#define FRX_TRACE 0
#endif
#if FRX_TRACE
__device__ __forceinline__ unsigned long long frx_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define FRX_STAMP(k) do { if (A.trace && lane == 0 && !traced) A.trace[((size_t)cta_local * FRX_WARPS_PER_CTA + wib) * 16 + (k)] = frx_now(); } while (0)
#else
#define FRX_STAMP(k) do { } while (0)
#endif
#if FRX_TRACE
#define FRX_TRW ((A.trace && !traced) ? A.trace + ((size_t)cta_local * FRX_WARPS_PER_CTA + wib) * 16 : nullptr)
#else
#define FRX_TRW nullptr
#endif

enum { M_S = 0, M_SD, M_SDD, M_INTERP, M_KR, M_KRD, M_PX, M_PY, M_SN, M_CS, M_FIELDS };
// rows of the time-power table t, t^2 .. t^5 (shared by the CTA: the rounded powers of step i do not depend on the duration,
// np.arange(0, T + dt, dt) is 0 + i * dt for every T -- only the number of samples does; frx_set_time_tables checks it)
enum { T_1 = 0, T_2, T_3, T_4, T_5, T_ROWS };

struct FrxMemoHdr {        // one per memo slot (shared memory)
    double key[5];         // t1, s0, ss0, sss0, ss1 the slot was filled for
    double s_first, jerk_lon, goal;
    int traj_len, first_none, bits, valid;   // bits: 1 any s_d < -eps, 2 any |s_dd| > a_max, 4 duration not registered
};

#define FRX_MEMO_SLOTS 2
#define FRX_WALL_LIST 512      // static boxes a tile-level cull list can hold (more: every box is tested)

__host__ __device__ inline int frx_memo_pitch(int Nt) { return (Nt + 3) & ~3; }
__host__ __device__ inline size_t frx_tile_smem_bytes(int Mpad, int tpitch, int mpitch) {
    return (size_t)(6 * Mpad + FRX_MAX_T_VALUES + T_ROWS * tpitch + FRX_WARPS_PER_CTA * FRX_MEMO_SLOTS * M_FIELDS * mpitch) * sizeof(double) +
           FRX_WARPS_PER_CTA * FRX_MEMO_SLOTS * sizeof(FrxMemoHdr) + 16 + FRX_WARPS_PER_CTA * sizeof(FrxBest) +
           (size_t)FRX_WARPS_PER_CTA * 32 * 13 * sizeof(double);
}

// Simpson-rule accumulator (scipy simps, dx = dt; partial_cost_functions.py:24-46, :141-151, :189-196), fed one
// sample at a time: `part` is the plain composite sum over the first nb samples, `corr` the last-interval correction
// scipy applies when the number of samples is even.
struct FrxSimpson {
    double part, corr;
    __device__ __forceinline__ void add(int jx, int n, int nb, double yv, double alpha, double beta, double eta) {
        if (jx >= 0 && jx < n) {
            if (jx < nb) {
                double wgt = (jx == 0 || jx == nb - 1) ? 1.0 : ((jx & 1) ? 4.0 : 2.0);
                part += wgt * yv;
            }
            if (!(n & 1) && n > 2) {
                if (jx == n - 1) corr += alpha * yv;
                else if (jx == n - 2) corr += beta * yv;
                else if (jx == n - 3) corr -= eta * yv;
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------------------
// memo fill: warp-cooperative, lane = time step (chunks of 32).  reactive_planner.py:296-322, :350-355, :415-420,
// :457-460, :536-547; polynomial_trajectory.py:452-488 (quartic, closed form)
// ------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void frx_memo_fill(const FrxKernelArgs& A, const double* __restrict__ s_ref, const double* __restrict__ s_Ttab,
                                           const int* __restrict__ s_Tlen, const double* __restrict__ s_tp, unsigned long long* trw,
                                           double* __restrict__ mt, FrxMemoHdr* __restrict__ hdr, const double T, const double s0,
                                           const double ss0, const double sss0, const double ss1) {
    const int lane = threadIdx.x & 31;
    const int Mpad = A.Mpad, M = A.M, Nt = A.Nt, TP = A.tpitch, MP = A.mpitch;
    const double* __restrict__ rp = s_ref;
    const double* __restrict__ rth = s_ref + Mpad;
    const double* __restrict__ rc = s_ref + 2 * Mpad;
    const double* __restrict__ rcd = s_ref + 3 * Mpad;
    const double* __restrict__ rx = s_ref + 4 * Mpad;
    const double* __restrict__ ry = s_ref + 5 * Mpad;
    const double dT = A.dt;
    const double pos_first = rp[0], pos_last = rp[M - 1];
    const double inv_step = A.inv_step;       // (M - 1) / (pos_last - pos_first), computed by the host with the same IEEE operations
    __syncwarp();
    if (lane == 0) {
        hdr->key[0] = T; hdr->key[1] = s0; hdr->key[2] = ss0; hdr->key[3] = sss0; hdr->key[4] = ss1;
        hdr->valid = 1;
    }
#if FRX_TRACE
#define FRX_MSTAMP(k) do { if (trw && lane == 0) trw[k] = frx_now(); } while (0)
#else
#define FRX_MSTAMP(k) do { } while (0)
#endif
    FRX_MSTAMP(8);
    // ---------------- time table of this duration (reactive_planner.py:296-303)
    int tix = -1;
    for (int b0 = 0; b0 < A.nT; b0 += 32) {
        unsigned m = __ballot_sync(FULL, (b0 + lane < A.nT) && (s_Ttab[b0 + lane] == T));
        if (m) { tix = b0 + __ffs(m) - 1; break; }
    }
    if (tix < 0) {   // the host did not register this duration: the rows of this key are reported dead
        if (lane == 0) { hdr->bits = 4; hdr->traj_len = 0; hdr->first_none = 0; hdr->s_first = hdr->jerk_lon = hdr->goal = 0.0; }
        __syncwarp();
        return;
    }
    const int traj_len = s_Tlen[tix];      // (shared memory: a global load here would put a second round trip in front of
                                           // the loads at index traj_len - 1 below)
    Poly L;
    {
        double T2 = T * T, T3 = T2 * T;
        double b0 = (ss1 - ss0) - sss0 * T;
        double b1 = -sss0;
        L.c0 = s0; L.c1 = ss0; L.c2 = sss0 * 0.5;   // == sss0 / 2.0 exactly
        L.c3 = ddivf(3 * b0 - T * b1, 3 * T2);
        L.c4 = ddivf(T * b1 - 2 * b0, 4 * T3);
        L.c5 = 0.0;
    }
    const int il = traj_len - 1;
    double s_last = 0.0, sd_last = 0.0, s_inc = 0.0;
    const double s_first = poly_pos(L, s_tp[T_1 * TP], s_tp[T_2 * TP], s_tp[T_3 * TP], s_tp[T_4 * TP], s_tp[T_5 * TP]);
    if (traj_len < Nt) {   // values of the last polynomial sample feed the extension of every later step
        double tl = s_tp[T_1 * TP + il], tl2 = s_tp[T_2 * TP + il], tl3 = s_tp[T_3 * TP + il], tl4 = s_tp[T_4 * TP + il],
               tl5 = s_tp[T_5 * TP + il];
        s_last = poly_pos(L, tl, tl2, tl3, tl4, tl5);
        sd_last = poly_vel(L, tl, tl2, tl3, tl4);
        s_inc = dT * sd_last;
    }
    FRX_MSTAMP(9);       // coefficients, s_first, s_last (time-table loads) done
    bool any_neg = false, any_acc = false;
    int first_none = Nt;
    for (int c0 = 0; c0 < TP; c0 += 32) {
        const int i = c0 + lane;
        const bool act = i < Nt;
        double vs = 0, vsd = 0, vsdd = 0;
        {
            const double t = s_tp[T_1 * TP + i], t2 = s_tp[T_2 * TP + i], t3 = s_tp[T_3 * TP + i], t4 = s_tp[T_4 * TP + i],
                         t5 = s_tp[T_5 * TP + i];
            if (i < traj_len) {
                vs = poly_pos(L, t, t2, t3, t4, t5);
                vsd = poly_vel(L, t, t2, t3, t4);
                vsdd = poly_acc(L, t, t2, t3);
            }
        }
        if (i >= traj_len && act) {
            vs = s_last;                        // s[ext] = s[ext-1] + dt * s_velocity[traj_len-1]
            for (int k = il; k < i; ++k) vs += s_inc;
            vsd = sd_last; vsdd = 0.0;
        }
        FRX_MSTAMP(10);  // samples + extension
        any_neg |= __any_sync(FULL, act && (vsd < -FRX_EPS));
        any_acc |= __any_sync(FULL, act && (fabs(vsdd) > A.a_max));
        if (fabs(vsd) < FRX_EPS) vsd = 0.0;     // :355
        // :415-420 segment lookup (python negative-index wrap reproduced), :457-460 curvature
        int j = first_greater(rp, M, vs, pos_first, inv_step);
        int ia = (j == 0) ? (M - 1) : (j - 1);
        FRX_MSTAMP(11);  // segment search
        double pa = rp[ia], pb = rp[j];
        double lam = ddivf(vs - pa, pb - pa);
        double tha = rth[ia], thb = rth[j];
        double interp = make_valid_orientation(ddivf((thb - tha) * (vs - pa), pb - pa) + tha);
        double k_r = (rc[j] - rc[ia]) * lam + rc[ia];
        double k_r_d = (rcd[j] - rcd[ia]) * lam + rcd[ia];
        // :536-547 foot point and normal of the Cartesian conversion (library definition of CCosy)
        bool none = !(vs >= pos_first) || !(vs < pos_last);
        unsigned nm = __ballot_sync(FULL, none && act);
        if (nm && first_none == Nt) first_none = c0 + __ffs(nm) - 1;
        double px = (1.0 - lam) * rx[ia] + lam * rx[j];
        double py = (1.0 - lam) * ry[ia] + lam * ry[j];
        double thr = tha + lam * (thb - tha);
        FRX_MSTAMP(12);  // interpolation
        double sn, cs;
        sincos(thr, &sn, &cs);
        FRX_MSTAMP(13);  // sincos
        if (i < MP) {
            mt[M_S * MP + i] = vs; mt[M_SD * MP + i] = vsd; mt[M_SDD * MP + i] = vsdd;
            mt[M_INTERP * MP + i] = interp;
            mt[M_KR * MP + i] = k_r; mt[M_KRD * MP + i] = k_r_d;
            mt[M_PX * MP + i] = px; mt[M_PY * MP + i] = py;
            mt[M_SN * MP + i] = sn; mt[M_CS * MP + i] = cs;
        }
    }
    if (lane == 0) {
        hdr->traj_len = traj_len;
        hdr->first_none = first_none;
        hdr->bits = (any_neg ? 1 : 0) | (any_acc ? 2 : 0);
        hdr->s_first = s_first;
        hdr->jerk_lon = sq_jerk_integral(L, dT);
        // reactive_planner.py:161-166 (evaluate_state_at_tau at tau = delta_tau), used in low-velocity mode
        double t2 = T * T, t3 = t2 * T, t4 = t2 * t2, t5 = t3 * t2;
        hdr->goal = poly_pos(L, T, t2, t3, t4, t5) - s0;
    }
    __syncwarp();
    FRX_MSTAMP(14);
}

// curvature-rate limit (:513-521): the Python path's constant 0.4, or (use_cpp flavour) v_delta_max / (wheelbase cos^2(delta))
// with delta = atan(wheelbase kappa), i.e. (v_delta_max / wheelbase) (1 + (wheelbase kappa)^2)
__device__ __forceinline__ double frx_kappa_dot_max(const FrxKernelArgs& A, const double kappa) {
    if (!A.kd_from_v_delta) return 0.4;
    const double wk = A.wheelbase * kappa;
    return A.v_delta_over_wb * (1.0 + wk * wk);
}

// ------------------------------------------------------------------------------------------------------------
// one candidate, one thread
// ------------------------------------------------------------------------------------------------------------
struct FrxLaneOut {
    unsigned ev;         // event bits (CNT_* order) of this candidate
    double total;
    bool winner_ok;      // candidate && !collide && !boundary
    bool t_missing;
};

// `pass` = the lanes executing this call together (all SEG lanes of a candidate are among them): the mask of every
// shuffle below, which all of them reach -- there is no early return.
template <int SEG, bool OBS, bool XCOST>
__device__ __forceinline__ FrxLaneOut frx_candidate(const FrxKernelArgs& A, const unsigned cost_mask, const unsigned pass,
                                                    const long long r, const double T, const double d0, const double dd0,
                                                    const double ddd0, const double d1, const double dd1, const double ddd1,
                                                    const double* __restrict__ mt, const FrxMemoHdr* __restrict__ H,
                                                    const double* __restrict__ s_tp, unsigned short* __restrict__ wall_list) {
    constexpr int C = 32 / SEG;                 // candidates per tile
    const int lane = threadIdx.x & 31;
    const int cand = lane & (C - 1), seg = lane / C;
    FrxLaneOut out;
    out.ev = 0; out.total = 0.0; out.winner_ok = false; out.t_missing = false;
    const int Nt = A.Nt, TP = A.tpitch, MP = A.mpitch;
    const double dT = A.dt;
    const bool low = A.low != 0, draw = A.draw != 0, debug = A.debug != 0;
    const bool brk = !draw && !debug;
    const int hbits = H->bits;
    const bool dead = (hbits & 4) != 0;         // duration missing from the time tables: reported, row marked dead
    // this lane's segment of the time steps
    const int seglen = (Nt + SEG - 1) / SEG;
    const int i0 = seg * seglen;
    const int i1 = dead ? i0 : ((i0 + seglen < Nt) ? (i0 + seglen) : Nt);
    const int traj_len = H->traj_len;
    const int il = traj_len - 1;
    const bool any_neg = (hbits & 1) != 0, any_acc = (hbits & 2) != 0;
    const double s_first = H->s_first;

    // ---------------- lateral quintic (polynomial_trajectory.py:293-343; closed form)
    Poly Q;
    {
        double tau = T;
        if (low) { const double goal = H->goal; tau = (goal <= 0) ? T : goal; }
        double u2 = tau * tau, u3 = u2 * tau, u4 = u2 * u2, u5 = u4 * tau;
        double b0 = ((d1 - d0) - dd0 * tau) - (.5 * ddd0) * u2;
        double b1 = (dd1 - dd0) - ddd0 * tau;
        double b2 = ddd1 - ddd0;
        Q.c0 = d0; Q.c1 = dd0; Q.c2 = .5 * ddd0;
        Q.c3 = ddivf((10 * b0 - (4 * b1) * tau) + (0.5 * b2) * u2, u3);
        Q.c4 = ddivf((-15 * b0 + (7 * b1) * tau) - b2 * u2, u4);
        Q.c5 = ddivf((6 * b0 - (3 * b1) * tau) + (0.5 * b2) * u2, u5);
    }
    double d_last = 0.0;
    if (traj_len < Nt && !dead) {
        if (!low) {
            d_last = poly_pos(Q, s_tp[T_1 * TP + il], s_tp[T_2 * TP + il], s_tp[T_3 * TP + il], s_tp[T_4 * TP + il], s_tp[T_5 * TP + il]);
        } else {
            double q1 = mt[M_S * MP + il] - s_first, q2 = q1 * q1, q3 = q2 * q1, q4 = q2 * q2, q5 = q4 * q1;
            d_last = poly_pos(Q, q1, q2, q3, q4, q5);
        }
    }

    // ---------------- validity / pre-filter bookkeeping (:350-386)
    bool valid = !any_neg;
    bool feasible = true;
    uint32_t reasons = 0;
    bool in_list = true, stored = true;
    if (any_neg) {
        reasons |= FRX_FLAG_REASON(10);
        if (brk) { in_list = false; stored = false; }
    }
    if (in_list && !draw) {
        if (any_acc) { feasible = false; reasons |= FRX_FLAG_REASON(1); stored = false; }
        else if (any_neg) { feasible = false; reasons |= FRX_FLAG_REASON(2); stored = false; }
    }
    const bool evaluate = in_list && stored && !dead;    // reaches the per-step loop of :389

    // ---------------- the time-step loop: lateral samples (:325-346), back-projection + gates (:389-533), x/y
    //                  (:536-547), running cost sums, 14 coalesced stores per step
    uint32_t gate_or = 0;
    bool gate_hit = false;
    const int first_none = H->first_none;
    const bool seen_none = evaluate && (first_none < Nt);
    double th_prev = A.x0_orientation;   // theta_gl[i-1]
    double ka_prev = 0.0;
    // first step of a later segment: gates that need step i0 - 1 are completed after the loop
    uint32_t g_first = 0;
    double th_first = 0.0, ka_first = 0.0, vi_first = 0.0, a_first = 0.0, thc_first = 0.0;
    if (SEG > 1 && seg > 0 && evaluate && !low && i0 < i1 && !(mt[M_SD * MP + i0] > 0.001)) {
        // the segment starts in stand-still: theta_gl of the last moving step before it (:423-454), or x_0's
        for (int j = i0 - 1; j >= 0; --j) {
            const double sdj = mt[M_SD * MP + j];
            if (sdj > 0.001) {
                double ddj = 0.0;
                if (j < traj_len) ddj = poly_vel(Q, s_tp[T_1 * TP + j], s_tp[T_2 * TP + j], s_tp[T_3 * TP + j], s_tp[T_4 * TP + j]);
                th_prev = datan(ddivf(ddj, sdj)) + mt[M_INTERP * MP + j];
                break;
            }
        }
    }
    double vo_sum = 0.0, v_last = 0.0, dr_sum = 0.0, dr_last = 0.0;
    // bounding box of this lane's stored positions (fused obstacle pass: tile-level cull of the static boxes)
    float bb_x0 = 3e38f, bb_x1 = -3e38f, bb_y0 = 3e38f, bb_y1 = -3e38f;
    const int half = Nt / 2;
    // state tensor: blocks of 32 candidates, [block][step][field][32] -- the 14 fields of a step are 14 consecutive
    // 256-byte rows, so a warp's stores of one step are one 3.5 KB span addressed as base + immediate
    constexpr size_t fstride = 32;                                   // doubles between two fields of the same step
    const size_t sstride = (size_t)A.nf_store * 32;                  // doubles between two steps
    double* sp = A.states + (size_t)(r >> 5) * (size_t)Nt * sstride + (size_t)(r & 31);
    const bool st_all = A.store_states != 0;
    const bool st_xyt = st_all || A.keep_xyt;
    // Simpson-rule terms
    FrxSimpson S_acc, S_jerk, S_ori, S_len;
    S_acc.part = S_acc.corr = S_jerk.part = S_jerk.corr = S_ori.part = S_ori.corr = S_len.part = S_len.corr = 0.0;
    double a_prev = 0.0, thc_prev = 0.0;
    const double alpha = (2 * dT * dT + 3 * dT * dT) / (6 * (dT + dT));
    const double beta = (dT * dT + 3.0 * dT * dT) / (6 * dT);
    const double eta = (1 * dT * dT * dT) / (6 * dT * (dT + dT));
    const int nA = Nt, nbA = (nA & 1) ? nA : (nA - 1);            // integrands sampled at every step
    const int nD = Nt - 1, nbD = (nD & 1) ? nD : (nD - 1);        // integrands built from np.diff

    for (int i = i0; i < i1; ++i) {
        // ---- straight-line code with selects instead of branches wherever the two sides are cheap, so that the
        // independent fp64 chains (the two slope divisions, the reciprocal square root, the 1/qc refinement, the a_max(v)
        // quotient) are scheduled together; the unchecked reciprocal sequences are bit-identical to IEEE division inside
        // their range, a divisor outside it is detected and redone with IEEE division.
        const double si = mt[M_S * MP + i], sdi = mt[M_SD * MP + i], sddi = mt[M_SDD * MP + i];
        const double interp = mt[M_INTERP * MP + i];
        const double k_r = mt[M_KR * MP + i], k_r_d = mt[M_KRD * MP + i];
        double di, ddi, dddi;
        {
            const double q1 = si - s_first, q2 = q1 * q1, q3 = q2 * q1, q4 = q2 * q2, q5 = q4 * q1;
            const double u1 = low ? q1 : s_tp[T_1 * TP + i], u2 = low ? q2 : s_tp[T_2 * TP + i], u3 = low ? q3 : s_tp[T_3 * TP + i],
                         u4 = low ? q4 : s_tp[T_4 * TP + i], u5 = low ? q5 : s_tp[T_5 * TP + i];
            const bool inpoly = i < traj_len;
            const double pd = poly_pos(Q, u1, u2, u3, u4, u5), pv = poly_vel(Q, u1, u2, u3, u4), pa = poly_acc(Q, u1, u2, u3);
            di = inpoly ? pd : d_last; ddi = inpoly ? pv : 0.0; dddi = inpoly ? pa : 0.0;
        }
        const bool mov = sdi > 0.001;
        const bool direct = mov || low;
        const bool deferred = (SEG > 1) && (i == i0) && (seg > 0);
        const bool has_prev = (i > 0) && !deferred;
        double dp, dpp;
        {
            const double qa = ddivf(ddi, sdi);
            const double dph = mov ? qa : 0.;
            const double ddot = dddi - dph * sddi;
            const double qb = ddivf(ddot, sdi * sdi);
            dp = low ? ddi : dph;
            dpp = low ? dddi : (mov ? qb : 0.);
        }
        // :423-454 orientations.  Which steps stand still depends on s(t) only, so `direct` is uniform over the lanes of a
        // memo slot: the branch below costs (almost) no divergence, and a stand-still step runs the trigonometric version
        // INSTEAD of the algebraic one (tiles of stopping candidates were the kernel's 20 % slower tail when both ran).
        const double oneKrD = 1 - k_r * di;
        double th_cl, th_gl, cosT, secT, tanT;
        if (direct || !evaluate) {
            th_cl = datan(dp);                        // np.arctan2(dp, 1.0)
            th_gl = th_cl + interp;
            // cos, tan and 1/cos of theta_cl = atan(dp): with w = 1 + dp^2, cos = 1/sqrt(w), 1/cos = sqrt(w), tan = dp hold
            // algebraically (same <= 1-2 ulp error class as libm's cos/tan of the rounded angle)
            const double w = 1.0 + dp * dp;
            cosT = drsqrt_ge1(w); secT = w * cosT; tanT = dp;
        } else {
            // stand-still in high-velocity mode keeps the previous global orientation and needs real trigonometry
            th_gl = th_prev; th_cl = th_gl - interp;
            double sT;
            sincos(th_cl, &sT, &cosT);
            secT = ddivg(1.0, cosT);
            tanT = sT * secT;
        }
        // :457-478
        double qc = oneKrD * secT;                    // oneKrD / cos(theta_cl)
        double cq = drcp_unchecked(qc);               // cos(theta_cl) / oneKrD
        double kappa = (dpp + (k_r * dp + k_r_d * di) * tanT) * cosT * (cq * cq) + cq * k_r;
        double vi = sdi * qc;
        double ai = sddi * qc + ((sdi * sdi) * secT) * (oneKrD * tanT * (kappa * qc - k_r) - (k_r_d * di + k_r * dp));
        const double a_q = ddivf(A.a_max * A.v_switch, vi);
        double a_hi = (vi > A.v_switch) ? a_q : A.a_max;
        if (evaluate && (!drcp_in_range(qc) || ((vi > A.v_switch) && !drcp_in_range(vi)))) {
            // a divisor outside the fast reciprocal's range: IEEE division
            cq = ddivg(1.0, qc);
            kappa = (dpp + (k_r * dp + k_r_d * di) * tanT) * cosT * (cq * cq) + cq * k_r;
            ai = sddi * qc + ((sdi * sdi) * secT) * (oneKrD * tanT * (kappa * qc - k_r) - (k_r_d * di + k_r * dp));
            a_hi = (vi > A.v_switch) ? ddivg(A.a_max * A.v_switch, vi) : A.a_max;
        }
        // :483-533 gates
        uint32_t g = 0;
        {
            const double yaw_rate = has_prev ? ddivc(th_gl - th_prev, dT, A.inv_dt) : 0.;
            const double theta_dot_max = A.kappa_max * vi;
            const double kappa_dot = has_prev ? ddivc(kappa - ka_prev, dT, A.inv_dt) : 0.;
            g |= (vi < -FRX_EPS) ? 1u : 0u;
            g |= (fabs(kappa) > A.kappa_max) ? 2u : 0u;
            g |= (fabs(ddivc(rint(yaw_rate * 100000.0), 100000.0, 1e-5)) > theta_dot_max) ? 4u : 0u;
            g |= (fabs(kappa_dot) > frx_kappa_dot_max(A, kappa)) ? 8u : 0u;
            g |= (!(-A.a_max <= ai && ai <= a_hi)) ? 16u : 0u;
        }
        double kd = has_prev ? (kappa - ka_prev) : 0.0;   // np.append([0], np.diff(kappa_gl))
        // :536-547 Cartesian position: zero from the first out-of-domain step on
        const bool inside = i < first_none;
        double xi = inside ? (mt[M_PX * MP + i] - di * mt[M_SN * MP + i]) : 0.0;
        double yi = inside ? (mt[M_PY * MP + i] + di * mt[M_CS * MP + i]) : 0.0;
        if (evaluate) {
            if (deferred) {
                g_first = g; th_first = th_gl; ka_first = kappa; vi_first = vi; a_first = ai; thc_first = th_cl;
            } else if (brk) {
                if (!gate_hit && g) { gate_or = g & (~g + 1u); gate_hit = true; }   // first violating step, its first gate only
            } else {
                gate_or |= g;
            }
            th_prev = th_gl;
            ka_prev = kappa;
        } else {
            xi = 0.0; yi = 0.0; th_gl = 0.0; th_cl = 0.0; vi = 0.0; ai = 0.0; kappa = 0.0; kd = 0.0;
        }
        // running sums of the two default reductions (velocity_offset :120-130, distance_to_reference_path :154-169)
        if (i >= half && i < Nt - 1) { const double dv = vi - A.v_des; vo_sum += A.vo_norm2 ? dv * dv : fabs(dv); }
        dr_sum += fabs(di);
        v_last = vi; dr_last = di;       // what is left after the loop are the values of the segment's last step (Nt - 1 for its owner)
        if (XCOST) {
            S_acc.add(i, nA, nbA, ai * ai, alpha, beta, eta);
            S_len.add(i, nA, nbA, vi, alpha, beta, eta);
            if (i > i0 || (i > 0 && seg == 0)) {
                double qa = ddivc(ai - a_prev, dT, A.inv_dt), qo = ddivc(th_cl - thc_prev, dT, A.inv_dt);
                S_jerk.add(i - 1, nD, nbD, qa * qa, alpha, beta, eta);
                S_ori.add(i - 1, nD, nbD, qo * qo, alpha, beta, eta);
            }
            a_prev = ai; thc_prev = th_cl;
        }
        if (OBS) {
            const float fxi = (float)(xi - A.origin_x), fyi = (float)(yi - A.origin_y);
            bb_x0 = fminf(bb_x0, fxi); bb_x1 = fmaxf(bb_x1, fxi); bb_y0 = fminf(bb_y0, fyi); bb_y1 = fmaxf(bb_y1, fyi);
        }
        // the 14 fields of this step: lanes = 32 consecutive candidates -> one coalesced 256-byte store each
        double* p = sp + (size_t)i * sstride;
        if (st_xyt) {           // x, y, theta are re-read by the obstacle pass: keep them in L2
            __stcg(p, xi); __stcg(p + fstride, yi); __stcg(p + 2 * fstride, th_gl);
        }
        if (st_all) {
            __stcs(p + 3 * fstride, vi); __stcs(p + 4 * fstride, ai); __stcs(p + 5 * fstride, kappa); __stcs(p + 6 * fstride, kd);
            __stcs(p + 7 * fstride, si); __stcs(p + 8 * fstride, di); __stcs(p + 9 * fstride, th_cl); __stcs(p + 10 * fstride, sdi);
            __stcs(p + 11 * fstride, sddi); __stcs(p + 12 * fstride, ddi); __stcs(p + 13 * fstride, dddi);
        }
    }
    if (SEG > 1) {
        // ---- segment boundaries: the previous segment's last step completes this segment's first one
        const double th_in = __shfl_up_sync(pass, th_prev, C), ka_in = __shfl_up_sync(pass, ka_prev, C);
        if (seg > 0 && evaluate && i0 < i1) {
            double yaw_rate = ddivc(th_first - th_in, dT, A.inv_dt);
            if (fabs(ddivc(rint(yaw_rate * 100000.0), 100000.0, 1e-5)) > A.kappa_max * vi_first) g_first |= 4u;
            double kappa_dot = ddivc(ka_first - ka_in, dT, A.inv_dt);
            if (fabs(kappa_dot) > frx_kappa_dot_max(A, ka_first)) g_first |= 8u;
            if (st_all) __stcs(sp + (size_t)FRX_F_KAPPA_DOT * fstride + (size_t)i0 * sstride, ka_first - ka_in);
        }
        if (XCOST) {
            const double a_in = __shfl_up_sync(pass, a_prev, C), thc_in = __shfl_up_sync(pass, thc_prev, C);
            if (seg > 0 && i0 < i1) {
                // !evaluate lanes carry zeros here like in the loop
                double qa = ddivc((evaluate ? a_first : 0.0) - a_in, dT, A.inv_dt);
                double qo = ddivc((evaluate ? thc_first : 0.0) - thc_in, dT, A.inv_dt);
                S_jerk.add(i0 - 1, nD, nbD, qa * qa, alpha, beta, eta);
                S_ori.add(i0 - 1, nD, nbD, qo * qo, alpha, beta, eta);
            }
        }
        // ---- this segment's verdict, then the candidate's: first violating step wins (brk) / union of all gates
        uint32_t loc = brk ? (g_first ? (g_first & (~g_first + 1u)) : gate_or) : (g_first | gate_or);
        uint32_t all = 0;
#pragma unroll
        for (int sgm = 0; sgm < SEG; ++sgm) {
            const uint32_t v = __shfl_sync(pass, loc, cand + sgm * C);
            if (brk) { if (!all) all = v; } else all |= v;
        }
        gate_or = all;
        // ---- running sums
#pragma unroll
        for (int off = C; off < 32; off <<= 1) {
            vo_sum += __shfl_xor_sync(pass, vo_sum, off);
            dr_sum += __shfl_xor_sync(pass, dr_sum, off);
            if (XCOST) {
                S_acc.part += __shfl_xor_sync(pass, S_acc.part, off); S_acc.corr += __shfl_xor_sync(pass, S_acc.corr, off);
                S_jerk.part += __shfl_xor_sync(pass, S_jerk.part, off); S_jerk.corr += __shfl_xor_sync(pass, S_jerk.corr, off);
                S_ori.part += __shfl_xor_sync(pass, S_ori.part, off); S_ori.corr += __shfl_xor_sync(pass, S_ori.corr, off);
                S_len.part += __shfl_xor_sync(pass, S_len.part, off); S_len.corr += __shfl_xor_sync(pass, S_len.corr, off);
            }
        }
        const int owner = cand + ((Nt - 1) / seglen) * C;     // the lane whose segment holds the last step
        v_last = __shfl_sync(pass, v_last, owner);
        dr_last = __shfl_sync(pass, dr_last, owner);
    }
    if (evaluate) {
        if (gate_or) {
            feasible = false;
            if (gate_or & 1u) reasons |= FRX_FLAG_REASON(4);
            if (gate_or & 2u) reasons |= FRX_FLAG_REASON(5);
            if (gate_or & 4u) reasons |= FRX_FLAG_REASON(6);
            if (gate_or & 8u) reasons |= FRX_FLAG_REASON(7);
            if (gate_or & 16u) reasons |= FRX_FLAG_REASON(8);
        }
        stored = feasible || draw;
        in_list = stored;
        if (stored && seen_none) { valid = false; reasons |= FRX_FLAG_REASON(9); }
    }
    const bool costed = draw ? in_list : (in_list && valid && feasible && stored);
    const bool candidate = draw ? (in_list && feasible) : costed;

    // ---------------- second pass over the steps, only for the lanes that need it: prediction cost
    // (get_inv_mahalanobis_dist, collision_probability.py:264-299), distance_to_obstacles (:172-186) and the
    // collision sweep (planner.py:329-378, collision_check.py:110-200); obstacle data is warp-uniform per step
    double pred_sum = 0.0, d2o_sum = 0.0;
    bool collide = false, boundary = false;
    int col_k = 64, bnd_k = 64;          // ego hull index of the first hit (64 = none)
    if (SEG > 1 && (OBS || XCOST)) __syncwarp(pass);     // the x, y, theta planes of the other segments are visible now
    if (OBS || XCOST) {
        const bool need_pred = OBS && costed && (cost_mask & (1u << FRX_COST_PREDICTION)) && A.O > 0;
        const bool need_d2o = XCOST && costed && (cost_mask & (1u << FRX_COST_DISTANCE_TO_OBSTACLES)) && A.n_obs_pos > 0;
        const bool need_col = OBS && candidate && A.check_collisions && (A.O > 0 || A.B > 0);
        // ---- static boxes (road boundary): the warp culls them ONCE per tile against the bounding box of all positions
        // of its candidates, lane b testing box b (fp32, inflated like frx_cull_radius); the boxes that can be reached go
        // into a per-warp list the step loop walks instead of all A.B of them -- a lanelet network has hundreds of
        // boundary segments, a tile of neighbouring candidates comes near a dozen
        int n_near = -1;                                  // -1: no list, walk all boxes
        if (OBS && A.B > 0 && A.B <= FRX_WALL_LIST && wall_list != nullptr && !A.defer_obs && __any_sync(pass, need_col)) {
            const int kx0 = __reduce_min_sync(pass, need_col ? frx_f32_key(bb_x0) : 0x7fffffff);
            const int kx1 = __reduce_max_sync(pass, need_col ? frx_f32_key(bb_x1) : (int)0x80000000);
            const int ky0 = __reduce_min_sync(pass, need_col ? frx_f32_key(bb_y0) : 0x7fffffff);
            const int ky1 = __reduce_max_sync(pass, need_col ? frx_f32_key(bb_y1) : (int)0x80000000);
            const float x0 = frx_key_f32(kx0), x1 = frx_key_f32(kx1), y0 = frx_key_f32(ky0), y1 = frx_key_f32(ky1);
            // an ego hull spans two consecutive positions of the box: it stays within sqrt(2) (rear-axle offset + half
            // diagonal + half the largest step) of them, and a step is no longer than the box diagonal
            const float diag = sqrtf((x1 - x0) * (x1 - x0) + (y1 - y0) * (y1 - y0));
            const float grow = 1.4143f * ((float)A.wb_rear + sqrtf((float)(A.half_len * A.half_len + A.half_wid * A.half_wid)) + 0.5f * diag) *
                               (1.f + 1e-5f) + 1e-6f * (fabsf(x0) + fabsf(x1) + fabsf(y0) + fabsf(y1)) + 1e-3f;
            const float mx = 0.5f * (x0 + x1), my = 0.5f * (y0 + y1), hx = 0.5f * (x1 - x0) + grow, hy = 0.5f * (y1 - y0) + grow;
            const bool finite = (fabsf(x0) + fabsf(x1) + fabsf(y0) + fabsf(y1)) < 3e38f;
            int cnt = 0;
            __syncwarp(pass);
            for (int b0 = 0; b0 < A.B; b0 += 32) {
                bool near = false;
                if (b0 + lane < A.B) {
                    const float4 c = __ldg(A.sobb32 + b0 + lane);
                    near = !finite || ((fabsf(c.x - mx) <= hx + c.z) && (fabsf(c.y - my) <= hy + c.z));
                }
                // the lanes of `pass` are not always 0..31: every lane of the pass tests the boxes of its own index, the
                // boxes of the absent lanes are kept unconditionally
                const unsigned wm = __ballot_sync(pass, near) | (~pass & (b0 + 32 <= A.B ? 0xffffffffu : ((1u << (A.B - b0)) - 1u)));
                if ((wm >> lane) & 1u) wall_list[cnt + __popc(wm & ((1u << lane) - 1u))] = (unsigned short)(b0 + lane);
                // slots of absent lanes are written by the first lane of the pass
                if (lane == __ffs(pass) - 1) {
                    unsigned missing = wm & ~pass;
                    while (missing) {
                        const int l2 = __ffs(missing) - 1;
                        missing &= missing - 1;
                        wall_list[cnt + __popc(wm & ((1u << l2) - 1u))] = (unsigned short)(b0 + l2);
                    }
                }
                cnt += __popc(wm);
            }
            __syncwarp(pass);
            n_near = cnt;
        }
        if ((need_pred || need_d2o || need_col) && !A.defer_obs) {      // deferred: frx_obstacle_kernel does this pass
            double pbx = 0, pby = 0, pux = 0, puy = 0;   // ego box of the previous step
            if (SEG > 1 && need_col && i0 >= 1 && i0 < i1) {      // ... which another lane wrote for a later segment
                const double* qp = sp + (size_t)(i0 - 1) * sstride;
                double sn, cs;
                sincos(__ldcg(qp + 2 * fstride), &sn, &cs);
                pbx = __ldcg(qp) + A.wb_rear * cs; pby = __ldcg(qp + fstride) + A.wb_rear * sn; pux = cs; puy = sn;
            }
            // x, y, theta of the candidate come back from the state tensor (L2), loaded ONE STEP AHEAD of their use
            const double* q = sp + (size_t)i0 * sstride;
            double x_n = 0.0, y_n = 0.0, th_n = 0.0;
            const bool need_th = need_col;
            if (i0 < i1) { x_n = __ldcg(q); y_n = __ldcg(q + fstride); if (need_th) th_n = __ldcg(q + 2 * fstride); }
            for (int i = i0; i < i1; ++i, q += sstride) {
                const double x = x_n, y = y_n, th = th_n;
                if (i + 1 < i1) {
                    x_n = __ldcg(q + sstride); y_n = __ldcg(q + sstride + fstride);
                    if (need_th) th_n = __ldcg(q + sstride + 2 * fstride);
                }
                if (need_pred && i >= 1) {
                    // the obstacles predicted at this step: 48-byte records, warp-uniform 16-byte loads (frx_pred_step)
                    // (the collision-probability flavour of this term always runs in frx_obstacle_kernel<1>)
                    const double xs[1] = {x}, ys[1] = {y};
                    const bool nd[1] = {true};
                    double acc[1] = {pred_sum};
                    frx_pred_step<1>(reinterpret_cast<const double2*>(A.opred + (size_t)(i - 1) * A.O * FRX_PRED_REC), __ldg(A.on_pred + (i - 1)), xs, ys, nd, acc, A.origin_x,
                                     A.origin_y, A.obs, A.Tp, i - 1);
                    pred_sum = acc[0];
                }
                if (need_d2o) {
                    for (int o = 0; o < A.n_obs_pos; ++o) {
                        double ex = x - __ldg(A.obs_pos + 2 * o), ey = y - __ldg(A.obs_pos + 2 * o + 1);
                        double dist = sqrt(ex * ex + ey * ey);
                        d2o_sum += ddivg(1.0, dist * dist);
                    }
                }
                if (need_col) {
                    double sn, cs;
                    sincos(th, &sn, &cs);
                    const double bx = x + A.wb_rear * cs, by = y + A.wb_rear * sn;     // state.py:30-39 rear axle -> centre
                    if (i >= 1 && !(collide && (boundary || A.B == 0))) {
                        const int k = i - 1;                                            // hull of boxes k, k + 1
                        Hull e = obb_sum_hull(pbx, pby, pux, puy, bx, by, cs, sn, A.half_len, A.half_wid);
                        const double er = sqrt(e.ha * e.ha + e.hb * e.hb) * (1.0 + 1e-9);
                        if (k >= 1 && !collide && A.O > 0) {
                            // obstacle hulls of step k - 1 (hull record: cx, cy, r | ux, uy | ha, hb)
                            const int n = __ldg(A.on_hull + (k - 1));
                            const double2* __restrict__ rec = reinterpret_cast<const double2*>(A.ohull + (size_t)(k - 1) * A.O * 8);
                            // conservative broad phase over all hulls of the step, branch-free (bounding circles, 32
                            // hulls per mask word); the exact separating-axis test runs for the few that pass
                            for (int o0 = 0; o0 < n && !collide; o0 += 32) {
                                const int nn = (n - o0 < 32) ? (n - o0) : 32;
                                unsigned near_mask = 0;
#pragma unroll 4
                                for (int o = 0; o < nn; ++o) {
                                    const double2 cc = __ldg(rec + 4 * (o0 + o));
                                    const double hr = __ldg(reinterpret_cast<const double*>(rec + 4 * (o0 + o) + 1));
                                    double rr = er + hr;
                                    double ddx = cc.x - e.cx, ddy = cc.y - e.cy;
                                    near_mask |= (ddx * ddx + ddy * ddy > rr * rr) ? 0u : (1u << o);
                                }
                                while (near_mask) {
                                    const int o = o0 + __ffs(near_mask) - 1;
                                    near_mask &= near_mask - 1;
                                    const double2 cc = __ldg(rec + 4 * o), ru = __ldg(rec + 4 * o + 1), uh = __ldg(rec + 4 * o + 2);
                                    if (obb_overlap(e, cc.x, cc.y, ru.y, uh.x, uh.y, __ldg(reinterpret_cast<const double*>(rec + 4 * o + 3)))) {
                                        collide = true; col_k = k;
                                        break;
                                    }
                                }
                            }
                        }
                        if (!boundary) {
                            const int nb = (n_near >= 0) ? n_near : A.B;
                            for (int q2 = 0; q2 < nb; ++q2) {
                                const int b = (n_near >= 0) ? (int)wall_list[q2] : q2;
                                const double* __restrict__ sb = A.sobb + b * 8;
                                double rr = er + __ldg(sb + 6);
                                double ddx = __ldg(sb) - e.cx, ddy = __ldg(sb + 1) - e.cy;
                                if (ddx * ddx + ddy * ddy > rr * rr) continue;
                                if (obb_overlap(e, __ldg(sb), __ldg(sb + 1), __ldg(sb + 2), __ldg(sb + 3), __ldg(sb + 4), __ldg(sb + 5))) {
                                    boundary = true; bnd_k = k;
                                    break;
                                }
                            }
                        }
                    }
                    pbx = bx; pby = by; pux = cs; puy = sn;
                }
            }
        }
    }

    if (SEG > 1 && (OBS || XCOST)) {
#pragma unroll
        for (int off = C; off < 32; off <<= 1) {
            if (OBS) pred_sum += __shfl_xor_sync(pass, pred_sum, off);
            if (XCOST) d2o_sum += __shfl_xor_sync(pass, d2o_sum, off);
        }
        if (OBS) {
            unsigned cm2 = __ballot_sync(pass, collide), bm2 = __ballot_sync(pass, boundary);
            unsigned mine = 0;
#pragma unroll
            for (int sgm = 0; sgm < SEG; ++sgm) mine |= 1u << (cand + sgm * C);
            collide = (cm2 & mine) != 0;
            boundary = (bm2 & mine) != 0;
            // every segment lane found the first hit of ITS steps: the candidate's first hit is the earliest of them
#pragma unroll
            for (int off = C; off < 32; off <<= 1) {
                col_k = min(col_k, __shfl_xor_sync(pass, col_k, off));
                bnd_k = min(bnd_k, __shfl_xor_sync(pass, bnd_k, off));
            }
        }
    }
    if (dead) {
        if (seg == 0) { A.flags[r] = 0u; A.total[r] = 0.0; A.traj_len[r] = 0; out.t_missing = true; }
        return out;
    }
    if (SEG > 1 && seg != 0) return out;     // one lane per candidate writes the scalars and reports

    // ---------------- costs (cost_function.py:78-91, partial_cost_functions.py): weighted sum in name-sorted order
    double total = 0.0;
    double* cp = A.costs + (size_t)r * A.n_costs;
    if (costed) {
        for (int k = 0; k < A.n_costs; ++k) {
            const int id = A.cost_ids[k];
            double cv = 0.0;
            switch (id) {
                case FRX_COST_LATERAL_JERK: cv = sq_jerk_integral(Q, dT); break;
                case FRX_COST_LONGITUDINAL_JERK: cv = H->jerk_lon; break;
                case FRX_COST_VELOCITY_OFFSET: { double dv = v_last - A.v_des; cv = vo_sum + fabs(dv * dv); } break;   // :120-130
                case FRX_COST_DISTANCE_TO_REFERENCE_PATH: cv = ddivc(dr_sum + fabs(dr_last) * 5, (double)Nt, A.inv_Nt); break;   // :154-169
                case FRX_COST_PREDICTION: cv = pred_sum; break;
                case FRX_COST_DISTANCE_TO_OBSTACLES: cv = d2o_sum; break;
                case FRX_COST_ACCELERATION: cv = dT / 3.0 * S_acc.part + S_acc.corr; break;
                case FRX_COST_JERK: cv = dT / 3.0 * S_jerk.part + S_jerk.corr; break;
                case FRX_COST_ORIENTATION_OFFSET: cv = dT / 3.0 * S_ori.part + S_ori.corr; break;
                case FRX_COST_PATH_LENGTH: cv = dT / 3.0 * S_len.part + S_len.corr; break;
                default: break;
            }
            total += A.w[k] * cv;
            cp[k] = cv;
        }
    } else {
        for (int k = 0; k < A.n_costs; ++k) cp[k] = 0.0;
    }

    // ---------------- per-candidate scalars
    uint32_t fl = reasons;
    if (valid) fl |= FRX_FLAG_VALID;
    if (feasible) fl |= FRX_FLAG_FEASIBLE;
    if (stored) fl |= FRX_FLAG_STORED;
    if (in_list) fl |= FRX_FLAG_IN_LIST;
    if (costed) fl |= FRX_FLAG_COSTED;
    if (candidate) fl |= FRX_FLAG_CANDIDATE;
    if (collide) fl |= FRX_FLAG_COLLIDE | ((uint32_t)(col_k & 63) << FRX_FLAG_COLLIDE_STEP_SHIFT);
    if (boundary) fl |= FRX_FLAG_BOUNDARY | ((uint32_t)(bnd_k & 63) << FRX_FLAG_BOUNDARY_STEP_SHIFT);
    A.total[r] = total;
    A.flags[r] = fl;
    A.traj_len[r] = traj_len;
    // statistics (reactive_planner.py:229-235): one event bit per counter
    unsigned ev = 0;
    if (in_list) ev |= 1u << CNT_IN_LIST;
    if (in_list && valid && feasible) ev |= 1u << CNT_FEASIBLE;
    if (in_list && !(valid && feasible)) ev |= 1u << CNT_INFEASIBLE_IN_LIST;
    if (candidate) ev |= 1u << CNT_CANDIDATES;
    if (candidate && collide) ev |= 1u << CNT_COLLIDE;
    if (candidate && boundary) ev |= 1u << CNT_BOUNDARY;
    ev |= ((fl >> 2) & 0x3ffu) << CNT_REASON1;      // reason bits 1..10 -> slots CNT_REASON1..+9
    out.ev = ev;
    out.total = total;
    out.winner_ok = candidate && !collide && !boundary && !A.defer_obs;    // deferred: the obstacle kernel selects
    return out;
}

// ------------------------------------------------------------------------------------------------------------
// kernel body: prologue (TMA staging of the reference tables), tile loop, per-CTA / last-CTA arg-min
// ------------------------------------------------------------------------------------------------------------
static_assert(FRX_THREADS >= 96, "the last-CTA epilogue assigns roles to threads 0, 32..47 and 64..69");
template <int SEG, bool OBS, bool XCOST>
__device__ __forceinline__ void frx_tile_body(const FrxKernelArgs& A, const int cta_local, unsigned char* smem_raw) {
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int Mpad = A.Mpad, TP = A.tpitch, MP = A.mpitch;
#if FRX_TRACE
    bool traced = false;
#endif
    FRX_STAMP(0);                                                            // kernel entry
    double* s_ref = reinterpret_cast<double*>(smem_raw);                    // [6][Mpad]
    double* s_Ttab = s_ref + 6 * Mpad;                                       // [FRX_MAX_T_VALUES]
    double* s_tp = s_Ttab + FRX_MAX_T_VALUES;                                // [T_ROWS][TP] rounded powers of the step times
    double* s_memo = s_tp + T_ROWS * TP;                                     // [WARPS][SLOTS][M_FIELDS][MP]
    FrxMemoHdr* s_hdr = reinterpret_cast<FrxMemoHdr*>(s_memo + (size_t)FRX_WARPS_PER_CTA * FRX_MEMO_SLOTS * M_FIELDS * MP);
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_hdr + FRX_WARPS_PER_CTA * FRX_MEMO_SLOTS);
    FrxBest* s_best = reinterpret_cast<FrxBest*>(s_bar + 2);                 // [WARPS]
    double* s_rows = reinterpret_cast<double*>(s_best + FRX_WARPS_PER_CTA);  // [WARPS][32 * 13] sampling rows of a tile

    // ---- stage the reference tables with one TMA bulk copy (UBLKCP) guarded by an mbarrier
    const uint32_t ref_bytes = (uint32_t)(6 * Mpad * sizeof(double));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(s_bar)), "r"(ref_bytes)
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                smem_u32(s_ref)),
            "l"(A.ref), "r"(ref_bytes), "r"(smem_u32(s_bar))
            : "memory");
    }
    for (int k = threadIdx.x; k < FRX_MAX_T_VALUES; k += FRX_THREADS)
        s_Ttab[k] = (k < A.nT) ? A.Ttab[k] : __longlong_as_double(0x7ff8000000000000LL);
    __shared__ int s_Tlen[FRX_MAX_T_VALUES];
    for (int k = threadIdx.x; k < FRX_MAX_T_VALUES; k += FRX_THREADS) s_Tlen[k] = (k < A.nT) ? A.Tlen[k] : 0;
    for (int k = threadIdx.x; k < T_ROWS * TP; k += FRX_THREADS) s_tp[k] = __ldg(A.tpow + k);     // visible after the barrier below
    if (lane < FRX_MEMO_SLOTS) s_hdr[wib * FRX_MEMO_SLOTS + lane].valid = 0;
    __shared__ unsigned short s_walls[OBS ? FRX_WARPS_PER_CTA * FRX_WALL_LIST : 1];     // per-warp list of reachable static boxes
    __shared__ unsigned int s_cnt[CNT_REASON1 + 10];     // per-CTA event counters
    __shared__ unsigned long long s_part[FRX_THREADS];   // last CTA: partial counter sums
    if (threadIdx.x < CNT_REASON1 + 10) s_cnt[threadIdx.x] = 0u;

    unsigned cost_mask = 0;
    for (int k = 0; k < A.n_costs; ++k) cost_mask |= 1u << A.cost_ids[k];
    const long long N = A.N;
    constexpr int C = 32 / SEG;
    const long long n_tiles = (N + C - 1) / C;

    // Work distribution: the first tile of a warp is its global index, further tiles of C consecutive rows come
    // from a global ticket counter, requested one tile ahead.  The 13-column rows of a tile are one contiguous span
    // of the sampling matrix: the warp copies it with coalesced 8-byte cp.async into its staging buffer -- the NEXT
    // tile's rows while the current tile is evaluated, so neither HBM nor (for a pinned host matrix the kernel reads
    // in place, zero-copy) PCIe latency is ever waited on.
    const double* __restrict__ samp = A.sampling;
    double* rows = s_rows + wib * (32 * 13);
    const long long total_warps = (long long)A.n_cta * FRX_WARPS_PER_CTA;
    auto stage_rows = [&](long long tile) {
        const long long base = tile * (C * 13), lim = N * 13;
#pragma unroll
        for (int k = 0; k < (C * 13 + 31) / 32; ++k) {
            const int e = k * 32 + lane;
            if (e < C * 13 && base + e < lim)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(rows + e)), "l"(samp + base + e) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    long long cur_tile = (long long)wib * A.n_cta + cta_local;      // warp-major: idle warps (few tiles) spread over all CTAs
    if (samp != nullptr && cur_tile < n_tiles) stage_rows(cur_tile);
    unsigned long long next_tile = 0;
    if (lane == 0) next_tile = atomicAdd(A.counters + CNT_WORK, 1ULL) + (unsigned long long)total_warps;

    {   // wait for the bulk copy (phase 0)
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(smem_u32(s_bar))
                : "memory");
        }
    }
    __syncthreads();

    FRX_STAMP(1);                                                            // reference tables staged
    double* memo = s_memo + (size_t)wib * FRX_MEMO_SLOTS * M_FIELDS * MP;
    FrxMemoHdr* hdr = s_hdr + wib * FRX_MEMO_SLOTS;
    double best_cost = __longlong_as_double(0x7ff0000000000000LL);  // +inf
    long long best_idx = -1;
    unsigned int my_cnt = 0;          // lane k counts event k (CNT_* enum), 32-bit is ample per warp
    unsigned int t_missing = 0;

    for (;;) {
        const long long tile = cur_tile;
        if (tile >= n_tiles) break;
        const long long r = tile * C + (lane & (C - 1));
        const bool in_range = r < N;
        const long long rl = in_range ? r : (N - 1);
        // ---------------- sampling row (sampling_matrix.py:85-121 column order)
        double T, s0, ss0, sss0, ss1, d0, dd0, ddd0, d1, dd1, ddd1;
        if (samp != nullptr) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            const double* row = rows + (int)(rl - tile * C) * 13;
            T = row[1]; s0 = row[2]; ss0 = row[3]; sss0 = row[4]; ss1 = row[5];
            d0 = row[7]; dd0 = row[8]; ddd0 = row[9]; d1 = row[10]; dd1 = row[11]; ddd1 = row[12];
            __syncwarp();
        }
        FRX_STAMP(2);                                                        // rows of the first tile in registers
        cur_tile = (long long)__shfl_sync(FULL, next_tile, 0);
        if (samp != nullptr && cur_tile < n_tiles) stage_rows(cur_tile);       // prefetch the next tile's rows
        if (lane == 0) next_tile = atomicAdd(A.counters + CNT_WORK, 1ULL) + (unsigned long long)total_warps;
        if (samp == nullptr) {
            const long long g = A.row_first + rl;
            const long long per_t = (long long)A.g_nv * A.g_nd;
            const int it = (int)(g / per_t);
            const int rem = (int)(g - (long long)it * per_t);
            const int iv = rem / A.g_nd, id = rem - iv * A.g_nd;
            T = __ldg(A.g_t1 + it); ss1 = __ldg(A.g_v1 + iv); d1 = __ldg(A.g_d1 + id);
            s0 = A.xcl[0]; ss0 = A.xcl[1]; sss0 = A.xcl[2]; d0 = A.xcl[3]; dd0 = A.xcl[4]; ddd0 = A.xcl[5];
            dd1 = 0.0; ddd1 = 0.0;
        }
        // ---------------- passes of (at most) two memo keys each
        unsigned todo = __ballot_sync(FULL, in_range);
        while (todo) {
            const int la = __ffs(todo) - 1;
            const double aT = __shfl_sync(FULL, T, la), as0 = __shfl_sync(FULL, s0, la), ass0 = __shfl_sync(FULL, ss0, la),
                         asss0 = __shfl_sync(FULL, sss0, la), ass1 = __shfl_sync(FULL, ss1, la);
            const bool eqA = (T == aT) && (s0 == as0) && (ss0 == ass0) && (sss0 == asss0) && (ss1 == ass1);
            const unsigned mA = (__ballot_sync(FULL, eqA) & todo) | (1u << la);
            const unsigned rest = todo & ~mA;
            unsigned mB = 0;
            double bT = 0, bs0 = 0, bss0 = 0, bsss0 = 0, bss1 = 0;
            if (rest) {
                const int lb = __ffs(rest) - 1;
                bT = __shfl_sync(FULL, T, lb); bs0 = __shfl_sync(FULL, s0, lb); bss0 = __shfl_sync(FULL, ss0, lb);
                bsss0 = __shfl_sync(FULL, sss0, lb); bss1 = __shfl_sync(FULL, ss1, lb);
                const bool eqB = (T == bT) && (s0 == bs0) && (ss0 == bss0) && (sss0 == bsss0) && (ss1 == bss1);
                mB = (__ballot_sync(FULL, eqB) & rest) | (1u << lb);
            }
            // which slot holds which key (warp-uniform): reuse a slot filled for the same key by an earlier tile
            int sa = -1, sb = -1;
#pragma unroll
            for (int s = 0; s < FRX_MEMO_SLOTS; ++s) {
                const FrxMemoHdr& h = hdr[s];
                if (h.valid) {
                    if (h.key[0] == aT && h.key[1] == as0 && h.key[2] == ass0 && h.key[3] == asss0 && h.key[4] == ass1) sa = s;
                    else if (mB && h.key[0] == bT && h.key[1] == bs0 && h.key[2] == bss0 && h.key[3] == bsss0 && h.key[4] == bss1) sb = s;
                }
            }
            if (sa < 0) {
                sa = (sb == 0) ? 1 : 0;
                frx_memo_fill(A, s_ref, s_Ttab, s_Tlen, s_tp, FRX_TRW, memo + (size_t)sa * M_FIELDS * MP, hdr + sa, aT, as0, ass0, asss0, ass1);
            }
            if (mB && sb < 0) {
                sb = 1 - sa;
                frx_memo_fill(A, s_ref, s_Ttab, s_Tlen, s_tp, nullptr, memo + (size_t)sb * M_FIELDS * MP, hdr + sb, bT, bs0, bss0, bsss0, bss1);
            }
            __syncwarp();
            FRX_STAMP(3);                                                    // memo slots ready
            const unsigned pass = mA | mB;
            FrxLaneOut o;
            o.ev = 0; o.total = 0.0; o.winner_ok = false; o.t_missing = false;
            if ((pass >> lane) & 1u) {
                const int slot = ((mB >> lane) & 1u) ? sb : sa;
                o = frx_candidate<SEG, OBS, XCOST>(A, cost_mask, pass, r, T, d0, dd0, ddd0, d1, dd1, ddd1, memo + (size_t)slot * M_FIELDS * MP,
                                              hdr + slot, s_tp, OBS ? s_walls + wib * FRX_WALL_LIST : nullptr);
            }
            __syncwarp();
            FRX_STAMP(4);                                                    // candidates of the pass evaluated
            // the running arg-min (planner.py:384-392); rows of a lane only grow, so `<` keeps the lowest row
            if (o.winner_ok && o.total < best_cost) { best_cost = o.total; best_idx = r; }
            if (o.t_missing) t_missing++;
            if (__any_sync(FULL, o.ev != 0)) {
#pragma unroll
                for (int e = 0; e < CNT_REASON1 + 10; ++e) {
                    const unsigned m = __ballot_sync(FULL, (o.ev >> e) & 1u);
                    if (lane == e) my_cnt += __popc(m);
                }
            }
            todo &= ~pass;
        }
#if FRX_TRACE
        FRX_STAMP(5);                                                        // first tile done
        traced = true;
#endif
    }

    // ---------------- per-warp, per-CTA reduction of (min cost, lowest row) and the counters
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double oc = __shfl_xor_sync(FULL, best_cost, off);
        const long long oi = __shfl_xor_sync(FULL, best_idx, off);
        if (oi >= 0 && (best_idx < 0 || oc < best_cost || (oc == best_cost && oi < best_idx))) { best_cost = oc; best_idx = oi; }
    }
#if FRX_TRACE
    traced = false;
#endif
    FRX_STAMP(6);                                                            // all tiles of this warp done
    t_missing = __reduce_add_sync(FULL, t_missing);
    if (lane == 0) {
        s_best[wib].cost = best_cost;
        s_best[wib].idx = best_idx;
        if (t_missing) atomicAdd(A.counters + CNT_T_NOT_FOUND, (unsigned long long)t_missing);
    }
    // event counters: summed per CTA in shared memory and written as one row of A.blockcnt (plain stores) -- thousands
    // of same-address global atomics at the end of the kernel would queue up right in front of the done-counter
    if (lane < CNT_REASON1 + 10 && my_cnt) atomicAdd(&s_cnt[lane], my_cnt);
    __syncthreads();   // thread 0's fence below is cumulative over what the barrier made visible to it
    if (threadIdx.x < CNT_REASON1 + 10)
        A.blockcnt[(size_t)cta_local * (CNT_REASON1 + 10) + threadIdx.x] = s_cnt[threadIdx.x];
    __syncthreads();
    if (A.defer_obs) return;     // the obstacle kernel finishes the plan (arg-min, counters, result record)
    __shared__ int s_is_last;
    if (threadIdx.x == 0) {
        FrxBest b = s_best[0];
#pragma unroll
        for (int w = 1; w < FRX_WARPS_PER_CTA; ++w) {
            FrxBest o = s_best[w];
            if (o.idx >= 0 && (b.idx < 0 || o.cost < b.cost || (o.cost == b.cost && o.idx < b.idx))) b = o;
        }
        A.blockbest[cta_local] = b;
        __threadfence();
        unsigned long long done = atomicAdd(A.counters + CNT_DONE, 1ULL);
        s_is_last = (done == (unsigned long long)(A.n_cta - 1));
    }
    __syncthreads();
    // ---------------- the last CTA of this plan reduces the per-CTA winners, publishes the result record to the
    // mapped host struct (no memcpy node) and re-arms the counters for the next launch
    if (s_is_last) {
        __threadfence();
        // one round of loads: this thread's share of the per-CTA winners and of the per-CTA counter rows
        constexpr int NC = CNT_REASON1 + 10;
        constexpr int NPART = FRX_THREADS / NC;
        FrxBest b; b.cost = __longlong_as_double(0x7ff0000000000000LL); b.idx = -1;
        for (int k = threadIdx.x; k < A.n_cta; k += FRX_THREADS) {
            FrxBest o;
            o.cost = __ldcg(&A.blockbest[k].cost);
            o.idx = __ldcg(&A.blockbest[k].idx);
            if (o.idx >= 0 && (b.idx < 0 || o.cost < b.cost || (o.cost == b.cost && o.idx < b.idx))) b = o;
        }
        {   // thread t: counter t % NC, CTAs t / NC, t / NC + NPART, ...
            const int c = threadIdx.x % NC, part = threadIdx.x / NC;
            unsigned long long acc = 0;
            if (part < NPART)
                for (int k = part; k < A.n_cta; k += NPART) acc += __ldcg(A.blockcnt + (size_t)k * NC + c);
            s_part[threadIdx.x] = acc;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            FrxBest o;
            o.cost = __shfl_xor_sync(FULL, b.cost, off);
            o.idx = __shfl_xor_sync(FULL, b.idx, off);
            if (o.idx >= 0 && (b.idx < 0 || o.cost < b.cost || (o.cost == b.cost && o.idx < b.idx))) b = o;
        }
        if (lane == 0) s_best[wib] = b;
        __syncthreads();
        if (threadIdx.x == 0) {
            b = s_best[0];
#pragma unroll
            for (int w = 1; w < FRX_WARPS_PER_CTA; ++w) {
                FrxBest o = s_best[w];
                if (o.idx >= 0 && (b.idx < 0 || o.cost < b.cost || (o.cost == b.cost && o.idx < b.idx))) b = o;
            }
            s_best[0] = b;                          // local row, for the state copy below
            FRX_PUBLISH_WINNER_SCALARS(A, b.idx);
            if (b.idx >= 0) b.idx += A.row_base;   // the winner record carries the GLOBAL row index
            *A.winner = b;
            A.host_res->winner = b;
            frx_publish_exchange(A.xchg, A.xchg_rank, A.xchg_epoch, b.cost, b.idx);
        } else if (threadIdx.x >= 32 && threadIdx.x < 32 + NC) {
            const int c = threadIdx.x - 32;
            unsigned long long tot = 0;
            for (int q = 0; q < NPART; ++q) tot += s_part[q * NC + c];
            A.host_res->counters[c] = tot;
        } else if (threadIdx.x >= 64 && threadIdx.x < 64 + (FRX_NUM_COUNTERS - NC)) {
            // the few global counters (collision counter of the previous plan's second kernel, unknown durations,
            // tickets, done): snapshot + reset in one step
            const int c = NC + (threadIdx.x - 64);
            unsigned long long v = atomicExch(A.counters + c, 0ULL);
            A.host_res->counters[c] = v;
        }
        __syncthreads();
        {   // the selected trajectory's 14 state rows go into the mapped result record as well: the host reads the
            // optimal trajectory without a second round trip
            const long long wi = s_best[0].idx;
            if (wi >= 0 && A.store_states) {
                const int Nt = A.Nt;
                for (int q = threadIdx.x; q < FRX_NUM_FIELDS * Nt; q += FRX_THREADS) {
                    const int f = q / Nt, i = q - f * Nt;
                    A.host_res->winner_states[f][i] = __ldcg(A.states + frx_state_index(wi, Nt, FRX_NUM_FIELDS, f, i));
                }
            }
        }
        // no system-scope fence: the host reads the record after synchronising with the stream, and kernel completion
        // makes every write (mapped host memory included) visible to it
    }
    FRX_STAMP(7);                                                            // kernel exit
}
