// frx_obstacle.cuh -- the obstacle pass as kernels of its own (large plans; included by frx_kernels.cu): prediction cost
// (collision_probability.py:264-299), distance to obstacles (partial_cost_functions.py:172-186), collision sweep
// (planner.py:329-378, collision_check.py:110-200), then the weighted sum, the arg-min and the result record.
//
// Why split from the eval kernel: this pass is pure fp64 arithmetic on warp-uniform obstacle records.  Inside the eval
// kernel it runs at 12 warps per SM (168 registers, 175 KB of shared memory per SM, ~50 KB of L1 left for the records);
// here it runs 16 warps per SM at 128 registers with the records in shared memory.
//
// Work unit = (R = 2 candidates per thread, a CHUNK of the time steps): R * THREADS consecutive rows owned by a block and
// dealt round-robin (large plans), or R * 32 rows owned by one warp and dealt by ticket (small plans) -- see the kernel.
// One thread per candidate; x / y / theta come back from the state planes through a cp.async ring, one step ahead.
// Per step:
//   * prediction cost: frx_pred_step -- 9.25 fp64 instructions per (candidate, obstacle) instead of 20, records read with
//     LDS from the block's staged copy (DESIGN.md section 5 for what was measured on the way);
//   * collision: every lane builds its exact ego hull (obb-sum of boxes k, k + 1); the warp then culls the step's
//     obstacle hulls COOPERATIVELY: the bounding box of the 32 ego hull circles comes from four REDUX min/max on
//     order-preserving integer images of fp32 coordinates, lane o tests obstacle o against it (fp32, conservatively
//     inflated: frx_cull_radius) and a ballot yields the few hulls any lane can touch.  Only those go through the
//     per-lane exact fp64 circle test and the separating-axis test -- the same decisions as testing all of them, at
//     ~1/5 of the instructions (50 obstacles: 400 -> 70 per lane and step).  Static boxes (road boundary) are culled the
//     same way, so a wall costs one lane-test per warp and step instead of one per candidate and step.
//
// Step chunks: neither term couples the steps of a candidate (the prediction cost is a sum, the sweep only wants the
// FIRST hit), so a plan that would leave the GPU with one or two long units per warp -- 200,000 rows are 1.3 units per
// resident warp -- is cut into C chunks of steps: C times more units of 1/C the length, each writes
// its partial sum and its first hits to scratch, and frx_obstacle_finish_kernel adds them up in chunk order (fixed
// order: the result does not depend on scheduling).  A chunk that starts at step i0 > 0 evaluates step i0 - 1 for the
// ego box only (the hull of boxes i0 - 1, i0 needs it).  Plans with >= 8 units per warp run one chunk and finish inline.
#pragma once

#ifndef FRX_OBS_THREADS
#define FRX_OBS_THREADS 256
#endif
#ifndef FRX_OBS_ROWS
#define FRX_OBS_ROWS 2
#endif
#define FRX_OBS_NOHIT 127u
#define FRX_OBS_RING_BYTES(threads) ((threads) / 32 * 2 * FRX_OBS_ROWS * 3 * 32 * 8)      // two slots of R * 3 planes of 32 doubles per warp


struct FrxObsAcc {      // what a thread carries to the end of the kernel
    double best_cost;
    long long best_idx;
    unsigned n_col, n_bnd;
};

// one candidate: fill the cost terms this pass owns, weighted sum in name-sorted order (cost_function.py:78-91), flags,
// running arg-min (planner.py:384-392: lowest row wins ties)
__device__ __forceinline__ void frx_obs_row_finish(const FrxKernelArgs& A, const long long r, const uint32_t fl, const bool need_pred,
                                                   const double pred_sum, const double d2o_sum, const bool collide, const int col_k,
                                                   const bool boundary, const int bnd_k, FrxObsAcc& acc) {
    const bool costed = (fl & FRX_FLAG_COSTED) != 0, candidate = (fl & FRX_FLAG_CANDIDATE) != 0;
    double total = 0.0;
    if (costed) {
        double* cp = A.costs + (size_t)r * A.n_costs;
        for (int k = 0; k < A.n_costs; ++k) {
            const int id = A.cost_ids[k];
            double cv;
            if (id == FRX_COST_PREDICTION) { cv = need_pred ? pred_sum : 0.0; cp[k] = cv; }
            else if (id == FRX_COST_DISTANCE_TO_OBSTACLES) { cv = d2o_sum; cp[k] = cv; }
            else cv = cp[k];
            total += A.w[k] * cv;
        }
        A.total[r] = total;
    }
    if (collide || boundary) {
        uint32_t f2 = fl;
        if (collide) { f2 |= FRX_FLAG_COLLIDE | ((uint32_t)col_k << FRX_FLAG_COLLIDE_STEP_SHIFT); ++acc.n_col; }
        if (boundary) { f2 |= FRX_FLAG_BOUNDARY | ((uint32_t)bnd_k << FRX_FLAG_BOUNDARY_STEP_SHIFT); ++acc.n_bnd; }
        A.flags[r] = f2;
    }
    if (candidate && !collide && !boundary && (total < acc.best_cost || (total == acc.best_cost && r < acc.best_idx))) {
        acc.best_cost = total; acc.best_idx = r;
    }
}

// block reduction of (min cost, lowest row) and the two counters of this pass; the last block finishes the plan: winners
// of all blocks, counter rows of the eval kernel's CTAs (A.n_cta of them) plus this pass's two global counters, result
// record + winner state rows to mapped host memory
template <int THREADS>
__device__ __forceinline__ void frx_obs_block_finish(const FrxKernelArgs& A, FrxObsAcc acc) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    constexpr int NW = THREADS / 32;
    __shared__ FrxBest s_best[NW];
    __shared__ unsigned int s_hits[2];
    __shared__ unsigned long long s_part[THREADS];
    __shared__ int s_is_last;
    if (threadIdx.x < 2) s_hits[threadIdx.x] = 0u;
    __syncthreads();
    double best_cost = acc.best_cost;
    long long best_idx = acc.best_idx;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double oc = __shfl_xor_sync(FULL, best_cost, off);
        const long long oi = __shfl_xor_sync(FULL, best_idx, off);
        if (oi >= 0 && (best_idx < 0 || oc < best_cost || (oc == best_cost && oi < best_idx))) { best_cost = oc; best_idx = oi; }
    }
    const unsigned n_col = __reduce_add_sync(FULL, acc.n_col), n_bnd = __reduce_add_sync(FULL, acc.n_bnd);
    if (lane == 0) {
        s_best[wib].cost = best_cost; s_best[wib].idx = best_idx;
        if (n_col) atomicAdd(&s_hits[0], n_col);
        if (n_bnd) atomicAdd(&s_hits[1], n_bnd);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        FrxBest b = s_best[0];
#pragma unroll
        for (int w = 1; w < NW; ++w) {
            FrxBest o = s_best[w];
            if (o.idx >= 0 && (b.idx < 0 || o.cost < b.cost || (o.cost == b.cost && o.idx < b.idx))) b = o;
        }
        A.blockbest[blockIdx.x] = b;
        if (s_hits[0]) atomicAdd(A.counters + CNT_COLLIDE, (unsigned long long)s_hits[0]);
        if (s_hits[1]) atomicAdd(A.counters + CNT_BOUNDARY, (unsigned long long)s_hits[1]);
        __threadfence();
        unsigned long long done = atomicAdd(A.counters + CNT_DONE, 1ULL);
        s_is_last = (done == (unsigned long long)(gridDim.x - 1));
    }
    __syncthreads();
    if (!s_is_last) return;
    __threadfence();
    constexpr int NC = CNT_REASON1 + 10;
    constexpr int NPART = THREADS / NC;
    FrxBest b; b.cost = __longlong_as_double(0x7ff0000000000000LL); b.idx = -1;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += THREADS) {
        FrxBest o;
        o.cost = __ldcg(&A.blockbest[k].cost);
        o.idx = __ldcg(&A.blockbest[k].idx);
        if (o.idx >= 0 && (b.idx < 0 || o.cost < b.cost || (o.cost == b.cost && o.idx < b.idx))) b = o;
    }
    {
        const int c = threadIdx.x % NC, part = threadIdx.x / NC;
        unsigned long long a2 = 0;
        if (part < NPART)
            for (int k = part; k < A.n_cta; k += NPART) a2 += __ldcg(A.blockcnt + (size_t)k * NC + c);
        s_part[threadIdx.x] = a2;
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
        for (int w = 1; w < NW; ++w) {
            FrxBest o = s_best[w];
            if (o.idx >= 0 && (b.idx < 0 || o.cost < b.cost || (o.cost == b.cost && o.idx < b.idx))) b = o;
        }
        s_best[0] = b;
        FRX_PUBLISH_WINNER_SCALARS(A, b.idx);
        if (b.idx >= 0) b.idx += A.row_base;   // the winner record carries the GLOBAL row index
        *A.winner = b;
        A.host_res->winner = b;
        frx_publish_exchange(A.xchg, A.xchg_rank, A.xchg_epoch, b.cost, b.idx);
    } else if (threadIdx.x >= 32 && threadIdx.x < 32 + NC) {
        const int c = threadIdx.x - 32;
        unsigned long long tot = 0;
        for (int q = 0; q < NPART; ++q) tot += s_part[q * NC + c];
        if (c == CNT_COLLIDE || c == CNT_BOUNDARY) tot += atomicExch(A.counters + c, 0ULL);
        A.host_res->counters[c] = tot;
    } else if (threadIdx.x >= 64 && threadIdx.x < 64 + (FRX_NUM_COUNTERS - NC)) {
        const int c = NC + (threadIdx.x - 64);
        unsigned long long v = atomicExch(A.counters + c, 0ULL);
        A.host_res->counters[c] = v;
    }
    __syncthreads();
    const long long wi = s_best[0].idx;
    const int Nt = A.Nt;
    if (wi >= 0 && A.store_states) {
        for (int q = threadIdx.x; q < FRX_NUM_FIELDS * Nt; q += THREADS) {
            const int f = q / Nt, i = q - f * Nt;
            A.host_res->winner_states[f][i] = __ldcg(A.states + frx_state_index(wi, Nt, FRX_NUM_FIELDS, f, i));
        }
    }
}

// ---- x / y / theta of the next step: 16-byte asynchronous copies (LDGSTS, L2 -> shared memory, no register, no L1
// allocation) into a two-slot ring per warp.  Held in registers across a step -- twelve doubles per thread -- the values
// were spilled by ptxas the moment they were loaded, which makes the "prefetch" wait for DRAM on the spot (a fifth of all
// stall samples, profiles/r02_ncu_obstacle_config5_1250k_stage_report.txt).  Chunk c of a warp's R * 3 planes of 256 bytes:
// row set c / 48, field (c % 48) / 16, candidates 2 * (c % 16) and + 1.
__device__ __forceinline__ void frx_obs_prefetch(double* slot, const double* const (&wbase)[FRX_OBS_ROWS], const size_t step_off,
                                                 const bool with_theta, const int lane) {
#pragma unroll
    for (int c0 = 0; c0 < FRX_OBS_ROWS * 48; c0 += 32) {
        const int c = c0 + lane, u = c / 48, f = (c % 48) >> 4, pos = (c & 15) * 2;
        if (c < FRX_OBS_ROWS * 48 && (f < 2 || with_theta)) {
            const double* wb = wbase[0];                   // (a select, not an indexed read: the array stays in registers)
#pragma unroll
            for (int k = 1; k < FRX_OBS_ROWS; ++k) wb = (u == k) ? wbase[k] : wb;
            const double* g = wb + step_off + f * 32 + pos;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(slot + (u * 3 + f) * 32 + pos);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void frx_obs_prefetch_wait() {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
}

// PMODE: 0 = inverse-Mahalanobis prediction cost (python path), 1 = collision probability (cpp flavour) -- separate
// instances so that the default one keeps its register budget
template <int PMODE, int THREADS, bool TICKET>
__global__ void __launch_bounds__(THREADS, 512 / THREADS)
frx_obstacle_kernel(const __grid_constant__ FrxKernelArgs A) {
    const int lane = threadIdx.x & 31;
    const int Nt = A.Nt;
    const long long N = A.N;
    const size_t Np = (size_t)A.nf_store * 32;             // doubles between two steps of a candidate
    __shared__ int s_npred[64], s_nhull[64];               // records per step (Nt <= 64)
    // The prediction records of the first obs_stage_steps steps live in shared memory for the life of the block (filled once,
    // read by every warp at every step); later steps are read through the L1, whose working set shrinks accordingly.
    // Through the L1 alone the 16 warps of an SM, each at its own step, keep evicting each other's records (hit rate 77 %,
    // a third of all stall samples on the record loads: profiles/r02_ncu_obstacle_config5_1250k_report.txt).
    extern __shared__ double2 s_dyn[];                     // state ring of every warp (frx_obs_prefetch), then the records
    double* const s_ring = reinterpret_cast<double*>(s_dyn);
    const int stage = (PMODE == 0) ? A.obs_stage_steps : 0;
    for (int k = threadIdx.x; k < 64; k += THREADS) {
        s_npred[k] = (A.O > 0 && k < A.Tp) ? A.on_pred[k] : 0;
        s_nhull[k] = (A.O > 0 && k < A.Tp) ? A.on_hull[k] : 0;
    }
    {
        const double2* __restrict__ g = reinterpret_cast<const double2*>(A.opred);
        const int n2 = stage > 0 ? stage * A.O * (FRX_PRED_REC / 2) + FRX_PRED_REC : 0;    // + 2 records: read-ahead of the last group
        for (int k = threadIdx.x; k < n2; k += THREADS) s_dyn[FRX_OBS_RING_BYTES(THREADS) / sizeof(double2) + k] = __ldg(g + k);
    }
    __syncthreads();
    unsigned cost_mask = 0;
    for (int k = 0; k < A.n_costs; ++k) cost_mask |= 1u << A.cost_ids[k];
    const bool pred_on = (cost_mask & (1u << FRX_COST_PREDICTION)) && A.O > 0;
    const bool d2o_on = (cost_mask & (1u << FRX_COST_DISTANCE_TO_OBSTACLES)) && A.n_obs_pos > 0;
    const bool col_on = A.check_collisions && (A.O > 0 || A.B > 0);
    const double ox = A.origin_x, oy = A.origin_y;
    FrxObsAcc acc;
    acc.best_cost = __longlong_as_double(0x7ff0000000000000LL); acc.best_idx = -1; acc.n_col = acc.n_bnd = 0;
    constexpr int R = FRX_OBS_ROWS;
    const int C = A.obs_chunks;                            // step chunks (1: finish inline)
    const int clen = (Nt + C - 1) / C;
    // Two ways of dealing the (row group, step chunk) units:
    //  * TICKET = false, plans with dozens of units per warp: a unit is R * THREADS consecutive rows, owned by the BLOCK and
    //    dealt round-robin.  Everything that describes the unit is block-uniform and lives in uniform registers (loop
    //    bounds, record base addresses: LDS [UR + imm]); measured 7 % faster on the 10^7-row plan than the other way.
    //  * TICKET = true, plans with few units per warp (a 1/8 shard of that plan is eight): a unit is R * 32 rows, owned by
    //    ONE warp; the first one is the warp's global index, further ones come from a ticket counter.  Warps whose
    //    candidates collide early are done with a unit in a fraction of the time of a collision-free one; static dealing
    //    measured 18 % off linear at 8 GPUs, tickets 8 %.
    constexpr int USTRIDE = TICKET ? 32 : THREADS;                  // rows between the R candidates of a thread
    const int tin = TICKET ? lane : (int)threadIdx.x;               // thread's place in the unit
    const unsigned n_units = (unsigned)((N + (long long)R * USTRIDE - 1) / ((long long)R * USTRIDE)) * (unsigned)C;   // < 2^31 (launcher)
    unsigned unit = TICKET ? blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5) : blockIdx.x;
    while (unit < n_units) {
        const long long b0 = (long long)(unit / (unsigned)C) * (R * USTRIDE);
        const int chunk = (int)(unit % (unsigned)C);
        const int i0 = chunk * clen, i1 = (i0 + clen < Nt) ? (i0 + clen) : Nt;
        long long rr_[R];
        uint32_t fl[R];
        bool need_pred[R], need_col[R], need_d2o[R], live[R];
        bool any_pred = false, any_col = false, any_d2o = false;
#pragma unroll
        for (int u = 0; u < R; ++u) {
            const long long r = b0 + (long long)u * USTRIDE + tin;
            live[u] = r < N;
            rr_[u] = live[u] ? r : (N - 1);
            fl[u] = live[u] ? A.flags[rr_[u]] : 0u;
            const bool costed = (fl[u] & FRX_FLAG_COSTED) != 0, candidate = (fl[u] & FRX_FLAG_CANDIDATE) != 0;
            need_pred[u] = costed && pred_on; need_d2o[u] = costed && d2o_on; need_col[u] = candidate && col_on;
            any_pred |= need_pred[u]; any_col |= need_col[u]; any_d2o |= need_d2o[u];
        }
        double pred_sum[R], d2o_sum[R];
        bool collide[R], boundary[R];
        int col_k[R], bnd_k[R];          // ego hull index of the first hit (planner.py:370-372 reads the velocity there)
#pragma unroll
        for (int u = 0; u < R; ++u) { pred_sum[u] = 0.0; d2o_sum[u] = 0.0; collide[u] = false; boundary[u] = false; col_k[u] = bnd_k[u] = 0; }
        const bool w_pred = __any_sync(FULL, any_pred), w_d2o = __any_sync(FULL, any_d2o);
        const bool w_sweep = __any_sync(FULL, any_col);            // some lane runs the collision sweep
        const bool w_col = w_sweep || (w_pred && PMODE == 1);         // theta is needed (sweep, or the probability cost)
        if ((w_pred || w_d2o || w_col) && i0 < i1) {
            double pbx[R], pby[R], pux[R], puy[R];   // ego box of the previous step
            const double* wbase[R];                  // x plane of the warp's 32 candidates at step 0 (warp-uniform)
            // a later chunk starts one step early: that step only yields the ego box the first hull needs
            const int ifirst = (i0 > 0 && w_sweep) ? (i0 - 1) : i0;
#pragma unroll
            for (int u = 0; u < R; ++u) {
                pbx[u] = pby[u] = pux[u] = puy[u] = 0.0;
                long long r0 = b0 + (long long)u * USTRIDE + (tin & ~31);
                if (r0 >= N) r0 = (N - 1) & ~31LL;                              // a warp past the end re-reads the last block
                wbase[u] = A.states + frx_state_index(r0, Nt, A.nf_store, 0, 0);
            }
            double* ring = s_ring + (threadIdx.x >> 5) * (2 * R * 3 * 32);
            __syncwarp();                                                       // the previous unit's last reads of the ring
            frx_obs_prefetch(ring, wbase, (size_t)ifirst * Np, w_col, lane);
            for (int i = ifirst; i < i1; ++i) {
                double x[R], y[R], th[R];
                double* cur = ring + ((i - ifirst) & 1) * (R * 3 * 32);
                frx_obs_prefetch_wait();
#pragma unroll
                for (int u = 0; u < R; ++u) {
                    x[u] = cur[(u * 3 + 0) * 32 + lane]; y[u] = cur[(u * 3 + 1) * 32 + lane];
                    th[u] = w_col ? cur[(u * 3 + 2) * 32 + lane] : 0.0;
                }
                if (i + 1 < i1) frx_obs_prefetch(ring + ((i + 1 - ifirst) & 1) * (R * 3 * 32), wbase, (size_t)(i + 1) * Np, w_col, lane);
                const bool warm = i < i0;                   // box-only step in front of a later chunk
                if (w_pred && i >= 1 && !warm) {
                    if (PMODE == 0 && i - 1 < stage)
                        frx_pred_step<R, FRX_REC_SHARED>(s_dyn + FRX_OBS_RING_BYTES(THREADS) / sizeof(double2) + (size_t)(i - 1) * A.O * (FRX_PRED_REC / 2),
                                                         s_npred[i - 1], x, y, need_pred, pred_sum, ox, oy, A.obs, A.Tp, i - 1);
                    else if (PMODE == 0)
                        frx_pred_step<R, FRX_REC_GLOBAL>(reinterpret_cast<const double2*>(A.opred + (size_t)(i - 1) * A.O * FRX_PRED_REC),
                                                         s_npred[i - 1], x, y, need_pred, pred_sum, ox, oy, A.obs, A.Tp, i - 1);
                    else
                        frx_prob_step<R>(A.oprob + (size_t)(i - 1) * A.O * FRX_PROB_REC, s_npred[i - 1], x, y, th, need_pred, pred_sum,
                                         2 * A.half_len, 2 * A.half_wid);
                }
                if (w_d2o && !warm) {
#pragma unroll
                    for (int u = 0; u < R; ++u) {
                        if (need_d2o[u]) {
                            for (int o = 0; o < A.n_obs_pos; ++o) {
                                double ex = x[u] - __ldg(A.obs_pos + 2 * o), ey = y[u] - __ldg(A.obs_pos + 2 * o + 1);
                                double dist = sqrt(ex * ex + ey * ey);
                                d2o_sum[u] += ddivg(1.0, dist * dist);
                            }
                        }
                    }
                }
                if (!w_sweep) continue;
#pragma unroll
                for (int u = 0; u < R; ++u) {
                    // lanes that still have something to find; the set only shrinks, so a warp without one is done with
                    // the sweep of this unit for good
                    const bool act = need_col[u] && !(collide[u] && (boundary[u] || A.B == 0));
                    if (!__any_sync(FULL, act)) continue;
                    double sn, cs;
                    sincos(th[u], &sn, &cs);
                    const double bx = x[u] + A.wb_rear * cs, by = y[u] + A.wb_rear * sn;     // state.py:30-39 rear axle -> centre
                    if (i >= 1 && !warm) {
                        const int k = i - 1;                                            // hull of boxes k, k + 1
                        Hull e = obb_sum_hull(pbx[u], pby[u], pux[u], puy[u], bx, by, cs, sn, A.half_len, A.half_wid);
                        const double er = sqrt(e.ha * e.ha + e.hb * e.hb) * (1.0 + 1e-9);
                        // ---- warp bounding box of the active lanes' hull circles (fp32, inflated; frx_cull_radius)
                        const double rx = e.cx - ox, ry = e.cy - oy;
                        const float fx = (float)rx, fy = (float)ry, fr = frx_cull_radius(er, rx, ry);
                        const int kx0 = __reduce_min_sync(FULL, act ? frx_f32_key(fx - fr) : 0x7fffffff);
                        const int kx1 = __reduce_max_sync(FULL, act ? frx_f32_key(fx + fr) : (int)0x80000000);
                        const int ky0 = __reduce_min_sync(FULL, act ? frx_f32_key(fy - fr) : 0x7fffffff);
                        const int ky1 = __reduce_max_sync(FULL, act ? frx_f32_key(fy + fr) : (int)0x80000000);
                        // one more ulp-scale pad for the roundings of fx -+ fr and of the centre / half-extent below
                        const float bx0 = frx_key_f32(kx0), bx1 = frx_key_f32(kx1), by0 = frx_key_f32(ky0), by1 = frx_key_f32(ky1);
                        float mx = 0.5f * (bx0 + bx1), my = 0.5f * (by0 + by1);
                        const float pad = 1e-6f * (fabsf(bx0) + fabsf(bx1) + fabsf(by0) + fabsf(by1)) + 1e-4f;
                        float hx = 0.5f * (bx1 - bx0) + pad, hy = 0.5f * (by1 - by0) + pad;
                        // a non-finite hull (cannot come out of finite inputs) must not hide anything from the exact test
                        if (__any_sync(FULL, act && !(fabsf(fx) + fabsf(fy) + fr < 3e38f))) {
                            mx = my = 0.f; hx = hy = __int_as_float(0x7f800000);
                        }
                        if (k >= 1 && __any_sync(FULL, act && !collide[u])) {
                            // obstacle hulls of step k - 1 (hull record: cx, cy, r | ux, uy | ha, hb)
                            const int n = s_nhull[k - 1];
                            const float4* __restrict__ c32 = A.ohull32 + (size_t)(k - 1) * A.O;
                            const double2* __restrict__ rec = reinterpret_cast<const double2*>(A.ohull + (size_t)(k - 1) * A.O * 8);
                            for (int o0 = 0; o0 < n; o0 += 32) {
                                bool near = false;
                                if (o0 + lane < n) {
                                    const float4 c = __ldg(c32 + o0 + lane);
                                    near = (fabsf(c.x - mx) <= hx + c.z) && (fabsf(c.y - my) <= hy + c.z);
                                }
                                unsigned wm = __ballot_sync(FULL, near);
                                while (wm) {                                  // warp-uniform: the hulls some lane may touch
                                    const int o = o0 + __ffs(wm) - 1;
                                    wm &= wm - 1;
                                    if (act && !collide[u]) {
                                        const double2 cc = __ldg(rec + 4 * o);
                                        const double hr = __ldg(reinterpret_cast<const double*>(rec + 4 * o + 1));
                                        const double rr = er + hr;
                                        const double ddx = cc.x - e.cx, ddy = cc.y - e.cy;
                                        if (!(ddx * ddx + ddy * ddy > rr * rr)) {
                                            const double2 ru = __ldg(rec + 4 * o + 1), uh = __ldg(rec + 4 * o + 2);
                                            if (obb_overlap(e, cc.x, cc.y, ru.y, uh.x, uh.y, __ldg(reinterpret_cast<const double*>(rec + 4 * o + 3)))) {
                                                collide[u] = true; col_k[u] = k;
                                            }
                                        }
                                    }
                                }
                            }
                        }
                        if (A.B > 0 && __any_sync(FULL, act && !boundary[u])) {
                            for (int b0s = 0; b0s < A.B; b0s += 32) {
                                bool near = false;
                                if (b0s + lane < A.B) {
                                    const float4 c = __ldg(A.sobb32 + b0s + lane);
                                    near = (fabsf(c.x - mx) <= hx + c.z) && (fabsf(c.y - my) <= hy + c.z);
                                }
                                unsigned wm = __ballot_sync(FULL, near);
                                while (wm) {
                                    const int b = b0s + __ffs(wm) - 1;
                                    wm &= wm - 1;
                                    if (act && !boundary[u]) {
                                        const double* __restrict__ sb = A.sobb + b * 8;
                                        const double rr = er + __ldg(sb + 6);
                                        const double ddx = __ldg(sb) - e.cx, ddy = __ldg(sb + 1) - e.cy;
                                        if (!(ddx * ddx + ddy * ddy > rr * rr) &&
                                            obb_overlap(e, __ldg(sb), __ldg(sb + 1), __ldg(sb + 2), __ldg(sb + 3), __ldg(sb + 4), __ldg(sb + 5))) {
                                            boundary[u] = true; bnd_k[u] = k;
                                        }
                                    }
                                }
                            }
                        }
                    }
                    pbx[u] = bx; pby[u] = by; pux[u] = cs; puy[u] = sn;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < R; ++u) {
            if (!live[u]) continue;
            if (C == 1) {
                frx_obs_row_finish(A, rr_[u], fl[u], need_pred[u], pred_sum[u], d2o_sum[u], collide[u], col_k[u], boundary[u],
                                   bnd_k[u], acc);
            } else {
                // partial results of this chunk: frx_obstacle_finish_kernel combines them in chunk order
                const size_t slot = (size_t)chunk * (size_t)N + (size_t)rr_[u];
                A.obs_part[slot] = pred_sum[u];
                A.obs_hit[slot] = (collide[u] ? (uint32_t)col_k[u] : FRX_OBS_NOHIT) | ((boundary[u] ? (uint32_t)bnd_k[u] : FRX_OBS_NOHIT) << 8);
            }
        }
        // next ticket (a unit is hundreds of microseconds of work: nothing to gain from asking ahead, and nothing live across it)
        if (!TICKET) { unit += gridDim.x; continue; }
        unsigned next = 0;
        if (lane == 0) next = (unsigned)atomicAdd(A.counters + CNT_OBS_WORK, 1ULL) + gridDim.x * (THREADS / 32);
        unit = __shfl_sync(FULL, next, 0);
    }
    if (C == 1) frx_obs_block_finish<THREADS>(A, acc);
}

// chunked plans: add the partial sums up in chunk order, earliest hit over the chunks, then the same finish as the
// one-chunk kernel (grid-stride over the rows, one thread per candidate)
__global__ void __launch_bounds__(FRX_OBS_THREADS)
frx_obstacle_finish_kernel(const __grid_constant__ FrxKernelArgs A) {
    const long long N = A.N;
    const int C = A.obs_chunks;
    unsigned cost_mask = 0;
    for (int k = 0; k < A.n_costs; ++k) cost_mask |= 1u << A.cost_ids[k];
    const bool pred_on = (cost_mask & (1u << FRX_COST_PREDICTION)) && A.O > 0;
    FrxObsAcc acc;
    acc.best_cost = __longlong_as_double(0x7ff0000000000000LL); acc.best_idx = -1; acc.n_col = acc.n_bnd = 0;
    for (long long r = (long long)blockIdx.x * FRX_OBS_THREADS + threadIdx.x; r < N; r += (long long)gridDim.x * FRX_OBS_THREADS) {
        const uint32_t fl = A.flags[r];
        double pred = 0.0;
        uint32_t ck = FRX_OBS_NOHIT, bk = FRX_OBS_NOHIT;
        for (int c = 0; c < C; ++c) {
            const size_t slot = (size_t)c * (size_t)N + (size_t)r;
            pred += __ldcg(A.obs_part + slot);
            const uint32_t h = __ldcg(A.obs_hit + slot);
            ck = min(ck, h & 0xffu); bk = min(bk, (h >> 8) & 0xffu);
        }
        frx_obs_row_finish(A, r, fl, (fl & FRX_FLAG_COSTED) && pred_on, pred, 0.0, ck != FRX_OBS_NOHIT, (int)ck, bk != FRX_OBS_NOHIT,
                           (int)bk, acc);
    }
    frx_obs_block_finish<FRX_OBS_THREADS>(A, acc);
}

// How many step chunks: enough units for >= 4 per resident warp, at least 4 steps per chunk; FRX_OBS_CHUNKS overrides
// (tests pin 1 to compare with the fused pass bit for bit).  The distance-to-obstacles term keeps one chunk.
static int frx_obstacle_chunks(const FrxKernelArgs& a, long long warps_resident) {
    if (a.obs_part == nullptr || a.obs_hit == nullptr) return 1;     // no scratch reserved (very large plans)
    if (const char* e = getenv("FRX_OBS_CHUNKS")) {
        const int f = atoi(e);
        if (f >= 1 && f <= 16) return (f < a.Nt) ? f : 1;
    }
    for (int k = 0; k < a.n_costs; ++k)
        if (a.cost_ids[k] == FRX_COST_DISTANCE_TO_OBSTACLES && a.n_obs_pos > 0) return 1;
    const long long units = (a.N + 32 * FRX_OBS_ROWS - 1) / (32 * FRX_OBS_ROWS);
    int c = 1;
    while (c < 8 && units * c < 4 * warps_resident && (a.Nt + 2 * c - 1) / (2 * c) >= 4) c *= 2;
    return c;
}

// scratch the chunked pass needs (elements of obs_part / obs_hit)
size_t frx_obstacle_scratch_elems(long long N) { return (size_t)8 * (size_t)N; }

// Two block shapes, 16 warps per SM at 128 registers either way: two blocks of 256 threads when the prediction records of
// the whole plan fit twice into the SM's shared memory next to the state rings (84 KB each), else one block of 512 threads
// with up to 160 KB of records (measured, configs[4]: 50 obstacles x 50 steps = 120 KB, 21.2 ms against 23.6 ms with 35 of
// the 50 steps staged per 256-thread block; configs[2]: 29 KB, 0.238 ms against 0.252 ms the other way round).  What does
// not fit is read through what is left of the L1.
#define FRX_OBS_STAGE_BYTES_2 (84 * 1024)
#define FRX_OBS_STAGE_BYTES_1 (160 * 1024)

template <int PMODE, int THREADS, bool TICKET>
static cudaError_t frx_launch_obstacle_shape(FrxKernelArgs& a, int sm_count, size_t stage_bytes, cudaStream_t st, int* launches) {
    // function attributes are sticky per function AND device: the cache is keyed by the device this launch goes to
    struct Entry { int dev, occ, carve; };
    static thread_local Entry cache[16];
    static thread_local int n_cache = 0;
    int dev = -1;
    cudaGetDevice(&dev);
    Entry* e = nullptr;
    for (int k = 0; k < n_cache; ++k)
        if (cache[k].dev == dev) { e = &cache[k]; break; }
    if (e == nullptr) {
        if (n_cache == 16) n_cache = 0;                     // more devices than slots: start over (attributes are set again)
        e = &cache[n_cache++];
        e->dev = dev; e->occ = 0; e->carve = -1;
    }
    constexpr size_t RING = FRX_OBS_RING_BYTES(THREADS);
    constexpr size_t MAXDYN = RING + (PMODE == 0 ? (THREADS == 256 ? FRX_OBS_STAGE_BYTES_2 : FRX_OBS_STAGE_BYTES_1) : 0);
    if (e->occ == 0) {
        cudaError_t rc = cudaFuncSetAttribute(frx_obstacle_kernel<PMODE, THREADS, TICKET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAXDYN);
        if (rc != cudaSuccess) return rc;
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, frx_obstacle_kernel<PMODE, THREADS, TICKET>, THREADS, MAXDYN) != cudaSuccess || occ < 1) occ = 1;
        e->occ = occ;
        if (getenv("FRX_DEBUG")) fprintf(stderr, "[frx] obstacle kernel<%d>: %d blocks of %d threads per SM\n", PMODE, occ, THREADS);
    }
    const int occ = e->occ;
    const size_t dyn = RING + stage_bytes;
    const int want_carve = (int)(((512 / THREADS) * (dyn + 5 * 1024)) * 100 / (228 * 1024)) + 1;       // the rest stays L1
    if (want_carve != e->carve) {
        cudaFuncSetAttribute(frx_obstacle_kernel<PMODE, THREADS, TICKET>, cudaFuncAttributePreferredSharedMemoryCarveout, want_carve > 100 ? 100 : want_carve);
        e->carve = want_carve;
    }
    const long long full = (long long)sm_count * occ;
    a.obs_chunks = frx_obstacle_chunks(a, full * (THREADS / 32));
    long long want;                                                                              // blocks that have a first unit
    if (TICKET) want = ((a.N + 32 * FRX_OBS_ROWS - 1) / (32 * FRX_OBS_ROWS) * a.obs_chunks + THREADS / 32 - 1) / (THREADS / 32);
    else want = (a.N + THREADS * FRX_OBS_ROWS - 1) / (THREADS * FRX_OBS_ROWS) * a.obs_chunks;
    const int grid = (int)(want < full ? want : full);
    frx_obstacle_kernel<PMODE, THREADS, TICKET><<<grid, THREADS, dyn, st>>>(a);
    *launches = 1;
    if (a.obs_chunks > 1) {
        long long fg = (a.N + FRX_OBS_THREADS - 1) / FRX_OBS_THREADS;
        if (fg > (long long)sm_count * 8) fg = (long long)sm_count * 8;
        frx_obstacle_finish_kernel<<<(int)fg, FRX_OBS_THREADS, 0, st>>>(a);
        *launches = 2;
    }
    return cudaGetLastError();
}

cudaError_t frx_launch_obstacle_pass(FrxKernelArgs& a, int sm_count, cudaStream_t st, int* launches) {
    a.obs_stage_steps = 0;
    // per-warp tickets below 32 units per resident warp (16 warps per SM in either shape)
    bool ticket = (a.N + 32 * FRX_OBS_ROWS - 1) / (32 * FRX_OBS_ROWS) < 32LL * 16 * sm_count;
    if (const char* e = getenv("FRX_OBS_TICKET")) ticket = atoi(e) != 0;     // tuning: force one way of dealing
    if (a.pred_mode == 1) return frx_launch_obstacle_shape<1, 256, true>(a, sm_count, 0, st, launches);
    // prediction records staged in shared memory: as many leading steps as fit
    bool wide = false;
    size_t per_step = 0;
    if (a.O > 0 && a.opred != nullptr) {
        per_step = (size_t)a.O * FRX_PRED_REC * sizeof(double);
        const int used = a.Tp < a.Nt - 1 ? a.Tp : a.Nt - 1;                    // record lists the steps 1 .. Nt - 1 read
        wide = (size_t)used * per_step > FRX_OBS_STAGE_BYTES_2;
        if (const char* e = getenv("FRX_OBS_WIDE")) wide = atoi(e) != 0;       // tuning: force a shape
        size_t budget = wide ? FRX_OBS_STAGE_BYTES_1 : FRX_OBS_STAGE_BYTES_2;
        if (const char* e = getenv("FRX_OBS_STAGE_KB")) {                      // tuning: 0 = records through the L1 only
            const size_t b = (size_t)atoll(e) * 1024;
            if (b < budget) budget = b;
        }
        const long long steps = budget > 96 ? (long long)((budget - 96) / per_step) : 0;
        a.obs_stage_steps = (int)(steps < used ? steps : used);
    }
    const size_t stage_bytes = a.obs_stage_steps > 0 ? (size_t)a.obs_stage_steps * per_step + 2 * FRX_PRED_REC * sizeof(double) : 0;
    if (wide) return ticket ? frx_launch_obstacle_shape<0, 512, true>(a, sm_count, stage_bytes, st, launches)
                            : frx_launch_obstacle_shape<0, 512, false>(a, sm_count, stage_bytes, st, launches);
    return ticket ? frx_launch_obstacle_shape<0, 256, true>(a, sm_count, stage_bytes, st, launches)
                  : frx_launch_obstacle_shape<0, 256, false>(a, sm_count, stage_bytes, st, launches);
}
int frx_obstacle_pass_max_grid(int sm_count) { return sm_count * 8; }

