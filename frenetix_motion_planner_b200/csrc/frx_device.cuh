// frx_device.cuh -- device-side data structures shared by the kernels and the C-ABI host code.
#pragma once
#include <stdint.h>
#include "frx.h"

#ifndef FRX_WARPS_PER_CTA
#define FRX_WARPS_PER_CTA 6   // measured on B200: 6 warps x 2 CTAs/SM beats 4 x 3 and 8 x 2 (shared memory per SM stays
#endif                        // under the 164 KB carve-out, which leaves ~90 KB of L1 for the obstacle records)
#define FRX_THREADS (FRX_WARPS_PER_CTA * 32)
#ifndef FRX_MIN_CTAS
#define FRX_MIN_CTAS 2   // resident CTAs per SM the eval kernel is compiled for (register cap 168)
#endif
#define FRX_MAX_T_VALUES 128
#define FRX_EPS 1e-5

// obstacle table, SoA with step pitch Tp: arr[(o * FRX_OBS_NARR + k) * Tp + t]
enum {
    OB_PX = 0, OB_PY, OB_IV00, OB_IV01, OB_IV10, OB_IV11,           // mean + inverse covariance at step t
    OB_HCX, OB_HCY, OB_HUX, OB_HUY, OB_HHA, OB_HHB, OB_HR,          // obb-sum hull of boxes t, t+1
    FRX_OBS_NARR
};

// counters written by the eval kernel (unsigned long long each)
enum {
    CNT_IN_LIST = 0, CNT_FEASIBLE, CNT_CANDIDATES, CNT_COLLIDE, CNT_BOUNDARY, CNT_INFEASIBLE_IN_LIST,
    CNT_REASON1,  // .. CNT_REASON1 + 9 = reason 10
    CNT_COLLISION_COUNTER = CNT_REASON1 + 10,
    CNT_T_NOT_FOUND,
    CNT_WORK,          // chunk ticket counter of the eval kernel's dynamic scheduler
    CNT_DONE,          // CTAs of this plan that have finished (last one publishes the result)
    CNT_OBS_WORK,      // unit ticket counter of the obstacle kernel's warps
    FRX_NUM_COUNTERS
};

struct FrxBest {
    double cost;
    long long idx;
};

// result record the eval kernel's last CTA writes into mapped (pinned) host memory
struct FrxHostResult {
    FrxBest winner;
    unsigned long long counters[FRX_NUM_COUNTERS];
    double winner_states[FRX_NUM_FIELDS][64];   // the selected candidate's state rows (Nt <= 64 samples each)
    double winner_costs[FRX_MAX_COSTS];         // ... its unweighted cost terms, flags and sample count: everything a
    unsigned int winner_flags;                  //     TrajectorySample of the winner can be asked for, without a second
    int winner_traj_len;                        //     device round trip
};

// the last CTA copies the selected candidate's scalars into the mapped result record (thread 0 of the finishing block)
#define FRX_PUBLISH_WINNER_SCALARS(A, wi)                                                          \
    do {                                                                                           \
        if ((wi) >= 0) {                                                                           \
            (A).host_res->winner_flags = __ldcg((A).flags + (wi));                                 \
            (A).host_res->winner_traj_len = __ldcg((A).traj_len + (wi));                           \
            for (int k_ = 0; k_ < (A).n_costs; ++k_)                                               \
                (A).host_res->winner_costs[k_] = __ldcg((A).costs + (size_t)(wi) * (A).n_costs + k_); \
        }                                                                                          \
    } while (0)

// Multi-GPU arg-min exchange without a collective kernel: ONE page of pinned host memory shared by the ranks of a node
// (POSIX shm, registered with every rank's CUDA context).  The last CTA of a plan stores its rank's winner record into its
// slot -- a posted PCIe write, the same mechanism as the result record -- and every rank's host reads all slots.  Slots are
// double-buffered by the parity of the plan epoch (a fast rank can be at most one plan ahead of a slow reader).
struct FrxXchgSlot {
    double cost;
    long long idx;                  // global row, -1 = no candidate
    unsigned long long epoch;       // written LAST (after a system-scope fence): the record of plan `epoch` is complete
    unsigned long long pad[5];      // one slot per 64-byte line
};
#define FRX_XCHG_MAX_RANKS 64
#define FRX_XCHG_PAGE_BYTES (2 * FRX_XCHG_MAX_RANKS * sizeof(FrxXchgSlot))

__device__ __forceinline__ void frx_publish_exchange(FrxXchgSlot* page, int rank, unsigned long long epoch, double cost, long long idx) {
    if (page == nullptr) return;
    volatile FrxXchgSlot* s = page + (epoch & 1ULL) * FRX_XCHG_MAX_RANKS + rank;
    s->cost = cost; s->idx = idx;
    __threadfence_system();
    s->epoch = epoch;
}

// state tensor index: blocks of 32 candidates, [block][step][field][32] (nf = fields stored per step: 14, or 3 when only
// x, y, theta are kept for the obstacle pass)
__host__ __device__ inline size_t frx_state_index(long long row, int Nt, int nf, int f, int i) {
    return (((size_t)(row >> 5) * (size_t)Nt + (size_t)i) * (size_t)nf + (size_t)f) * 32 + (size_t)(row & 31);
}

struct FrxKernelArgs {
    // ---- scalars (frx_params, pre-digested on the host)
    double dt, a_max, v_switch, kappa_max, wb_rear, half_len, half_wid, x0_orientation, v_des;
    double inv_dt, inv_Nt;  // correctly rounded 1/dt and 1/Nt (host), for ddivc
    double inv_step;        // (M - 1) / (ref_pos[M-1] - ref_pos[0]): initial guess of the reference-segment search
    double w[FRX_MAX_COSTS];
    int cost_ids[FRX_MAX_COSTS];
    int n_costs;
    int Nt, Ntp;            // samples per candidate, step pitch of the state tensor
    int low, draw, debug;
    int store_states, check_collisions;
    int kd_from_v_delta, vo_norm2;      // cpp-path variants (frx_params.curvature_rate_from_v_delta / velocity_offset_norm == 2)
    double v_delta_over_wb, wheelbase;  // v_delta_max / wheelbase and the wheelbase, for the curvature-rate limit
    // ---- reference path: 6 tables of Mpad doubles, contiguous (pos, theta, curv, curv_d, x, y)
    const double* ref;
    int M, Mpad;
    // ---- time tables
    const double* Ttab;     // [nT] distinct durations
    const int* Tlen;        // [nT] traj_len
    const double* tpow;     // [5][tpitch] rounded powers t, t^2 .. t^5 of the step times (the same for every duration)
    int nT, tpitch;
    int mpitch;             // step pitch of the per-warp memo rows in shared memory (Nt rounded up to 4)
    // ---- predictions / obstacles
    const double* obs;      // [O][FRX_OBS_NARR][Tp]
    const int* obs_len;     // [O] valid steps
    const double* opred;    // [Tp][O][6] per-step compact prediction records (frx_obstacle_compact_kernel): px, py, iv00,
                            //            iv01 + iv10, iv11, 0 -- the quadratic form of the inverse covariance
    const double* ohull;    // [Tp][O][8] per-step compact hull records
    const double* oprob;    // [Tp][O][8] records of the collision-probability cost (pred_mode 1): px, py, devx, devy, sx, sy, rho
    int pred_mode;          // 0: inverse Mahalanobis (python path), 1: collision probability (cpp flavour)
    const float4* ohull32;  // [Tp][O]    fp32 copies (cx - origin, cy - origin, inflated radius, 0) for the warp-level cull
    const int* on_pred;     // [Tp] records per step
    const int* on_hull;     // [Tp]
    int O, Tp;
    const double* obs_pos;  // [n_obs_pos][2] current obstacle positions (distance_to_obstacles)
    int n_obs_pos;
    const double* sobb;     // [B][8]: cx, cy, ux, uy, ha, hb, r, pad
    const float4* sobb32;   // [B] fp32 copies (cx - origin, cy - origin, inflated radius, 0) for the warp-level cull
    int B;
    double origin_x, origin_y;   // frame of the fp32 cull records (first reference-path vertex)
    // ---- candidates
    const double* sampling; // [N][13] or nullptr (grid mode)
    const double* g_t1; const double* g_v1; const double* g_d1;
    int g_nv, g_nd;
    double xcl[6];
    long long row_first;    // global index of local row 0 (grid mode: also the first generated row)
    long long row_base;     // added to the local index when reporting argmin
    long long N;
    // ---- outputs
    double* states;         // [Np / 32][Nt][nf_store][32]: block of 32 candidates, step, field, candidate (frx_state_index)
    long long Np;           // N rounded up to 32
    int nf_store;           // fields stored per step: 14 (store_states) or 3 (x, y, theta only)
    int seg;                // lanes per candidate (1, 2 or 4): which kernel instance runs, tile = 32 / seg rows
    int defer_obs;          // 1: the eval kernel skips the obstacle pass, arg-min and result record; frx_obstacle_kernel
                            //    (launched right behind it) does them
    int keep_xyt;           // store_states == 0 but the obstacle pass needs the x, y, theta planes
    int obs_stage_steps;    // leading steps whose prediction records every block of the obstacle pass keeps in shared memory
    int obs_chunks;         // step chunks of the split obstacle pass (1: frx_obstacle_kernel finishes the plan inline)
    double* obs_part;       // [obs_chunks][N] partial prediction cost of a chunk
    uint32_t* obs_hit;      // [obs_chunks][N] first colliding hull of a chunk: collide | boundary << 8, 127 = none
    double* costs;          // [N][n_costs]
    double* total;          // [N]
    uint32_t* flags;        // [N]
    int* traj_len;          // [N]
    FrxBest* blockbest;     // [n_cta]
    unsigned long long* blockcnt;   // [n_cta][CNT_REASON1 + 10] per-CTA event counters (summed by the last CTA)
    unsigned long long* counters;
    FrxBest* winner;        // device copy of the winner record (multi-GPU exchange payload)
    FrxHostResult* host_res;// device address of the mapped host result struct
    unsigned long long* trace;  // FRX_TRACE tuning builds: [warps][8] globaltimer stamps, else null
    int n_cta;              // CTAs working on this plan (grid size, or this agent's share of a batched grid)
    FrxXchgSlot* xchg;      // device address of the node's shared exchange page, or null
    int xchg_rank;
    unsigned long long xchg_epoch;
};
