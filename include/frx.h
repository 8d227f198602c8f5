/*
 * frx.h -- C ABI of the B200 (sm_100a) reactive-planner hot path.
 *
 * One shared library, libfrx_b200.so, exports exactly the entry points below.  No C++ or torch
 * types cross this boundary: plain pointers, sizes and POD structs only.  Every function returns
 * FRX_OK (0) or a negative error code and never throws; frx_last_error() gives the message.
 *
 * What each entry point replaces in the reference (TUM-AVS/Frenetix-Motion-Planner):
 *
 *   frx_create / frx_destroy      frenetix.TrajectoryHandler(dt) life cycle
 *                                 (frenetix_motion_planner/reactive_planner_cpp.py:49)
 *   frx_set_reference             CoordinateSystem tables ref_pos/ref_theta/ref_curv/ref_curv_d
 *                                 (cr_scenario_handler/utils/utils_coordinate_system.py:203-207) and
 *                                 frenetix.CoordinateSystemWrapper(reference_path)
 *                                 (reactive_planner_cpp.py:192)
 *   frx_set_params                vehicle / planning / debug scalars read by check_feasibility
 *                                 (frenetix_motion_planner/reactive_planner.py:274-577), the
 *                                 add_feasability_function / add_cost_function set-up
 *                                 (reactive_planner_cpp.py:96-141) and the name-sorted weight list of
 *                                 AdaptableCostFunction (cost_functions/cost_function.py:55-60)
 *   frx_set_time_tables           the rounded time-power tables of reactive_planner.py:296-300
 *                                 (computed with numpy by the host layer, one per distinct duration)
 *   frx_set_predictions           Planner.set_predictions / CalculateCollisionProbabilityFast set-up
 *                                 (reactive_planner.py:49-51, reactive_planner_cpp.py:151-155) and
 *                                 the obstacle side of collision_check_prediction
 *                                 (cr_scenario_handler/utils/collision_check.py:110-200)
 *   frx_set_obstacle_positions    CalculateDistanceToObstacleCost(positions) (reactive_planner_cpp.py:166)
 *   frx_set_static_obbs           road-boundary collision objects (frenetix_motion_planner/planner.py:362-368)
 *   frx_plan                      handler.generate_trajectories(sampling_matrix[N x 13], low_vel_mode) +
 *                                 handler.evaluate_all_current_functions(True) +
 *                                 get_sorted_trajectories()[0] after trajectory_collision_check
 *                                 (reactive_planner_cpp.py:256,347-374; reactive_planner.py:89-94,184-272)
 *   frx_plan_grid                 same, rows generated on the device from the three 1-D ranges that
 *                                 generate_sampling_matrix would expand
 *                                 (frenetix_motion_planner/sampling_matrix.py:85-121)
 *   frx_plan_device               same as frx_plan with the sampling matrix already resident in HBM
 *   frx_plan_batched              AgentBatch._step_agents: one launch for all agents' plan() calls
 *                                 (cr_scenario_handler/simulation/agent_batch.py:186-189, agent.py:185-270)
 *   frx_get_*                     lazy read-back of what the reference keeps in TrajectorySample objects
 *                                 (frenetix_motion_planner/trajectories.py:56-478)
 *
 * Ownership: the caller owns every host buffer (row-major, float64 unless noted, may be pinned);
 * the library owns all device buffers, valid until the next frx_plan* / frx_destroy on that ctx.
 * Threading: one ctx = one CUDA stream; a ctx is NOT thread-safe; use one ctx per planner instance.
 */
#ifndef FRX_H
#define FRX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRX_ABI_VERSION 2

/* error codes */
#define FRX_OK 0
#define FRX_ERR_INVALID (-1)   /* bad argument / missing set-up call */
#define FRX_ERR_CUDA (-2)      /* CUDA runtime error, see frx_last_error */
#define FRX_ERR_NOMEM (-3)
#define FRX_ERR_UNSUPPORTED (-4)

/* state tensor fields.  Read-back (frx_get_states*, frx_winner_states) delivers out[field][candidate][step] with step
 * pitch frx_state_pitch(); in HBM the tensor is laid out in blocks of 32 candidates, [block][step][field][32], so that the
 * 32 candidates of a warp store the 14 fields of a step as one contiguous 3.5 KB span (14 coalesced 256-byte rows). */
enum {
    FRX_F_X = 0, FRX_F_Y, FRX_F_THETA, FRX_F_V, FRX_F_A, FRX_F_KAPPA, FRX_F_KAPPA_DOT,
    FRX_F_S, FRX_F_D, FRX_F_THETA_CL, FRX_F_S_DOT, FRX_F_S_DDOT, FRX_F_D_DOT, FRX_F_D_DDOT,
    FRX_NUM_FIELDS
};

/* per-candidate flag bits */
#define FRX_FLAG_VALID (1u << 0)
#define FRX_FLAG_FEASIBLE (1u << 1)
#define FRX_FLAG_REASON(r) (1u << (1 + (r))) /* r = 1..10: slots of _infeasible_count_kinematics */
#define FRX_FLAG_COLLIDE (1u << 12)
#define FRX_FLAG_BOUNDARY (1u << 13)
#define FRX_FLAG_STORED (1u << 14)    /* has Cartesian/curvilinear samples (reactive_planner.py:551-567) */
#define FRX_FLAG_IN_LIST (1u << 15)   /* member of trajectories_all */
#define FRX_FLAG_COSTED (1u << 16)    /* cost function evaluated */
#define FRX_FLAG_CANDIDATE (1u << 17) /* handed to the collision check / arg-min */
/* index k of the first ego hull (obb-sum of boxes k, k + 1; time index t0 + k) that met a predicted obstacle / a static box:
 * what pycrcc's trajectories_collision_static_obstacles returns as "leaving_road_at" (planner.py:362-372, where the
 * velocity at that index feeds boundary_harm).  Valid when FRX_FLAG_COLLIDE / FRX_FLAG_BOUNDARY is set. */
#define FRX_FLAG_COLLIDE_STEP_SHIFT 18
#define FRX_FLAG_BOUNDARY_STEP_SHIFT 24
#define FRX_FLAG_COLLIDE_STEP(f) (((f) >> FRX_FLAG_COLLIDE_STEP_SHIFT) & 63u)
#define FRX_FLAG_BOUNDARY_STEP(f) (((f) >> FRX_FLAG_BOUNDARY_STEP_SHIFT) & 63u)

/* cost term ids, alphabetical like the reference's evaluation order */
enum {
    FRX_COST_ACCELERATION = 0, FRX_COST_DISTANCE_TO_OBSTACLES, FRX_COST_DISTANCE_TO_REFERENCE_PATH,
    FRX_COST_JERK, FRX_COST_LATERAL_JERK, FRX_COST_LONGITUDINAL_JERK, FRX_COST_ORIENTATION_OFFSET,
    FRX_COST_PATH_LENGTH, FRX_COST_PREDICTION, FRX_COST_VELOCITY_OFFSET, FRX_NUM_COST_TERMS
};
#define FRX_MAX_COSTS 10

typedef struct frx_ctx frx_ctx;

typedef struct frx_params {
    double dt;               /* planning.dt */
    int32_t N;               /* int(horizon / dt); Nt = N + 1 samples per candidate; N <= 63 */
    int32_t low_vel_mode;    /* Planner._LOW_VEL_MODE */
    int32_t draw_traj_set;   /* debug.draw_traj_set */
    int32_t kinematic_debug; /* debug.kinematic_debug */
    double a_max, v_switch, delta_max, wheelbase, wb_rear_axle, length, width;
    double x0_orientation;   /* x_0.orientation */
    double desired_velocity;
    int32_t n_costs;                       /* active cost terms, name-sorted */
    int32_t cost_ids[FRX_MAX_COSTS];       /* FRX_COST_* */
    double cost_weights[FRX_MAX_COSTS];
    int32_t store_states;    /* 1: materialise the 14 state fields of every candidate and step in HBM (default) */
    int32_t check_collisions;/* 1: OBB sweep vs predictions / static boxes for every candidate */
    /* ---- the use_cpp = True flavour of the reference (reactive_planner_cpp.py:96-178); all 0 = the Python path */
    int32_t curvature_rate_from_v_delta; /* 1: |kappa_dot| <= v_delta_max / (wheelbase cos^2(steering angle)), what
                                            CheckCurvatureRateConstraint(wheelbase, velocityDeltaMax) stands for
                                            (reactive_planner_cpp.py:109-112; reactive_planner.py:513-515), instead of 0.4 */
    int32_t velocity_offset_norm;        /* 2: CalculateVelocityOffsetCost(..., norm_order=2) (reactive_planner_cpp.py:170-178):
                                            squared instead of absolute offsets over the second half of the horizon */
    double v_delta_max;                  /* vehicle.v_delta_max (steering rate limit), used when curvature_rate_from_v_delta */
    int32_t prediction_cost_mode;        /* 1: the prediction term is CalculateCollisionProbabilityFast (reactive_planner_cpp.py:
                                            151-155), i.e. get_collision_probability_fast (risk_assessment/
                                            collision_probability.py:141-261) summed over steps and obstacles; 0: inverse
                                            Mahalanobis distance (python path) */
    int32_t reserved_;
} frx_params;

typedef struct frx_result {
    int64_t argmin;            /* global row index of the selected candidate, -1 if none */
    double min_cost;           /* its total cost (+inf if none) */
    int64_t n_rows;            /* rows evaluated by this call */
    int64_t n_in_list;         /* len(trajectories_all) */
    int64_t n_feasible;        /* valid and feasible members of the list */
    int64_t n_candidates;      /* handed to the collision check */
    int64_t n_collide;         /* candidates overlapping a predicted obstacle */
    int64_t n_boundary;        /* candidates overlapping a static box */
    int64_t collision_counter; /* colliding candidates the lazy reference loop would have visited */
    int64_t reason_counts[11]; /* _infeasible_count_kinematics (slot 0 = infeasible-or-invalid in list) */
    float eval_kernel_ms;      /* device time of the plan's kernels (eval kernel + obstacle kernel when the obstacle pass
                                  runs as a kernel of its own), CUDA events on the ctx stream */
    float total_device_ms;     /* first H2D to last D2H of this call, CUDA events */
    float obstacle_kernel_ms;  /* share of eval_kernel_ms spent in the obstacle kernel; 0 when the pass is fused */
    float reserved_;
} frx_result;

int frx_abi_version(void);
int frx_create(int device_ordinal, frx_ctx** out);
int frx_destroy(frx_ctx* ctx);
const char* frx_last_error(const frx_ctx* ctx);

int frx_set_reference(frx_ctx* ctx, int32_t M, const double* ref_pos, const double* ref_theta,
                      const double* ref_curv, const double* ref_curv_d, const double* ref_x,
                      const double* ref_y);
/* Same tables built ON THE DEVICE from the polyline [M][2] (CoordinateSystem.__init__, utils_coordinate_system.py:203-207):
 * a host without numpy / CCosy hands over the smoothed reference path only.  frx_get_reference reads the six tables back
 * (out[6][M]: pos, theta, curv, curv_d, x, y). */
int frx_set_reference_polyline(frx_ctx* ctx, int32_t M, const double* xy);
int frx_get_reference(frx_ctx* ctx, int32_t M, double* out);
/* Planner._compute_initial_states (frenetix_motion_planner/planner.py:567-635) on the device, including the projection
 * (x, y) -> (s, d) the reference asks CCosy for.  x0 = {x, y, orientation, velocity, acceleration, steering_angle} of the
 * rear axle; x_cl = {s, s', s'', d, d', d''}.  FRX_ERR_INVALID if the curvilinear velocity comes out negative (the
 * reference raises). */
int frx_initial_state(frx_ctx* ctx, const double* x0, int32_t low_vel_mode, double wheelbase, double* x_cl);
int frx_set_params(frx_ctx* ctx, const frx_params* p);
/* nT distinct durations; traj_len[k] samples are valid for T_values[k]; tpow[k][p][i] (p = 0..4 for
 * t, t^2 .. t^5; i < Nt) row-major [nT][5][Nt] */
int frx_set_time_tables(frx_ctx* ctx, int32_t nT, const double* T_values, const int32_t* traj_len,
                        const double* tpow);
/* O obstacles x T steps: pos [O][T][2], cov [O][T][2][2], theta [O][T]; half_len/half_wid [O];
 * len_valid[O] = number of valid steps (<= T).  O = 0 clears. */
int frx_set_predictions(frx_ctx* ctx, int32_t O, int32_t T, const double* pos, const double* cov,
                        const double* theta, const double* half_len, const double* half_wid,
                        const int32_t* len_valid);
int frx_set_obstacle_positions(frx_ctx* ctx, int32_t n, const double* pos_xy);
/* B boxes [B][5] = cx, cy, theta, half_len, half_wid.  B = 0 clears. */
int frx_set_static_obbs(frx_ctx* ctx, int32_t B, const double* obb);

/* rows: host [N][13] = t0,t1,s0,ss0,sss0,ss1,sss1,d0,dd0,ddd0,d1,dd1,ddd1.  row_index_base is
 * added to local row numbers in frx_result.argmin (shards of a larger matrix).  A pageable matrix is staged with
 * one cudaMemcpyAsync; a pinned (page-locked) one is read by the eval kernel in place -- every warp prefetches the
 * rows of its next tile over PCIe while it evaluates the current one (FRX_ZEROCOPY=0 forces the staged copy). */
int frx_plan(frx_ctx* ctx, int64_t N, const double* sampling, int64_t row_index_base, frx_result* out);
/* same, sampling already in device memory (pointer from cudaMalloc / torch) */
int frx_plan_device(frx_ctx* ctx, int64_t N, const void* d_sampling, int64_t row_index_base,
                    frx_result* out);
/* Asynchronous form: enqueue the plan on the context's stream and return; frx_plan_wait synchronises and
 * fills the result.  Lets a caller queue the multi-GPU exchange (or its own kernels) behind the plan
 * before blocking once. */
int frx_plan_device_async(frx_ctx* ctx, int64_t N, const void* d_sampling, int64_t row_index_base);
int frx_plan_wait(frx_ctx* ctx, frx_result* out);
/* rows generated on device in itertools.product order (t1 slowest, then ss1, then d1);
 * x_cl = s0,ss0,sss0,d0,dd0,ddd0; evaluates global rows [row_first, row_first + row_count) */
int frx_plan_grid(frx_ctx* ctx, int32_t nt, const double* t1, int32_t nv, const double* ss1,
                  int32_t nd, const double* d1, const double* x_cl, int64_t row_first,
                  int64_t row_count, frx_result* out);

/* Multi-agent batch (main_multiagent.py; cr_scenario_handler/simulation/agent_batch.py:186-189 loops
 * agent.step_agent() one after the other): ONE eval-kernel launch evaluates the sampling matrices of all
 * agents.  ctxs[a] is agent a's own context (reference path, params, predictions set as for frx_plan);
 * all contexts must live on one device and share the planning horizon; the launch runs on ctxs[0]'s stream;
 * results[a] is agent a's frx_result, read-back per context as usual. */
int frx_plan_batched(int32_t n_agents, frx_ctx** ctxs, const int64_t* n_rows, const double* const* samplings,
                     frx_result* results);

/* read-back (rows are LOCAL indices of the last plan call) */
int32_t frx_state_pitch(const frx_ctx* ctx);
/* kernels of this library the last plan on ctx launched (eval kernel, obstacle kernel, collision counter): what a
 * benchmark reports as its launch count */
int32_t frx_last_launches(const frx_ctx* ctx);
int frx_get_states(frx_ctx* ctx, int64_t n_idx, const int64_t* idx, uint32_t field_mask, double* out);
int frx_get_states_range(frx_ctx* ctx, int64_t first, int64_t count, uint32_t field_mask, double* out);
int frx_get_costs(frx_ctx* ctx, int64_t first, int64_t count, double* costs, double* total);
int frx_get_flags(frx_ctx* ctx, int64_t first, int64_t count, uint32_t* flags, int32_t* traj_len);
/* state rows of the selected candidate (frx_result.argmin) of the last plan: out[n_fields][frx_state_pitch()].
 * The eval kernel's last CTA writes them into the mapped result record together with the arg-min, so this is a
 * host-side copy -- no device round trip (reference: the optimal trajectory handed back by plan(),
 * frenetix_motion_planner/reactive_planner.py:89-94) */
int frx_winner_states(frx_ctx* ctx, uint32_t field_mask, double* out);
/* ... and its scalars from the same record: flags, traj_len, total cost, the n_costs unweighted cost terms */
int frx_winner_record(frx_ctx* ctx, uint32_t* flags, int32_t* traj_len, double* total, double* costs);
/* raw device pointers of the last plan for zero-copy consumers: states [ceil(N / 32)][Nt][14][32] (element (field f,
 * step i) of candidate r at (((r / 32) * Nt + i) * 14 + f) * 32 + r % 32), costs [N][n_costs], total [N], flags [N] */
int frx_device_pointers(frx_ctx* ctx, void** states, void** costs, void** total, void** flags);
/* device address of the 16-byte winner record {double min_cost; int64 global_row} of the last plan:
 * the payload of the multi-GPU arg-min exchange (all-gather of 16 B per rank, no host round trip) */
int frx_winner_device_pointer(frx_ctx* ctx, void** winner);
/* Multi-GPU arg-min exchange without a collective kernel (one node): `page` is FRX_EXCHANGE_PAGE_BYTES of host memory
 * SHARED by the ranks (POSIX shm), zero-initialised once.  After frx_set_exchange the last CTA of every plan on this
 * context also stores {min_cost, global row, epoch} into this rank's slot of the page (a posted PCIe write next to the
 * result record), and frx_exchange_wait -- called after the plan has been waited for -- spins until every rank's slot
 * shows the same plan epoch, then reduces them deterministically (lowest cost, ties -> lowest global row): semantically
 * the all-reduce(argmin) of SURVEY 8e, with no NCCL launch, no second device round trip and 64 B per rank on the wire.
 * All ranks must issue the same sequence of plans.  page = NULL detaches. */
#define FRX_EXCHANGE_PAGE_BYTES 8192
int frx_set_exchange(frx_ctx* ctx, void* page, int32_t rank, int32_t world);
int frx_exchange_wait(frx_ctx* ctx, int64_t timeout_us, double* min_cost, int64_t* global_row, int32_t* owner_rank);
/* diagnostics: the kernels' slow-path-free fp64 division next to IEEE division (tests only) */
int frx_selftest_fdiv(frx_ctx* ctx, int64_t n, const double* a, const double* b, double* q_fdiv, double* q_ieee);
/* diagnostics: the kernels' division by a plan constant b (dt, 100000, Nt; reciprocal precomputed on the host)
 * next to IEEE division (tests only) */
int frx_selftest_divc(frx_ctx* ctx, int64_t n, const double* a, double b, double* q_divc, double* q_ieee);
/* diagnostics: measured fp64 throughput of this GPU (dependent-free DFMA streams on every SM, 2 flop per FMA) in TFLOP/s:
 * the roofline denominator of the fp64-bound obstacle kernel (bench.py) */
int frx_selftest_fp64_peak(frx_ctx* ctx, double* tflops);
/* use an externally created stream (e.g. torch's current stream); 0 restores the private stream */
int frx_set_stream(frx_ctx* ctx, void* cuda_stream);
int frx_synchronize(frx_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* FRX_H */
