/*
 * rp_b200.h -- C-ABI of the B200-native candidate-trajectory engine.
 *
 * Drop-in boundary for ONE hot path of commonroad-reactive-planner (reference paths relative to
 * /root/reference): everything ReactivePlanner.plan() does between sampling the (t, lon, d) end
 * states and returning the optimal trajectory sample:
 *
 *   sampling.py:202-242            enumeration  t x lon x (d U {d0})
 *   polynomial_trajectory.py:292-360  quintic / quartic coefficient solve
 *   reactive_planner.py:715-969    _check_kinematics (evaluation, Frenet->Cartesian, limits, extension)
 *   reactive_planner.py:971-1017   _check_constraints
 *   cost_function.py:51-71, :85-92 DefaultCostFunction / DefaultCostFunctionFailSafe
 *   trajectories.py:502-510        TrajectoryBundle.sort  (here: feasible arg-min)
 *   reactive_planner.py:1019-1063  _check_collisions (discrete ego-vs-obstacle check)
 *   reactive_planner.py:1065-1136  _get_optimal_trajectory
 *
 * The reference has no FFI today (its operator API is the Python class surface); these are the
 * entry points a binding for that path would call.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions: plain pointers and sizes, host buffers unless the name says "_dev"; all floating
 * point is IEEE fp64; every function returns 0 on success and a negative rp_status on failure,
 * with a message available from rp_last_error().  "No feasible candidate" is NOT an error
 * (winner == -1).  A context is bound to one CUDA device and one stream; calls on one context
 * are not re-entrant, distinct contexts may be used from distinct threads.
 */
#ifndef RP_B200_H
#define RP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rp_ctx rp_ctx;

enum rp_status {
    RP_OK = 0,
    RP_ERR_ARG = -1,      /* invalid argument            */
    RP_ERR_CUDA = -2,     /* CUDA runtime failure        */
    RP_ERR_STATE = -3,    /* call order (missing tables) */
    RP_ERR_NOMEM = -4,
    RP_ERR_PEER = -5      /* multi-GPU peer exchange failed; the group is closed and must be re-created on every rank */
};

/* candidate status (FeasibilityStatus, trajectories.py:18-22; FILTERED = filter_goals_behind, :545-550) */
enum rp_cand_status { RP_FEASIBLE = 0, RP_INFEASIBLE_KINEMATIC = 1, RP_INFEASIBLE_COLLISION = 2, RP_FILTERED = 3,
                      RP_FEASIBLE_UNCHECKED = 4 /* lazy collision mode: kinematically feasible, never visited */ };

/* reason of a kinematic rejection: keys of infeasible_reason_dict (reactive_planner.py:799,803,979-1016),
 * plus the projection-domain rejection (:911-917) which the reference counts but does not name. */
enum rp_reason {
    RP_R_NONE = 0, RP_R_VELOCITY = 1, RP_R_ACCELERATION = 2, RP_R_KAPPA = 3, RP_R_KAPPA_DOT = 4,
    RP_R_YAW_RATE = 5, RP_R_PROJECTION = 6, RP_R_REF_RANGE = 7, RP_N_REASONS = 8
};

/* constraints_to_check bit mask (utility/config.py:127-128) */
enum rp_constraint_bits {
    RP_C_VELOCITY = 1, RP_C_ACCELERATION = 2, RP_C_KAPPA = 4, RP_C_KAPPA_DOT = 8, RP_C_YAW_RATE = 16
};

enum rp_lon_mode { RP_VELOCITY_KEEPING = 0, RP_STOPPING = 1 };          /* sampling.py:244-266 */
enum rp_cost_kind { RP_COST_DEFAULT = 0, RP_COST_FAILSAFE = 1, RP_COST_NONE = 2 };

/* order of the 14 state rows returned by rp_fetch_states (CartesianSample / CurviLinearSample,
 * trajectories.py:61-332) */
enum rp_state_row {
    RP_X = 0, RP_Y, RP_THETA, RP_V, RP_A, RP_KAPPA, RP_KAPPA_DOT,
    RP_S, RP_D, RP_THETA_CL, RP_S_DOT, RP_S_DDOT, RP_D_DOT, RP_D_DDOT, RP_N_STATE_ROWS
};

/* VehicleConfiguration (utility/config.py:194-222); kappa_max = tan(delta_max) / wheelbase is
 * computed by the HOST (same expression as :222) so both sides compare against the same double. */
typedef struct rp_vehicle_params {
    double length, width;
    double wb_rear_axle;
    double wheelbase;
    double a_max, v_switch;
    double delta_max, v_delta_max;
    double kappa_max;
} rp_vehicle_params;

/* per-cycle scalar inputs of plan() (reactive_planner.py:570-665) */
typedef struct rp_plan_inputs {
    double x0_lon[3];            /* s, s_dot, s_ddot                       (:591)           */
    double x0_lat[3];            /* d, d_dot, d_ddot                                         */
    double x0_orientation;       /* x_0.orientation                        (:866)           */
    int32_t x0_time_step;        /* x_0.time_step                          (:1040)          */
    int32_t low_vel_mode;        /* x_0.velocity < low_vel_mode_threshold  (:594)           */
    int32_t lon_mode;            /* rp_lon_mode                                              */
    int32_t N;                   /* time_steps_computation                                   */
    double dt;
    int32_t factor;              /* config.planning.factor                 (:1040)          */
    int32_t draw_all;            /* _draw_traj_set: no pre-filter, no early exit (:97, :796, :903) */
    uint32_t constraint_mask;    /* rp_constraint_bits                                       */
    int32_t cost_kind;           /* rp_cost_kind                                             */
    int32_t has_desired_speed;   /* DefaultCostFunction.desired_speed is not None            */
    int32_t has_desired_s;       /* DefaultCostFunction.desired_s is not None                */
    double desired_speed, desired_s, desired_d, w_a;
    int32_t want_all_states;     /* 1: keep the 14 x (N+1) state block of EVERY candidate     */
    int32_t check_collision;     /* 0: skip a13; 1: check every kinematically feasible candidate; 2: lazy, like the
                                    reference's cost-ordered pass (:1031-1063): candidates costlier than the best
                                    collision-free one found so far need not be visited and stay
                                    RP_FEASIBLE_UNCHECKED (the step-parallel schedule skips them as it goes; the
                                    candidate-major schedule and the scenario batch store the ego boxes while they
                                    march and check afterwards only what can be ranked before the winner).  Every
                                    candidate ranked up to the winner has its verdict; winner, winner_cost,
                                    n_infeasible_collision and the kinematic counters are identical in modes 1
                                    and 2; n_collision_total counts the colliders that were visited.      */
    int32_t continuous_collision_check;   /* config.planning.continuous_collision_check (:1049-1058): the first discretely
                                    collision-free candidate in cost order is re-checked with the OBB-sum hulls of its
                                    consecutive poses (time indices x0_time_step + i); a hit ends the level without a
                                    trajectory, as the reference's `break` does.  The caller uploads dynamic obstacles
                                    already replaced by their hulls (:240-241). */
    int32_t reserved_;
} rp_plan_inputs;

typedef struct rp_plan_result {
    int32_t winner;                        /* enumeration index of the optimal candidate, -1 if none   */
    int32_t n_candidates;
    int32_t n_feasible;                    /* kinematically feasible (before the collision check)     */
    int32_t n_infeasible_kinematics;       /* ReactivePlanner.infeasible_count_kinematics (:1119)     */
    int32_t n_infeasible_collision;        /* colliders ranked before the winner (:1043; App. B#12)   */
    int32_t n_collision_total;             /* kinematically feasible candidates found colliding (all of them in mode 1) */
    int32_t reason_counts[RP_N_REASONS];   /* infeasible_reason_dict (+ projection)                   */
    double winner_cost;
} rp_plan_result;

/* ---- lifetime ------------------------------------------------------------------------------ */
const char* rp_last_error(void);
int rp_version(void);
/* stream: a cudaStream_t created by the caller on `device` (e.g. torch.cuda.current_stream().cuda_stream),
 * or NULL to let the context create its own. */
int rp_ctx_create(int device, void* stream, rp_ctx** out);
int rp_ctx_destroy(rp_ctx* ctx);
int rp_ctx_synchronize(rp_ctx* ctx);

/* ---- scenario-static tables (set once, device resident) ------------------------------------- */
/* VehicleConfiguration -> constants of a6/a7/a13 */
int rp_ctx_set_vehicle(rp_ctx* ctx, const rp_vehicle_params* vp);

/* CoordinateSystem (utility/utils_coordinate_system.py:113-118, :167-174): the planner's four
 * reference arrays plus the curvilinear frame's own polyline tables for (s,d)->(x,y):
 * path_xy[n][2] vertices, path_s[n] cumulative length, path_normal_xy[n][2] per-vertex
 * pseudo-normals, proj_limit = lateral projection-domain limit. */
int rp_ctx_set_reference(rp_ctx* ctx, int n_pts, const double* ref_pos, const double* ref_theta,
                         const double* ref_curv, const double* ref_curv_d, const double* path_xy,
                         const double* path_s, const double* path_normal_xy, double proj_limit);

/* The same tables DERIVED ON THE DEVICE from the (smoothed, de-duplicated) reference polyline xy[n_pts][2] -- what
 * CoordinateSystem.__init__ does on the host (utility/utils_coordinate_system.py:101-118: frame construction with one
 * extension vertex eps2 beyond each end and per-vertex pseudo-normals, then compute_pathlength / orientation (+ np.unwrap)
 * / curvature_from_polyline and np.gradient for the curvature rate, in numpy's operation order).  The context then
 * holds n_pts + 2 table rows.  For batches of thousands of scenarios the per-scenario host numpy pass is the set-up
 * cost this replaces. */
int rp_ctx_set_reference_polyline(rp_ctx* ctx, int n_pts, const double* xy, double proj_limit, double eps2);
/* read the context's reference tables back (any of the array pointers may be NULL); capacity = rows the buffers hold;
 * capacity <= 0: only *n_pts is written (size query) */
int rp_ctx_get_reference(rp_ctx* ctx, int capacity, int* n_pts, double* ref_pos, double* ref_theta, double* ref_curv,
                         double* ref_curv_d, double* path_xy, double* path_s, double* path_normal_xy);

/* pycrcc.CollisionChecker content (reactive_planner.py:234-251):
 *   static_obb[n_static][5]   = cx, cy, theta, half_length, half_width   (static obstacles + OBB road boundary)
 *   dynamic obstacle o: time indices dyn_t0[o] .. dyn_t0[o]+dyn_len[o]-1, boxes dyn_obb[sum(len)][5] concatenated
 *   tris[n_tri][6]            = x1, y1, x2, y2, x3, y3                   (triangulated road boundary)
 * cell_size: edge of the uniform broad-phase grid over the static primitives (<= 0: automatic, 0.5 m or the
 * smallest power-of-two multiple that keeps the table below 2^20 cells). */
int rp_ctx_set_obstacles(rp_ctx* ctx, int n_static, const double* static_obb, int n_dyn,
                         const int32_t* dyn_t0, const int32_t* dyn_len, const double* dyn_obb,
                         int n_tri, const double* tris, double cell_size);

/* Which kernel evaluates a4-a13 for the main launch.  Both give identical bits; they differ in schedule:
 *   STEP_PARALLEL    one thread per (candidate, time step)  -- replanning-size bundles, state output, draw mode
 *   CANDIDATE_MAJOR  one thread per candidate, sequential in time like reactive_planner.py:790-935 -- large
 *                    bundles in select-only mode (falls back to STEP_PARALLEL for want_all_states / draw_all /
 *                    N + 1 > 128)
 *   AUTO             CANDIDATE_MAJOR from ~24k candidates up */
enum rp_kernel_policy { RP_KERNEL_AUTO = 0, RP_KERNEL_STEP_PARALLEL = 1, RP_KERNEL_CANDIDATE_MAJOR = 2 };
int rp_ctx_set_kernel_policy(rp_ctx* ctx, int policy);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* Grid form (FixedIntervalSampling.generate_trajectories_at_level, sampling.py:202-242): the three
 * ORDERED sample lists as the host's Python sets iterate them; traj_len[i] =
 * len(np.arange(0, np.round(t[i] + dt, 5), dt)) (reactive_planner.py:733).  Candidate enumeration
 * index = (i_t * n_lon + i_lon) * n_d + i_d.  lon = target velocities (velocity_keeping) or
 * target positions (stopping). */
int rp_plan_grid(rp_ctx* ctx, const rp_plan_inputs* in, int n_t, const double* t, const int32_t* traj_len,
                 int n_lon, const double* lon, int n_d, const double* d, rp_plan_result* out);

/* The same three stages separately, so that a caller can keep inputs resident in HBM:
 *   upload: host -> device copy of the inputs (asynchronous on the context's stream)
 *   launch: coefficient solve + fused evaluation + arg-min + winner states (no host sync)
 *   result: device -> host copy of rp_plan_result and synchronisation */
int rp_grid_upload(rp_ctx* ctx, const rp_plan_inputs* in, int n_t, const double* t, const int32_t* traj_len,
                   int n_lon, const double* lon, int n_d, const double* d);
int rp_grid_launch(rp_ctx* ctx);
int rp_grid_result(rp_ctx* ctx, rp_plan_result* out);

/* One replanning cycle in ONE launch.  plan() escalates the sampling level until a level yields a feasible,
 * collision-free candidate (reactive_planner.py:616-636); the levels are independent bundles, so several of them can be
 * evaluated together and the LOWEST level with a winner selected -- the result the loop would have produced.  The
 * cycle's whole input (this call's arguments) travels inside the kernel launch, each candidate's two polynomials are
 * solved in shared memory, the last block of the launch selects level by level, gathers the winner's state block and
 * writes the records into mapped host memory: no host->device copy, no coefficient / selection launches, no
 * device->host copy, no event.
 *   n_t / n_lon / n_d [n_levels]; t_cat / traj_len_cat, lon_cat, d_cat: the levels' ORDERED sample lists concatenated in
 *   level order (same meaning as in rp_plan_grid);
 *   out[j], j < *n_evaluated: the record of level j (winner = enumeration index WITHIN the level); *chosen = the level
 *   whose record is final = *n_evaluated - 1: the first level with a winner, else the last level.  Levels above it are
 *   not "evaluated" in the reference's sense (the loop would not have reached them) and have no record.
 * Afterwards rp_fetch_states / rp_fetch_candidates / rp_fetch_coeffs address the chosen level; rp_select_level switches
 * to another evaluated level.  Limits (rp_cycle_limits): <= 4 levels, sum of samples <= max_samples, sampled horizons
 * <= max_segments, (sum of candidates) x (N + 1) <= max_work, N + 1 <= 256; no sharding, no continuous collision check,
 * and cost_kind NONE only with one level.  Callers fall back to one rp_plan_grid per level beyond them. */
int rp_plan_levels(rp_ctx* ctx, const rp_plan_inputs* in, int n_levels, const int32_t* n_t, const int32_t* n_lon,
                   const int32_t* n_d, const double* t_cat, const int32_t* traj_len_cat, const double* lon_cat,
                   const double* d_cat, rp_plan_result* out, int32_t* n_evaluated, int32_t* chosen);
int rp_select_level(rp_ctx* ctx, int level);
/* the winner-state area of the mapped result block of rp_plan_levels: 14 x (N + 1) doubles of the CHOSEN level's
 * winner after every call (the address changes only when N grows) -- callers that poll results at replanning rate read
 * it in place instead of calling rp_fetch_states */
int rp_cycle_host_block(rp_ctx* ctx, void** states, int64_t* n_doubles);
int rp_cycle_limits(int32_t* max_levels, int32_t* max_samples, int32_t* max_segments, int64_t* max_work);

/* List form (any SamplingSpace, e.g. CorridorSampling, sampling.py:340-397): per-candidate
 * polynomial coefficients as the host built them; skip[i] != 0 marks candidates dropped by
 * filter_goals_behind (may be NULL). */
int rp_plan_list(rp_ctx* ctx, const rp_plan_inputs* in, int n_cand, const double* coeffs_lon,
                 const double* coeffs_lat, const int32_t* traj_len, const uint8_t* skip,
                 rp_plan_result* out);

/* Shard of the enumeration space [first, first+count) for multi-GPU bundles; affects the next
 * rp_grid_launch / rp_plan_grid.  count < 0 resets to the whole bundle. */
int rp_set_candidate_range(rp_ctx* ctx, int first, int count);

/* Shard `rank` of `world` of a GRID bundle by interleaving the lon samples: the rank owns lon indices rank, rank + world,
 * ... of every sampled t (and all d), so every rank's shard has the same mix of horizons (contiguous t-major tiles
 * differ in traj_len and the exchange waits for the slowest).  Replaces a candidate range and vice versa; world <= 1
 * resets to the whole bundle.  Results, counters and the peer exchange work as for ranges; winner indices are the
 * bundle's enumeration indices. */
int rp_set_candidate_stripe(rp_ctx* ctx, int rank, int world);

/* Multi-GPU arg-min plumbing for sharded bundles (device pointers, asynchronous on the context's stream):
 * rp_export_record_dev writes [best cost (+inf if none), best enumeration index as double (+inf if none),
 * n_infeasible_kinematics, n_feasible] of this rank's shard to dev_dst4 for an NCCL all-gather;
 * rp_count_colliders_before_dev writes to dev_out1 how many of this shard's colliding candidates rank
 * before the GLOBAL winner dev_winner2 = [cost, index] (lazy collision count, reactive_planner.py:1031-1063). */
int rp_export_record_dev(rp_ctx* ctx, double* dev_dst4);
/* merge of the all-gathered records dev_gathered[world][4]: dev_winner2 = lexicographic min on (cost, index),
 * dev_totals2 = [sum n_infeasible_kinematics, sum n_feasible] */
int rp_merge_records_dev(rp_ctx* ctx, const double* dev_gathered, int world, double* dev_winner2, double* dev_totals2);
int rp_count_colliders_before_dev(rp_ctx* ctx, const double* dev_winner2, double* dev_out1);

/* The same exchange over peer-mapped memory instead of NCCL (one process per GPU of ONE box, NVLink / NVSwitch):
 * the reference gathers its workers' results through fork + pickle (reactive_planner.py:1084-1111); here every rank
 * owns a small mailbox that all ranks map through CUDA IPC.  While a peer group is open and a candidate range is set,
 * rp_grid_launch / rp_plan_grid end with two kernels that store the shard's record and the shard's collider count
 * straight into the peers' mailboxes and merge what arrived -- no collective call on the data path -- and
 * rp_grid_result returns the GLOBAL result on every rank (winner, cost, all counters; rp_fetch_states(winner) the
 * global winner's states).  Every rank must launch the same cycles.  Failure protocol: the per-launch epochs of the
 * ranks must stay equal, so a rank that never arrives (or whose launch failed before it reached the exchange) makes
 * the others' waits time out after ~8 s; a rank whose first wait fails still publishes the second phase's flag with
 * a poison count, so its peers fail that cycle at once instead of waiting another 8 s.  rp_grid_result then returns
 * RP_ERR_PEER and the context leaves the group (launches go back to shard-local results).  Recovery: EVERY rank calls
 * rp_peer_close, rp_peer_create and rp_peer_open again (epochs restart at zero).  rp_peer_create zeroes the mailbox:
 * it must not be called while any rank still has a cycle of the old group in flight.
 *   rp_peer_create  allocates this rank's mailbox and returns its 64-byte CUDA IPC handle in handle64;
 *   rp_peer_open    maps the mailboxes of all ranks (handles[world][64] in rank order, e.g. from one all-gather at
 *                   set-up); every rank must have returned from rp_peer_create before any rank launches a cycle;
 *   rp_peer_close   unmaps / frees (also done by rp_ctx_destroy); launches go back to shard-local results. */
int rp_peer_create(rp_ctx* ctx, unsigned char* handle64);
int rp_peer_open(rp_ctx* ctx, int rank, int world, const unsigned char* handles);
int rp_peer_close(rp_ctx* ctx);

/* ---- batches of independent scenarios (reactive_planner.py has no counterpart: one ReactivePlanner per scenario
 * and process; BASELINE configs[4]) ------------------------------------------------------------------------------
 * A batch groups contexts of ONE device -- each with its own vehicle, reference and obstacle tables -- and
 * evaluates one replanning cycle of all of them with one host->device copy, four launches -- seven when every scenario
 * asks for the lazy collision pass (check_collision = 2: ego boxes stored by the march, deferred checker) -- (no per-scenario launch)
 * and one device->host copy.  Select-only mode (no draw_all / want_all_states), N + 1 <= 128.  stream: as in
 * rp_ctx_create (NULL: the first context's stream).  The contexts stay usable on their own between batch cycles. */
typedef struct rp_batch rp_batch;
int rp_batch_create(rp_ctx* const* ctxs, int n, void* stream, rp_batch** out);
int rp_batch_destroy(rp_batch* b);
int rp_batch_size(rp_batch* b);
/* inputs of scenario k for the next rp_batch_launch (same arguments as rp_grid_upload; host side only) */
int rp_batch_set_inputs(rp_batch* b, int k, const rp_plan_inputs* in, int n_t, const double* t, const int32_t* traj_len,
                        int n_lon, const double* lon, int n_d, const double* d);
/* the same for ALL scenarios in one call: in[rp_batch_size], per-scenario counts, and the sample lists of all scenarios
 * concatenated in scenario order (t_cat / traj_len_cat: sum n_t entries, lon_cat: sum n_lon, d_cat: sum n_d) */
int rp_batch_set_inputs_all(rp_batch* b, const rp_plan_inputs* in, const int32_t* n_t, const int32_t* n_lon, const int32_t* n_d,
                            const double* t_cat, const int32_t* traj_len_cat, const double* lon_cat, const double* d_cat);
int rp_batch_launch(rp_batch* b);                              /* asynchronous on the batch's stream */
int rp_batch_results(rp_batch* b, rp_plan_result* out);        /* out[rp_batch_size]; synchronises */
int rp_batch_fetch_candidates(rp_batch* b, int k, double* cost, int32_t* status, int32_t* reason, int32_t* step);
/* closed-loop batches (run_planner.py:84-107 for every scenario at once): the state of each scenario's winner at time step
 * `step` of its trajectory, where the next replanning cycle starts -- one launch.  out[rp_batch_size][16] = x, y, theta, v,
 * a, kappa, s, s_dot, s_ddot, d, d_dot, d_ddot, valid (0: the scenario has no winner), 3 unused */
int rp_batch_winner_states(rp_batch* b, int step, double* out);
/* device time of the last rp_batch_launch (all its launches) and the candidates it evaluated */
int rp_batch_last_ms(rp_batch* b, float* ms, long long* n_candidates);

/* ---- results of the last plan call ----------------------------------------------------------- */
/* 14 x (N+1) state rows (rp_state_row order) of candidate idx; available for the winner always,
 * for every candidate when want_all_states was set.  Other indices are re-evaluated on demand. */
int rp_fetch_states(rp_ctx* ctx, int idx, double* out);
/* per-candidate arrays of length n_candidates (any pointer may be NULL) */
int rp_fetch_candidates(rp_ctx* ctx, double* cost, int32_t* status, int32_t* reason, int32_t* step);
/* coefficients as solved on the device: lon[n][6], lat[n][6], delta_tau_lat[n] */
int rp_fetch_coeffs(rp_ctx* ctx, double* coeffs_lon, double* coeffs_lat, double* delta_tau_lat);

/* ---- the step before the path ------------------------------------------------ */
/* ReactivePlanner._compute_initial_states (reactive_planner.py:446-512) incl. pycrccosy's
 * convert_to_curvilinear_coords, batched: x0[n][6] = x, y, orientation, velocity, acceleration, steering_angle of the
 * rear-axle state; low_vel_mode[n] (d derivatives w.r.t. arc length instead of time, :498-505) ->
 * out_lon[n][3] = s, s_dot, s_ddot; out_lat[n][3] = d, d_dot, d_ddot; status[n]: 0 ok, 1 outside the projection
 * domain (the reference raises ValueError, :459-461), 2 negative s_dot (the reference raises Exception, :489-491). */
int rp_initial_states(rp_ctx* ctx, int n, const double* x0, const int32_t* low_vel_mode, double* out_lon, double* out_lat,
                      int32_t* status);
/* the same for one state per scenario of a batch (each against its own reference tables): arrays of rp_batch_size rows */
int rp_batch_initial_states(rp_batch* b, const double* x0, const int32_t* low_vel_mode, double* out_lon, double* out_lat,
                            int32_t* status);

/* ---- stand-alone pieces ---------------------------------------------------------------------- */
/* batched coefficient solve (polynomial_trajectory.py:292-360): kind 0 = quartic
 * (x_d = target velocity in xd[i][0]), 1 = quintic; x0[n][3], xd[n][3], tau[n] -> coeffs[n][6] */
int rp_solve_coeffs(rp_ctx* ctx, int n, const int32_t* kind, const double* x0, const double* xd,
                    const double* tau, double* coeffs);
/* pycrcc.CollisionChecker.collide for n ego boxes: pose[n][3] = cx, cy, theta (box centre),
 * time_idx[n]; half extents from the arguments -> hit[n] (0/1) */
int rp_collide_poses(rp_ctx* ctx, int n, const double* pose, const int32_t* time_idx,
                     double half_length, double half_width, uint8_t* hit);

/* self-test of the kernels' division: q_shared[i] = what the candidate-major kernel computes (quotient from a shared
 * refined reciprocal with a range check, plain division when the check rejects), q_plain[i] = a[i] / b[i] as the
 * compiler emits it -- the two must agree bit for bit for every input; rejected[i] (may be NULL) = 1 where the range
 * check sent the pair to the plain division */
int rp_selftest_divide(rp_ctx* ctx, int n, const double* a, const double* b, double* q_shared, double* q_plain,
                       int32_t* rejected);

/* device timing of the last rp_grid_launch stages in milliseconds: [coeff, fused, argmin, winner] */
int rp_last_stage_ms(rp_ctx* ctx, float* ms4);
/* the same for the launch `back` launches ago (0 = last; the context keeps the last 64) */
int rp_stage_ms(rp_ctx* ctx, int back, float* ms4);
/* stage timing is opt-in (off by default): the five event records per launch cost a replanning-size cycle 16 of its
 * 82 us.  rp_last_stage_ms / rp_stage_ms fail with RP_ERR_STATE for launches made while it was off. */
int rp_ctx_set_stage_timing(rp_ctx* ctx, int on);
/* measurement aid: FP64 FMA peak of the device in TFLOP/s from a DFMA micro-benchmark (roofline denominator) */
int rp_measure_fp64_peak(rp_ctx* ctx, double* tflops);
/* which kernel evaluated the main launch of the last plan: RP_KERNEL_STEP_PARALLEL or RP_KERNEL_CANDIDATE_MAJOR */
int rp_last_main_kernel(rp_ctx* ctx);
/* number of kernels rp_grid_launch enqueues (for bench.py's gpu_launches) */
int rp_launches_per_plan(rp_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* RP_B200_H */
