// rp_kernels.cuh -- the kernels of the candidate-trajectory hot path (sm_100a, fp64, no tensor cores:
// nothing on this path is a dense contraction).
//
//   coeff_kernel     a3   batched quartic / quintic end-state solves       (polynomial_trajectory.py:292-360)
//   fused_kernel     a4-a13  one thread per (candidate, time step): polynomial evaluation, reference
//                    segment lookup in shared memory, Frenet->Cartesian, kinematic limits, horizon
//                    extension, cost (numpy summation order) and the ego-vs-obstacle SAT with the
//                    dynamic obstacles of each time step staged in shared memory
//                    (reactive_planner.py:715-1063, cost_function.py:51-92, trajectories.py:168-332)
//   finalize_kernel  a12/a14  warp-shuffle feasible arg-min on (cost, enumeration index) + the cycle's
//                    counters (trajectories.py:502-510, reactive_planner.py:1065-1136)
//   collide_kernel   pycrcc.CollisionChecker.collide for a batch of ego boxes
#pragma once
#include "rp_device.cuh"
#include "rp_b200.h"
#include "rp_fused.cuh"
#include "rp_cand.cuh"

namespace rp {

struct PlanResultDev {
    rp_plan_result r;
    int n_filtered;
    int peer_error;        // multi-GPU exchange over peer-mapped memory: a wait on another rank timed out
};

// ------------------------------------------------------------------------------------------------
// a3: one thread per polynomial.  Threads [0, n_lon_sys) solve the longitudinal systems, the rest
// the lateral ones.  In low-velocity mode the lateral parameter range is the longitudinal end
// position (sampling.py:229-234), so a lateral thread first redoes its longitudinal solve.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void coeff_thread(int gid, int n_t, int n_lon, int n_d, int low_vel, int lon_mode,
                                             const double* __restrict__ t, const double* __restrict__ lon,
                                             const double* __restrict__ d, const double x0s, const double x0sd,
                                             const double x0sdd, const double x0d, const double x0dd, const double x0ddd,
                                             double* __restrict__ lon_coef, double* __restrict__ lat_coef,
                                             double* __restrict__ lat_tau) {
    const int n_lon_sys = n_t * n_lon;
    const int n_lat_sys = low_vel ? n_t * n_lon * n_d : n_t * n_d;
    if (gid >= n_lon_sys + n_lat_sys) return;
    double c[6];
    if (gid < n_lon_sys) {
        int it = gid / n_lon, il = gid - it * n_lon;
        if (lon_mode == RP_VELOCITY_KEEPING) solve_quartic(x0s, x0sd, x0sdd, t[it], lon[il], c);
        else solve_quintic(x0s, x0sd, x0sdd, lon[il], 0.0, 0.0, t[it], c);
#pragma unroll
        for (int q = 0; q < 6; ++q) lon_coef[(size_t)gid * 6 + q] = c[q];
        return;
    }
    const int q = gid - n_lon_sys;
    int it, id;
    double tau;
    if (low_vel) {
        it = q / (n_lon * n_d);
        int rem = q - it * (n_lon * n_d);
        int il = rem / n_d;
        id = rem - il * n_d;
        double cl[6];
        double tt = t[it];
        if (lon_mode == RP_VELOCITY_KEEPING) solve_quartic(x0s, x0sd, x0sdd, tt, lon[il], cl);
        else solve_quintic(x0s, x0sd, x0sdd, lon[il], 0.0, 0.0, tt, cl);
        double s_goal = position_at_end(cl, tt) - x0s;
        if (s_goal <= 0) s_goal = tt;
        tau = s_goal;
    } else {
        it = q / n_d;
        id = q - it * n_d;
        tau = t[it];
    }
    solve_quintic(x0d, x0dd, x0ddd, d[id], 0.0, 0.0, tau, c);
#pragma unroll
    for (int k = 0; k < 6; ++k) lat_coef[(size_t)q * 6 + k] = c[k];
    if (lat_tau) lat_tau[q] = tau;
}

__global__ void coeff_kernel(int n_t, int n_lon, int n_d, int low_vel, int lon_mode, const double* __restrict__ t,
                             const double* __restrict__ lon, const double* __restrict__ d, const double x0s,
                             const double x0sd, const double x0sdd, const double x0d, const double x0dd,
                             const double x0ddd, double* __restrict__ lon_coef, double* __restrict__ lat_coef,
                             double* __restrict__ lat_tau) {
    coeff_thread(blockIdx.x * blockDim.x + threadIdx.x, n_t, n_lon, n_d, low_vel, lon_mode, t, lon, d, x0s, x0sd, x0sdd, x0d,
                 x0dd, x0ddd, lon_coef, lat_coef, lat_tau);
}

// coefficient solves and the dynamic-obstacle rows of the candidate-major kernel are independent: one launch, the first
// n_coeff_blocks blocks solve, the rest write rows
struct PrepArgs {
    int n_t, n_lon, n_d, low_vel, lon_mode;
    const double *t, *lon, *d;
    double x0s, x0sd, x0sdd, x0d, x0dd, x0ddd;
    double *lon_coef, *lat_coef, *lat_tau;
    int n_coeff_blocks;
    ObstacleTables obs;
    int x0_time_step, factor, Np1;
    float r_ego_f_up, wb_rear_f_up;
    float4* dyn_rows;
    int n_dyn_blocks;
    // lateral table of the candidate-major kernel (lat_rows_thread; null: not used)
    double* lat_rows;
    const int* traj_len;
    double dt;
    // per-cycle scratch words this launch resets for the kernels after it (instead of separate memset nodes)
    int* argmin_counts;                 // [16]
    int* work_counter;                  // [kWorkWords] chunk dispenser of the candidate-major kernel + its deferred-check list lengths (may be null)
    unsigned long long* best_bits;      // lazy collision bound of the step-parallel kernel (may be null)
};

constexpr int kWorkWords = 4;          // chunk dispenser, the two list lengths of the deferred collision check, spare

__global__ void __launch_bounds__(128) prep_kernel(const __grid_constant__ PrepArgs A) {
    if (blockIdx.x == 0) {
        if (threadIdx.x < 16 && A.argmin_counts) A.argmin_counts[threadIdx.x] = 0;
        if (threadIdx.x >= 20 && threadIdx.x < 20 + kWorkWords && A.work_counter) A.work_counter[threadIdx.x - 20] = 0;
        if (threadIdx.x == 17 && A.best_bits) *A.best_bits = 0x7f7f7f7f7f7f7f7fULL;     // ~1.4e306
    }
    if ((int)blockIdx.x < A.n_coeff_blocks)
        coeff_thread(blockIdx.x * blockDim.x + threadIdx.x, A.n_t, A.n_lon, A.n_d, A.low_vel, A.lon_mode, A.t, A.lon, A.d, A.x0s,
                     A.x0sd, A.x0sdd, A.x0d, A.x0dd, A.x0ddd, A.lon_coef, A.lat_coef, A.lat_tau);
    else if ((int)blockIdx.x < A.n_coeff_blocks + A.n_dyn_blocks)
        dyn_rows_thread((blockIdx.x - A.n_coeff_blocks) * blockDim.x + threadIdx.x, A.obs, A.x0_time_step, A.factor, A.Np1,
                        A.r_ego_f_up, A.wb_rear_f_up, A.dyn_rows);
    else
        lat_rows_thread((blockIdx.x - A.n_coeff_blocks - A.n_dyn_blocks) * blockDim.x + threadIdx.x, A.n_t, A.n_d, A.Np1, A.t, A.d,
                        A.traj_len, A.x0d, A.x0dd, A.x0ddd, A.dt, A.lat_rows);
}

// ---- batch of independent scenarios (blockIdx.y = scenario; see cand_batch_kernel) --------------------------------
__global__ void coeff_batch_kernel(const PlanParams* __restrict__ params) {
    const PlanParams& P = params[blockIdx.y];
    const rp_plan_inputs& in = P.in;
    coeff_thread(blockIdx.x * blockDim.x + threadIdx.x, P.n_t, P.n_lon, P.n_d, in.low_vel_mode, in.lon_mode, P.t_samples,
                 P.lon_samples, P.d_samples, in.x0_lon[0], in.x0_lon[1], in.x0_lon[2], in.x0_lat[0], in.x0_lat[1], in.x0_lat[2],
                 const_cast<double*>(P.lon_coef), const_cast<double*>(P.lat_coef), nullptr);
}

__global__ void dyn_rows_batch_kernel(const PlanParams* __restrict__ params) {
    const PlanParams& P = params[blockIdx.y];
    const ObstacleTables& O = P.obs;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (P.dyn_rows == nullptr || q >= P.Np1 * O.n_dyn) return;
    const int step = q / O.n_dyn, o = q - step * O.n_dyn;
    const int kk = P.in.x0_time_step + step * P.in.factor - O.dyn_t0[o];
    const bool present = kk >= 0 && kk < O.dyn_len[o];
    float4 r = make_float4(1.0e30f, 1.0e30f, 0.0f, 0.0f);
    if (present) {
        const double* b = O.dyn_box + (size_t)(O.dyn_off[o] + kk) * kBoxStride;
        const float reach = (P.r_ego_f_up + (float)b[6]) * 1.000001f + O.dyn_margin;
        r = make_float4((float)(b[0] - O.org_x), (float)(b[1] - O.org_y), reach * reach * 1.00001f,
                        (reach + P.wb_rear_f_up) * 1.0001f + 1.0e-3f);
    }
    const_cast<float4*>(P.dyn_rows)[q] = r;
}

// generic batched solve (rp_solve_coeffs)
__global__ void solve_kernel(int n, const int* __restrict__ kind, const double* __restrict__ x0,
                             const double* __restrict__ xd, const double* __restrict__ tau,
                             double* __restrict__ coeffs, int* __restrict__ ok) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    double c[6];
    bool good;
    if (kind[g] == 0) good = solve_quartic(x0[3 * g], x0[3 * g + 1], x0[3 * g + 2], tau[g], xd[3 * g], c);
    else good = solve_quintic(x0[3 * g], x0[3 * g + 1], x0[3 * g + 2], xd[3 * g], xd[3 * g + 1], xd[3 * g + 2],
                              tau[g], c);
#pragma unroll
    for (int q = 0; q < 6; ++q) coeffs[(size_t)g * 6 + q] = c[q];
    ok[g] = good ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// a12/a14: arg-min over feasible, collision-free candidates on (cost, enumeration index) -- the
// element a stable ascending sort puts first -- then the counters of the cycle.  Three small launches:
// per-block partial minima + winner-independent counters, a one-block merge, and the count of colliders
// ranked before the winner (what the reference's lazy collision pass counts, App. B#12).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool lex_less(double c1, int i1, double c2, int i2) {
    return (c1 < c2) || (c1 == c2 && i1 < i2);
}

struct ArgminPartial {
    double cost;
    int idx;
    int pad;
};

struct ArgminScratch {
    int counts[16];        // [0] feasible(kin), [2] colliders total, [3] filtered, [8..15] reasons
    ArgminPartial part[512];
};

__device__ __forceinline__ void warp_lexmin(double& bc, int& bi) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double oc = __shfl_down_sync(0xffffffffu, bc, off);
        const int oi = __shfl_down_sync(0xffffffffu, bi, off);
        if (lex_less(oc, oi, bc, bi)) { bc = oc; bi = oi; }
    }
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    return v;
}

// out != nullptr: the LAST block to finish merges the partials and writes the result record (one launch fewer; the
// peer-exchange form keeps its own merge kernel, argmin_merge_kernel).  counts[1] is the blocks' ticket.
__global__ void __launch_bounds__(256) argmin_partial_kernel(const double* __restrict__ cost, const int* __restrict__ info,
                                                             int first, int count, ArgminScratch* sc, Stripe sm,
                                                             PlanResultDev* out = nullptr) {
    __shared__ double w_cost[8];
    __shared__ int w_idx[8];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double bc = __longlong_as_double(0x7ff0000000000000LL);       // +inf
    int bi = 0x7fffffff;
    int l_feas = 0, l_colt = 0, l_filt = 0;
    int l_reason[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int q = blockIdx.x * blockDim.x + tid; q < count; q += gridDim.x * blockDim.x) {
        const int k = sm.real(first + q);
        const int w = info[k];
        const int st = w & 0xFF;
        if (st == ST_FEASIBLE) {
            ++l_feas;
            const double cc = cost[k];
            if (lex_less(cc, k, bc, bi)) { bc = cc; bi = k; }
        } else if (st == ST_COLLISION) {
            ++l_feas;
            ++l_colt;
        } else if (st == ST_UNCHECKED) {
            ++l_feas;
        } else if (st == ST_KINEMATIC) {
            const int r = (w >> 8) & 0x7;
#pragma unroll
            for (int z = 0; z < 8; ++z) l_reason[z] += (r == z) ? 1 : 0;
        } else {
            ++l_filt;
        }
    }
    warp_lexmin(bc, bi);
    if (lane == 0) { w_cost[warp] = bc; w_idx[warp] = bi; }
    l_feas = warp_sum(l_feas); l_colt = warp_sum(l_colt); l_filt = warp_sum(l_filt);
#pragma unroll
    for (int z = 0; z < 8; ++z) l_reason[z] = warp_sum(l_reason[z]);
    if (lane == 0) {
        if (l_feas) atomicAdd(&sc->counts[0], l_feas);
        if (l_colt) atomicAdd(&sc->counts[2], l_colt);
        if (l_filt) atomicAdd(&sc->counts[3], l_filt);
#pragma unroll
        for (int z = 0; z < 8; ++z)
            if (l_reason[z]) atomicAdd(&sc->counts[8 + z], l_reason[z]);
    }
    __syncthreads();
    if (warp == 0) {
        bc = lane < 8 ? w_cost[lane] : __longlong_as_double(0x7ff0000000000000LL);
        bi = lane < 8 ? w_idx[lane] : 0x7fffffff;
        warp_lexmin(bc, bi);
        if (lane == 0) { sc->part[blockIdx.x].cost = bc; sc->part[blockIdx.x].idx = bi; }
    }
    if (out == nullptr) return;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&sc->counts[1], 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    bc = __longlong_as_double(0x7ff0000000000000LL);
    bi = 0x7fffffff;
    for (int q = tid; q < (int)gridDim.x; q += blockDim.x) {
        const double c = __ldcg(&sc->part[q].cost);
        const int k = __ldcg(&sc->part[q].idx);
        if (lex_less(c, k, bc, bi)) { bc = c; bi = k; }
    }
    warp_lexmin(bc, bi);
    __syncthreads();
    if (lane == 0) { w_cost[warp] = bc; w_idx[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
        bc = lane < 8 ? w_cost[lane] : __longlong_as_double(0x7ff0000000000000LL);
        bi = lane < 8 ? w_idx[lane] : 0x7fffffff;
        warp_lexmin(bc, bi);
        if (lane == 0) {
            const bool has_winner = bi != 0x7fffffff;
            rp_plan_result& r = out->r;
            const int n_feas = __ldcg(&sc->counts[0]), n_filt = __ldcg(&sc->counts[3]);
            r.winner = has_winner ? bi : -1;
            r.winner_cost = has_winner ? bc : __longlong_as_double(0x7ff8000000000000LL);
            r.n_candidates = count;
            r.n_feasible = n_feas;
            r.n_infeasible_kinematics = count - n_filt - n_feas;
            r.n_infeasible_collision = 0;              // filled by count_before_result_kernel
            r.n_collision_total = __ldcg(&sc->counts[2]);
            for (int z = 0; z < 8; ++z) r.reason_counts[z] = __ldcg(&sc->counts[8 + z]);
            out->n_filtered = n_filt;
        }
    }
}

struct PeerTable;
__device__ __forceinline__ void peer_merge_warp(const PeerTable& T, unsigned long long epoch, PlanResultDev* res);

// peer != nullptr (sharded bundle with an open peer group): the block's first warp goes straight on to the exchange of
// the shard records (peer_merge_warp) -- one launch fewer per cycle than a separate exchange kernel
__global__ void __launch_bounds__(512) argmin_merge_kernel(const ArgminScratch* __restrict__ sc, int n_part, int count,
                                                           PlanResultDev* out, const PeerTable* __restrict__ peer = nullptr,
                                                           unsigned long long epoch = 0ULL) {
    __shared__ double w_cost[16];
    __shared__ int w_idx[16];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double bc = tid < n_part ? sc->part[tid].cost : __longlong_as_double(0x7ff0000000000000LL);
    int bi = tid < n_part ? sc->part[tid].idx : 0x7fffffff;
    warp_lexmin(bc, bi);
    if (lane == 0) { w_cost[warp] = bc; w_idx[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
        bc = lane < 16 ? w_cost[lane] : __longlong_as_double(0x7ff0000000000000LL);
        bi = lane < 16 ? w_idx[lane] : 0x7fffffff;
        warp_lexmin(bc, bi);
        if (lane == 0) {
            const bool has_winner = bi != 0x7fffffff;
            rp_plan_result& r = out->r;
            r.winner = has_winner ? bi : -1;
            r.winner_cost = has_winner ? bc : __longlong_as_double(0x7ff8000000000000LL);
            r.n_candidates = count;
            r.n_feasible = sc->counts[0];
            r.n_infeasible_kinematics = count - sc->counts[3] - sc->counts[0];
            r.n_infeasible_collision = 0;              // filled by count_before_result_kernel
            r.n_collision_total = sc->counts[2];
            for (int z = 0; z < 8; ++z) r.reason_counts[z] = sc->counts[8 + z];
            out->n_filtered = sc->counts[3];
        }
        if (peer != nullptr) {
            __syncwarp();
            peer_merge_warp(*peer, epoch, out);
        }
    }
}

// colliders ranked before the winner (all colliders when there is no winner)
__global__ void __launch_bounds__(256) count_before_result_kernel(const double* __restrict__ cost,
                                                                  const int* __restrict__ info, int first, int count,
                                                                  PlanResultDev* out, Stripe sm) {
    if (out->r.n_collision_total == 0) return;
    const int wi = out->r.winner;
    const double wc = out->r.winner_cost;
    const bool none = wi < 0;
    int local = 0;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < count; q += gridDim.x * blockDim.x) {
        const int k = sm.real(first + q);
        if ((info[k] & 0xFF) == ST_COLLISION) {
            if (none || lex_less(cost[k], k, wc, wi)) ++local;
        }
    }
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&out->r.n_infeasible_collision, local);
}

// (loads go to L2 -- __ldcg: in cycle_kernel the verdicts were written by OTHER blocks of the same launch)
// The whole selection of ONE bundle (or shard [first, first + count)) by ONE block: lexicographic arg-min on
// (cost, enumeration index) over feasible, collision-free candidates, the cycle's counters, and the colliders ranked
// before the winner.  Used per scenario of a batch and for replanning-size bundles (one launch instead of three).
__device__ __forceinline__ void block_select(const double* __restrict__ cost, const int* __restrict__ info, int first,
                                             int count, PlanResultDev* __restrict__ out) {
    __shared__ double w_cost[32];
    __shared__ int w_idx[32];
    __shared__ int s_counts[16];
    __shared__ double s_wc;
    __shared__ int s_wi, s_before;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 16) s_counts[tid] = 0;
    if (tid == 0) s_before = 0;
    __syncthreads();
    double bc = __longlong_as_double(0x7ff0000000000000LL);
    int bi = 0x7fffffff;
    int l_feas = 0, l_colt = 0, l_filt = 0;
    int l_reason[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int q = tid; q < count; q += blockDim.x) {
        const int k = first + q;
        const int w = __ldcg(info + k);
        const int st = w & 0xFF;
        if (st == ST_FEASIBLE) {
            ++l_feas;
            const double cc = __ldcg(cost + k);
            if (lex_less(cc, k, bc, bi)) { bc = cc; bi = k; }
        } else if (st == ST_COLLISION) {
            ++l_feas;
            ++l_colt;
        } else if (st == ST_UNCHECKED) {
            ++l_feas;
        } else if (st == ST_KINEMATIC) {
            const int r = (w >> 8) & 0x7;
#pragma unroll
            for (int z = 0; z < 8; ++z) l_reason[z] += (r == z) ? 1 : 0;
        } else {
            ++l_filt;
        }
    }
    warp_lexmin(bc, bi);
    if (lane == 0) { w_cost[warp] = bc; w_idx[warp] = bi; }
    l_feas = warp_sum(l_feas); l_colt = warp_sum(l_colt); l_filt = warp_sum(l_filt);
#pragma unroll
    for (int z = 0; z < 8; ++z) l_reason[z] = warp_sum(l_reason[z]);
    if (lane == 0) {
        atomicAdd(&s_counts[0], l_feas);
        atomicAdd(&s_counts[2], l_colt);
        atomicAdd(&s_counts[3], l_filt);
#pragma unroll
        for (int z = 0; z < 8; ++z) atomicAdd(&s_counts[8 + z], l_reason[z]);
    }
    __syncthreads();
    if (warp == 0) {
        const int n_warps = (int)(blockDim.x >> 5);             // (blocks of 64 .. 1024 threads)
        bc = lane < n_warps ? w_cost[lane] : __longlong_as_double(0x7ff0000000000000LL);
        bi = lane < n_warps ? w_idx[lane] : 0x7fffffff;
        warp_lexmin(bc, bi);
        if (lane == 0) { s_wc = bc; s_wi = bi; }
    }
    __syncthreads();
    const bool none = s_wi == 0x7fffffff;
    if (s_counts[2] > 0) {                      // colliders ranked before the winner (all of them without a winner)
        const double wc = s_wc;
        const int wi = s_wi;
        int local = 0;
        for (int q = tid; q < count; q += blockDim.x) {
            const int k = first + q;
            if ((__ldcg(info + k) & 0xFF) == ST_COLLISION && (none || lex_less(__ldcg(cost + k), k, wc, wi))) ++local;
        }
        local = warp_sum(local);
        if (lane == 0 && local) atomicAdd(&s_before, local);
        __syncthreads();
    }
    if (tid == 0) {
        rp_plan_result& r = out->r;
        r.winner = none ? -1 : s_wi;
        r.winner_cost = none ? __longlong_as_double(0x7ff8000000000000LL) : s_wc;
        r.n_candidates = count;
        r.n_feasible = s_counts[0];
        r.n_infeasible_kinematics = count - s_counts[3] - s_counts[0];
        r.n_infeasible_collision = s_before;
        r.n_collision_total = s_counts[2];
        for (int z = 0; z < 8; ++z) r.reason_counts[z] = s_counts[8 + z];
        out->n_filtered = s_counts[3];
    }
    __syncthreads();
}

// one block per scenario of a batch
// (1 024 threads: a scenario's few thousand candidates are two passes of a handful of iterations per thread)
__global__ void __launch_bounds__(1024) argmin_batch_kernel(const PlanParams* __restrict__ params, PlanResultDev* __restrict__ results) {
    const PlanParams& P = params[blockIdx.x];
    block_select(P.cost, P.info, 0, P.n_cand, results + blockIdx.x);
}

// Closed-loop batches (run_planner.py:84-107 for many scenarios at once): the state of every scenario's winner at time
// step `step` -- where the next replanning cycle starts -- in ONE launch, one thread per scenario marching 0 .. step with
// the exact form of the per-step code (standstill carry, horizon extension included).
// out[sc][16] = x, y, theta, v, a, kappa, s, s_dot, s_ddot, d, d_dot, d_ddot, valid (1 / 0: no winner), -, -, -
__global__ void __launch_bounds__(64) batch_winner_step_kernel(const PlanParams* __restrict__ params,
                                                               const PlanResultDev* __restrict__ results, int n, int step,
                                                               double* __restrict__ out) {
    const int sc = blockIdx.x * blockDim.x + threadIdx.x;
    if (sc >= n) return;
    const PlanParams& P = params[sc];
    double* o = out + (size_t)sc * 16;
    const int k = results[sc].r.winner;
    for (int q = 0; q < 16; ++q) o[q] = 0.;
    if (k < 0 || P.mode != 0) return;
    const rp_plan_inputs& in = P.in;
    const bool low_vel = in.low_vel_mode != 0;
    const int per_t = P.n_lon * P.n_d;
    const int it = k / per_t, rem = k - it * per_t, il = rem / P.n_d, id = rem - il * P.n_d;
    StepIn I;
    I.cs = P.lon_coef + (size_t)(it * P.n_lon + il) * 6;
    I.cd = P.lat_coef + (size_t)(low_vel ? k : it * P.n_d + id) * 6;
    I.lr = nullptr;
    I.lr_stride = 0;
    int tl = P.traj_len[it];
    tl = tl > P.Np1 ? P.Np1 : tl;
    LimitRcp Y{0., 0., 0.};                                  // (unused by the exact form)
    const int last = step < tl ? step : tl - 1;
    StepOut so{};
    double th = 0., kap = 0.;
    for (int i = 0; i <= last; ++i) {
        I.i = i; I.th_prev = th; I.kap_prev = kap;
        so = poly_step<true, false>(P, P.ref, Y, I);
        th = so.th_gl; kap = so.kappa;
    }
    const double tt = (double)last * in.dt;
    const double t2 = tt * tt, t3 = t2 * tt;
    double cs[6], cd[6];
    for (int q = 0; q < 6; ++q) { cs[q] = I.cs[q]; cd[q] = I.cd[q]; }
    const double sa = poly_acc(cs, tt, t2, t3);
    double da;
    if (!low_vel) {
        da = poly_acc(cd, tt, t2, t3);
    } else {
        const double s1 = so.s - cs[0];
        da = poly_acc(cd, s1, s1 * s1, s1 * s1 * s1);
    }
    double x = so.x, y = so.y, v = so.v, s = so.s, d = so.d;
    if (step >= tl) {
        // horizon extension (trajectories.py:168-197, :302-332), as in cand_march
        double ax = 0., ay = 0.;
        for (int i = tl; i <= step; ++i) {
            const double tau = (double)(i - tl + 1) * in.dt;
            double v_tmp = so.v + tau * so.a;
            v_tmp = v_tmp * (v_tmp >= 0 ? 1.0 : 0.0);
            const double ix = in.dt * v_tmp * so.cn, iy = in.dt * v_tmp * so.sn;
            if (i == tl) { ax = ix; ay = iy; } else { ax += ix; ay += iy; }
            if (i == step) { v = v_tmp; s = so.s + tau * so.sv; d = so.d + tau * so.dv; }
        }
        x = so.x + ax; y = so.y + ay;
    }
    o[0] = x; o[1] = y; o[2] = so.th_gl; o[3] = v; o[4] = so.a; o[5] = so.kappa;
    o[6] = s; o[7] = so.sv; o[8] = sa; o[9] = d; o[10] = so.dv; o[11] = da; o[12] = 1.0;
}

// replanning-size bundle: selection + the winner's 14 x (N + 1) state block gathered from the states the main launch
// wrote for every kept candidate -- one launch instead of partial / merge / count / winner-state re-evaluation
__global__ void __launch_bounds__(256) select_small_kernel(const double* __restrict__ cost, const int* __restrict__ info, int first,
                                                           int count, const double* __restrict__ states_all, int Np1,
                                                           PlanResultDev* __restrict__ out, double* __restrict__ states_one) {
    block_select(cost, info, first, count, out);
    const int winner = out->r.winner;           // written by thread 0 before the closing barrier of block_select
    if (winner < 0) return;
    const double* src = states_all + (size_t)winner * 14 * Np1;
    for (int q = threadIdx.x; q < 14 * Np1; q += blockDim.x) states_one[q] = src[q];
}

// result block of a launch chain (record + the winner's states) -> pinned host memory, then the epoch into a host flag the
// caller spins on: no copy-engine hop and no event between the last kernel and the host (rp_grid_result)
__global__ void __launch_bounds__(256) publish_result_kernel(const int4* __restrict__ src, int4* __restrict__ dst_host, int n16,
                                                             unsigned long long* flag_host, unsigned long long epoch) {
    for (int q = threadIdx.x; q < n16; q += blockDim.x) dst_host[q] = __ldcg(src + q);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        *reinterpret_cast<volatile unsigned long long*>(flag_host) = epoch;
        __threadfence_system();
    }
}

// ---- one replanning cycle in one launch (rp_plan_levels; structs in rp_fused.cuh) ------------------------------
// Result block in MAPPED PINNED HOST memory: the last block of cycle_kernel writes it over PCIe and raises `flag` to the
// cycle's epoch after a system-wide fence; the host spins on the flag -- no device->host copy node, no event.
struct CycleOut {
    unsigned long long flag;
    int chosen;                         // level whose record is final: the lowest with a winner, else the last one
    int n_evaluated;                    // records written: levels 0 .. chosen (reactive_planner.py:618 stops there)
    PlanResultDev res[kMaxLevels];      // winner = enumeration index WITHIN the level
    double states[1];                   // winner's 14 x (N + 1) block follows (offset kCycleStatesOffset)
};
constexpr size_t kCycleStatesOffset = 512;
static_assert(sizeof(CycleOut) <= kCycleStatesOffset + sizeof(double), "CycleOut header grew past its slot");

// (two blocks per SM at most: a cycle launch is latency bound and runs blocks of 64 .. 256 threads in one wave; the
// register room keeps the deferred state values of defer_states out of local memory)
template <int MAXT, class ARGS>
__global__ void __launch_bounds__(MAXT, 2)
cycle_kernel(const __grid_constant__ PlanParams P, const __grid_constant__ ARGS A, PlanResultDev* __restrict__ d_res) {
    extern __shared__ double smem[];
    RP_STAMP(0);
#ifdef RP_CYCLE_TIMING
    if (threadIdx.x == 0 && blockIdx.x < 1024) {
        unsigned smid; asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        g_block_t[3 * blockIdx.x] = rp_globaltimer(); g_block_t[3 * blockIdx.x + 2] = smid;
    }
#endif
    fused_body<MAXT, true, ARGS>(P, &A, smem);
    RP_STAMP(11);
#ifdef RP_CYCLE_TIMING
    if (threadIdx.x == 0 && blockIdx.x < 1024) g_block_t[3 * blockIdx.x + 1] = rp_globaltimer();
#endif
    // ---- the last block to finish selects (a12 / a14, trajectories.py:502-510, reactive_planner.py:616-636, :1065-1136)
    __shared__ unsigned s_ticket;
    __threadfence();
    __syncthreads();
#ifdef RP_CYCLE_TIMING
    if (threadIdx.x == 0) atomicMax((unsigned long long*)&g_stamps[31], (unsigned long long)rp_globaltimer());   // last body end + fence
#endif
    if (threadIdx.x == 0) s_ticket = atomicAdd(A.ticket, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();
#ifdef RP_CYCLE_TIMING
    if (threadIdx.x == 0) { g_stamps[12] = clock64(); g_stamps[28] = rp_globaltimer(); }
#endif
    int chosen = A.n_levels - 1;
    for (int lv = 0; lv < A.n_levels; ++lv) {
        block_select(P.cost, P.info, A.lv[lv].k0, A.lv[lv].count, d_res + lv);          // (ends with a barrier)
        if (d_res[lv].r.winner >= 0) { chosen = lv; break; }
    }
    CycleOut* const out = static_cast<CycleOut*>(A.out_host);
    const int Np1 = P.Np1;
    const int winner = d_res[chosen].r.winner;                                           // global enumeration index
    if (winner >= 0) {
        const double* src = P.states + (size_t)winner * 14 * Np1;
        double* dst = reinterpret_cast<double*>(reinterpret_cast<char*>(out) + kCycleStatesOffset);
        for (int q = threadIdx.x; q < 14 * Np1; q += blockDim.x) dst[q] = __ldcg(src + q);
    }
    if ((int)threadIdx.x <= chosen) {
        PlanResultDev r = d_res[threadIdx.x];
        if (r.r.winner >= 0) r.r.winner -= A.lv[threadIdx.x].k0;
        out->res[threadIdx.x] = r;
    }
    if (threadIdx.x == 32) {
        out->chosen = chosen;
        out->n_evaluated = chosen + 1;
        *A.ticket = 0u;                                                                  // ready for the next launch
        if (P.best_bits)
            for (int lv = 0; lv < kMaxLevels; ++lv) P.best_bits[lv] = 0x7f7f7f7f7f7f7f7fULL;
    }
#ifdef RP_CYCLE_TIMING
    if (threadIdx.x == 0) { g_stamps[13] = clock64(); g_stamps[29] = rp_globaltimer(); }
#endif
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
#ifdef RP_CYCLE_TIMING
        g_stamps[14] = clock64(); g_stamps[30] = rp_globaltimer();
#endif
        *reinterpret_cast<volatile unsigned long long*>(&out->flag) = A.epoch;
        __threadfence_system();
    }
}

// ---- multi-GPU bundle shards: each rank owns a contiguous tile of the enumeration space --------
// record = [local best cost (+inf if none), its enumeration index (as double, +inf if none),
//           kinematically infeasible count, kinematically feasible count]
__global__ void export_record_kernel(const PlanResultDev* __restrict__ res, double* __restrict__ dst) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const bool has = res->r.winner >= 0;
    dst[0] = has ? res->r.winner_cost : inf;
    dst[1] = has ? (double)res->r.winner : inf;
    dst[2] = (double)res->r.n_infeasible_kinematics;
    dst[3] = (double)res->r.n_feasible;
}

// lexicographic min on (cost, enumeration index) over the ranks' gathered records [world][4] and the summed counters:
// winner[2] = [cost, index] (+inf = none), totals[2] = [n_infeasible_kinematics, n_feasible]; one warp
__global__ void merge_records_kernel(const double* __restrict__ gathered, int world, double* __restrict__ winner,
                                     double* __restrict__ totals) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double bc = inf, bi = inf, kin = 0., feas = 0.;
    for (int r = threadIdx.x; r < world; r += 32) {
        const double c = gathered[4 * r], i = gathered[4 * r + 1];
        if (c < bc || (c == bc && i < bi)) { bc = c; bi = i; }
        kin += gathered[4 * r + 2];
        feas += gathered[4 * r + 3];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double oc = __shfl_down_sync(0xffffffffu, bc, off), oi = __shfl_down_sync(0xffffffffu, bi, off);
        if (oc < bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
        kin += __shfl_down_sync(0xffffffffu, kin, off);
        feas += __shfl_down_sync(0xffffffffu, feas, off);
    }
    if (threadIdx.x == 0) {
        winner[0] = bc; winner[1] = bi;
        totals[0] = kin; totals[1] = feas;
    }
}

// colliders of this shard ranked before the GLOBAL winner (lazy collision count, App. B#12)
__global__ void __launch_bounds__(256) count_before_kernel(const double* __restrict__ cost, const int* __restrict__ info,
                                                           int first, int count, const double* __restrict__ winner,
                                                           double* __restrict__ out, Stripe sm) {
    __shared__ int total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    const double wc = winner[0];
    const double wi = winner[1];
    const bool none = !(wi < __longlong_as_double(0x7ff0000000000000LL));
    int local = 0;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < count; q += gridDim.x * blockDim.x) {
        const int k = sm.real(first + q);
        if ((info[k] & 0xFF) == ST_COLLISION) {
            const double c = cost[k];
            if (none || c < wc || (c == wc && (double)k < wi)) ++local;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) local += __shfl_down_sync(0xffffffffu, local, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(&total, local);
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(out, (double)total);
}

// ------------------------------------------------------------------------------------------------
// reference polyline -> the nine device tables, on the device (one block per path).
//   utils_coordinate_system.py:113-118 (ref_pos / ref_curv / ref_theta / ref_curv_d from the frame's stored polyline)
//   and the frame construction of pycrccosy as restated in oracle/third_party.py:126-145 (two extension vertices eps2
//   beyond the ends, per-vertex pseudo-normals from the normalised chords p[i+1] - p[i-1]);
//   commonroad_dc.geometry.util compute_pathlength / orientation / curvature_from_polyline (third_party.py:44-70);
//   np.gradient with non-uniform spacing (second-order interior, first-order edges), np.cumsum and np.unwrap in
//   numpy's operation order (the two scans are sequential in one thread; everything else is one thread per vertex).
// in:  xy[n_in][2] (smoothed, de-duplicated reference), out: 9 SoA arrays of n = n_in + 2 doubles
//      [pos | theta | curv | curv_d | px | py | nx | ny | ps], scratch: 4 arrays of n doubles.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double np_gradient_at(const double* __restrict__ f, const double* __restrict__ x, int n, int i,
                                                 bool uniform) {
    if (i == 0) return (f[1] - f[0]) / (x[1] - x[0]);
    if (i == n - 1) return (f[n - 1] - f[n - 2]) / (x[n - 1] - x[n - 2]);
    const double dx1 = x[i] - x[i - 1], dx2 = x[i + 1] - x[i];
    if (uniform) return (f[i + 1] - f[i - 1]) / (2. * dx1);
    const double a = -(dx2) / (dx1 * (dx1 + dx2));
    const double b = (dx2 - dx1) / (dx1 * dx2);
    const double c = dx1 / (dx2 * (dx1 + dx2));
    return a * f[i - 1] + b * f[i] + c * f[i + 1];
}

__global__ void __launch_bounds__(256) ref_tables_kernel(int n_in, const double* __restrict__ xy, double eps2,
                                                         double* __restrict__ out, double* __restrict__ scratch) {
    const int n = n_in + 2;
    double* pos = out;
    double* theta = out + n;
    double* curv = out + 2 * (size_t)n;
    double* curv_d = out + 3 * (size_t)n;
    double* px = out + 4 * (size_t)n;
    double* py = out + 5 * (size_t)n;
    double* nx = out + 6 * (size_t)n;
    double* ny = out + 7 * (size_t)n;
    double* ps = out + 8 * (size_t)n;
    double* xd = scratch;
    double* yd = scratch + n;
    double* seg = scratch + 2 * (size_t)n;
    double* th_raw = scratch + 3 * (size_t)n;
    __shared__ int s_uniform;
    const int tid = threadIdx.x, T = blockDim.x;
    // ---- the frame's stored polyline: the input plus one vertex eps2 beyond each end --------------------------
    for (int i = tid; i < n; i += T) {
        double x, y;
        if (i == 0) {
            const double hx = xy[2] - xy[0], hy = xy[3] - xy[1];
            const double nrm = sqrt(hx * hx + hy * hy);
            x = xy[0] - eps2 * hx / nrm; y = xy[1] - eps2 * hy / nrm;
        } else if (i == n - 1) {
            const double* e = xy + 2 * (size_t)(n_in - 1);
            const double tx = e[0] - e[-2], ty = e[1] - e[-1];
            const double nrm = sqrt(tx * tx + ty * ty);
            x = e[0] + eps2 * tx / nrm; y = e[1] + eps2 * ty / nrm;
        } else {
            x = xy[2 * (size_t)(i - 1)]; y = xy[2 * (size_t)(i - 1) + 1];
        }
        px[i] = x; py[i] = y;
    }
    __syncthreads();
    // ---- segment lengths, raw orientations, pseudo-normals ----------------------------------------------------
    for (int i = tid; i < n; i += T) {
        if (i < n - 1) {
            const double dx = px[i + 1] - px[i], dy = py[i + 1] - py[i];
            seg[i] = sqrt(dx * dx + dy * dy);
            th_raw[i] = atan2(dy, dx);
        }
        const int lo = i == 0 ? 0 : i - 1, hi = i == n - 1 ? n - 1 : i + 1;
        const double cx = px[hi] - px[lo], cy = py[hi] - py[lo];
        const double nrm = sqrt(cx * cx + cy * cy);
        nx[i] = -(cy / nrm);
        ny[i] = cx / nrm;
    }
    __syncthreads();
    // ---- np.cumsum / np.unwrap: sequential, numpy's order -------------------------------------------------------
    if (tid == 0) {
        double acc = 0.;
        pos[0] = 0.;
        for (int i = 0; i < n - 1; ++i) {
            acc = i == 0 ? seg[0] : acc + seg[i];
            pos[i + 1] = acc;
        }
        int uni = 1;
        const double d0 = pos[1] - pos[0];
        for (int i = 1; i < n - 1 && uni; ++i) uni = (pos[i + 1] - pos[i]) == d0;
        s_uniform = uni;
    }
    if (tid == 32) {
        th_raw[n - 1] = th_raw[n - 2];                      // compute_orientation_from_polyline repeats the last value
        const double pi = 3.141592653589793, two_pi = 6.283185307179586;
        theta[0] = th_raw[0];
        double corr = 0.;
        bool first = true;
        for (int i = 1; i < n; ++i) {
            const double dd = th_raw[i] - th_raw[i - 1];
            // np.mod(dd + pi, 2 pi) - pi with Python's sign convention (result has the sign of the divisor)
            double m = fmod(dd + pi, two_pi);
            if (m != 0. && m < 0.) m += two_pi;
            double ddmod = m - pi;
            if (ddmod == -pi && dd > 0.) ddmod = pi;
            double ph = ddmod - dd;
            if (fabs(dd) < pi) ph = 0.;
            corr = first ? ph : corr + ph;
            first = false;
            theta[i] = th_raw[i] + corr;
        }
    }
    __syncthreads();
    const bool uniform = s_uniform != 0;
    for (int i = tid; i < n; i += T) {
        ps[i] = pos[i];
        xd[i] = np_gradient_at(px, pos, n, i, uniform);
        yd[i] = np_gradient_at(py, pos, n, i, uniform);
    }
    __syncthreads();
    for (int i = tid; i < n; i += T) {
        const double xdd = np_gradient_at(xd, pos, n, i, uniform);
        const double ydd = np_gradient_at(yd, pos, n, i, uniform);
        curv[i] = (xd[i] * ydd - xdd * yd[i]) / pow(xd[i] * xd[i] + yd[i] * yd[i], 1.5);
    }
    __syncthreads();
    for (int i = tid; i < n; i += T) curv_d[i] = np_gradient_at(curv, pos, n, i, uniform);
}

// ---- the same exchange over peer-mapped memory (NVLink / NVSwitch, CUDA IPC) instead of two NCCL collectives ------
// Every rank owns one PeerMailbox that all ranks of the box have mapped.  When a peer group is open and the bundle is
// sharded, the selection chain of rp_grid_launch is argmin_partial -> argmin_merge (the shard's result) ->
//   peer_merge_kernel (one warp): lane r stores this shard's record (best cost, best index, the counters) into rank
//     r's mailbox, fences, then raises its flag there; the warp waits until every rank's flag in its OWN mailbox shows
//     the cycle's epoch and merges the records -- lexicographic min on (cost, index), summed counters -- INTO the
//     result block: every rank ends up with the global result, and the winner-state launch that follows materialises
//     the GLOBAL winner on every rank (each rank holds all coefficient systems and tables).
//   peer_count_kernel: the shard's colliders ranked before the global winner; the last block to finish exchanges the
//     shard counts the same way and sums them (integers in doubles: exact) into n_infeasible_collision.
// Slots are double-buffered by epoch parity: a rank can be at most one phase ahead of the slowest one, because it
// cannot leave a phase before all ranks have written that phase's flags.  Waits are bounded (kPeerSpinCycles); a
// timeout sets PlanResultDev::peer_error, which rp_grid_result reports.
constexpr int kMaxPeers = 16;
constexpr int kPeerRecord = 16;          // cost, index, n_kin, n_feasible, n_collision_total, n_candidates, n_filtered, reasons[8]
constexpr long long kPeerSpinCycles = 16000000000LL;            // ~8 s at 1.9 GHz
struct PeerMailbox {
    double rec[2][kMaxPeers][kPeerRecord];
    unsigned long long flag1[2][kMaxPeers];
    double cnt[2][kMaxPeers];
    unsigned long long flag2[2][kMaxPeers];
    int local_count;                    // this rank only: colliders before the winner, accumulated over the blocks
    unsigned int ticket;                // this rank only: blocks of peer_count_kernel that have finished
};
struct PeerTable {
    PeerMailbox* box[kMaxPeers];
    int rank, world;
};

__device__ __forceinline__ bool peer_wait(const unsigned long long* flag, unsigned long long epoch) {
    const volatile unsigned long long* f = flag;
    const long long t0 = clock64();
    while (*f != epoch) {
        if (clock64() - t0 > kPeerSpinCycles) return false;
        __nanosleep(64);
    }
    __threadfence_system();
    return true;
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    return v;
}

__device__ __forceinline__ void peer_merge_warp(const PeerTable& T, unsigned long long epoch, PlanResultDev* res) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const int lane = threadIdx.x & 31, par = (int)(epoch & 1ULL);
    PeerMailbox* const mine = T.box[T.rank];
    if (lane == 0) { mine->local_count = 0; mine->ticket = 0u; res->peer_error = 0; }
    bool ok = true;
    if (lane < T.world) {
        const rp_plan_result r0 = res->r;
        const bool has = r0.winner >= 0;
        PeerMailbox* const dst = T.box[lane];
        double* r = dst->rec[par][T.rank];
        r[0] = has ? r0.winner_cost : inf;
        r[1] = has ? (double)r0.winner : inf;
        r[2] = (double)r0.n_infeasible_kinematics;
        r[3] = (double)r0.n_feasible;
        r[4] = (double)r0.n_collision_total;
        r[5] = (double)r0.n_candidates;
        r[6] = (double)res->n_filtered;
#pragma unroll
        for (int z = 0; z < 8; ++z) r[7 + z] = (double)r0.reason_counts[z];
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>(&dst->flag1[par][T.rank]) = epoch;
        ok = peer_wait(&mine->flag1[par][lane], epoch);
    }
    if (!__all_sync(0xffffffffu, ok)) {
        if (lane == 0) { res->peer_error = 1; res->r.winner = -1; }
        // publish the second phase with a poison count: the peers fail this cycle at once instead of timing out again
        if (lane < T.world) {
            PeerMailbox* const dst = T.box[lane];
            dst->cnt[par][T.rank] = -1.0;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long*>(&dst->flag2[par][T.rank]) = epoch;
        }
        return;
    }
    double bc = inf, bi = inf, sums[kPeerRecord - 2];
#pragma unroll
    for (int z = 0; z < kPeerRecord - 2; ++z) sums[z] = 0.;
    if (lane < T.world) {
        const volatile double* g = mine->rec[par][lane];
        bc = g[0]; bi = g[1];
#pragma unroll
        for (int z = 0; z < kPeerRecord - 3; ++z) sums[z] = g[2 + z];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double oc = __shfl_down_sync(0xffffffffu, bc, off), oi = __shfl_down_sync(0xffffffffu, bi, off);
        if (oc < bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
    }
#pragma unroll
    for (int z = 0; z < kPeerRecord - 3; ++z) sums[z] = warp_sum_f64(sums[z]);
    if (lane == 0) {
        const bool has_winner = bi < inf;
        rp_plan_result& r = res->r;
        r.winner = has_winner ? (int)bi : -1;
        r.winner_cost = has_winner ? bc : __longlong_as_double(0x7ff8000000000000LL);
        r.n_infeasible_kinematics = (int)sums[0];
        r.n_feasible = (int)sums[1];
        r.n_collision_total = (int)sums[2];
        r.n_candidates = (int)sums[3];
        res->n_filtered = (int)sums[4];
#pragma unroll
        for (int z = 0; z < 8; ++z) r.reason_counts[z] = (int)sums[5 + z];
        r.n_infeasible_collision = 0;                  // filled by peer_count_kernel
    }
}

__global__ void __launch_bounds__(32) peer_merge_kernel(const __grid_constant__ PeerTable T, unsigned long long epoch,
                                                        PlanResultDev* __restrict__ res) {
    peer_merge_warp(T, epoch, res);
}

__global__ void __launch_bounds__(256) peer_count_kernel(const __grid_constant__ PeerTable T, unsigned long long epoch,
                                                         const double* __restrict__ cost, const int* __restrict__ info,
                                                         int first, int count, PlanResultDev* res, Stripe sm) {
    __shared__ int total;
    __shared__ unsigned int s_ticket;
    PeerMailbox* const mine = T.box[T.rank];
    if (res->peer_error) return;                     // (uniform over the grid: set before this launch started)
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    const int wi = res->r.winner;
    const double wc = res->r.winner_cost;
    const bool none = wi < 0;
    int local = 0;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < count; q += gridDim.x * blockDim.x) {
        const int k = sm.real(first + q);
        if ((info[k] & 0xFF) == ST_COLLISION) {
            if (none || lex_less(cost[k], k, wc, wi)) ++local;
        }
    }
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&total, local);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (total) atomicAdd(&mine->local_count, total);
        __threadfence();
        s_ticket = atomicAdd(&mine->ticket, 1u);
    }
    __syncthreads();
    if (s_ticket != gridDim.x - 1 || threadIdx.x >= 32) return;
    // ---- last block, first warp: exchange the shard counts -----------------------------------------------------
    const int lane = threadIdx.x, par = (int)(epoch & 1ULL);
    bool ok = true;
    if (lane < T.world) {
        const double mine_cnt = (double)*reinterpret_cast<volatile int*>(&mine->local_count);
        PeerMailbox* const dst = T.box[lane];
        dst->cnt[par][T.rank] = mine_cnt;
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>(&dst->flag2[par][T.rank]) = epoch;
        ok = peer_wait(&mine->flag2[par][lane], epoch);
    }
    if (!__all_sync(0xffffffffu, ok)) {
        if (lane == 0) res->peer_error = 1;
        return;
    }
    double sum = lane < T.world ? *reinterpret_cast<const volatile double*>(&mine->cnt[par][lane]) : 0.;
    if (__any_sync(0xffffffffu, sum < 0.)) {          // a peer's first phase failed (poison count)
        if (lane == 0) { res->peer_error = 1; res->r.winner = -1; }
        return;
    }
    sum = warp_sum_f64(sum);
    if (lane == 0) res->r.n_infeasible_collision = (int)sum;
}

// ------------------------------------------------------------------------------------------------
// Cartesian initial state -> curvilinear (lon, lat) initial states, batched.
//   pycrccosy convert_to_curvilinear_coords as restated in oracle/third_party.py (:188-221): per segment the foot
//   parameter of the pseudo-normal map solves a quadratic; the solution with the smallest |d| wins (first in
//   (segment, root) order on ties); then reactive_planner.py:446-512 in the reference's operation order.
// One block per state: threads scan the segments, a lexicographic (|d|, 2 j + root) block minimum picks the foot.
// status: 0 ok, 1 outside the projection domain (the reference raises ValueError), 2 negative s_dot (Exception).
// ------------------------------------------------------------------------------------------------
struct InitFrame {
    RefTables ref;
    double wheelbase;
};

__global__ void __launch_bounds__(128) initial_states_kernel(int n_states, const double* __restrict__ x0,
                                                             const int* __restrict__ low_vel,
                                                             const InitFrame* __restrict__ frames, int frame_stride,
                                                             double* __restrict__ out_lon, double* __restrict__ out_lat,
                                                             int* __restrict__ status) {
    __shared__ double w_key[4], w_s[4], w_d[4];
    __shared__ int w_ord[4];
    const int st = blockIdx.x;
    if (st >= n_states) return;
    const InitFrame& F = frames[(size_t)st * frame_stride];
    const RefTables& R = F.ref;
    const double px = x0[6 * st], py = x0[6 * st + 1];
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double best_key = inf, best_s = 0., best_d = 0.;
    int best_ord = 0x7fffffff;
    for (int j = threadIdx.x; j < R.n - 1; j += blockDim.x) {
        const double p0x = R.px[j], p0y = R.py[j];
        const double ex = R.px[j + 1] - p0x, ey = R.py[j + 1] - p0y;
        const double mx = R.nx[j], my = R.ny[j];
        const double dnx = R.nx[j + 1] - mx, dny = R.ny[j + 1] - my;
        const double ax = px - p0x, ay = py - p0y;
        const double qa = -(ex * dny - ey * dnx);
        const double qb = (ax * dny - ay * dnx) - (ex * my - ey * mx);
        const double qc = ax * my - ay * mx;
        double roots[2];
        int n_roots = 0;
        if (fabs(qa) < 1e-14) {
            if (fabs(qb) > 0.0) roots[n_roots++] = -qc / qb;
        } else {
            const double disc = qb * qb - 4.0 * qa * qc;
            if (disc >= 0.0) {
                // cancellation-free form of (-qb +- sqrt(disc)) / (2 qa): qa is tiny on gently curved paths
                const double sq = sqrt(disc);
                const double qq = -0.5 * (qb + copysign(sq, qb));
                if (qq == 0.0) { roots[0] = 0.0; roots[1] = 0.0; }
                else if (qb >= 0.0) { roots[0] = qc / qq; roots[1] = qq / qa; }
                else { roots[0] = qq / qa; roots[1] = qc / qq; }
                n_roots = 2;
            }
        }
        for (int r = 0; r < n_roots; ++r) {
            double lam = roots[r];
            if (!(-1e-12 <= lam && lam <= 1.0 + 1e-12)) continue;
            lam = fmin(fmax(lam, 0.0), 1.0);
            const double bx = p0x + lam * ex, by = p0y + lam * ey;
            const double pnx = mx + lam * dnx, pny = my + lam * dny;
            const double dd = ((px - bx) * pnx + (py - by) * pny) / (pnx * pnx + pny * pny);
            const int ord = 2 * j + r;
            if (fabs(dd) <= R.limit && (fabs(dd) < best_key || (fabs(dd) == best_key && ord < best_ord))) {
                best_key = fabs(dd);
                best_ord = ord;
                best_d = dd;
                best_s = R.ps[j] + lam * (R.ps[j + 1] - R.ps[j]);
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ok = __shfl_down_sync(0xffffffffu, best_key, off);
        const int oo = __shfl_down_sync(0xffffffffu, best_ord, off);
        const double os = __shfl_down_sync(0xffffffffu, best_s, off);
        const double od = __shfl_down_sync(0xffffffffu, best_d, off);
        if (ok < best_key || (ok == best_key && oo < best_ord)) { best_key = ok; best_ord = oo; best_s = os; best_d = od; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { w_key[warp] = best_key; w_ord[warp] = best_ord; w_s[warp] = best_s; w_d[warp] = best_d; }
    __syncthreads();
    if (threadIdx.x != 0) return;
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
        if (w_key[w] < best_key || (w_key[w] == best_key && w_ord[w] < best_ord)) {
            best_key = w_key[w]; best_ord = w_ord[w]; best_s = w_s[w]; best_d = w_d[w];
        }
    double* lon = out_lon + 3 * st;
    double* lat = out_lat + 3 * st;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (best_ord == 0x7fffffff) {
        status[st] = 1;
        lon[0] = lon[1] = lon[2] = lat[0] = lat[1] = lat[2] = nan;
        return;
    }
    // ---- reactive_planner.py:465-512 ----------------------------------------------------------------
    const double s = best_s, d = best_d;
    const double orientation = x0[6 * st + 2], velocity = x0[6 * st + 3], acceleration = x0[6 * st + 4];
    const double steering = x0[6 * st + 5];
    const int ub = upper_bound(R.pos, R.n, s);               // np.argmax(ref_pos > s) (0 if none)
    const int ub_np = ub == R.n ? 0 : ub;
    const int i0 = ub_np == 0 ? R.n - 1 : ub_np - 1;          // s_idx = argmax - 1, python index wrap for -1
    const int i1 = ub_np == 0 ? 0 : ub_np;                    // s_idx + 1
    const double s_lambda = (s - R.pos[i0]) / (R.pos[i1] - R.pos[i0]);
    const double theta_cl = orientation - interpolate_angle(s, R.pos[i0], R.pos[i1], R.theta[i0], R.theta[i1]);
    const double kr = (R.curv[i1] - R.curv[i0]) * s_lambda + R.curv[i0];
    const double kr_d = (R.curv_d[i1] - R.curv_d[i0]) * s_lambda + R.curv_d[i0];
    const double kappa_0 = tan(steering) / F.wheelbase;
    const double one_krd = 1 - kr * d;
    const double tan_t = tan(theta_cl), cos_t = cos(theta_cl);
    const double d_p = one_krd * tan_t;
    const double d_pp = -(kr_d * d + kr * d_p) * tan_t + (one_krd / (cos_t * cos_t)) * (kappa_0 * one_krd / cos_t - kr);
    const double s_velocity = velocity * cos_t / one_krd;
    double s_acceleration = acceleration;
    s_acceleration -= (s_velocity * s_velocity / cos_t) * (one_krd * tan_t * (kappa_0 * one_krd / cos_t - kr) - (kr_d * d + kr * d_p));
    s_acceleration /= (one_krd / cos_t);
    double d_velocity, d_acceleration;
    if (low_vel[st]) {
        d_velocity = d_p;
        d_acceleration = d_pp;
    } else {
        d_velocity = velocity * sin(theta_cl);
        d_acceleration = s_acceleration * d_p + s_velocity * s_velocity * d_pp;
    }
    lon[0] = s; lon[1] = s_velocity; lon[2] = s_acceleration;
    lat[0] = d; lat[1] = d_velocity; lat[2] = d_acceleration;
    status[st] = s_velocity < 0 ? 2 : 0;
}

// ------------------------------------------------------------------------------------------------
// the continuous collision check of the selected candidate (reactive_planner.py:1049-1058).
// Hull k = OBB-sum of the ego boxes at steps k, k + 1 (oracle/third_party.py obb_sum_hull: tight box along the
// first box's axes), tested at time index x0.time_step + k against every dynamic box of that index and every static
// primitive (brute force: one trajectory, <= N hulls; the broad-phase grid is sized for the vehicle box, not hulls).
// A hit relabels the winner, counts it and leaves the cycle without a trajectory -- the reference breaks out of its
// candidate loop there.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) continuous_check_kernel(ObstacleTables O, const double* __restrict__ st, int Np1,
                                                               int x0_time_step, double half_len, double half_wid,
                                                               double wb_rear, PlanResultDev* res, int* __restrict__ info) {
    __shared__ int s_hit;
    const int winner = res->r.winner;
    if (winner < 0) return;
    if (threadIdx.x == 0) s_hit = 0;
    __syncthreads();
    const double* X = st;
    const double* Y = st + Np1;
    const double* TH = st + 2 * Np1;
    bool hit = false;
    for (int i = threadIdx.x; i < Np1 - 1 && !hit; i += blockDim.x) {
        double sa, ca, sb, cb;
        sincos(TH[i], &sa, &ca);
        sincos(TH[i + 1], &sb, &cb);
        const double acx = X[i] + wb_rear * ca, acy = Y[i] + wb_rear * sa;
        const double bcx = X[i + 1] + wb_rear * cb, bcy = Y[i + 1] + wb_rear * sb;
        const double dx = bcx - acx, dy = bcy - acy;
        const double u0 = dx * ca + dy * sa;
        const double v0 = -dx * sa + dy * ca;
        const double c = ca * cb + sa * sb;
        const double s = ca * sb - sa * cb;
        const double eu = half_len * fabs(c) + half_wid * fabs(s);
        const double ev = half_len * fabs(s) + half_wid * fabs(c);
        const double umin = fmin(-half_len, u0 - eu), umax = fmax(half_len, u0 + eu);
        const double vmin = fmin(-half_wid, v0 - ev), vmax = fmax(half_wid, v0 + ev);
        const double um = 0.5 * (umin + umax), vm = 0.5 * (vmin + vmax);
        const double hcx = acx + um * ca - vm * sa, hcy = acy + um * sa + vm * ca;
        const double hl = 0.5 * (umax - umin), hw = 0.5 * (vmax - vmin);
        const int tidx = x0_time_step + i;
        for (int o = 0; o < O.n_dyn && !hit; ++o) {
            const int k = tidx - O.dyn_t0[o];
            if (k < 0 || k >= O.dyn_len[o]) continue;
            hit = obb_obb_overlap(hcx, hcy, ca, sa, hl, hw, O.dyn_box + (size_t)(O.dyn_off[o] + k) * kBoxStride);
        }
        for (int q = 0; q < O.n_obb && !hit; ++q) hit = obb_obb_overlap(hcx, hcy, ca, sa, hl, hw, O.obb + (size_t)q * kBoxStride);
        for (int q = 0; q < O.n_tri && !hit; ++q) hit = obb_triangle_overlap(hcx, hcy, ca, sa, hl, hw, O.tri + (size_t)q * 6);
    }
    if (hit) atomicOr(&s_hit, 1);
    __syncthreads();
    if (threadIdx.x == 0 && s_hit) {
        info[winner] = pack_info(ST_COLLISION, R_NONE, -1);
        res->r.winner = -1;
        res->r.winner_cost = __longlong_as_double(0x7ff8000000000000LL);
        res->r.n_infeasible_collision += 1;
        res->r.n_collision_total += 1;
    }
}

// pycrcc.CollisionChecker.collide for a batch of ego boxes (rp_collide_poses)
__global__ void collide_kernel(int n, const double* __restrict__ pose, const int* __restrict__ tidx, double hl,
                               double hw, double r_ego, ObstacleTables O, int vehicle_box, uint8_t* __restrict__ hit) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    double st, ct;
    sincos(pose[3 * g + 2], &st, &ct);
    const double cx = pose[3 * g], cy = pose[3 * g + 1];
    hit[g] = (dyn_collides_global(O, tidx[g], cx, cy, ct, st, hl, hw, r_ego) || static_collides(O, cx, cy, ct, st, hl, hw, vehicle_box != 0)) ? 1 : 0;
}

}  // namespace rp
