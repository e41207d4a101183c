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

namespace rp {

struct PlanParams {
    // ---- candidate source ----
    int mode;                    // 0: grid (t x lon x d), 1: list (per-candidate coefficients)
    int n_t, n_lon, n_d;
    int n_cand;                  // size of the enumeration space
    int first, count;            // slots processed by this launch: candidate = first + slot ...
    const int* index;            // ... or index[slot] when non-null (negative = empty slot)
    const double* lon_samples;   // grid: [n_lon] (filter_goals_behind in stopping mode)
    const int* traj_len;         // grid: [n_t], list: [n_cand]
    const double* lon_coef;      // grid: [n_t*n_lon][6], list: [n_cand][6]
    const double* lat_coef;      // grid: [n_t*n_d][6] (or [n_cand][6] in low-velocity mode), list: [n_cand][6]
    const uint8_t* skip;         // list: filter_goals_behind flags (may be null)
    // ---- per-cycle scalars / tables ----
    rp_plan_inputs in;
    Limits lim;
    double half_len, half_wid, wb_rear, r_ego;
    RefTables ref;
    ObstacleTables obs;
    // ---- outputs ----
    double* cost;                // [n_cand]   (null: not written)
    int* info;                   // [n_cand]   status | reason << 8 | (step + 1) << 16
    double* states;              // [.][14][N+1] indexed by candidate (states_by_slot = 0) or slot
    int states_by_slot;
    // ---- launch geometry ----
    int C;                       // candidates per block
    int Np1;                     // N + 1
    int n_groups;
    int stage_ref, stage_dyn;
};

struct PlanResultDev {
    rp_plan_result r;
    int n_filtered;
};

__device__ __forceinline__ int pack_info(int status, int reason, int step) {
    return status | (reason << 8) | ((step + 1) << 16);
}

// ------------------------------------------------------------------------------------------------
// a3: one thread per polynomial.  Threads [0, n_lon_sys) solve the longitudinal systems, the rest
// the lateral ones.  In low-velocity mode the lateral parameter range is the longitudinal end
// position (sampling.py:229-234), so a lateral thread first redoes its longitudinal solve.
// ------------------------------------------------------------------------------------------------
__global__ void coeff_kernel(int n_t, int n_lon, int n_d, int low_vel, int lon_mode, const double* __restrict__ t,
                             const double* __restrict__ lon, const double* __restrict__ d, const double x0s,
                             const double x0sd, const double x0sdd, const double x0d, const double x0dd,
                             const double x0ddd, double* __restrict__ lon_coef, double* __restrict__ lat_coef,
                             double* __restrict__ lat_tau) {
    const int n_lon_sys = n_t * n_lon;
    const int n_lat_sys = low_vel ? n_t * n_lon * n_d : n_t * n_d;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
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
    lat_tau[q] = tau;
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
// a4-a13 fused.  Block = C candidates x (N+1) time steps (flat: thread -> (slot c, step i)); the
// block is persistent over candidate groups so that the reference-path table and the per-step
// dynamic-obstacle rows are staged into shared memory once.
// ------------------------------------------------------------------------------------------------
#ifndef RP_FUSED_MIN_BLOCKS
#define RP_FUSED_MIN_BLOCKS 2
#endif
template <int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT == 256 ? RP_FUSED_MIN_BLOCKS : 1) fused_kernel(const __grid_constant__ PlanParams P) {
    extern __shared__ double smem[];
    const int Np1 = P.Np1;
    const int C = P.C;
    const int tid = threadIdx.x;
    const int c = tid / Np1;
    const int i = tid - c * Np1;
    const bool lane = c < C;

    // ---- shared memory carve-up -----------------------------------------------------------------
    double* sp = smem;
    RefTables R = P.ref;
    if (P.stage_ref) {
        const int n = R.n;
        double* base = sp;
        const double* src[9] = {R.pos, R.theta, R.curv, R.curv_d, R.px, R.py, R.nx, R.ny, R.ps};
        const int n_arr = R.same_s ? 8 : 9;
        for (int a = 0; a < n_arr; ++a)
            for (int q = tid; q < n; q += blockDim.x) base[a * n + q] = src[a][q];
        R.pos = base; R.theta = base + n; R.curv = base + 2 * n; R.curv_d = base + 3 * n;
        R.px = base + 4 * n; R.py = base + 5 * n; R.nx = base + 6 * n; R.ny = base + 7 * n;
        R.ps = R.same_s ? R.pos : base + 8 * n;
        sp += n_arr * n;
    }
    const ObstacleTables& O = P.obs;
    const double* dyn_stage = nullptr;
    if (P.stage_dyn && O.n_dyn > 0) {
        // rows of the dynamic obstacles present at time index x0.time_step + i * factor (reactive_planner.py:1040)
        double* dst = sp;
        const int total = Np1 * O.n_dyn;
        for (int q = tid; q < total; q += blockDim.x) {
            int step = q / O.n_dyn, o = q - step * O.n_dyn;
            int k = P.in.x0_time_step + step * P.in.factor - O.dyn_t0[o];
            bool present = k >= 0 && k < O.dyn_len[o];
            const double* b = O.dyn_box + (size_t)(O.dyn_off[o] + (present ? k : 0)) * kBoxStride;
#pragma unroll
            for (int w = 0; w < kBoxStride; ++w) dst[(size_t)q * kBoxStride + w] = present ? b[w] : 0.0;
        }
        dyn_stage = dst;
        sp += (size_t)total * kBoxStride;
    }
    // per-slot scratch
    const int per_slot = 9 * Np1 + 14 + 40 + 6;
    double* scratch = sp + (size_t)(lane ? c : 0) * per_slot;
    double* s_th = scratch;                 // theta_gl per step
    double* s_kap = scratch + Np1;          // kappa_gl per step
    double* s_tx = scratch + 2 * Np1;       // extension increments (x)
    double* s_ty = scratch + 3 * Np1;       // extension increments (y)
    double* s_ct = scratch + 4 * Np1;       // 5 cost term rows
    double* s_last = scratch + 9 * Np1;     // the 14 state values at the last polynomial step
    double* s_acc = s_last + 14;            // 5 x 8 pairwise accumulators
    double* s_sum = s_acc + 40;             // 5 sums (+1 pad)
    sp += (size_t)C * per_slot;
    int* ip = reinterpret_cast<int*>(sp);
    int* s_carry = ip + (size_t)(lane ? c : 0) * Np1;    // standstill carry flags per step
    unsigned* s_flags = reinterpret_cast<unsigned*>(ip + (size_t)C * Np1) + (size_t)(lane ? c : 0) * 4;
    // s_flags[0] first kinematic violation ((step << 8) | reason), [1] first projection failure step,
    // [2] pre-filter bits, [3] first colliding step

    const rp_plan_inputs& in = P.in;
    const bool low_vel = in.low_vel_mode != 0;
    const bool draw = in.draw_all != 0;
    const double dt = in.dt;
    const unsigned NONE = 0xFFFFFFFFu;
    __syncthreads();

    for (int g = blockIdx.x; g < P.n_groups; g += gridDim.x) {
        const int slot = g * C + c;
        bool valid = lane && slot < P.count;
        int k = -1;
        if (valid) {
            k = P.index ? P.index[slot] : P.first + slot;
            if (k < 0 || k >= P.n_cand) valid = false;
        }
        if (lane && i == 0) { s_flags[0] = NONE; s_flags[1] = NONE; s_flags[2] = 0u; s_flags[3] = NONE; }

        double cs[6], cd[6];
        int tl = 0;
        bool filtered = false;
        if (valid) {
            const double *pl, *pt;
            if (P.mode == 0) {
                const int per_t = P.n_lon * P.n_d;
                int it = k / per_t;
                int rem = k - it * per_t;
                int il = rem / P.n_d;
                int id = rem - il * P.n_d;
                pl = P.lon_coef + (size_t)(it * P.n_lon + il) * 6;
                pt = P.lat_coef + (size_t)(low_vel ? k : it * P.n_d + id) * 6;
                tl = P.traj_len[it];
                filtered = (in.lon_mode == RP_STOPPING) && !(in.x0_lon[0] < P.lon_samples[il]);
            } else {
                pl = P.lon_coef + (size_t)k * 6;
                pt = P.lat_coef + (size_t)k * 6;
                tl = P.traj_len[k];
                filtered = P.skip != nullptr && P.skip[k] != 0;
            }
#pragma unroll
            for (int q = 0; q < 6; ++q) { cs[q] = pl[q]; cd[q] = pt[q]; }
            if (tl > Np1) tl = Np1;
        }
        const bool live = valid && !filtered;
        const bool in_traj = live && i < tl;
        __syncthreads();                                                            // flags initialised

        // ---- polynomial evaluation (reactive_planner.py:733-777) -------------------------------
        double s = 0., sv = 0., sa = 0., d = 0., dv = 0., da = 0.;
        if (in_traj) {
            const double tt = (double)i * dt;
            const double t2 = tt * tt, t3 = t2 * tt, t4 = t2 * t2, t5 = t4 * tt;
            s = poly_pos(cs, tt, t2, t3, t4, t5);
            sv = poly_vel(cs, tt, t2, t3, t4);
            sa = poly_acc(cs, tt, t2, t3);
            if (!low_vel) {
                d = poly_pos(cd, tt, t2, t3, t4, t5);
                dv = poly_vel(cd, tt, t2, t3, t4);
                da = poly_acc(cd, tt, t2, t3);
            } else {
                const double s1 = s - cs[0];                    // s - s[0]; s[0] == c0 exactly
                const double s2 = s1 * s1, s3 = s2 * s1, s4 = s2 * s2, s5 = s4 * s1;
                d = poly_pos(cd, s1, s2, s3, s4, s5);
                dv = poly_vel(cd, s1, s2, s3, s4);
                da = poly_acc(cd, s1, s2, s3);
            }
            if (fabs(sv) < kEps) sv = 0.0;
            if (fabs(dv) < kEps) dv = 0.0;
            if (!draw) {                                         // pre-filter (:796-805)
                unsigned bits = 0u;
                if (fabs(sa) > P.lim.a_max) bits |= 1u;
                if (sv < -kEps) bits |= 2u;
                if (bits) atomicOr(&s_flags[2], bits);
            }
        }
        __syncthreads();                                                            // pre-filter known
        const unsigned pre = lane ? s_flags[2] : 0u;
        const bool alive = live && pre == 0u;
        const bool act = alive && i < tl;

        // ---- orientation (reactive_planner.py:810-873) -------------------------------------------
        double dp = 0., dpp = 0., lam = 0., th_ref = 0., th_cl = 0., th_gl = 0.;
        int j0 = 0, j1 = 0, ub = 0;
        bool carry = false;
        if (act) {
            if (!low_vel) {
                if (sv > 0.001) dp = dv / sv; else dp = 0.;
                const double ddot = da - dp * sa;
                if (sv > 0.001) dpp = ddot / (sv * sv); else dpp = 0.;
            } else {
                dp = dv;
                dpp = da;
            }
            ub = upper_bound(R.pos, R.n, s);
            const bool wrap = (ub == R.n) || (ub == 0);          // s_idx == -1: python index wrap (App. B#8)
            j0 = wrap ? R.n - 1 : ub - 1;
            j1 = wrap ? 0 : ub;
            const double p0 = R.pos[j0], p1 = R.pos[j1];
            lam = (s - p0) / (p1 - p0);
            th_ref = interpolate_angle(s, p0, p1, R.theta[j0], R.theta[j1]);
            carry = !(sv > 0.001) && !low_vel;
            if (!carry) {
                th_cl = atan2(dp, 1.0);
                th_gl = th_cl + th_ref;
                s_th[i] = th_gl;
            }
            s_carry[i] = carry ? 1 : 0;
        }
        if (__syncthreads_or(carry ? 1 : 0)) {
            // standstill in high-velocity mode keeps the previous global orientation (:866-873)
            if (carry) {
                int j = i - 1;
                while (j >= 0 && s_carry[j]) --j;
                th_gl = j < 0 ? in.x0_orientation : s_th[j];
                th_cl = th_gl - th_ref;
            }
            __syncthreads();
            if (carry) s_th[i] = th_gl;
            __syncthreads();
        }

        // ---- curvature, velocity, acceleration (reactive_planner.py:876-896) -----------------------
        double kappa = 0., v = 0., a = 0.;
        if (act) {
            const double k0 = R.curv[j0], kd0 = R.curv_d[j0];
            const double k_r = (R.curv[j1] - k0) * lam + k0;
            const double k_r_d = (R.curv_d[j1] - kd0) * lam + kd0;
            const double oneKrD = (1 - k_r * d);
            const double cosT = cos(th_cl);
            const double tanT = tan(th_cl);
            const double q = cosT / oneKrD;
            kappa = (dpp + (k_r * dp + k_r_d * d) * tanT) * cosT * (q * q) + q * k_r;
            v = sv * (oneKrD / cosT);
            a = sa * oneKrD / cosT + ((sv * sv) / cosT) * (oneKrD * tanT * (kappa * oneKrD / cosT - k_r) -
                                                            (k_r_d * d + k_r * dp));
            s_kap[i] = kappa;
        }
        __syncthreads();                                                            // theta/kappa rows complete

        // ---- limits (reactive_planner.py:971-1017) + projection (:908-917) -------------------------
        double x = 0., y = 0., kdot = 0.;
        if (act) {
            const double th_prev = i > 0 ? s_th[i - 1] : 0.;
            const double kap_prev = i > 0 ? s_kap[i - 1] : 0.;
            const int r = check_constraints(P.lim, in.constraint_mask, dt, i, v, kappa, kap_prev, th_gl, th_prev, a);
            if (r != R_NONE) atomicMin(&s_flags[0], ((unsigned)i << 8) | (unsigned)r);
            kdot = i > 0 ? kappa - kap_prev : 0.;                  // np.append([0], np.diff(kappa_gl)) (:923)
            if (!project_to_cartesian(R, s, d, ub, x, y)) {
                atomicMin(&s_flags[1], (unsigned)i);
                x = 0.; y = 0.;
            }
        }
        __syncthreads();                                                            // verdict known
        const unsigned bad = lane ? s_flags[0] : NONE;
        const unsigned pbad = lane ? s_flags[1] : NONE;
        const bool kin_ok = alive && bad == NONE && pbad == NONE;
        const bool keep = alive && (kin_ok || draw);              // states are produced for these
        if (act && pbad != NONE && (unsigned)i > pbad) { x = 0.; y = 0.; }   // draw mode: loop broke at pbad

        // ---- horizon extension (trajectories.py:168-197, :302-332) ---------------------------------
        if (keep && i == tl - 1) {
            s_last[0] = x; s_last[1] = y; s_last[2] = th_gl; s_last[3] = v; s_last[4] = a; s_last[5] = kappa;
            s_last[6] = kdot; s_last[7] = s; s_last[8] = d; s_last[9] = th_cl; s_last[10] = sv; s_last[11] = sa;
            s_last[12] = dv; s_last[13] = da;
        }
        __syncthreads();
        const bool tail = keep && i >= tl;
        double cosL = 0., sinL = 0.;
        if (tail) {
            const double tau = (double)(i - tl + 1) * dt;         // np.arange(1, steps + 1) * dt
            a = s_last[4];
            double v_tmp = s_last[3] + tau * a;
            v_tmp = v_tmp * (v_tmp >= 0 ? 1.0 : 0.0);
            v = v_tmp;
            th_gl = s_last[2];
            kappa = s_last[5];
            kdot = s_last[6];
            cosL = cos(th_gl);
            sinL = sin(th_gl);
            s_tx[i] = dt * v_tmp * cosL;
            s_ty[i] = dt * v_tmp * sinL;
            // curvilinear: the velocity extrapolation multiplies the still-zero tail acceleration (App. B#7)
            double sv_tmp = s_last[10] + tau * 0.0;
            sv = sv_tmp * (sv_tmp >= 0 ? 1.0 : 0.0);
            dv = s_last[12] + tau * 0.0;
            sa = s_last[11];
            da = s_last[13];
            th_cl = s_last[9];
            s = s_last[7] + tau * s_last[10];
            d = s_last[8] + tau * s_last[12];
        }
        __syncthreads();
        if (tail) {                                               // np.cumsum: sequential adds
            double ax = 0., ay = 0.;
            for (int j = tl; j <= i; ++j) {
                if (j == tl) { ax = s_tx[j]; ay = s_ty[j]; }
                else { ax += s_tx[j]; ay += s_ty[j]; }
            }
            x = s_last[0] + ax;
            y = s_last[1] + ay;
        }

        // ---- cost (cost_function.py:51-71, :85-92); np.sum order per SURVEY App. B#5 ----------------
        const bool costed = kin_ok && in.cost_kind != RP_COST_NONE;
        if (costed) {
            const bool fs = in.cost_kind == RP_COST_FAILSAFE;
            const double wa = fs ? 1.0 : in.w_a;
            const double dd = fs ? 0.0 : in.desired_d;
            const double t0 = wa * a;
            s_ct[i] = t0 * t0;
            const double t1 = 5 * (v - in.desired_speed);
            s_ct[Np1 + i] = t1 * t1;
            const double t2 = 0.25 * (in.desired_s - s);
            s_ct[2 * Np1 + i] = t2 * t2;
            const double t3 = 0.25 * (dd - d);
            s_ct[3 * Np1 + i] = t3 * t3;
            const double t4 = 0.25 * fabs(th_cl);
            s_ct[4 * Np1 + i] = t4 * t4;
        }
        __syncthreads();
        const bool par_sum = Np1 >= 8 && Np1 <= 128;
        if (costed && par_sum) {
            for (int q = i; q < 40; q += Np1) {
                const double* row = s_ct + (q >> 3) * Np1;
                const int j = q & 7;
                double r = row[j];
                for (int m = 8 + j; m < Np1 - (Np1 % 8); m += 8) r += row[m];
                s_acc[q] = r;
            }
        }
        __syncthreads();
        if (costed) {
            for (int u = i; u < 5; u += Np1) {
                const double* row = s_ct + u * Np1;
                double res;
                if (par_sum) {
                    const double* r = s_acc + u * 8;
                    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
                    for (int m = Np1 - (Np1 % 8); m < Np1; ++m) res += row[m];
                } else {
                    res = np_pairwise_sum(row, Np1);
                }
                s_sum[u] = res;
            }
        }
        // the scalar end/mid terms need v[-1], v[mid], s[-1], d[-1], theta_cl[-1]: park them in s_acc tail
        __syncthreads();
        if (costed) {
            if (i == Np1 - 1) { s_last[0] = v; s_last[1] = s; s_last[2] = d; s_last[3] = th_cl; }
            if (i == Np1 / 2) s_last[4] = v;                      // v[int(len(v) / 2)]
        }

        // ---- ego-vs-obstacle check (reactive_planner.py:1026-1046) ---------------------------------
        if (kin_ok && in.check_collision) {
            double st, ct;
            sincos(th_gl, &st, &ct);
            const double ecx = x + P.wb_rear * ct;
            const double ecy = y + P.wb_rear * st;
            const int tidx = in.x0_time_step + i * in.factor;
            const double* rows = dyn_stage ? dyn_stage + (size_t)i * O.n_dyn * kBoxStride : nullptr;
            if (ego_collides(O, rows, tidx, ecx, ecy, ct, st, P.half_len, P.half_wid, P.r_ego))
                atomicMin(&s_flags[3], (unsigned)i);
        }

        // ---- state block --------------------------------------------------------------------------
        if (keep && P.states != nullptr) {
            double* o = P.states + (size_t)(P.states_by_slot ? slot : k) * 14 * Np1 + i;
            o[0] = x; o[Np1] = y; o[2 * Np1] = th_gl; o[3 * Np1] = v; o[4 * Np1] = a; o[5 * Np1] = kappa;
            o[6 * Np1] = kdot; o[7 * Np1] = s; o[8 * Np1] = d; o[9 * Np1] = th_cl; o[10 * Np1] = sv;
            o[11 * Np1] = sa; o[12 * Np1] = dv; o[13 * Np1] = da;
        }
        __syncthreads();

        // ---- per-candidate verdict ------------------------------------------------------------------
        if (valid && i == 0 && P.info != nullptr) {
            int status, reason = R_NONE, step = -1;
            double cost = __longlong_as_double(0x7ff8000000000000LL);   // NaN
            if (filtered) {
                status = ST_FILTERED;
            } else if (pre != 0u) {
                status = ST_KINEMATIC;
                reason = (pre & 1u) ? R_ACCELERATION : R_VELOCITY;
            } else if (bad != NONE) {
                status = ST_KINEMATIC;
                reason = (int)(bad & 0xFFu);
                step = (int)(bad >> 8);
            } else if (pbad != NONE) {
                status = ST_KINEMATIC;
                reason = R_PROJECTION;
                step = (int)pbad;
            } else {
                status = ST_FEASIBLE;
                if (in.cost_kind != RP_COST_NONE) {
                    const bool fs = in.cost_kind == RP_COST_FAILSAFE;
                    const double dd = fs ? 0.0 : in.desired_d;
                    double costs = 0.0;
                    costs += s_sum[0];
                    if (!fs && in.has_desired_speed) {
                        const double e1 = s_last[0] - in.desired_speed, e2 = s_last[4] - in.desired_speed;
                        costs += s_sum[1] + (50 * (e1 * e1)) + (100 * (e2 * e2));
                    }
                    if (!fs && in.has_desired_s) {
                        const double e = 20 * (in.desired_s - s_last[1]);
                        costs += s_sum[2] + e * e;
                    }
                    {
                        const double e = 20 * (dd - s_last[2]);
                        costs += s_sum[3] + e * e;
                    }
                    {
                        const double e = 5 * fabs(s_last[3]);
                        costs += s_sum[4] + e * e;
                    }
                    cost = costs;
                }
                const unsigned cstep = s_flags[3];
                if (cstep != NONE) { status = ST_COLLISION; step = (int)cstep; }
            }
            P.info[k] = pack_info(status, reason, step);
            if (P.cost) P.cost[k] = cost;
        }
        __syncthreads();                                                            // scratch reusable
    }
}

// ------------------------------------------------------------------------------------------------
// a12/a14: arg-min over feasible, collision-free candidates on (cost, enumeration index) -- the
// element a stable ascending sort puts first -- then the counters of the cycle.  One block.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool lex_less(double c1, int i1, double c2, int i2) {
    return (c1 < c2) || (c1 == c2 && i1 < i2);
}

__global__ void __launch_bounds__(1024) finalize_kernel(const double* __restrict__ cost, const int* __restrict__ info,
                                                        int first, int count, int n_cand, PlanResultDev* out) {
    __shared__ double w_cost[32];
    __shared__ int w_idx[32];
    __shared__ int counts[16];
    __shared__ double best_cost_s;
    __shared__ int best_idx_s;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid < 16) counts[tid] = 0;

    double bc = __longlong_as_double(0x7ff0000000000000LL);       // +inf
    int bi = 0x7fffffff;
    for (int q = tid; q < count; q += blockDim.x) {
        const int k = first + q;
        if ((info[k] & 0xFF) == ST_FEASIBLE) {
            const double cc = cost[k];
            if (lex_less(cc, k, bc, bi)) { bc = cc; bi = k; }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double oc = __shfl_down_sync(0xffffffffu, bc, off);
        const int oi = __shfl_down_sync(0xffffffffu, bi, off);
        if (lex_less(oc, oi, bc, bi)) { bc = oc; bi = oi; }
    }
    if (lane == 0) { w_cost[warp] = bc; w_idx[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        bc = lane < nw ? w_cost[lane] : __longlong_as_double(0x7ff0000000000000LL);
        bi = lane < nw ? w_idx[lane] : 0x7fffffff;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double oc = __shfl_down_sync(0xffffffffu, bc, off);
            const int oi = __shfl_down_sync(0xffffffffu, bi, off);
            if (lex_less(oc, oi, bc, bi)) { bc = oc; bi = oi; }
        }
        if (lane == 0) { best_cost_s = bc; best_idx_s = bi; }
    }
    __syncthreads();
    const double wc = best_cost_s;
    const int wi = best_idx_s;
    const bool has_winner = wi != 0x7fffffff;

    // counters: [0] feasible(kin), [1] colliders before winner, [2] colliders total, [3] filtered, [8..15] reasons
    int l_feas = 0, l_colb = 0, l_colt = 0, l_filt = 0;
    int l_reason[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int q = tid; q < count; q += blockDim.x) {
        const int k = first + q;
        const int w = info[k];
        const int st = w & 0xFF;
        if (st == ST_FEASIBLE) {
            ++l_feas;
        } else if (st == ST_COLLISION) {
            ++l_feas;
            ++l_colt;
            // lazy check visits candidates in (cost, index) order up to the winner (App. B#12)
            if (!has_winner || lex_less(cost[k], k, wc, wi)) ++l_colb;
        } else if (st == ST_KINEMATIC) {
            const int r = (w >> 8) & 0x7;
#pragma unroll
            for (int z = 0; z < 8; ++z) l_reason[z] += (r == z) ? 1 : 0;
        } else {
            ++l_filt;
        }
    }
    auto wsum = [](int v) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        return v;
    };
    l_feas = wsum(l_feas); l_colb = wsum(l_colb); l_colt = wsum(l_colt); l_filt = wsum(l_filt);
#pragma unroll
    for (int z = 0; z < 8; ++z) l_reason[z] = wsum(l_reason[z]);
    if (lane == 0) {
        atomicAdd(&counts[0], l_feas); atomicAdd(&counts[1], l_colb); atomicAdd(&counts[2], l_colt);
        atomicAdd(&counts[3], l_filt);
#pragma unroll
        for (int z = 0; z < 8; ++z) atomicAdd(&counts[8 + z], l_reason[z]);
    }
    __syncthreads();
    if (tid == 0) {
        rp_plan_result& r = out->r;
        r.winner = has_winner ? wi : -1;
        r.winner_cost = has_winner ? wc : __longlong_as_double(0x7ff8000000000000LL);
        r.n_candidates = count;
        r.n_feasible = counts[0];
        r.n_infeasible_kinematics = count - counts[3] - counts[0];
        r.n_infeasible_collision = counts[1];
        r.n_collision_total = counts[2];
        for (int z = 0; z < 8; ++z) r.reason_counts[z] = counts[8 + z];
        out->n_filtered = counts[3];
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

// colliders of this shard ranked before the GLOBAL winner (lazy collision count, App. B#12)
__global__ void __launch_bounds__(256) count_before_kernel(const double* __restrict__ cost, const int* __restrict__ info,
                                                           int first, int count, const double* __restrict__ winner,
                                                           double* __restrict__ out) {
    __shared__ int total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    const double wc = winner[0];
    const double wi = winner[1];
    const bool none = !(wi < __longlong_as_double(0x7ff0000000000000LL));
    int local = 0;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < count; q += gridDim.x * blockDim.x) {
        const int k = first + q;
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

// pycrcc.CollisionChecker.collide for a batch of ego boxes (rp_collide_poses)
__global__ void collide_kernel(int n, const double* __restrict__ pose, const int* __restrict__ tidx, double hl,
                               double hw, double r_ego, ObstacleTables O, uint8_t* __restrict__ hit) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    double st, ct;
    sincos(pose[3 * g + 2], &st, &ct);
    hit[g] = ego_collides(O, nullptr, tidx[g], pose[3 * g], pose[3 * g + 1], ct, st, hl, hw, r_ego) ? 1 : 0;
}

}  // namespace rp
