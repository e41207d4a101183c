// rp_cand.cuh -- the candidate-major form of the fused path (a4-a13 of SURVEY.md section 8a) for LARGE bundles and
// batches of scenarios in select-only mode: one thread per candidate, marching through the horizon sequentially.
// (reactive_planner.py:715-1063, cost_function.py:51-92, trajectories.py:168-332)
//
// Why a second mapping.  fused_kernel (rp_fused.cuh) gives every (candidate, step) pair a thread; that is what a
// 120-3000 candidate replanning bundle needs to fill 148 SMs, but it pays ~12 block barriers per group, exchanges
// every sequential-in-time quantity (theta[i-1], kappa[i-1], standstill carry, cumsum, numpy's summation order)
// through shared-memory rows and re-derives indices per element: ~2200 warp instructions per candidate-timestep of
// which 27 % are FP64.  A bundle of >= ~25 000 candidates fills the machine with candidates alone, and then the
// reference's own loop structure is the cheapest one: everything sequential-in-time lives in registers, there are
// no barriers, and the lanes of a warp (adjacent candidates of one sampled t) share traj_len, the dynamic
// obstacles of the step (one broadcast load per obstacle) and -- computed ONCE per warp, see cand_march -- everything
// that depends on the longitudinal polynomial alone.  ~720 warp instructions per 32 candidate-timesteps in round 1;
// ~370 at the end of round 2 (lateral table, two loops per march, collision checks out of the march).
//
// Collision checks.  check_collision = 1 (a flag for every feasible candidate): inside the march, per step.
// check_collision = 2 (the reference's lazy pass, reactive_planner.py:1031-1063): the march stores the ego box of every
// step (PlanParams::pose) and the verdicts come afterwards from deferred_collision_kernel /
// deferred_collision_list_kernel for the candidates that can be ranked before the winner (see there).
//
// The arithmetic is expression-for-expression the one of fused_kernel (same rp_device.cuh functions; divisions are
// IEEE quotients in both), so both kernels give identical bits; only the schedule differs.  Not handled here (the
// host routes these to fused_kernel): state output, draw mode, index mode, N + 1 > 128 (numpy's recursive pairwise
// split), and bundles too small to fill the machine (one march of a warp is ~0.08 ms whatever the load).
//
// Work distribution: the host sorts the segments (one per sampled t) by traj_len, longest first, and cuts them
// into chunks of 32 candidates; warps of a persistent grid (one 512-thread block per SM) draw chunks from a global
// counter.  cand_batch_kernel does the same over the chunks of MANY scenarios (rp_batch_*).
#pragma once
#include "rp_fused.cuh"

namespace rp {

#ifndef RP_CAND_THREADS
#define RP_CAND_THREADS 512
#endif
#ifndef RP_CAND_MIN_BLOCKS
#define RP_CAND_MIN_BLOCKS 1
#endif

// Dynamic obstacles of every time step as single-precision bounding circles, rows [step][obstacle] of
// (cx, cy, squared reach, -) relative to the obstacle-table origin, written once per launch by dyn_rows_kernel and
// read through L1: all lanes of a warp are at the same step, so a row is one broadcast load, and one copy serves
// every block of the SM (shared memory is left to the cost accumulators).  The pre-reject is conservative (reach
// inflated by the fp32 rounding bound, build_obstacle_tables) and branch-free over the obstacles -- the circle tests of one step are independent
// instructions, not a serial chain; survivors go to the exact fp64 SAT, which alone decides a hit.
// An obstacle absent at a step is parked at 1e30 with zero reach (inf <= 0 is false).
// `relevant`: bit o clear = obstacle o (< 32) was already excluded for this step by the row-level line test (lon_part).
__device__ __forceinline__ bool dyn_collides_f32(const ObstacleTables& O, const float4* __restrict__ row, int tidx,
                                                 double cx, double cy, double ca, double sa, double ahl, double ahw,
                                                 unsigned relevant) {
    const float ex = (float)(cx - O.org_x), ey = (float)(cy - O.org_y);
    for (int base = 0; base < O.n_dyn; base += 32) {
        const int cnt = min(32, O.n_dyn - base);
        unsigned mask = 0u;
        if (base == 0 && relevant != 0xffffffffu) {
            for (unsigned todo = relevant; todo;) {             // warp-uniform: the lanes share the row
                const int o = __ffs(todo) - 1;
                todo &= todo - 1u;
                const float4 r = row[o];
                const float dx = r.x - ex, dy = r.y - ey;
                if (dx * dx + dy * dy <= r.z) mask |= 1u << o;
            }
        } else {
#pragma unroll 8
            for (int o = 0; o < cnt; ++o) {
                const float4 r = row[base + o];
                const float dx = r.x - ex, dy = r.y - ey;
                if (dx * dx + dy * dy <= r.z) mask |= 1u << o;
            }
        }
        while (mask) {
            const int o = base + __ffs(mask) - 1;
            mask &= mask - 1u;
            const double* b = O.dyn_box + (size_t)(O.dyn_off[o] + (tidx - O.dyn_t0[o])) * kBoxStride;
            if (obb_obb_overlap(cx, cy, ca, sa, ahl, ahw, b)) return true;
        }
    }
    return false;
}

// first index with a[idx] > x (identical to upper_bound()), from the O(1) guess of a uniform table; returns the
// bracketing values lo = a[idx - 1] (-inf if idx == 0) and hi = a[idx] (+inf if idx == n) it verified the guess with
__device__ __forceinline__ int locate_segment(const double* __restrict__ a, int n, double x, double inv_step, double& lo,
                                              double& hi) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double f = (x - a[0]) * inv_step;
    int j = f > 0.0 ? (f < (double)(n - 1) ? (int)f + 1 : n) : 0;
    hi = j < n ? a[j] : inf;
    lo = j > 0 ? a[j - 1] : -inf;
    while (hi <= x) { lo = hi; ++j; hi = j < n ? a[j] : inf; }
    while (lo > x) { hi = lo; --j; lo = j > 0 ? a[j - 1] : -inf; }
    return j;
}

// ---- one polynomial step of one candidate (reactive_planner.py:733-935 loop body) as straight-line code --------
// Everything data-dependent that is cheap is a select, not a branch (standstill guards, the five ordered limit
// checks, the projection-domain test), and the sixteen divisions go through Divider<EXACT>: shared refined
// reciprocals with a sticky reject word (EXACT = false) or plain a / b (EXACT = true).  Large basic blocks let the
// compiler interleave the independent FP64 chains -- the kernel is bound by FP64 dependency latency, not issue.
// Remaining branches: low-velocity mode (launch-uniform), the reference-segment walk, the standstill carry (rare),
// orientation folding (never taken on sane inputs), the second segment search when the tables differ.
struct StepIn {
    const double *cs, *cd;       // the candidate's longitudinal / lateral coefficients (re-read every step: L1-resident,
                                 // the longitudinal row is a warp-wide broadcast; keeps 24 registers free)
    const double* lr;            // LATROWS: the candidate's column of the lateral table (+ i * lr_stride per step)
    int lr_stride;
    double th_prev, kap_prev;
    int i;
};
struct StepOut {
    double x, y, th_gl, th_cl, v, a, kappa, s, sv, d, dv;
    double cn, sn;       // cos / sin of th_gl
    unsigned pre;        // pre-filter bits of this step (:796-805)
    int reason;          // first violated limit of this step (R_NONE if none)
    int proj_fail;       // projection domain left at this step (:911-917)
    unsigned reject;     // sign bit set: a division left the fast path's window -> redo with EXACT = true
};

// Everything of a step that depends on the LONGITUDINAL polynomial alone: s, s_dot, s_ddot, the reference segment and
// what is interpolated on it, the projection base point, two reciprocals.  In a grid bundle it is identical for all
// candidates of one (t, lon) pair -- a whole warp when n_d is a multiple of 32 -- which is what cand_march
// exploits (rows computed cooperatively by the warp and read back from shared memory).
struct LonRow {
    double s, sv, sa;            // sv after the eps clamp
    double y_sv, y_sv2;          // refined reciprocals of the guarded s_dot and its square
    double th_ref, k_r, k_r_d;   // reference heading / curvature / curvature rate at s
    double c_ref, s_ref;         // cos / sin of th_ref
    double bx, by, nx, ny;       // frame base point and interpolated pseudo-normal: (x, y) = b + d * n
    unsigned flags;
    unsigned dynmask;            // dynamic obstacles of this step that can reach ANY pose on the lateral line b + d * n
};
enum : unsigned { LR_PRE = 3u, LR_MOVING = 4u, LR_OK_S = 8u, LR_REJECT = 16u };
constexpr int kLonRowDoubles = 14;
// row slots per warp: 32 in general; the ONE_GROUP instantiation (one longitudinal polynomial per warp) may refresh fewer
// steps at a time to leave shared memory for more resident warps
#ifndef RP_CAND_ONE_GROUP_SLOTS
#define RP_CAND_ONE_GROUP_SLOTS 32
#endif
__host__ __device__ constexpr int warp_row_doubles(int slots) { return kLonRowDoubles * slots + slots; }   // + flag words + dynamic-obstacle masks
constexpr int kWarpRowDoubles = warp_row_doubles(32);

// DYNMASK: also compute the row's mask of dynamic obstacles that can reach its lateral line (not needed when the march
// does not check collisions)
template <bool EXACT, bool DYNMASK = true>
__device__ __forceinline__ LonRow lon_part(const PlanParams& P, const RefTables& R, const double* __restrict__ cs_ptr, int i) {
    Divider<EXACT> D;
    LonRow o;
    double cs[6];
    {
        const double2* a = reinterpret_cast<const double2*>(cs_ptr);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const double2 u = __ldg(a + q);
            cs[2 * q] = u.x; cs[2 * q + 1] = u.y;
        }
    }
    // ---- polynomial evaluation (reactive_planner.py:733-777) ---------------------------------------
    const double tt = (double)i * P.in.dt;
    const double t2 = tt * tt, t3 = t2 * tt, t4 = t2 * t2, t5 = t4 * tt;
    const double s = poly_pos(cs, tt, t2, t3, t4, t5);
    double sv = poly_vel(cs, tt, t2, t3, t4);
    const double sa = poly_acc(cs, tt, t2, t3);
    sv = fabs(sv) < kEps ? 0.0 : sv;
    unsigned flags = (fabs(sa) > P.lim.a_max ? 1u : 0u) | (sv < -kEps ? 2u : 0u);    // pre-filter (:796-805)
    const bool moving = sv > 0.001;
    flags |= moving ? LR_MOVING : 0u;
    const double svs = moving ? sv : 1.0;                       // guarded divisor: quotients are discarded at standstill
    o.y_sv = D.rcp(svs);
    o.y_sv2 = D.rcp(svs * svs);
    // reference segment (:835): first index with ref_pos > s, from an O(1) guess on the near-uniform table; the two
    // loads that verify the guess ARE ref_pos[j0], ref_pos[j1]
    double p0, p1;
    const int ub = locate_segment(R.pos, R.n, s, P.ref_inv_step, p0, p1);
    const bool wrap = (ub == R.n) || (ub == 0);                 // s_idx == -1: python index wrap (App. B#8)
    const int j0 = wrap ? R.n - 1 : ub - 1;
    const int j1 = wrap ? 0 : ub;
    if (wrap) { p0 = R.pos[j0]; p1 = R.pos[j1]; }
    const double seg_len = p1 - p0;
    const double y_seg = D.rcp(seg_len);
    const double lam = D.div(s - p0, seg_len, y_seg);
    // interpolate_angle (utility/utils_coordinate_system.py:25-43)
    const double th0 = R.theta[j0];
    double th_ref = D.div((R.theta[j1] - th0) * (s - p0), seg_len, y_seg) + th0;
    if (th_ref > kTwoPi || th_ref < -kTwoPi) th_ref = make_valid_orientation(th_ref);
    const double k0 = R.curv[j0], kd0 = R.curv_d[j0];
    o.k_r = (R.curv[j1] - k0) * lam + k0;
    o.k_r_d = (R.curv_d[j1] - kd0) * lam + kd0;
    // ---- base of (s, d) -> (x, y) (:908-917; project_to_cartesian) ----------------------------------------
    {
        const int n = R.n;
        flags |= (s >= R.ps[0] && s <= R.ps[n - 1]) ? LR_OK_S : 0u;
        const int ub_ps = R.same_s ? ub : upper_bound_guess(R.ps, n, s, P.ps_inv_step);
        int j = ub_ps - 1;
        j = j > n - 2 ? n - 2 : j;
        j = j < 0 ? 0 : j;                                      // only outside the domain
        double lam2;
        if (R.same_s && j == j0 && j1 == j0 + 1) {
            lam2 = lam;                                         // same dividend, same divisor
        } else {
            const double sl = R.ps[j + 1] - R.ps[j];
            lam2 = D.div(s - R.ps[j], sl, D.rcp(sl));
        }
        const double p0x = R.px[j], p0y = R.py[j];
        o.bx = p0x + lam2 * (R.px[j + 1] - p0x);
        o.by = p0y + lam2 * (R.py[j + 1] - p0y);
        const double n0x = R.nx[j], n0y = R.ny[j];
        o.nx = n0x + lam2 * (R.nx[j + 1] - n0x);
        o.ny = n0y + lam2 * (R.ny[j + 1] - n0y);
    }
    o.s = s; o.sv = sv; o.sa = sa; o.th_ref = th_ref;
    sincos(th_ref, &o.s_ref, &o.c_ref);
    // Dynamic obstacles that matter at this step for ANY candidate sharing this longitudinal motion: the rear axle of
    // every such candidate lies on the line b + d * n, its box centre within wb_rear of that, so an obstacle whose
    // centre is farther than (reach + wb_rear + margin) from the LINE cannot touch any of them.  Conservative, single
    // precision (row.w carries that radius, dyn_rows_kernel); the lanes test only the surviving obstacles.
    o.dynmask = 0xffffffffu;
    if (DYNMASK && P.dyn_rows != nullptr && P.obs.n_dyn <= 32) {
        const ObstacleTables& O = P.obs;
        const float4* row = P.dyn_rows + (size_t)i * O.n_dyn;
        const float fbx = (float)(o.bx - O.org_x), fby = (float)(o.by - O.org_y), fnx = (float)o.nx, fny = (float)o.ny;
        const float n2 = (fnx * fnx + fny * fny) * 1.0001f;
        unsigned m = 0u;
        for (int q = 0; q < O.n_dyn; ++q) {
            const float4 r = row[q];
            const float dx = r.x - fbx, dy = r.y - fby;
            const float cr = dx * fny - dy * fnx;
            if (cr * cr <= r.w * r.w * n2) m |= 1u << q;
        }
        o.dynmask = m;
    }
    o.flags = flags | ((D.reject & 0x80000000u) ? LR_REJECT : 0u);
    return o;
}

// the candidate's own part of the step, given the longitudinal row
// LATROWS: d, d_dot, d_ddot of the step come from the lateral table (lat_rows_thread; high-velocity grid bundles: the
// lateral polynomial over the time grid is the same for every lon sample) instead of three polynomial evaluations
template <bool EXACT, bool LATROWS>
__device__ __forceinline__ StepOut lat_part(const PlanParams& P, const RefTables& R, const LimitRcp& Y, const LonRow& L,
                                            const double* __restrict__ cd_ptr, double cs0, double th_prev, double kap_prev,
                                            int i) {
    const rp_plan_inputs& in = P.in;
    const bool low_vel = in.low_vel_mode != 0;
    const double dt = in.dt;
    Divider<EXACT> D;
    StepOut o;
    const double s = L.s, sv = L.sv, sa = L.sa;
    double d, dv, da;
    if (LATROWS) {
        // (the caller passes a pointer to the step's three values: the table entry itself, or a copy it fetched one
        // step ahead -- the table lives in L2, and the march would wait for it at the head of every step otherwise)
        d = cd_ptr[0]; dv = cd_ptr[1]; da = cd_ptr[2];
    } else {
    double cd[6];
    {
        const double2* b = reinterpret_cast<const double2*>(cd_ptr);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const double2 w = __ldg(b + q);
            cd[2 * q] = w.x; cd[2 * q + 1] = w.y;
        }
    }
    if (!low_vel) {
        const double tt = (double)i * dt;
        const double t2 = tt * tt, t3 = t2 * tt, t4 = t2 * t2, t5 = t4 * tt;
        d = poly_pos(cd, tt, t2, t3, t4, t5);
        dv = poly_vel(cd, tt, t2, t3, t4);
        da = poly_acc(cd, tt, t2, t3);
    } else {
        const double s1 = s - cs0;                              // s - s[0]; s[0] == c0 exactly
        const double s2 = s1 * s1, s3 = s2 * s1, s4 = s2 * s2, s5 = s4 * s1;
        d = poly_pos(cd, s1, s2, s3, s4, s5);
        dv = poly_vel(cd, s1, s2, s3, s4);
        da = poly_acc(cd, s1, s2, s3);
    }
    dv = fabs(dv) < kEps ? 0.0 : dv;
    }
    o.pre = L.flags & LR_PRE;

    // ---- orientation (:810-873) ----------------------------------------------------------------------
    const bool moving = (L.flags & LR_MOVING) != 0u;
    double dp, dpp;
    if (!low_vel) {
        const double svs = moving ? sv : 1.0;
        const double dp_q = D.div(dv, svs, L.y_sv);
        dp = moving ? dp_q : 0.;
        const double ddot = da - dp * sa;
        const double sv2 = svs * svs;
        const double dpp_q = D.div(ddot, sv2, L.y_sv2);
        dpp = moving ? dpp_q : 0.;
    } else {
        dp = dv;
        dpp = da;
    }
    const double th_ref = L.th_ref;
    const bool carry = !moving && !low_vel;
    const double k_r = L.k_r, k_r_d = L.k_r_d;
    double th_cl, th_gl, cosT, kappa, v, a;
    // ---- orientation (:842-873), curvature, velocity, acceleration (:876-896) ----------------------------------------
    if (!carry) {
        th_cl = rp_atan(dp);                                       // np.arctan2(dp, 1.0)
        th_gl = th_cl + th_ref;
        motion_moving(D, dp, dpp, d, k_r, k_r_d, sv, sa, cosT, kappa, v, a);
        heading_cos_sin(cosT, dp, L.c_ref, L.s_ref, o.cn, o.sn);
    } else {
        // standstill in high-velocity mode keeps the previous global orientation (:866-873)
        th_gl = i > 0 ? th_prev : in.x0_orientation;
        th_cl = th_gl - th_ref;
        double tanT;
        motion_carry(D, th_cl, dp, dpp, d, k_r, k_r_d, sv, sa, cosT, tanT, kappa, v, a);
        sincos(th_gl, &o.sn, &o.cn);
    }

    // ---- the five ordered limit checks (:971-1017), all evaluated, first violation selected ---------------
    {
        const Limits& Lm = P.lim;
        const unsigned mask = in.constraint_mask;
        const bool c_v = v < -kEps;
        const bool c_k = fabs(kappa) > Lm.kappa_max;
        const double yaw_q = D.div(th_gl - th_prev, dt, Y.y_dt);
        const double yaw_rate = i > 0 ? yaw_q : 0.;
        // round(np.float64, 5) == rint(x * 1e5) / 1e5   
        const bool c_y = fabs(D.div(rint(yaw_rate * 100000.0), 100000.0, Y.y_1e5)) > Lm.kappa_max * v;
        // cos(atan2(wb * kappa, 1))^2 == 1 / (1 + (wb * kappa)^2)  (see check_constraints)
        const double tk = Lm.wheelbase * kappa;
        const double kappa_dot_max = D.div_nz(Lm.v_delta_max * (1.0 + tk * tk), Lm.wheelbase, Y.y_wb);
        const double kd_q = D.div(kappa - kap_prev, dt, Y.y_dt);
        const double kappa_dot = i > 0 ? kd_q : 0.;
        const bool c_kd = fabs(kappa_dot) > kappa_dot_max;
        const bool fast = v > Lm.v_switch;
        const double vv = fast ? v : 1.0;
        const double a_hi_q = D.div_nz(Lm.a_max * Lm.v_switch, vv, D.rcp(vv));
        const double a_hi = fast ? a_hi_q : Lm.a_max;
        const bool c_a = !(-Lm.a_max <= a && a <= a_hi);
        int r = R_NONE;
        r = ((mask & C_ACCELERATION) && c_a) ? R_ACCELERATION : r;
        r = ((mask & C_KAPPA_DOT) && c_kd) ? R_KAPPA_DOT : r;
        r = ((mask & C_YAW_RATE) && c_y) ? R_YAW_RATE : r;
        r = ((mask & C_KAPPA) && c_k) ? R_KAPPA : r;
        r = ((mask & C_VELOCITY) && c_v) ? R_VELOCITY : r;
        o.reason = r;
    }
    // ---- (s, d) -> (x, y) ---------------------------------------------------------------------------------
    const bool ok = (L.flags & LR_OK_S) && (fabs(d) <= R.limit);
    o.x = ok ? L.bx + d * L.nx : 0.;
    o.y = ok ? L.by + d * L.ny : 0.;
    o.proj_fail = ok ? 0 : 1;
    o.th_gl = th_gl; o.th_cl = th_cl; o.v = v; o.a = a; o.kappa = kappa; o.s = s; o.sv = sv; o.d = d; o.dv = dv;
    o.reject = D.reject | ((L.flags & LR_REJECT) ? 0x80000000u : 0u);
    return o;
}

template <bool EXACT, bool LATROWS>
__device__ __forceinline__ StepOut poly_step(const PlanParams& P, const RefTables& R, const LimitRcp& Y, const StepIn& I) {
    const LonRow L = lon_part<EXACT>(P, R, I.cs, I.i);
    return lat_part<EXACT, LATROWS>(P, R, Y, L, LATROWS ? I.lr + (size_t)I.i * I.lr_stride : I.cd, __ldg(I.cs), I.th_prev,
                                    I.kap_prev, I.i);
}

// the rare exact redo: out of line, so its plain divisions (each with a slow-path call) stay out of the hot loop
template <bool LATROWS>
__device__ __noinline__ StepOut poly_step_exact(const PlanParams& P, const RefTables& R, const LimitRcp& Y, StepIn I) {
    return poly_step<true, LATROWS>(P, R, Y, I);
}

// the lateral table of a high-velocity grid bundle: entry [it][i][id] = d, d_dot (eps-clamped, :777), d_ddot of the lateral
// polynomial (t[it], d[id]) at time i * dt, i < traj_len[it] -- exactly what lat_part evaluates per candidate, once per
// (t, d) pair instead of once per (t, lon, d) candidate.  One thread per (it, chunk of 8 steps, id); it re-solves its own
// polynomial (coeff_thread's expressions) because the coefficient blocks of the same launch may not have finished.
__device__ __forceinline__ void lat_rows_thread(int q, int n_t, int n_d, int Np1, const double* __restrict__ t,
                                                const double* __restrict__ dsamp, const int* __restrict__ traj_len,
                                                double x0d, double x0dd, double x0ddd, double dt, double* __restrict__ out) {
    const int n_chunks = (Np1 + 7) / 8;
    if (q >= n_t * n_chunks * n_d) return;
    const int id = q % n_d;
    const int rest = q / n_d;
    const int chunk = rest % n_chunks, it = rest / n_chunks;
    int tl = traj_len[it];
    tl = tl > Np1 ? Np1 : tl;
    if (chunk * 8 >= tl) return;
    double c[6];
    solve_quintic(x0d, x0dd, x0ddd, dsamp[id], 0.0, 0.0, t[it], c);
    for (int i = chunk * 8; i < chunk * 8 + 8 && i < tl; ++i) {
        const double tt = (double)i * dt;
        const double t2 = tt * tt, t3 = t2 * tt, t4 = t2 * t2, t5 = t4 * tt;
        double dv = poly_vel(c, tt, t2, t3, t4);
        dv = fabs(dv) < kEps ? 0.0 : dv;
        double2* o = reinterpret_cast<double2*>(out + (((size_t)it * Np1 + i) * n_d + id) * 4);
        o[0] = make_double2(poly_pos(c, tt, t2, t3, t4, t5), dv);
        o[1] = make_double2(poly_acc(c, tt, t2, t3), 0.0);
    }
}

// rows [step][obstacle] for the launch's time window x0.time_step + step * factor (reactive_planner.py:1040)
__device__ __forceinline__ void dyn_rows_thread(int q, const ObstacleTables& O, int x0_time_step, int factor, int Np1,
                                                float r_ego_f_up, float wb_rear_f_up, float4* __restrict__ out) {
    if (q >= Np1 * O.n_dyn) return;
    const int step = q / O.n_dyn, o = q - step * O.n_dyn;
    const int kk = x0_time_step + step * factor - O.dyn_t0[o];
    const bool present = kk >= 0 && kk < O.dyn_len[o];
    float4 r = make_float4(1.0e30f, 1.0e30f, 0.0f, 0.0f);
    if (present) {
        const double* b = O.dyn_box + (size_t)(O.dyn_off[o] + kk) * kBoxStride;
        const float reach = (r_ego_f_up + (float)b[6]) * 1.000001f + O.dyn_margin;   // >= r_ego + r_obs + margin
        // .w: radius of the row-level test against the lateral LINE of the rear-axle positions (lon_part)
        r = make_float4((float)(b[0] - O.org_x), (float)(b[1] - O.org_y), reach * reach * 1.00001f,
                        (reach + wb_rear_f_up) * 1.0001f + 1.0e-3f);
    }
    out[q] = r;
}

__global__ void dyn_rows_kernel(ObstacleTables O, int x0_time_step, int factor, int Np1, float r_ego_f_up, float wb_rear_f_up,
                                float4* __restrict__ out) {
    dyn_rows_thread(blockIdx.x * blockDim.x + threadIdx.x, O, x0_time_step, factor, Np1, r_ego_f_up, wb_rear_f_up, out);
}

// ---- one candidate through the whole horizon (the per-lane body of both kernels below) ----------------------
// acc: this thread's accumulator column (+ (row * 8 + j) * BLOCK), s_vmid: this thread's parked v[mid] slot
// The warp evaluates the longitudinal rows cooperatively.  Its 32 lanes hold G distinct longitudinal polynomials
// ("groups": grid bundle -> the distinct lon samples among 32 consecutive candidates of one t, G = 1 when n_d is a
// multiple of 32; list form -> one per lane, G = 32).  Every W = 32 / G steps the warp computes the rows of the next
// W steps of all groups at once -- lane q does (group q / W, step base + q % W) -- into `rows` ([field][32] doubles +
// [32] flag words of this warp's shared memory); each lane then reads the row of its own group.  The (t, lon)-
// invariant part of a step therefore costs G / 32 of a per-lane evaluation.
// ALL 32 lanes of the warp call this function; lanes without a candidate (valid == false) mirror the chunk's last
// candidate and write nothing.
// ONE_GROUP: the host guarantees G == 1 for every chunk (grid form, n_d a multiple of 32, aligned shard): the group
// arithmetic folds away at compile time.
// the tiles (32 consecutive candidates) of the first pass of the deferred collision check: a pseudo-random sixteenth
// (the bound on the winner's cost comes from them)
#ifndef RP_DEFER_SAMPLE_SHIFT
#define RP_DEFER_SAMPLE_SHIFT 28          // 1 tile in 2^(32 - shift)
#endif
__device__ __forceinline__ bool defer_sampled(int tile) { return (((unsigned)tile * 0x9E3779B1u) >> RP_DEFER_SAMPLE_SHIFT) == 0u; }
// put candidate k on a pass's list: its bit in the tile's mask; whoever sets the first bit lists the tile
__device__ __forceinline__ void defer_enlist(const PlanParams& P, int k, int pass) {
    const int tile = k >> 5;
    if (atomicOr(P.defer_mask + tile, 1u << (k & 31)) == 0u)
        (pass ? P.defer_list2 : P.defer_list)[atomicAdd(P.defer_count + pass, 1)] = P.defer_tag | tile;
}

// DEFER: the march does not check collisions; it stores the ego box of every step while the candidate is still
// kinematically feasible (P.pose) and leaves feasible candidates ST_UNCHECKED for deferred_collision_kernel.
// (Letting the march itself check every 16th chunk -- taken first from the queue -- instead of a first deferred pass was
// measured: the slower chunks stretch the march by 19 us, the saved pass is worth 17.)
// SPLIT: all lanes of a warp share one traj_len (grid form: a chunk never crosses a sampled t) -- two loops, see below
template <int BLOCK, bool ONE_GROUP, int PF, bool LATROWS = false, int SLOTS = 32, bool DEFER = false, bool SPLIT = ONE_GROUP>
__device__ __forceinline__ void cand_march(const PlanParams& P, const RefTables& R, const LimitRcp& Y, int k, bool valid,
                                           double* __restrict__ acc, double* __restrict__ s_vmid,
                                           double* __restrict__ rows) {
    const int Np1 = P.Np1;
    const ObstacleTables& O = P.obs;
    const rp_plan_inputs& in = P.in;
    const bool low_vel = in.low_vel_mode != 0;
    const bool store_boxes = DEFER && P.pose != nullptr;
    const double dt = in.dt;
    const unsigned NONE = 0xFFFFFFFFu;
    const bool fs = in.cost_kind == RP_COST_FAILSAFE;
    const bool costed = in.cost_kind != RP_COST_NONE;
    const double w_a = fs ? 1.0 : in.w_a;
    const double des_d = fs ? 0.0 : in.desired_d;
    const bool use_v = costed && !fs && in.has_desired_speed;
    const bool use_s = costed && !fs && in.has_desired_s;
    // accumulator rows: 0 acceleration, 1 lateral offset, 2 orientation, then velocity / position when used
    const int row_v = 3, row_s = use_v ? 4 : 3;
    const int n8 = Np1 - (Np1 & 7);                      // numpy: elements [8, n8) go to the 8 accumulators
    const int mid = Np1 / 2;
    // np.sum in time order: n < 8 plain loop from 0.; otherwise 8 accumulators over [0, n8)
    // (shared memory), their pairwise tree, then the remainder added sequentially (registers).
    auto tree = [&](int row) {
        const double* r8 = acc + (size_t)row * 8 * BLOCK;
        return ((r8[0] + r8[BLOCK]) + (r8[2 * BLOCK] + r8[3 * BLOCK])) +
               ((r8[4 * BLOCK] + r8[5 * BLOCK]) + (r8[6 * BLOCK] + r8[7 * BLOCK]));
    };

    // ---- candidate decode (sampling.py:202-242 enumeration order) + the warp's longitudinal groups -------------
    const int lane = threadIdx.x & 31;
    StepIn I;
    int tl, grp, G;
    bool filtered;
    const double* cs_group0;                             // grid form: coefficients of group 0 (group g: + 6 g)
    if (P.mode == 0) {
        const int per_t = P.n_lon * P.n_d;
        const int it = k / per_t;
        const int rem = k - it * per_t;
        const int il = rem / P.n_d;
        const int id = rem - il * P.n_d;
        I.cs = P.lon_coef + (size_t)(it * P.n_lon + il) * 6;
        I.cd = P.lat_coef + (size_t)(low_vel ? k : it * P.n_d + id) * 6;
        I.lr = LATROWS ? P.lat_rows + ((size_t)it * Np1 * P.n_d + id) * 4 : nullptr;
        I.lr_stride = P.n_d * 4;
        tl = P.traj_len[it];
        filtered = (in.lon_mode == RP_STOPPING) && !(in.x0_lon[0] < P.lon_samples[il]);
        if (ONE_GROUP) {
            grp = 0;
            G = 1;
            cs_group0 = I.cs;
        } else {
            const int il0 = __shfl_sync(0xffffffffu, il, 0);
            grp = il - il0;
            G = __shfl_sync(0xffffffffu, grp, 31) + 1;   // lanes are consecutive candidates of one t: il is monotone
            cs_group0 = P.lon_coef + (size_t)(it * P.n_lon + il0) * 6;
        }
    } else {
        I.cs = P.lon_coef + (size_t)k * 6;
        I.cd = P.lat_coef + (size_t)k * 6;
        I.lr = nullptr;
        I.lr_stride = 0;
        tl = P.traj_len[k];
        filtered = P.skip != nullptr && P.skip[k] != 0;
        grp = lane;
        G = 32;
        cs_group0 = nullptr;
    }
    if (tl > Np1) tl = Np1;
    const int tl_warp = P.mode == 0 ? tl : __reduce_max_sync(0xffffffffu, tl);
    const int W = ONE_GROUP ? SLOTS : 32 / G;            // steps per refresh of the rows
    // the (group, step offset) this lane computes (idle if item_g >= G)
    const int item_g = (ONE_GROUP && SLOTS == 32) ? 0 : lane / W;
    const int item_w = (ONE_GROUP && SLOTS == 32) ? lane : lane - item_g * W;
    unsigned* const rflags = reinterpret_cast<unsigned*>(rows + kLonRowDoubles * SLOTS);
    int row_base = 0, row_next = 0;

    unsigned pre = 0u, bad = NONE, pbad = NONE, col = NONE;
    const double cs0 = low_vel ? __ldg(I.cs) : 0.;
    // (check_collision == 2, the reference's lazy pass: this schedule checks every feasible candidate while it marches.
    // A gate on the running cost -- skip the remaining poses once the partial cost exceeds the best collision-free
    // cost published so far -- was measured: the first wave of candidates finishes before any bound exists, and the
    // running sum costs the march more than the skipped checks of the second wave save: 0.360 vs 0.353 ms.)
    // values of the current / last polynomial step (the extension reads them after step tl - 1)
    double x = 0., y = 0., th_gl = 0., v = 0., a = 0., kappa = 0., s = 0., sv = 0., d = 0., dv = 0., th_cl = 0.;
    double cn = 1., sn = 0.;                       // cos / sin of th_gl
    double ax = 0., ay = 0.;                       // np.cumsum of the extension increments

    double lat_next[3] = {0., 0., 0.};             // LATROWS: the lateral values of the NEXT polynomial step
    if (LATROWS && tl > 0) {
        const double2* r = reinterpret_cast<const double2*>(I.lr);
        const double2 u = __ldg(r), w = __ldg(r + 1);
        lat_next[0] = u.x; lat_next[1] = u.y; lat_next[2] = w.x;
    }
    // the part of a step that the polynomial phase and the horizon extension share: cost terms, ego box / collision check,
    // end values
    auto step_tail = [&](const int i, const double px, const double py, const double c_a, const double c_v, const double c_s,
                         const double c_d, const double c_th, const unsigned heavy_dynmask) {
        // ---- cost terms in numpy's np.sum order (cost_function.py:51-71) ------------------------------
        if (costed) {
            const double t0 = w_a * c_a, t3 = 0.25 * (des_d - c_d), t4 = 0.25 * fabs(c_th);
            const double q0 = t0 * t0, q3 = t3 * t3, q4 = t4 * t4;
            double q1 = 0., q2 = 0.;
            if (use_v) { const double t1 = 5 * (c_v - in.desired_speed); q1 = t1 * t1; }
            if (use_s) { const double t2 = 0.25 * (in.desired_s - c_s); q2 = t2 * t2; }
            if (Np1 >= 8 && i < n8) {
                double* ap = acc + (size_t)(i & 7) * BLOCK;
                if (i < 8) {
                    ap[0] = q0; ap[8 * BLOCK] = q3; ap[16 * BLOCK] = q4;
                    if (use_v) ap[row_v * 8 * BLOCK] = q1;
                    if (use_s) ap[row_s * 8 * BLOCK] = q2;
                } else {
                    ap[0] += q0; ap[8 * BLOCK] += q3; ap[16 * BLOCK] += q4;
                    if (use_v) ap[row_v * 8 * BLOCK] += q1;
                    if (use_s) ap[row_s * 8 * BLOCK] += q2;
                }
            } else {
                // remainder of np.sum (or the plain loop of n < 8): the running results live in slot 0 of each row --
                // five doubles that would otherwise sit in registers through the whole march
                if (Np1 >= 8 && i == n8) {
                    const double ta = tree(0), td = tree(1), tt = tree(2);
                    acc[0] = ta; acc[8 * BLOCK] = td; acc[16 * BLOCK] = tt;
                    if (use_v) { const double tv = tree(row_v); acc[row_v * 8 * BLOCK] = tv; }
                    if (use_s) { const double ts = tree(row_s); acc[row_s * 8 * BLOCK] = ts; }
                } else if (Np1 < 8 && i == 0) {
                    acc[0] = 0.; acc[8 * BLOCK] = 0.; acc[16 * BLOCK] = 0.;
                    if (use_v) acc[row_v * 8 * BLOCK] = 0.;
                    if (use_s) acc[row_s * 8 * BLOCK] = 0.;
                }
                acc[0] += q0; acc[8 * BLOCK] += q3; acc[16 * BLOCK] += q4;
                if (use_v) acc[row_v * 8 * BLOCK] += q1;
                if (use_s) acc[row_s * 8 * BLOCK] += q2;
            }
            if (i == mid) s_vmid[0] = c_v;                   // v[int(len(v) / 2)]
        }

        // ---- ego-vs-obstacle check (reactive_planner.py:1026-1046) ------------------------------------
        if (DEFER) {
            // the box of this step for the deferred check (only candidates that end feasible are read back)
            if (store_boxes && valid && !filtered && bad == NONE && pbad == NONE && pre == 0u) {
                double2* rec = reinterpret_cast<double2*>(P.pose) + (((size_t)(k >> 5) * Np1 + i) * 32 + (k & 31)) * 2;
                rec[0] = make_double2(px + P.wb_rear * cn, py + P.wb_rear * sn);
                rec[1] = make_double2(cn, sn);
            }
        } else if (in.check_collision && col == NONE && bad == NONE && pbad == NONE && pre == 0u) {
            const double ecx = px + P.wb_rear * cn;
            const double ecy = py + P.wb_rear * sn;
            const int tidx = in.x0_time_step + i * in.factor;
            const bool hit = (P.dyn_rows ? dyn_collides_f32(O, P.dyn_rows + (size_t)i * O.n_dyn, tidx, ecx, ecy, cn, sn, P.half_len, P.half_wid,
                                                               i < tl ? heavy_dynmask : 0xffffffffu)
                                         : dyn_collides_global(O, tidx, ecx, ecy, cn, sn, P.half_len, P.half_wid, P.r_ego)) ||
                             static_collides<PF>(O, ecx, ecy, cn, sn, P.half_len, P.half_wid);
            if (hit) col = (unsigned)i;
        }
        if (i == Np1 - 1) {                                   // park the end values for the terminal terms
            v = c_v; s = c_s; d = c_d; th_cl = c_th;
        }
    };
    if (SPLIT) {
        // every lane of the warp has the chunk's traj_len: the polynomial steps and the horizon extension are two loops, so
        // the values the extension starts from are results of the first loop, not state carried through it
        for (int i = 0; i < tl; ++i) {
            double px, py;                             // rear-axle position of this step
            unsigned heavy_dynmask = 0xffffffffu;      // polynomial steps: the obstacles that can reach this step's lateral line
            double c_a, c_v, c_s, c_d, c_th;           // values entering the cost terms
            if (i == row_next && i < tl_warp) {            // warp-uniform: refresh the rows of steps i .. i + W - 1
                __syncwarp();
                if (item_g < G) {
                    const int step = i + item_w;
                    // grid form: every group has the chunk's traj_len; list form: W == 1, the item is the lane's own candidate
                    if (step < tl) {
                        const double* cs_item = P.mode == 0 ? cs_group0 + (size_t)item_g * 6 : I.cs;
                        const LonRow w = lon_part<false, !DEFER>(P, R, cs_item, step);
                        double* c = rows + lane;                   // (lane == item index: item_g * W + item_w)
                        c[0] = w.s; c[SLOTS] = w.sv; c[2 * SLOTS] = w.sa; c[3 * SLOTS] = w.y_sv; c[4 * SLOTS] = w.y_sv2;
                        c[5 * SLOTS] = w.th_ref; c[6 * SLOTS] = w.k_r; c[7 * SLOTS] = w.k_r_d; c[8 * SLOTS] = w.bx; c[9 * SLOTS] = w.by;
                        c[10 * SLOTS] = w.nx; c[11 * SLOTS] = w.ny; c[12 * SLOTS] = w.c_ref; c[13 * SLOTS] = w.s_ref;
                        rflags[lane] = w.flags;
                        if (!DEFER) rflags[SLOTS + lane] = w.dynmask;
                    }
                }
                __syncwarp();
                row_base = i;
                row_next = i + W;
            }
            {
                I.i = i; I.th_prev = th_gl; I.kap_prev = kappa;
                const int item = grp * W + (i - row_base);
                const double* c = rows + item;
                LonRow L;
                L.s = c[0]; L.sv = c[SLOTS]; L.sa = c[2 * SLOTS]; L.y_sv = c[3 * SLOTS]; L.y_sv2 = c[4 * SLOTS]; L.th_ref = c[5 * SLOTS];
                L.k_r = c[6 * SLOTS]; L.k_r_d = c[7 * SLOTS]; L.bx = c[8 * SLOTS]; L.by = c[9 * SLOTS]; L.nx = c[10 * SLOTS];
                L.ny = c[11 * SLOTS]; L.c_ref = c[12 * SLOTS]; L.s_ref = c[13 * SLOTS];
                L.flags = rflags[item];
                if (!DEFER) heavy_dynmask = rflags[SLOTS + item];
                double lat_now[3] = {lat_next[0], lat_next[1], lat_next[2]};
                if (LATROWS && i + 1 < tl) {                       // next step's lateral values: in flight during this step
                    const double2* r = reinterpret_cast<const double2*>(I.lr + (size_t)(i + 1) * I.lr_stride);
                    const double2 u = __ldg(r), w = __ldg(r + 1);
                    lat_next[0] = u.x; lat_next[1] = u.y; lat_next[2] = w.x;
                }
                StepOut o = lat_part<false, LATROWS>(P, R, Y, L, LATROWS ? lat_now : I.cd, cs0, th_gl, kappa, i);
                if (o.reject & 0x80000000u) o = poly_step_exact<LATROWS>(P, R, Y, I);
                pre |= o.pre;
                if (o.reason != R_NONE && bad == NONE) bad = ((unsigned)i << 8) | (unsigned)o.reason;
                if (o.proj_fail && pbad == NONE) pbad = (unsigned)i;
                x = o.x; y = o.y; th_gl = o.th_gl; th_cl = o.th_cl; v = o.v; a = o.a; kappa = o.kappa;
                s = o.s; sv = o.sv; d = o.d; dv = o.dv;
                cn = o.cn; sn = o.sn;
                px = x; py = y;
                c_a = a; c_v = v; c_s = s; c_d = d; c_th = th_cl;
            }
            step_tail(i, px, py, c_a, c_v, c_s, c_d, c_th, heavy_dynmask);
        }
        for (int i = tl; i < Np1; ++i) {
            double px, py, c_a, c_v, c_s, c_d, c_th;
                // ---- horizon extension (trajectories.py:168-197, :302-332) ---------------------------
                const double tau = (double)(i - tl + 1) * dt;      // np.arange(1, steps + 1) * dt
                double v_tmp = v + tau * a;
                v_tmp = v_tmp * (v_tmp >= 0 ? 1.0 : 0.0);
                const double ix = dt * v_tmp * cn, iy = dt * v_tmp * sn;
                if (i == tl) { ax = ix; ay = iy; } else { ax += ix; ay += iy; }   // np.cumsum: sequential adds
                px = x + ax; py = y + ay;
                c_a = a; c_v = v_tmp;
                c_s = s + tau * sv;                                // curvilinear tail (App. B#7)
                c_d = d + tau * dv;
                c_th = th_cl;
            step_tail(i, px, py, c_a, c_v, c_s, c_d, c_th, 0xffffffffu);
        }
    } else {
        for (int i = 0; i < Np1; ++i) {
            double px, py;                             // rear-axle position of this step
            unsigned heavy_dynmask = 0xffffffffu;      // polynomial steps: the obstacles that can reach this step's lateral line
            double c_a, c_v, c_s, c_d, c_th;           // values entering the cost terms
            if (i == row_next && i < tl_warp) {            // warp-uniform: refresh the rows of steps i .. i + W - 1
                __syncwarp();
                if (item_g < G) {
                    const int step = i + item_w;
                    // grid form: every group has the chunk's traj_len; list form: W == 1, the item is the lane's own candidate
                    if (step < tl) {
                        const double* cs_item = P.mode == 0 ? cs_group0 + (size_t)item_g * 6 : I.cs;
                        const LonRow w = lon_part<false, !DEFER>(P, R, cs_item, step);
                        double* c = rows + lane;                   // (lane == item index: item_g * W + item_w)
                        c[0] = w.s; c[SLOTS] = w.sv; c[2 * SLOTS] = w.sa; c[3 * SLOTS] = w.y_sv; c[4 * SLOTS] = w.y_sv2;
                        c[5 * SLOTS] = w.th_ref; c[6 * SLOTS] = w.k_r; c[7 * SLOTS] = w.k_r_d; c[8 * SLOTS] = w.bx; c[9 * SLOTS] = w.by;
                        c[10 * SLOTS] = w.nx; c[11 * SLOTS] = w.ny; c[12 * SLOTS] = w.c_ref; c[13 * SLOTS] = w.s_ref;
                        rflags[lane] = w.flags;
                        if (!DEFER) rflags[SLOTS + lane] = w.dynmask;
                    }
                }
                __syncwarp();
                row_base = i;
                row_next = i + W;
            }
            if (i < tl) {
                I.i = i; I.th_prev = th_gl; I.kap_prev = kappa;
                const int item = grp * W + (i - row_base);
                const double* c = rows + item;
                LonRow L;
                L.s = c[0]; L.sv = c[SLOTS]; L.sa = c[2 * SLOTS]; L.y_sv = c[3 * SLOTS]; L.y_sv2 = c[4 * SLOTS]; L.th_ref = c[5 * SLOTS];
                L.k_r = c[6 * SLOTS]; L.k_r_d = c[7 * SLOTS]; L.bx = c[8 * SLOTS]; L.by = c[9 * SLOTS]; L.nx = c[10 * SLOTS];
                L.ny = c[11 * SLOTS]; L.c_ref = c[12 * SLOTS]; L.s_ref = c[13 * SLOTS];
                L.flags = rflags[item];
                if (!DEFER) heavy_dynmask = rflags[SLOTS + item];
                double lat_now[3] = {lat_next[0], lat_next[1], lat_next[2]};
                if (LATROWS && i + 1 < tl) {                       // next step's lateral values: in flight during this step
                    const double2* r = reinterpret_cast<const double2*>(I.lr + (size_t)(i + 1) * I.lr_stride);
                    const double2 u = __ldg(r), w = __ldg(r + 1);
                    lat_next[0] = u.x; lat_next[1] = u.y; lat_next[2] = w.x;
                }
                StepOut o = lat_part<false, LATROWS>(P, R, Y, L, LATROWS ? lat_now : I.cd, cs0, th_gl, kappa, i);
                if (o.reject & 0x80000000u) o = poly_step_exact<LATROWS>(P, R, Y, I);
                pre |= o.pre;
                if (o.reason != R_NONE && bad == NONE) bad = ((unsigned)i << 8) | (unsigned)o.reason;
                if (o.proj_fail && pbad == NONE) pbad = (unsigned)i;
                x = o.x; y = o.y; th_gl = o.th_gl; th_cl = o.th_cl; v = o.v; a = o.a; kappa = o.kappa;
                s = o.s; sv = o.sv; d = o.d; dv = o.dv;
                cn = o.cn; sn = o.sn;
                px = x; py = y;
                c_a = a; c_v = v; c_s = s; c_d = d; c_th = th_cl;
            } else {
                // ---- horizon extension (trajectories.py:168-197, :302-332) ---------------------------
                const double tau = (double)(i - tl + 1) * dt;      // np.arange(1, steps + 1) * dt
                double v_tmp = v + tau * a;
                v_tmp = v_tmp * (v_tmp >= 0 ? 1.0 : 0.0);
                const double ix = dt * v_tmp * cn, iy = dt * v_tmp * sn;
                if (i == tl) { ax = ix; ay = iy; } else { ax += ix; ay += iy; }   // np.cumsum: sequential adds
                px = x + ax; py = y + ay;
                c_a = a; c_v = v_tmp;
                c_s = s + tau * sv;                                // curvilinear tail (App. B#7)
                c_d = d + tau * dv;
                c_th = th_cl;
            }
            step_tail(i, px, py, c_a, c_v, c_s, c_d, c_th, heavy_dynmask);
        }
    }

    // ---- per-candidate verdict --------------------------------------------------------------------
    int status, reason = R_NONE, step = -1;
    double cost = __longlong_as_double(0x7ff8000000000000LL);   // NaN
    if (!valid) return;                                  // mirror lane: it only helped with the rows
    if (filtered) {
        status = ST_FILTERED;                            // filter_goals_behind (trajectories.py:545-550)
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
        status = (store_boxes && in.check_collision) ? ST_UNCHECKED : ST_FEASIBLE;
        if (costed) {
            double res_a, res_d, res_th, res_v = 0., res_s = 0.;     // np.sum results
            if (Np1 >= 8 && n8 == Np1) {                       // no remainder: the tree was not taken in the loop
                res_a = tree(0); res_d = tree(1); res_th = tree(2);
                if (use_v) res_v = tree(row_v);
                if (use_s) res_s = tree(row_s);
            } else {
                res_a = acc[0]; res_d = acc[8 * BLOCK]; res_th = acc[16 * BLOCK];
                if (use_v) res_v = acc[row_v * 8 * BLOCK];
                if (use_s) res_s = acc[row_s * 8 * BLOCK];
            }
            double costs = 0.0;
            costs += res_a;
            if (!fs && in.has_desired_speed) {
                const double e1 = v - in.desired_speed, e2 = s_vmid[0] - in.desired_speed;
                costs += res_v + (50 * (e1 * e1)) + (100 * (e2 * e2));
            }
            if (!fs && in.has_desired_s) {
                const double e = 20 * (in.desired_s - s);
                costs += res_s + e * e;
            }
            {
                const double e = 20 * (des_d - d);
                costs += res_d + e * e;
            }
            {
                const double e = 5 * fabs(th_cl);
                costs += res_th + e * e;
            }
            cost = costs;
        }
        if (col != NONE) { status = ST_COLLISION; step = (int)col; }
    }
    P.info[k] = pack_info(status, reason, step);
    if (P.cost) P.cost[k] = cost;
    // first pass of the deferred check: the feasible candidates of a pseudo-random sixteenth of the tiles
    if (DEFER && status == ST_UNCHECKED && defer_sampled(k >> 5)) defer_enlist(P, k, 0);
}

// locate chunk g of 32 candidates in the segment table (sorted by traj_len, longest first): the lane's candidate,
// clamped to the chunk's last one for lanes past the end of the segment (valid = false: mirror lanes)
__device__ __forceinline__ int chunk_candidate(const Segment* __restrict__ segs, int n_segs, int g, int lane, bool& valid) {
    int lo = 0, hi = n_segs - 1;
    while (lo < hi) {
        const int m = (lo + hi + 1) >> 1;
        if (segs[m].g_begin <= g) lo = m; else hi = m - 1;
    }
    const int k = segs[lo].k_begin + (g - segs[lo].g_begin) * 32 + lane;
    valid = k < segs[lo].k_end;
    return valid ? k : segs[lo].k_end - 1;
}

// ---- one bundle: persistent grid, warps draw chunks of 32 candidates from a counter -----------------------------
// SPLIT: grid form (every warp shares one traj_len); the list form keeps polynomial steps and extension in one loop
template <int BLOCK, bool ONE_GROUP, bool LATROWS = false, bool DEFER = false, bool SPLIT = ONE_GROUP>
__global__ void __launch_bounds__(BLOCK, RP_CAND_MIN_BLOCKS)
cand_kernel(const __grid_constant__ PlanParams P) {
    extern __shared__ double smem[];
    const int tid = threadIdx.x;

    // ---- shared memory: [reference tables (optional)] [np.sum accumulators] [v_mid] [limit reciprocals] [segments]
    double* sp = smem;
    RefTables R = P.ref;
    if (P.stage_ref) {
        const int n = R.n;
        double* base = sp;
        const double* src[9] = {R.pos, R.theta, R.curv, R.curv_d, R.px, R.py, R.nx, R.ny, R.ps};
        const int n_arr = R.same_s ? 8 : 9;
        for (int a = 0; a < n_arr; ++a)
            for (int q = tid; q < n; q += BLOCK) base[a * n + q] = src[a][q];
        R.pos = base; R.theta = base + n; R.curv = base + 2 * n; R.curv_d = base + 3 * n;
        R.px = base + 4 * n; R.py = base + 5 * n; R.nx = base + 6 * n; R.ny = base + 7 * n;
        R.ps = R.same_s ? R.pos : base + 8 * n;
        sp += n_arr * n;
    }
    double* const acc = sp + tid;                        // + (row * 8 + j) * BLOCK
    sp += (size_t)P.n_acc_rows * 8 * BLOCK;
    double* const s_vmid = sp + tid;
    sp += BLOCK;
    LimitRcp* const s_Y = reinterpret_cast<LimitRcp*>(sp);
    sp += 4;
    constexpr int SLOTS = ONE_GROUP ? RP_CAND_ONE_GROUP_SLOTS : 32;
    double* const s_rows = sp + (size_t)(tid >> 5) * warp_row_doubles(SLOTS);   // this warp's longitudinal rows
    sp += (size_t)(BLOCK / 32) * warp_row_doubles(SLOTS);
    Segment* const s_segs = reinterpret_cast<Segment*>(sp);
    for (int q = tid; q < P.n_segs; q += BLOCK) s_segs[q] = P.segs[q];
    if (tid == 0) {
        s_Y->y_dt = rcp_window(P.in.dt);
        s_Y->y_1e5 = rcp_window(100000.0);
        s_Y->y_wb = rcp_window(P.lim.wheelbase);
    }
    __syncthreads();
    const int lane = tid & 31;
    for (;;) {
        int g = 0;
        if (lane == 0) g = atomicAdd(P.work_counter, 1);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= P.n_groups) break;
        bool valid;
        int k = chunk_candidate(s_segs, P.n_segs, g, lane, valid);
        if (P.stripe_world > 1) k = Stripe{P.stripe_rank, P.stripe_world, P.n_lon, P.n_d}.real(k);     // lon-interleaved shard
        cand_march<BLOCK, ONE_GROUP, 2, LATROWS, SLOTS, DEFER, SPLIT>(P, R, *s_Y, k, valid, acc, s_vmid, s_rows);
    }
}

// ---- deferred collision check of one bundle (after cand_kernel<.., DEFER = true>; check_collision = 2) -------------
// The reference's collision pass is lazy (reactive_planner.py:1031-1063): it walks the feasible candidates in cost order
// and stops at the first collision-free one.  Here:
//   pass 1  the feasible candidates of a pseudo-random sixteenth of the tiles (listed by the march) are checked; the
//           cheapest collision-free one bounds the winner's cost (best_bits)
//   gather  every still unchecked candidate at or below the bound is listed
//   pass 2  those are checked: now every candidate ranked before the winner has its verdict, which is all the
//           reference looks at; the rest stay ST_UNCHECKED
// One BLOCK per listed tile of 32 consecutive candidates, lane = candidate, warp w = time steps w, w + 16, ...: the lanes
// of a warp are neighbours in d at the same time step -- the same cells and obstacle rows, as in the march -- and the
// N + 1 boxes of a candidate are checked by 16 warps at once instead of an (N + 1)-step march.  The verdict (status, first
// colliding step) is what the march itself would have recorded: same boxes (stored by the march), same tests.
constexpr int kDeferThreads = 512;
#ifndef RP_DEFER_MIN_BLOCKS
#define RP_DEFER_MIN_BLOCKS 2          // 64 registers: two 16-warp blocks per SM (3 / 4 blocks spill: measured slower)
#endif
// all threads of the block; s_first: 32 ints of shared memory
__device__ __forceinline__ void deferred_check_tile(const PlanParams& P, int tile, int* s_first) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const ObstacleTables& O = P.obs;
    const int Np1 = P.Np1;
    const unsigned mask = P.defer_mask[tile];
    if (threadIdx.x < 32) s_first[threadIdx.x] = 0x7fffffff;
    __syncthreads();
    if (threadIdx.x == 0) P.defer_mask[tile] = 0u;                  // clean for the next pass / cycle
    const bool active = (mask >> lane) & 1u;
    const double2* rec0 = reinterpret_cast<const double2*>(P.pose) + ((size_t)tile * Np1 * 32 + lane) * 2;
    // the boxes were written to DRAM by the march: start all of this warp's fetches now (one sector per lane and step)
    if (active)
        for (int i = warp + kDeferThreads / 32; i < Np1; i += kDeferThreads / 32)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rec0 + (size_t)i * 64));
    const float4* dyn_rows = P.dyn_rows;
    for (int i = warp; i < Np1; i += kDeferThreads / 32) {
        if (!active || reinterpret_cast<volatile int*>(s_first)[lane] < i) continue;        // (a smaller colliding step is already known)
        const double2* rec = rec0 + (size_t)i * 64;
        const double2 c = __ldcs(rec), h = __ldcs(rec + 1);
        const int tidx = P.in.x0_time_step + i * P.in.factor;
        const bool hit = (dyn_rows ? dyn_collides_f32(O, dyn_rows + (size_t)i * O.n_dyn, tidx, c.x, c.y, h.x, h.y, P.half_len, P.half_wid,
                                                      0xffffffffu)
                                   : dyn_collides_global(O, tidx, c.x, c.y, h.x, h.y, P.half_len, P.half_wid, P.r_ego)) ||
                         static_collides<2>(O, c.x, c.y, h.x, h.y, P.half_len, P.half_wid);
        if (hit) atomicMin(&s_first[lane], i);
    }
    __syncthreads();
    if (warp == 0 && active) {
        const int k = tile * 32 + lane;
        const int step = s_first[lane];
        P.info[k] = step != 0x7fffffff ? pack_info(ST_COLLISION, R_NONE, step) : pack_info(ST_FEASIBLE, R_NONE, -1);
        if (step == 0x7fffffff) atomicMin(P.best_bits, (unsigned long long)__double_as_longlong(P.cost[k]));
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kDeferThreads, RP_DEFER_MIN_BLOCKS) deferred_collision_kernel(const __grid_constant__ PlanParams P, const int* __restrict__ list,
                                                                           const int* __restrict__ n_listed) {
    __shared__ int s_first[32];
    const int n = *n_listed;
    for (int b = blockIdx.x; b < n; b += gridDim.x) deferred_check_tile(P, list[b], s_first);
}

// Second pass of a single bundle: the list names CANDIDATES (deferred_gather_kernel), 32 per block -- a listed tile has
// only a few eligible candidates, so tile-wise blocks ran half empty.  The 32-byte box records make any lane order
// sector-exact; neighbours in the list are mostly neighbours in d (the gather appends warp-wise, in order).
__global__ void __launch_bounds__(kDeferThreads, RP_DEFER_MIN_BLOCKS) deferred_collision_list_kernel(const __grid_constant__ PlanParams P,
                                                                                                    const int* __restrict__ list,
                                                                                                    const int* __restrict__ n_listed) {
    __shared__ int s_first[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = *n_listed;
    const ObstacleTables& O = P.obs;
    const int Np1 = P.Np1;
    const float4* dyn_rows = P.dyn_rows;
    for (int b = blockIdx.x; b * 32 < n; b += gridDim.x) {
        const int q = b * 32 + lane;
        const bool active = q < n;
        const int k = active ? list[q] : 0;
        if (threadIdx.x < 32) s_first[threadIdx.x] = 0x7fffffff;
        __syncthreads();
        const double2* rec0 = reinterpret_cast<const double2*>(P.pose) + ((size_t)(k >> 5) * Np1 * 32 + (k & 31)) * 2;
        if (active)
            for (int i = warp + kDeferThreads / 32; i < Np1; i += kDeferThreads / 32)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rec0 + (size_t)i * 64));
        for (int i = warp; i < Np1; i += kDeferThreads / 32) {
            if (!active || reinterpret_cast<volatile int*>(s_first)[lane] < i) continue;
            const double2* rec = rec0 + (size_t)i * 64;
            const double2 c = __ldcs(rec), h = __ldcs(rec + 1);
            const int tidx = P.in.x0_time_step + i * P.in.factor;
            const bool hit = (dyn_rows ? dyn_collides_f32(O, dyn_rows + (size_t)i * O.n_dyn, tidx, c.x, c.y, h.x, h.y, P.half_len, P.half_wid,
                                                          0xffffffffu)
                                       : dyn_collides_global(O, tidx, c.x, c.y, h.x, h.y, P.half_len, P.half_wid, P.r_ego)) ||
                             static_collides<2>(O, c.x, c.y, h.x, h.y, P.half_len, P.half_wid);
            if (hit) atomicMin(&s_first[lane], i);
        }
        __syncthreads();
        if (warp == 0 && active) {
            const int step = s_first[lane];
            P.info[k] = step != 0x7fffffff ? pack_info(ST_COLLISION, R_NONE, step) : pack_info(ST_FEASIBLE, R_NONE, -1);
            if (step == 0x7fffffff) atomicMin(P.best_bits, (unsigned long long)__double_as_longlong(P.cost[k]));
        }
        __syncthreads();
    }
}

// the same for a batch of scenarios: the lists are shared, an entry names scenario and tile
__global__ void __launch_bounds__(kDeferThreads, RP_DEFER_MIN_BLOCKS) deferred_collision_batch_kernel(const PlanParams* __restrict__ params, const int* __restrict__ list,
                                                                                 const int* __restrict__ n_listed) {
    __shared__ int s_first[32];
    const int n = *n_listed;
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        const unsigned e = (unsigned)list[b];
        deferred_check_tile(params[e >> kDeferTileBits], (int)(e & ((1u << kDeferTileBits) - 1u)), s_first);
    }
}

// second pass of a batch: the shared list names (scenario, candidate) pairs; the lanes of a block may belong to different
// scenarios (their tables, horizons and bounds are read per lane)
__global__ void __launch_bounds__(kDeferThreads, RP_DEFER_MIN_BLOCKS) deferred_collision_list_batch_kernel(const PlanParams* __restrict__ params,
                                                                                                          const int* __restrict__ list,
                                                                                                          const int* __restrict__ n_listed) {
    __shared__ int s_first[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = *n_listed;
    for (int b = blockIdx.x; b * 32 < n; b += gridDim.x) {
        const int q = b * 32 + lane;
        const bool active = q < n;
        const unsigned e = active ? (unsigned)list[q] : 0u;
        const PlanParams& P = params[e >> kDeferTileBits];
        const int k = (int)(e & ((1u << kDeferTileBits) - 1u));
        const ObstacleTables& O = P.obs;
        const int Np1 = active ? P.Np1 : 0;
        const float4* dyn_rows = P.dyn_rows;
        if (threadIdx.x < 32) s_first[threadIdx.x] = 0x7fffffff;
        __syncthreads();
        const double2* rec0 = reinterpret_cast<const double2*>(P.pose) + ((size_t)(k >> 5) * Np1 * 32 + (k & 31)) * 2;
        for (int i = warp + kDeferThreads / 32; i < Np1; i += kDeferThreads / 32)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rec0 + (size_t)i * 64));
        for (int i = warp; i < Np1; i += kDeferThreads / 32) {
            if (reinterpret_cast<volatile int*>(s_first)[lane] < i) continue;
            const double2* rec = rec0 + (size_t)i * 64;
            const double2 c = __ldcs(rec), h = __ldcs(rec + 1);
            const int tidx = P.in.x0_time_step + i * P.in.factor;
            const bool hit = (dyn_rows ? dyn_collides_f32(O, dyn_rows + (size_t)i * O.n_dyn, tidx, c.x, c.y, h.x, h.y, P.half_len, P.half_wid,
                                                          0xffffffffu)
                                       : dyn_collides_global(O, tidx, c.x, c.y, h.x, h.y, P.half_len, P.half_wid, P.r_ego)) ||
                             static_collides<2>(O, c.x, c.y, h.x, h.y, P.half_len, P.half_wid);
            if (hit) atomicMin(&s_first[lane], i);
        }
        __syncthreads();
        if (warp == 0 && active) {
            const int step = s_first[lane];
            P.info[k] = step != 0x7fffffff ? pack_info(ST_COLLISION, R_NONE, step) : pack_info(ST_FEASIBLE, R_NONE, -1);
            if (step == 0x7fffffff) atomicMin(P.best_bits, (unsigned long long)__double_as_longlong(P.cost[k]));
        }
        __syncthreads();
    }
}

// second list: unchecked candidates of the shard whose cost does not exceed the bound of the first pass -- appended warp by
// warp, in order (one atomic per warp)
__global__ void __launch_bounds__(256) deferred_gather_kernel(const __grid_constant__ PlanParams P, int first, int count) {
    const int q = (int)(blockIdx.x * (unsigned)blockDim.x + threadIdx.x);
    const int lane = threadIdx.x & 31;
    int k = 0;
    bool eligible = false;
    if (q < count) {
        k = first + q;
        if (P.stripe_world > 1) k = Stripe{P.stripe_rank, P.stripe_world, P.n_lon, P.n_d}.real(k);
        eligible = (P.info[k] & 0xFF) == ST_UNCHECKED && (unsigned long long)__double_as_longlong(P.cost[k]) <= *P.best_bits;
    }
    const unsigned m = __ballot_sync(0xffffffffu, eligible);
    if (m == 0u) return;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(P.defer_count + 1, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (eligible) P.defer_list2[base + __popc(m & ((1u << lane) - 1u))] = k;
}

__global__ void __launch_bounds__(256) deferred_gather_batch_kernel(const PlanParams* __restrict__ params) {
    const PlanParams& P = params[blockIdx.y];
    const int k = (int)(blockIdx.x * (unsigned)blockDim.x + threadIdx.x);
    const int lane = threadIdx.x & 31;
    const bool eligible = k < P.n_cand && P.pose != nullptr && (P.info[k] & 0xFF) == ST_UNCHECKED &&
                          (unsigned long long)__double_as_longlong(P.cost[k]) <= *P.best_bits;
    const unsigned m = __ballot_sync(0xffffffffu, eligible);
    if (m == 0u) return;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(P.defer_count + 1, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (eligible) P.defer_list2[base + __popc(m & ((1u << lane) - 1u))] = P.defer_tag | k;        // scenario << 20 | candidate
}

// ---- a batch of independent scenarios (BASELINE configs[4]): ONE launch over all (scenario, chunk) pairs --------
// params[sc] is the scenario's PlanParams in device memory (its own tables, samples, coefficients, verdict arrays),
// chunk_prefix[sc] the index of its first chunk in the global queue.  Scenario-static data is read through L1; the
// limit reciprocals are per warp (dt / wheelbase may differ between scenarios).
struct BatchTable {
    const PlanParams* params;
    const int* chunk_prefix;     // [n_scenarios + 1]
    int n_scenarios;
    int n_acc_rows;              // max over the scenarios (sizes shared memory)
    int* work_counter;
};

template <int BLOCK, bool DEFER = false>
__global__ void __launch_bounds__(BLOCK, RP_CAND_MIN_BLOCKS)
cand_batch_kernel(const __grid_constant__ BatchTable B) {
    extern __shared__ double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* sp = smem;
    double* const acc = sp + tid;
    sp += (size_t)B.n_acc_rows * 8 * BLOCK;
    double* const s_vmid = sp + tid;
    sp += BLOCK;
    LimitRcp* const s_Y = reinterpret_cast<LimitRcp*>(sp) + warp;
    sp += (size_t)(BLOCK / 32) * 3;
    double* const s_rows = sp + (size_t)warp * kWarpRowDoubles;
    const int total = B.chunk_prefix[B.n_scenarios];
    int sc_cached = -1;
    for (;;) {
        int g = 0;
        if (lane == 0) g = atomicAdd(B.work_counter, 1);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= total) break;
        int lo = 0, hi = B.n_scenarios - 1;              // last scenario with chunk_prefix[sc] <= g
        while (lo < hi) {
            const int m = (lo + hi + 1) >> 1;
            if (B.chunk_prefix[m] <= g) lo = m; else hi = m - 1;
        }
        const PlanParams& P = B.params[lo];
        if (lo != sc_cached) {                           // warp-uniform
            __syncwarp();
            if (lane == 0) {
                s_Y->y_dt = rcp_window(P.in.dt);
                s_Y->y_1e5 = rcp_window(100000.0);
                s_Y->y_wb = rcp_window(P.lim.wheelbase);
            }
            __syncwarp();
            sc_cached = lo;
        }
        bool valid;
        const int k = chunk_candidate(P.segs, P.n_segs, g - B.chunk_prefix[lo], lane, valid);
        cand_march<BLOCK, false, 1, false, 32, DEFER, true>(P, P.ref, *s_Y, k, valid, acc, s_vmid, s_rows);     // (batches are grid form)
    }
}

}  // namespace rp
