// rp_cand.cuh -- the candidate-major form of the fused path (a4-a13 of SURVEY.md section 8a) for LARGE bundles
// in select-only mode: one thread per candidate, marching through the horizon sequentially.
// (reactive_planner.py:715-1063, cost_function.py:51-92, trajectories.py:168-332)
//
// Why a second mapping.  fused_kernel (rp_fused.cuh) gives every (candidate, step) pair a thread; that is what a
// 120-3000 candidate replanning bundle needs to fill 148 SMs, but it pays ~12 block barriers per group, exchanges
// every sequential-in-time quantity (theta[i-1], kappa[i-1], standstill carry, cumsum, numpy's summation order)
// through shared-memory rows and re-derives indices per element: ~2200 warp instructions per candidate-timestep of
// which 27 % are FP64.  A bundle of >= ~30 000 candidates fills the machine with candidates alone, and then the
// reference's own loop structure is the cheapest one: everything sequential-in-time lives in registers, the
// reference-segment lookup advances incrementally (s is monotone up to the eps clamp), there are no barriers, and
// lanes of a warp (adjacent candidates of one sampled t) share traj_len, the dynamic obstacles of the step
// (shared-memory broadcast) and the longitudinal polynomial.
//
// The arithmetic is expression-for-expression the one of fused_kernel (same rp_device.cuh functions), so both
// kernels give identical bits; only the schedule differs.  Not handled here (the host routes these to fused_kernel):
// state output, draw mode, index mode, N + 1 > 128 (numpy's recursive pairwise split).
//
// Work distribution: the host sorts the segments (one per sampled t) by traj_len, longest first, and cuts them
// into chunks of 32 candidates; warps of a persistent grid draw chunks from a global counter.
#pragma once
#include "rp_fused.cuh"

namespace rp {

#ifndef RP_CAND_THREADS
#define RP_CAND_THREADS 128
#endif
#ifndef RP_CAND_MIN_BLOCKS
#define RP_CAND_MIN_BLOCKS 4
#endif

// cost accumulator rows kept in shared memory: acc[(row * 8 + j) * BLOCK + tid], numpy's 8 partial sums per np.sum
constexpr int kAccRowsMax = 5;

// Dynamic obstacles of every time step staged as single-precision bounding circles, rows [step][obstacle] of
// (cx, cy, squared reach, -) relative to the obstacle-table origin: all lanes of a warp are at the same step, so a row
// is a shared-memory broadcast.  The pre-reject is conservative (reach inflated by the fp32 rounding bound,
// build_obstacle_tables) and branch-free over the obstacles -- the circle tests of one step are independent
// instructions, not a serial chain; survivors go to the exact fp64 SAT, which alone decides a hit.
// An obstacle absent at a step is parked at 1e30 with zero reach (inf <= 0 is false).
__device__ __forceinline__ bool dyn_collides_f32(const ObstacleTables& O, const float4* __restrict__ row, int tidx,
                                                 double cx, double cy, double ca, double sa, double ahl, double ahw) {
    const float ex = (float)(cx - O.org_x), ey = (float)(cy - O.org_y);
    for (int base = 0; base < O.n_dyn; base += 32) {
        const int cnt = min(32, O.n_dyn - base);
        unsigned mask = 0u;
#pragma unroll 8
        for (int o = 0; o < cnt; ++o) {
            const float4 r = row[base + o];
            const float dx = r.x - ex, dy = r.y - ey;
            if (dx * dx + dy * dy <= r.z) mask |= 1u << o;
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

// first index with a[idx] > x, walking from a guess (the previous step's answer); identical to upper_bound()
__device__ __forceinline__ int upper_bound_from(const double* __restrict__ a, int n, double x, int j) {
    while (j < n && a[j] <= x) ++j;
    while (j > 0 && a[j - 1] > x) --j;
    return j;
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, RP_CAND_MIN_BLOCKS)
cand_kernel(const __grid_constant__ PlanParams P) {
    extern __shared__ double smem[];
    const int Np1 = P.Np1;
    const int tid = threadIdx.x;

    // ---- shared memory carve-up -----------------------------------------------------------------
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
    const ObstacleTables& O = P.obs;
    const float4* dyn_stage = nullptr;
    if (P.stage_dyn && O.n_dyn > 0) {
        if (reinterpret_cast<uintptr_t>(sp) & 8u) ++sp;  // 16-byte rows (the host's size includes the slack)
        float4* dst = reinterpret_cast<float4*>(sp);
        const int total = Np1 * O.n_dyn;
        for (int q = tid; q < total; q += BLOCK) {
            const int step = q / O.n_dyn, o = q - step * O.n_dyn;
            const int kk = P.in.x0_time_step + step * P.in.factor - O.dyn_t0[o];
            const bool present = kk >= 0 && kk < O.dyn_len[o];
            float4 r = make_float4(1.0e30f, 1.0e30f, 0.0f, 0.0f);
            if (present) {
                const double* b = O.dyn_box + (size_t)(O.dyn_off[o] + kk) * kBoxStride;
                const float reach = (P.r_ego_f_up + (float)b[6]) * 1.000001f + O.dyn_margin;   // >= r_ego + r_obs + margin
                r = make_float4((float)(b[0] - O.org_x), (float)(b[1] - O.org_y), reach * reach * 1.00001f, 0.0f);
            }
            dst[q] = r;
        }
        dyn_stage = dst;
        sp += (size_t)total * 2;                          // float4 = 2 doubles
    }
    double* const acc = sp + tid;                        // + (row * 8 + j) * BLOCK
    sp += (size_t)P.n_acc_rows * 8 * BLOCK;
    Segment* const s_segs = reinterpret_cast<Segment*>(sp);
    for (int q = tid; q < P.n_segs; q += BLOCK) s_segs[q] = P.segs[q];
    __syncthreads();

    const rp_plan_inputs& in = P.in;
    const bool low_vel = in.low_vel_mode != 0;
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
    const int lane = tid & 31;

    for (;;) {
        int g = 0;
        if (lane == 0) g = atomicAdd(P.work_counter, 1);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= P.n_groups) break;
        int lo = 0, hi = P.n_segs - 1;
        while (lo < hi) {
            const int m = (lo + hi + 1) >> 1;
            if (s_segs[m].g_begin <= g) lo = m; else hi = m - 1;
        }
        const int k = s_segs[lo].k_begin + (g - s_segs[lo].g_begin) * 32 + lane;
        if (k >= s_segs[lo].k_end) continue;

        // ---- candidate decode (sampling.py:202-242 enumeration order) ----------------------------
        double cs[6], cd[6];
        int tl;
        bool filtered;
        {
            const double *pl, *pt;
            if (P.mode == 0) {
                const int per_t = P.n_lon * P.n_d;
                const int it = k / per_t;
                const int rem = k - it * per_t;
                const int il = rem / P.n_d;
                const int id = rem - il * P.n_d;
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
        if (filtered) {
            P.info[k] = pack_info(ST_FILTERED, R_NONE, -1);
            if (P.cost) P.cost[k] = __longlong_as_double(0x7ff8000000000000LL);
            continue;
        }

        const LimitRcp Y = {rcp_refined(dt), rcp_refined(100000.0), rcp_refined(P.lim.wheelbase)};
        unsigned pre = 0u, bad = NONE, pbad = NONE, col = NONE;
        int ub = -1;
        // values of the current / last polynomial step (the extension reads them after step tl - 1)
        double x = 0., y = 0., th_gl = 0., v = 0., a = 0., kappa = 0., s = 0., sv = 0., d = 0., dv = 0., th_cl = 0.;
        double cn = 1., sn = 0.;                       // cos / sin of th_gl
        double th_prev = 0., kap_prev = 0.;
        double ax = 0., ay = 0.;                       // np.cumsum of the extension increments
        double res_a = 0., res_d = 0., res_th = 0., res_v = 0., res_s = 0.;   // np.sum results
        double v_mid = 0.;

        for (int i = 0; i < Np1; ++i) {
            double px, py;                             // rear-axle position of this step
            double c_a, c_v, c_s, c_d, c_th;           // values entering the cost terms
            if (i < tl) {
                // ---- polynomial evaluation (reactive_planner.py:733-777) ---------------------------
                const double tt = (double)i * dt;
                const double t2 = tt * tt, t3 = t2 * tt, t4 = t2 * t2, t5 = t4 * tt;
                s = poly_pos(cs, tt, t2, t3, t4, t5);
                sv = poly_vel(cs, tt, t2, t3, t4);
                const double sa = poly_acc(cs, tt, t2, t3);
                double da;
                if (!low_vel) {
                    d = poly_pos(cd, tt, t2, t3, t4, t5);
                    dv = poly_vel(cd, tt, t2, t3, t4);
                    da = poly_acc(cd, tt, t2, t3);
                } else {
                    const double s1 = s - cs[0];
                    const double s2 = s1 * s1, s3 = s2 * s1, s4 = s2 * s2, s5 = s4 * s1;
                    d = poly_pos(cd, s1, s2, s3, s4, s5);
                    dv = poly_vel(cd, s1, s2, s3, s4);
                    da = poly_acc(cd, s1, s2, s3);
                }
                if (fabs(sv) < kEps) sv = 0.0;
                if (fabs(dv) < kEps) dv = 0.0;
                if (fabs(sa) > P.lim.a_max) pre |= 1u;   // pre-filter (:796-805); this kernel never runs draw mode
                if (sv < -kEps) pre |= 2u;

                // ---- orientation (:810-873) --------------------------------------------------------
                // every division below is IEEE a / b (div_rcp, rp_device.cuh); divisors that serve several
                // quotients get ONE refined reciprocal
                double dp, dpp;
                if (!low_vel) {
                    if (sv > 0.001) {
                        dp = div_rcp(dv, sv, rcp_refined(sv));
                        const double ddot = da - dp * sa;
                        const double sv2 = sv * sv;
                        dpp = div_rcp(ddot, sv2, rcp_refined(sv2));
                    } else {
                        dp = 0.;
                        dpp = 0.;
                    }
                } else {
                    dp = dv;
                    dpp = da;
                }
                ub = ub < 0 ? upper_bound_guess(R.pos, R.n, s, P.ref_inv_step) : upper_bound_from(R.pos, R.n, s, ub);
                const bool wrap = (ub == R.n) || (ub == 0);      // s_idx == -1: python index wrap (App. B#8)
                const int j0 = wrap ? R.n - 1 : ub - 1;
                const int j1 = wrap ? 0 : ub;
                const double p0 = R.pos[j0], p1 = R.pos[j1];
                const double seg_len = p1 - p0;
                const double y_seg = rcp_refined(seg_len);
                const double lam = div_rcp(s - p0, seg_len, y_seg);
                // interpolate_angle (utility/utils_coordinate_system.py:25-43)
                const double th0 = R.theta[j0];
                const double th_ref = make_valid_orientation(div_rcp((R.theta[j1] - th0) * (s - p0), seg_len, y_seg) + th0);
                const bool carry = !(sv > 0.001) && !low_vel;
                if (!carry) {
                    th_cl = atan(dp);                            // np.arctan2(dp, 1.0)
                    th_gl = th_cl + th_ref;
                } else {
                    // standstill in high-velocity mode keeps the previous global orientation (:866-873)
                    th_gl = i > 0 ? th_prev : in.x0_orientation;
                    th_cl = th_gl - th_ref;
                }

                // ---- curvature, velocity, acceleration (:876-896) -----------------------------------
                const double k0 = R.curv[j0], kd0 = R.curv_d[j0];
                const double k_r = (R.curv[j1] - k0) * lam + k0;
                const double k_r_d = (R.curv_d[j1] - kd0) * lam + kd0;
                const double oneKrD = (1 - k_r * d);
                double cosT, tanT;
                if (!carry) {
                    const double hyp = sqrt(1.0 + dp * dp);
                    cosT = div_rcp(1.0, hyp, rcp_refined(hyp));
                    tanT = dp;
                } else {
                    cosT = cos(th_cl);
                    tanT = tan(th_cl);
                }
                const double y_cos = rcp_refined(cosT);
                const double q = div_rcp(cosT, oneKrD, rcp_refined(oneKrD));
                kappa = (dpp + (k_r * dp + k_r_d * d) * tanT) * cosT * (q * q) + q * k_r;
                v = sv * div_rcp(oneKrD, cosT, y_cos);
                a = div_rcp(sa * oneKrD, cosT, y_cos) +
                    div_rcp(sv * sv, cosT, y_cos) * (oneKrD * tanT * (div_rcp(kappa * oneKrD, cosT, y_cos) - k_r) -
                                                     (k_r_d * d + k_r * dp));

                // ---- limits (:971-1017) + projection (:908-917) --------------------------------------
                const int r = check_constraints_rcp(P.lim, Y, in.constraint_mask, dt, i, v, kappa, kap_prev, th_gl, th_prev, a);
                if (r != R_NONE && bad == NONE) bad = ((unsigned)i << 8) | (unsigned)r;
                {
                    // project_to_cartesian (rp_device.cuh) with the segment reciprocal shared when both tables coincide
                    const int n = R.n;
                    bool ok = (s >= R.ps[0] && s <= R.ps[n - 1]) && (fabs(d) <= R.limit);
                    if (ok) {
                        const int ub_ps = R.same_s ? ub : upper_bound_guess(R.ps, n, s, P.ps_inv_step);
                        int j = ub_ps - 1;
                        if (j > n - 2) j = n - 2;
                        double lam2;
                        if (R.same_s && j == j0 && j1 == j0 + 1) {
                            lam2 = lam;                          // same dividend, same divisor
                        } else {
                            const double sl = R.ps[j + 1] - R.ps[j];
                            lam2 = div_rcp(s - R.ps[j], sl, rcp_refined(sl));
                        }
                        const double p0x = R.px[j], p0y = R.py[j];
                        const double bx = p0x + lam2 * (R.px[j + 1] - p0x);
                        const double by = p0y + lam2 * (R.py[j + 1] - p0y);
                        const double n0x = R.nx[j], n0y = R.ny[j];
                        const double nx = n0x + lam2 * (R.nx[j + 1] - n0x);
                        const double ny = n0y + lam2 * (R.ny[j + 1] - n0y);
                        x = bx + d * nx;
                        y = by + d * ny;
                    } else {
                        if (pbad == NONE) pbad = (unsigned)i;
                        x = 0.; y = 0.;
                    }
                }
                th_prev = th_gl;
                kap_prev = kappa;
                if (in.check_collision || i == tl - 1) sincos(th_gl, &sn, &cn);
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

            // ---- cost terms in numpy's np.sum order (cost_function.py:51-71) ---------
            if (costed) {
                const double t0 = w_a * c_a, t3 = 0.25 * (des_d - c_d), t4 = 0.25 * fabs(c_th);
                double q0 = t0 * t0, q3 = t3 * t3, q4 = t4 * t4, q1 = 0., q2 = 0.;
                if (use_v) { const double t1 = 5 * (c_v - in.desired_speed); q1 = t1 * t1; }
                if (use_s) { const double t2 = 0.25 * (in.desired_s - c_s); q2 = t2 * t2; }
                if (Np1 < 8) {
                    res_a += q0; res_d += q3; res_th += q4; res_v += q1; res_s += q2;
                } else if (i < n8) {
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
                    if (i == n8) {
                        auto tree = [&](int row) {
                            const double* r8 = acc + (size_t)row * 8 * BLOCK;
                            return ((r8[0] + r8[BLOCK]) + (r8[2 * BLOCK] + r8[3 * BLOCK])) +
                                   ((r8[4 * BLOCK] + r8[5 * BLOCK]) + (r8[6 * BLOCK] + r8[7 * BLOCK]));
                        };
                        res_a = tree(0); res_d = tree(1); res_th = tree(2);
                        if (use_v) res_v = tree(row_v);
                        if (use_s) res_s = tree(row_s);
                    }
                    res_a += q0; res_d += q3; res_th += q4; res_v += q1; res_s += q2;
                }
                if (i == mid) v_mid = c_v;                       // v[int(len(v) / 2)]
            }

            // ---- ego-vs-obstacle check (reactive_planner.py:1026-1046), speculative ---------------------
            if (in.check_collision && col == NONE && bad == NONE && pbad == NONE && pre == 0u) {
                const double ecx = px + P.wb_rear * cn;
                const double ecy = py + P.wb_rear * sn;
                const int tidx = in.x0_time_step + i * in.factor;
                const bool hit = (dyn_stage ? dyn_collides_f32(O, dyn_stage + (size_t)i * O.n_dyn, tidx, ecx, ecy, cn, sn, P.half_len, P.half_wid)
                                            : dyn_collides_global(O, tidx, ecx, ecy, cn, sn, P.half_len, P.half_wid, P.r_ego)) ||
                                 static_collides(O, ecx, ecy, cn, sn, P.half_len, P.half_wid);
                if (hit) col = (unsigned)i;
            }
            if (i == Np1 - 1) {                                   // park the end values for the terminal terms
                v = c_v; s = c_s; d = c_d; th_cl = c_th;
            }
        }

        // ---- per-candidate verdict --------------------------------------------------------------------
        int status, reason = R_NONE, step = -1;
        double cost = __longlong_as_double(0x7ff8000000000000LL);   // NaN
        if (pre != 0u) {
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
            if (costed) {
                if (Np1 >= 8 && n8 == Np1) {                       // no remainder: the tree was not taken in the loop
                    auto tree = [&](int row) {
                        const double* r8 = acc + (size_t)row * 8 * BLOCK;
                        return ((r8[0] + r8[BLOCK]) + (r8[2 * BLOCK] + r8[3 * BLOCK])) +
                               ((r8[4 * BLOCK] + r8[5 * BLOCK]) + (r8[6 * BLOCK] + r8[7 * BLOCK]));
                    };
                    res_a = tree(0); res_d = tree(1); res_th = tree(2);
                    if (use_v) res_v = tree(row_v);
                    if (use_s) res_s = tree(row_s);
                }
                double costs = 0.0;
                costs += res_a;
                if (!fs && in.has_desired_speed) {
                    const double e1 = v - in.desired_speed, e2 = v_mid - in.desired_speed;
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
    }
}

}  // namespace rp
