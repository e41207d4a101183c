// rp_fused.cuh -- the fused per-(candidate, time step) kernel (a4-a13 of SURVEY.md section 8a):
// polynomial evaluation, reference-segment lookup in shared memory, Frenet->Cartesian, kinematic limits,
// horizon extension, cost (numpy summation order) and the ego-vs-obstacle SAT with the dynamic obstacles
// of each time step staged in shared memory.
// (reactive_planner.py:715-1063, cost_function.py:51-92, trajectories.py:168-332)
//
// Thread mapping.  The heavy phases (evaluation, orientation, curvature, limits, projection) only exist
// for steps i < traj_len, so a block maps its threads to (slot c, step i < tl) pairs: C = 256 / tl
// candidates per group.  Candidates are enumerated t-major, hence the candidates of a group share tl;
// the host describes the work as SEGMENTS of equal tl (one per sampled t).  Everything that covers the
// whole horizon (extension tail, cost terms, collision, state rows) runs as element-strided loops over
// the C x (N+1) (slot, step) pairs of the group, fed from shared-memory rows.  Blocks are persistent over
// groups so that the reference tables and the dynamic-obstacle rows are staged once.
#pragma once
#include "rp_device.cuh"
#include "rp_b200.h"

namespace rp {

struct Segment {
    int k_begin, k_end;   // candidate ids (or slots in index mode) [k_begin, k_end)
    int tl;               // threads per candidate in the heavy phases (= traj_len in grid mode, N+1 otherwise)
    int C;                // candidates per group
    int g_begin;          // index of the segment's first group
    int level;            // cycle mode (several sampling levels in one launch): index of the segment's level
};

// ---- one replanning cycle in ONE launch (rp_plan_levels) -------------------------------------------------------
// The level-escalation loop of reactive_planner.py:616-636 evaluates sampling levels 1, 2, 3 one after the other and
// stops at the first that yields a feasible, collision-free candidate; the levels are independent bundles (SURVEY
// App. B#13).  cycle_kernel evaluates several levels as ONE concatenated enumeration space; everything a cycle needs
// travels in the launch itself (this struct is a __grid_constant__ kernel parameter: no host->device copy, no
// coefficient launch -- each candidate slot solves its two polynomials in shared memory), and the last block to finish
// selects level by level, gathers the winner's state block and writes the records straight into mapped host memory.
constexpr int kMaxLevels = 4;
constexpr int kCycleSamples = 448;      // doubles: t | lon | d of every level
constexpr int kCycleSegs = 128;         // one segment per (level, sampled t)

struct LevelDesc {
    int k0, count;                      // the level's slice of the concatenated enumeration space
    int n_t, n_lon, n_d;
    int off_t, off_lon, off_d;          // offsets into CycleArgs::samples
};

// (the parameter block rides in the launch packet: 7.4 KB cost a null launch 2.8 us more than 64 bytes on this part, so
// cycles that fit -- e.g. levels 1-3 of the bundled N = 20 configurations -- use the SMALL instantiation)
constexpr int kCycleSamplesSmall = 160;
constexpr int kCycleSegsSmall = 32;

template <int NSEG, int NSAMP>
struct CycleArgsT {
    int n_levels, n_segs;
    LevelDesc lv[kMaxLevels];
    Segment segs[NSEG];
    double samples[NSAMP];
    void* out_host;                     // CycleOut in mapped pinned host memory (rp_kernels.cuh)
    unsigned* ticket;                   // blocks that have finished (reset by the last one)
    unsigned long long epoch;           // written to CycleOut::flag when the records are complete
};
using CycleArgs = CycleArgsT<kCycleSegs, kCycleSamples>;
using CycleArgsSmall = CycleArgsT<kCycleSegsSmall, kCycleSamplesSmall>;
struct NoCycleArgs {                    // fused_kernel: never dereferenced
    Segment segs[1];
    LevelDesc lv[1];
    double samples[1];
};

struct PlanParams {
    // ---- candidate source ----
    int mode;                    // 0: grid (t x lon x d), 1: list (per-candidate coefficients)
    int n_t, n_lon, n_d;
    int n_cand;                  // size of the enumeration space
    const int* index;            // index mode: candidate id per slot (negative = empty slot); else null
    const double* lon_samples;   // grid: [n_lon] (filter_goals_behind in stopping mode)
    const double* t_samples;     // grid: [n_t], [n_d]: read by the batch coefficient kernel only
    const double* d_samples;
    const int* traj_len;         // grid: [n_t], list: [n_cand]
    const double* lon_coef;      // grid: [n_t*n_lon][6], list: [n_cand][6]
    const double* lat_coef;      // grid: [n_t*n_d][6] (or [n_cand][6] in low-velocity mode), list: [n_cand][6]
    const uint8_t* skip;         // list: filter_goals_behind flags (may be null)
    // ---- work decomposition ----
    const Segment* segs;
    int n_segs, n_groups, Cmax;
    // ---- per-cycle scalars / tables ----
    rp_plan_inputs in;
    Limits lim;
    double half_len, half_wid, wb_rear, r_ego;
    float r_ego_f_up;            // r_ego rounded up to fp32 (single-precision pre-reject of rp_cand.cuh)
    float wb_rear_f_up;          // |wb_rear| rounded up to fp32
    RefTables ref;
    double ref_inv_step, ps_inv_step;     // (n-1) / (last - first): index guess of the segment lookup
    ObstacleTables obs;
    // ---- outputs ----
    double* cost;                // [n_cand]   (null: not written)
    int* info;                   // [n_cand]   status | reason << 8 | (step + 1) << 16
    double* states;              // [.][14][N+1] indexed by candidate (states_by_slot = 0) or slot
    int states_by_slot;
    unsigned long long* best_bits;   // lazy collision mode: bit pattern of the best collision-free cost so far
    int Np1;                     // N + 1
    int stage_ref, stage_dyn;
    // ---- candidate-major kernel (rp_cand.cuh) ----
    int* work_counter;           // chunk dispenser (zeroed before the launch)
    int n_acc_rows;              // np.sum accumulator rows kept in shared memory
    const float4* dyn_rows;      // [Np1][n_dyn] single-precision circles of the launch's time window (null: none staged)
    // lon-interleaved shard of a grid bundle (rp_set_candidate_stripe; world <= 1: off): the launch enumerates the shard's
    // candidates compactly ("virtual" index) and stripe_to_real() maps them into the bundle's enumeration space
    int stripe_rank, stripe_world;
    const double* lat_rows;      // [n_t][Np1][n_d][4] = d, d_dot (clamped), d_ddot, - of the lateral polynomials on the time
                                 // grid (high-velocity grid bundles: shared by all lon samples; null: evaluated per candidate)
    // deferred collision check of the candidate-major kernel (deferred_collision_kernel): the march stores the ego box of
    // every step as a 32-byte record (centre x, centre y, cos, sin) at [(k >> 5)][step][k & 31] -- the warp's store of
    // one step is 1 KB contiguous, the checker's read of one (candidate, step) is exactly one sector
    double* pose;
    int* defer_list;             // tiles (32 consecutive candidates, k >> 5) awaiting the deferred check, first pass ...
    int* defer_list2;            // ... and second pass; entries are defer_tag | tile (a scenario batch shares the lists)
    int* defer_count;            // [2] lengths of the two lists (zeroed with the work counter)
    unsigned* defer_mask;        // [n_tiles] which candidates of a listed tile are to be checked (cleared by the checker)
    int defer_tag;               // scenario index << kDeferTileBits in a batch, 0 for a single bundle
};
constexpr int kDeferTileBits = 20;         // tiles per bundle < 2^20 (33 M candidates), scenarios per batch < 2^11

// Multi-GPU shards that are ALIKE instead of contiguous: rank r of `world` owns the lon samples il = r, r + world, ... of
// EVERY sampled t, so all ranks see the same mix of traj_len (contiguous t-major tiles differ in it and the exchange waits
// for the slowest rank).  Virtual index kv = (it * n_lon_r + jl) * n_d + id, il = rank + jl * world.
struct Stripe {
    int rank, world, n_lon, n_d;
    __device__ __forceinline__ int real(int kv) const {
        if (world <= 1) return kv;
        const int n_lon_r = (n_lon - rank + world - 1) / world;
        const int per_t_r = n_lon_r * n_d;
        const int it = kv / per_t_r;
        const int rem = kv - it * per_t_r;
        const int jl = rem / n_d;
        return (it * n_lon + rank + jl * world) * n_d + (rem - jl * n_d);
    }
};

__device__ __forceinline__ int pack_info(int status, int reason, int step) {
    return status | (reason << 8) | ((step + 1) << 16);
}

// first index with a[idx] > x: O(1) guess on the (nearly) uniform table + exact fix-up, binary search
// if the table turns out not to be uniform.  Identical result to upper_bound().
__device__ __forceinline__ int upper_bound_guess(const double* __restrict__ a, int n, double x, double inv_step) {
    const double f = (x - a[0]) * inv_step;
    int j = f > 0.0 ? (f < (double)n ? (int)f : n) : 0;
    int moved = 0;
    while (j < n && a[j] <= x) { ++j; if (++moved > 6) return upper_bound(a, n, x); }
    while (j > 0 && a[j - 1] > x) { --j; if (++moved > 6) return upper_bound(a, n, x); }
    return j;
}

// per-slot integer scratch
enum : int { F_BAD = 0, F_PBAD = 1, F_PRE = 2, F_COL = 3, F_TL = 4, F_STATE = 5, F_K = 6, F_GATE = 7, F_WORDS = 8 };
// F_STATE bits
enum : unsigned { S_VALID = 1u, S_FILTERED = 2u, S_ALIVE = 4u, S_KINOK = 8u, S_KEEP = 16u };
// per-slot double scratch beyond the rows: last[16] acc[40] sums[5] park[5] (+pad)
constexpr int kRows = 11;         // th, kap, x, y, tx, ty, ct0..ct4
constexpr int kSlotExtra = 16 + 40 + 5 + 5 + 2;

// development aid (-DRP_CYCLE_TIMING, tools/probe_cycle.py): cycle stamps of block 0 at the phase boundaries
#ifdef RP_CYCLE_TIMING
__device__ long long g_stamps[32];
__device__ long long g_block_t[3 * 1024];       // per block: entry, body end, smid (globaltimer ns)
__device__ __forceinline__ long long rp_globaltimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define RP_STAMP(n) do { if (blockIdx.x == 0 && threadIdx.x == 0) { g_stamps[n] = clock64(); if ((n) == 0 || (n) == 11) g_stamps[16 + (n)] = rp_globaltimer(); } } while (0)
#else
#define RP_STAMP(n) do { } while (0)
#endif

#ifndef RP_FUSED_MIN_BLOCKS
#define RP_FUSED_MIN_BLOCKS 3
#endif

// CYCLE = true: several sampling levels in one launch (cycle_kernel): candidates are decoded through the segment's
// level, and the two polynomials of a slot are solved HERE (threads 0 / 1 of the slot, into the slot's still unused
// accumulator scratch) with the same device functions coeff_kernel uses -- identical bits, no coefficient launch.
template <int MAXT, bool CYCLE, class ARGS>
__device__ __forceinline__ void fused_body(const PlanParams& P, const ARGS* __restrict__ A, double* smem) {
    const int Np1 = P.Np1;
    const int Cmax = P.Cmax;
    const int tid = threadIdx.x;
    const int T = blockDim.x;

    // ---- shared memory carve-up -----------------------------------------------------------------
    double* sp = smem;
    RefTables R = P.ref;
    if (P.stage_ref) {
        const int n = R.n;
        double* base = sp;
        const double* src[9] = {R.pos, R.theta, R.curv, R.curv_d, R.px, R.py, R.nx, R.ny, R.ps};
        const int n_arr = R.same_s ? 8 : 9;
        for (int a = 0; a < n_arr; ++a)
            for (int q = tid; q < n; q += T) base[a * n + q] = src[a][q];
        R.pos = base; R.theta = base + n; R.curv = base + 2 * n; R.curv_d = base + 3 * n;
        R.px = base + 4 * n; R.py = base + 5 * n; R.nx = base + 6 * n; R.ny = base + 7 * n;
        R.ps = R.same_s ? R.pos : base + 8 * n;
        sp += n_arr * n;
    }
    const ObstacleTables& O = P.obs;
    const double* dyn_stage = nullptr;
    if (P.stage_dyn && O.n_dyn > 0) {
        // dynamic obstacles present at time index x0.time_step + i * factor (reactive_planner.py:1040),
        // laid out [obstacle][field][step]
        double* dst = sp;
        const int total = Np1 * O.n_dyn;
        for (int q = tid; q < total; q += T) {
            const int o = q / Np1, step = q - o * Np1;
            const int k = P.in.x0_time_step + step * P.in.factor - O.dyn_t0[o];
            const bool present = k >= 0 && k < O.dyn_len[o];
            const double* b = O.dyn_box + (size_t)(O.dyn_off[o] + (present ? k : 0)) * kBoxStride;
            double* row = dst + (size_t)o * kDynFields * Np1 + step;
            row[0] = present ? b[0] : 1.0e300;               // absent: parked far away, the circle reject drops it
            row[Np1] = present ? b[1] : 1.0e300;
            row[2 * Np1] = present ? reach2(P.r_ego, b[6]) : 0.0;
        }
        dyn_stage = dst;
        sp += (size_t)total * kDynFields;
    }
    const int per_slot = kRows * Np1 + kSlotExtra;
    double* const slots = sp;
    sp += (size_t)Cmax * per_slot;
    int* const s_carry_all = reinterpret_cast<int*>(sp);                    // [Cmax][Np1]
    unsigned* const s_flags_all = reinterpret_cast<unsigned*>(s_carry_all + (size_t)Cmax * Np1);   // [Cmax][F_WORDS]
    Segment* const s_segs = reinterpret_cast<Segment*>(s_flags_all + (size_t)Cmax * F_WORDS);
    for (int q = tid; q < P.n_segs; q += T) s_segs[q] = CYCLE ? A->segs[q] : P.segs[q];

    const rp_plan_inputs& in = P.in;
    const bool low_vel = in.low_vel_mode != 0;
    const bool draw = in.draw_all != 0;
    const double dt = in.dt;
    const unsigned NONE = 0xFFFFFFFFu;
    const bool fs = in.cost_kind == RP_COST_FAILSAFE;
    const double w_a = fs ? 1.0 : in.w_a;
    const double des_d = fs ? 0.0 : in.desired_d;
    // cycle launches in select-only mode: only the BEST candidate of each group (feasible, collision-free, lowest
    // (cost, index)) writes its state block -- the cycle's winner is the best of its own group, and the stores of all
    // the other kept candidates (2.3 KB each at N = 20) were what the publish phase and the closing fence waited for
    const bool defer_states = CYCLE && P.states != nullptr && !in.want_all_states;
    __shared__ unsigned long long s_gbest;
    __shared__ int s_gk;
    // element-strided loops over (slot, step) pairs advance without integer division
    const int e_c0 = tid / Np1, e_i0 = tid - e_c0 * Np1;
    const int e_dc = T / Np1, e_di = T - e_dc * Np1;
    __syncthreads();
    RP_STAMP(1);

    for (int g = blockIdx.x; g < P.n_groups; g += gridDim.x) {
        // ---- locate the group's segment (block-uniform) ---------------------------------------------
        int lo = 0, hi = P.n_segs - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (s_segs[mid].g_begin <= g) lo = mid; else hi = mid - 1;
        }
        const Segment seg = s_segs[lo];
        // lazy collision bound: one per sampling level (levels are independent bundles)
        unsigned long long* const best_ptr = P.best_bits ? P.best_bits + (CYCLE ? seg.level : 0) : nullptr;
        const int C = seg.C;
        const int tlg = seg.tl;
        const int c = tid / tlg;
        const int i = tid - c * tlg;
        const bool lane = c < C;
        const int slot = seg.k_begin + (g - seg.g_begin) * C + c;
        bool valid = lane && slot < seg.k_end;
        int k = -1;
        if (valid) {
            k = P.index ? P.index[slot] : (P.stripe_world > 1 ? Stripe{P.stripe_rank, P.stripe_world, P.n_lon, P.n_d}.real(slot) : slot);
            if (k < 0 || k >= P.n_cand) valid = false;
        }
        double* const scratch = slots + (size_t)(lane ? c : 0) * per_slot;
        double* const s_th = scratch;
        double* const s_kap = scratch + Np1;
        double* const s_x = scratch + 2 * Np1;
        double* const s_y = scratch + 3 * Np1;
        double* const s_ct = scratch + 6 * Np1;               // rows 6..10
        double* const s_last = scratch + kRows * Np1;         // 14 states at step tl-1, cos/sin of its heading
        double* const s_park = s_last + 16 + 40 + 5;          // v[-1], s[-1], d[-1], theta_cl[-1], v[mid]
        int* const s_carry = s_carry_all + (size_t)(lane ? c : 0) * Np1;
        unsigned* const s_flags = s_flags_all + (size_t)(lane ? c : 0) * F_WORDS;

        double cs[6], cd[6];
        int tl = 0;
        bool filtered = false;
        if (CYCLE) {
            double* const s_coef = scratch + kRows * Np1 + 16;       // the slot's accumulator scratch (cost phase only)
            if (valid) {
                const LevelDesc& L = A->lv[seg.level];
                const int kk = k - L.k0;
                const int per_t = L.n_lon * L.n_d;
                const int it = kk / per_t;
                const int rem = kk - it * per_t;
                const int il = rem / L.n_d;
                const int id = rem - il * L.n_d;
                tl = tlg;                                            // one segment per (level, sampled t)
                const double tt = A->samples[L.off_t + it], lon = A->samples[L.off_lon + il];
                filtered = (in.lon_mode == RP_STOPPING) && !(in.x0_lon[0] < lon);
                if (i == 0) {
                    // a3 (polynomial_trajectory.py:292-360), exactly as coeff_thread does it.  One thread solves both
                    // systems: straight-line code, so the two independent elimination chains overlap (on two lanes of
                    // one warp they would run one after the other)
                    double cl[6], ct[6];
                    if (in.lon_mode == RP_VELOCITY_KEEPING) solve_quartic(in.x0_lon[0], in.x0_lon[1], in.x0_lon[2], tt, lon, cl);
                    else solve_quintic(in.x0_lon[0], in.x0_lon[1], in.x0_lon[2], lon, 0.0, 0.0, tt, cl);
                    double tau = tt;
                    if (low_vel) {
                        double s_goal = position_at_end(cl, tt) - in.x0_lon[0];
                        if (s_goal <= 0) s_goal = tt;
                        tau = s_goal;
                    }
                    solve_quintic(in.x0_lat[0], in.x0_lat[1], in.x0_lat[2], A->samples[L.off_d + id], 0.0, 0.0, tau, ct);
#pragma unroll
                    for (int q = 0; q < 6; ++q) { s_coef[q] = cl[q]; s_coef[6 + q] = ct[q]; }
                }
            }
            __syncthreads();                                         // coefficients of every slot solved
        RP_STAMP(2);
            if (valid) {
#pragma unroll
                for (int q = 0; q < 6; ++q) { cs[q] = s_coef[q]; cd[q] = s_coef[6 + q]; }
            }
        } else if (valid) {
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
            if (tl > tlg) tl = tlg;
        }
        if (lane && i == 0) {
            s_flags[F_BAD] = NONE; s_flags[F_PBAD] = NONE; s_flags[F_PRE] = 0u; s_flags[F_COL] = NONE;
            s_flags[F_TL] = (unsigned)tl; s_flags[F_K] = (unsigned)k;
            s_flags[F_STATE] = (valid ? S_VALID : 0u) | (filtered ? S_FILTERED : 0u);
        }
        const bool live = valid && !filtered;
        const bool in_traj = live && i < tl;
        if (tid == 0) { s_gbest = ~0ULL; s_gk = 0x7fffffff; }
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
                if (bits) atomicOr(&s_flags[F_PRE], bits);
            }
        }
        __syncthreads();                                                            // pre-filter known
        RP_STAMP(3);
        const unsigned pre = lane ? s_flags[F_PRE] : 0u;
        const bool alive = live && pre == 0u;
        const bool act = alive && i < tl;

        // ---- orientation (reactive_planner.py:810-873) -------------------------------------------
        double dp = 0., dpp = 0., lam = 0., th_ref = 0., th_cl = 0., th_gl = 0.;
        int j0 = 0, j1 = 0, ub = 0;
        bool carry = false;
        if (act) {
            if (!low_vel) {
                if (sv > 0.001) dp = ddiv(dv, sv); else dp = 0.;
                const double ddot = da - dp * sa;
                if (sv > 0.001) dpp = ddiv(ddot, sv * sv); else dpp = 0.;
            } else {
                dp = dv;
                dpp = da;
            }
            ub = upper_bound_guess(R.pos, R.n, s, P.ref_inv_step);
            const bool wrap = (ub == R.n) || (ub == 0);          // s_idx == -1: python index wrap (App. B#8)
            j0 = wrap ? R.n - 1 : ub - 1;
            j1 = wrap ? 0 : ub;
            const double p0 = R.pos[j0], p1 = R.pos[j1];
            lam = ddiv(s - p0, p1 - p0);
            th_ref = interpolate_angle(s, p0, p1, R.theta[j0], R.theta[j1]);
            carry = !(sv > 0.001) && !low_vel;
            if (!carry) {
                th_cl = rp_atan(dp);                                // np.arctan2(dp, 1.0)
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
        double kappa = 0., v = 0., a = 0., cn = 1., sn = 0.;
        if (act) {
            const double k0 = R.curv[j0], kd0 = R.curv_d[j0];
            const double k_r = (R.curv[j1] - k0) * lam + k0;
            const double k_r_d = (R.curv_d[j1] - kd0) * lam + kd0;
            double cosT;
            if (!carry) {
                // shared with the candidate-major kernel (rp_device.cuh motion_moving): shared-reciprocal quotients first,
                // plain divisions if an operand left their proven window -- identical bits either way
                Divider<false> D;
                motion_moving(D, dp, dpp, d, k_r, k_r_d, sv, sa, cosT, kappa, v, a);
                if (D.reject & 0x80000000u) {
                    Divider<true> E;
                    motion_moving(E, dp, dpp, d, k_r, k_r_d, sv, sa, cosT, kappa, v, a);
                }
                double s_ref, c_ref;
                sincos(th_ref, &s_ref, &c_ref);
                heading_cos_sin(cosT, dp, c_ref, s_ref, cn, sn);
            } else {
                double tanT;
                Divider<true> E;
                motion_carry(E, th_cl, dp, dpp, d, k_r, k_r_d, sv, sa, cosT, tanT, kappa, v, a);
                sincos(th_gl, &sn, &cn);
            }
            // cos / sin of the heading for the collision phase: the head of the (tail-only) increment rows is free
            scratch[4 * Np1 + i] = cn;
            scratch[5 * Np1 + i] = sn;
            s_kap[i] = kappa;
        }
        __syncthreads();                                                            // theta/kappa rows complete
        RP_STAMP(4);

        // ---- limits (reactive_planner.py:971-1017) + projection (:908-917) -------------------------
        double x = 0., y = 0., kdot = 0.;
        if (act) {
            const double th_prev = i > 0 ? s_th[i - 1] : 0.;
            const double kap_prev = i > 0 ? s_kap[i - 1] : 0.;
            const int r = check_constraints(P.lim, in.constraint_mask, dt, i, v, kappa, kap_prev, th_gl, th_prev, a);
            if (r != R_NONE) atomicMin(&s_flags[F_BAD], ((unsigned)i << 8) | (unsigned)r);
            kdot = i > 0 ? kappa - kap_prev : 0.;                  // np.append([0], np.diff(kappa_gl)) (:923)
            const int ub_ps = R.same_s ? ub : upper_bound_guess(R.ps, R.n, s, P.ps_inv_step);
            if (!project_to_cartesian(R, s, d, ub_ps, x, y)) {
                atomicMin(&s_flags[F_PBAD], (unsigned)i);
                x = 0.; y = 0.;
            }
        }
        __syncthreads();                                                            // verdict known
        RP_STAMP(5);
        const unsigned bad = lane ? s_flags[F_BAD] : NONE;
        const unsigned pbad = lane ? s_flags[F_PBAD] : NONE;
        const bool kin_ok = alive && bad == NONE && pbad == NONE;
        const bool keep = alive && (kin_ok || draw);              // states are produced for these
        const bool costed = kin_ok && in.cost_kind != RP_COST_NONE;
        if (lane && i == 0)
            s_flags[F_STATE] |= (alive ? S_ALIVE : 0u) | (kin_ok ? S_KINOK : 0u) | (keep ? S_KEEP : 0u);

        // ---- publish the polynomial part of the horizon -------------------------------------------
        if (keep && i < tl) {
            if (pbad != NONE && (unsigned)i > pbad) { x = 0.; y = 0.; }     // draw mode: loop broke at pbad
            s_x[i] = x;
            s_y[i] = y;
            if (i == tl - 1) {
                s_last[0] = x; s_last[1] = y; s_last[2] = th_gl; s_last[3] = v; s_last[4] = a; s_last[5] = kappa;
                s_last[6] = kdot; s_last[7] = s; s_last[8] = d; s_last[9] = th_cl; s_last[10] = sv; s_last[11] = sa;
                s_last[12] = dv; s_last[13] = da;
                s_last[14] = cn; s_last[15] = sn;
            }
            if (costed) {
                const double t0 = w_a * a, t1 = 5 * (v - in.desired_speed), t2 = 0.25 * (in.desired_s - s);
                const double t3 = 0.25 * (des_d - d), t4 = 0.25 * fabs(th_cl);
                s_ct[i] = t0 * t0; s_ct[Np1 + i] = t1 * t1; s_ct[2 * Np1 + i] = t2 * t2;
                s_ct[3 * Np1 + i] = t3 * t3; s_ct[4 * Np1 + i] = t4 * t4;
                if (i == Np1 - 1) { s_park[0] = v; s_park[1] = s; s_park[2] = d; s_park[3] = th_cl; }
                if (i == Np1 / 2) s_park[4] = v;                  // v[int(len(v) / 2)]
            }
            if (P.states != nullptr && !defer_states) {
                double* o = P.states + (size_t)(P.states_by_slot ? slot : k) * 14 * Np1 + i;
                o[0] = x; o[Np1] = y; o[2 * Np1] = th_gl; o[3 * Np1] = v; o[4 * Np1] = a; o[5 * Np1] = kappa;
                o[6 * Np1] = kdot; o[7 * Np1] = s; o[8 * Np1] = d; o[9 * Np1] = th_cl; o[10 * Np1] = sv;
                o[11 * Np1] = sa; o[12 * Np1] = dv; o[13 * Np1] = da;
            }
        }
        __syncthreads();
        RP_STAMP(6);

        // ---- horizon extension (trajectories.py:168-197, :302-332), element-strided over the tails ----
        const int n_elem = C * Np1;
        for (int e = tid, c2 = e_c0, i2 = e_i0; e < n_elem; e += T, c2 += e_dc, i2 += e_di) {
            if (i2 >= Np1) { i2 -= Np1; ++c2; }
            const unsigned* fl = s_flags_all + (size_t)c2 * F_WORDS;
            const int tl2 = (int)fl[F_TL];
            if (!(fl[F_STATE] & S_KEEP) || i2 < tl2) continue;
            double* sc = slots + (size_t)c2 * per_slot;
            const double* last = sc + kRows * Np1;
            const double tau = (double)(i2 - tl2 + 1) * dt;       // np.arange(1, steps + 1) * dt
            double v_tmp = last[3] + tau * last[4];
            v_tmp = v_tmp * (v_tmp >= 0 ? 1.0 : 0.0);
            sc[4 * Np1 + i2] = dt * v_tmp * last[14];             // increments of np.cumsum (x, y)
            sc[5 * Np1 + i2] = dt * v_tmp * last[15];
        }
        __syncthreads();
        for (int e = tid, c2 = e_c0, i2 = e_i0; e < n_elem; e += T, c2 += e_dc, i2 += e_di) {
            if (i2 >= Np1) { i2 -= Np1; ++c2; }
            const unsigned* fl = s_flags_all + (size_t)c2 * F_WORDS;
            const int tl2 = (int)fl[F_TL];
            const unsigned st2 = fl[F_STATE];
            if (!(st2 & S_KEEP) || i2 < tl2) continue;
            double* sc = slots + (size_t)c2 * per_slot;
            const double* last = sc + kRows * Np1;
            const double* tx = sc + 4 * Np1;
            const double* ty = sc + 5 * Np1;
            double ax = tx[tl2], ay = ty[tl2];
            for (int j = tl2 + 1; j <= i2; ++j) { ax += tx[j]; ay += ty[j]; }     // np.cumsum: sequential adds
            const double xe = last[0] + ax, ye = last[1] + ay;
            sc[2 * Np1 + i2] = xe;
            sc[3 * Np1 + i2] = ye;
            sc[i2] = last[2];                                     // theta_gl row for the collision phase
            const double tau = (double)(i2 - tl2 + 1) * dt;
            const double ae = last[4];
            double ve = last[3] + tau * ae;
            ve = ve * (ve >= 0 ? 1.0 : 0.0);
            // curvilinear: the velocity extrapolation multiplies the still-zero tail acceleration (App. B#7)
            double sve = last[10] + tau * 0.0;
            sve = sve * (sve >= 0 ? 1.0 : 0.0);
            const double dve = last[12] + tau * 0.0;
            const double se = last[7] + tau * last[10];
            const double de = last[8] + tau * last[12];
            const double the = last[9];
            if ((st2 & S_KINOK) && in.cost_kind != RP_COST_NONE) {
                double* ct = sc + 6 * Np1;
                const double t0 = w_a * ae, t1 = 5 * (ve - in.desired_speed), t2 = 0.25 * (in.desired_s - se);
                const double t3 = 0.25 * (des_d - de), t4 = 0.25 * fabs(the);
                ct[i2] = t0 * t0; ct[Np1 + i2] = t1 * t1; ct[2 * Np1 + i2] = t2 * t2;
                ct[3 * Np1 + i2] = t3 * t3; ct[4 * Np1 + i2] = t4 * t4;
                double* park = sc + kRows * Np1 + 16 + 40 + 5;
                if (i2 == Np1 - 1) { park[0] = ve; park[1] = se; park[2] = de; park[3] = the; }
                if (i2 == Np1 / 2) park[4] = ve;
            }
            if (P.states != nullptr && !defer_states) {
                const int k2 = (int)fl[F_K];
                const int slot2 = seg.k_begin + (g - seg.g_begin) * C + c2;
                double* o = P.states + (size_t)(P.states_by_slot ? slot2 : k2) * 14 * Np1 + i2;
                o[0] = xe; o[Np1] = ye; o[2 * Np1] = last[2]; o[3 * Np1] = ve; o[4 * Np1] = ae; o[5 * Np1] = last[5];
                o[6 * Np1] = last[6]; o[7 * Np1] = se; o[8 * Np1] = de; o[9 * Np1] = the; o[10 * Np1] = sve;
                o[11 * Np1] = last[11]; o[12 * Np1] = dve; o[13 * Np1] = last[13];
            }
        }
        __syncthreads();
        RP_STAMP(7);

        // ---- cost (cost_function.py:51-71, :85-92); np.sum order per SURVEY App. B#5 ----------------
        const bool par_sum = Np1 >= 8 && Np1 <= 128;
        if (in.cost_kind != RP_COST_NONE) {
            if (par_sum) {
                for (int e = tid; e < C * 40; e += T) {           // numpy's 8 accumulators per sum
                    const int c2 = e / 40, q = e - c2 * 40;
                    if (!(s_flags_all[(size_t)c2 * F_WORDS + F_STATE] & S_KINOK)) continue;
                    double* sc = slots + (size_t)c2 * per_slot;
                    const double* row = sc + (6 + (q >> 3)) * Np1;
                    const int j = q & 7;
                    double r = row[j];
                    for (int m = 8 + j; m < Np1 - (Np1 % 8); m += 8) r += row[m];
                    sc[kRows * Np1 + 16 + q] = r;
                }
            }
            __syncthreads();
            for (int e = tid; e < C * 5; e += T) {
                const int c2 = e / 5, u = e - c2 * 5;
                if (!(s_flags_all[(size_t)c2 * F_WORDS + F_STATE] & S_KINOK)) continue;
                double* sc = slots + (size_t)c2 * per_slot;
                const double* row = sc + (6 + u) * Np1;
                double res;
                if (par_sum) {
                    const double* r = sc + kRows * Np1 + 16 + u * 8;
                    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
                    for (int m = Np1 - (Np1 % 8); m < Np1; ++m) res += row[m];
                } else {
                    res = np_pairwise_sum(row, Np1);
                }
                sc[kRows * Np1 + 16 + 40 + u] = res;
            }
        }

        // total cost of slot c2 from its five sums and the parked end / mid values, in the reference's order
        auto total_cost = [&](int c2) {
            const double* sums = slots + (size_t)c2 * per_slot + kRows * Np1 + 16 + 40;
            const double* park = sums + 5;
            double costs = 0.0;
            costs += sums[0];
            if (!fs && in.has_desired_speed) {
                const double e1 = park[0] - in.desired_speed, e2 = park[4] - in.desired_speed;
                costs += sums[1] + (50 * (e1 * e1)) + (100 * (e2 * e2));
            }
            if (!fs && in.has_desired_s) {
                const double e = 20 * (in.desired_s - park[1]);
                costs += sums[2] + e * e;
            }
            {
                const double e = 20 * (des_d - park[2]);
                costs += sums[3] + e * e;
            }
            {
                const double e = 5 * fabs(park[3]);
                costs += sums[4] + e * e;
            }
            return costs;
        };
        // ---- lazy collision mode: the reference checks candidates in cost order and stops at the first
        // collision-free one (reactive_planner.py:1031-1063).  In parallel form: a candidate whose cost exceeds
        // the best collision-free cost found so far can be neither the winner nor a collider ranked before it,
        // so its check is skipped (status RP_FEASIBLE_UNCHECKED).  Winner and lazy collision count are exact.
        const bool lazy = in.check_collision == 2 && in.cost_kind != RP_COST_NONE && P.best_bits != nullptr;
        if (lazy) {
            __syncthreads();                                   // sums complete
            if (tid < C) {
                unsigned* fl = s_flags_all + (size_t)tid * F_WORDS;
                unsigned gate = 0u;
                if (fl[F_STATE] & S_KINOK) {
                    const double cst = total_cost(tid);
                    const unsigned long long best = *reinterpret_cast<volatile unsigned long long*>(best_ptr);
                    gate = !((unsigned long long)__double_as_longlong(cst) > best) ? 1u : 0u;   // costs are >= 0: bit order == value order
                    if (cst != cst) gate = 1u;                 // NaN cost: keep full semantics
                }
                fl[F_GATE] = gate;
            }
            __syncthreads();
        }

        RP_STAMP(8);
        // ---- ego-vs-obstacle check (reactive_planner.py:1026-1046), element-strided ----------------
        if (in.check_collision) {
            for (int e = tid, c2 = e_c0, i2 = e_i0; e < n_elem; e += T, c2 += e_dc, i2 += e_di) {
                if (i2 >= Np1) { i2 -= Np1; ++c2; }
                unsigned* fl = s_flags_all + (size_t)c2 * F_WORDS;
                if (!(fl[F_STATE] & S_KINOK)) continue;
                if (lazy && fl[F_GATE] == 0u) continue;
                const double* sc = slots + (size_t)c2 * per_slot;
                const bool tail = i2 >= (int)fl[F_TL];
                const double ct = tail ? sc[kRows * Np1 + 14] : sc[4 * Np1 + i2];
                const double st = tail ? sc[kRows * Np1 + 15] : sc[5 * Np1 + i2];
                const double ecx = sc[2 * Np1 + i2] + P.wb_rear * ct;
                const double ecy = sc[3 * Np1 + i2] + P.wb_rear * st;
                const int tidx = in.x0_time_step + i2 * in.factor;
                const bool hit = (dyn_stage ? dyn_collides_staged(O, dyn_stage, Np1, i2, tidx, ecx, ecy, ct, st, P.half_len, P.half_wid)
                                            : dyn_collides_global(O, tidx, ecx, ecy, ct, st, P.half_len, P.half_wid, P.r_ego)) ||
                                 static_collides(O, ecx, ecy, ct, st, P.half_len, P.half_wid);
                if (hit)
                    atomicMin(&fl[F_COL], (unsigned)i2);
            }
        }
        __syncthreads();
        RP_STAMP(9);

        // ---- per-candidate verdict ------------------------------------------------------------------
        unsigned long long my_bits = ~0ULL;                  // cost bits of a feasible, collision-free candidate (defer_states)
        int my_k = 0x7fffffff;
        if (tid < C && P.info != nullptr) {
            const unsigned* fl = s_flags_all + (size_t)tid * F_WORDS;
            const unsigned st2 = fl[F_STATE];
            if (st2 & S_VALID) {
                int status, reason = R_NONE, step = -1;
                double cost = __longlong_as_double(0x7ff8000000000000LL);   // NaN
                const unsigned pre2 = fl[F_PRE], bad2 = fl[F_BAD], pbad2 = fl[F_PBAD];
                if (st2 & S_FILTERED) {
                    status = ST_FILTERED;
                } else if (pre2 != 0u) {
                    status = ST_KINEMATIC;
                    reason = (pre2 & 1u) ? R_ACCELERATION : R_VELOCITY;
                } else if (bad2 != NONE) {
                    status = ST_KINEMATIC;
                    reason = (int)(bad2 & 0xFFu);
                    step = (int)(bad2 >> 8);
                } else if (pbad2 != NONE) {
                    status = ST_KINEMATIC;
                    reason = R_PROJECTION;
                    step = (int)pbad2;
                } else {
                    status = ST_FEASIBLE;
                    if (in.cost_kind != RP_COST_NONE) cost = total_cost(tid);
                    const unsigned cstep = fl[F_COL];
                    if (cstep != NONE) { status = ST_COLLISION; step = (int)cstep; }
                    else if (lazy) {
                        if (fl[F_GATE] == 0u) status = ST_UNCHECKED;
                        else if (cost == cost) {
                            const unsigned long long cb = (unsigned long long)__double_as_longlong(cost);
                            if (cb < *reinterpret_cast<volatile unsigned long long*>(best_ptr)) atomicMin(best_ptr, cb);
                        }
                    }
                }
                const int k2 = (int)fl[F_K];
                P.info[k2] = pack_info(status, reason, step);
                if (P.cost) P.cost[k2] = cost;
                if (defer_states && status == ST_FEASIBLE && cost == cost) {
                    my_bits = (unsigned long long)__double_as_longlong(cost);       // costs are >= 0: bit order == value order
                    my_k = k2;
                    atomicMin(&s_gbest, my_bits);
                }
            }
        }
        if (defer_states) {
            __syncthreads();
            if (my_bits == s_gbest && my_k != 0x7fffffff) atomicMin(&s_gk, my_k);   // ties: the lowest enumeration index
            __syncthreads();
            if (valid && k == s_gk) {
                // this slot holds the group's best candidate: its threads write the polynomial part from their
                // registers and the extension part from the slot's shared-memory rows (same expressions as above)
                double* o = P.states + (size_t)k * 14 * Np1;
                if (i < tl) {
                    double* q = o + i;
                    q[0] = x; q[Np1] = y; q[2 * Np1] = th_gl; q[3 * Np1] = v; q[4 * Np1] = a; q[5 * Np1] = kappa;
                    q[6 * Np1] = kdot; q[7 * Np1] = s; q[8 * Np1] = d; q[9 * Np1] = th_cl; q[10 * Np1] = sv;
                    q[11 * Np1] = sa; q[12 * Np1] = dv; q[13 * Np1] = da;
                }
                const double* last = s_last;
                for (int i2 = tl + i; i2 < Np1; i2 += tl) {
                    const double tau = (double)(i2 - tl + 1) * dt;
                    const double ae = last[4];
                    double ve = last[3] + tau * ae;
                    ve = ve * (ve >= 0 ? 1.0 : 0.0);
                    double sve = last[10] + tau * 0.0;
                    sve = sve * (sve >= 0 ? 1.0 : 0.0);
                    const double dve = last[12] + tau * 0.0;
                    double* q = o + i2;
                    q[0] = s_x[i2]; q[Np1] = s_y[i2]; q[2 * Np1] = last[2]; q[3 * Np1] = ve; q[4 * Np1] = ae; q[5 * Np1] = last[5];
                    q[6 * Np1] = last[6]; q[7 * Np1] = last[7] + tau * last[10]; q[8 * Np1] = last[8] + tau * last[12];
                    q[9 * Np1] = last[9]; q[10 * Np1] = sve; q[11 * Np1] = last[11]; q[12 * Np1] = dve; q[13 * Np1] = last[13];
                }
            }
        }
        __syncthreads();                                                            // scratch reusable
        RP_STAMP(10);
    }
}

template <int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT == 256 ? RP_FUSED_MIN_BLOCKS : 1)
fused_kernel(const __grid_constant__ PlanParams P) {
    extern __shared__ double smem[];
    fused_body<MAXT, false, NoCycleArgs>(P, nullptr, smem);
}

}  // namespace rp
