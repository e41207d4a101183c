// rp_capi.cu -- host side of the C-ABI declared in include/rp_b200.h: device-resident scenario
// tables, per-cycle upload / launch / result, and the small stand-alone entry points.
// No torch types cross this boundary; PyTorch only supplies the stream handle.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

#include "rp_kernels.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define RP_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(RP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));         \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return RP_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = std::max<size_t>(bytes, 256);
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) return fail(RP_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        cap = want;
        return RP_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return RP_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = std::max<size_t>(bytes, 4096);
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) return fail(RP_ERR_NOMEM, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
        cap = want;
        return RP_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

// largest (candidates x time steps) of one cycle launch (rp_plan_levels): the states of every kept candidate are written
constexpr long long kCycleMaxWork = 1LL << 20;

// launch geometry of the fused kernel
struct Geometry {
    int threads, Cmax, n_groups, n_segs, grid, stage_ref, stage_dyn;
    size_t smem;
    bool big;       // needs the 1024-thread instantiation
};

}  // namespace

struct rp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 148;
    int max_smem_optin = 0;

    bool have_vehicle = false, have_ref = false;
    rp_vehicle_params veh{};

    // reference tables
    int ref_n = 0, ref_same_s = 0;
    double ref_limit = 0;
    DevBuf d_ref;                       // 9 arrays of ref_n doubles

    // obstacles (host copies; device tables rebuilt lazily when the vehicle or obstacles change)
    std::vector<double> h_static, h_dyn, h_tri;
    std::vector<int> h_dyn_t0, h_dyn_len;
    double cell_size = 0.0;             // <= 0: automatic
    bool obstacles_dirty = true;
    DevBuf d_obb, d_tri, d_cell_start, d_cell_items, d_dyn_box, d_dyn_meta, d_clr;
    rp::ObstacleTables obs{};

    // per-cycle
    rp_plan_inputs in{};
    int mode = 0, n_t = 0, n_lon = 0, n_d = 0, n_cand = 0;
    int range_first = 0, range_count = -1;
    int stripe_rank = 0, stripe_world = 0;       // lon-interleaved shard (rp_set_candidate_stripe); world <= 1: off
    bool have_inputs = false, have_plan = false, states_all_valid = false;
    DevBuf d_samples;                   // t | lon | d (doubles) then traj_len (ints)
    DevBuf d_lon_coef, d_lat_coef, d_lat_tau, d_skip;
    DevBuf d_cost, d_info, d_states_all, d_result, d_index;
    PinBuf h_stage, h_result, h_flag;
    unsigned long long res_epoch = 0;
    // the result block: [PlanResultDev, padded to kResBytes][winner states 14 x (N+1)] -- ONE device->host copy per cycle
    static constexpr size_t kResBytes = 256;
    size_t res_states_bytes = 0;          // bytes of winner states in the block of the last launch
    bool h_states_valid = false;          // the pinned copy holds the last launch's winner states
    double* d_states_one() const { return reinterpret_cast<double*>(static_cast<char*>(d_result.p) + kResBytes); }
    size_t off_t = 0, off_lon = 0, off_d = 0, off_len = 0;

    // fused-kernel work decomposition (segments of equal traj_len)
    std::vector<int> h_traj_len;
    bool segs_dirty = true;
    bool geom_key_valid = false;        // shape of the inputs the current work decomposition was built for
    int geom_key_mode = -1, geom_key_n[9] = {};
    long long tables_version = 0;
    DevBuf d_segs, d_segs_index, d_argmin, d_best;
    PinBuf h_segs, h_segs_index;
    Geometry main_geom{}, index_geom{}, cand_geom{};
    bool main_is_cand = false, main_one_group = false, small_path_last = false;          // geometry in main_geom/d_segs belongs to the candidate-major kernel
    int kernel_policy = RP_KERNEL_AUTO;
    DevBuf d_work, d_dyn_rows, d_lat_rows, d_pose, d_defer_list, d_defer_mask;
    int index_geom_np1 = -1, index_geom_count = -1;
    long long index_geom_tables = -1;
    double ref_inv_step = 1.0, ps_inv_step = 1.0;

    // "copy done" events of the pinned staging buffers: a buffer is rewritten only after ITS last copy finished
    // (waiting on the whole stream would serialise back-to-back cycles of many contexts sharing one stream)
    cudaEvent_t ev_stage = nullptr, ev_segs = nullptr, ev_result = nullptr;
    cudaStream_t side_stream = nullptr;             // the collider count of a cycle runs beside the winner-state launch
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool stage_pending = false, segs_pending = false;

    // multi-GPU exchange over peer-mapped memory (rp_peer_*)
    rp::PeerMailbox* peer_mine = nullptr;           // cudaMalloc'ed, exported through CUDA IPC
    void* peer_opened[rp::kMaxPeers] = {};          // the other ranks' mailboxes (cudaIpcOpenMemHandle)
    rp::PeerTable peer_table{};
    bool peer_ready = false, peer_mode_last = false, peer_table_dirty = true;
    DevBuf d_peer_table;                            // the table in device memory (read by the merge block)
    unsigned long long peer_epoch = 0;

    // one replanning cycle in one launch (rp_plan_levels)
    bool cycle_valid = false;              // the last plan was a cycle launch: fetch_* address the selected level
    int cyc_n_levels = 0, cyc_chosen = 0, cyc_n_eval = 0, cyc_sel = 0, cyc_coeff_level = -1;
    rp::LevelDesc cyc_lv[rp::kMaxLevels] = {};
    std::vector<double> cyc_t[rp::kMaxLevels], cyc_lon[rp::kMaxLevels], cyc_d[rp::kMaxLevels];
    std::vector<int> cyc_tl[rp::kMaxLevels];
    rp::CycleOut* h_cycle = nullptr;       // mapped pinned host memory, written by the kernel's last block
    void* d_cycle = nullptr;               // its device alias
    size_t cycle_bytes = 0;
    unsigned long long cyc_epoch = 0;
    DevBuf d_cycle_res, d_ticket, d_best4;
    Geometry cycle_geom{};
    std::vector<int> cyc_geom_key;         // shapes the cached work decomposition was built for
    std::vector<rp::Segment> cyc_geom_segs;

    // a batch (rp_batch_*) may run this context's tables on ANOTHER stream: the end-of-launch event of the last batch
    // cycle that read them (owned by the batch); table updates wait for it before overwriting device memory
    cudaEvent_t ext_busy = nullptr;

    static constexpr int kEvRing = 64;
    cudaEvent_t ev_ring[kEvRing][5] = {};
    cudaEvent_t* ev = ev_ring[0];       // event set of the launch in flight
    bool ev_recorded[kEvRing] = {};     // the launch that last used the slot ran with stage timing on
    bool stage_timing = false;          // rp_ctx_set_stage_timing: five event records per launch cost a replanning-size
                                        // cycle 16 of its 82 us, so they are opt-in
    long long n_launches = 0;
};

namespace {

using rp::PlanParams;

int bind(rp_ctx* ctx) {
    if (!ctx) return fail(RP_ERR_ARG, "null context");
    RP_CUDA(cudaSetDevice(ctx->device));
    return RP_OK;
}

// ---- static broad-phase grid ------------------------------------------------------------------
int wait_external_readers(rp_ctx* ctx) {
    if (ctx->ext_busy) RP_CUDA(cudaEventSynchronize(ctx->ext_busy));
    return RP_OK;
}

int build_obstacle_tables(rp_ctx* ctx) {
    if (!ctx->have_vehicle) return fail(RP_ERR_STATE, "rp_ctx_set_vehicle must precede planning");
    if (int rc = wait_external_readers(ctx)) return rc;      // a batch cycle in flight on another stream still reads the old tables
    RP_CUDA(cudaStreamSynchronize(ctx->stream));
    const double hl = 0.5 * ctx->veh.length, hw = 0.5 * ctx->veh.width;
    const double r_ego = std::sqrt(hl * hl + hw * hw);
    const double infl = r_ego * (1.0 + 1e-9) + 1e-6;
    const int n_obb = (int)(ctx->h_static.size() / 5);
    const int n_tri = (int)(ctx->h_tri.size() / 6);
    const int n_prim = n_obb + n_tri;

    std::vector<double> obb((size_t)std::max(n_obb, 1) * rp::kBoxStride, 0.0);
    std::vector<double> lo_x(n_prim), hi_x(n_prim), lo_y(n_prim), hi_y(n_prim);
    for (int q = 0; q < n_obb; ++q) {
        const double* s = &ctx->h_static[(size_t)q * 5];
        const double c = std::cos(s[2]), sn = std::sin(s[2]);
        double* o = &obb[(size_t)q * rp::kBoxStride];
        o[0] = s[0]; o[1] = s[1]; o[2] = c; o[3] = sn; o[4] = s[3]; o[5] = s[4];
        o[6] = std::sqrt(s[3] * s[3] + s[4] * s[4]);
        {
            const double rr = (r_ego + o[6]) * 1.000000001 + 1e-9;     // same formula as rp::reach2
            o[7] = rr * rr;
        }
        const double ex = std::fabs(c) * s[3] + std::fabs(sn) * s[4];
        const double ey = std::fabs(sn) * s[3] + std::fabs(c) * s[4];
        lo_x[q] = s[0] - ex - infl; hi_x[q] = s[0] + ex + infl;
        lo_y[q] = s[1] - ey - infl; hi_y[q] = s[1] + ey + infl;
    }
    for (int q = 0; q < n_tri; ++q) {
        const double* t = &ctx->h_tri[(size_t)q * 6];
        lo_x[n_obb + q] = std::min({t[0], t[2], t[4]}) - infl; hi_x[n_obb + q] = std::max({t[0], t[2], t[4]}) + infl;
        lo_y[n_obb + q] = std::min({t[1], t[3], t[5]}) - infl; hi_y[n_obb + q] = std::max({t[1], t[3], t[5]}) + infl;
    }
    rp::ObstacleTables& O = ctx->obs;
    O = rp::ObstacleTables{};
    O.n_obb = n_obb;
    O.n_tri = n_tri;
    // ---- single-precision pre-reject frame ------------------------------------------------------------
    // Positions are taken relative to the centre of everything's bounding box.  With M = the largest relative
    // coordinate of an obstacle plus the largest reach plus 100 m, a relative coordinate rounds to fp32 with an
    // error <= M * 2^-24, a difference of two with <= 3 * M * 2^-24, the distance with <= 1.5 * that, and the
    // fp32 square / sum add <= 2^-21 relative.  margin = 1e-3 + 1e-6 * M (> 5 * M * 2^-24) metres on the reach and
    // a factor 1.00001 on its square cover all of it, so a pair the exact test would accept is never rejected.
    // An ego pose farther out than M is >= 100 m + reach away from every obstacle: far beyond any rounding.
    double bx0 = 1e300, bx1 = -1e300, by0 = 1e300, by1 = -1e300, r_max = 0.0;
    auto grow = [&](double x, double y, double r) {
        bx0 = std::min(bx0, x - r); bx1 = std::max(bx1, x + r); by0 = std::min(by0, y - r); by1 = std::max(by1, y + r);
        r_max = std::max(r_max, r);
    };
    for (int q = 0; q < n_obb; ++q) {
        const double* sq = &ctx->h_static[(size_t)q * 5];
        grow(sq[0], sq[1], std::sqrt(sq[3] * sq[3] + sq[4] * sq[4]));
    }
    for (int q = 0; q < n_tri; ++q) {
        const double* t = &ctx->h_tri[(size_t)q * 6];
        for (int c = 0; c < 3; ++c) grow(t[2 * c], t[2 * c + 1], 0.0);
    }
    for (size_t q = 0; q < ctx->h_dyn.size() / 5; ++q) {
        const double* sq = &ctx->h_dyn[q * 5];
        grow(sq[0], sq[1], std::sqrt(sq[3] * sq[3] + sq[4] * sq[4]));
    }
    if (bx0 > bx1) { bx0 = bx1 = by0 = by1 = 0.0; }
    O.org_x = 0.5 * (bx0 + bx1);
    O.org_y = 0.5 * (by0 + by1);
    const double M = std::max(bx1 - bx0, by1 - by0) * 0.5 + (r_ego + r_max) + 100.0;
    const bool f32_ok = M < 1.0e6;                       // beyond ~1000 km extent the pre-reject is switched off
    const double margin = 1e-3 + 1e-6 * M;
    auto r2_f32 = [&](double reach) -> float {           // squared fp32 reach, rounded up
        if (!f32_ok) return std::numeric_limits<float>::infinity();
        const double v = (reach + margin) * (reach + margin) * 1.00001;
        return std::nextafterf((float)v, std::numeric_limits<float>::infinity());
    };
    O.dyn_margin = f32_ok ? (float)margin : std::numeric_limits<float>::infinity();
    O.r_ego_f = (float)r_ego;
    if (n_prim > 0) {
        double gx0 = lo_x[0], gx1 = hi_x[0], gy0 = lo_y[0], gy1 = hi_y[0];
        for (int q = 1; q < n_prim; ++q) {
            gx0 = std::min(gx0, lo_x[q]); gx1 = std::max(gx1, hi_x[q]);
            gy0 = std::min(gy0, lo_y[q]); gy1 = std::max(gy1, hi_y[q]);
        }
        // default edge 0.5 m: with the ego circumradius folded into the lists, most poses on a free road land in an
        // EMPTY cell (one 4-byte load); doubled until the table stays below 2^20 cells
        double cell = ctx->cell_size > 0 ? ctx->cell_size : 0.5;
        double w = gx1 - gx0, h = gy1 - gy0;
        while ((std::ceil(w / cell) + 1) * (std::ceil(h / cell) + 1) > (ctx->cell_size > 0 ? 4.0e6 : 1048576.0)) cell *= 2.0;
        const int gnx = (int)std::ceil(w / cell) + 1, gny = (int)std::ceil(h / cell) + 1;
        const double inv = 1.0 / cell;
        std::vector<int> start((size_t)gnx * gny + 1, 0);
        auto span = [&](int q, int& ix0, int& ix1, int& iy0, int& iy1) {
            ix0 = std::max(0, (int)std::floor((lo_x[q] - gx0) * inv) - 0);
            ix1 = std::min(gnx - 1, (int)std::floor((hi_x[q] - gx0) * inv));
            iy0 = std::max(0, (int)std::floor((lo_y[q] - gy0) * inv));
            iy1 = std::min(gny - 1, (int)std::floor((hi_y[q] - gy0) * inv));
        };
        for (int q = 0; q < n_prim; ++q) {
            int ix0, ix1, iy0, iy1;
            span(q, ix0, ix1, iy0, iy1);
            for (int iy = iy0; iy <= iy1; ++iy)
                for (int ix = ix0; ix <= ix1; ++ix) ++start[(size_t)iy * gnx + ix + 1];
        }
        for (size_t q = 1; q < start.size(); ++q) start[q] += start[q - 1];
        std::vector<rp::CellItem> items((size_t)std::max(start.back(), 1));
        std::vector<int> fill(start.begin(), start.end() - 1);
        for (int q = 0; q < n_prim; ++q) {
            int ix0, ix1, iy0, iy1;
            span(q, ix0, ix1, iy0, iy1);
            rp::CellItem rec;
            rec.id = q;
            if (q < n_obb) {
                const double* o = &obb[(size_t)q * rp::kBoxStride];
                rec.cx = (float)(o[0] - O.org_x);
                rec.cy = (float)(o[1] - O.org_y);
                rec.r2 = r2_f32((r_ego + o[6]) * 1.000000001 + 1e-9);
            } else {
                const double* t = &ctx->h_tri[(size_t)(q - n_obb) * 6];
                const double tx0 = std::min({t[0], t[2], t[4]}), tx1 = std::max({t[0], t[2], t[4]});
                const double ty0 = std::min({t[1], t[3], t[5]}), ty1 = std::max({t[1], t[3], t[5]});
                const double mx = 0.5 * (tx0 + tx1), my = 0.5 * (ty0 + ty1);
                const double rt = 0.5 * std::sqrt((tx1 - tx0) * (tx1 - tx0) + (ty1 - ty0) * (ty1 - ty0));
                rec.cx = (float)(mx - O.org_x);
                rec.cy = (float)(my - O.org_y);
                rec.r2 = r2_f32((r_ego + rt) * 1.000000001 + 1e-9);
            }
            for (int iy = iy0; iy <= iy1; ++iy)
                for (int ix = ix0; ix <= ix1; ++ix) items[(size_t)fill[(size_t)iy * gnx + ix]++] = rec;
        }
        // clearance bitmask: cells touching a primitive's AABB inflated by the radius of the three covering circles
        const double clr_off = 2.0 * hl / 3.0;
        const double clr_r = std::sqrt((hl / 3.0) * (hl / 3.0) + hw * hw) * (1.0 + 1e-9) + 1e-6;
        std::vector<unsigned> bits(((size_t)gnx * gny + 31) / 32 + 1, 0u);
        for (int q = 0; q < n_prim; ++q) {
            const double shrink = infl - clr_r;              // lo/hi carry the r_ego inflation
            const int ix0 = std::max(0, (int)std::floor((lo_x[q] + shrink - gx0) * inv));
            const int ix1 = std::min(gnx - 1, (int)std::floor((hi_x[q] - shrink - gx0) * inv));
            const int iy0 = std::max(0, (int)std::floor((lo_y[q] + shrink - gy0) * inv));
            const int iy1 = std::min(gny - 1, (int)std::floor((hi_y[q] - shrink - gy0) * inv));
            for (int iy = iy0; iy <= iy1; ++iy)
                for (int ix = ix0; ix <= ix1; ++ix) {
                    const size_t c = (size_t)iy * gnx + ix;
                    bits[c >> 5] |= 1u << (c & 31);
                }
        }
        // second half of the buffer: occupancy bits (cell list non-empty)
        const size_t n_words = bits.size();
        bits.resize(2 * n_words, 0u);
        for (size_t c = 0; c + 1 < start.size(); ++c)
            if (start[c + 1] > start[c]) bits[n_words + (c >> 5)] |= 1u << (c & 31);
        if (int rc = ctx->d_clr.ensure(bits.size() * sizeof(unsigned))) return rc;
        RP_CUDA(cudaMemcpyAsync(ctx->d_clr.p, bits.data(), bits.size() * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
        O.clr_bits = ctx->d_clr.as<unsigned>();
        O.occ_bits = ctx->d_clr.as<unsigned>() + n_words;
        O.clr_off = clr_off;
        if (int rc = ctx->d_obb.ensure(obb.size() * sizeof(double))) return rc;
        if (int rc = ctx->d_tri.ensure(std::max<size_t>(ctx->h_tri.size(), 6) * sizeof(double))) return rc;
        if (int rc = ctx->d_cell_start.ensure(start.size() * sizeof(int))) return rc;
        if (int rc = ctx->d_cell_items.ensure(items.size() * sizeof(rp::CellItem))) return rc;
        RP_CUDA(cudaMemcpyAsync(ctx->d_obb.p, obb.data(), obb.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        if (n_tri)
            RP_CUDA(cudaMemcpyAsync(ctx->d_tri.p, ctx->h_tri.data(), ctx->h_tri.size() * sizeof(double),
                                    cudaMemcpyHostToDevice, ctx->stream));
        RP_CUDA(cudaMemcpyAsync(ctx->d_cell_start.p, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        RP_CUDA(cudaMemcpyAsync(ctx->d_cell_items.p, items.data(), items.size() * sizeof(rp::CellItem), cudaMemcpyHostToDevice, ctx->stream));
        RP_CUDA(cudaStreamSynchronize(ctx->stream));     // host vectors go out of scope
        O.obb = ctx->d_obb.as<double>();
        O.tri = ctx->d_tri.as<double>();
        O.gnx = gnx; O.gny = gny; O.gx0 = gx0; O.gy0 = gy0; O.inv_cell = inv;
        O.cell_start = ctx->d_cell_start.as<int>();
        O.cell_items = ctx->d_cell_items.as<rp::CellItem>();
    }
    // dynamic obstacles
    const int n_dyn = (int)ctx->h_dyn_t0.size();
    O.n_dyn = n_dyn;
    if (n_dyn > 0) {
        const size_t total = ctx->h_dyn.size() / 5;
        std::vector<double> box(total * rp::kBoxStride);
        for (size_t q = 0; q < total; ++q) {
            const double* s = &ctx->h_dyn[q * 5];
            double* o = &box[q * rp::kBoxStride];
            o[0] = s[0]; o[1] = s[1]; o[2] = std::cos(s[2]); o[3] = std::sin(s[2]); o[4] = s[3]; o[5] = s[4];
            o[6] = std::sqrt(s[3] * s[3] + s[4] * s[4]); o[7] = 1.0;
        }
        std::vector<int> meta((size_t)3 * n_dyn);
        int off = 0;
        for (int o = 0; o < n_dyn; ++o) {
            meta[o] = ctx->h_dyn_t0[o];
            meta[n_dyn + o] = ctx->h_dyn_len[o];
            meta[2 * n_dyn + o] = off;
            off += ctx->h_dyn_len[o];
        }
        if (int rc = ctx->d_dyn_box.ensure(std::max<size_t>(box.size(), 8) * sizeof(double))) return rc;
        if (int rc = ctx->d_dyn_meta.ensure(meta.size() * sizeof(int))) return rc;
        if (!box.empty())
            RP_CUDA(cudaMemcpyAsync(ctx->d_dyn_box.p, box.data(), box.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        RP_CUDA(cudaMemcpyAsync(ctx->d_dyn_meta.p, meta.data(), meta.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        RP_CUDA(cudaStreamSynchronize(ctx->stream));
        O.dyn_t0 = ctx->d_dyn_meta.as<int>();
        O.dyn_len = ctx->d_dyn_meta.as<int>() + n_dyn;
        O.dyn_off = ctx->d_dyn_meta.as<int>() + 2 * n_dyn;
        O.dyn_box = ctx->d_dyn_box.as<double>();
    }
    ctx->obstacles_dirty = false;
    ctx->segs_dirty = true;            // shared-memory staging depends on the obstacle counts
    ++ctx->tables_version;
    return RP_OK;
}

// deferred collision check: upper bound of the ego-box store of one launch (32 bytes per candidate-timestep; the dense
// sweep needs 256 MB, 64 dense scenarios in one batch 16 GB)
constexpr size_t kDeferMaxBytes = (size_t)48 << 30;

// ---- launch geometry of the fused kernel ---------------------------------------------------------
// Fill C / g_begin of the segments (k ranges and tl given), size shared memory, query occupancy.
int plan_geometry(rp_ctx* ctx, int Np1, std::vector<rp::Segment>& segs, Geometry& G, bool cycle = false, int threads = 0) {
    if (Np1 < 2 || Np1 > 1024) return fail(RP_ERR_ARG, "N + 1 must be in [2, 1024]");
    G.big = Np1 > 256;
    constexpr int kScratchKb = 48;                       // per-block cap of the per-slot scratch (measured, profiles/README.md)
    G.threads = G.big ? ((Np1 + 31) / 32) * 32 : (threads > 0 ? threads : 256);
    const size_t per_slot = (size_t)(rp::kRows * Np1 + rp::kSlotExtra) * sizeof(double) +
                            (size_t)(Np1 + rp::F_WORDS) * sizeof(int);
    // Shared-memory policy (measured on B200, profiles/README.md): three resident blocks per SM beat two,
    // so a block may use ~1/3 of the SM's shared memory.  Priority: the per-step dynamic-obstacle rows
    // (bank-conflict-free staging is worth 35 %), then per-slot scratch (lane utilisation for short
    // trajectories), then the reference tables (no measurable gain over L1-resident global reads).
    const size_t budget = (size_t)ctx->max_smem_optin;
    const size_t target = std::min<size_t>(budget, G.big ? budget : (size_t)74 * 1024);
    const size_t ref_bytes = (size_t)(ctx->ref_same_s ? 8 : 9) * ctx->ref_n * sizeof(double);
    const size_t dyn_bytes = (size_t)Np1 * ctx->obs.n_dyn * rp::kDynFields * sizeof(double);
    const size_t fixed = (size_t)segs.size() * sizeof(rp::Segment) + 64;
    G.stage_dyn = (ctx->obs.n_dyn > 0 && dyn_bytes <= 32 * 1024 && fixed + per_slot + dyn_bytes <= target) ? 1 : 0;
    const size_t avail = target - fixed - (G.stage_dyn ? dyn_bytes : 0);
    const size_t scratch_cap = std::min<size_t>(avail, (size_t)kScratchKb * 1024);
    const int c_budget = std::max<int>(1, (int)(scratch_cap / per_slot));
    G.Cmax = 1;
    G.n_groups = 0;
    for (auto& sg : segs) {
        const int tl = std::max(1, std::min(sg.tl, Np1));
        sg.tl = tl;
        sg.C = G.big ? 1 : std::max(1, std::min(c_budget, G.threads / tl));
        sg.g_begin = G.n_groups;
        G.n_groups += (std::max(0, sg.k_end - sg.k_begin) + sg.C - 1) / sg.C;
        G.Cmax = std::max(G.Cmax, sg.C);
    }
    G.n_segs = (int)segs.size();
    G.smem = (size_t)G.Cmax * per_slot + fixed + (G.stage_dyn ? dyn_bytes : 0);
    if (G.smem > budget) return fail(RP_ERR_ARG, "horizon too long for shared-memory scratch");
    G.stage_ref = (G.smem + ref_bytes <= target) ? 1 : 0;
    if (G.stage_ref) G.smem += ref_bytes;
    // (decided below, once the grid is known: a block that runs only a group or two reads the tables through L1 --
    // staging 13 KB per block was 18 % of a replanning-size launch, profiles/README.md round 2)
    const size_t smem_without_ref = G.smem - (G.stage_ref ? ref_bytes : 0);
    int occ = 0;
    // the attribute is a per-function maximum: only ever raise it (main and index launches share the kernel)
    // (per device, shared by all contexts of the process: the attribute belongs to the function, not the context)
    static int g_granted[64][3] = {};
    static std::mutex g_granted_mutex;                      // distinct contexts may plan from distinct threads
    std::lock_guard<std::mutex> granted_lock(g_granted_mutex);
    if (cycle && G.big) return fail(RP_ERR_ARG, "cycle launch: N + 1 must be <= 256");
    int& granted = g_granted[ctx->device & 63][cycle ? 2 : (G.big ? 1 : 0)];
    if ((int)G.smem > granted) {
        if (cycle) {
            RP_CUDA(cudaFuncSetAttribute(rp::cycle_kernel<256, rp::CycleArgs>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem));
            RP_CUDA(cudaFuncSetAttribute(rp::cycle_kernel<256, rp::CycleArgsSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem));
        }
        else if (G.big) RP_CUDA(cudaFuncSetAttribute(rp::fused_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem));
        else RP_CUDA(cudaFuncSetAttribute(rp::fused_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem));
        granted = (int)G.smem;
    }
    if (cycle) RP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rp::cycle_kernel<256, rp::CycleArgs>, G.threads, G.smem));
    else if (G.big) RP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rp::fused_kernel<1024>, G.threads, G.smem));
    else RP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rp::fused_kernel<256>, G.threads, G.smem));
    if (occ < 1) return fail(RP_ERR_CUDA, "fused kernel does not fit on an SM");
    G.grid = std::max(1, std::min(G.n_groups, occ * ctx->num_sms));
    if (G.stage_ref && G.n_groups < 3 * G.grid) {
        G.stage_ref = 0;
        G.smem = smem_without_ref;
    }
    return RP_OK;
}

// Launch geometry of the candidate-major kernel (rp_cand.cuh): segments sorted by traj_len, longest first,
// cut into chunks of 32 candidates that the warps of a persistent grid draw from a counter.
int plan_cand_geometry(rp_ctx* ctx, int Np1, std::vector<rp::Segment>& segs, Geometry& G, int n_acc_rows, bool one_group) {
    std::stable_sort(segs.begin(), segs.end(), [](const rp::Segment& a, const rp::Segment& b) { return a.tl > b.tl; });
    G.big = false;
    G.threads = RP_CAND_THREADS;
    G.Cmax = 32;
    G.n_groups = 0;
    for (auto& sg : segs) {
        sg.tl = std::max(1, std::min(sg.tl, Np1));
        sg.C = 32;
        sg.g_begin = G.n_groups;
        G.n_groups += (std::max(0, sg.k_end - sg.k_begin) + 31) / 32;
    }
    G.n_segs = (int)segs.size();
    const size_t budget = (size_t)ctx->max_smem_optin;
    const size_t acc_bytes = (size_t)(n_acc_rows * 8 + 1) * G.threads * sizeof(double);    // + v_mid
    const size_t fixed = (size_t)segs.size() * sizeof(rp::Segment) + 128 +
                         (size_t)(G.threads / 32) * rp::warp_row_doubles(one_group ? RP_CAND_ONE_GROUP_SLOTS : 32) * sizeof(double);   // the warps' longitudinal rows
    G.stage_dyn = 0;                                   // dynamic-obstacle rows are read through L1 (dyn_rows_kernel)
    G.smem = acc_bytes + fixed;
    G.stage_ref = 0;                                   // reference tables through L1 (staging them cost occupancy, profiles/README.md)
    if (G.smem > budget) return fail(RP_ERR_ARG, "candidate-major kernel: shared-memory need exceeds the SM");
    static int g_granted[64] = {};
    static std::mutex g_granted_mutex;
    std::lock_guard<std::mutex> granted_lock(g_granted_mutex);
    int& granted = g_granted[ctx->device & 63];
    if ((int)G.smem > granted) {
        RP_CUDA(cudaFuncSetAttribute(rp::cand_kernel<RP_CAND_THREADS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem));
        RP_CUDA(cudaFuncSetAttribute(rp::cand_kernel<RP_CAND_THREADS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem));
        RP_CUDA(cudaFuncSetAttribute(rp::cand_kernel<RP_CAND_THREADS, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem));
        RP_CUDA(cudaFuncSetAttribute(rp::cand_kernel<RP_CAND_THREADS, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem));
        RP_CUDA(cudaFuncSetAttribute(rp::cand_kernel<RP_CAND_THREADS, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem));
        RP_CUDA(cudaFuncSetAttribute(rp::cand_kernel<RP_CAND_THREADS, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem));
        RP_CUDA(cudaFuncSetAttribute(rp::cand_kernel<RP_CAND_THREADS, false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem));
        RP_CUDA(cudaFuncSetAttribute(rp::cand_kernel<RP_CAND_THREADS, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem));
        granted = (int)G.smem;
    }
    int occ = 0;
    RP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rp::cand_kernel<RP_CAND_THREADS, false>, G.threads, G.smem));
    if (occ < 1) return fail(RP_ERR_CUDA, "candidate-major kernel does not fit on an SM");
    const int warps_per_block = G.threads / 32;
    G.grid = std::max(1, std::min((G.n_groups + warps_per_block - 1) / warps_per_block, occ * ctx->num_sms));
    return RP_OK;
}

// np.sum accumulator rows of the candidate-major kernel: acceleration, lateral offset, orientation (+ velocity, + position)
int cand_acc_rows(const rp_plan_inputs& in) {
    if (in.cost_kind == RP_COST_NONE) return 0;
    const bool fs = in.cost_kind == RP_COST_FAILSAFE;
    return 3 + ((!fs && in.has_desired_speed) ? 1 : 0) + ((!fs && in.has_desired_s) ? 1 : 0);
}

bool stripe_on(const rp_ctx* ctx) { return ctx->stripe_world > 1 && ctx->mode == 0; }
int stripe_n_lon(const rp_ctx* ctx) { return (ctx->n_lon - ctx->stripe_rank + ctx->stripe_world - 1) / ctx->stripe_world; }
rp::Stripe stripe_of(const rp_ctx* ctx) {
    return stripe_on(ctx) ? rp::Stripe{ctx->stripe_rank, ctx->stripe_world, ctx->n_lon, ctx->n_d} : rp::Stripe{0, 0, 0, 0};
}
// the shard of the next launch: [first, first + count) in the launch's (virtual, for stripes) enumeration
void shard_extent(const rp_ctx* ctx, int& first, int& count) {
    const int n = ctx->n_cand;
    first = 0;
    count = n;
    if (stripe_on(ctx)) {
        count = ctx->n_t * std::max(0, stripe_n_lon(ctx)) * ctx->n_d;
    } else if (ctx->range_count >= 0) {
        first = std::min(ctx->range_first, n);
        count = std::min(ctx->range_count, n - first);
    }
}

// which kernel evaluates the main launch (the winner-state / on-demand launches always use fused_kernel)
bool use_cand_kernel(const rp_ctx* ctx, int count) {
    constexpr int kCandMin = 24576;                    // AUTO: candidate-major from this many candidates up
    const int Np1 = ctx->in.N + 1;
    if (ctx->in.want_all_states || ctx->in.draw_all || Np1 > 128) return false;     // not handled by cand_kernel
    const int policy = ctx->kernel_policy;
    if (policy == RP_KERNEL_STEP_PARALLEL) return false;
    if (policy == RP_KERNEL_CANDIDATE_MAJOR) return true;
    return count >= kCandMin;
}

void fill_common(rp_ctx* ctx, PlanParams& P, const Geometry& G, const rp::Segment* d_segs) {
    P.in = ctx->in;
    P.lim.a_max = ctx->veh.a_max;
    P.lim.v_switch = ctx->veh.v_switch;
    P.lim.wheelbase = ctx->veh.wheelbase;
    P.lim.v_delta_max = ctx->veh.v_delta_max;
    P.lim.kappa_max = ctx->veh.kappa_max;
    P.half_len = 0.5 * ctx->veh.length;
    P.half_wid = 0.5 * ctx->veh.width;
    P.wb_rear = ctx->veh.wb_rear_axle;
    P.r_ego = std::sqrt(P.half_len * P.half_len + P.half_wid * P.half_wid);
    P.r_ego_f_up = std::nextafterf((float)P.r_ego, std::numeric_limits<float>::infinity());
    P.wb_rear_f_up = std::nextafterf((float)std::fabs(P.wb_rear), std::numeric_limits<float>::infinity());
    const double* base = ctx->d_ref.as<double>();
    const int n = ctx->ref_n;
    P.ref.n = n;
    P.ref.same_s = ctx->ref_same_s;
    P.ref.limit = ctx->ref_limit;
    P.ref.pos = base; P.ref.theta = base + n; P.ref.curv = base + 2 * n; P.ref.curv_d = base + 3 * n;
    P.ref.px = base + 4 * n; P.ref.py = base + 5 * n; P.ref.nx = base + 6 * n; P.ref.ny = base + 7 * n;
    P.ref.ps = base + 8 * n;
    P.ref_inv_step = ctx->ref_inv_step;
    P.ps_inv_step = ctx->ps_inv_step;
    P.obs = ctx->obs;
    P.segs = d_segs;
    P.n_segs = G.n_segs;
    P.n_groups = G.n_groups;
    P.Cmax = G.Cmax;
    P.Np1 = ctx->in.N + 1;
    P.stage_ref = G.stage_ref;
    P.stage_dyn = G.stage_dyn;
    P.mode = ctx->mode;
    P.n_t = ctx->n_t; P.n_lon = ctx->n_lon; P.n_d = ctx->n_d;
    P.n_cand = ctx->n_cand;
    if (ctx->mode == 2) {                      // cycle launch: samples, segments and coefficients live in CycleArgs / shared memory
        P.lon_samples = nullptr;
        P.traj_len = nullptr;
        P.skip = nullptr;
    } else if (ctx->mode == 0) {
        const char* sb = static_cast<const char*>(ctx->d_samples.p);
        P.lon_samples = reinterpret_cast<const double*>(sb + ctx->off_lon);
        P.traj_len = reinterpret_cast<const int*>(sb + ctx->off_len);
        P.skip = nullptr;
    } else {
        P.lon_samples = nullptr;
        P.traj_len = reinterpret_cast<const int*>(static_cast<const char*>(ctx->d_samples.p) + ctx->off_len);
        P.skip = ctx->d_skip.p ? ctx->d_skip.as<uint8_t>() : nullptr;
    }
    P.lon_coef = ctx->d_lon_coef.as<double>();
    P.lat_coef = ctx->d_lat_coef.as<double>();
    P.stripe_rank = 0;                         // (set by the main launch of a striped shard only)
    P.stripe_world = 0;
}

int launch_fused(rp_ctx* ctx, const PlanParams& P, const Geometry& G) {
    if (G.big) rp::fused_kernel<1024><<<G.grid, G.threads, G.smem, ctx->stream>>>(P);
    else rp::fused_kernel<256><<<G.grid, G.threads, G.smem, ctx->stream>>>(P);
    RP_CUDA(cudaGetLastError());
    return RP_OK;
}

// segment table of the main launch: one segment per sampled t (grid) clipped to the candidate range,
// or one segment over the list.  Rebuilt (and re-uploaded) only when inputs / range / tables changed.
int prepare_main_geometry(rp_ctx* ctx, int first, int count) {
    if (!ctx->segs_dirty) return RP_OK;
    const int Np1 = ctx->in.N + 1;
    std::vector<rp::Segment> segs;
    if (ctx->mode == 0) {
        const int per_t = (stripe_on(ctx) ? std::max(0, stripe_n_lon(ctx)) : ctx->n_lon) * ctx->n_d;   // (virtual enumeration of a stripe)
        for (int it = 0; it < ctx->n_t && per_t > 0; ++it) {
            const int b = std::max(first, it * per_t), e = std::min(first + count, (it + 1) * per_t);
            if (b < e) segs.push_back(rp::Segment{b, e, ctx->h_traj_len[it], 0, 0, 0});
        }
    } else if (count > 0) {
        segs.push_back(rp::Segment{first, first + count, Np1, 0, 0, 0});
    }
    if (segs.empty()) segs.push_back(rp::Segment{0, 0, Np1, 0, 0, 0});
    ctx->main_is_cand = use_cand_kernel(ctx, count);
    if (ctx->main_is_cand) {
        // (list form: one segment in list order; traj_len varies per candidate, lanes diverge on i < tl only)
        // one longitudinal group per chunk of 32: grid form, n_d a multiple of 32, shard boundaries on multiples of 32
        ctx->main_one_group = ctx->mode == 0 && ctx->n_d > 0 && ctx->n_d % 32 == 0 && first % 32 == 0 && count % 32 == 0;
        std::vector<rp::Segment> sorted = segs;
        if (plan_cand_geometry(ctx, Np1, sorted, ctx->main_geom, cand_acc_rows(ctx->in), ctx->main_one_group) == RP_OK) {
            segs.swap(sorted);
        } else {
            ctx->main_is_cand = false;          // e.g. a segment table too long for its shared memory: the other schedule
        }
    }
    if (!ctx->main_is_cand)
        if (int rc = plan_geometry(ctx, Np1, segs, ctx->main_geom)) return rc;
    const size_t bytes = segs.size() * sizeof(rp::Segment);
    if (int rc = ctx->h_segs.ensure(bytes)) return rc;
    if (int rc = ctx->d_segs.ensure(bytes)) return rc;
    if (ctx->segs_pending) RP_CUDA(cudaEventSynchronize(ctx->ev_segs));      // the previous copy of this buffer
    std::memcpy(ctx->h_segs.p, segs.data(), bytes);
    RP_CUDA(cudaMemcpyAsync(ctx->d_segs.p, ctx->h_segs.p, bytes, cudaMemcpyHostToDevice, ctx->stream));
    RP_CUDA(cudaEventRecord(ctx->ev_segs, ctx->stream));
    ctx->segs_pending = true;
    ctx->segs_dirty = false;
    return RP_OK;
}

int prepare_index_geometry(rp_ctx* ctx, int count) {
    const int Np1 = ctx->in.N + 1;
    if (ctx->index_geom_np1 == Np1 && ctx->index_geom_count == count && ctx->index_geom_tables == ctx->tables_version)
        return RP_OK;
    std::vector<rp::Segment> segs{rp::Segment{0, count, Np1, 0, 0, 0}};
    if (int rc = plan_geometry(ctx, Np1, segs, ctx->index_geom)) return rc;
    if (int rc = ctx->h_segs_index.ensure(sizeof(rp::Segment))) return rc;
    if (int rc = ctx->d_segs_index.ensure(sizeof(rp::Segment))) return rc;
    RP_CUDA(cudaStreamSynchronize(ctx->stream));
    std::memcpy(ctx->h_segs_index.p, segs.data(), sizeof(rp::Segment));
    RP_CUDA(cudaMemcpyAsync(ctx->d_segs_index.p, ctx->h_segs_index.p, sizeof(rp::Segment), cudaMemcpyHostToDevice, ctx->stream));
    ctx->index_geom_np1 = Np1;
    ctx->index_geom_count = count;
    ctx->index_geom_tables = ctx->tables_version;
    return RP_OK;
}

int check_ready(rp_ctx* ctx) {
    if (!ctx->have_vehicle) return fail(RP_ERR_STATE, "vehicle parameters not set");
    if (!ctx->have_ref) return fail(RP_ERR_STATE, "reference tables not set");
    if (ctx->obstacles_dirty)
        if (int rc = build_obstacle_tables(ctx)) return rc;
    return RP_OK;
}

int check_inputs(const rp_plan_inputs* in) {
    if (!in) return fail(RP_ERR_ARG, "null inputs");
    if (in->N < 1 || in->N > 1023) return fail(RP_ERR_ARG, "N out of range");
    if (!(in->dt > 0)) return fail(RP_ERR_ARG, "dt must be positive");
    if (in->lon_mode != RP_VELOCITY_KEEPING && in->lon_mode != RP_STOPPING) return fail(RP_ERR_ARG, "invalid lon_mode");
    if (in->cost_kind < RP_COST_DEFAULT || in->cost_kind > RP_COST_NONE) return fail(RP_ERR_ARG, "invalid cost_kind");
    return RP_OK;
}

// winner (or arbitrary candidate) state block via an index-mode launch
int launch_states_for_index(rp_ctx* ctx, const int* d_index, int count, double* d_out) {
    if (int rc = prepare_index_geometry(ctx, count)) return rc;
    PlanParams P{};
    fill_common(ctx, P, ctx->index_geom, ctx->d_segs_index.as<rp::Segment>());
    P.index = d_index;
    P.cost = nullptr;
    P.info = nullptr;
    P.states = d_out;
    P.states_by_slot = 1;
    P.in.draw_all = 1;          // produce states whatever the verdict was
    P.in.check_collision = 0;
    P.in.cost_kind = RP_COST_NONE;
    // a handful of candidates: nothing to amortise a staging pass over (tables through L1, no collision phase)
    P.stage_dyn = 0;
    if (count <= 4) P.stage_ref = 0;
    return launch_fused(ctx, P, ctx->index_geom);
}

// cycle launch: the coefficient arrays / device sample lists of the SELECTED level, built on demand (only the lazy paths
// need them: rp_fetch_coeffs and the re-evaluation of a non-winner's states; the cycle kernel itself solves in shared memory)
int cycle_ensure_coeffs(rp_ctx* ctx) {
    const int lv = ctx->cyc_sel;
    if (ctx->cyc_coeff_level == lv) return RP_OK;
    const int n_t = (int)ctx->cyc_t[lv].size(), n_lon = (int)ctx->cyc_lon[lv].size(), n_d = (int)ctx->cyc_d[lv].size();
    ctx->off_t = 0;
    ctx->off_lon = (size_t)n_t * sizeof(double);
    ctx->off_d = ctx->off_lon + (size_t)n_lon * sizeof(double);
    ctx->off_len = ctx->off_d + (size_t)n_d * sizeof(double);
    const size_t bytes = ctx->off_len + (size_t)n_t * sizeof(int);
    if (int rc = ctx->h_stage.ensure(bytes)) return rc;
    if (int rc = ctx->d_samples.ensure(bytes)) return rc;
    if (ctx->stage_pending) RP_CUDA(cudaEventSynchronize(ctx->ev_stage));
    char* hs = static_cast<char*>(ctx->h_stage.p);
    std::memcpy(hs + ctx->off_t, ctx->cyc_t[lv].data(), (size_t)n_t * sizeof(double));
    std::memcpy(hs + ctx->off_lon, ctx->cyc_lon[lv].data(), (size_t)n_lon * sizeof(double));
    std::memcpy(hs + ctx->off_d, ctx->cyc_d[lv].data(), (size_t)n_d * sizeof(double));
    std::memcpy(hs + ctx->off_len, ctx->cyc_tl[lv].data(), (size_t)n_t * sizeof(int));
    RP_CUDA(cudaMemcpyAsync(ctx->d_samples.p, hs, bytes, cudaMemcpyHostToDevice, ctx->stream));
    RP_CUDA(cudaEventRecord(ctx->ev_stage, ctx->stream));
    ctx->stage_pending = true;
    const int n = n_t * n_lon * n_d;
    const int n_lon_sys = n_t * n_lon, n_lat_sys = ctx->in.low_vel_mode ? n : n_t * n_d;
    if (int rc = ctx->d_lon_coef.ensure((size_t)std::max(n_lon_sys, 1) * 6 * sizeof(double))) return rc;
    if (int rc = ctx->d_lat_coef.ensure((size_t)std::max(n_lat_sys, 1) * 6 * sizeof(double))) return rc;
    if (int rc = ctx->d_lat_tau.ensure((size_t)std::max(n_lat_sys, 1) * sizeof(double))) return rc;
    if (n_lon_sys + n_lat_sys > 0) {
        const char* sb = static_cast<const char*>(ctx->d_samples.p);
        rp::coeff_kernel<<<(n_lon_sys + n_lat_sys + 127) / 128, 128, 0, ctx->stream>>>(
            n_t, n_lon, n_d, ctx->in.low_vel_mode, ctx->in.lon_mode, reinterpret_cast<const double*>(sb + ctx->off_t),
            reinterpret_cast<const double*>(sb + ctx->off_lon), reinterpret_cast<const double*>(sb + ctx->off_d), ctx->in.x0_lon[0],
            ctx->in.x0_lon[1], ctx->in.x0_lon[2], ctx->in.x0_lat[0], ctx->in.x0_lat[1], ctx->in.x0_lat[2], ctx->d_lon_coef.as<double>(),
            ctx->d_lat_coef.as<double>(), ctx->d_lat_tau.as<double>());
        RP_CUDA(cudaGetLastError());
    }
    ctx->cyc_coeff_level = lv;
    return RP_OK;
}

}  // namespace

// =====================================================================================================
extern "C" {

const char* rp_last_error(void) { return g_err.c_str(); }
int rp_version(void) { return 100; }

int rp_ctx_create(int device, void* stream, rp_ctx** out) {
    if (!out) return fail(RP_ERR_ARG, "null out pointer");
    *out = nullptr;
    int n_dev = 0;
    RP_CUDA(cudaGetDeviceCount(&n_dev));
    if (device < 0 || device >= n_dev) return fail(RP_ERR_ARG, "invalid CUDA device index");
    RP_CUDA(cudaSetDevice(device));
    rp_ctx* ctx = new rp_ctx();
    ctx->device = device;
    if (stream) {
        ctx->stream = static_cast<cudaStream_t>(stream);
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete ctx; return fail(RP_ERR_CUDA, cudaGetErrorString(e)); }
        ctx->own_stream = true;
    }
    cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    for (auto& set : ctx->ev_ring)
        for (auto& e : set) cudaEventCreate(&e);
    cudaEventCreateWithFlags(&ctx->ev_stage, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_segs, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_result, cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    static_assert(sizeof(rp::PlanResultDev) <= rp_ctx::kResBytes, "result block header too small");
    if (ctx->d_result.ensure(rp_ctx::kResBytes) || ctx->h_result.ensure(rp_ctx::kResBytes) ||
        ctx->d_index.ensure(sizeof(int))) {
        rp_ctx_destroy(ctx);
        return RP_ERR_NOMEM;
    }
    *out = ctx;
    return RP_OK;
}

int rp_ctx_destroy(rp_ctx* ctx) {
    if (!ctx) return RP_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    rp_peer_close(ctx);
    for (DevBuf* b : {&ctx->d_ref, &ctx->d_obb, &ctx->d_tri, &ctx->d_cell_start, &ctx->d_cell_items, &ctx->d_dyn_box,
                      &ctx->d_dyn_meta, &ctx->d_samples, &ctx->d_lon_coef, &ctx->d_lat_coef, &ctx->d_lat_tau,
                      &ctx->d_skip, &ctx->d_cost, &ctx->d_info, &ctx->d_states_all,
                      &ctx->d_result, &ctx->d_index, &ctx->d_segs, &ctx->d_segs_index, &ctx->d_argmin, &ctx->d_best,
                      &ctx->d_work, &ctx->d_clr, &ctx->d_dyn_rows, &ctx->d_lat_rows, &ctx->d_pose, &ctx->d_defer_list,
                      &ctx->d_defer_mask})
        b->release();
    for (DevBuf* b : {&ctx->d_cycle_res, &ctx->d_ticket, &ctx->d_best4, &ctx->d_peer_table}) b->release();
    if (ctx->h_cycle) cudaFreeHost(ctx->h_cycle);
    ctx->h_stage.release();
    ctx->h_result.release();
    ctx->h_flag.release();
    ctx->h_segs.release();
    ctx->h_segs_index.release();
    for (auto& set : ctx->ev_ring)
        for (auto& e : set)
            if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : {ctx->ev_stage, ctx->ev_segs, ctx->ev_result, ctx->ev_fork, ctx->ev_join})
        if (e) cudaEventDestroy(e);
    if (ctx->side_stream) {
        cudaStreamSynchronize(ctx->side_stream);
        cudaStreamDestroy(ctx->side_stream);
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return RP_OK;
}

int rp_ctx_synchronize(rp_ctx* ctx) {
    if (int rc = bind(ctx)) return rc;
    RP_CUDA(cudaStreamSynchronize(ctx->stream));
    return RP_OK;
}

int rp_ctx_set_vehicle(rp_ctx* ctx, const rp_vehicle_params* vp) {
    if (int rc = bind(ctx)) return rc;
    if (!vp) return fail(RP_ERR_ARG, "null vehicle parameters");
    if (!(vp->length > 0 && vp->width > 0 && vp->wheelbase > 0)) return fail(RP_ERR_ARG, "vehicle dimensions must be positive");
    ctx->veh = *vp;
    ctx->have_vehicle = true;
    ctx->obstacles_dirty = true;       // broad-phase inflation depends on the ego circumradius
    return RP_OK;
}

int rp_ctx_set_reference(rp_ctx* ctx, int n, const double* ref_pos, const double* ref_theta, const double* ref_curv,
                         const double* ref_curv_d, const double* path_xy, const double* path_s,
                         const double* path_normal_xy, double proj_limit) {
    if (int rc = bind(ctx)) return rc;
    if (n < 2) return fail(RP_ERR_ARG, "reference path needs at least 2 points");
    if (!ref_pos || !ref_theta || !ref_curv || !ref_curv_d || !path_xy || !path_s || !path_normal_xy)
        return fail(RP_ERR_ARG, "null reference array");
    for (int q = 1; q < n; ++q)
        if (!(ref_pos[q] > ref_pos[q - 1]) || !(path_s[q] > path_s[q - 1]))
            return fail(RP_ERR_ARG, "ref_pos / path_s must be strictly increasing");
    std::vector<double> pack((size_t)9 * n);
    std::memcpy(&pack[0], ref_pos, n * sizeof(double));
    std::memcpy(&pack[(size_t)n], ref_theta, n * sizeof(double));
    std::memcpy(&pack[(size_t)2 * n], ref_curv, n * sizeof(double));
    std::memcpy(&pack[(size_t)3 * n], ref_curv_d, n * sizeof(double));
    for (int q = 0; q < n; ++q) {
        pack[(size_t)4 * n + q] = path_xy[2 * q];
        pack[(size_t)5 * n + q] = path_xy[2 * q + 1];
        pack[(size_t)6 * n + q] = path_normal_xy[2 * q];
        pack[(size_t)7 * n + q] = path_normal_xy[2 * q + 1];
    }
    std::memcpy(&pack[(size_t)8 * n], path_s, n * sizeof(double));
    if (int rc = wait_external_readers(ctx)) return rc;
    RP_CUDA(cudaStreamSynchronize(ctx->stream));           // launches in flight may still read the old tables
    if (int rc = ctx->d_ref.ensure(pack.size() * sizeof(double))) return rc;
    RP_CUDA(cudaMemcpyAsync(ctx->d_ref.p, pack.data(), pack.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    RP_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->ref_n = n;
    ctx->ref_same_s = std::memcmp(ref_pos, path_s, n * sizeof(double)) == 0 ? 1 : 0;
    ctx->ref_limit = proj_limit;
    ctx->ref_inv_step = (double)(n - 1) / (ref_pos[n - 1] - ref_pos[0]);
    ctx->ps_inv_step = (double)(n - 1) / (path_s[n - 1] - path_s[0]);
    ctx->segs_dirty = true;
    ++ctx->tables_version;
    ctx->have_ref = true;
    return RP_OK;
}

int rp_ctx_set_reference_polyline(rp_ctx* ctx, int n_pts, const double* xy, double proj_limit, double eps2) {
    if (int rc = bind(ctx)) return rc;
    if (n_pts < 3 || !xy) return fail(RP_ERR_ARG, "reference polyline needs at least 3 points");
    if (!(eps2 > 0) || !(proj_limit > 0)) return fail(RP_ERR_ARG, "eps2 and proj_limit must be positive");
    for (int q = 1; q < n_pts; ++q)
        if (xy[2 * q] == xy[2 * q - 2] && xy[2 * q + 1] == xy[2 * q - 1])
            return fail(RP_ERR_ARG, "reference polyline has a repeated vertex");
    const int n = n_pts + 2;
    if (int rc = wait_external_readers(ctx)) return rc;
    RP_CUDA(cudaStreamSynchronize(ctx->stream));           // launches in flight may still read the old tables
    DevBuf d_xy, d_scratch;
    if (int rc = d_xy.ensure((size_t)n_pts * 2 * sizeof(double))) return rc;
    if (int rc = d_scratch.ensure((size_t)n * 4 * sizeof(double))) { d_xy.release(); return rc; }
    if (int rc = ctx->d_ref.ensure((size_t)9 * n * sizeof(double))) { d_xy.release(); d_scratch.release(); return rc; }
    cudaError_t e = cudaMemcpyAsync(d_xy.p, xy, (size_t)n_pts * 2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        rp::ref_tables_kernel<<<1, 256, 0, ctx->stream>>>(n_pts, d_xy.as<double>(), eps2, ctx->d_ref.as<double>(), d_scratch.as<double>());
        e = cudaGetLastError();
    }
    double ends[2] = {0., 0.};           // ref_pos[0], ref_pos[n - 1]: the O(1) segment guess needs their span on the host
    if (e == cudaSuccess) e = cudaMemcpyAsync(&ends[0], ctx->d_ref.as<double>(), sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&ends[1], ctx->d_ref.as<double>() + (n - 1), sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    d_xy.release();
    d_scratch.release();
    if (e != cudaSuccess) return fail(RP_ERR_CUDA, cudaGetErrorString(e));
    if (!(ends[1] > ends[0])) return fail(RP_ERR_ARG, "degenerate reference polyline");
    ctx->ref_n = n;
    ctx->ref_same_s = 1;                 // the frame's arc length IS ref_pos (utils_coordinate_system.py:113)
    ctx->ref_limit = proj_limit;
    ctx->ref_inv_step = (double)(n - 1) / (ends[1] - ends[0]);
    ctx->ps_inv_step = ctx->ref_inv_step;
    ctx->segs_dirty = true;
    ++ctx->tables_version;
    ctx->have_ref = true;
    return RP_OK;
}

int rp_ctx_get_reference(rp_ctx* ctx, int capacity, int* n_pts, double* ref_pos, double* ref_theta, double* ref_curv,
                         double* ref_curv_d, double* path_xy, double* path_s, double* path_normal_xy) {
    if (int rc = bind(ctx)) return rc;
    if (!ctx->have_ref) return fail(RP_ERR_STATE, "reference tables not set");
    if (!n_pts) return fail(RP_ERR_ARG, "null n_pts");
    const int n = ctx->ref_n;
    *n_pts = n;
    if (capacity <= 0) return RP_OK;                       // size query
    if (capacity < n) return fail(RP_ERR_ARG, "reference buffers too small");
    std::vector<double> pack((size_t)9 * n);
    RP_CUDA(cudaMemcpyAsync(pack.data(), ctx->d_ref.p, pack.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    RP_CUDA(cudaStreamSynchronize(ctx->stream));
    // (layout of d_ref: pos | theta | curv | curv_d | px | py | nx | ny | ps; ps is omitted when it equals pos)
    auto col = [&](int a) { return &pack[(size_t)a * n]; };
    if (ref_pos) std::memcpy(ref_pos, col(0), n * sizeof(double));
    if (ref_theta) std::memcpy(ref_theta, col(1), n * sizeof(double));
    if (ref_curv) std::memcpy(ref_curv, col(2), n * sizeof(double));
    if (ref_curv_d) std::memcpy(ref_curv_d, col(3), n * sizeof(double));
    for (int q = 0; q < n; ++q) {
        if (path_xy) { path_xy[2 * q] = col(4)[q]; path_xy[2 * q + 1] = col(5)[q]; }
        if (path_normal_xy) { path_normal_xy[2 * q] = col(6)[q]; path_normal_xy[2 * q + 1] = col(7)[q]; }
    }
    if (path_s) std::memcpy(path_s, col(8), n * sizeof(double));
    return RP_OK;
}

int rp_ctx_set_obstacles(rp_ctx* ctx, int n_static, const double* static_obb, int n_dyn, const int32_t* dyn_t0,
                         const int32_t* dyn_len, const double* dyn_obb, int n_tri, const double* tris,
                         double cell_size) {
    if (int rc = bind(ctx)) return rc;
    if (n_static < 0 || n_dyn < 0 || n_tri < 0) return fail(RP_ERR_ARG, "negative obstacle count");
    if ((n_static && !static_obb) || (n_dyn && (!dyn_t0 || !dyn_len)) || (n_tri && !tris))
        return fail(RP_ERR_ARG, "null obstacle array");
    size_t total = 0;
    for (int o = 0; o < n_dyn; ++o) {
        if (dyn_len[o] < 0) return fail(RP_ERR_ARG, "negative dynamic obstacle length");
        total += (size_t)dyn_len[o];
    }
    if (total && !dyn_obb) return fail(RP_ERR_ARG, "null dynamic obstacle boxes");
    ctx->h_static.assign(static_obb, static_obb + (size_t)n_static * 5);
    ctx->h_dyn_t0.assign(dyn_t0, dyn_t0 + n_dyn);
    ctx->h_dyn_len.assign(dyn_len, dyn_len + n_dyn);
    ctx->h_dyn.assign(dyn_obb, dyn_obb + total * 5);
    ctx->h_tri.assign(tris, tris + (size_t)n_tri * 6);
    ctx->cell_size = cell_size > 0 ? cell_size : 0.0;
    ctx->obstacles_dirty = true;
    return RP_OK;
}

int rp_ctx_set_kernel_policy(rp_ctx* ctx, int policy) {
    if (!ctx) return fail(RP_ERR_ARG, "null context");
    if (policy < RP_KERNEL_AUTO || policy > RP_KERNEL_CANDIDATE_MAJOR) return fail(RP_ERR_ARG, "invalid kernel policy");
    ctx->kernel_policy = policy;
    ctx->segs_dirty = true;
    return RP_OK;
}

int rp_set_candidate_range(rp_ctx* ctx, int first, int count) {
    if (!ctx) return fail(RP_ERR_ARG, "null context");
    if (count >= 0 && first < 0) return fail(RP_ERR_ARG, "negative range start");
    ctx->range_first = count < 0 ? 0 : first;
    ctx->range_count = count;
    ctx->stripe_world = 0;                     // a contiguous range replaces a stripe
    ctx->stripe_rank = 0;
    ctx->segs_dirty = true;
    return RP_OK;
}

int rp_set_candidate_stripe(rp_ctx* ctx, int rank, int world) {
    if (!ctx) return fail(RP_ERR_ARG, "null context");
    if (world > 1 && (rank < 0 || rank >= world)) return fail(RP_ERR_ARG, "stripe rank out of range");
    ctx->stripe_rank = world > 1 ? rank : 0;
    ctx->stripe_world = world > 1 ? world : 0;
    ctx->range_first = 0;
    ctx->range_count = -1;
    ctx->segs_dirty = true;
    return RP_OK;
}

int rp_grid_upload(rp_ctx* ctx, const rp_plan_inputs* in, int n_t, const double* t, const int32_t* traj_len, int n_lon,
                   const double* lon, int n_d, const double* d) {
    if (int rc = bind(ctx)) return rc;
    if (int rc = check_inputs(in)) return rc;
    if (n_t < 0 || n_lon < 0 || n_d < 0) return fail(RP_ERR_ARG, "negative sample count");
    if ((n_t && (!t || !traj_len)) || (n_lon && !lon) || (n_d && !d)) return fail(RP_ERR_ARG, "null sample array");
    const long long n_cand = (long long)n_t * n_lon * n_d;
    if (n_cand > 0x7fffffffLL / 8) return fail(RP_ERR_ARG, "bundle too large");
    for (int q = 0; q < n_t; ++q)
        if (traj_len[q] < 1 || traj_len[q] > in->N + 1) return fail(RP_ERR_ARG, "traj_len out of [1, N+1]");
    ctx->in = *in;
    ctx->mode = 0;
    ctx->cycle_valid = false;
    ctx->n_t = n_t; ctx->n_lon = n_lon; ctx->n_d = n_d;
    ctx->n_cand = (int)n_cand;
    ctx->off_t = 0;
    ctx->off_lon = ctx->off_t + (size_t)n_t * sizeof(double);
    ctx->off_d = ctx->off_lon + (size_t)n_lon * sizeof(double);
    ctx->off_len = ctx->off_d + (size_t)n_d * sizeof(double);
    const size_t bytes = ctx->off_len + (size_t)n_t * sizeof(int);
    if (int rc = ctx->h_stage.ensure(bytes)) return rc;
    if (int rc = ctx->d_samples.ensure(bytes)) return rc;
    if (ctx->stage_pending) RP_CUDA(cudaEventSynchronize(ctx->ev_stage));     // last copy out of this buffer
    char* hs = static_cast<char*>(ctx->h_stage.p);
    if (n_t) std::memcpy(hs + ctx->off_t, t, (size_t)n_t * sizeof(double));
    if (n_lon) std::memcpy(hs + ctx->off_lon, lon, (size_t)n_lon * sizeof(double));
    if (n_d) std::memcpy(hs + ctx->off_d, d, (size_t)n_d * sizeof(double));
    if (n_t) std::memcpy(hs + ctx->off_len, traj_len, (size_t)n_t * sizeof(int));
    if (bytes) {
        RP_CUDA(cudaMemcpyAsync(ctx->d_samples.p, hs, bytes, cudaMemcpyHostToDevice, ctx->stream));
        RP_CUDA(cudaEventRecord(ctx->ev_stage, ctx->stream));
        ctx->stage_pending = true;
    }
    // the work decomposition depends on the grid's shape, the horizon and what selects the kernel -- not on the sample
    // values: consecutive replanning cycles of one planner reuse it
    const bool same_shape = ctx->geom_key_valid && ctx->geom_key_mode == 0 && ctx->geom_key_n[0] == n_t && ctx->geom_key_n[1] == n_lon &&
                            ctx->geom_key_n[2] == n_d && ctx->geom_key_n[3] == in->N && ctx->geom_key_n[4] == in->want_all_states &&
                            ctx->geom_key_n[5] == in->draw_all && ctx->geom_key_n[6] == in->cost_kind &&
                            ctx->geom_key_n[7] == in->has_desired_speed && ctx->geom_key_n[8] == in->has_desired_s &&
                            ctx->h_traj_len.size() == (size_t)n_t && std::equal(traj_len, traj_len + n_t, ctx->h_traj_len.begin());
    if (!same_shape) {
        ctx->h_traj_len.assign(traj_len, traj_len + n_t);
        ctx->segs_dirty = true;
        ctx->geom_key_valid = true;
        ctx->geom_key_mode = 0;
        const int key[9] = {n_t, n_lon, n_d, in->N, in->want_all_states, in->draw_all, in->cost_kind, in->has_desired_speed,
                            in->has_desired_s};
        std::copy(key, key + 9, ctx->geom_key_n);
    }
    ctx->have_inputs = true;
    ctx->have_plan = false;
    return RP_OK;
}

static int launch_plan(rp_ctx* ctx) {
    if (int rc = check_ready(ctx)) return rc;
    const int Np1 = ctx->in.N + 1;
    const int n = ctx->n_cand;
    int first = 0, count = n;
    shard_extent(ctx, first, count);
    const bool sharded = ctx->range_count >= 0 || stripe_on(ctx);
    const rp::Stripe sm = stripe_of(ctx);
    // every argument check comes before the first kernel of the chain: a rank that fails here has enqueued nothing
    // and has not advanced the peer epoch
    if (count > 0 && ctx->in.continuous_collision_check && ctx->in.check_collision && sharded)
        return fail(RP_ERR_ARG, "continuous collision check is not available for sharded bundles");
    if (int rc = ctx->d_cost.ensure((size_t)std::max(n, 1) * sizeof(double))) return rc;
    if (int rc = ctx->d_info.ensure((size_t)std::max(n, 1) * sizeof(int))) return rc;
    ctx->res_states_bytes = (size_t)14 * Np1 * sizeof(double);
    ctx->h_states_valid = false;
    if (ctx->d_result.cap < rp_ctx::kResBytes + ctx->res_states_bytes) {
        RP_CUDA(cudaStreamSynchronize(ctx->stream));           // a previous result copy may still read the old block
        if (int rc = ctx->d_result.ensure(rp_ctx::kResBytes + ctx->res_states_bytes)) return rc;
        if (int rc = ctx->h_result.ensure(rp_ctx::kResBytes + ctx->res_states_bytes)) return rc;
    }
    ctx->states_all_valid = false;
    if (ctx->in.want_all_states) {
        if (int rc = ctx->d_states_all.ensure((size_t)std::max(n, 1) * 14 * Np1 * sizeof(double))) return rc;
    }
    // replanning-size bundles: the main launch writes the states of every kept candidate (a few MB at most) and ONE
    // block selects the winner and gathers its states -- 3 launches per cycle instead of 6
    // sharded bundle with an open peer group: the selection chain ends with the exchange over peer-mapped memory
    const bool peer_mode = ctx->peer_ready && sharded;
    ctx->peer_mode_last = peer_mode;
    const bool small_path = !peer_mode && !stripe_on(ctx) && count > 0 && (long long)n * Np1 <= 262144 && !use_cand_kernel(ctx, count);
    ctx->small_path_last = small_path;
    if (small_path && !ctx->in.want_all_states) {
        if (int rc = ctx->d_states_all.ensure((size_t)n * 14 * Np1 * sizeof(double))) return rc;
    }
    ctx->ev = ctx->ev_ring[ctx->n_launches % rp_ctx::kEvRing];
    ctx->ev_recorded[ctx->n_launches % rp_ctx::kEvRing] = ctx->stage_timing;
    if (ctx->stage_timing) cudaEventRecord(ctx->ev[0], ctx->stream);
    // dynamic-obstacle rows of the candidate-major kernel ride along with the coefficient solve (one launch)
    const bool cand_main = count > 0 && use_cand_kernel(ctx, count);
    const int dyn_total = (cand_main && ctx->obs.n_dyn > 0 && ctx->in.check_collision) ? Np1 * ctx->obs.n_dyn : 0;
    if (dyn_total > 0)
        if (int rc = ctx->d_dyn_rows.ensure((size_t)dyn_total * sizeof(float4))) return rc;
    const double hl_ = 0.5 * ctx->veh.length, hw_ = 0.5 * ctx->veh.width;
    const float r_ego_f_up = std::nextafterf((float)std::sqrt(hl_ * hl_ + hw_ * hw_), std::numeric_limits<float>::infinity());
    const float wb_rear_f_up = std::nextafterf((float)std::fabs(ctx->veh.wb_rear_axle), std::numeric_limits<float>::infinity());
    // lateral table of the candidate-major kernel: high-velocity grid bundles whose chunks of 32 candidates share one
    // (t, lon) pair (the ONE_GROUP instantiation) read d, d_dot, d_ddot per step instead of evaluating three polynomials
    const bool use_lat_rows = cand_main && ctx->mode == 0 && !ctx->in.low_vel_mode && ctx->n_d > 0 && ctx->n_d % 32 == 0 &&
                              first % 32 == 0 && count % 32 == 0;
    if (use_lat_rows)
        if (int rc = ctx->d_lat_rows.ensure((size_t)ctx->n_t * Np1 * ctx->n_d * 4 * sizeof(double))) return rc;
    bool dyn_rows_done = false;         // the prep launch ran: coefficients, dynamic-obstacle rows, scratch words reset
    if (ctx->mode == 0 && n > 0) {
        const int n_lon_sys = ctx->n_t * ctx->n_lon;
        const int n_lat_sys = ctx->in.low_vel_mode ? n : ctx->n_t * ctx->n_d;
        if (int rc = ctx->d_lon_coef.ensure((size_t)n_lon_sys * 6 * sizeof(double))) return rc;
        if (int rc = ctx->d_lat_coef.ensure((size_t)n_lat_sys * 6 * sizeof(double))) return rc;
        if (int rc = ctx->d_lat_tau.ensure((size_t)n_lat_sys * sizeof(double))) return rc;
        const char* sb = static_cast<const char*>(ctx->d_samples.p);
        rp::PrepArgs A{};
        A.n_t = ctx->n_t; A.n_lon = ctx->n_lon; A.n_d = ctx->n_d; A.low_vel = ctx->in.low_vel_mode; A.lon_mode = ctx->in.lon_mode;
        A.t = reinterpret_cast<const double*>(sb + ctx->off_t);
        A.lon = reinterpret_cast<const double*>(sb + ctx->off_lon);
        A.d = reinterpret_cast<const double*>(sb + ctx->off_d);
        A.x0s = ctx->in.x0_lon[0]; A.x0sd = ctx->in.x0_lon[1]; A.x0sdd = ctx->in.x0_lon[2];
        A.x0d = ctx->in.x0_lat[0]; A.x0dd = ctx->in.x0_lat[1]; A.x0ddd = ctx->in.x0_lat[2];
        A.lon_coef = ctx->d_lon_coef.as<double>(); A.lat_coef = ctx->d_lat_coef.as<double>(); A.lat_tau = ctx->d_lat_tau.as<double>();
        A.n_coeff_blocks = (n_lon_sys + n_lat_sys + 127) / 128;
        A.obs = ctx->obs;
        A.x0_time_step = ctx->in.x0_time_step; A.factor = ctx->in.factor; A.Np1 = Np1;
        A.r_ego_f_up = r_ego_f_up; A.wb_rear_f_up = wb_rear_f_up;
        A.dyn_rows = ctx->d_dyn_rows.as<float4>();
        A.n_dyn_blocks = (dyn_total + 127) / 128;
        A.lat_rows = use_lat_rows ? ctx->d_lat_rows.as<double>() : nullptr;
        A.traj_len = reinterpret_cast<const int*>(sb + ctx->off_len);
        A.dt = ctx->in.dt;
        const int n_lat_row_threads = use_lat_rows ? ctx->n_t * ((Np1 + 7) / 8) * ctx->n_d : 0;
        if (int rc = ctx->d_argmin.ensure(sizeof(rp::ArgminScratch))) return rc;
        if (int rc = ctx->d_work.ensure(sizeof(int) * rp::kWorkWords)) return rc;
        if (int rc = ctx->d_best.ensure(sizeof(unsigned long long))) return rc;
        A.argmin_counts = reinterpret_cast<int*>(ctx->d_argmin.p);
        A.work_counter = ctx->d_work.as<int>();
        A.best_bits = ctx->d_best.as<unsigned long long>();
        rp::prep_kernel<<<A.n_coeff_blocks + A.n_dyn_blocks + (n_lat_row_threads + 127) / 128, 128, 0, ctx->stream>>>(A);
        RP_CUDA(cudaGetLastError());
        dyn_rows_done = true;
    }
    if (ctx->stage_timing) cudaEventRecord(ctx->ev[1], ctx->stream);
    if (count > 0) {
        if (int rc = prepare_main_geometry(ctx, first, count)) return rc;
        PlanParams P{};
        fill_common(ctx, P, ctx->main_geom, ctx->d_segs.as<rp::Segment>());
        P.index = nullptr;
        P.cost = ctx->d_cost.as<double>();
        P.info = ctx->d_info.as<int>();
        P.states = (ctx->in.want_all_states || small_path) ? ctx->d_states_all.as<double>() : nullptr;
        P.states_by_slot = 0;
        P.stripe_rank = sm.rank;
        P.stripe_world = sm.world;
        if (ctx->in.check_collision == 2) {
            if (int rc = ctx->d_best.ensure(sizeof(unsigned long long))) return rc;
            if (!dyn_rows_done)             // (the prep launch of the grid form resets the per-cycle scratch words)
                RP_CUDA(cudaMemsetAsync(ctx->d_best.p, 0x7f, sizeof(unsigned long long), ctx->stream));   // ~1.4e306
            P.best_bits = ctx->d_best.as<unsigned long long>();
        }
        if (ctx->main_is_cand) {
            if (int rc = ctx->d_work.ensure(sizeof(int) * rp::kWorkWords)) return rc;
            if (!dyn_rows_done) RP_CUDA(cudaMemsetAsync(ctx->d_work.p, 0, sizeof(int) * rp::kWorkWords, ctx->stream));
            P.work_counter = ctx->d_work.as<int>();
            P.n_acc_rows = cand_acc_rows(ctx->in);
            P.dyn_rows = nullptr;
            if (dyn_total > 0) {
                if (!dyn_rows_done)         // list form: no coefficient launch to ride along with
                    rp::dyn_rows_kernel<<<(dyn_total + 127) / 128, 128, 0, ctx->stream>>>(ctx->obs, ctx->in.x0_time_step, ctx->in.factor,
                                                                                           Np1, P.r_ego_f_up, P.wb_rear_f_up,
                                                                                           ctx->d_dyn_rows.as<float4>());
                P.dyn_rows = ctx->d_dyn_rows.as<float4>();
            }
            const Geometry& G = ctx->main_geom;
            P.lat_rows = (use_lat_rows && ctx->main_one_group) ? ctx->d_lat_rows.as<double>() : nullptr;
            // the reference's lazy collision pass (check_collision = 2): the march stores the ego boxes (32 bytes per
            // candidate-timestep) and the checks run afterwards, all time steps of a candidate at once, for the
            // candidates that can be ranked before the winner (rp_cand.cuh, deferred_collision_kernel)
            const bool defer = ctx->in.check_collision == 2 && ctx->in.cost_kind != RP_COST_NONE &&
                               (size_t)((n + 31) / 32) * Np1 * 32 * 4 * sizeof(double) <= kDeferMaxBytes &&
                               (n + 31) / 32 < (1 << rp::kDeferTileBits);
            P.pose = nullptr;
            if (defer) {
                if (int rc = ctx->d_pose.ensure((size_t)((n + 31) / 32) * Np1 * 32 * 4 * sizeof(double))) return rc;
                const size_t n_tiles = (size_t)(n + 31) / 32;
                if (int rc = ctx->d_defer_list.ensure((n_tiles + (size_t)std::max(n, 1)) * sizeof(int))) return rc;    // tiles, then candidates
                if (ctx->d_defer_mask.cap < n_tiles * sizeof(unsigned)) {       // (the checker leaves the masks it used clear)
                    if (int rc = ctx->d_defer_mask.ensure(n_tiles * sizeof(unsigned))) return rc;
                    RP_CUDA(cudaMemsetAsync(ctx->d_defer_mask.p, 0, ctx->d_defer_mask.cap, ctx->stream));
                }
                P.pose = ctx->d_pose.as<double>();
                P.defer_list = ctx->d_defer_list.as<int>();
                P.defer_list2 = P.defer_list + n_tiles;
                P.defer_tag = 0;
                P.defer_count = ctx->d_work.as<int>() + 1;
                P.defer_mask = ctx->d_defer_mask.as<unsigned>();
            }
            if (defer) {
                if (P.lat_rows) rp::cand_kernel<RP_CAND_THREADS, true, true, true><<<G.grid, G.threads, G.smem, ctx->stream>>>(P);
                else if (ctx->main_one_group) rp::cand_kernel<RP_CAND_THREADS, true, false, true><<<G.grid, G.threads, G.smem, ctx->stream>>>(P);
                else if (ctx->mode == 0) rp::cand_kernel<RP_CAND_THREADS, false, false, true, true><<<G.grid, G.threads, G.smem, ctx->stream>>>(P);
                else rp::cand_kernel<RP_CAND_THREADS, false, false, true><<<G.grid, G.threads, G.smem, ctx->stream>>>(P);
            } else {
                if (P.lat_rows) rp::cand_kernel<RP_CAND_THREADS, true, true><<<G.grid, G.threads, G.smem, ctx->stream>>>(P);
                else if (ctx->main_one_group) rp::cand_kernel<RP_CAND_THREADS, true><<<G.grid, G.threads, G.smem, ctx->stream>>>(P);
                else if (ctx->mode == 0) rp::cand_kernel<RP_CAND_THREADS, false, false, false, true><<<G.grid, G.threads, G.smem, ctx->stream>>>(P);
                else rp::cand_kernel<RP_CAND_THREADS, false><<<G.grid, G.threads, G.smem, ctx->stream>>>(P);
            }
            RP_CUDA(cudaGetLastError());
            if (defer) {
                const int n_tiles = (n + 31) / 32;
                const int check_blocks = std::max(1, std::min(n_tiles, 8 * ctx->num_sms));
                rp::deferred_collision_kernel<<<check_blocks, rp::kDeferThreads, 0, ctx->stream>>>(P, P.defer_list, P.defer_count);
                rp::deferred_gather_kernel<<<(count + 255) / 256, 256, 0, ctx->stream>>>(P, first, count);
                rp::deferred_collision_list_kernel<<<check_blocks, rp::kDeferThreads, 0, ctx->stream>>>(P, P.defer_list2, P.defer_count + 1);
                RP_CUDA(cudaGetLastError());
            }
        } else if (int rc = launch_fused(ctx, P, ctx->main_geom)) return rc;
        ctx->states_all_valid = ctx->in.want_all_states != 0;
    }
    if (ctx->stage_timing) cudaEventRecord(ctx->ev[2], ctx->stream);
    rp::PlanResultDev* dres = ctx->d_result.as<rp::PlanResultDev>();
    const bool side_count = !small_path && !ctx->stage_timing && ctx->side_stream != nullptr;
    if (small_path) {
        rp::select_small_kernel<<<1, 256, 0, ctx->stream>>>(ctx->d_cost.as<double>(), ctx->d_info.as<int>(), first, count,
                                                            ctx->d_states_all.as<double>(), Np1, dres, ctx->d_states_one());
    } else {
        if (int rc = ctx->d_argmin.ensure(sizeof(rp::ArgminScratch))) return rc;
        rp::ArgminScratch* sc = ctx->d_argmin.as<rp::ArgminScratch>();
        if (!dyn_rows_done) RP_CUDA(cudaMemsetAsync(sc, 0, sizeof(int) * 16, ctx->stream));
        const int nb = std::max(1, std::min(512, std::min(2 * ctx->num_sms, (count + 255) / 256)));
        // (without a peer exchange the last partial block merges: no separate merge launch)
        rp::argmin_partial_kernel<<<nb, 256, 0, ctx->stream>>>(ctx->d_cost.as<double>(), ctx->d_info.as<int>(), first, count, sc, sm,
                                                               peer_mode ? nullptr : dres);
        // the count of colliders ranked before the winner (and, sharded, its exchange) needs the merged winner only, like
        // the winner-state launch below: the two run side by side (the count on the side stream, joined before anything
        // else touches the result block).  With stage timing on, the chain stays serial so that the stages add up.
        cudaStream_t cs = ctx->stream;
        if (peer_mode) {
            // the shard's merge block goes straight on to the exchange of the shard records (no separate launch)
            const unsigned long long epoch = ++ctx->peer_epoch;
            if (int rc = ctx->d_peer_table.ensure(sizeof(rp::PeerTable))) return rc;
            if (ctx->peer_table_dirty) {
                RP_CUDA(cudaMemcpyAsync(ctx->d_peer_table.p, &ctx->peer_table, sizeof(rp::PeerTable), cudaMemcpyHostToDevice, ctx->stream));
                ctx->peer_table_dirty = false;
            }
            rp::argmin_merge_kernel<<<1, 512, 0, ctx->stream>>>(sc, nb, count, dres, ctx->d_peer_table.as<rp::PeerTable>(), epoch);
            if (side_count) {
                RP_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
                RP_CUDA(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
                cs = ctx->side_stream;
            }
            rp::peer_count_kernel<<<nb, 256, 0, cs>>>(ctx->peer_table, epoch, ctx->d_cost.as<double>(), ctx->d_info.as<int>(),
                                                      first, count, dres, sm);
        } else {
            if (side_count) {
                RP_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
                RP_CUDA(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
                cs = ctx->side_stream;
            }
            rp::count_before_result_kernel<<<nb, 256, 0, cs>>>(ctx->d_cost.as<double>(), ctx->d_info.as<int>(), first, count, dres, sm);
        }
        if (side_count) RP_CUDA(cudaEventRecord(ctx->ev_join, ctx->side_stream));
    }
    RP_CUDA(cudaGetLastError());
    if (ctx->stage_timing) cudaEventRecord(ctx->ev[3], ctx->stream);
    // winner's 14 x (N+1) state block; the winner index never leaves the device
    if ((count > 0 || peer_mode) && !small_path) {          // peer mode: the GLOBAL winner, on every rank
        if (int rc = launch_states_for_index(ctx, &dres->r.winner, 1, ctx->d_states_one())) return rc;
    }
    if (side_count) RP_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    if (count > 0 && ctx->in.continuous_collision_check && ctx->in.check_collision) {
        rp::continuous_check_kernel<<<1, 128, 0, ctx->stream>>>(ctx->obs, ctx->d_states_one(), Np1, ctx->in.x0_time_step,
                                                                0.5 * ctx->veh.length, 0.5 * ctx->veh.width, ctx->veh.wb_rear_axle,
                                                                dres, ctx->d_info.as<int>());
        RP_CUDA(cudaGetLastError());
    }
    if (ctx->stage_timing) cudaEventRecord(ctx->ev[4], ctx->stream);
    ++ctx->n_launches;
    ctx->have_plan = true;
    return RP_OK;
}

int rp_grid_launch(rp_ctx* ctx) {
    if (int rc = bind(ctx)) return rc;
    if (!ctx->have_inputs || ctx->mode != 0) return fail(RP_ERR_STATE, "rp_grid_upload must precede rp_grid_launch");
    return launch_plan(ctx);
}

int rp_grid_result(rp_ctx* ctx, rp_plan_result* out) {
    if (int rc = bind(ctx)) return rc;
    if (!out) return fail(RP_ERR_ARG, "null result");
    if (!ctx->have_plan) return fail(RP_ERR_STATE, "no plan launched");
    // result and winner states in ONE block: rp_fetch_states(winner) is then served from the pinned copy.  A one-block
    // kernel behind the chain writes the block into pinned host memory and raises a host flag; the host spins on the flag
    // (no copy-engine hop, no event wake-up: ~6 us less between the last kernel and the caller)
    if (!ctx->h_flag.p) {
        if (int rc = ctx->h_flag.ensure(64)) return rc;
        std::memset(ctx->h_flag.p, 0, 64);
    }
    const unsigned long long epoch = ++ctx->res_epoch;
    const size_t bytes = rp_ctx::kResBytes + ctx->res_states_bytes;
    static_assert(rp_ctx::kResBytes % 16 == 0, "result block is copied in 16-byte words");
    rp::publish_result_kernel<<<1, 256, 0, ctx->stream>>>(static_cast<const int4*>(ctx->d_result.p), static_cast<int4*>(ctx->h_result.p),
                                                         (int)(bytes / 16), static_cast<unsigned long long*>(ctx->h_flag.p), epoch);
    RP_CUDA(cudaGetLastError());
    const volatile unsigned long long* flag = static_cast<const volatile unsigned long long*>(ctx->h_flag.p);
    for (unsigned spins = 0; *flag != epoch; ++spins) {
        if ((spins & 0x3fffu) == 0x3fffu) {
            const cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q == cudaSuccess) {
                if (*flag == epoch) break;
                return fail(RP_ERR_CUDA, "rp_grid_result: the chain finished without publishing its result");
            }
            if (q != cudaErrorNotReady) return fail(RP_ERR_CUDA, std::string("rp_grid_result: ") + cudaGetErrorString(q));
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    ctx->h_states_valid = true;
    *out = static_cast<rp::PlanResultDev*>(ctx->h_result.p)->r;
    if (ctx->peer_mode_last && static_cast<rp::PlanResultDev*>(ctx->h_result.p)->peer_error) {
        // the ranks' epochs can no longer be trusted to agree: the group is broken until every rank re-creates it
        ctx->peer_ready = false;
        return fail(RP_ERR_PEER, "peer exchange: a wait on another rank timed out or a peer reported a failed cycle; the "
                                 "group is closed -- rp_peer_close / rp_peer_create / rp_peer_open on EVERY rank to resume");
    }
    return RP_OK;
}

int rp_plan_grid(rp_ctx* ctx, const rp_plan_inputs* in, int n_t, const double* t, const int32_t* traj_len, int n_lon,
                 const double* lon, int n_d, const double* d, rp_plan_result* out) {
    if (int rc = rp_grid_upload(ctx, in, n_t, t, traj_len, n_lon, lon, n_d, d)) return rc;
    if (int rc = rp_grid_launch(ctx)) return rc;
    return rp_grid_result(ctx, out);
}

// ---- one replanning cycle in one launch --------------------------------------------------------------------------
int rp_cycle_limits(int32_t* max_levels, int32_t* max_samples, int32_t* max_segments, int64_t* max_work) {
    if (max_levels) *max_levels = rp::kMaxLevels;
    if (max_samples) *max_samples = rp::kCycleSamples;
    if (max_segments) *max_segments = rp::kCycleSegs;
    if (max_work) *max_work = kCycleMaxWork;
    return RP_OK;
}

int rp_cycle_host_block(rp_ctx* ctx, void** states, int64_t* n_doubles) {
    if (!ctx || !states || !n_doubles) return fail(RP_ERR_ARG, "null argument");
    if (!ctx->h_cycle) return fail(RP_ERR_STATE, "no cycle launch yet");
    *states = reinterpret_cast<char*>(ctx->h_cycle) + rp::kCycleStatesOffset;
    *n_doubles = (int64_t)((ctx->cycle_bytes - rp::kCycleStatesOffset) / sizeof(double));
    return RP_OK;
}

#ifdef RP_CYCLE_TIMING
}  // extern "C"
#include <chrono>
namespace {
struct BigParam { char bytes[7424]; };
struct SmallParam { char bytes[64]; };
template <class PT>
__global__ void null_flag_kernel(const __grid_constant__ PT p, unsigned long long* flag_host, unsigned long long epoch) {
    if (threadIdx.x == 0 && blockIdx.x == gridDim.x - 1) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>(flag_host) = epoch + (unsigned long long)p.bytes[0];
    }
}
}  // namespace
extern "C" {
// launch floor: null kernel with a small / large parameter block, completion seen through a mapped flag or a stream sync
int rp_debug_launch_floor(rp_ctx* ctx, int iters, double* us4) {
    if (int rc = bind(ctx)) return rc;
    unsigned long long* h = nullptr;
    cudaHostAlloc((void**)&h, 64, cudaHostAllocMapped);
    unsigned long long* d = nullptr;
    cudaHostGetDevicePointer((void**)&d, h, 0);
    *h = 0;
    BigParam big{};
    SmallParam small{};
    unsigned long long epoch = 0;
    for (int variant = 0; variant < 4; ++variant) {
        double total = 0;
        for (int it = 0; it < iters + 20; ++it) {
            ++epoch;
            auto t0 = std::chrono::steady_clock::now();
            if (variant & 1) null_flag_kernel<BigParam><<<9, 64, 0, ctx->stream>>>(big, d, epoch);
            else null_flag_kernel<SmallParam><<<9, 64, 0, ctx->stream>>>(small, d, epoch);
            if (variant & 2) cudaStreamSynchronize(ctx->stream);
            else while (*reinterpret_cast<volatile unsigned long long*>(h) != epoch) { }
            auto t1 = std::chrono::steady_clock::now();
            if (it >= 20) total += std::chrono::duration<double, std::micro>(t1 - t0).count();
        }
        us4[variant] = total / iters;
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFreeHost(h);
    return RP_OK;
}
int rp_debug_block_times(rp_ctx* ctx, long long* out, int n_blocks, int* grid, int* threads) {
    if (grid) *grid = ctx->cycle_geom.grid;
    if (threads) *threads = ctx->cycle_geom.threads;
    return cudaMemcpyFromSymbol(out, rp::g_block_t, sizeof(long long) * 3 * n_blocks) == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}
int rp_debug_stamps(long long* out32) {
    return cudaMemcpyFromSymbol(out32, rp::g_stamps, sizeof(long long) * 32) == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}
#endif

int rp_select_level(rp_ctx* ctx, int level) {
    if (!ctx) return fail(RP_ERR_ARG, "null context");
    if (!ctx->cycle_valid) return fail(RP_ERR_STATE, "rp_select_level: the last plan was not an rp_plan_levels call");
    if (level < 0 || level >= ctx->cyc_n_eval) return fail(RP_ERR_ARG, "rp_select_level: that level was not evaluated");
    ctx->cyc_sel = level;
    const rp::LevelDesc& L = ctx->cyc_lv[level];
    ctx->mode = 0;
    ctx->n_t = L.n_t; ctx->n_lon = L.n_lon; ctx->n_d = L.n_d;
    ctx->n_cand = L.count;
    return RP_OK;
}

int rp_plan_levels(rp_ctx* ctx, const rp_plan_inputs* in, int n_levels, const int32_t* n_t, const int32_t* n_lon,
                   const int32_t* n_d, const double* t_cat, const int32_t* traj_len_cat, const double* lon_cat,
                   const double* d_cat, rp_plan_result* out, int32_t* n_evaluated, int32_t* chosen) {
    if (int rc = bind(ctx)) return rc;
    if (int rc = check_inputs(in)) return rc;
    if (n_levels < 1 || n_levels > rp::kMaxLevels) return fail(RP_ERR_ARG, "rp_plan_levels: 1 .. 4 levels");
    if (!n_t || !n_lon || !n_d || !t_cat || !traj_len_cat || !lon_cat || !d_cat || !out) return fail(RP_ERR_ARG, "null array");
    if (ctx->range_count >= 0 || ctx->stripe_world > 1) return fail(RP_ERR_ARG, "rp_plan_levels does not shard (reset rp_set_candidate_range)");
    if (in->continuous_collision_check) return fail(RP_ERR_ARG, "rp_plan_levels does not run the continuous collision check");
    if (in->cost_kind == RP_COST_NONE && n_levels > 1)
        return fail(RP_ERR_ARG, "rp_plan_levels: without a device cost the host selects, one level at a time");
    const int Np1 = in->N + 1;
    if (Np1 > 256) return fail(RP_ERR_ARG, "rp_plan_levels: N + 1 must be <= 256");
    if (int rc = check_ready(ctx)) return rc;
    // ---- the launch's argument block: levels, segments (one per level and sampled t), samples -------------------
    rp::CycleArgs A{};
    A.n_levels = n_levels;
    std::vector<rp::Segment> segs;
    int k0 = 0, n_samp = 0;
    size_t ot = 0, ol = 0, od = 0;
    for (int lv = 0; lv < n_levels; ++lv) {
        if (n_t[lv] < 0 || n_lon[lv] < 0 || n_d[lv] < 0) return fail(RP_ERR_ARG, "negative sample count");
        const long long cnt = (long long)n_t[lv] * n_lon[lv] * n_d[lv];
        if (n_samp + n_t[lv] + n_lon[lv] + n_d[lv] > rp::kCycleSamples) return fail(RP_ERR_ARG, "rp_plan_levels: too many samples for one launch");
        if ((long long)(k0 + cnt) * Np1 > kCycleMaxWork) return fail(RP_ERR_ARG, "rp_plan_levels: too many candidates for one launch");
        rp::LevelDesc& L = A.lv[lv];
        L.k0 = k0; L.count = (int)cnt;
        L.n_t = n_t[lv]; L.n_lon = n_lon[lv]; L.n_d = n_d[lv];
        L.off_t = n_samp; L.off_lon = n_samp + n_t[lv]; L.off_d = L.off_lon + n_lon[lv];
        std::memcpy(A.samples + L.off_t, t_cat + ot, (size_t)n_t[lv] * sizeof(double));
        std::memcpy(A.samples + L.off_lon, lon_cat + ol, (size_t)n_lon[lv] * sizeof(double));
        std::memcpy(A.samples + L.off_d, d_cat + od, (size_t)n_d[lv] * sizeof(double));
        n_samp += n_t[lv] + n_lon[lv] + n_d[lv];
        const int per_t = n_lon[lv] * n_d[lv];
        for (int it = 0; it < n_t[lv] && per_t > 0; ++it) {
            const int tl = traj_len_cat[ot + it];
            if (tl < 1 || tl > Np1) return fail(RP_ERR_ARG, "traj_len out of [1, N+1]");
            segs.push_back(rp::Segment{k0 + it * per_t, k0 + (it + 1) * per_t, tl, 0, 0, lv});
        }
        ctx->cyc_t[lv].assign(t_cat + ot, t_cat + ot + n_t[lv]);
        ctx->cyc_tl[lv].assign(traj_len_cat + ot, traj_len_cat + ot + n_t[lv]);
        ctx->cyc_lon[lv].assign(lon_cat + ol, lon_cat + ol + n_lon[lv]);
        ctx->cyc_d[lv].assign(d_cat + od, d_cat + od + n_d[lv]);
        ctx->cyc_lv[lv] = L;
        ot += (size_t)n_t[lv]; ol += (size_t)n_lon[lv]; od += (size_t)n_d[lv];
        k0 += (int)cnt;
    }
    const int n_total = k0;
    if ((int)segs.size() > rp::kCycleSegs) return fail(RP_ERR_ARG, "rp_plan_levels: too many sampled horizons for one launch");
    if (segs.empty()) segs.push_back(rp::Segment{0, 0, Np1, 0, 0, 0});
    ctx->in = *in;
    ctx->mode = 2;
    ctx->n_t = ctx->n_lon = ctx->n_d = 0;
    ctx->n_cand = n_total;
    ctx->geom_key_valid = false;
    ctx->segs_dirty = true;                       // the grid form's cached decomposition no longer describes the inputs
    // work decomposition: depends on the shapes only (not on the sample values), so consecutive cycles reuse it.
    // A replanning-size launch is latency bound: the smallest blocks that still give ONE wave (one or two warps per
    // scheduler instead of two blocks' worth) -- measured, profiles/README.md round 2.
    {
        std::vector<int> key;
        key.reserve(segs.size() * 3 + 4);
        key.push_back(Np1);
        key.push_back((int)ctx->tables_version);
        key.push_back(ctx->obs.n_dyn);
        for (const auto& sg : segs) { key.push_back(sg.k_begin); key.push_back(sg.k_end); key.push_back(sg.tl); }
        if (key != ctx->cyc_geom_key) {
            int max_tl = 1;
            for (const auto& sg : segs) max_tl = std::max(max_tl, std::min(sg.tl, Np1));
            int rc = RP_OK;
            for (int threads : {64, 128, 256}) {
                if (threads < max_tl) continue;
                std::vector<rp::Segment> trial = segs;
                rc = plan_geometry(ctx, Np1, trial, ctx->cycle_geom, true, threads);
                if (rc != RP_OK) break;
                if (ctx->cycle_geom.n_groups <= ctx->cycle_geom.grid || threads == 256) { segs.swap(trial); break; }
            }
            if (rc != RP_OK) return rc;
            ctx->cyc_geom_key.swap(key);
            ctx->cyc_geom_segs = segs;
        } else {
            segs = ctx->cyc_geom_segs;
        }
    }
    A.n_segs = (int)segs.size();
    std::memcpy(A.segs, segs.data(), segs.size() * sizeof(rp::Segment));
    // ---- buffers ---------------------------------------------------------------------------------------------
    if (int rc = ctx->d_cost.ensure((size_t)std::max(n_total, 1) * sizeof(double))) return rc;
    if (int rc = ctx->d_info.ensure((size_t)std::max(n_total, 1) * sizeof(int))) return rc;
    if (int rc = ctx->d_states_all.ensure((size_t)std::max(n_total, 1) * 14 * Np1 * sizeof(double))) return rc;
    if (int rc = ctx->d_cycle_res.ensure(sizeof(rp::PlanResultDev) * rp::kMaxLevels)) return rc;
    if (!ctx->d_ticket.p) {
        if (int rc = ctx->d_ticket.ensure(sizeof(unsigned))) return rc;
        if (int rc = ctx->d_best4.ensure(sizeof(unsigned long long) * rp::kMaxLevels)) return rc;
        RP_CUDA(cudaMemsetAsync(ctx->d_ticket.p, 0, sizeof(unsigned), ctx->stream));
        RP_CUDA(cudaMemsetAsync(ctx->d_best4.p, 0x7f, sizeof(unsigned long long) * rp::kMaxLevels, ctx->stream));
    }
    const size_t need = rp::kCycleStatesOffset + (size_t)14 * Np1 * sizeof(double);
    if (ctx->cycle_bytes < need) {
        RP_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->h_cycle) cudaFreeHost(ctx->h_cycle);
        ctx->h_cycle = nullptr;
        void* p = nullptr;
        RP_CUDA(cudaHostAlloc(&p, need, cudaHostAllocMapped));
        std::memset(p, 0, need);
        ctx->h_cycle = static_cast<rp::CycleOut*>(p);
        RP_CUDA(cudaHostGetDevicePointer(&ctx->d_cycle, p, 0));
        ctx->cycle_bytes = need;
    }
    A.out_host = ctx->d_cycle;
    A.ticket = ctx->d_ticket.as<unsigned>();
    A.epoch = ++ctx->cyc_epoch;
    PlanParams P{};
    fill_common(ctx, P, ctx->cycle_geom, nullptr);
    P.index = nullptr;
    P.cost = ctx->d_cost.as<double>();
    P.info = ctx->d_info.as<int>();
    P.states = ctx->d_states_all.as<double>();        // every kept candidate: the last block gathers the winner's block
    P.states_by_slot = 0;
    P.best_bits = in->check_collision == 2 ? ctx->d_best4.as<unsigned long long>() : nullptr;
    const Geometry& G = ctx->cycle_geom;
    if (A.n_segs <= rp::kCycleSegsSmall && n_samp <= rp::kCycleSamplesSmall) {
        rp::CycleArgsSmall S{};
        S.n_levels = A.n_levels; S.n_segs = A.n_segs;
        std::memcpy(S.lv, A.lv, sizeof(A.lv));
        std::memcpy(S.segs, A.segs, (size_t)A.n_segs * sizeof(rp::Segment));
        std::memcpy(S.samples, A.samples, (size_t)n_samp * sizeof(double));
        S.out_host = A.out_host; S.ticket = A.ticket; S.epoch = A.epoch;
        rp::cycle_kernel<256, rp::CycleArgsSmall><<<G.grid, G.threads, G.smem, ctx->stream>>>(P, S, ctx->d_cycle_res.as<rp::PlanResultDev>());
    } else {
        rp::cycle_kernel<256, rp::CycleArgs><<<G.grid, G.threads, G.smem, ctx->stream>>>(P, A, ctx->d_cycle_res.as<rp::PlanResultDev>());
    }
    RP_CUDA(cudaGetLastError());
    ++ctx->n_launches;
    // ---- wait for the result block (written by the kernel into mapped host memory) -----------------------------
    const volatile unsigned long long* flag = &ctx->h_cycle->flag;
    for (unsigned spins = 0; *flag != A.epoch; ++spins) {
        if ((spins & 0x3fffu) == 0x3fffu) {
            const cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q == cudaSuccess) {
                if (*flag == A.epoch) break;
                return fail(RP_ERR_CUDA, "rp_plan_levels: the cycle kernel finished without publishing its result");
            }
            if (q != cudaErrorNotReady) return fail(RP_ERR_CUDA, std::string("rp_plan_levels: ") + cudaGetErrorString(q));
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    const rp::CycleOut* H = ctx->h_cycle;
    ctx->cyc_n_levels = n_levels;
    ctx->cyc_chosen = H->chosen;
    ctx->cyc_n_eval = H->n_evaluated;
    ctx->cyc_coeff_level = -1;
    for (int lv = 0; lv < H->n_evaluated; ++lv) out[lv] = H->res[lv].r;
    if (n_evaluated) *n_evaluated = H->n_evaluated;
    if (chosen) *chosen = H->chosen;
    ctx->cycle_valid = true;
    ctx->have_inputs = false;                     // rp_grid_launch needs a fresh rp_grid_upload
    ctx->have_plan = true;
    ctx->peer_mode_last = false;
    ctx->small_path_last = false;
    ctx->states_all_valid = in->want_all_states != 0;
    ctx->h_states_valid = false;
    return rp_select_level(ctx, H->chosen);
}

int rp_plan_list(rp_ctx* ctx, const rp_plan_inputs* in, int n_cand, const double* coeffs_lon, const double* coeffs_lat,
                 const int32_t* traj_len, const uint8_t* skip, rp_plan_result* out) {
    if (int rc = bind(ctx)) return rc;
    if (int rc = check_inputs(in)) return rc;
    if (n_cand < 0) return fail(RP_ERR_ARG, "negative candidate count");
    if (n_cand && (!coeffs_lon || !coeffs_lat || !traj_len)) return fail(RP_ERR_ARG, "null candidate array");
    for (int q = 0; q < n_cand; ++q)
        if (traj_len[q] < 1 || traj_len[q] > in->N + 1) return fail(RP_ERR_ARG, "traj_len out of [1, N+1]");
    ctx->in = *in;
    ctx->mode = 1;
    ctx->cycle_valid = false;
    ctx->geom_key_valid = false;
    ctx->n_t = ctx->n_lon = ctx->n_d = 0;
    ctx->n_cand = n_cand;
    ctx->off_len = 0;
    const size_t nb = (size_t)std::max(n_cand, 1);
    if (int rc = ctx->d_samples.ensure(nb * sizeof(int))) return rc;
    if (int rc = ctx->d_lon_coef.ensure(nb * 6 * sizeof(double))) return rc;
    if (int rc = ctx->d_lat_coef.ensure(nb * 6 * sizeof(double))) return rc;
    if (n_cand) {
        RP_CUDA(cudaMemcpyAsync(ctx->d_samples.p, traj_len, (size_t)n_cand * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        RP_CUDA(cudaMemcpyAsync(ctx->d_lon_coef.p, coeffs_lon, (size_t)n_cand * 6 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        RP_CUDA(cudaMemcpyAsync(ctx->d_lat_coef.p, coeffs_lat, (size_t)n_cand * 6 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (skip && n_cand) {
        if (int rc = ctx->d_skip.ensure(nb)) return rc;
        RP_CUDA(cudaMemcpyAsync(ctx->d_skip.p, skip, (size_t)n_cand, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        ctx->d_skip.release();
    }
    RP_CUDA(cudaStreamSynchronize(ctx->stream));       // pageable host buffers are the caller's
    ctx->segs_dirty = true;
    ctx->have_inputs = true;
    if (int rc = launch_plan(ctx)) return rc;
    return rp_grid_result(ctx, out);
}

int rp_fetch_states(rp_ctx* ctx, int idx, double* out) {
    if (int rc = bind(ctx)) return rc;
    if (!out) return fail(RP_ERR_ARG, "null output");
    if (!ctx->have_plan) return fail(RP_ERR_STATE, "no plan launched");
    if (idx < 0 || idx >= ctx->n_cand) return fail(RP_ERR_ARG, "candidate index out of range");
    const size_t bytes = (size_t)14 * (ctx->in.N + 1) * sizeof(double);
    const rp::PlanResultDev* hres = static_cast<rp::PlanResultDev*>(ctx->h_result.p);
    const int k0 = ctx->cycle_valid ? ctx->cyc_lv[ctx->cyc_sel].k0 : 0;         // cycle launch: the selected level's slice
    if (ctx->states_all_valid) {
        RP_CUDA(cudaMemcpyAsync(out, ctx->d_states_all.as<double>() + (size_t)(k0 + idx) * 14 * (ctx->in.N + 1), bytes,
                                cudaMemcpyDeviceToHost, ctx->stream));
    } else if (ctx->cycle_valid) {
        if (ctx->cyc_sel == ctx->cyc_chosen && ctx->h_cycle->res[ctx->cyc_chosen].r.winner == idx) {
            std::memcpy(out, reinterpret_cast<const char*>(ctx->h_cycle) + rp::kCycleStatesOffset, bytes);   // written by the kernel
            return RP_OK;
        }
        // any other candidate: re-evaluated on demand (needs the level's coefficient arrays)
        if (int rc = cycle_ensure_coeffs(ctx)) return rc;
        RP_CUDA(cudaMemcpyAsync(ctx->d_index.p, &idx, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        DevBuf tmp;
        if (int rc = tmp.ensure(bytes)) return rc;
        int rc = launch_states_for_index(ctx, ctx->d_index.as<int>(), 1, tmp.as<double>());
        cudaError_t e = rc ? cudaSuccess : cudaMemcpyAsync(out, tmp.p, bytes, cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
        tmp.release();
        if (rc) return rc;
        if (e != cudaSuccess || e2 != cudaSuccess) return fail(RP_ERR_CUDA, cudaGetErrorString(e != cudaSuccess ? e : e2));
        return RP_OK;
    } else {
        if (!ctx->h_states_valid) {
            RP_CUDA(cudaMemcpyAsync(ctx->h_result.p, ctx->d_result.p, rp_ctx::kResBytes + ctx->res_states_bytes, cudaMemcpyDeviceToHost,
                                    ctx->stream));
            RP_CUDA(cudaStreamSynchronize(ctx->stream));
            ctx->h_states_valid = true;
        }
        if (hres->r.winner == idx) {
            std::memcpy(out, static_cast<const char*>(ctx->h_result.p) + rp_ctx::kResBytes, bytes);
            return RP_OK;
        }
        // lazily re-evaluate one candidate (TrajectorySample views of non-winners)
        RP_CUDA(cudaMemcpyAsync(ctx->d_index.p, &idx, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        if (int rc = ctx->d_states_all.ensure(bytes)) return rc;
        if (int rc = launch_states_for_index(ctx, ctx->d_index.as<int>(), 1, ctx->d_states_all.as<double>())) return rc;
        RP_CUDA(cudaMemcpyAsync(out, ctx->d_states_all.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    RP_CUDA(cudaStreamSynchronize(ctx->stream));
    return RP_OK;
}

int rp_fetch_candidates(rp_ctx* ctx, double* cost, int32_t* status, int32_t* reason, int32_t* step) {
    if (int rc = bind(ctx)) return rc;
    if (!ctx->have_plan) return fail(RP_ERR_STATE, "no plan launched");
    const int n = ctx->n_cand;
    if (n == 0) return RP_OK;
    const int k0 = ctx->cycle_valid ? ctx->cyc_lv[ctx->cyc_sel].k0 : 0;         // cycle launch: the selected level's slice
    if (cost) RP_CUDA(cudaMemcpyAsync(cost, ctx->d_cost.as<double>() + k0, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<int> info;
    if (status || reason || step) {
        info.resize(n);
        RP_CUDA(cudaMemcpyAsync(info.data(), ctx->d_info.as<int>() + k0, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    }
    RP_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int q = 0; q < (int)info.size(); ++q) {
        if (status) status[q] = info[q] & 0xFF;
        if (reason) reason[q] = (info[q] >> 8) & 0xFF;
        if (step) step[q] = ((info[q] >> 16) & 0xFFFF) - 1;
    }
    return RP_OK;
}

int rp_fetch_coeffs(rp_ctx* ctx, double* coeffs_lon, double* coeffs_lat, double* delta_tau_lat) {
    if (int rc = bind(ctx)) return rc;
    if (!ctx->have_plan) return fail(RP_ERR_STATE, "no plan launched");
    const int n = ctx->n_cand;
    if (n == 0) return RP_OK;
    if (ctx->cycle_valid)
        if (int rc = cycle_ensure_coeffs(ctx)) return rc;
    if (ctx->mode == 1) {
        if (coeffs_lon) RP_CUDA(cudaMemcpyAsync(coeffs_lon, ctx->d_lon_coef.p, (size_t)n * 48, cudaMemcpyDeviceToHost, ctx->stream));
        if (coeffs_lat) RP_CUDA(cudaMemcpyAsync(coeffs_lat, ctx->d_lat_coef.p, (size_t)n * 48, cudaMemcpyDeviceToHost, ctx->stream));
        RP_CUDA(cudaStreamSynchronize(ctx->stream));
        return RP_OK;
    }
    const int n_lon_sys = ctx->n_t * ctx->n_lon;
    const bool lv = ctx->in.low_vel_mode != 0;
    const int n_lat_sys = lv ? n : ctx->n_t * ctx->n_d;
    std::vector<double> hl((size_t)n_lon_sys * 6), ht((size_t)n_lat_sys * 6), htau((size_t)n_lat_sys);
    RP_CUDA(cudaMemcpyAsync(hl.data(), ctx->d_lon_coef.p, hl.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    RP_CUDA(cudaMemcpyAsync(ht.data(), ctx->d_lat_coef.p, ht.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    RP_CUDA(cudaMemcpyAsync(htau.data(), ctx->d_lat_tau.p, htau.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    RP_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < n; ++k) {
        const int per_t = ctx->n_lon * ctx->n_d;
        const int it = k / per_t, rem = k - it * per_t, il = rem / ctx->n_d, id = rem - il * ctx->n_d;
        const size_t ql = (size_t)(it * ctx->n_lon + il), qt = lv ? (size_t)k : (size_t)(it * ctx->n_d + id);
        if (coeffs_lon) std::memcpy(coeffs_lon + (size_t)k * 6, &hl[ql * 6], 48);
        if (coeffs_lat) std::memcpy(coeffs_lat + (size_t)k * 6, &ht[qt * 6], 48);
        if (delta_tau_lat) delta_tau_lat[k] = htau[qt];
    }
    return RP_OK;
}

int rp_solve_coeffs(rp_ctx* ctx, int n, const int32_t* kind, const double* x0, const double* xd, const double* tau,
                    double* coeffs) {
    if (int rc = bind(ctx)) return rc;
    if (n < 0) return fail(RP_ERR_ARG, "negative count");
    if (n == 0) return RP_OK;
    if (!kind || !x0 || !xd || !tau || !coeffs) return fail(RP_ERR_ARG, "null array");
    DevBuf dk, dx0, dxd, dtau, dc, dok;
    int rc = RP_OK;
    auto cleanup = [&]() { dk.release(); dx0.release(); dxd.release(); dtau.release(); dc.release(); dok.release(); };
    if ((rc = dk.ensure((size_t)n * 4)) || (rc = dx0.ensure((size_t)n * 24)) || (rc = dxd.ensure((size_t)n * 24)) ||
        (rc = dtau.ensure((size_t)n * 8)) || (rc = dc.ensure((size_t)n * 48)) || (rc = dok.ensure((size_t)n * 4))) {
        cleanup();
        return rc;
    }
    cudaMemcpyAsync(dk.p, kind, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(dx0.p, x0, (size_t)n * 24, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(dxd.p, xd, (size_t)n * 24, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(dtau.p, tau, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
    rp::solve_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, dk.as<int>(), dx0.as<double>(), dxd.as<double>(),
                                                               dtau.as<double>(), dc.as<double>(), dok.as<int>());
    std::vector<int> ok(n);
    cudaMemcpyAsync(coeffs, dc.p, (size_t)n * 48, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(ok.data(), dok.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cleanup();
    if (e != cudaSuccess) return fail(RP_ERR_CUDA, cudaGetErrorString(e));
    for (int q = 0; q < n; ++q)
        if (!ok[q]) return fail(RP_ERR_ARG, "singular system in rp_solve_coeffs (numpy raises LinAlgError)");
    return RP_OK;
}

int rp_collide_poses(rp_ctx* ctx, int n, const double* pose, const int32_t* time_idx, double half_length,
                     double half_width, uint8_t* hit) {
    if (int rc = bind(ctx)) return rc;
    if (n < 0) return fail(RP_ERR_ARG, "negative count");
    if (n == 0) return RP_OK;
    if (!pose || !time_idx || !hit) return fail(RP_ERR_ARG, "null array");
    if (!ctx->have_vehicle) return fail(RP_ERR_STATE, "vehicle parameters not set");
    const double r_q = std::sqrt(half_length * half_length + half_width * half_width);
    const double r_v = std::sqrt(0.25 * ctx->veh.length * ctx->veh.length + 0.25 * ctx->veh.width * ctx->veh.width);
    if (r_q > r_v * (1.0 + 1e-12))
        return fail(RP_ERR_ARG, "query box larger than the vehicle the broad-phase grid was built for");
    if (ctx->obstacles_dirty)
        if (int rc = build_obstacle_tables(ctx)) return rc;
    DevBuf dp, dt, dh;
    int rc = RP_OK;
    auto cleanup = [&]() { dp.release(); dt.release(); dh.release(); };
    if ((rc = dp.ensure((size_t)n * 24)) || (rc = dt.ensure((size_t)n * 4)) || (rc = dh.ensure((size_t)n))) {
        cleanup();
        return rc;
    }
    cudaMemcpyAsync(dp.p, pose, (size_t)n * 24, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(dt.p, time_idx, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream);
    // the clearance bitmask is only valid for boxes inside the vehicle's own box
    const int vehicle_box = (half_length <= 0.5 * ctx->veh.length && half_width <= 0.5 * ctx->veh.width) ? 1 : 0;
    rp::collide_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, dp.as<double>(), dt.as<int>(), half_length,
                                                                 half_width, r_q, ctx->obs, vehicle_box, dh.as<uint8_t>());
    cudaMemcpyAsync(hit, dh.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cleanup();
    if (e != cudaSuccess) return fail(RP_ERR_CUDA, cudaGetErrorString(e));
    return RP_OK;
}

namespace {
__global__ void divide_kernel(int n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ q_shared,
                              double* __restrict__ q_plain, int* __restrict__ rejected) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    // exactly what poly_step does: fast quotient with a sticky reject word, plain division on reject
    unsigned reject = 0u;
    double q = rp::div_fast<true>(a[g], b[g], rp::rcp_window(b[g]), reject);
    const bool rej = (reject & 0x80000000u) != 0u;
    if (rej) q = a[g] / b[g];
    q_shared[g] = q;
    q_plain[g] = a[g] / b[g];
    if (rejected) rejected[g] = rej ? 1 : 0;
}
}  // namespace

int rp_selftest_divide(rp_ctx* ctx, int n, const double* a, const double* b, double* q_shared, double* q_plain, int32_t* rejected) {
    if (int rc = bind(ctx)) return rc;
    if (n < 0) return fail(RP_ERR_ARG, "negative count");
    if (n == 0) return RP_OK;
    if (!a || !b || !q_shared || !q_plain) return fail(RP_ERR_ARG, "null array");
    DevBuf da, db, d1, d2, d3;
    int rc = RP_OK;
    auto cleanup = [&]() { da.release(); db.release(); d1.release(); d2.release(); d3.release(); };
    const size_t bytes = (size_t)n * sizeof(double);
    if ((rc = da.ensure(bytes)) || (rc = db.ensure(bytes)) || (rc = d1.ensure(bytes)) || (rc = d2.ensure(bytes)) ||
        (rc = d3.ensure((size_t)n * sizeof(int)))) {
        cleanup();
        return rc;
    }
    cudaMemcpyAsync(da.p, a, bytes, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(db.p, b, bytes, cudaMemcpyHostToDevice, ctx->stream);
    divide_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, da.as<double>(), db.as<double>(), d1.as<double>(), d2.as<double>(),
                                                             d3.as<int>());
    cudaMemcpyAsync(q_shared, d1.p, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(q_plain, d2.p, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (rejected) cudaMemcpyAsync(rejected, d3.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cleanup();
    if (e != cudaSuccess) return fail(RP_ERR_CUDA, cudaGetErrorString(e));
    return RP_OK;
}

int rp_stage_ms(rp_ctx* ctx, int back, float* ms4) {
    if (int rc = bind(ctx)) return rc;
    if (!ms4) return fail(RP_ERR_ARG, "null output");
    if (back < 0 || back >= rp_ctx::kEvRing || back >= ctx->n_launches) return fail(RP_ERR_STATE, "no such launch in the event ring");
    const int slot = (int)((ctx->n_launches - 1 - back) % rp_ctx::kEvRing);
    if (!ctx->ev_recorded[slot]) return fail(RP_ERR_STATE, "that launch ran without stage timing (rp_ctx_set_stage_timing)");
    cudaEvent_t* ev = ctx->ev_ring[slot];
    RP_CUDA(cudaEventSynchronize(ev[4]));
    for (int q = 0; q < 4; ++q) RP_CUDA(cudaEventElapsedTime(&ms4[q], ev[q], ev[q + 1]));
    return RP_OK;
}

int rp_last_stage_ms(rp_ctx* ctx, float* ms4) { return rp_stage_ms(ctx, 0, ms4); }

int rp_ctx_set_stage_timing(rp_ctx* ctx, int on) {
    if (!ctx) return fail(RP_ERR_ARG, "null context");
    ctx->stage_timing = on != 0;
    return RP_OK;
}

// FP64 pipe peak by a DFMA micro-benchmark (the roofline denominator SURVEY 8d asks to measure)
namespace {
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
}  // namespace

int rp_measure_fp64_peak(rp_ctx* ctx, double* tflops) {
    if (int rc = bind(ctx)) return rc;
    if (!tflops) return fail(RP_ERR_ARG, "null output");
    DevBuf sink;
    const int blocks = ctx->num_sms * 8, threads = 256, iters = 20000;
    if (int rc = sink.ensure((size_t)blocks * threads * sizeof(double))) return rc;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, ctx->stream);
        dfma_kernel<<<blocks, threads, 0, ctx->stream>>>(sink.as<double>(), iters, 1.0 + rep);
        cudaEventRecord(e1, ctx->stream);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { sink.release(); return fail(RP_ERR_CUDA, cudaGetErrorString(e)); }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    sink.release();
    *tflops = best;
    return RP_OK;
}

int rp_export_record_dev(rp_ctx* ctx, double* dev_dst4) {
    if (int rc = bind(ctx)) return rc;
    if (!dev_dst4) return fail(RP_ERR_ARG, "null device pointer");
    if (!ctx->have_plan) return fail(RP_ERR_STATE, "no plan launched");
    rp::export_record_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_result.as<rp::PlanResultDev>(), dev_dst4);
    RP_CUDA(cudaGetLastError());
    return RP_OK;
}

int rp_merge_records_dev(rp_ctx* ctx, const double* dev_gathered, int world, double* dev_winner2, double* dev_totals2) {
    if (int rc = bind(ctx)) return rc;
    if (!dev_gathered || !dev_winner2 || !dev_totals2 || world < 1) return fail(RP_ERR_ARG, "invalid merge arguments");
    rp::merge_records_kernel<<<1, 32, 0, ctx->stream>>>(dev_gathered, world, dev_winner2, dev_totals2);
    RP_CUDA(cudaGetLastError());
    return RP_OK;
}

int rp_count_colliders_before_dev(rp_ctx* ctx, const double* dev_winner2, double* dev_out1) {
    if (int rc = bind(ctx)) return rc;
    if (!dev_winner2 || !dev_out1) return fail(RP_ERR_ARG, "null device pointer");
    if (!ctx->have_plan) return fail(RP_ERR_STATE, "no plan launched");
    int first = 0, count = 0;
    shard_extent(ctx, first, count);
    RP_CUDA(cudaMemsetAsync(dev_out1, 0, sizeof(double), ctx->stream));
    const int blocks = std::max(1, std::min(ctx->num_sms, (count + 255) / 256));
    rp::count_before_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_cost.as<double>(), ctx->d_info.as<int>(), first,
                                                              count, dev_winner2, dev_out1, stripe_of(ctx));
    RP_CUDA(cudaGetLastError());
    return RP_OK;
}

int rp_peer_create(rp_ctx* ctx, unsigned char* handle64) {
    if (int rc = bind(ctx)) return rc;
    if (!handle64) return fail(RP_ERR_ARG, "null handle buffer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    if (!ctx->peer_mine) {
        void* p = nullptr;
        RP_CUDA(cudaMalloc(&p, sizeof(rp::PeerMailbox)));
        ctx->peer_mine = static_cast<rp::PeerMailbox*>(p);
    }
    RP_CUDA(cudaMemset(ctx->peer_mine, 0, sizeof(rp::PeerMailbox)));
    RP_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    RP_CUDA(cudaIpcGetMemHandle(&h, ctx->peer_mine));
    std::memcpy(handle64, &h, 64);
    ctx->peer_ready = false;
    ctx->peer_epoch = 0;
    return RP_OK;
}

int rp_peer_open(rp_ctx* ctx, int rank, int world, const unsigned char* handles) {
    if (int rc = bind(ctx)) return rc;
    if (!ctx->peer_mine) return fail(RP_ERR_STATE, "rp_peer_create must precede rp_peer_open");
    if (!handles || world < 1 || world > rp::kMaxPeers || rank < 0 || rank >= world) return fail(RP_ERR_ARG, "invalid peer arguments");
    for (int r = 0; r < rp::kMaxPeers; ++r)
        if (ctx->peer_opened[r]) { cudaIpcCloseMemHandle(ctx->peer_opened[r]); ctx->peer_opened[r] = nullptr; }
    ctx->peer_table = rp::PeerTable{};
    ctx->peer_table.rank = rank;
    ctx->peer_table.world = world;
    for (int r = 0; r < world; ++r) {
        if (r == rank) { ctx->peer_table.box[r] = ctx->peer_mine; continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)r * 64, 64);
        void* p = nullptr;
        RP_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_opened[r] = p;
        ctx->peer_table.box[r] = static_cast<rp::PeerMailbox*>(p);
    }
    ctx->peer_ready = true;
    ctx->peer_table_dirty = true;
    return RP_OK;
}

int rp_peer_close(rp_ctx* ctx) {
    if (!ctx) return RP_OK;
    cudaSetDevice(ctx->device);
    if (ctx->peer_mine) cudaStreamSynchronize(ctx->stream);        // a cycle in flight still stores into the mailboxes
    for (int r = 0; r < rp::kMaxPeers; ++r)
        if (ctx->peer_opened[r]) { cudaIpcCloseMemHandle(ctx->peer_opened[r]); ctx->peer_opened[r] = nullptr; }
    if (ctx->peer_mine) { cudaFree(ctx->peer_mine); ctx->peer_mine = nullptr; }
    ctx->peer_ready = false;
    return RP_OK;
}

int rp_last_main_kernel(rp_ctx* ctx) {
    if (!ctx) return 0;
    if (ctx->cycle_valid) return RP_KERNEL_STEP_PARALLEL;
    return ctx->main_is_cand ? RP_KERNEL_CANDIDATE_MAJOR : RP_KERNEL_STEP_PARALLEL;
}

int rp_launches_per_plan(rp_ctx* ctx) {
    if (!ctx) return 0;
    // coeff, fused, argmin partial / merge / count, winner states (+ the dynamic-obstacle rows of the candidate-major kernel)
    if (ctx->cycle_valid) return 1;                               // rp_plan_levels: the whole cycle is one kernel
    if (ctx->small_path_last) return ctx->mode == 0 ? 3 : 2;      // coeff, fused (states of every kept candidate), select
    // the candidate-major kernel defers the checks of the lazy collision pass: check, gather, check
    const int deferred = (ctx->main_is_cand && ctx->in.check_collision == 2 && ctx->in.cost_kind != RP_COST_NONE) ? 3 : 0;
    if (ctx->peer_mode_last)        // prep, main kernel, argmin partial / merge + record exchange, peer count, winner states
        return deferred + (ctx->mode == 0 ? 6 : 5 + ((ctx->main_is_cand && ctx->obs.n_dyn > 0 && ctx->in.check_collision) ? 1 : 0));
    // prep (coefficients + dynamic-obstacle rows), main kernel, argmin partial + merge (last block) / count, winner states; the list
    // form has no coefficient solve but, for the candidate-major kernel, its own dynamic-obstacle rows launch
    return deferred + (ctx->mode == 0 ? 5 : 4 + ((ctx->main_is_cand && ctx->obs.n_dyn > 0 && ctx->in.check_collision) ? 1 : 0));
}

}  // extern "C"

// =====================================================================================================
// Batch of independent scenarios (BASELINE configs[4]): every scenario keeps its own context (device-resident
// reference / obstacle tables); one replanning cycle of ALL of them is one host->device copy, four launches
// (coefficients, dynamic-obstacle rows, candidate-major evaluation over a global (scenario, chunk) queue,
// one arg-min block per scenario) and one device->host copy.
// =====================================================================================================
struct rp_batch {
    int device = 0;
    cudaStream_t stream = nullptr;
    int num_sms = 148;
    std::vector<rp_ctx*> ctxs;
    struct Slot {
        rp_plan_inputs in{};
        std::vector<double> t, lon, d;
        std::vector<int> traj_len;
        bool set = false;
    };
    std::vector<Slot> slots;
    PinBuf h_stage, h_results;
    DevBuf d_stage, d_lon_coef, d_lat_coef, d_cost, d_info, d_dyn_rows, d_results, d_work;
    DevBuf d_pose, d_defer_list, d_defer_mask, d_best;         // deferred collision check (all scenarios in lazy mode)
    cudaEvent_t ev_stage = nullptr, ev_results = nullptr, ev0 = nullptr, ev1 = nullptr;
    bool stage_pending = false, launched = false;
    long long total_cand = 0;
    int smem_granted = 0;
    bool deferred_last = false;
};

extern "C" {

int rp_batch_create(rp_ctx* const* ctxs, int n, void* stream, rp_batch** out) {
    if (!out) return fail(RP_ERR_ARG, "null out pointer");
    *out = nullptr;
    if (n < 1 || !ctxs) return fail(RP_ERR_ARG, "a batch needs at least one context");
    for (int k = 0; k < n; ++k) {
        if (!ctxs[k]) return fail(RP_ERR_ARG, "null context in batch");
        if (ctxs[k]->device != ctxs[0]->device) return fail(RP_ERR_ARG, "all contexts of a batch must live on one device");
    }
    RP_CUDA(cudaSetDevice(ctxs[0]->device));
    rp_batch* b = new rp_batch();
    b->device = ctxs[0]->device;
    b->stream = stream ? static_cast<cudaStream_t>(stream) : ctxs[0]->stream;
    b->num_sms = ctxs[0]->num_sms;
    b->ctxs.assign(ctxs, ctxs + n);
    b->slots.resize(n);
    cudaEventCreateWithFlags(&b->ev_stage, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&b->ev_results, cudaEventDisableTiming);
    cudaEventCreate(&b->ev0);
    cudaEventCreate(&b->ev1);
    *out = b;
    return RP_OK;
}

int rp_batch_destroy(rp_batch* b) {
    if (!b) return RP_OK;
    cudaSetDevice(b->device);
    cudaStreamSynchronize(b->stream);
    for (rp_ctx* c : b->ctxs)
        if (c->ext_busy == b->ev1) c->ext_busy = nullptr;
    for (DevBuf* q : {&b->d_stage, &b->d_lon_coef, &b->d_lat_coef, &b->d_cost, &b->d_info, &b->d_dyn_rows, &b->d_results, &b->d_work,
                      &b->d_pose, &b->d_defer_list, &b->d_defer_mask, &b->d_best})
        q->release();
    b->h_stage.release();
    b->h_results.release();
    for (cudaEvent_t e : {b->ev_stage, b->ev_results, b->ev0, b->ev1})
        if (e) cudaEventDestroy(e);
    delete b;
    return RP_OK;
}

int rp_batch_size(rp_batch* b) { return b ? (int)b->ctxs.size() : 0; }

int rp_batch_set_inputs(rp_batch* b, int k, const rp_plan_inputs* in, int n_t, const double* t, const int32_t* traj_len,
                        int n_lon, const double* lon, int n_d, const double* d) {
    if (!b) return fail(RP_ERR_ARG, "null batch");
    if (k < 0 || k >= (int)b->ctxs.size()) return fail(RP_ERR_ARG, "scenario index out of range");
    if (int rc = check_inputs(in)) return rc;
    if (n_t < 0 || n_lon < 0 || n_d < 0) return fail(RP_ERR_ARG, "negative sample count");
    if ((n_t && (!t || !traj_len)) || (n_lon && !lon) || (n_d && !d)) return fail(RP_ERR_ARG, "null sample array");
    if ((long long)n_t * n_lon * n_d > 0x7fffffffLL / 8) return fail(RP_ERR_ARG, "bundle too large");
    if (in->want_all_states || in->draw_all) return fail(RP_ERR_ARG, "a batch runs in select-only mode (no draw_all / want_all_states)");
    if (in->continuous_collision_check) return fail(RP_ERR_ARG, "a batch does not run the continuous collision check");
    if (in->N + 1 > 128) return fail(RP_ERR_ARG, "a batch supports N + 1 <= 128");
    for (int q = 0; q < n_t; ++q)
        if (traj_len[q] < 1 || traj_len[q] > in->N + 1) return fail(RP_ERR_ARG, "traj_len out of [1, N+1]");
    rp_batch::Slot& s = b->slots[k];
    s.in = *in;
    s.t.assign(t, t + n_t);
    s.lon.assign(lon, lon + n_lon);
    s.d.assign(d, d + n_d);
    s.traj_len.assign(traj_len, traj_len + n_t);
    s.set = true;
    return RP_OK;
}

int rp_batch_set_inputs_all(rp_batch* b, const rp_plan_inputs* in, const int32_t* n_t, const int32_t* n_lon, const int32_t* n_d,
                            const double* t_cat, const int32_t* traj_len_cat, const double* lon_cat, const double* d_cat) {
    if (!b) return fail(RP_ERR_ARG, "null batch");
    if (!in || !n_t || !n_lon || !n_d) return fail(RP_ERR_ARG, "null array");
    size_t ot = 0, ol = 0, od = 0;
    for (int k = 0; k < (int)b->ctxs.size(); ++k) {
        if (int rc = rp_batch_set_inputs(b, k, in + k, n_t[k], t_cat ? t_cat + ot : nullptr, traj_len_cat ? traj_len_cat + ot : nullptr,
                                         n_lon[k], lon_cat ? lon_cat + ol : nullptr, n_d[k], d_cat ? d_cat + od : nullptr))
            return rc;
        ot += (size_t)std::max(n_t[k], 0);
        ol += (size_t)std::max(n_lon[k], 0);
        od += (size_t)std::max(n_d[k], 0);
    }
    return RP_OK;
}

int rp_batch_launch(rp_batch* b) {
    if (!b) return fail(RP_ERR_ARG, "null batch");
    RP_CUDA(cudaSetDevice(b->device));
    const int n = (int)b->ctxs.size();
    // ---- layout of the staging buffer: params | chunk_prefix | per scenario: t, lon, d, traj_len, segments -------
    auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    size_t off = align16((size_t)n * sizeof(PlanParams));
    const size_t off_prefix = off;
    off = align16(off + (size_t)(n + 1) * sizeof(int));
    std::vector<size_t> off_samples(n), off_segs(n);
    std::vector<std::vector<rp::Segment>> segs(n);
    std::vector<int> prefix(n + 1, 0);
    std::vector<size_t> off_lon(n), off_lat(n), off_cand(n), off_rows(n), off_tiles(n), off_pose(n);
    size_t n_lon_tot = 0, n_lat_tot = 0, n_cand_tot = 0, n_rows_tot = 0, n_tiles_tot = 0, n_pose_tot = 0;
    int max_sys = 0, max_rows = 0, acc_rows = 0, max_cand = 0;
    // the lazy collision pass is deferred (rp_cand.cuh, deferred_collision_batch_kernel) when EVERY scenario asks for it
    bool defer = n < (1 << (32 - rp::kDeferTileBits - 1));
    for (int k = 0; k < n; ++k) {
        rp_ctx* c = b->ctxs[k];
        const rp_batch::Slot& s = b->slots[k];
        if (!s.set) return fail(RP_ERR_STATE, "rp_batch_set_inputs missing for a scenario");
        if (int rc = check_ready(c)) return rc;
        const int n_t = (int)s.t.size(), n_lon = (int)s.lon.size(), n_d = (int)s.d.size();
        const int Np1 = s.in.N + 1;
        off_samples[k] = off;
        off = align16(off + (size_t)(n_t + n_lon + n_d) * sizeof(double) + (size_t)n_t * sizeof(int));
        const int per_t = n_lon * n_d;
        for (int it = 0; it < n_t && per_t > 0; ++it)
            segs[k].push_back(rp::Segment{it * per_t, (it + 1) * per_t, std::max(1, std::min(s.traj_len[it], Np1)), 32, 0, 0});
        if (segs[k].empty()) segs[k].push_back(rp::Segment{0, 0, Np1, 32, 0, 0});
        std::stable_sort(segs[k].begin(), segs[k].end(), [](const rp::Segment& x, const rp::Segment& y) { return x.tl > y.tl; });
        int groups = 0;
        for (auto& sg : segs[k]) {
            sg.g_begin = groups;
            groups += (sg.k_end - sg.k_begin + 31) / 32;
        }
        prefix[k + 1] = prefix[k] + groups;
        off_segs[k] = off;
        off = align16(off + segs[k].size() * sizeof(rp::Segment));
        const int n_cand = n_t * per_t;
        const int n_lon_sys = n_t * n_lon, n_lat_sys = s.in.low_vel_mode ? n_cand : n_t * n_d;
        off_lon[k] = n_lon_tot; off_lat[k] = n_lat_tot; off_cand[k] = n_cand_tot; off_rows[k] = n_rows_tot;
        off_tiles[k] = n_tiles_tot; off_pose[k] = n_pose_tot;
        n_lon_tot += (size_t)n_lon_sys; n_lat_tot += (size_t)n_lat_sys; n_cand_tot += (size_t)n_cand;
        const size_t tiles = (size_t)(n_cand + 31) / 32;
        n_tiles_tot += tiles;
        n_pose_tot += tiles * Np1 * 32 * 4;                   // doubles
        max_cand = std::max(max_cand, n_cand);
        defer = defer && s.in.check_collision == 2 && s.in.cost_kind != RP_COST_NONE && (size_t)n_cand < ((size_t)1 << rp::kDeferTileBits);
        const int rows = (c->obs.n_dyn > 0 && s.in.check_collision) ? Np1 * c->obs.n_dyn : 0;
        n_rows_tot += (size_t)rows;
        max_sys = std::max(max_sys, n_lon_sys + n_lat_sys);
        max_rows = std::max(max_rows, rows);
        acc_rows = std::max(acc_rows, cand_acc_rows(s.in));
    }
    // (the stored ego boxes: 32 bytes per candidate-timestep; beyond the cap the march checks collisions itself)
    if (n_pose_tot * sizeof(double) > kDeferMaxBytes) defer = false;
    const size_t stage_bytes = off;
    if (b->stage_pending) RP_CUDA(cudaEventSynchronize(b->ev_stage));          // last copy out of the pinned buffer (before it may be re-allocated)
    if (int rc = b->h_stage.ensure(stage_bytes)) return rc;
    if (int rc = b->d_stage.ensure(stage_bytes)) return rc;
    if (int rc = b->d_lon_coef.ensure(std::max<size_t>(n_lon_tot, 1) * 6 * sizeof(double))) return rc;
    if (int rc = b->d_lat_coef.ensure(std::max<size_t>(n_lat_tot, 1) * 6 * sizeof(double))) return rc;
    if (int rc = b->d_cost.ensure(std::max<size_t>(n_cand_tot, 1) * sizeof(double))) return rc;
    if (int rc = b->d_info.ensure(std::max<size_t>(n_cand_tot, 1) * sizeof(int))) return rc;
    if (int rc = b->d_dyn_rows.ensure(std::max<size_t>(n_rows_tot, 1) * sizeof(float4))) return rc;
    if (int rc = b->d_results.ensure((size_t)n * sizeof(rp::PlanResultDev))) return rc;
    if (int rc = b->h_results.ensure((size_t)n * sizeof(rp::PlanResultDev))) return rc;
    if (int rc = b->d_work.ensure(sizeof(int) * rp::kWorkWords)) return rc;
    if (defer) {
        if (int rc = b->d_pose.ensure(std::max<size_t>(n_pose_tot, 1) * sizeof(double))) return rc;
        if (int rc = b->d_defer_list.ensure((std::max<size_t>(n_tiles_tot, 1) + std::max<size_t>(n_cand_tot, 1)) * sizeof(int))) return rc;   // tiles, then (scenario, candidate) pairs
        if (int rc = b->d_best.ensure((size_t)n * sizeof(unsigned long long))) return rc;
        if (b->d_defer_mask.cap < n_tiles_tot * sizeof(unsigned)) {            // (the checker leaves the masks it used clear)
            if (int rc = b->d_defer_mask.ensure(std::max<size_t>(n_tiles_tot, 1) * sizeof(unsigned))) return rc;
            RP_CUDA(cudaMemsetAsync(b->d_defer_mask.p, 0, b->d_defer_mask.cap, b->stream));
        }
    }
    char* hs = static_cast<char*>(b->h_stage.p);
    const char* ds = static_cast<const char*>(b->d_stage.p);
    std::memcpy(hs + off_prefix, prefix.data(), prefix.size() * sizeof(int));
    PlanParams* hp = reinterpret_cast<PlanParams*>(hs);
    Geometry G{};
    G.stage_ref = 0; G.stage_dyn = 0; G.Cmax = 32;
    for (int k = 0; k < n; ++k) {
        rp_ctx* c = b->ctxs[k];
        const rp_batch::Slot& s = b->slots[k];
        const int n_t = (int)s.t.size(), n_lon = (int)s.lon.size(), n_d = (int)s.d.size();
        char* p = hs + off_samples[k];
        if (n_t) std::memcpy(p, s.t.data(), (size_t)n_t * 8);
        if (n_lon) std::memcpy(p + (size_t)n_t * 8, s.lon.data(), (size_t)n_lon * 8);
        if (n_d) std::memcpy(p + (size_t)(n_t + n_lon) * 8, s.d.data(), (size_t)n_d * 8);
        if (n_t) std::memcpy(p + (size_t)(n_t + n_lon + n_d) * 8, s.traj_len.data(), (size_t)n_t * 4);
        std::memcpy(hs + off_segs[k], segs[k].data(), segs[k].size() * sizeof(rp::Segment));
        // the scenario's context describes its tables; the cycle's inputs come from the slot
        c->in = s.in;
        c->mode = 0;
        c->n_t = n_t; c->n_lon = n_lon; c->n_d = n_d;
        c->n_cand = n_t * n_lon * n_d;
        c->have_plan = false;
        G.n_segs = (int)segs[k].size();
        G.n_groups = prefix[k + 1] - prefix[k];
        PlanParams P{};
        fill_common(c, P, G, reinterpret_cast<const rp::Segment*>(ds + off_segs[k]));
        const char* dsamp = ds + off_samples[k];
        P.t_samples = reinterpret_cast<const double*>(dsamp);
        P.lon_samples = reinterpret_cast<const double*>(dsamp + (size_t)n_t * 8);
        P.d_samples = reinterpret_cast<const double*>(dsamp + (size_t)(n_t + n_lon) * 8);
        P.traj_len = reinterpret_cast<const int*>(dsamp + (size_t)(n_t + n_lon + n_d) * 8);
        P.skip = nullptr;
        P.lon_coef = b->d_lon_coef.as<double>() + off_lon[k] * 6;
        P.lat_coef = b->d_lat_coef.as<double>() + off_lat[k] * 6;
        P.index = nullptr;
        P.cost = b->d_cost.as<double>() + off_cand[k];
        P.info = b->d_info.as<int>() + off_cand[k];
        P.states = nullptr;
        P.best_bits = nullptr;
        P.work_counter = nullptr;
        P.n_acc_rows = cand_acc_rows(s.in);
        P.dyn_rows = (c->obs.n_dyn > 0 && s.in.check_collision) ? b->d_dyn_rows.as<float4>() + off_rows[k] : nullptr;
        P.pose = nullptr;
        if (defer) {
            P.pose = b->d_pose.as<double>() + off_pose[k];
            P.defer_list = b->d_defer_list.as<int>();
            P.defer_list2 = P.defer_list + n_tiles_tot;
            P.defer_count = b->d_work.as<int>() + 1;
            P.defer_mask = b->d_defer_mask.as<unsigned>() + off_tiles[k];
            P.defer_tag = k << rp::kDeferTileBits;
            P.best_bits = b->d_best.as<unsigned long long>() + k;
        }
        hp[k] = P;
    }
    RP_CUDA(cudaMemcpyAsync(b->d_stage.p, b->h_stage.p, stage_bytes, cudaMemcpyHostToDevice, b->stream));
    RP_CUDA(cudaEventRecord(b->ev_stage, b->stream));
    b->stage_pending = true;
    RP_CUDA(cudaMemsetAsync(b->d_work.p, 0, sizeof(int) * rp::kWorkWords, b->stream));
    if (defer) RP_CUDA(cudaMemsetAsync(b->d_best.p, 0x7f, (size_t)n * sizeof(unsigned long long), b->stream));      // ~1.4e306
    cudaEventRecord(b->ev0, b->stream);
    const PlanParams* dparams = reinterpret_cast<const PlanParams*>(ds);
    if (max_sys > 0) rp::coeff_batch_kernel<<<dim3((max_sys + 127) / 128, n), 128, 0, b->stream>>>(dparams);
    if (max_rows > 0) rp::dyn_rows_batch_kernel<<<dim3((max_rows + 127) / 128, n), 128, 0, b->stream>>>(dparams);
    if (prefix[n] > 0) {
        rp::BatchTable T{};
        T.params = dparams;
        T.chunk_prefix = reinterpret_cast<const int*>(ds + off_prefix);
        T.n_scenarios = n;
        T.n_acc_rows = acc_rows;
        T.work_counter = b->d_work.as<int>();
        const int threads = RP_CAND_THREADS;
        const size_t smem = (size_t)(acc_rows * 8 + 1) * threads * sizeof(double) + (size_t)(threads / 32) * sizeof(rp::LimitRcp) +
                            (size_t)(threads / 32) * rp::kWarpRowDoubles * sizeof(double) + 64;
        if ((int)smem > b->smem_granted) {
            RP_CUDA(cudaFuncSetAttribute(rp::cand_batch_kernel<RP_CAND_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            RP_CUDA(cudaFuncSetAttribute(rp::cand_batch_kernel<RP_CAND_THREADS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            b->smem_granted = (int)smem;
        }
        int occ = 0;
        RP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rp::cand_batch_kernel<RP_CAND_THREADS>, threads, smem));
        if (occ < 1) return fail(RP_ERR_CUDA, "batch kernel does not fit on an SM");
        const int wpb = threads / 32;
        const int grid = std::max(1, std::min((prefix[n] + wpb - 1) / wpb, occ * b->num_sms));
        if (defer) {
            rp::cand_batch_kernel<RP_CAND_THREADS, true><<<grid, threads, smem, b->stream>>>(T);
            const int check_blocks = (int)std::max<size_t>(1, std::min<size_t>(n_tiles_tot, (size_t)8 * b->num_sms));
            int* lists = b->d_defer_list.as<int>();
            int* counts = b->d_work.as<int>() + 1;
            rp::deferred_collision_batch_kernel<<<check_blocks, rp::kDeferThreads, 0, b->stream>>>(dparams, lists, counts);
            rp::deferred_gather_batch_kernel<<<dim3((max_cand + 255) / 256, n), 256, 0, b->stream>>>(dparams);
            rp::deferred_collision_list_batch_kernel<<<check_blocks, rp::kDeferThreads, 0, b->stream>>>(dparams, lists + n_tiles_tot, counts + 1);
        } else {
            rp::cand_batch_kernel<RP_CAND_THREADS><<<grid, threads, smem, b->stream>>>(T);
        }
    }
    b->deferred_last = defer && prefix[n] > 0;
    rp::argmin_batch_kernel<<<n, 1024, 0, b->stream>>>(dparams, b->d_results.as<rp::PlanResultDev>());
    RP_CUDA(cudaGetLastError());
    cudaEventRecord(b->ev1, b->stream);
    for (rp_ctx* c : b->ctxs) c->ext_busy = b->ev1;          // table updates of a member context wait for this cycle
    b->total_cand = (long long)n_cand_tot;
    b->launched = true;
    return RP_OK;
}

int rp_batch_results(rp_batch* b, rp_plan_result* out) {
    if (!b || !out) return fail(RP_ERR_ARG, "null argument");
    if (!b->launched) return fail(RP_ERR_STATE, "no batch launched");
    RP_CUDA(cudaSetDevice(b->device));
    const int n = (int)b->ctxs.size();
    RP_CUDA(cudaMemcpyAsync(b->h_results.p, b->d_results.p, (size_t)n * sizeof(rp::PlanResultDev), cudaMemcpyDeviceToHost, b->stream));
    RP_CUDA(cudaEventRecord(b->ev_results, b->stream));
    RP_CUDA(cudaEventSynchronize(b->ev_results));
    const rp::PlanResultDev* h = static_cast<const rp::PlanResultDev*>(b->h_results.p);
    for (int k = 0; k < n; ++k) out[k] = h[k].r;
    return RP_OK;
}

int rp_batch_fetch_candidates(rp_batch* b, int k, double* cost, int32_t* status, int32_t* reason, int32_t* step) {
    if (!b) return fail(RP_ERR_ARG, "null batch");
    if (!b->launched) return fail(RP_ERR_STATE, "no batch launched");
    if (k < 0 || k >= (int)b->ctxs.size()) return fail(RP_ERR_ARG, "scenario index out of range");
    RP_CUDA(cudaSetDevice(b->device));
    const PlanParams* hp = reinterpret_cast<const PlanParams*>(b->h_stage.p);
    const int n = hp[k].n_cand;
    if (n == 0) return RP_OK;
    if (cost) RP_CUDA(cudaMemcpyAsync(cost, hp[k].cost, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    std::vector<int> info(n);
    RP_CUDA(cudaMemcpyAsync(info.data(), hp[k].info, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    RP_CUDA(cudaStreamSynchronize(b->stream));
    for (int q = 0; q < n; ++q) {
        if (status) status[q] = info[q] & 0xFF;
        if (reason) reason[q] = (info[q] >> 8) & 0xFF;
        if (step) step[q] = ((info[q] >> 16) & 0xFFFF) - 1;
    }
    return RP_OK;
}

int rp_batch_winner_states(rp_batch* b, int step, double* out) {
    if (!b || !out) return fail(RP_ERR_ARG, "null argument");
    if (!b->launched) return fail(RP_ERR_STATE, "no batch launched");
    if (step < 0) return fail(RP_ERR_ARG, "negative time step");
    RP_CUDA(cudaSetDevice(b->device));
    const int n = (int)b->ctxs.size();
    for (int k = 0; k < n; ++k)
        if (step > b->slots[k].in.N) return fail(RP_ERR_ARG, "time step beyond a scenario's horizon");
    DevBuf d_out;
    if (int rc = d_out.ensure((size_t)n * 16 * sizeof(double))) return rc;
    rp::batch_winner_step_kernel<<<(n + 63) / 64, 64, 0, b->stream>>>(reinterpret_cast<const PlanParams*>(b->d_stage.p),
                                                                     b->d_results.as<rp::PlanResultDev>(), n, step, d_out.as<double>());
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out.p, (size_t)n * 16 * sizeof(double), cudaMemcpyDeviceToHost, b->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
    d_out.release();
    if (e != cudaSuccess) return fail(RP_ERR_CUDA, cudaGetErrorString(e));
    return RP_OK;
}

int rp_batch_last_ms(rp_batch* b, float* ms, long long* n_candidates) {
    if (!b || !ms) return fail(RP_ERR_ARG, "null argument");
    if (!b->launched) return fail(RP_ERR_STATE, "no batch launched");
    RP_CUDA(cudaSetDevice(b->device));
    RP_CUDA(cudaEventSynchronize(b->ev1));
    RP_CUDA(cudaEventElapsedTime(ms, b->ev0, b->ev1));
    if (n_candidates) *n_candidates = b->total_cand;
    return RP_OK;
}

}  // extern "C"

// ---- Cartesian -> curvilinear initial states on the device ---------------------------------------
namespace {
int run_initial_states(int device, cudaStream_t stream, int n, const double* x0, const int32_t* low_vel,
                       const std::vector<rp::InitFrame>& frames, double* out_lon, double* out_lat, int32_t* status) {
    RP_CUDA(cudaSetDevice(device));
    DevBuf dx, dl, df, do1, do2, ds;
    int rc = RP_OK;
    auto cleanup = [&]() { dx.release(); dl.release(); df.release(); do1.release(); do2.release(); ds.release(); };
    if ((rc = dx.ensure((size_t)n * 48)) || (rc = dl.ensure((size_t)n * 4)) || (rc = df.ensure(frames.size() * sizeof(rp::InitFrame))) ||
        (rc = do1.ensure((size_t)n * 24)) || (rc = do2.ensure((size_t)n * 24)) || (rc = ds.ensure((size_t)n * 4))) {
        cleanup();
        return rc;
    }
    cudaMemcpyAsync(dx.p, x0, (size_t)n * 48, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(dl.p, low_vel, (size_t)n * 4, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(df.p, frames.data(), frames.size() * sizeof(rp::InitFrame), cudaMemcpyHostToDevice, stream);
    rp::initial_states_kernel<<<n, 128, 0, stream>>>(n, dx.as<double>(), dl.as<int>(), df.as<rp::InitFrame>(),
                                                     frames.size() > 1 ? 1 : 0, do1.as<double>(), do2.as<double>(), ds.as<int>());
    cudaMemcpyAsync(out_lon, do1.p, (size_t)n * 24, cudaMemcpyDeviceToHost, stream);
    cudaMemcpyAsync(out_lat, do2.p, (size_t)n * 24, cudaMemcpyDeviceToHost, stream);
    cudaMemcpyAsync(status, ds.p, (size_t)n * 4, cudaMemcpyDeviceToHost, stream);
    cudaError_t e = cudaStreamSynchronize(stream);
    cleanup();
    if (e != cudaSuccess) return fail(RP_ERR_CUDA, cudaGetErrorString(e));
    return RP_OK;
}

rp::InitFrame frame_of(rp_ctx* ctx) {
    rp::InitFrame F{};
    const double* base = ctx->d_ref.as<double>();
    const int n = ctx->ref_n;
    F.ref.n = n;
    F.ref.same_s = ctx->ref_same_s;
    F.ref.limit = ctx->ref_limit;
    F.ref.pos = base; F.ref.theta = base + n; F.ref.curv = base + 2 * n; F.ref.curv_d = base + 3 * n;
    F.ref.px = base + 4 * n; F.ref.py = base + 5 * n; F.ref.nx = base + 6 * n; F.ref.ny = base + 7 * n;
    F.ref.ps = base + 8 * n;
    F.wheelbase = ctx->veh.wheelbase;
    return F;
}
}  // namespace

extern "C" {

int rp_initial_states(rp_ctx* ctx, int n, const double* x0, const int32_t* low_vel_mode, double* out_lon, double* out_lat,
                      int32_t* status) {
    if (int rc = bind(ctx)) return rc;
    if (n < 0) return fail(RP_ERR_ARG, "negative count");
    if (n == 0) return RP_OK;
    if (!x0 || !low_vel_mode || !out_lon || !out_lat || !status) return fail(RP_ERR_ARG, "null array");
    if (!ctx->have_vehicle) return fail(RP_ERR_STATE, "vehicle parameters not set");
    if (!ctx->have_ref) return fail(RP_ERR_STATE, "reference tables not set");
    std::vector<rp::InitFrame> frames{frame_of(ctx)};
    return run_initial_states(ctx->device, ctx->stream, n, x0, low_vel_mode, frames, out_lon, out_lat, status);
}

int rp_batch_initial_states(rp_batch* b, const double* x0, const int32_t* low_vel_mode, double* out_lon, double* out_lat,
                            int32_t* status) {
    if (!b) return fail(RP_ERR_ARG, "null batch");
    if (!x0 || !low_vel_mode || !out_lon || !out_lat || !status) return fail(RP_ERR_ARG, "null array");
    std::vector<rp::InitFrame> frames;
    for (rp_ctx* c : b->ctxs) {
        if (!c->have_vehicle || !c->have_ref) return fail(RP_ERR_STATE, "vehicle / reference tables not set for a scenario");
        frames.push_back(frame_of(c));
    }
    if (frames.size() == 1) frames.push_back(frames[0]);      // stride 1 addressing
    return run_initial_states(b->device, b->stream, (int)b->ctxs.size(), x0, low_vel_mode, frames, out_lon, out_lat, status);
}

}  // extern "C"
