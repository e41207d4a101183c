"""ctypes binding of the C-ABI in include/rp_b200.h (librp_b200.so, built in-tree by build.py).

There is deliberately NO CPU fallback: if the CUDA library is missing or a call fails, an
exception is raised.  ``Engine`` is a thin owner of one ``rp_ctx`` (one CUDA device + stream).
"""
import ctypes as C
import struct
import functools
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
try:        # optional CPython helper (csrc/rp_pack.c, built by build.py): a faster binding of the per-cycle call
    from commonroad_rp_b200 import _rp_pack
except ImportError:
    _rp_pack = None
LIB_PATH = os.environ.get("RP_B200_LIB", os.path.join(_HERE, "librp_b200.so"))   # override: kernel A/B experiments

N_REASONS = 8
N_STATE_ROWS = 14
STATE_ROWS = ("x", "y", "theta", "v", "a", "kappa", "kappa_dot",
              "s", "d", "theta_cl", "s_dot", "s_ddot", "d_dot", "d_ddot")
REASON_NAMES = ("none", "velocity", "acceleration", "kappa", "kappa_dot", "yaw_rate", "projection", "ref_range")
CONSTRAINT_BITS = {"velocity": 1, "acceleration": 2, "kappa": 4, "kappa_dot": 8, "yaw_rate": 16}
ST_FEASIBLE, ST_KINEMATIC, ST_COLLISION, ST_FILTERED, ST_UNCHECKED = 0, 1, 2, 3, 4
COLLISION_OFF, COLLISION_ALL, COLLISION_LAZY = 0, 1, 2
VELOCITY_KEEPING, STOPPING = 0, 1
COST_DEFAULT, COST_FAILSAFE, COST_NONE = 0, 1, 2
KERNEL_AUTO, KERNEL_STEP_PARALLEL, KERNEL_CANDIDATE_MAJOR = 0, 1, 2


class RpError(RuntimeError):
    pass


class VehicleParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("length", "width", "wb_rear_axle", "wheelbase", "a_max", "v_switch",
                                          "delta_max", "v_delta_max", "kappa_max")]


class PlanInputs(C.Structure):
    _fields_ = [
        ("x0_lon", C.c_double * 3), ("x0_lat", C.c_double * 3), ("x0_orientation", C.c_double),
        ("x0_time_step", C.c_int32), ("low_vel_mode", C.c_int32), ("lon_mode", C.c_int32), ("N", C.c_int32),
        ("dt", C.c_double), ("factor", C.c_int32), ("draw_all", C.c_int32), ("constraint_mask", C.c_uint32),
        ("cost_kind", C.c_int32), ("has_desired_speed", C.c_int32), ("has_desired_s", C.c_int32),
        ("desired_speed", C.c_double), ("desired_s", C.c_double), ("desired_d", C.c_double), ("w_a", C.c_double),
        ("want_all_states", C.c_int32), ("check_collision", C.c_int32),
        ("continuous_collision_check", C.c_int32), ("reserved_", C.c_int32),
    ]


# the same layout for struct.pack_into (one call fills the whole struct; reactive_planner._plan_inputs)
PLAN_INPUTS_STRUCT = struct.Struct("@7d4id2iI3i4d4i")
assert PLAN_INPUTS_STRUCT.size == C.sizeof(PlanInputs)


class PlanResult(C.Structure):
    _fields_ = [
        ("winner", C.c_int32), ("n_candidates", C.c_int32), ("n_feasible", C.c_int32),
        ("n_infeasible_kinematics", C.c_int32), ("n_infeasible_collision", C.c_int32),
        ("n_collision_total", C.c_int32), ("reason_counts", C.c_int32 * N_REASONS), ("winner_cost", C.c_double),
    ]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)

# every symbol include/rp_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "rp_last_error": (C.c_char_p, []),
    "rp_version": (C.c_int, []),
    "rp_ctx_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "rp_ctx_destroy": (C.c_int, [C.c_void_p]),
    "rp_ctx_synchronize": (C.c_int, [C.c_void_p]),
    "rp_ctx_set_vehicle": (C.c_int, [C.c_void_p, C.POINTER(VehicleParams)]),
    "rp_ctx_set_reference": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_double]),
    "rp_ctx_set_reference_polyline": (C.c_int, [C.c_void_p, C.c_int, _dp, C.c_double, C.c_double]),
    "rp_ctx_get_reference": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "rp_ctx_set_obstacles": (C.c_int, [C.c_void_p, C.c_int, _dp, C.c_int, _ip, _ip, _dp, C.c_int, _dp, C.c_double]),
    "rp_plan_grid": (C.c_int, [C.c_void_p, C.POINTER(PlanInputs), C.c_int, _dp, _ip, C.c_int, _dp, C.c_int, _dp,
                               C.POINTER(PlanResult)]),
    "rp_grid_upload": (C.c_int, [C.c_void_p, C.POINTER(PlanInputs), C.c_int, _dp, _ip, C.c_int, _dp, C.c_int, _dp]),
    "rp_grid_launch": (C.c_int, [C.c_void_p]),
    "rp_grid_result": (C.c_int, [C.c_void_p, C.POINTER(PlanResult)]),
    "rp_plan_levels": (C.c_int, [C.c_void_p, C.POINTER(PlanInputs), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(PlanResult), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_int32)]),
    "rp_select_level": (C.c_int, [C.c_void_p, C.c_int]),
    "rp_cycle_host_block": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "rp_cycle_limits": (C.c_int, [C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "rp_plan_list": (C.c_int, [C.c_void_p, C.POINTER(PlanInputs), C.c_int, _dp, _dp, _ip, _bp, C.POINTER(PlanResult)]),
    "rp_set_candidate_range": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "rp_set_candidate_stripe": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "rp_ctx_set_kernel_policy": (C.c_int, [C.c_void_p, C.c_int]),
    "rp_export_record_dev": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rp_count_colliders_before_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "rp_merge_records_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "rp_peer_create": (C.c_int, [C.c_void_p, C.c_char_p]),
    "rp_peer_open": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    "rp_peer_close": (C.c_int, [C.c_void_p]),
    "rp_fetch_states": (C.c_int, [C.c_void_p, C.c_int, _dp]),
    "rp_fetch_candidates": (C.c_int, [C.c_void_p, _dp, _ip, _ip, _ip]),
    "rp_fetch_coeffs": (C.c_int, [C.c_void_p, _dp, _dp, _dp]),
    "rp_solve_coeffs": (C.c_int, [C.c_void_p, C.c_int, _ip, _dp, _dp, _dp, _dp]),
    "rp_collide_poses": (C.c_int, [C.c_void_p, C.c_int, _dp, _ip, C.c_double, C.c_double, _bp]),
    "rp_batch_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "rp_batch_destroy": (C.c_int, [C.c_void_p]),
    "rp_batch_size": (C.c_int, [C.c_void_p]),
    "rp_batch_set_inputs": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(PlanInputs), C.c_int, _dp, _ip, C.c_int, _dp, C.c_int, _dp]),
    "rp_batch_set_inputs_all": (C.c_int, [C.c_void_p, C.POINTER(PlanInputs), _ip, _ip, _ip, _dp, _ip, _dp, _dp]),
    "rp_batch_launch": (C.c_int, [C.c_void_p]),
    "rp_batch_results": (C.c_int, [C.c_void_p, C.POINTER(PlanResult)]),
    "rp_batch_fetch_candidates": (C.c_int, [C.c_void_p, C.c_int, _dp, _ip, _ip, _ip]),
    "rp_batch_winner_states": (C.c_int, [C.c_void_p, C.c_int, _dp]),
    "rp_batch_last_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_longlong)]),
    "rp_initial_states": (C.c_int, [C.c_void_p, C.c_int, _dp, _ip, _dp, _dp, _ip]),
    "rp_batch_initial_states": (C.c_int, [C.c_void_p, _dp, _ip, _dp, _dp, _ip]),
    "rp_selftest_divide": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _ip]),
    "rp_last_stage_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "rp_stage_ms": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "rp_ctx_set_stage_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "rp_measure_fp64_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "rp_launches_per_plan": (C.c_int, [C.c_void_p]),
    "rp_last_main_kernel": (C.c_int, [C.c_void_p]),
}

_lib = None


def load_library():
    """Load librp_b200.so and declare every signature.  Raises RpError if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RpError("CUDA library %s not built -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None and a.size else C.cast(None, t)


@functools.lru_cache(maxsize=8192)
def _traj_len_cached(delta_tau, dt):
    return len(np.arange(0, np.round(delta_tau + dt, 5), dt))


def traj_len_of(delta_tau, dt):
    """len(np.arange(0, np.round(delta_tau + dt, 5), dt))  (reactive_planner.py:733, :748).  The sampled horizons of a
    level repeat every replanning cycle, so the numpy expression is evaluated once per distinct (delta_tau, dt)."""
    return _traj_len_cached(float(delta_tau), float(dt))


class Engine:
    """One device-resident planning context."""

    def __init__(self, device=0, stream=None):
        self._lib = load_library()
        self._ctx = C.c_void_p()
        self._check(self._lib.rp_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(self._ctx)))
        self.device = int(device)
        self._N = None
        self._n_cand = 0
        self.plan_generation = 0      # bumped by every upload; lazy TrajectorySample views check it

    def _check(self, rc):
        if rc != 0:
            msg = self._lib.rp_last_error()
            raise RpError("rp_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self._lib.rp_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- scenario-static tables ----
    def set_vehicle(self, length, width, wb_rear_axle, wheelbase, a_max, v_switch, delta_max, v_delta_max,
                    kappa_max=None):
        if kappa_max is None:
            kappa_max = np.tan(delta_max) / wheelbase            # utility/config.py:222
        vp = VehicleParams(length, width, wb_rear_axle, wheelbase, a_max, v_switch, delta_max, v_delta_max,
                           float(kappa_max))
        self._check(self._lib.rp_ctx_set_vehicle(self._ctx, C.byref(vp)))

    def set_reference(self, ref_pos, ref_theta, ref_curv, ref_curv_d, path_xy, path_s, path_normals, proj_limit):
        arrs = [_f64(a) for a in (ref_pos, ref_theta, ref_curv, ref_curv_d, path_xy, path_s, path_normals)]
        n = arrs[0].shape[0]
        assert all(a.shape[0] == n for a in arrs), "reference arrays must have equal length"
        self._check(self._lib.rp_ctx_set_reference(self._ctx, n, *[_p(a, _dp) for a in arrs], float(proj_limit)))

    def set_reference_polyline(self, xy, proj_limit=20.0, eps2=1e-4):
        """Reference tables derived on the device from the (smoothed, de-duplicated) polyline xy[n][2]."""
        xy = _f64(xy).reshape(-1, 2)
        self._check(self._lib.rp_ctx_set_reference_polyline(self._ctx, len(xy), _p(xy, _dp), float(proj_limit),
                                                            float(eps2)))

    def get_reference(self):
        """The context's reference tables as a dict with the keys of ``CoordinateSystem.device_tables()``."""
        n = C.c_int(0)
        self._check(self._lib.rp_ctx_get_reference(self._ctx, 0, C.byref(n), None, None, None, None, None, None, None))
        n = n.value
        a = {k: np.empty(n, dtype=np.float64) for k in ("ref_pos", "ref_theta", "ref_curv", "ref_curv_d", "path_s")}
        xy, nm = np.empty((n, 2), dtype=np.float64), np.empty((n, 2), dtype=np.float64)
        m = C.c_int(0)
        self._check(self._lib.rp_ctx_get_reference(
            self._ctx, n, C.byref(m), _p(a["ref_pos"], _dp), _p(a["ref_theta"], _dp),
            _p(a["ref_curv"], _dp), _p(a["ref_curv_d"], _dp), _p(xy, _dp), _p(a["path_s"], _dp),
            _p(nm, _dp)))
        a["path_xy"], a["path_normals"] = xy, nm
        return a

    def set_obstacles(self, static_obb=None, dyn_t0=None, dyn_boxes=None, tris=None, cell_size=0.0):
        """static_obb (n,5): cx, cy, theta, half_len, half_wid.  dyn_boxes: list of (K_i,5) arrays."""
        so = _f64(static_obb if static_obb is not None else np.zeros((0, 5))).reshape(-1, 5)
        t0 = _i32(dyn_t0 if dyn_t0 is not None else np.zeros(0))
        boxes = [] if dyn_boxes is None else [_f64(b).reshape(-1, 5) for b in dyn_boxes]
        ln = _i32([len(b) for b in boxes])
        cat = _f64(np.concatenate(boxes, axis=0)) if boxes else np.zeros((0, 5))
        tr = _f64(tris if tris is not None else np.zeros((0, 6))).reshape(-1, 6)
        self._check(self._lib.rp_ctx_set_obstacles(self._ctx, so.shape[0], _p(so, _dp), len(boxes), _p(t0, _ip),
                                                   _p(ln, _ip), _p(cat, _dp), tr.shape[0], _p(tr, _dp),
                                                   float(cell_size)))

    # ---- planning ----
    @staticmethod
    def make_inputs(x0_lon, x0_lat, x0_orientation, x0_time_step, low_vel_mode, lon_mode, N, dt, factor=1,
                    draw_all=False, constraints=("velocity", "acceleration", "kappa", "kappa_dot", "yaw_rate"),
                    cost_kind=COST_DEFAULT, desired_speed=None, desired_s=None, desired_d=0.0, w_a=5.0,
                    want_all_states=False, check_collision=True, continuous_collision_check=False):
        pi = PlanInputs()
        pi.x0_lon[:] = [float(v) for v in x0_lon]
        pi.x0_lat[:] = [float(v) for v in x0_lat]
        pi.x0_orientation = float(x0_orientation)
        pi.x0_time_step = int(x0_time_step)
        pi.low_vel_mode = int(bool(low_vel_mode))
        pi.lon_mode = STOPPING if lon_mode in (STOPPING, "stopping") else VELOCITY_KEEPING
        pi.N = int(N)
        pi.dt = float(dt)
        pi.factor = int(factor)
        pi.draw_all = int(bool(draw_all))
        mask = 0
        for name in constraints:
            mask |= CONSTRAINT_BITS[name]
        pi.constraint_mask = mask
        pi.cost_kind = int(cost_kind)
        pi.has_desired_speed = int(desired_speed is not None)
        pi.has_desired_s = int(desired_s is not None)
        pi.desired_speed = float(desired_speed) if desired_speed is not None else 0.0
        pi.desired_s = float(desired_s) if desired_s is not None else 0.0
        pi.desired_d = float(desired_d)
        pi.w_a = float(w_a)
        pi.want_all_states = int(bool(want_all_states))
        pi.check_collision = int(check_collision) if not isinstance(check_collision, bool) else int(check_collision)
        pi.continuous_collision_check = int(bool(continuous_collision_check))
        return pi

    def _grid_args(self, inputs, t, lon, d, traj_len):
        self._N = inputs.N
        self._cyc_selected = None
        self.plan_generation += 1
        # the same four arrays as in the last call (a replanning loop re-using its sample arrays): the pointer objects are
        # kept -- building them costs more than the rest of the call's host work
        keep = getattr(self, "_keep", None)
        if keep is not None and keep[0] is t and keep[1] is lon and keep[2] is d and keep[3] is traj_len and keep[5] is inputs:
            self._n_cand = keep[6]
            return keep[4]
        t_in, lon_in, d_in, tl_in = t, lon, d, traj_len
        t = _f64(t)
        lon = _f64(lon)
        d = _f64(d)
        if traj_len is None:
            traj_len = [traj_len_of(tt, inputs.dt) for tt in t]
        tl = _i32(traj_len)
        self._n_cand = t.size * lon.size * d.size
        args = (C.byref(inputs), t.size, _p(t, _dp), _p(tl, _ip), lon.size, _p(lon, _dp), d.size, _p(d, _dp))
        same = t is t_in and lon is lon_in and d is d_in and tl is tl_in        # (already contiguous arrays of the right type)
        self._keep = (t, lon, d, tl, args, inputs, self._n_cand) if same else (object(), None, None, None, args, inputs, 0, (t, lon, d, tl))
        return args

    def plan_grid(self, inputs, t, lon, d, traj_len=None):
        res = PlanResult()
        self._check(self._lib.rp_plan_grid(self._ctx, *self._grid_args(inputs, t, lon, d, traj_len), C.byref(res)))
        return res

    def grid_upload(self, inputs, t, lon, d, traj_len=None):
        self._check(self._lib.rp_grid_upload(self._ctx, *self._grid_args(inputs, t, lon, d, traj_len)))

    def grid_launch(self):
        self._check(self._lib.rp_grid_launch(self._ctx))

    def grid_result(self):
        res = PlanResult()
        self._check(self._lib.rp_grid_result(self._ctx, C.byref(res)))
        return res

    # ---- one replanning cycle in one launch (rp_plan_levels) ----
    def cycle_limits(self):
        """(max_levels, max_samples, max_segments, max_work) of one ``plan_levels`` launch."""
        lim = getattr(Engine, "_cycle_limits", None)
        if lim is None:
            a, b, c, d = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
            self._check(self._lib.rp_cycle_limits(C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
            lim = Engine._cycle_limits = (a.value, b.value, c.value, d.value)
        return lim

    def plan_levels(self, inputs, levels):
        """Several sampling levels in ONE launch: ``levels`` = [(t, lon, d, traj_len), ...] in escalation order.
        Returns (records, chosen): ``records[j]`` is the PlanResult of level j for j <= chosen, the first level with a
        winner (else the last).  fetch_* afterwards address the chosen level (``select_level`` switches)."""
        b = getattr(self, "_cyc", None)
        if b is None:
            max_levels, max_samples = self.cycle_limits()[:2]
            b = self._cyc = {"n": np.zeros((3, max_levels), dtype=np.int32), "t": np.zeros(max_samples), "lon": np.zeros(max_samples),
                             "d": np.zeros(max_samples), "tl": np.zeros(max_samples, dtype=np.int32),
                             "res": (PlanResult * max_levels)(), "n_eval": C.c_int32(), "chosen": C.c_int32()}
            b["ptr"] = (b["n"][0].ctypes.data, b["n"][1].ctypes.data, b["n"][2].ctypes.data, b["t"].ctypes.data,
                        b["tl"].ctypes.data, b["lon"].ctypes.data, b["d"].ctypes.data)
            b["refs"] = (C.byref(b["n_eval"]), C.byref(b["chosen"]))
            b["fast"] = None
            if _rp_pack is not None and hasattr(_rp_pack, "plan_levels"):
                # (csrc/rp_pack.c: the same rp_plan_levels call entered through its exported address, no ctypes marshalling)
                b["fast"] = (C.cast(self._lib.rp_plan_levels, C.c_void_p).value,
                             self._ctx.value if hasattr(self._ctx, "value") else int(self._ctx), C.addressof(b["res"]))
        fast = b["fast"]
        if fast is not None:
            try:
                rc, chosen, _, counts = _rp_pack.plan_levels(fast[0], fast[1], C.addressof(inputs), levels, fast[2])
            except (TypeError, ValueError):
                rc = None                  # unusual array types / sizes: the ctypes path below converts or reports them
            if rc is not None:
                self._N = inputs.N
                self.plan_generation += 1
                b["last"] = ()
                if rc != 0:
                    self._check(rc)
                self._level_counts = counts
                self._n_cand = counts[chosen]
                self._cyc_chosen = self._cyc_selected = chosen
                return b["res"], chosen
        n, bt, bl, bd, btl = b["n"], b["t"], b["lon"], b["d"], b["tl"]
        ot = ol = od = 0
        counts, seen = [], []
        last = b.get("last", ())
        for j, (t, lon, d, tl) in enumerate(levels):
            nt, nl, nd = len(t), len(lon), len(d)
            # the sampled horizons / lon samples of a level are usually the very same (cached) arrays as in the last cycle
            # at the same place of the staging buffers: copied again only when they are not
            prev = last[j] if j < len(last) else None
            if prev is not None and (t.flags.writeable or tl.flags.writeable or lon.flags.writeable):
                prev = None                  # (only read-only arrays -- the sampling caches -- are trusted to be unchanged)
            if prev is None or prev[0] is not t or prev[1] is not tl or prev[4] != ot:
                bt[ot:ot + nt] = t
                btl[ot:ot + nt] = tl
                n[0, j] = nt
            if prev is None or prev[2] is not lon or prev[5] != ol:
                bl[ol:ol + nl] = lon
                n[1, j] = nl
            bd[od:od + nd] = d
            if prev is None or prev[3] != nd:
                n[2, j] = nd
            counts.append(nt * nl * nd)
            seen.append((t, tl, lon, nd, ot, ol))
            ot += nt
            ol += nl
            od += nd
        b["last"] = seen
        self._N = inputs.N
        self.plan_generation += 1
        rc = self._lib.rp_plan_levels(self._ctx, C.byref(inputs), len(levels), *b["ptr"], b["res"], *b["refs"])
        if rc != 0:
            self._check(rc)
        chosen = b["chosen"].value
        self._level_counts = counts
        self._n_cand = counts[chosen]
        self._cyc_chosen = self._cyc_selected = chosen
        return b["res"], chosen

    def cycle_winner_states(self):
        """The chosen level's winner states (14 x (N + 1)) of the last ``plan_levels`` call, copied out of the mapped
        result block the kernel wrote (no library call)."""
        n = N_STATE_ROWS * (self._N + 1)
        view = getattr(self, "_cyc_view", None)
        if view is None or view[0] != n:
            ptr, cnt = C.c_void_p(), C.c_int64()
            self._check(self._lib.rp_cycle_host_block(self._ctx, C.byref(ptr), C.byref(cnt)))
            buf = (C.c_double * n).from_address(ptr.value)
            view = self._cyc_view = (n, ptr.value, np.frombuffer(buf, dtype=np.float64).reshape(N_STATE_ROWS, self._N + 1))
        else:
            # the block is re-allocated only when N grows, which changes n as well
            pass
        return view[2].copy()

    def select_level(self, level):
        """After ``plan_levels``: make fetch_states / fetch_candidates / fetch_coeffs address that evaluated level."""
        if getattr(self, "_cyc_selected", None) == level:
            return
        self._check(self._lib.rp_select_level(self._ctx, int(level)))
        self._n_cand = self._level_counts[level]
        self._cyc_selected = level

    def plan_list(self, inputs, coeffs_lon, coeffs_lat, traj_len, skip=None):
        cl = _f64(coeffs_lon).reshape(-1, 6)
        ct = _f64(coeffs_lat).reshape(-1, 6)
        tl = _i32(traj_len)
        sk = np.ascontiguousarray(skip, dtype=np.uint8) if skip is not None else None
        self._N = inputs.N
        self._n_cand = cl.shape[0]
        self._cyc_selected = None
        self.plan_generation += 1
        res = PlanResult()
        self._check(self._lib.rp_plan_list(self._ctx, C.byref(inputs), cl.shape[0], _p(cl, _dp), _p(ct, _dp),
                                           _p(tl, _ip), _p(sk, _bp) if sk is not None else C.cast(None, _bp),
                                           C.byref(res)))
        return res

    def initial_states(self, x0, low_vel_mode):
        """Cartesian rear-axle states x0[n][6] = (x, y, orientation, velocity, acceleration, steering_angle) ->
        (x_0_lon[n][3], x_0_lat[n][3], status[n]) on the device (reactive_planner.py:446-512)."""
        x0 = _f64(x0).reshape(-1, 6)
        lv = _i32(np.broadcast_to(np.asarray(low_vel_mode, dtype=np.int32), (x0.shape[0],)))
        lon, lat = np.empty((x0.shape[0], 3)), np.empty((x0.shape[0], 3))
        status = np.empty(x0.shape[0], dtype=np.int32)
        self._check(self._lib.rp_initial_states(self._ctx, x0.shape[0], _p(x0, _dp), _p(lv, _ip), _p(lon, _dp), _p(lat, _dp),
                                                _p(status, _ip)))
        return lon, lat, status

    def selftest_divide(self, a, b):
        """(shared-reciprocal quotient as the kernels form it, plain a / b, 1 where the range check fell back to the plain
        division) as computed on the device."""
        a, b = _f64(a).ravel(), _f64(b).ravel()
        q1, q2 = np.empty_like(a), np.empty_like(a)
        rej = np.empty(len(a), dtype=np.int32)
        self._check(self._lib.rp_selftest_divide(self._ctx, len(a), _p(a, _dp), _p(b, _dp), _p(q1, _dp), _p(q2, _dp), _p(rej, _ip)))
        return q1, q2, rej

    def set_kernel_policy(self, policy):
        """KERNEL_AUTO / KERNEL_STEP_PARALLEL / KERNEL_CANDIDATE_MAJOR (identical results, different schedule)."""
        self._check(self._lib.rp_ctx_set_kernel_policy(self._ctx, int(policy)))

    def set_candidate_range(self, first, count):
        self._check(self._lib.rp_set_candidate_range(self._ctx, int(first), int(count)))

    def set_candidate_stripe(self, rank, world):
        """lon-interleaved shard of a grid bundle (every rank gets the same mix of horizons); world <= 1 resets"""
        self._check(self._lib.rp_set_candidate_stripe(self._ctx, int(rank), int(world)))

    def export_record_dev(self, dev_ptr):
        self._check(self._lib.rp_export_record_dev(self._ctx, C.c_void_p(int(dev_ptr))))

    def merge_records_dev(self, dev_gathered_ptr, world, dev_winner_ptr, dev_totals_ptr):
        self._check(self._lib.rp_merge_records_dev(self._ctx, C.c_void_p(int(dev_gathered_ptr)), int(world),
                                                   C.c_void_p(int(dev_winner_ptr)), C.c_void_p(int(dev_totals_ptr))))

    def count_colliders_before_dev(self, dev_winner_ptr, dev_out_ptr):
        self._check(self._lib.rp_count_colliders_before_dev(self._ctx, C.c_void_p(int(dev_winner_ptr)),
                                                            C.c_void_p(int(dev_out_ptr))))

    # ---- multi-GPU exchange over peer-mapped memory (rp_peer_*) ----
    def peer_create(self):
        """Allocate this rank's mailbox; returns its 64-byte CUDA IPC handle."""
        buf = C.create_string_buffer(64)
        self._check(self._lib.rp_peer_create(self._ctx, buf))
        return buf.raw

    def peer_open(self, rank, world, handles):
        """handles: the ranks' 64-byte handles concatenated in rank order."""
        handles = bytes(handles)
        if len(handles) != 64 * int(world):
            raise RpError("peer_open: expected %d handle bytes, got %d" % (64 * int(world), len(handles)))
        self._check(self._lib.rp_peer_open(self._ctx, int(rank), int(world), handles))

    def peer_close(self):
        self._check(self._lib.rp_peer_close(self._ctx))

    def synchronize(self):
        self._check(self._lib.rp_ctx_synchronize(self._ctx))

    # ---- results ----
    def fetch_states(self, idx):
        buf = getattr(self, "_states_buf", None)
        if buf is None or buf[0] != self._N:
            arr = np.empty((N_STATE_ROWS, self._N + 1), dtype=np.float64)
            buf = self._states_buf = (self._N, arr, _p(arr, _dp))
        rc = self._lib.rp_fetch_states(self._ctx, int(idx), buf[2])
        if rc != 0:
            self._check(rc)
        return buf[1].copy()

    def fetch_candidates(self):
        n = self._n_cand
        cost = np.empty(n, dtype=np.float64)
        status = np.empty(n, dtype=np.int32)
        reason = np.empty(n, dtype=np.int32)
        step = np.empty(n, dtype=np.int32)
        self._check(self._lib.rp_fetch_candidates(self._ctx, _p(cost, _dp), _p(status, _ip), _p(reason, _ip),
                                                  _p(step, _ip)))
        return cost, status, reason, step

    def fetch_coeffs(self):
        n = self._n_cand
        cl = np.empty((n, 6), dtype=np.float64)
        ct = np.empty((n, 6), dtype=np.float64)
        tau = np.full(n, np.nan, dtype=np.float64)
        self._check(self._lib.rp_fetch_coeffs(self._ctx, _p(cl, _dp), _p(ct, _dp), _p(tau, _dp)))
        return cl, ct, tau

    def solve_coeffs(self, kind, x0, xd, tau):
        kind = _i32(kind)
        x0 = _f64(x0).reshape(-1, 3)
        xd = _f64(xd).reshape(-1, 3)
        tau = _f64(tau)
        out = np.empty((kind.size, 6), dtype=np.float64)
        self._check(self._lib.rp_solve_coeffs(self._ctx, kind.size, _p(kind, _ip), _p(x0, _dp), _p(xd, _dp),
                                              _p(tau, _dp), _p(out, _dp)))
        return out

    def collide_poses(self, pose, time_idx, half_length, half_width):
        pose = _f64(pose).reshape(-1, 3)
        ti = _i32(time_idx)
        hit = np.zeros(pose.shape[0], dtype=np.uint8)
        self._check(self._lib.rp_collide_poses(self._ctx, pose.shape[0], _p(pose, _dp), _p(ti, _ip),
                                               float(half_length), float(half_width), _p(hit, _bp)))
        return hit.astype(bool)

    def set_stage_timing(self, on=True):
        """CUDA-event stage timing of the launches that follow (``stage_ms``); off by default."""
        self._check(self._lib.rp_ctx_set_stage_timing(self._ctx, int(bool(on))))

    def last_stage_ms(self):
        ms = (C.c_float * 4)()
        self._check(self._lib.rp_last_stage_ms(self._ctx, ms))
        return [float(v) for v in ms]

    def stage_ms(self, back=0):
        ms = (C.c_float * 4)()
        self._check(self._lib.rp_stage_ms(self._ctx, int(back), ms))
        return [float(v) for v in ms]

    def measure_fp64_peak(self):
        tf = C.c_double()
        self._check(self._lib.rp_measure_fp64_peak(self._ctx, C.byref(tf)))
        return float(tf.value)

    def launches_per_plan(self):
        return int(self._lib.rp_launches_per_plan(self._ctx))

    def last_main_kernel(self):
        """KERNEL_STEP_PARALLEL or KERNEL_CANDIDATE_MAJOR: which schedule evaluated the last plan's main launch."""
        return int(self._lib.rp_last_main_kernel(self._ctx))


class Batch:
    """Independent scenarios (``Engine`` objects of one device) evaluated by ONE set of launches per replanning
    cycle (rp_batch_* in include/rp_b200.h; BASELINE configs[4])."""

    def __init__(self, engines, stream=None):
        self._lib = load_library()
        self.engines = list(engines)
        arr = (C.c_void_p * len(self.engines))(*[e._ctx for e in self.engines])
        self._b = C.c_void_p()
        self._check(self._lib.rp_batch_create(arr, len(self.engines), C.c_void_p(stream) if stream else None, C.byref(self._b)))
        self._n_cand = [0] * len(self.engines)

    def _check(self, rc):
        if rc != 0:
            msg = self._lib.rp_last_error()
            raise RpError("rp_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))

    def __len__(self):
        return len(self.engines)

    def set_inputs(self, k, inputs, t, lon, d, traj_len=None):
        t, lon, d = _f64(t), _f64(lon), _f64(d)
        if traj_len is None:
            traj_len = [traj_len_of(tt, inputs.dt) for tt in t]
        tl = _i32(traj_len)
        self._n_cand[k] = t.size * lon.size * d.size
        self._check(self._lib.rp_batch_set_inputs(self._b, int(k), C.byref(inputs), t.size, _p(t, _dp), _p(tl, _ip),
                                                  lon.size, _p(lon, _dp), d.size, _p(d, _dp)))

    @staticmethod
    def pack(cycle_inputs):
        """[(rp_plan_inputs, t, lon, d[, traj_len]), ...] -> the packed host form of ``set_inputs_all``: a ctypes array
        of PlanInputs (update x0 / targets in place between cycles), count arrays and concatenated sample lists."""
        n = len(cycle_inputs)
        arr = (PlanInputs * n)()
        ts, lons, ds, tls = [], [], [], []
        for k, item in enumerate(cycle_inputs):
            inputs, t, lon, d = item[:4]
            C.memmove(C.byref(arr[k]), C.byref(inputs), C.sizeof(PlanInputs))
            t = _f64(t)
            ts.append(t); lons.append(_f64(lon)); ds.append(_f64(d))
            tls.append(_i32(item[4]) if len(item) > 4 and item[4] is not None else _i32([traj_len_of(x, inputs.dt) for x in t]))
        cat = lambda xs, dt: np.ascontiguousarray(np.concatenate(xs) if xs else np.zeros(0), dtype=dt)
        return {"inputs": arr, "n_t": _i32([len(x) for x in ts]), "n_lon": _i32([len(x) for x in lons]),
                "n_d": _i32([len(x) for x in ds]), "t": cat(ts, np.float64), "traj_len": cat(tls, np.int32),
                "lon": cat(lons, np.float64), "d": cat(ds, np.float64)}

    def set_inputs_all(self, packed):
        """inputs of every scenario in one call (``pack``)"""
        assert len(packed["inputs"]) == len(self.engines)
        self._n_cand = [int(a) * int(b) * int(c) for a, b, c in zip(packed["n_t"], packed["n_lon"], packed["n_d"])]
        self._check(self._lib.rp_batch_set_inputs_all(self._b, packed["inputs"], _p(packed["n_t"], _ip), _p(packed["n_lon"], _ip),
                                                      _p(packed["n_d"], _ip), _p(packed["t"], _dp), _p(packed["traj_len"], _ip),
                                                      _p(packed["lon"], _dp), _p(packed["d"], _dp)))

    def launch(self):
        self._check(self._lib.rp_batch_launch(self._b))

    def results(self):
        out = (PlanResult * len(self.engines))()
        self._check(self._lib.rp_batch_results(self._b, out))
        return list(out)

    def fetch_candidates(self, k):
        n = self._n_cand[k]
        cost = np.empty(n, dtype=np.float64)
        status, reason, step = (np.empty(n, dtype=np.int32) for _ in range(3))
        self._check(self._lib.rp_batch_fetch_candidates(self._b, int(k), _p(cost, _dp), _p(status, _ip), _p(reason, _ip), _p(step, _ip)))
        return cost, status, reason, step

    def initial_states(self, x0, low_vel_mode):
        """one Cartesian state per scenario, x0[n_scenarios][6] -> (x_0_lon, x_0_lat, status), each scenario against its
        own reference tables (the batched reset() of SURVEY 8f rank 1)"""
        n = len(self.engines)
        x0 = _f64(x0).reshape(n, 6)
        lv = _i32(np.broadcast_to(np.asarray(low_vel_mode, dtype=np.int32), (n,)))
        lon, lat, status = np.empty((n, 3)), np.empty((n, 3)), np.empty(n, dtype=np.int32)
        self._check(self._lib.rp_batch_initial_states(self._b, _p(x0, _dp), _p(lv, _ip), _p(lon, _dp), _p(lat, _dp), _p(status, _ip)))
        return lon, lat, status

    def winner_states(self, step):
        """state of every scenario's winner at time step ``step`` (n, 16): x, y, theta, v, a, kappa, s, s_dot, s_ddot, d,
        d_dot, d_ddot, valid -- where the next replanning cycle of a closed-loop batch starts"""
        out = np.empty((len(self.engines), 16), dtype=np.float64)
        self._check(self._lib.rp_batch_winner_states(self._b, int(step), _p(out, _dp)))
        return out

    def last_ms(self):
        """(device milliseconds of the last launch, candidates it evaluated)"""
        ms, n = C.c_float(), C.c_longlong()
        self._check(self._lib.rp_batch_last_ms(self._b, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def close(self):
        if self._b:
            self._lib.rp_batch_destroy(self._b)
            self._b = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
