"""Shared test plumbing: build oracle-format problems from synthetic scenarios, feed the SAME
arrays through the C-ABI, and compare.  The oracle is only ever the checker here."""
import numpy as np

from commonroad_rp_b200.utility import synthetic
from oracle import rp_oracle as O

STATE_RTOL = 1e-9       # BASELINE.json north_star: states and costs within 1e-9 relative in fp64


def set_order(values):
    """iteration order of ``set(ndarray)`` -- the expression the reference samples with (sampling.py:83,98,116)."""
    return [float(v) for v in set(np.asarray(values, dtype=np.float64))]


def d_order(values, d0):
    """``samples_d.union({x_0_lat[0]})`` iteration order (sampling.py:226)."""
    return [float(v) for v in set(np.asarray(values, dtype=np.float64)).union({d0})]


def level_sets(level, t_min, horizon, dt, v_lo, v_hi, d_lo=-3.0, d_hi=3.0):
    """The reference's per-level sample sets (sampling.py:80-118) as ordered lists."""
    n = 3
    for _ in range(level):
        n = n * 2 - 1
    step = int((1 / (level + 1)) / dt)
    samp = set(np.arange(t_min, round(horizon + dt, 2), step * dt))
    samp.discard(round(horizon + dt, 2))
    t = [float(x) for x in samp]
    v = set_order(np.linspace(v_lo, v_hi, n))
    d = set(np.linspace(d_lo, d_hi, n))
    return t, v, d


def make_problem(scn, t, lon, d, x0_lon, x0_lat, N=20, dt=0.1, lon_mode="velocity_keeping", low_vel_mode=False,
                 x0_orientation=None, x0_time_step=0, factor=1, draw_all=False, constraints=O.CONSTRAINTS,
                 desired_speed=15.0, desired_s=None, desired_d=0.0, w_a=5, cost_kind="default", smooth=True,
                 tables=None):
    ref, ccosy, _ = tables if tables is not None else O.reference_tables(scn["ref_path"], smooth=smooth)
    if x0_orientation is None:
        j = int(np.argmax(ref["ref_pos"] > x0_lon[0])) - 1
        x0_orientation = float(ref["ref_theta"][j])
    return {
        "t": np.asarray(t, dtype=np.float64), "lon": np.asarray(lon, dtype=np.float64),
        "d": np.asarray(d, dtype=np.float64),
        "x0_lon": np.asarray(x0_lon, dtype=np.float64), "x0_lat": np.asarray(x0_lat, dtype=np.float64),
        "x0_orientation": float(x0_orientation), "x0_time_step": int(x0_time_step),
        "lon_mode": lon_mode, "low_vel_mode": bool(low_vel_mode), "dt": dt, "N": N, "factor": factor,
        "draw_all": bool(draw_all), "constraints": tuple(constraints),
        "cost": {"kind": cost_kind, "desired_speed": desired_speed, "desired_s": desired_s, "desired_d": desired_d,
                 "w_a": w_a},
        "vehicle": O.vehicle_dict(), "ref": ref, "ccosy": ccosy,
        "obstacles": {k: scn[k] for k in ("static_boxes", "dyn_t0", "dyn_states", "dyn_lw", "boundary_boxes",
                                          "boundary_tris")},
    }


def obstacle_arrays(obst, continuous=False):
    """scenario-dict obstacles -> the C-ABI's (static_obb, dyn_t0, dyn_boxes, tris); with the continuous collision
    check the dynamic obstacles are uploaded as their OBB-sum hulls (reference :240-241)."""
    sb = np.asarray(obst.get("static_boxes", np.zeros((0, 5))), dtype=np.float64).reshape(-1, 5)
    static = np.stack([sb[:, 0], sb[:, 1], sb[:, 2], 0.5 * sb[:, 3], 0.5 * sb[:, 4]], axis=1)
    bb = np.asarray(obst.get("boundary_boxes", np.zeros((0, 5))), dtype=np.float64).reshape(-1, 5)
    static = np.concatenate([static, bb], axis=0)
    dyn_boxes = []
    for st, lw in zip(obst.get("dyn_states", ()), obst.get("dyn_lw", ())):
        st = np.asarray(st, dtype=np.float64).reshape(-1, 3)
        boxes = np.concatenate([st, np.full((len(st), 1), 0.5 * lw[0]), np.full((len(st), 1), 0.5 * lw[1])], axis=1)
        if continuous:
            from commonroad_rp_b200 import collision
            tvo = collision.TimeVariantCollisionObject(0)
            for cx, cy, th, hl, hw in boxes:
                tvo.append_obstacle(collision.RectOBB(hl, hw, th, cx, cy))
            hull, err = collision.trajectory_preprocess_obb_sum(tvo)
            boxes = np.asarray([hull.obstacle_at_time(k).row() for k in range(len(boxes) - 1)], dtype=np.float64).reshape(-1, 5)
        dyn_boxes.append(boxes)
    tris = np.asarray(obst.get("boundary_tris", np.zeros((0, 6))), dtype=np.float64).reshape(-1, 6)
    return static, np.asarray(obst.get("dyn_t0", ()), dtype=np.int32), dyn_boxes, tris


def engine_for(prob, device=0, stream=None):
    from commonroad_rp_b200._lib import Engine
    eng = Engine(device, stream)
    v = prob["vehicle"]
    eng.set_vehicle(v["length"], v["width"], v["wb_rear_axle"], v["wheelbase"], v["a_max"], v["v_switch"],
                    v["delta_max"], v["v_delta_max"])
    r, c = prob["ref"], prob["ccosy"]
    eng.set_reference(r["ref_pos"], r["ref_theta"], r["ref_curv"], r["ref_curv_d"], c["path"], c["S"], c["normals"],
                      c["limit"])
    static, t0, dyn_boxes, tris = obstacle_arrays(prob["obstacles"], prob.get("continuous", False))
    eng.set_obstacles(static, t0, dyn_boxes, tris)
    return eng


def inputs_for(prob, want_all_states=False, check_collision=True):
    from commonroad_rp_b200 import _lib
    cost = prob["cost"]
    kind = {"default": _lib.COST_DEFAULT, "failsafe": _lib.COST_FAILSAFE}[cost.get("kind", "default")]
    return _lib.Engine.make_inputs(
        prob["x0_lon"], prob["x0_lat"], prob["x0_orientation"], prob["x0_time_step"], prob["low_vel_mode"],
        prob["lon_mode"], prob["N"], prob["dt"], factor=prob["factor"], draw_all=prob.get("draw_all", False),
        constraints=prob["constraints"], cost_kind=kind, desired_speed=cost.get("desired_speed"),
        desired_s=cost.get("desired_s"), desired_d=cost.get("desired_d", 0.0), w_a=cost.get("w_a", 5),
        want_all_states=want_all_states, check_collision=check_collision,
        continuous_collision_check=prob.get("continuous", False))


def run_engine_grid(eng, prob, want_all_states=True, kernel=None, check_collision=True):
    """kernel: None (context default) or _lib.KERNEL_* -- which schedule evaluates the main launch."""
    if kernel is not None:
        eng.set_kernel_policy(kernel)
    res = eng.plan_grid(inputs_for(prob, want_all_states, check_collision), prob["t"], prob["lon"], prob["d"])
    cost, status, reason, step = eng.fetch_candidates()
    cl, ct, tau = eng.fetch_coeffs()
    out = {
        "n": res.n_candidates, "winner": res.winner, "winner_cost": res.winner_cost,
        "n_feasible": res.n_feasible, "n_infeasible_kinematics": res.n_infeasible_kinematics,
        "n_infeasible_collision": res.n_infeasible_collision, "n_collision_total": res.n_collision_total,
        "reason_counts": list(res.reason_counts), "cost": cost, "status": status, "reason": reason, "step": step,
        "coeffs_lon": cl, "coeffs_lat": ct, "delta_tau_lat": tau,
    }
    if res.winner >= 0:
        out["winner_states"] = eng.fetch_states(res.winner)
    if want_all_states:
        out["states"] = np.stack([eng.fetch_states(k) for k in range(res.n_candidates)]) if res.n_candidates else None
    return out


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(a), np.abs(b))
    err = np.abs(a - b)
    # absolute floor: quantities that are differences of O(1..100) numbers (angles, offsets near zero)
    return float(np.max(err / np.maximum(scale, 1.0))) if err.size else 0.0


def assert_parity(o, g, prob, check_states=True, tag=""):
    """oracle output ``o`` (rp_oracle.plan_grid, full_collision=True) vs engine output ``g``."""
    n = o["n"]
    assert g["n"] == n, tag
    # coefficients: LAPACK LU vs in-register LU; conditioning bounds this by ~1e-10 relative
    assert rel_err(o["coeffs_lon"], g["coeffs_lon"]) < 1e-9, tag
    assert rel_err(o["coeffs_lat"], g["coeffs_lat"]) < 1e-9, tag
    # flags: bit exact
    o_status = o["status"]
    assert np.array_equal(o_status, g["status"]), "%s status mismatch at %s" % (tag, np.nonzero(o_status != g["status"])[0][:10])
    kin = o_status == O.ST_KINEMATIC
    assert np.array_equal(o["reason"][kin], g["reason"][kin]), tag
    assert np.array_equal(o["bad_step"][kin], g["step"][kin]), tag
    col = o_status == O.ST_COLLISION
    assert np.array_equal(o["collide_step"][col], g["step"][col]), tag
    # selected index and counters: bit exact
    assert o["winner"] == g["winner"], tag
    assert o["n_infeasible_kinematics"] == g["n_infeasible_kinematics"], tag
    assert o["n_infeasible_collision"] == g["n_infeasible_collision"], tag
    from commonroad_rp_b200._lib import REASON_NAMES
    for name, cnt in o["reasons"].items():
        assert g["reason_counts"][REASON_NAMES.index(name)] == cnt, (tag, name)
    # costs within 1e-9 relative
    feas = (o_status == O.ST_FEASIBLE) | (o_status == O.ST_COLLISION)
    if feas.any():
        assert rel_err(o["cost"][feas], g["cost"][feas]) < STATE_RTOL, tag
    if check_states and o.get("states") is not None and g.get("states") is not None:
        have = ~np.isnan(o["states"][:, 0, 0])
        if have.any():
            assert rel_err(o["states"][have], g["states"][have]) < STATE_RTOL, tag
    if o["winner"] >= 0 and o.get("states") is not None:
        assert rel_err(o["states"][o["winner"]], g["winner_states"]) < STATE_RTOL, tag


class EmptyScenario:
    """config.scenario stand-in for planners that receive their collision checker explicitly"""
    static_obstacles, dynamic_obstacles = (), ()
    lanelet_network = type("LN", (), {"lanelets": ()})()


def planner_from_fixture(prob, ref_path_raw, x0_row, desired_velocity=None, continuous=False, draw=False, cost=None,
                         low_vel_mode_threshold=4.0, t_min=0.4):
    """The drop-in ``ReactivePlanner`` set up on a fixture's problem: reference path from the RAW polyline (the
    product's own CoordinateSystem), obstacles from the scenario arrays, x0 = (x, y, orientation, velocity,
    acceleration, yaw_rate, steering_angle, time_step) with the fixture's curvilinear initial state."""
    from commonroad_rp_b200 import collision
    from commonroad_rp_b200.reactive_planner import ReactivePlanner
    from commonroad_rp_b200.state import ReactivePlannerState
    from commonroad_rp_b200.utility.config import ReactivePlannerConfiguration
    from commonroad_rp_b200.utility.utils_coordinate_system import CoordinateSystem
    cfg = ReactivePlannerConfiguration()
    cfg.planning.time_steps_computation = prob["N"]
    cfg.planning.dt = prob["dt"]
    cfg.planning.planning_horizon = prob["dt"] * prob["N"]
    cfg.planning.factor = prob["factor"]
    cfg.planning.low_vel_mode_threshold = low_vel_mode_threshold
    cfg.planning.continuous_collision_check = bool(continuous)
    cfg.planning.constraints_to_check = list(prob["constraints"])
    cfg.sampling.longitudinal_mode = prob["lon_mode"]
    cfg.sampling.t_min = t_min
    cfg.debug.draw_traj_set = draw
    cfg.debug.save_plots = draw
    ob = prob["obstacles"]
    scn = EmptyScenario()
    if continuous:
        # the planner builds the checker itself (dynamic obstacles replaced by their OBB-sum hulls, reference :239-243)
        class _St:
            def __init__(self, pos, th, t):
                self.position, self.orientation, self.time_step = pos, th, t

        class _Ob:
            pass

        statics, dyns = [], []
        for cx, cy, th, l, w in np.asarray(ob["static_boxes"], dtype=np.float64).reshape(-1, 5):
            o = _Ob()
            o.obstacle_shape = type("R", (), {"length": l, "width": w})()
            o.initial_state = _St(np.array([cx, cy]), th, 0)
            o.prediction = None
            statics.append(o)
        for t0, st, lw in zip(ob["dyn_t0"], ob["dyn_states"], ob["dyn_lw"]):
            o = _Ob()
            o.obstacle_shape = type("R", (), {"length": lw[0], "width": lw[1]})()
            st = np.asarray(st, dtype=np.float64).reshape(-1, 3)
            o.initial_state = _St(st[0, :2], st[0, 2], int(t0))
            traj = type("T", (), {"state_list": [_St(r[:2], r[2], int(t0) + 1 + k) for k, r in enumerate(st[1:])]})()
            o.prediction = type("P", (), {"trajectory": traj})()
            dyns.append(o)
        scn = type("Scn", (), {"static_obstacles": statics, "dynamic_obstacles": dyns,
                               "lanelet_network": EmptyScenario.lanelet_network})()
    cfg.update(scenario=scn, planning_problem=None)
    planner = ReactivePlanner(cfg)
    co = CoordinateSystem(np.asarray(ref_path_raw, dtype=np.float64))
    if continuous:
        sg = collision.ShapeGroup()
        for cx, cy, th, hl, hw in np.asarray(ob["boundary_boxes"], dtype=np.float64).reshape(-1, 5):
            sg.add_shape(collision.RectOBB(hl, hw, th, cx, cy))
        for tri in np.asarray(ob.get("boundary_tris", np.zeros((0, 6))), dtype=np.float64).reshape(-1, 6):
            sg.add_shape(collision.Triangle(*tri))
        planner.set_collision_checker(scenario=scn, road_boundary_obstacle=sg)
        cc = planner.collision_checker
    else:
        cc = collision.checker_from_arrays(**{k: ob[k] for k in ("static_boxes", "dyn_t0", "dyn_states", "dyn_lw",
                                                                  "boundary_boxes", "boundary_tris")})
    if cost == "failsafe":
        from commonroad_rp_b200.cost_function import DefaultCostFunctionFailSafe
        planner.set_cost_function(DefaultCostFunctionFailSafe())
    x = np.asarray(x0_row, dtype=np.float64)
    x0 = ReactivePlannerState(time_step=int(x[7]), position=np.array([x[0], x[1]]), orientation=x[2], velocity=x[3],
                              acceleration=x[4], yaw_rate=x[5], steering_angle=x[6])
    planner.reset(initial_state_cart=x0, initial_state_curv=(list(prob["x0_lon"]), list(prob["x0_lat"])),
                  collision_checker=cc, coordinate_system=co)
    if prob["lon_mode"] == "stopping":
        planner.set_desired_lon_position(prob["cost"]["desired_s"])
    else:
        dv = prob["cost"].get("desired_speed") if desired_velocity is None else desired_velocity
        planner.set_desired_velocity(desired_velocity=dv if dv is not None else float(x[3]), current_speed=float(x[3]))
    return planner


def plan_output_arrays(result):
    """plan()'s return value as the arrays of oracle/make_golden.py:plan_output_arrays"""
    cart, curv, lon_list, lat_list = result
    oc = np.array([[st.position[0], st.position[1], st.orientation, st.velocity, st.acceleration, st.yaw_rate,
                    st.steering_angle, float(st.time_step)] for st in cart.state_list]).T
    ou = np.array([[st.position[0], st.position[1], st.orientation, st.velocity, st.acceleration, st.yaw_rate,
                    float(st.time_step)] for st in curv.state_list]).T
    return {"out_cart": oc, "out_curv": ou, "out_lon": np.array(lon_list, dtype=np.float64),
            "out_lat": np.array(lat_list, dtype=np.float64)}
