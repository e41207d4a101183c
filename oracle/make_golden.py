"""TEST INFRASTRUCTURE ONLY -- generates the fixtures under tests/golden/ by executing the REFERENCE's
own code (/root/reference/commonroad_rp, imported read-only through oracle/ref_shims.py).  Run it in
the build container:  python -m oracle.make_golden

  syn_*.npz   one sampling level of a seeded synthetic scenario: the complete problem (inputs) and the
              reference's per-candidate coefficients / feasibility / costs / winner / counters, plus the
              state blocks of the winner and of a few feasible candidates
  cyc_*.npz   cyclic replanning (run_planner.py loop) on the reference's bundled scenarios, read from
              /root/reference/example_scenarios with oracle/scenario_xml.py: per cycle the initial state and
              every evaluated level's verdicts, the selected index and the winner's states

The fixtures pin (a) oracle/rp_oracle.py and (b) the CUDA path on machines where the reference is absent.
Third-party arithmetic inside them comes from oracle/third_party.py (parity unpinned).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from commonroad_rp_b200.utility import synthetic  # noqa: E402
from oracle import ref_harness as H  # noqa: E402
from oracle import ref_shims, scenario_xml  # noqa: E402
from tests import golden_io  # noqa: E402

OUT = os.environ.get("RP_GOLDEN_OUT", os.path.join(ROOT, "tests", "golden"))
N_STATE_SAMPLES = 24

SYN_CASES = {
    "lvl1_N20": dict(seed=0, level=1, N=20),
    "lvl2_N20": dict(seed=0, level=2, N=20),
    "lvl3_N20": dict(seed=1, level=3, N=20, d0=-0.4),
    "lvl2_N60": dict(seed=2, level=2, N=60, s_dot0=12.0),
    "lowvel": dict(seed=3, level=2, N=20, s_dot0=3.0),
    "standstill_carry": dict(seed=4, level=2, N=30, s_dot0=1.0, low_vel_threshold=0.5),
    "draw_all": dict(seed=5, level=2, N=20, draw=True),
    "stopping": dict(seed=6, level=2, N=30, s_dot0=8.0, mode="stopping", desired_s=24.0),
    "straight": dict(seed=7, level=2, N=20, amplitude=0.0, static_offset=1.0),
    "time_offset": dict(seed=8, level=1, N=20, time_step=30),
    "dense_small": dict(seed=0, level=1, N=60, dense=(6, 9, 9)),
    # continuous collision check (reactive_planner.py:240-241, :1049-1058): OBB-sum hulls of consecutive poses
    "continuous_pass": dict(seed=0, level=2, N=20, amplitude=0.0, continuous=True, crossing=(40.0, 1.0)),
    "continuous_hit": dict(seed=0, level=2, N=20, amplitude=0.0, continuous=True, crossing=(40.0, 2.0)),
    # DefaultCostFunctionFailSafe (cost_function.py:74-92): N + 1 = 21 (np.sum remainder of 5) and N + 1 = 8 (no remainder)
    "failsafe": dict(seed=9, level=2, N=20, cost="failsafe"),
    "failsafe_N7": dict(seed=9, level=1, N=7, cost="failsafe", t_min=0.2),
    # road boundary as TRIANGLES close enough to the lane that outer candidates hit it (box / triangle SAT in a bundle)
    "tri_boundary": dict(seed=10, level=2, N=20, boundary_kind="tris", boundary_offset=3.4),
    "tri_boundary_N60": dict(seed=11, level=1, N=60, boundary_kind="tris", boundary_offset=3.0, s_dot0=10.0),
}

# large bundles, stored compactly (no per-candidate coefficients, states of a seeded sample only):
#   dense_full   BASELINE configs[3] exactly as bench.py builds it: 64 d x 64 v x 32 t, N = 60 -> 131 072 candidates
#   batch_sid*   BASELINE configs[4]'s per-scenario bundle exactly as bench.py's scenario-batch leg builds it:
#                synthetic.scenario_seeded(sid), default level-3 grid at N = 60 -> 29 x 17 x 18 = 8 874 candidates
BIG_CASES = {
    "dense_full": dict(kind="dense"),
    "batch_sid0": dict(kind="batch", sid=0),
    "batch_sid5": dict(kind="batch", sid=5),
}


def reference_result_arrays(r, rng):
    feas = np.nonzero(r["kin_feasible"])[0]
    pick = feas if len(feas) <= N_STATE_SAMPLES else np.sort(rng.choice(feas, N_STATE_SAMPLES, replace=False))
    if r["winner"] >= 0 and r["winner"] not in pick:
        pick = np.sort(np.append(pick, r["winner"]))
    out = {"coeffs_lon": r["coeffs_lon"], "coeffs_lat": r["coeffs_lat"], "delta_tau_lat": r["delta_tau_lat"],
           "kept": r["kept"], "kin_feasible": r["kin_feasible"], "cost": r["cost"],
           "label": np.array([{"none": 0, "feasible": 1, "infeasible_kinematic": 2, "infeasible_collision": 3}[x]
                              for x in r["label"]], dtype=np.int8),
           "winner": np.array(r["winner"]), "n_inf_kin": np.array(r["n_infeasible_kinematics"]),
           "n_inf_col": np.array(r["n_infeasible_collision"]), "reasons": np.array(json.dumps(r["reasons"])),
           "state_idx": pick.astype(np.int64), "states": r["states"][pick]}
    return out


def make_synthetic(name, seed, level, N, s_dot0=15.0, d0=0.3, mode="velocity_keeping", desired_s=None, draw=False,
                   amplitude=20.0, static_offset=0.0, time_step=0, low_vel_threshold=4.0, dense=None, continuous=False,
                   factor=1, crossing=None, cost="default", boundary_kind="boxes", boundary_offset=5.25, t_min=0.4):
    scn = synthetic.make_scenario(seed=seed, amplitude=amplitude, static_offset=static_offset, boundary_kind=boundary_kind,
                                  boundary_offset=boundary_offset)
    if crossing is not None:
        scn = synthetic.add_crossing_obstacle(scn, speed=crossing[0], phase=crossing[1])
    p = H.build_planner(scn, N=N, longitudinal_mode=mode, draw_traj_set=draw, low_vel_mode_threshold=low_vel_threshold,
                        continuous_collision_check=continuous, factor=factor, t_min=t_min)
    if cost == "failsafe":
        from commonroad_rp.cost_function import DefaultCostFunctionFailSafe
        p.set_cost_function(DefaultCostFunctionFailSafe())
    s0 = float(p.coordinate_system.ref_pos[10])
    H.set_initial_state(p, [s0, s_dot0, 0.0], [d0, 0.0, 0.0], time_step=time_step)
    if mode == "stopping":
        p.set_desired_lon_position(desired_s)
    else:
        p.set_desired_velocity(desired_velocity=s_dot0, current_speed=s_dot0)
    if dense is not None:
        n_t, n_v, n_d = dense
        t, v, d, d_on = synthetic.dense_grid(n_t=n_t, n_v=n_v, n_d=n_d, t_first_step=N - n_t + 1)
        H.override_sample_sets(p, level, t=t, v=v, d=d)
    prob = H.problem_from_planner(p, level, scn)
    r = H.evaluate_level(p, level)
    arrays = golden_io.pack_problem(prob)
    arrays.update({"r_" + k: v for k, v in reference_result_arrays(r, np.random.default_rng(seed)).items()})
    np.savez_compressed(os.path.join(OUT, "syn_%s.npz" % name), **arrays)
    print("syn_%s: n=%d feasible=%d winner=%d n_col=%d" % (name, r["n"], int(r["kin_feasible"].sum()), r["winner"],
                                                          r["n_infeasible_collision"]))


def plan_output_arrays(result):
    """plan()'s return value (reactive_planner.py:514-568) as arrays: rows of out_cart = x, y, orientation, velocity,
    acceleration, yaw_rate, steering_angle, time_step of the Cartesian state list; rows of out_curv = s, d, orientation,
    velocity, acceleration, yaw_rate, time_step of the curvilinear one; out_lon / out_lat = the two state lists."""
    cart, curv, lon_list, lat_list = result
    oc = np.array([[st.position[0], st.position[1], st.orientation, st.velocity, st.acceleration, st.yaw_rate,
                    st.steering_angle, float(st.time_step)] for st in cart.state_list]).T
    ou = np.array([[st.position[0], st.position[1], st.orientation, st.velocity, st.acceleration, st.yaw_rate,
                    float(st.time_step)] for st in curv.state_list]).T
    return {"out_cart": oc, "out_curv": ou, "out_lon": np.array(lon_list, dtype=np.float64),
            "out_lat": np.array(lat_list, dtype=np.float64)}


def make_big(name, kind, sid=0):
    """big_<name>.npz: problem + the reference's verdicts of a LARGE bundle in compact form."""
    import time
    if kind == "dense":
        scn = synthetic.make_scenario(seed=0)
        t, v, d, d0 = synthetic.dense_grid()
        s_dot0, level, N = 15.0, 1, 60
    else:
        scn, s_dot0, d0 = synthetic.scenario_seeded(sid)
        level, N = 3, 60
    p = H.build_planner(scn, N=N)
    s0 = float(p.coordinate_system.ref_pos[10])
    H.set_initial_state(p, [s0, s_dot0, 0.0], [d0, 0.0, 0.0])
    p.set_desired_velocity(desired_velocity=s_dot0, current_speed=s_dot0)
    if kind == "dense":
        H.override_sample_sets(p, level, t=t, v=v, d=d)
    prob = H.problem_from_planner(p, level, scn)
    t0 = time.time()
    r = H.evaluate_level(p, level, want_states=False, state_sample=(48, 12345))
    arrays = golden_io.pack_problem(prob)
    lab = np.array([{"none": 0, "feasible": 1, "infeasible_kinematic": 2, "infeasible_collision": 3}[x] for x in r["label"]],
                   dtype=np.int8)
    arrays.update({"r_kin_feasible": r["kin_feasible"], "r_cost": r["cost"], "r_label": lab, "r_winner": np.array(r["winner"]),
                   "r_n_inf_kin": np.array(r["n_infeasible_kinematics"]), "r_n_inf_col": np.array(r["n_infeasible_collision"]),
                   "r_reasons": np.array(json.dumps(r["reasons"])), "r_state_idx": r["state_idx"],
                   "r_states": r["states_sampled"]})
    np.savez_compressed(os.path.join(OUT, "big_%s.npz" % name), **arrays)
    print("big_%s: n=%d feasible=%d winner=%d n_kin=%d n_col=%d  (%.0f s)" % (
        name, r["n"], int(r["kin_feasible"].sum()), r["winner"], r["n_infeasible_kinematics"], r["n_infeasible_collision"],
        time.time() - t0))


# plan()-level fixtures on synthetic scenarios: the complete return value of ReactivePlanner.plan()
# (reactive_planner.py:570-665) incl. the standstill branch (:638-653, :667-713)
PLAN_CASES = {
    # stopped ego with a wall across its own box: every candidate of levels 1..3 collides at step 0 -> no optimum ->
    # standstill trajectory
    "standstill_blocked": dict(seed=20, N=20, s_dot0=0.0, block_ahead=1.0),
    # stopped ego, wall 4 m ahead: the only collision-free candidates stay put (v[lookahead] <= 0.05) -> standstill
    "standstill_wall_ahead": dict(seed=20, N=20, s_dot0=0.0, block_ahead=4.0),
    # stopped ego with desired velocity 0: the optimum does not move (v[lookahead] <= 0.05) -> standstill trajectory
    "standstill_stay": dict(seed=21, N=20, s_dot0=0.0, desired_velocity=0.0),
    # moving ego, blocked: levels 1..3 all fail, plan() returns None
    "blocked_moving": dict(seed=22, N=20, s_dot0=6.0, block_ahead=5.0),
    # ordinary cycle (level escalation not needed)
    "free": dict(seed=23, N=30, s_dot0=11.0),
}


def make_plan(name, seed, N, s_dot0, block_ahead=None, desired_velocity=None, d0=0.2):
    scn = synthetic.make_scenario(seed=seed, amplitude=8.0, wavelength=60.0, n_dynamic=2)
    p = H.build_planner(scn, N=N)
    co = p.coordinate_system
    s0 = float(co.ref_pos[10])
    if block_ahead is not None:
        # a wall across the road: boxes side by side on the normal through s0 + block_ahead
        j = int(np.argmax(co.ref_pos > s0 + block_ahead))
        th = float(co.ref_theta[j])
        c = np.asarray(co.convert_to_cartesian_coords(float(co.ref_pos[j]), 0.0))
        scn = dict(scn)
        scn["static_boxes"] = np.vstack([scn["static_boxes"], [[c[0], c[1], th + np.pi / 2, 14.0, 1.0]]])
        p = H.build_planner(scn, N=N)
    H.set_initial_state(p, [s0, s_dot0, 0.0], [d0, 0.0, 0.0])
    dv = s_dot0 if desired_velocity is None else desired_velocity
    p.set_desired_velocity(desired_velocity=max(dv, 0.0) if desired_velocity is not None else max(s_dot0, 5.0), current_speed=s_dot0)
    level_log = []
    orig_opt = p._get_optimal_trajectory

    def optimal(bundle):
        n = len(bundle.trajectories)
        win = orig_opt(bundle)
        level_log.append({"n": n, "found": win is not None, "n_inf_kin": int(p.infeasible_count_kinematics),
                          "n_inf_col": int(p.infeasible_count_collision), "reasons": dict(p.infeasible_reason_dict)})
        return win

    p._get_optimal_trajectory = optimal
    try:
        out = p.plan()
        err = None
    except Exception as exc:                                  # noqa: BLE001 -- the reference's own failure is the fixture
        out, err = None, "%s: %s" % (type(exc).__name__, exc)
    prob = H.problem_from_planner(p, 1, scn)
    arrays = golden_io.pack_problem(prob)
    x0 = p.x_0
    arrays["x0"] = np.array([x0.position[0], x0.position[1], x0.orientation, x0.velocity, x0.acceleration, x0.yaw_rate,
                             x0.steering_angle, float(x0.time_step)])
    arrays["ref_path_raw"] = np.asarray(scn["ref_path"], dtype=np.float64)
    meta = {"name": name, "N": N, "ok": out is not None, "error": err, "levels": level_log,
            "desired_velocity": float(p._desired_speed), "optimal_cost": float(p.optimal_cost),
            "standstill": bool(out is not None and all(abs(st.velocity) == 0.0 for st in out[0].state_list))}
    if out is not None:
        arrays.update(plan_output_arrays(out))
    arrays["plan_meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, "plan_%s.npz" % name), **arrays)
    print("plan_%s: ok=%s standstill=%s err=%s levels=%s" % (name, meta["ok"], meta["standstill"], err,
                                                              [(lv["n"], lv["found"]) for lv in level_log]))


# ---- bundled scenarios, cyclic replanning (run_planner.py:55-107) -----------------------------------
CYC_CASES = {
    # every case runs until goal_reached() (or the planner gives up), like run_planner.py:61
    "ZAM_Over-1_1": dict(route=[1000], yaml="ZAM_Over-1_1.yaml", max_cycles=400),
    "DEU_Test-1_1_T-1": dict(route=[1, 3], yaml="DEU_Test-1_1_T-1.yaml", max_cycles=400),
    "ZAM_Tjunction-1_42_T-1": dict(route=[50195, 50209, 50203], yaml="ZAM_Tjunction-1_42_T-1.yaml", max_cycles=400),
    # no YAML in configurations/: the ego starts at v = 0 (low-velocity mode, standstill branch of plan(), :638-653);
    # road boundary as triangles; the centre line is extended 5 m backwards so that the rear axle projects onto it
    "ZAM-Ramp-1_1-T-1": dict(route=[5, 6, 7, 8], yaml=None, max_cycles=400, boundary_method="triangulation", extend_back=5.0,
                             config=dict(planning=dict(time_steps_computation=30, replanning_frequency=3),
                                         sampling=dict(t_min=0.4))),
}


def _yaml(path):
    import yaml
    with open(path) as f:
        return yaml.safe_load(f)


def _route_path(sc, route, extend_back=0.0):
    path = scenario_xml.route_centerline(sc["lanelets"], route)
    if extend_back > 0.0:
        u = (path[0] - path[1]) / np.hypot(*(path[0] - path[1]))
        extra = [path[0] + u * k for k in range(int(extend_back), 0, -1)]
        path = np.vstack([np.array(extra), path])
    return path


def make_cyclic(name, route, yaml, max_cycles, boundary_method="obb_rectangles", extend_back=0.0, config=None):
    from types import SimpleNamespace
    ref_root = ref_shims.REFERENCE_ROOT
    sc = scenario_xml.load(os.path.join(ref_root, "example_scenarios", name + ".xml"), boundary_method=boundary_method)
    cfg_y = _yaml(os.path.join(ref_root, "configurations", yaml)) if yaml else dict(config)
    scn = dict(sc["scn"])
    scn["ref_path"] = _route_path(sc, route, extend_back)
    plan_y, samp_y, dbg_y = cfg_y.get("planning", {}), cfg_y.get("sampling", {}), cfg_y.get("debug", {})
    draw = bool(dbg_y.get("draw_traj_set", False) and (dbg_y.get("show_plots", False) or dbg_y.get("save_plots", False)))
    planner = H.build_planner(scn, N=plan_y.get("time_steps_computation", 60), dt=plan_y.get("dt", 0.1),
                              t_min=samp_y.get("t_min", 0.4),
                              low_vel_mode_threshold=plan_y.get("low_vel_mode_threshold", 4.0), draw_traj_set=draw)
    cfg = planner.config
    freq = plan_y.get("replanning_frequency", 3)
    ini = sc["initial_state"]
    veh = planner.vehicle_params
    from commonroad_rp.state import ReactivePlannerState
    th = ini["orientation"]
    x0 = ReactivePlannerState(time_step=ini["time"],
                              position=np.array([ini["x"] - veh.wb_rear_axle * np.cos(th),
                                                 ini["y"] - veh.wb_rear_axle * np.sin(th)]),
                              steering_angle=np.arctan2(veh.wheelbase * ini["yaw_rate"], ini["velocity"]),
                              velocity=ini["velocity"], orientation=th, acceleration=ini["acceleration"],
                              yaw_rate=ini["yaw_rate"])
    planner.reset(initial_state_cart=x0, collision_checker=planner.collision_checker,
                  coordinate_system=planner.coordinate_system)
    cfg.planning_problem = SimpleNamespace(goal=scenario_xml.GoalStub(sc))
    desired_v = scenario_xml.desired_velocity(sc)

    records = []
    level_log = []
    orig_create = planner._create_trajectory_bundle
    orig_opt = planner._get_optimal_trajectory

    def create(x_0_lon, x_0_lat, samp_level):
        level_log.append({"level": samp_level})
        return orig_create(x_0_lon, x_0_lat, samp_level=samp_level)

    def optimal(bundle):
        cands = list(bundle.trajectories)
        win = orig_opt(bundle)
        feas = {id(t) for t in bundle.trajectories}
        lv = level_log[-1]
        lv["n"] = len(cands)
        lv["kin_feasible"] = np.array([id(c) in feas for c in cands], dtype=bool)
        lv["cost"] = np.array([float(c.cost) if id(c) in feas else np.nan for c in cands])
        lv["winner"] = -1 if win is None else next(i for i, c in enumerate(cands) if c is win)
        lv["n_inf_kin"] = int(planner.infeasible_count_kinematics)
        lv["n_inf_col"] = int(planner.infeasible_count_collision)
        lv["reasons"] = dict(planner.infeasible_reason_dict)
        if win is not None:
            ca, cu = win.cartesian, win.curvilinear
            lv["winner_states"] = np.stack([ca.x, ca.y, ca.theta, ca.v, ca.a, ca.kappa, ca.kappa_dot, cu.s, cu.d,
                                            cu.theta, cu.s_dot, cu.s_ddot, cu.d_dot, cu.d_ddot])
        return win

    planner._create_trajectory_bundle = create
    planner._get_optimal_trajectory = optimal

    planner.record_state_and_input(planner.x_0)
    optimal_traj = None
    n_cycles = 0
    while not planner.goal_reached() and n_cycles < max_cycles:
        count = len(planner.record_state_list) - 1
        if count % freq == 0:
            planner.set_desired_velocity(desired_velocity=desired_v if n_cycles == 0 else None,
                                         current_speed=planner.x_0.velocity)
            x0c = planner.x_0
            rec = {"x0": np.array([x0c.position[0], x0c.position[1], x0c.orientation, x0c.velocity, x0c.acceleration,
                                   x0c.yaw_rate, x0c.steering_angle, float(x0c.time_step)]),
                   "x0_lon": np.array(planner.x_0_cl[0], dtype=np.float64),
                   "x0_lat": np.array(planner.x_0_cl[1], dtype=np.float64)}
            del level_log[:]
            optimal_traj = planner.plan()
            rec["levels"] = [dict(lv) for lv in level_log]
            rec["ok"] = optimal_traj is not None
            rec["optimal_cost"] = float(planner.optimal_cost)
            if optimal_traj is not None:
                rec.update(plan_output_arrays(optimal_traj))
            records.append(rec)
            n_cycles += 1
            if not optimal_traj:
                break
            planner.record_state_and_input(optimal_traj[0].state_list[1])
            planner.reset(initial_state_cart=planner.record_state_list[-1],
                          initial_state_curv=(optimal_traj[2][1], optimal_traj[3][1]),
                          collision_checker=planner.collision_checker, coordinate_system=planner.coordinate_system)
        else:
            k = count % freq
            planner.record_state_and_input(optimal_traj[0].state_list[1 + k])
            planner.reset(initial_state_cart=planner.record_state_list[-1],
                          initial_state_curv=(optimal_traj[2][1 + k], optimal_traj[3][1 + k]),
                          collision_checker=planner.collision_checker, coordinate_system=planner.coordinate_system)

    co = planner.coordinate_system
    arrays = {"ref_path_raw": scn["ref_path"], "ref_pos": co.ref_pos, "ref_theta": co.ref_theta,
              "ref_curv": co.ref_curv, "ref_curv_d": co.ref_curv_d, "cc_path": co.ccosy.path,
              "cc_S": co.ccosy.pathlength, "cc_normals": co.ccosy.normals}
    arrays.update(golden_io.pack_obstacles(scn, "ob_"))
    meta = {"name": name, "N": planner.N, "dt": planner.dt, "t_min": samp_y.get("t_min", 0.4),
            "low_vel_mode_threshold": cfg.planning.low_vel_mode_threshold, "draw_traj_set": draw,
            "replanning_frequency": freq, "desired_velocity": desired_v, "n_cycles": len(records),
            "cycles": []}
    for ci, rec in enumerate(records):
        arrays["c%d_x0" % ci] = rec["x0"]
        arrays["c%d_x0_lon" % ci] = rec["x0_lon"]
        arrays["c%d_x0_lat" % ci] = rec["x0_lat"]
        cm = {"ok": bool(rec["ok"]), "levels": [], "optimal_cost": rec["optimal_cost"]}
        for k in ("out_cart", "out_curv", "out_lon", "out_lat"):
            if k in rec:
                arrays["c%d_%s" % (ci, k)] = rec[k]
        for li, lv in enumerate(rec["levels"]):
            key = "c%d_l%d_" % (ci, li)
            arrays[key + "kin_feasible"] = lv["kin_feasible"]
            arrays[key + "cost"] = lv["cost"]
            if "winner_states" in lv:
                arrays[key + "winner_states"] = lv["winner_states"]
            cm["levels"].append({"level": lv["level"], "n": lv["n"], "winner": lv["winner"],
                                 "n_inf_kin": lv["n_inf_kin"], "n_inf_col": lv["n_inf_col"], "reasons": lv["reasons"]})
        meta["cycles"].append(cm)
    arrays["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, "cyc_%s.npz" % name), **arrays)
    print("cyc_%s: %d cycles; winners %s" % (name, len(records),
                                             [[lv["winner"] for lv in r["levels"]] for r in records]))


def make_initial_states():
    """the reference's own ``_compute_initial_states`` (reactive_planner.py:446-512) on the Cartesian
    states of every recorded replanning cycle, in high- and low-velocity mode -> tests/golden/init_states.npz.
    (The x0_lon / x0_lat of the cyclic fixtures are NOT that: from cycle 1 on run_planner.py hands reset() the previous
    optimal trajectory's curvilinear states.)"""
    ref_root = ref_shims.REFERENCE_ROOT
    arrays = {}
    for name, kw in CYC_CASES.items():
        z = np.load(os.path.join(OUT, "cyc_%s.npz" % name))
        meta = json.loads(str(z["meta"]))
        sc = scenario_xml.load(os.path.join(ref_root, "example_scenarios", name + ".xml"),
                               boundary_method=kw.get("boundary_method", "obb_rectangles"))
        scn = dict(sc["scn"])
        scn["ref_path"] = _route_path(sc, kw["route"], kw.get("extend_back", 0.0))
        planner = H.build_planner(scn, N=meta["N"], dt=meta["dt"], t_min=meta["t_min"],
                                  low_vel_mode_threshold=meta["low_vel_mode_threshold"], draw_traj_set=False)
        from commonroad_rp.state import ReactivePlannerState        # importable once the shims are installed
        xs, out = [], {"lon_hv": [], "lat_hv": [], "lon_lv": [], "lat_lv": []}
        for ci in range(meta["n_cycles"]):
            x = z["c%d_x0" % ci]
            x0 = ReactivePlannerState(time_step=int(x[7]), position=np.array([x[0], x[1]]), orientation=x[2], velocity=x[3],
                                      acceleration=x[4], yaw_rate=x[5], steering_angle=x[6])
            xs.append(x)
            for tag, flag in (("hv", False), ("lv", True)):
                planner._low_vel_mode = flag
                lon, lat = planner._compute_initial_states(x0)
                out["lon_" + tag].append(np.array(lon, dtype=np.float64))
                out["lat_" + tag].append(np.array(lat, dtype=np.float64))
        arrays[name + "_x0"] = np.stack(xs)
        for k, v in out.items():
            arrays[name + "_" + k] = np.stack(v)
        print("init_states %s: %d states" % (name, len(xs)))
    np.savez_compressed(os.path.join(OUT, "init_states.npz"), **arrays)


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "init":
        make_initial_states()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "only":
        for name in sys.argv[2:]:
            make_synthetic(name, **SYN_CASES[name])
        return
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        for name in sys.argv[2:] or BIG_CASES:
            make_big(name, **BIG_CASES[name])
        return
    if len(sys.argv) > 1 and sys.argv[1] == "plan":
        for name in sys.argv[2:] or PLAN_CASES:
            make_plan(name, **PLAN_CASES[name])
        return
    if len(sys.argv) > 1 and sys.argv[1] == "cyc":
        for name in sys.argv[2:]:
            make_cyclic(name, **CYC_CASES[name])
        return
    for name, kw in SYN_CASES.items():
        make_synthetic(name, **kw)
    for name, kw in CYC_CASES.items():
        make_cyclic(name, **kw)
    make_initial_states()
    for name, kw in PLAN_CASES.items():
        make_plan(name, **kw)
    for name, kw in BIG_CASES.items():
        make_big(name, **kw)


if __name__ == "__main__":
    main()
