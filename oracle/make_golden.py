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

OUT = os.path.join(ROOT, "tests", "golden")
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
                   factor=1, crossing=None):
    scn = synthetic.make_scenario(seed=seed, amplitude=amplitude, static_offset=static_offset)
    if crossing is not None:
        scn = synthetic.add_crossing_obstacle(scn, speed=crossing[0], phase=crossing[1])
    p = H.build_planner(scn, N=N, longitudinal_mode=mode, draw_traj_set=draw, low_vel_mode_threshold=low_vel_threshold,
                        continuous_collision_check=continuous, factor=factor)
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


# ---- bundled scenarios, cyclic replanning (run_planner.py:55-107) -----------------------------------
CYC_CASES = {
    "ZAM_Over-1_1": dict(route=[1000], yaml="ZAM_Over-1_1.yaml", max_cycles=14),
    "DEU_Test-1_1_T-1": dict(route=[1, 3], yaml="DEU_Test-1_1_T-1.yaml", max_cycles=12),
    "ZAM_Tjunction-1_42_T-1": dict(route=[50195, 50209, 50203], yaml="ZAM_Tjunction-1_42_T-1.yaml", max_cycles=14),
}


def _yaml(path):
    import yaml
    with open(path) as f:
        return yaml.safe_load(f)


def make_cyclic(name, route, yaml, max_cycles):
    from types import SimpleNamespace
    ref_root = ref_shims.REFERENCE_ROOT
    sc = scenario_xml.load(os.path.join(ref_root, "example_scenarios", name + ".xml"))
    cfg_y = _yaml(os.path.join(ref_root, "configurations", yaml))
    scn = dict(sc["scn"])
    scn["ref_path"] = scenario_xml.route_centerline(sc["lanelets"], route)
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
        cm = {"ok": bool(rec["ok"]), "levels": []}
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
        sc = scenario_xml.load(os.path.join(ref_root, "example_scenarios", name + ".xml"))
        scn = dict(sc["scn"])
        scn["ref_path"] = scenario_xml.route_centerline(sc["lanelets"], kw["route"])
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
    for name, kw in SYN_CASES.items():
        make_synthetic(name, **kw)
    for name, kw in CYC_CASES.items():
        make_cyclic(name, **kw)
    make_initial_states()


if __name__ == "__main__":
    main()
