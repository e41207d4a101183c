"""TEST INFRASTRUCTURE ONLY -- runs the reference's own ``ReactivePlanner`` (imported
read-only from /root/reference through ``oracle/ref_shims.py``) on a scenario dict
(``commonroad_rp_b200.utility.synthetic`` format) and records everything the parity tests
compare: enumeration order, per-candidate coefficients / labels / costs / states, the
selected enumeration index and the per-cycle counters.

Only usable in this container (needs /root/reference); its outputs travel as fixtures under
tests/golden/ (see oracle/make_golden.py).
"""
import numpy as np

from . import ref_shims

STATE_FIELDS = ("x", "y", "theta", "v", "a", "kappa", "kappa_dot",
                "s", "d", "theta_cl", "s_dot", "s_ddot", "d_dot", "d_ddot")


def scenario_stub(scn):
    statics = [ref_shims.StaticObstacleStub(*row) for row in scn["static_boxes"]]
    dyns = [ref_shims.DynamicObstacleStub(t0, st, lw[0], lw[1])
            for t0, st, lw in zip(scn["dyn_t0"], scn["dyn_states"], scn["dyn_lw"])]
    return ref_shims.ScenarioStub(statics, dyns, scn["boundary_boxes"], scn.get("boundary_tris", ()))


def build_planner(scn, N=20, dt=0.1, t_min=0.4, low_vel_mode_threshold=4.0, longitudinal_mode="velocity_keeping",
                  draw_traj_set=False, constraints=None, factor=1, d_min=-3.0, d_max=3.0, smooth_reference=True,
                  continuous_collision_check=False):
    """Reference ReactivePlanner with multiproc off (canonical enumeration order)."""
    ref_shims.install()
    from commonroad_rp.reactive_planner import ReactivePlanner
    from commonroad_rp.utility.config import ReactivePlannerConfiguration
    from commonroad_rp.utility.utils_coordinate_system import CoordinateSystem

    cfg = ReactivePlannerConfiguration()
    cfg.planning.dt = dt
    cfg.planning.time_steps_computation = N
    cfg.planning.planning_horizon = dt * N
    cfg.planning.low_vel_mode_threshold = low_vel_mode_threshold
    cfg.planning.factor = factor
    cfg.planning.continuous_collision_check = bool(continuous_collision_check)
    if constraints is not None:
        cfg.planning.constraints_to_check = list(constraints)
    cfg.sampling.t_min = t_min
    cfg.sampling.d_min = d_min
    cfg.sampling.d_max = d_max
    cfg.sampling.longitudinal_mode = longitudinal_mode
    cfg.debug.multiproc = False
    cfg.debug.draw_traj_set = draw_traj_set
    cfg.debug.save_plots = draw_traj_set
    cfg.update(scenario=scenario_stub(scn), planning_problem=None)
    planner = ReactivePlanner(cfg)
    planner.set_reference_path(coordinate_system=CoordinateSystem(np.asarray(scn["ref_path"], dtype=np.float64),
                                                                  smooth_reference=smooth_reference))
    return planner


def set_initial_state(planner, x0_lon, x0_lat, orientation=None, velocity=None, time_step=0,
                      acceleration=0.0, yaw_rate=0.0, steering_angle=0.0):
    """x_0 (Cartesian, rear axle) consistent with the given curvilinear state."""
    from commonroad_rp.state import ReactivePlannerState
    co = planner.coordinate_system
    pos = co.convert_to_cartesian_coords(x0_lon[0], x0_lat[0])
    if orientation is None:
        j = int(np.argmax(co.ref_pos > x0_lon[0])) - 1
        orientation = float(co.ref_theta[j])
    if velocity is None:
        velocity = float(x0_lon[1])
    x0 = ReactivePlannerState(time_step=int(time_step), position=np.asarray(pos), steering_angle=steering_angle,
                              velocity=velocity, orientation=orientation, acceleration=acceleration,
                              yaw_rate=yaw_rate)
    planner.reset(initial_state_cart=x0, initial_state_curv=(list(x0_lon), list(x0_lat)),
                  collision_checker=planner.collision_checker, coordinate_system=planner.coordinate_system)
    return x0


def override_sample_sets(planner, level, t=None, v=None, d=None, s=None):
    """Replace the level's sample sets by explicit values (dense synthetic sweeps).  The sets are
    built with the same ``set(ndarray)`` expression the reference uses (sampling.py:83,98,116)."""
    sp = planner.sampling_space
    if t is not None:
        sp.samples_t._dict_level_to_sample_set[level] = set(np.asarray(t, dtype=np.float64))
    if v is not None:
        sp.samples_v._dict_level_to_sample_set[level] = set(np.asarray(v, dtype=np.float64))
    if d is not None:
        sp.samples_d._dict_level_to_sample_set[level] = set(np.asarray(d, dtype=np.float64))
    if s is not None:
        sp.samples_s._dict_level_to_sample_set[level] = set(np.asarray(s, dtype=np.float64))


def evaluate_level(planner, level, want_states=True, state_sample=None):
    """One pass of reactive_planner.py:620-624 for ``level`` with per-candidate bookkeeping.
    state_sample = (count, seed): keep the state blocks of a seeded sample of feasible candidates instead of all."""
    x0_lon, x0_lat = planner.x_0_cl
    planner._low_vel_mode = bool(planner.x_0.velocity < planner.config.planning.low_vel_mode_threshold)
    bundle = planner._create_trajectory_bundle(x0_lon, x0_lat, samp_level=level)
    cands = list(bundle.trajectories)
    if planner.config.sampling.longitudinal_mode == "stopping":
        # filter_goals_behind drops candidates before the kinematic check (trajectories.py:545-550)
        kept = [c.trajectory_long.x_0[0] < c.trajectory_long.x_d[0] for c in cands]
    else:
        kept = [True] * len(cands)
    winner = planner._get_optimal_trajectory(bundle)
    feasible_ids = {id(t) for t in bundle.trajectories}

    n = len(cands)
    Np1 = planner.N + 1
    out = {
        "n": n,
        "level": level,
        "coeffs_lon": np.array([c.trajectory_long.coeffs for c in cands]).reshape(n, 6),
        "coeffs_lat": np.array([c.trajectory_lat.coeffs for c in cands]).reshape(n, 6),
        "delta_tau_lon": np.array([c.trajectory_long.delta_tau for c in cands], dtype=np.float64),
        "delta_tau_lat": np.array([c.trajectory_lat.delta_tau for c in cands], dtype=np.float64),
        "kept": np.array(kept, dtype=bool),
        "kin_feasible": np.array([id(c) in feasible_ids for c in cands], dtype=bool),
        "cost": np.array([float(c.cost) if id(c) in feasible_ids else np.nan for c in cands]),
        "label": np.array([("none" if c.feasibility_label is None else c.feasibility_label.value) for c in cands]),
        "winner": -1 if winner is None else next(i for i, c in enumerate(cands) if c is winner),
        "n_infeasible_kinematics": int(planner.infeasible_count_kinematics),
        "n_infeasible_collision": int(planner.infeasible_count_collision),
        "reasons": dict(planner.infeasible_reason_dict),
    }
    def state_block(c):
        blk = np.full((len(STATE_FIELDS), Np1), np.nan)
        if c.cartesian is not None and c.curvilinear is not None:
            ca, cu = c.cartesian, c.curvilinear
            rows = (ca.x, ca.y, ca.theta, ca.v, ca.a, ca.kappa, ca.kappa_dot,
                    cu.s, cu.d, cu.theta, cu.s_dot, cu.s_ddot, cu.d_dot, cu.d_ddot)
            for f, arr in enumerate(rows):
                blk[f, :] = arr
        return blk

    if state_sample is not None:
        # large bundles: the state blocks of a seeded sample of the feasible candidates (+ the winner) only
        feas = np.nonzero(out["kin_feasible"])[0]
        rng = np.random.default_rng(state_sample[1])
        pick = feas if len(feas) <= state_sample[0] else np.sort(rng.choice(feas, state_sample[0], replace=False))
        if out["winner"] >= 0 and out["winner"] not in pick:
            pick = np.sort(np.append(pick, out["winner"]))
        out["state_idx"] = pick.astype(np.int64)
        out["states_sampled"] = np.stack([state_block(cands[i]) for i in pick]) if len(pick) else np.zeros((0, len(STATE_FIELDS), Np1))
    elif want_states:
        out["states"] = np.stack([state_block(c) for c in cands]) if n else np.zeros((0, len(STATE_FIELDS), Np1))
    return out


def problem_from_planner(planner, level, scn):
    """Pack the reference planner's current cycle into the plain-array problem dict that
    oracle/rp_oracle.py and the GPU C-ABI consume.  Enumeration order is taken from the
    reference's own set iteration (sampling.py:218,220,226)."""
    from . import rp_oracle as O
    sp = planner.sampling_space
    cfg = planner.config
    x0_lon, x0_lat = planner.x_0_cl
    mode = cfg.sampling.longitudinal_mode
    t_list = [float(t) for t in sp.samples_t.samples_at_level(level)]
    lon_set = sp.samples_v.samples_at_level(level) if mode == "velocity_keeping" else sp.samples_s.samples_at_level(level)
    lon_list = [float(v) for v in lon_set]
    d_list = [float(d) for d in sp.samples_d.samples_at_level(level).union({x0_lat[0]})]
    co = planner.coordinate_system
    cc = co.ccosy
    cf = planner.cost_function
    kind = "failsafe" if type(cf).__name__ == "DefaultCostFunctionFailSafe" else "default"
    veh = planner.vehicle_params
    return {
        "t": np.array(t_list), "lon": np.array(lon_list), "d": np.array(d_list),
        "x0_lon": np.array(x0_lon, dtype=np.float64), "x0_lat": np.array(x0_lat, dtype=np.float64),
        "x0_orientation": float(planner.x_0.orientation), "x0_time_step": int(planner.x_0.time_step),
        "lon_mode": mode,
        "low_vel_mode": bool(planner.x_0.velocity < cfg.planning.low_vel_mode_threshold),
        "dt": planner.dt, "N": planner.N, "factor": cfg.planning.factor,
        "draw_all": bool(planner._draw_traj_set),
        "continuous": bool(cfg.planning.continuous_collision_check),
        "constraints": tuple(cfg.planning.constraints_to_check),
        "cost": {"kind": kind, "desired_speed": getattr(cf, "desired_speed", None),
                 "desired_s": getattr(cf, "desired_s", None), "desired_d": getattr(cf, "desired_d", 0.0),
                 "w_a": getattr(cf, "w_a", 1)},
        "vehicle": {"length": veh.length, "width": veh.width, "wb_rear_axle": veh.wb_rear_axle,
                    "wheelbase": veh.wheelbase, "a_max": veh.a_max, "v_switch": veh.v_switch,
                    "delta_max": veh.delta_max, "v_delta_max": veh.v_delta_max},
        "ref": {"ref_pos": co.ref_pos, "ref_theta": co.ref_theta, "ref_curv": co.ref_curv,
                "ref_curv_d": co.ref_curv_d},
        "ccosy": {"path": cc.path, "S": cc.pathlength, "normals": cc.normals,
                  "limit": cc.projection_domain_limit},
        "obstacles": {k: scn[k] for k in ("static_boxes", "dyn_t0", "dyn_states", "dyn_lw", "boundary_boxes",
                                          "boundary_tris") if k in scn},
    }
