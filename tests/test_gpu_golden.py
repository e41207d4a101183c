"""GPU parity against the fixtures generated from the REFERENCE's own code (tests/golden/, made by
oracle/make_golden.py): single bundles through the C-ABI, and cyclic replanning of the reference's three
bundled scenarios through the drop-in ``ReactivePlanner`` API."""
import glob
import json
import os

import numpy as np
import pytest

from tests import golden_io
from tests import helpers as H

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SYN = sorted(glob.glob(os.path.join(GOLDEN, "syn_*.npz")))
CYC = sorted(glob.glob(os.path.join(GOLDEN, "cyc_*.npz")))
RTOL = 1e-9      # BASELINE.json: states and costs within 1e-9 relative


@pytest.mark.parametrize("path", SYN, ids=[os.path.basename(p)[4:-4] for p in SYN])
def test_bundle_matches_reference_fixture(path):
    from commonroad_rp_b200._lib import REASON_NAMES
    z = np.load(path)
    prob = golden_io.unpack_problem(z)
    eng = H.engine_for(prob)
    g = H.run_engine_grid(eng, prob, want_all_states=True)
    n = len(z["r_cost"])
    assert g["n"] == n
    assert H.rel_err(z["r_coeffs_lon"], g["coeffs_lon"]) < RTOL and H.rel_err(z["r_coeffs_lat"], g["coeffs_lat"]) < RTOL
    kept = z["r_kept"]
    assert np.array_equal(g["status"] == 3, ~kept)                                   # filter_goals_behind
    kin_ok = (g["status"] == 0) | (g["status"] == 2)
    assert np.array_equal(kin_ok, z["r_kin_feasible"])                               # flags: bit exact
    assert H.rel_err(z["r_cost"][kin_ok], g["cost"][kin_ok]) < RTOL
    assert g["winner"] == int(z["r_winner"])                                         # selected index: exact
    assert g["n_infeasible_kinematics"] == int(z["r_n_inf_kin"])
    assert g["n_infeasible_collision"] == int(z["r_n_inf_col"])
    for name, cnt in json.loads(str(z["r_reasons"])).items():
        assert g["reason_counts"][REASON_NAMES.index(name)] == cnt, name
    lab = z["r_label"]
    assert np.all(g["status"][lab == 3] == 2)          # every collider the lazy reference pass met collides here
    idx = z["r_state_idx"]
    assert H.rel_err(z["r_states"], g["states"][idx]) < RTOL
    eng.close()


@pytest.mark.parametrize("path", SYN, ids=[os.path.basename(p)[4:-4] for p in SYN])
def test_bundle_matches_reference_fixture_candidate_major(path):
    """the same fixtures through the candidate-major kernel (select-only mode: verdicts, costs, winner, counters)"""
    from commonroad_rp_b200 import _lib
    from commonroad_rp_b200._lib import REASON_NAMES
    z = np.load(path)
    prob = golden_io.unpack_problem(z)
    eng = H.engine_for(prob)
    g = H.run_engine_grid(eng, prob, want_all_states=False, kernel=_lib.KERNEL_CANDIDATE_MAJOR)
    assert g["n"] == len(z["r_cost"])
    assert np.array_equal(g["status"] == 3, ~z["r_kept"])
    kin_ok = (g["status"] == 0) | (g["status"] == 2)
    assert np.array_equal(kin_ok, z["r_kin_feasible"])
    assert H.rel_err(z["r_cost"][kin_ok], g["cost"][kin_ok]) < RTOL
    assert g["winner"] == int(z["r_winner"])
    assert g["n_infeasible_kinematics"] == int(z["r_n_inf_kin"])
    assert g["n_infeasible_collision"] == int(z["r_n_inf_col"])
    for name, cnt in json.loads(str(z["r_reasons"])).items():
        assert g["reason_counts"][REASON_NAMES.index(name)] == cnt, name
    assert np.all(g["status"][z["r_label"] == 3] == 2)
    if g["winner"] >= 0:
        idx = list(z["r_state_idx"])
        if g["winner"] in idx:
            assert H.rel_err(z["r_states"][idx.index(g["winner"])], g["winner_states"]) < RTOL
    eng.close()


@pytest.mark.parametrize("path", CYC, ids=[os.path.basename(p)[4:-4] for p in CYC])
def test_cyclic_replanning_matches_reference_fixture(path):
    from commonroad_rp_b200 import collision
    from commonroad_rp_b200._lib import REASON_NAMES
    from commonroad_rp_b200.reactive_planner import ReactivePlanner
    from commonroad_rp_b200.state import ReactivePlannerState
    from commonroad_rp_b200.utility.config import ReactivePlannerConfiguration
    from commonroad_rp_b200.utility.utils_coordinate_system import CoordinateSystem
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    cfg = ReactivePlannerConfiguration()
    cfg.planning.time_steps_computation = meta["N"]
    cfg.planning.dt = meta["dt"]
    cfg.planning.planning_horizon = meta["dt"] * meta["N"]
    cfg.planning.low_vel_mode_threshold = meta["low_vel_mode_threshold"]
    cfg.sampling.t_min = meta["t_min"]
    cfg.debug.draw_traj_set = meta["draw_traj_set"]
    cfg.debug.save_plots = meta["draw_traj_set"]
    cc = collision.checker_from_arrays(**golden_io.unpack_obstacles(z, "ob_"))

    class _EmptyScenario:
        static_obstacles, dynamic_obstacles = (), ()
        lanelet_network = type("LN", (), {"lanelets": ()})()

    cfg.update(scenario=_EmptyScenario(), planning_problem=None)
    planner = ReactivePlanner(cfg)
    co = CoordinateSystem(z["ref_path_raw"])
    # the product's own table construction agrees with the reference-side tables of the fixture
    for mine, ref in ((co.ref_pos, z["ref_pos"]), (co.ref_theta, z["ref_theta"]), (co.ref_curv, z["ref_curv"]),
                      (co.ref_curv_d, z["ref_curv_d"]), (co.ccosy.path, z["cc_path"]), (co.ccosy.normals, z["cc_normals"])):
        assert np.allclose(mine, ref, rtol=1e-9, atol=1e-9)
    planner.set_reference_path(coordinate_system=co)

    levels_seen = []
    orig = planner._get_optimal_trajectory

    def spy(bundle):
        win = orig(bundle)
        cost, status, reason, step = planner.engine.fetch_candidates()
        levels_seen.append({"res": planner.last_result, "cost": cost, "status": status,
                            "counts": (planner.infeasible_count_kinematics, planner.infeasible_count_collision),
                            "reasons": dict(planner.infeasible_reason_dict)})
        return win

    planner._get_optimal_trajectory = spy
    for ci, cm in enumerate(meta["cycles"]):
        x = z["c%d_x0" % ci]
        x0 = ReactivePlannerState(time_step=int(x[7]), position=np.array([x[0], x[1]]), orientation=x[2], velocity=x[3],
                                  acceleration=x[4], yaw_rate=x[5], steering_angle=x[6])
        planner.reset(initial_state_cart=x0, initial_state_curv=(list(z["c%d_x0_lon" % ci]), list(z["c%d_x0_lat" % ci])),
                      collision_checker=cc, coordinate_system=co)
        if ci == 0:
            # Cartesian -> curvilinear initial state (reference :446-512) from the product's own frame
            planner._low_vel_mode = bool(x0.velocity < cfg.planning.low_vel_mode_threshold)
            lon, lat = planner._compute_initial_states(x0)
            assert np.allclose(lon, z["c0_x0_lon"], rtol=1e-8, atol=1e-8) and np.allclose(lat, z["c0_x0_lat"], rtol=1e-8, atol=1e-8)
        planner.set_desired_velocity(desired_velocity=meta["desired_velocity"] if ci == 0 else None,
                                     current_speed=x0.velocity)
        del levels_seen[:]
        out = planner.plan()
        assert (out is not None) == cm["ok"], "cycle %d" % ci
        assert len(levels_seen) == len(cm["levels"]), "cycle %d: level escalation differs" % ci
        for li, (seen, lv) in enumerate(zip(levels_seen, cm["levels"])):
            tag = "cycle %d level %d" % (ci, lv["level"])
            key = "c%d_l%d_" % (ci, li)
            assert seen["res"].n_candidates == lv["n"], tag
            kin_ok = np.isin(seen["status"], (0, 2, 4))       # 4: feasible, not visited by the lazy collision pass
            assert np.array_equal(kin_ok, z[key + "kin_feasible"]), tag
            assert seen["res"].winner == lv["winner"], tag
            assert seen["counts"] == (lv["n_inf_kin"], lv["n_inf_col"]), tag
            assert seen["reasons"] == lv["reasons"], tag
            assert H.rel_err(z[key + "cost"][kin_ok], seen["cost"][kin_ok]) < RTOL, tag
        if out is not None:
            ws = z["c%d_l%d_winner_states" % (ci, len(cm["levels"]) - 1)]
            cart, curv, lon_list, lat_list = out
            got = np.array([[s.position[0], s.position[1], s.velocity, s.acceleration] for s in cart.state_list]).T
            assert H.rel_err(ws[[0, 1, 3, 4]], got) < RTOL, "cycle %d" % ci
            assert H.rel_err(ws[[7, 10, 11]], np.array(lon_list).T) < RTOL and H.rel_err(ws[[8, 12, 13]], np.array(lat_list).T) < RTOL
            # the complete return value (reactive_planner.py:514-568): orientation shift, yaw rate, steering angle, time
            # steps, the curvilinear state list
            got_out = H.plan_output_arrays(out)
            for key in ("out_cart", "out_curv", "out_lon", "out_lat"):
                want = z["c%d_%s" % (ci, key)]
                assert got_out[key].shape == want.shape and H.rel_err(want, got_out[key]) < RTOL, ("cycle %d" % ci, key)


def _frame_engine(z):
    from commonroad_rp_b200 import _lib
    from commonroad_rp_b200.utility.config import VehicleConfiguration
    from commonroad_rp_b200.utility.utils_coordinate_system import CoordinateSystem
    veh = VehicleConfiguration()
    co = CoordinateSystem(z["ref_path_raw"])
    eng = _lib.Engine(0)
    eng.set_vehicle(veh.length, veh.width, veh.wb_rear_axle, veh.wheelbase, veh.a_max, veh.v_switch, veh.delta_max,
                    veh.v_delta_max, veh.kappa_max)
    tb = co.device_tables()
    eng.set_reference(tb["ref_pos"], tb["ref_theta"], tb["ref_curv"], tb["ref_curv_d"], tb["path_xy"], tb["path_s"],
                      tb["path_normals"], tb["proj_limit"])
    return eng, co


@pytest.mark.parametrize("path", CYC, ids=[os.path.basename(p)[4:-4] for p in CYC])
def test_initial_states_match_reference_fixture(path):
    """the device's Cartesian -> curvilinear initial states (rp_initial_states) against the
    reference's own _compute_initial_states (:446-512) evaluated on the Cartesian state of every recorded replanning
    cycle, in high- and low-velocity mode (tests/golden/init_states.npz, oracle/make_golden.py init)"""
    z = np.load(path)
    g = np.load(os.path.join(GOLDEN, "init_states.npz"))
    name = json.loads(str(z["meta"]))["name"]
    eng, co = _frame_engine(z)
    x = g[name + "_x0"]
    n = len(x)
    x0 = x[:, [0, 1, 2, 3, 4, 6]]                               # x, y, orientation, velocity, acceleration, steering_angle
    for tag, flag in (("hv", 0), ("lv", 1)):
        lon, lat, status = eng.initial_states(x0, np.full(n, flag, dtype=np.int32))
        assert not status.any()
        assert H.rel_err(g[name + "_lon_" + tag], lon) < RTOL, (tag, H.rel_err(g[name + "_lon_" + tag], lon))
        assert H.rel_err(g[name + "_lat_" + tag], lat) < RTOL, (tag, H.rel_err(g[name + "_lat_" + tag], lat))
    # the host-side frame (CoordinateSystem.convert_to_curvilinear_coords) agrees, and error cases map to status codes
    s_d = co.convert_to_curvilinear_coords(x0[0, 0], x0[0, 1])
    assert H.rel_err(s_d, [lon[0, 0], lat[0, 0]]) < RTOL
    far = x0[:1].copy()
    far[0, :2] += 1.0e4
    assert eng.initial_states(far, [0])[2][0] == 1                   # outside the projection domain
    fastest = int(np.argmax(x0[:, 3]))                               # (ZAM-Ramp starts at v = 0: s_dot = -0 is not negative)
    back = x0[fastest:fastest + 1].copy()
    back[0, 2] += np.pi
    assert eng.initial_states(back, [0])[2][0] == 2                  # driving against the reference: negative s_dot
    eng.close()


def test_batched_initial_states_over_scenarios():
    """one state per scenario, each against its own reference tables, in one launch (rp_batch_initial_states)"""
    from commonroad_rp_b200 import _lib
    g = np.load(os.path.join(GOLDEN, "init_states.npz"))
    zs = [np.load(p) for p in CYC]
    names = [json.loads(str(z["meta"]))["name"] for z in zs]
    pairs = [_frame_engine(z) for z in zs]
    batch = _lib.Batch([e for e, _ in pairs])
    x0 = np.stack([g[nm + "_x0"][2][[0, 1, 2, 3, 4, 6]] for nm in names])
    lon, lat, status = batch.initial_states(x0, 0)
    assert not status.any()
    for k, nm in enumerate(names):
        assert H.rel_err(g[nm + "_lon_hv"][2], lon[k]) < RTOL and H.rel_err(g[nm + "_lat_hv"][2], lat[k]) < RTOL
        single = pairs[k][0].initial_states(x0[k], [0])
        assert np.array_equal(single[0][0], lon[k]) and np.array_equal(single[1][0], lat[k])
    batch.close()
    for e, _ in pairs:
        e.close()


# ---- round 2: large bundles, plan()-level fixtures, continuous check through the planner ---------------------------
BIG = sorted(glob.glob(os.path.join(GOLDEN, "big_*.npz")))
PLAN = sorted(glob.glob(os.path.join(GOLDEN, "plan_*.npz")))


def _check_big(z, g, tag):
    from commonroad_rp_b200._lib import REASON_NAMES
    n = len(z["r_cost"])
    assert g["n"] == n, tag
    kin_ok = (g["status"] == 0) | (g["status"] == 2) | (g["status"] == 4)          # (4: the lazy pass never visited it)
    assert np.array_equal(kin_ok, z["r_kin_feasible"]), tag                         # flags: bit exact, every candidate
    assert H.rel_err(z["r_cost"][kin_ok], g["cost"][kin_ok]) < RTOL, tag             # every cost
    assert g["winner"] == int(z["r_winner"]), tag
    assert g["n_infeasible_kinematics"] == int(z["r_n_inf_kin"]), tag
    assert g["n_infeasible_collision"] == int(z["r_n_inf_col"]), tag
    for name, cnt in json.loads(str(z["r_reasons"])).items():
        assert g["reason_counts"][REASON_NAMES.index(name)] == cnt, (tag, name)
    lab = z["r_label"]
    assert np.all(g["status"][lab == 3] == 2), tag         # every collider the reference's lazy pass met collides here
    # ... and nothing ranked before the winner is collision-free here that the reference saw colliding, or vice versa
    wc, wi = g["winner_cost"], g["winner"]
    before = kin_ok & ((g["cost"] < wc) | ((g["cost"] == wc) & (np.arange(n) < wi)))
    assert np.array_equal(before & (g["status"] == 2), lab == 3), tag


@pytest.mark.parametrize("path", BIG, ids=[os.path.basename(p)[4:-4] for p in BIG])
def test_big_bundle_lazy_collision_pass_matches_reference_fixture(path):
    """the same bundles with check_collision = 2 through the candidate-major schedule: the march stores the ego boxes and
    the deferred checker (rp_cand.cuh) gives a verdict to everything that can be ranked before the winner -- winner,
    counters and the colliders the reference's own lazy pass met are the reference's"""
    from commonroad_rp_b200 import _lib
    z = np.load(path)
    prob = golden_io.unpack_problem(z)
    eng = H.engine_for(prob)
    for rep in range(2):                                   # (second cycle: the checker's masks / lists were left clean)
        g = H.run_engine_grid(eng, prob, want_all_states=False, kernel=_lib.KERNEL_CANDIDATE_MAJOR,
                              check_collision=_lib.COLLISION_LAZY)
        assert eng.last_main_kernel() == _lib.KERNEL_CANDIDATE_MAJOR and eng.launches_per_plan() == 8
        _check_big(z, g, os.path.basename(path))
        wc, wi = g["winner_cost"], g["winner"]
        n = g["n"]
        before = (g["cost"] < wc) | ((g["cost"] == wc) & (np.arange(n) <= wi))
        assert not (before & (g["status"] == 4)).any()      # nothing ranked before the winner is left unchecked
        assert (g["status"] == 4).sum() > 0.5 * n           # ... and most of the bundle never needed a collision check
    q = int(np.nonzero(z["r_state_idx"] == g["winner"])[0][0])
    assert H.rel_err(z["r_states"][q], eng.fetch_states(g["winner"])) < RTOL
    eng.close()


@pytest.mark.parametrize("path", BIG, ids=[os.path.basename(p)[4:-4] for p in BIG])
@pytest.mark.parametrize("kernel", ["candidate_major", "step_parallel"])
def test_big_bundle_matches_reference_fixture(path, kernel):
    """BASELINE configs[3] (131 072 x 61, the bench workload itself) and configs[4]'s per-scenario bundle (8 874 x 61)
    against the verdicts of the reference's own code on the same inputs, through both schedules"""
    from commonroad_rp_b200 import _lib
    z = np.load(path)
    prob = golden_io.unpack_problem(z)
    eng = H.engine_for(prob)
    pol = _lib.KERNEL_CANDIDATE_MAJOR if kernel == "candidate_major" else _lib.KERNEL_STEP_PARALLEL
    g = H.run_engine_grid(eng, prob, want_all_states=False, kernel=pol)
    assert eng.last_main_kernel() == pol
    _check_big(z, g, os.path.basename(path))
    for q, k in enumerate(z["r_state_idx"]):                                         # sampled state blocks + the winner
        assert H.rel_err(z["r_states"][q], eng.fetch_states(int(k))) < RTOL, (path, int(k))
    eng.close()


@pytest.mark.parametrize("collision_mode", [1, 2], ids=["all_flags", "lazy_pass"])
def test_scenario_batch_matches_reference_fixtures(collision_mode):
    """the batch path (rp_batch_*: one launch chain for all scenarios) on the reference-generated configs[4] bundles, with a
    collision flag for every candidate and with the reference's lazy pass (deferred checker over the shared tile lists)"""
    from commonroad_rp_b200 import _lib
    paths = [p for p in BIG if "batch_" in os.path.basename(p)]
    assert len(paths) >= 2
    zs = [np.load(p) for p in paths]
    probs = [golden_io.unpack_problem(z) for z in zs]
    engines = [H.engine_for(pr) for pr in probs]
    batch = _lib.Batch(engines)
    for rep in range(2):
        for k, pr in enumerate(probs):
            batch.set_inputs(k, H.inputs_for(pr, check_collision=collision_mode), pr["t"], pr["lon"], pr["d"])
        batch.launch()
        res = batch.results()
        for k, (z, r) in enumerate(zip(zs, res)):
            cost, status, reason, step = batch.fetch_candidates(k)
            g = {"n": r.n_candidates, "winner": r.winner, "winner_cost": r.winner_cost, "cost": cost, "status": status,
                 "n_infeasible_kinematics": r.n_infeasible_kinematics, "n_infeasible_collision": r.n_infeasible_collision,
                 "reason_counts": list(r.reason_counts)}
            _check_big(z, g, "batch %d rep %d" % (k, rep))
    batch.close()
    for e in engines:
        e.close()


@pytest.mark.parametrize("path", PLAN, ids=[os.path.basename(p)[5:-4] for p in PLAN])
def test_plan_output_matches_reference_fixture(path):
    """the complete return value of ReactivePlanner.plan() (reactive_planner.py:570-665) on synthetic cycles, incl. the
    standstill branch (:638-653, :667-713), total failure (None) and the level escalation the reference went through"""
    z = np.load(path)
    meta = json.loads(str(z["plan_meta"]))
    prob = golden_io.unpack_problem(z)
    planner = H.planner_from_fixture(prob, z["ref_path_raw"], z["x0"], desired_velocity=meta["desired_velocity"])
    seen = []
    orig = planner._get_optimal_trajectory

    def spy(bundle):
        win = orig(bundle)
        seen.append({"n": planner.last_result.n_candidates, "found": win is not None,
                     "n_inf_kin": int(planner.infeasible_count_kinematics), "n_inf_col": int(planner.infeasible_count_collision),
                     "reasons": dict(planner.infeasible_reason_dict)})
        return win

    planner._get_optimal_trajectory = spy
    out = planner.plan()
    assert (out is not None) == meta["ok"]
    assert seen == meta["levels"]
    assert planner.optimal_cost == pytest.approx(meta["optimal_cost"], rel=1e-9, abs=1e-12)
    if out is not None:
        got = H.plan_output_arrays(out)
        for key in ("out_cart", "out_curv", "out_lon", "out_lat"):
            assert got[key].shape == z[key].shape, key                               # standstill: N states, not N + 1
            assert H.rel_err(z[key], got[key]) < RTOL, (key, H.rel_err(z[key], got[key]))
        assert meta["standstill"] == bool(np.all(got["out_cart"][3] == 0.0))


@pytest.mark.parametrize("name", ["continuous_pass", "continuous_hit"])
def test_continuous_collision_check_through_the_planner(name):
    """SURVEY 8f rank 2 through the drop-in API: ReactivePlanner(config.planning.continuous_collision_check=True) builds
    its checker from the scenario with the dynamic obstacles replaced by OBB-sum hulls (reference :239-243) and hull-checks
    the selected candidate (:1049-1058) -- same winner / counters as the reference's own run of the same cycle"""
    z = np.load(os.path.join(GOLDEN, "syn_%s.npz" % name))
    prob = golden_io.unpack_problem(z)
    assert prob["continuous"]
    raw = synthetic_raw_path(prob)
    r, c = prob["ref"], prob["ccosy"]
    j = int(np.argmax(r["ref_pos"] > prob["x0_lon"][0])) - 1
    from commonroad_rp_b200.utility.utils_coordinate_system import CoordinateSystem
    co = CoordinateSystem(raw)
    pos = co.convert_to_cartesian_coords(prob["x0_lon"][0], prob["x0_lat"][0])
    x0 = [pos[0], pos[1], prob["x0_orientation"], prob["x0_lon"][1], 0.0, 0.0, 0.0, prob["x0_time_step"]]
    planner = H.planner_from_fixture(prob, raw, x0, continuous=True)
    assert np.allclose(planner.coordinate_system.ref_pos, r["ref_pos"], rtol=1e-9, atol=1e-9)
    level = {120: 1, 540: 2, 630: 2}.get(len(z["r_cost"]), 2)
    out = planner.plan(current_sampling_level=level)
    res = planner.last_result
    assert res.n_candidates == len(z["r_cost"])
    assert res.winner == int(z["r_winner"]) and (out is not None) == (int(z["r_winner"]) >= 0)
    assert planner.infeasible_count_kinematics == int(z["r_n_inf_kin"])
    assert planner.infeasible_count_collision == int(z["r_n_inf_col"])


def synthetic_raw_path(prob):
    """the raw polyline of the synthetic fixtures (straight variant, amplitude 0): x = 0..299, y = 0"""
    from commonroad_rp_b200.utility import synthetic
    return synthetic.sine_path(0.0, 40.0, 300)
