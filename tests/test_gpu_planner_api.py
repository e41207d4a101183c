"""GPU tests of the public Python surface: lazy TrajectorySample views, the generic plug-in paths (custom
cost function, list-form sampling space), draw mode, stand-alone solves / collision queries, error
behaviour, and size-independent properties on the full BASELINE-size bundle."""
import copy

import numpy as np
import pytest

from commonroad_rp_b200.utility import synthetic
from oracle import rp_oracle as O
from oracle import third_party as tp
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _planner(seed=0, N=20, draw=False, s_dot0=15.0, d0=0.3, mode="velocity_keeping"):
    from commonroad_rp_b200 import collision
    from commonroad_rp_b200.reactive_planner import ReactivePlanner
    from commonroad_rp_b200.state import ReactivePlannerState
    from commonroad_rp_b200.utility.config import ReactivePlannerConfiguration
    from commonroad_rp_b200.utility.utils_coordinate_system import CoordinateSystem
    scn = synthetic.make_scenario(seed=seed)
    cfg = ReactivePlannerConfiguration()
    cfg.planning.time_steps_computation = N
    cfg.sampling.longitudinal_mode = mode
    cfg.debug.draw_traj_set = draw
    cfg.debug.save_plots = draw

    class _Empty:
        static_obstacles, dynamic_obstacles = (), ()
        lanelet_network = type("LN", (), {"lanelets": ()})()

    cfg.update(scenario=_Empty(), planning_problem=None)
    p = ReactivePlanner(cfg)
    co = CoordinateSystem(scn["ref_path"])
    cc = collision.checker_from_arrays(**{k: scn[k] for k in ("static_boxes", "dyn_t0", "dyn_states", "dyn_lw",
                                                              "boundary_boxes", "boundary_tris")})
    s0 = float(co.ref_pos[10])
    j = int(np.argmax(co.ref_pos > s0)) - 1
    pos = co.convert_to_cartesian_coords(s0, d0)
    x0 = ReactivePlannerState(time_step=0, position=pos, orientation=float(co.ref_theta[j]), velocity=s_dot0,
                              acceleration=0.0, yaw_rate=0.0, steering_angle=0.0)
    p.reset(initial_state_cart=x0, initial_state_curv=([s0, s_dot0, 0.0], [d0, 0.0, 0.0]), collision_checker=cc,
            coordinate_system=co)
    p.set_desired_velocity(desired_velocity=s_dot0, current_speed=s_dot0)
    return p, scn


def test_plan_returns_reference_shaped_tuple():
    p, _ = _planner()
    out = p.plan()
    assert out is not None and len(out) == 4
    cart, curv, lon_list, lat_list = out
    assert len(cart.state_list) == p.N + 1 == len(curv.state_list) == len(lon_list) == len(lat_list)
    st = cart.state_list[3]
    assert st.time_step == 3 and st.position.shape == (2,)
    assert st.yaw_rate == pytest.approx((cart.state_list[3].orientation - cart.state_list[2].orientation) / p.dt)
    assert len(p.planning_times) == 1 and p.infeasible_count_kinematics > 0
    assert set(p.infeasible_reason_dict) == set(p.config.planning.constraints_to_check)
    # re-planning from the planned state works like run_planner.py:84-86
    p.reset(initial_state_cart=cart.state_list[1], initial_state_curv=(lon_list[1], lat_list[1]),
            collision_checker=p.collision_checker, coordinate_system=p.coordinate_system)
    assert p.plan() is not None


def test_custom_cost_function_goes_through_python_evaluate():
    from commonroad_rp_b200.cost_function import CostFunction, DefaultCostFunction
    p, _ = _planner(seed=2)
    base = p.plan(current_sampling_level=2)
    calls = []

    class Mine(CostFunction):
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def evaluate(self, trajectory):
            calls.append(1)
            return self.inner.evaluate(trajectory)

    inner = DefaultCostFunction(15.0, desired_d=0.0, desired_s=None)
    p.set_cost_function(Mine(inner))
    p.reset(initial_state_cart=p.x_0, initial_state_curv=p.x_0_cl, collision_checker=p.collision_checker,
            coordinate_system=p.coordinate_system)
    mine = p.plan(current_sampling_level=2)
    assert len(calls) > 10
    a = np.array([s.position for s in base[0].state_list])
    b = np.array([s.position for s in mine[0].state_list])
    assert np.allclose(a, b, rtol=0, atol=1e-12)          # same winner as the fused device cost


def test_list_form_matches_grid_form():
    """a custom SamplingSpace (only generate_trajectories_at_level) reaches the GPU through rp_plan_list"""
    from commonroad_rp_b200.sampling import FixedIntervalSampling, SamplingSpace
    p, _ = _planner(seed=1)
    grid = p.plan(current_sampling_level=1)
    counts = (p.infeasible_count_kinematics, p.infeasible_count_collision)
    inner = p.sampling_space

    class ListOnly(SamplingSpace):
        def __init__(self):
            super().__init__(inner.num_sampling_levels)

        def generate_trajectories_at_level(self, level, x0_lon, x0_lat, mode, low_vel):
            return inner.generate_trajectories_at_level(level, x0_lon, x0_lat, mode, low_vel)

    p.set_sampling_space(ListOnly())
    p.reset(initial_state_cart=p.x_0, initial_state_curv=p.x_0_cl, collision_checker=p.collision_checker,
            coordinate_system=p.coordinate_system)
    lst = p.plan(current_sampling_level=1)
    assert (p.infeasible_count_kinematics, p.infeasible_count_collision) == counts
    a = np.array([[s.position[0], s.position[1], s.velocity] for s in grid[0].state_list])
    b = np.array([[s.position[0], s.position[1], s.velocity] for s in lst[0].state_list])
    assert np.array_equal(a, b)


def test_draw_mode_stores_all_trajectories_as_views():
    from commonroad_rp_b200.trajectories import FeasibilityStatus
    p, _ = _planner(seed=5, draw=True)
    assert p.plan(current_sampling_level=1) is not None
    st = p.stored_trajectories
    assert st is not None and len(st) == p.last_result.n_candidates
    labels = [t.feasibility_label for t in st]
    assert FeasibilityStatus.FEASIBLE in labels and FeasibilityStatus.INFEASIBLE_KINEMATIC in labels
    t_inf = next(t for t in st if t.feasibility_label == FeasibilityStatus.INFEASIBLE_KINEMATIC)
    assert t_inf.cartesian is not None and len(t_inf.cartesian.x) == p.N + 1           # states kept for plotting
    snap = copy.deepcopy(st[:5])
    p.reset(initial_state_cart=p.x_0, initial_state_curv=p.x_0_cl, collision_checker=p.collision_checker,
            coordinate_system=p.coordinate_system)
    p.plan(current_sampling_level=1)
    assert all(s.cartesian is not None for s in snap)                                    # deep copies survive re-planning


def test_bundle_views_and_sort():
    p, _ = _planner(seed=0)
    x0_lon, x0_lat = p.x_0_cl
    bundle = p._create_trajectory_bundle(x0_lon, x0_lat, samp_level=1)
    best = p._get_optimal_trajectory(bundle)
    assert len(bundle.trajectories) == p.last_result.n_candidates
    feas = [t for t in bundle.trajectories if t.feasibility_label is not None and t.feasibility_label.value == "feasible"]
    assert best.cost == min(t.cost for t in feas)
    assert feas[0].cartesian is not None and feas[0].trajectory_long.coeffs.shape == (6,)
    bundle.trajectories = feas
    bundle.sort()
    assert bundle.min_costs().cost == best.cost and bundle.max_costs().cost >= best.cost


def test_standalone_polynomial_solve_kat():
    """coefficient solve against np.linalg.solve on the reference's own matrices (polynomial_trajectory.py:305-315)"""
    from commonroad_rp_b200.polynomial_trajectory import QuarticTrajectory, QuinticTrajectory, solve_batch
    rng = np.random.default_rng(0)
    n = 500
    tau = rng.uniform(0.2, 6.0, n)
    x0 = rng.uniform(-5, 5, (n, 3)) * np.array([10, 3, 1])
    xd = rng.uniform(-5, 5, (n, 3)) * np.array([10, 3, 0])
    kind = rng.integers(0, 2, n)
    got = solve_batch(kind, x0, xd, tau)
    for q in range(n):
        want = O.solve_quartic(x0[q, 0], x0[q, 1], x0[q, 2], tau[q], xd[q, 0]) if kind[q] == 0 else \
            O.solve_quintic(x0[q, 0], x0[q, 1], x0[q, 2], xd[q, 0], xd[q, 1], xd[q, 2], tau[q])
        assert H.rel_err(want, got[q]) < 1e-9
    single = QuinticTrajectory(0, 2.0, np.array([0., 1., 0.]), np.array([1., 0., 0.]))
    assert np.allclose(single.coeffs, [0, 1, 0, -0.25, 0.0625, 0], atol=1e-12)
    assert np.allclose(QuarticTrajectory(0, 2.0, np.array([0., 1., 0.]), np.array([3., 0.])).coeffs, [0, 1, 0, 0.5, -0.125, 0])


def test_collision_queries_match_oracle():
    from commonroad_rp_b200 import collision
    scn = synthetic.make_scenario(seed=4)
    keys = ("static_boxes", "dyn_t0", "dyn_states", "dyn_lw", "boundary_boxes", "boundary_tris")
    cc = collision.checker_from_arrays(**{k: scn[k] for k in keys})
    cc.add_collision_object(collision.Triangle(100.0, -30.0, 104.0, -30.0, 102.0, -26.0))
    ocheck = O.build_checker({k: scn[k] for k in keys})
    sg = tp.ShapeGroup()
    sg.add_shape(tp.Triangle(100.0, -30.0, 104.0, -30.0, 102.0, -26.0))
    ocheck.add_collision_object(sg)
    rng = np.random.default_rng(1)
    n_hit = 0
    for _ in range(300):
        x = rng.uniform(0, 299)
        y = 20 * np.sin(x / 40) + rng.uniform(-7, 7)
        if rng.random() < 0.15:
            x, y = rng.uniform(98, 106), rng.uniform(-32, -24)
        th = rng.uniform(-np.pi, np.pi)
        t = int(rng.integers(0, 120))
        ego = collision.TimeVariantCollisionObject(t)
        ego.append_obstacle(collision.RectOBB(2.25, 0.8, th, x, y))
        oego = tp.TimeVariantCollisionObject(t)
        oego.append_obstacle(tp.RectOBB(2.25, 0.8, th, x, y))
        want = ocheck.collide(oego)
        assert cc.collide(ego) == want
        n_hit += want
    assert 20 < n_hit < 280


def test_error_behaviour():
    from commonroad_rp_b200._lib import Engine, RpError
    eng = Engine(0)
    with pytest.raises(RpError):
        eng.grid_launch()                                   # nothing uploaded
    inputs = Engine.make_inputs([0, 1, 0], [0, 0, 0], 0.0, 0, False, "velocity_keeping", 20, 0.1, desired_speed=1.0)
    with pytest.raises(RpError):
        eng.plan_grid(inputs, [1.0], [1.0], [0.0])          # tables missing
    with pytest.raises(RpError):
        eng.set_reference([0, 0], [0, 0], [0, 0], [0, 0], np.zeros((2, 2)), [0, 0], np.zeros((2, 2)), 20.0)   # not increasing
    with pytest.raises(RpError):
        Engine(99)
    # empty bundle: not an error, winner -1
    prob_scn = synthetic.make_scenario(seed=0)
    tables = O.reference_tables(prob_scn["ref_path"])
    prob = H.make_problem(prob_scn, [1.0], [10.0], [0.0], [float(tables[0]["ref_pos"][10]), 10.0, 0.0], [0.0, 0, 0], tables=tables)
    e2 = H.engine_for(prob)
    res = e2.plan_grid(H.inputs_for(prob), [], [], [])
    assert res.winner == -1 and res.n_candidates == 0
    # ragged traj_len / out-of-domain s: candidate leaves the reference path => projection rejection, no crash
    prob["x0_lon"][0] = float(tables[0]["ref_pos"][-3])
    res = e2.plan_grid(H.inputs_for(prob), [2.0], [10.0, 15.0], [0.0, 1.0])
    assert res.winner == -1 and res.n_infeasible_kinematics == 4
    e2.close()
    eng.close()


# ---- BASELINE-size bundle: size-independent properties ----------------------------------------------------
def test_full_size_dense_bundle_properties():
    import bench
    work = bench.dense_workload(1)
    eng = bench.make_engine(work, 0, None)
    # the bench's own inputs run the lazy collision pass (check_collision = 2): same winner and counters as full checking,
    # nothing ranked before the winner left without a verdict
    lazy = eng.plan_grid(bench.make_inputs(work), work["t"], work["lon"], work["d"])
    cost_l, status_l, _, _ = eng.fetch_candidates()
    inputs = bench.make_inputs(work, check_collision=1)
    res = eng.plan_grid(inputs, work["t"], work["lon"], work["d"])
    n = work["n_cand"]
    assert res.n_candidates == n == 131072
    cost, status, reason, step = eng.fetch_candidates()
    assert (lazy.winner, lazy.winner_cost, lazy.n_infeasible_collision, lazy.n_infeasible_kinematics, lazy.n_feasible) == \
        (res.winner, res.winner_cost, res.n_infeasible_collision, res.n_infeasible_kinematics, res.n_feasible)
    ranked_before = (cost < res.winner_cost) | ((cost == res.winner_cost) & (np.arange(n) <= res.winner))
    assert np.array_equal(status_l[ranked_before], status[ranked_before]) and np.array_equal(cost_l, cost, equal_nan=True)
    assert np.all(np.isin(status[status_l == 4], (0, 2))) and np.array_equal(status_l[status_l != 4], status[status_l != 4])
    ws = eng.fetch_states(res.winner)
    # (1) arg-min property: the winner is the lexicographic minimum over feasible, collision-free candidates
    ok = status == 0
    assert ok.any() and res.winner == int(np.flatnonzero(ok)[np.argmin(cost[ok])])     # first minimum = lowest index
    assert cost[res.winner] == res.winner_cost == cost[ok].min()
    # (2) counters are consistent with the per-candidate verdicts
    assert res.n_feasible == int(((status == 0) | (status == 2)).sum())
    assert res.n_infeasible_kinematics == int((status == 1).sum()) and res.n_collision_total == int((status == 2).sum())
    assert res.n_infeasible_collision == int(((status == 2) & ((cost < res.winner_cost) | ((cost == res.winner_cost) & (np.arange(n) < res.winner)))).sum())
    assert sum(res.reason_counts) == res.n_infeasible_kinematics
    # (3) idempotence: a second evaluation is bit-identical
    res2 = eng.plan_grid(inputs, work["t"], work["lon"], work["d"])
    cost2, status2, _, _ = eng.fetch_candidates()
    assert res2.winner == res.winner and np.array_equal(status, status2) and np.array_equal(cost, cost2, equal_nan=True)
    # (4) sharding: two t-major shards merge to the same winner / counters
    recs = []
    for first, count in ((0, n // 2), (n // 2, n - n // 2)):
        eng.set_candidate_range(first, count)
        r = eng.plan_grid(inputs, work["t"], work["lon"], work["d"])
        recs.append(r)
    eng.set_candidate_range(0, -1)
    best = min((r for r in recs if r.winner >= 0), key=lambda r: (r.winner_cost, r.winner))
    assert best.winner == res.winner and sum(r.n_infeasible_kinematics for r in recs) == res.n_infeasible_kinematics
    # (5) full-state mode gives the same verdicts and the winner's block equals the select-only one
    fin = bench.make_inputs(work, want_all_states=True)
    res3 = eng.plan_grid(fin, work["t"], work["lon"], work["d"])
    assert res3.winner == res.winner and np.array_equal(eng.fetch_states(res.winner), ws)
    # (6) a seeded sample of candidates agrees with the oracle (flags exact, states 1e-9)
    cl, ct, _ = eng.fetch_coeffs()
    rng = np.random.default_rng(0)
    pick = np.sort(rng.choice(n, 48, replace=False))
    tb = work["cosy"].device_tables()
    prob = H.make_problem(work["scn"], work["t"], work["lon"], work["d"], work["x0_lon"], work["x0_lat"], N=60,
                          x0_orientation=work["x0_orientation"], desired_speed=15.0,
                          tables=({"ref_pos": tb["ref_pos"], "ref_theta": tb["ref_theta"], "ref_curv": tb["ref_curv"],
                                   "ref_curv_d": tb["ref_curv_d"]},
                                  {"path": tb["path_xy"], "S": tb["path_s"], "normals": tb["path_normals"],
                                   "limit": tb["proj_limit"]}, None))
    ocl, oct_, odt, _, _ = O.enumerate_grid(prob["t"], prob["lon"], prob["d"], prob["x0_lon"], prob["x0_lat"],
                                            "velocity_keeping", False)
    o = O.plan_candidates(ocl[pick], oct_[pick], odt[pick], prob, want_states=True, full_collision=True)
    assert np.array_equal(o["status"], status[pick])
    kin = o["status"] != O.ST_KINEMATIC
    assert H.rel_err(o["cost"][kin], cost[pick][kin]) < 1e-9
    for q, k in enumerate(pick):
        if kin[q]:
            assert H.rel_err(o["states"][q], eng.fetch_states(int(k))) < 1e-9
    eng.close()


def test_lazy_collision_mode_is_exact_for_reference_outputs():
    """check_collision = 2 skips candidates costlier than the best collision-free one found so far; the
    winner, the lazy collision count (reactive_planner.py:1043) and every verdict ranked before the winner
    must equal full checking."""
    from commonroad_rp_b200 import _lib
    from tests.test_gpu_parity import _bundle
    cases = (dict(seed=2, level=2, N=60, s_dot0=12.0), dict(seed=1, level=3, N=20, d0=-0.4), dict(seed=0, level=1, N=20))
    # both schedules: the step-parallel kernel gates whole candidates, the candidate-major kernel checks every feasible
    # candidate while it marches (no RP_FEASIBLE_UNCHECKED there) -- the reference outputs are the same
    for case, kernel in [(c, k) for c in cases for k in (_lib.KERNEL_STEP_PARALLEL, _lib.KERNEL_CANDIDATE_MAJOR)]:
        prob = _bundle(**case)
        eng = H.engine_for(prob)
        eng.set_kernel_policy(kernel)
        full = eng.plan_grid(H.inputs_for(prob, check_collision=_lib.COLLISION_ALL), prob["t"], prob["lon"], prob["d"])
        cost_f, status_f, _, step_f = eng.fetch_candidates()
        lazy = eng.plan_grid(H.inputs_for(prob, check_collision=_lib.COLLISION_LAZY), prob["t"], prob["lon"], prob["d"])
        cost_l, status_l, _, step_l = eng.fetch_candidates()
        assert lazy.winner == full.winner and lazy.winner_cost == full.winner_cost
        assert lazy.n_infeasible_collision == full.n_infeasible_collision
        assert lazy.n_infeasible_kinematics == full.n_infeasible_kinematics and lazy.n_feasible == full.n_feasible
        assert np.array_equal(cost_f, cost_l, equal_nan=True)
        n = len(cost_f)
        idx = np.arange(n)
        before = (cost_f < full.winner_cost) | ((cost_f == full.winner_cost) & (idx <= full.winner)) if full.winner >= 0 \
            else np.ones(n, dtype=bool)
        assert np.array_equal(status_f[before], status_l[before]) and np.array_equal(step_f[before], step_l[before])
        unchecked = status_l == _lib.ST_UNCHECKED
        assert np.all(np.isin(status_f[unchecked], (0, 2))) and not (unchecked & before).any()
        others = ~unchecked
        assert np.array_equal(status_f[others], status_l[others])
        eng.close()


def test_scenario_batch_equals_individual_plans():
    """config-5 shape: independent seeded scenarios (different paths, obstacles, horizons, sample levels) evaluated
    by ONE set of launches give, candidate by candidate, what each scenario gives alone"""
    from commonroad_rp_b200 import _lib, collision
    from commonroad_rp_b200.parallel import ScenarioBatch
    from commonroad_rp_b200.utility.config import VehicleConfiguration
    from commonroad_rp_b200.utility.utils_coordinate_system import CoordinateSystem
    veh = VehicleConfiguration()
    batch = ScenarioBatch()
    cycle, single, arrays = [], [], []
    keys = ("static_boxes", "dyn_t0", "dyn_states", "dyn_lw", "boundary_boxes", "boundary_tris")
    for sid in range(7):
        scn, s_dot0, d0 = synthetic.scenario_seeded(sid)
        co = CoordinateSystem(scn["ref_path"])
        cc = collision.checker_from_arrays(**{k: scn[k] for k in keys})
        batch.add_scenario(veh, co, cc)
        N = (20, 30, 20, 60, 20, 25, 20)[sid]
        level = (2, 1, 3, 2, 2, 2, 1)[sid]
        if sid == 5:
            s_dot0 = 2.0                       # low-velocity mode: one lateral system per candidate
        lo = max(0.0, s_dot0 - 0.125 * 2.0 * veh.a_max)
        t, lon, dset = H.level_sets(level, 0.4, N * 0.1, 0.1, lo, max(lo + 5.0, s_dot0 + 2))
        d = [float(x) for x in dset.union({d0})]
        s0 = float(co.ref_pos[10])
        j = int(np.argmax(co.ref_pos > s0)) - 1
        inputs = _lib.Engine.make_inputs([s0, s_dot0, 0.0], [d0, 0.0, 0.0], float(co.ref_theta[j]), 3 * sid, s_dot0 < 4.0,
                                         "velocity_keeping", N, 0.1, desired_speed=s_dot0)
        cycle.append((inputs, t, lon, d))
        eng = _lib.Engine(0)
        eng.set_vehicle(veh.length, veh.width, veh.wb_rear_axle, veh.wheelbase, veh.a_max, veh.v_switch, veh.delta_max,
                        veh.v_delta_max, veh.kappa_max)
        tb = co.device_tables()
        eng.set_reference(tb["ref_pos"], tb["ref_theta"], tb["ref_curv"], tb["ref_curv_d"], tb["path_xy"], tb["path_s"],
                          tb["path_normals"], tb["proj_limit"])
        cc.upload(eng)
        r = eng.plan_grid(inputs, t, lon, d)
        single.append((r.winner, r.winner_cost, r.n_infeasible_kinematics, r.n_infeasible_collision, r.n_feasible,
                       r.n_collision_total, list(r.reason_counts)))
        arrays.append(eng.fetch_candidates())
        eng.close()
    same = lambda a, b: (a == b) or (np.isnan(a) and np.isnan(b))
    for rep in range(2):                      # second cycle reuses the resident tables and staging buffers
        res = batch.plan(cycle)
        got = [(r.winner, r.winner_cost, r.n_infeasible_kinematics, r.n_infeasible_collision, r.n_feasible,
                r.n_collision_total, list(r.reason_counts)) for r in res]
        assert [g[0] for g in got] == [s[0] for s in single]
        assert all(g[2:] == s[2:] for g, s in zip(got, single))
        assert all(same(g[1], s[1]) for g, s in zip(got, single))
        for k in range(len(cycle)):
            cost, status, reason, step = batch.batch.fetch_candidates(k)
            c0, st0, r0, sp0 = arrays[k]
            assert np.array_equal(status, st0) and np.array_equal(reason, r0) and np.array_equal(step, sp0), k
            assert np.array_equal(cost.view(np.int64), c0.view(np.int64)), k          # identical bits
    # the reference's lazy collision pass for every scenario (check_collision = 2): the batch march stores the ego boxes and
    # the deferred checker runs over the shared tile lists -- winner, counters and everything ranked before the winner as
    # with full checking; the rest either unchecked or identical
    import copy
    lazy_cycle = []
    for inputs, t, lon, d in cycle:
        li = copy.copy(inputs)
        li.check_collision = _lib.COLLISION_LAZY
        lazy_cycle.append((li, t, lon, d))
    n_unchecked = 0
    for rep in range(2):
        res = batch.plan(lazy_cycle)
        for k, (r, sgl) in enumerate(zip(res, single)):
            assert (r.winner, r.n_infeasible_kinematics, r.n_infeasible_collision, r.n_feasible, list(r.reason_counts)) == \
                (sgl[0], sgl[2], sgl[3], sgl[4], sgl[6]), (k, rep)
            assert same(r.winner_cost, sgl[1])
            cost, status, reason, step = batch.batch.fetch_candidates(k)
            c0, st0, r0, sp0 = arrays[k]
            assert np.array_equal(cost.view(np.int64), c0.view(np.int64)), k
            unchecked = status == _lib.ST_UNCHECKED
            assert np.all(np.isin(st0[unchecked], (0, 2))), k
            assert np.array_equal(status[~unchecked], st0[~unchecked]) and np.array_equal(step[~unchecked], sp0[~unchecked]), k
            if r.winner >= 0:
                before = (c0 < r.winner_cost) | ((c0 == r.winner_cost) & (np.arange(len(c0)) <= r.winner))
                assert not (unchecked & before).any(), k
            else:
                assert not unchecked.any(), k
            n_unchecked += int(unchecked.sum())
    assert n_unchecked > 0
    # a batch whose scenarios ask for different collision modes keeps the checks inside the march for all of them
    mixed = [lazy_cycle[k] if k % 2 else cycle[k] for k in range(len(cycle))]
    res = batch.plan(mixed)
    for k, (r, sgl) in enumerate(zip(res, single)):
        assert (r.winner, r.n_infeasible_kinematics, r.n_infeasible_collision, r.n_feasible) == (sgl[0], sgl[2], sgl[3], sgl[4]), k
        _, status, _, _ = batch.batch.fetch_candidates(k)
        assert np.array_equal(status, arrays[k][1]), k                               # (no unchecked candidates: every flag)
    assert [r.winner for r in batch.plan(cycle)] == [s[0] for s in single]          # and back to full checking
    packed = _lib.Batch.pack(cycle)           # all scenarios' inputs in one call
    assert [r.winner for r in batch.plan(packed)] == [s[0] for s in single]
    one = batch.plan_one_by_one(cycle)
    assert [r.winner for r in one] == [s[0] for s in single]
    assert len({g[0] for g in got}) > 1       # the scenarios really differ
    ms, n = batch.batch.last_ms()
    assert ms > 0 and n == sum(len(c[1]) * len(c[2]) * len(c[3]) for c in cycle)
    batch.close()


def test_obb_sum_hull_helpers():
    """host helpers of the continuous collision check (collision.obb_sum_hull / trajectory_preprocess_obb_sum); the
    planner-level test is tests/test_gpu_golden.py::test_continuous_collision_check_through_the_planner"""
    from commonroad_rp_b200 import collision
    a = collision.RectOBB(2.0, 1.0, 0.3, 0.0, 0.0)
    b = collision.RectOBB(2.0, 1.0, 0.5, 3.0, 1.0)
    h = collision.obb_sum_hull(a, b)
    assert h.orientation == a.orientation and h.r_x > a.r_x and h.r_y >= a.r_y
    # the hull contains the corners of both boxes
    ca, sa = np.cos(h.orientation), np.sin(h.orientation)
    for box in (a, b):
        cb, sb = np.cos(box.orientation), np.sin(box.orientation)
        for sx in (-1, 1):
            for sy in (-1, 1):
                px = box.cx + sx * box.r_x * cb - sy * box.r_y * sb - h.cx
                py = box.cy + sx * box.r_x * sb + sy * box.r_y * cb - h.cy
                assert abs(px * ca + py * sa) <= h.r_x + 1e-12 and abs(-px * sa + py * ca) <= h.r_y + 1e-12
    tvo = collision.TimeVariantCollisionObject(5)
    for k in range(4):
        tvo.append_obstacle(collision.RectOBB(2.0, 1.0, 0.1 * k, 1.0 * k, 0.0))
    hull, err = collision.trajectory_preprocess_obb_sum(tvo)
    assert err == 0 and hull.time_start_idx() == 5 and hull.time_end_idx() == 7


def test_merge_records_kernel_matches_host_merge():
    """the one-warp merge of the all-gathered shard records (multi-GPU arg-min) against the torch reference"""
    import torch
    from commonroad_rp_b200 import _lib
    from commonroad_rp_b200.parallel import merge_records
    eng = _lib.Engine(0)
    inf = float("inf")
    rng = np.random.default_rng(3)
    for world in (1, 2, 8, 40):
        g = rng.uniform(0, 10, (world, 4))
        g[:, 1] = rng.integers(0, 1000, world)
        g[:, 2:] = rng.integers(0, 50, (world, 2))
        if world > 2:
            g[1, 0] = g[0, 0]                 # cost tie: lowest index wins
            g[2] = [inf, inf, 3, 0]           # a shard without a feasible candidate
        gathered = torch.tensor(g, dtype=torch.float64, device="cuda:0")
        winner = torch.empty(2, dtype=torch.float64, device="cuda:0")
        totals = torch.empty(2, dtype=torch.float64, device="cuda:0")
        torch.cuda.synchronize()
        eng.merge_records_dev(gathered.data_ptr(), world, winner.data_ptr(), totals.data_ptr())
        eng.synchronize()
        w, t = merge_records(gathered)
        assert torch.equal(w.cpu(), winner.cpu()) and torch.equal(t.cpu(), totals.cpu())
    eng.close()


def test_reference_tables_derived_on_the_device():
    """rp_ctx_set_reference_polyline derives the nine reference tables from the smoothed polyline on
    the device (CoordinateSystem.__init__, utils_coordinate_system.py:101-118 + the frame construction); they equal the
    oracle's numpy derivation, and a bundle planned on them gives the same verdicts as on uploaded tables."""
    from tests.test_gpu_parity import _bundle
    u = np.linspace(0.0, 1.0, 160)
    paths = {
        "sine": synthetic.sine_path(20.0, 40.0, 300),
        "straight": synthetic.sine_path(0.0, 40.0, 120),
        # a left-hand loop of 1.5 turns: the heading crosses +-pi twice (np.unwrap)
        "loop": np.stack([(30.0 + 25.0 * u) * np.cos(3.0 * np.pi * u), (30.0 + 25.0 * u) * np.sin(3.0 * np.pi * u)], axis=1),
    }
    prob = _bundle(seed=2, level=2, N=30, s_dot0=12.0)
    eng = H.engine_for(prob)
    for name, raw in paths.items():
        ref, ccosy, _ = O.reference_tables(raw)
        polyline = np.asarray(ccosy["path"])[1:-1]                   # the smoothed reference without its extension vertices
        eng.set_reference_polyline(polyline, ccosy["limit"], 1e-4)
        got = eng.get_reference()
        assert len(got["ref_pos"]) == len(polyline) + 2
        for key, want in (("ref_pos", ref["ref_pos"]), ("ref_theta", ref["ref_theta"]), ("ref_curv", ref["ref_curv"]),
                          ("ref_curv_d", ref["ref_curv_d"]), ("path_xy", ccosy["path"]), ("path_s", ccosy["S"]),
                          ("path_normals", ccosy["normals"])):
            # the curvature rate at the two extension vertices divides a curvature difference by eps2 = 1e-4: a last-bit
            # difference in pow() / the vertex norm is amplified 1e4 times there (interior rows agree to ~1e-13)
            tol = 1e-8 if key == "ref_curv_d" else 1e-9
            assert H.rel_err(want, got[key]) < tol, (name, key, H.rel_err(want, got[key]))
            if key == "ref_curv_d":
                assert H.rel_err(want[2:-2], got[key][2:-2]) < 1e-10, (name, key)
        if name == "loop":
            assert np.ptp(ref["ref_theta"]) > 2.0 * np.pi               # the case really unwraps
    eng.close()
    # the same bundle on uploaded tables and on device-derived tables
    eng_a, eng_b = H.engine_for(prob), H.engine_for(prob)
    eng_b.set_reference_polyline(np.asarray(prob["ccosy"]["path"])[1:-1], prob["ccosy"]["limit"], 1e-4)
    ra = H.run_engine_grid(eng_a, prob, want_all_states=False)
    rb = H.run_engine_grid(eng_b, prob, want_all_states=False)
    assert ra["winner"] == rb["winner"] and np.array_equal(ra["status"], rb["status"])
    assert np.array_equal(ra["reason"], rb["reason"]) and np.array_equal(ra["step"], rb["step"])
    kin = ra["status"] != 1
    assert H.rel_err(ra["cost"][kin], rb["cost"][kin]) < 1e-9
    assert H.rel_err(ra["winner_states"], rb["winner_states"]) < 1e-9
    eng_a.close()
    eng_b.close()


def test_corridor_sampling_on_interval_arrays_matches_oracle():
    """CorridorSampling (sampling.py:273-397) fed with plain (s, d, v) interval arrays -- no
    CommonRoad-Reach -- reaches the GPU through the list form; verdicts, winner and counters equal the oracle's evaluation
    of the same candidate list"""
    from commonroad_rp_b200.sampling import CorridorSampling, IntervalCorridor
    p, scn = _planner(seed=3, N=20, s_dot0=12.0)
    s0 = p.x_0_cl[0][0]
    corridor = {}
    for k in range(0, 21):
        s_mid = s0 + 12.0 * 0.1 * k
        corridor[k] = np.array([[s_mid - 8.0, s_mid + 10.0, -2.5, 0.4, 8.0, 14.0],
                                [s_mid - 6.0, s_mid + 12.0, 0.1, 1.2, 10.0, 15.0],
                                [s_mid - 2.0, s_mid + 6.0, 2.2, 3.0, 9.0, 11.0]])
    cs = CorridorSampling(p.config)
    cs.driving_corridor = IntervalCorridor(corridor)
    p.set_sampling_space(cs)
    p.reset(initial_state_cart=p.x_0, initial_state_curv=p.x_0_cl, collision_checker=p.collision_checker,
            coordinate_system=p.coordinate_system)
    out = p.plan(current_sampling_level=2)
    res = p.last_result
    assert res.n_candidates > 100
    cands = cs.generate_trajectories_at_level(2, p.x_0_cl[0], p.x_0_cl[1], "velocity_keeping", False)
    assert len(cands) == res.n_candidates
    cl = np.array([c.trajectory_long.coeffs for c in cands])
    ct = np.array([c.trajectory_lat.coeffs for c in cands])
    dtau = np.array([c.trajectory_long.delta_tau for c in cands])
    tables = O.reference_tables(scn["ref_path"])
    prob = H.make_problem(scn, [1.0], [1.0], [0.0], p.x_0_cl[0], p.x_0_cl[1], N=20, x0_orientation=p.x_0.orientation,
                          desired_speed=12.0, tables=tables)
    o = O.plan_candidates(cl, ct, dtau, prob, want_states=True, full_collision=False)
    assert o["winner"] == res.winner and (out is not None) == (o["winner"] >= 0)
    assert o["n_infeasible_kinematics"] == p.infeasible_count_kinematics
    assert o["n_infeasible_collision"] == p.infeasible_count_collision
    if out is not None:
        got = np.array([[st.position[0], st.position[1], st.velocity] for st in out[0].state_list]).T
        assert H.rel_err(o["states"][o["winner"]][[0, 1, 3]], got) < 1e-9


def _write_commonroad_xml(path):
    """a small CommonRoad 2020a file of our own: a straight two-lane road (lanelets 1 -> 2, neighbour 3 -> 4 on the left)
    with a side road (5) branching off lanelet 1's end to the right, a parked car, a moving car, a planning problem"""
    def bound(xs, ys):
        return "".join("<point><x>%.6f</x><y>%.6f</y></point>" % (x, y) for x, y in zip(xs, ys))

    def lanelet(lid, xs, yl, yr, succ=(), pred=(), adj_l=None, adj_r=None):
        s = '<lanelet id="%d"><leftBound>%s</leftBound><rightBound>%s</rightBound>' % (lid, bound(xs, yl), bound(xs, yr))
        s += "".join('<predecessor ref="%d"/>' % p for p in pred) + "".join('<successor ref="%d"/>' % q for q in succ)
        if adj_l is not None:
            s += '<adjacentLeft ref="%d" drivingDir="same"/>' % adj_l
        if adj_r is not None:
            s += '<adjacentRight ref="%d" drivingDir="same"/>' % adj_r
        return s + "</lanelet>"

    x1, x2 = np.arange(0.0, 61.0, 1.0), np.arange(60.0, 161.0, 1.0)
    one = np.ones_like
    xml = ['<?xml version="1.0" ?><commonRoad commonRoadVersion="2020a" benchmarkID="ZAM_Synth-1_1_T-1" timeStepSize="0.1">']
    xml.append(lanelet(1, x1, 1.75 * one(x1), -1.75 * one(x1), succ=(2, 5), adj_l=3))
    xml.append(lanelet(2, x2, 1.75 * one(x2), -1.75 * one(x2), pred=(1,), adj_l=4))
    xml.append(lanelet(3, x1, 5.25 * one(x1), 1.75 * one(x1), succ=(4,), adj_r=1))
    xml.append(lanelet(4, x2, 5.25 * one(x2), 1.75 * one(x2), pred=(3,), adj_r=2))
    # side road: leaves lanelet 1 at x = 60 and bends away to the right
    u = np.linspace(0.0, 1.0, 31)
    cx, cy = 60.0 + 40.0 * u, -25.0 * u ** 2
    th = np.arctan2(-50.0 * u, 40.0)
    xml.append('<lanelet id="5"><leftBound>%s</leftBound><rightBound>%s</rightBound><predecessor ref="1"/></lanelet>' % (
        bound(cx - 1.75 * np.sin(th), cy + 1.75 * np.cos(th)), bound(cx + 1.75 * np.sin(th), cy - 1.75 * np.cos(th))))
    xml.append('<staticObstacle id="30"><type>parkedVehicle</type><shape><rectangle><length>4.5</length><width>2.0</width>'
               '</rectangle></shape><initialState><position><point><x>55.0</x><y>-0.9</y></point></position>'
               '<orientation><exact>0.05</exact></orientation><time><exact>0</exact></time></initialState></staticObstacle>')
    states = "".join('<state><position><point><x>%.4f</x><y>3.5</y></point></position><orientation><exact>0.0</exact></orientation>'
                     '<time><exact>%d</exact></time><velocity><exact>9.0</exact></velocity></state>' % (20.0 + 0.9 * k, k)
                     for k in range(1, 61))
    xml.append('<dynamicObstacle id="31"><type>car</type><shape><rectangle><length>4.8</length><width>1.9</width></rectangle></shape>'
               '<initialState><position><point><x>20.0</x><y>3.5</y></point></position><orientation><exact>0.0</exact></orientation>'
               '<time><exact>0</exact></time><velocity><exact>9.0</exact></velocity></initialState><trajectory>%s</trajectory>'
               '</dynamicObstacle>' % states)
    xml.append('<planningProblem id="100"><initialState><position><point><x>12.0</x><y>0.2</y></point></position>'
               '<orientation><exact>0.01</exact></orientation><time><exact>0</exact></time><velocity><exact>11.0</exact></velocity>'
               '<yawRate><exact>0.0</exact></yawRate><slipAngle><exact>0.0</exact></slipAngle></initialState>'
               '<goalState><position><lanelet ref="2"/></position><time><intervalStart>40</intervalStart><intervalEnd>80</intervalEnd></time>'
               '<velocity><intervalStart>8.0</intervalStart><intervalEnd>14.0</intervalEnd></velocity></goalState></planningProblem>')
    xml.append("</commonRoad>")
    with open(path, "w") as fh:
        fh.write("".join(xml))


def test_scenario_file_to_device_tables_through_the_planner(tmp_path):
    """SURVEY 8f rank 4, second half: a CommonRoad XML file -> scenario objects (the package's own reader when
    commonroad-io is absent) -> ReactivePlanner(config) builds checker, road boundary and device tables itself
    (reference utility/general.py:11-29, reactive_planner.py:218-256); the cycle equals the oracle's on the arrays the
    checker holds, and a ScenarioBatch fed from the same scenario object selects the same trajectory"""
    from commonroad_rp_b200 import _lib
    from commonroad_rp_b200.parallel import ScenarioBatch
    from commonroad_rp_b200.reactive_planner import ReactivePlanner
    from commonroad_rp_b200.utility import scenario_io
    from commonroad_rp_b200.utility.config import ReactivePlannerConfiguration
    from commonroad_rp_b200.utility.general import load_scenario_and_planning_problem
    path = str(tmp_path / "ZAM_Synth-1_1_T-1.xml")
    _write_commonroad_xml(path)
    scenario, problem, _ = load_scenario_and_planning_problem(path)
    assert len(scenario.lanelet_network.lanelets) == 5 and len(scenario.dynamic_obstacles) == 1
    route = scenario_io.find_route(scenario, problem)
    assert route == [1, 2]
    cfg = ReactivePlannerConfiguration()
    cfg.planning.time_steps_computation = 20
    cfg.planning.planning_horizon = 2.0
    cfg.update(scenario=scenario, planning_problem=problem)
    planner = ReactivePlanner(cfg)                     # reset(): x0 from the planning problem, checker from the scenario
    ref_path = scenario_io.route_reference_path(scenario, route, extend_back=5.0)
    planner.set_reference_path(ref_path)
    planner.set_desired_velocity(current_speed=planner.x_0.velocity)
    assert planner._desired_speed == 11.0              # mid-point of the goal interval (utility/general.py:32-46)
    out = planner.plan()
    assert out is not None and not planner.goal_reached()
    res = planner.last_result
    arr = planner.collision_checker.device_arrays()
    # road boundary: the junction mouth between lanelet 1 and the side road carries no wall, the outer borders do
    seg_mid = arr["static_obb"][1:, :2]
    on_right_border = np.abs(seg_mid[:, 1] + 1.75) < 1e-9
    assert not np.any(on_right_border & (seg_mid[:, 0] > 63.0) & (seg_mid[:, 0] < 75.0))     # the side road's mouth is open
    assert np.any(on_right_border & (np.abs(seg_mid[:, 0] - 55.5) < 1e-9)) and np.any(on_right_border & (seg_mid[:, 0] > 80.0))
    assert np.any(np.abs(seg_mid[:, 1] - 5.25) < 1e-9) and len(arr["dyn_boxes"][0]) == 61
    # the oracle on the same arrays
    co = planner.coordinate_system
    tb = co.device_tables()
    x0_lon, x0_lat = planner.x_0_cl
    bundle = planner._create_trajectory_bundle(x0_lon, x0_lat, samp_level=1)
    dev = bundle.device
    st = arr["static_obb"]
    obstacles = {"static_boxes": np.zeros((0, 5)), "boundary_boxes": st, "boundary_tris": arr["tris"], "dyn_t0": arr["dyn_t0"],
                 "dyn_states": [b[:, :3] for b in arr["dyn_boxes"]], "dyn_lw": [(2 * b[0, 3], 2 * b[0, 4]) for b in arr["dyn_boxes"]]}
    prob = {"t": dev["t"], "lon": dev["lon"], "d": dev["d"], "x0_lon": np.array(x0_lon), "x0_lat": np.array(x0_lat),
            "x0_orientation": planner.x_0.orientation, "x0_time_step": 0, "lon_mode": "velocity_keeping", "low_vel_mode": False,
            "dt": 0.1, "N": 20, "factor": 1, "draw_all": False, "constraints": O.CONSTRAINTS,
            "cost": {"kind": "default", "desired_speed": 11.0, "desired_s": None, "desired_d": 0.0, "w_a": 5},
            "vehicle": O.vehicle_dict(), "ref": {k: tb[k] for k in ("ref_pos", "ref_theta", "ref_curv", "ref_curv_d")},
            "ccosy": {"path": tb["path_xy"], "S": tb["path_s"], "normals": tb["path_normals"], "limit": tb["proj_limit"]},
            "obstacles": obstacles}
    o = O.plan_grid(prob, want_states=True, full_collision=False)
    assert o["winner"] == res.winner and o["n_infeasible_kinematics"] == res.n_infeasible_kinematics
    assert o["n_infeasible_collision"] == res.n_infeasible_collision
    got = np.array([[s.position[0], s.position[1], s.velocity] for s in out[0].state_list]).T
    assert H.rel_err(o["states"][o["winner"]][[0, 1, 3]], got) < 1e-9
    # the same scenario object through the batch path
    batch = ScenarioBatch()
    batch.add_scenario_from(scenario, ref_path, cfg.vehicle)
    inputs = planner._plan_inputs(x0_lon, x0_lat, planner._device_cost_spec(), False)
    r = batch.plan([(inputs, dev["t"], dev["lon"], dev["d"])])[0]
    assert r.winner == res.winner and r.n_infeasible_collision == res.n_infeasible_collision
    batch.close()


def test_closed_loop_scenario_batch_equals_individual_planners():
    """SURVEY 8f rank 1 ("batched multi-scenario reset()"): three replanning cycles of several scenarios as ONE batch --
    plan, advance every scenario three steps along its winner (rp_batch_winner_states), next cycle -- against one
    ReactivePlanner per scenario running run_planner.py's loop (plan, reset to state_list[3] / lon_list[3] / lat_list[3])"""
    from commonroad_rp_b200 import collision
    from commonroad_rp_b200.parallel import ScenarioBatch
    from commonroad_rp_b200.state import ReactivePlannerState
    planners, scns = [], []
    for sid in range(5):
        p, scn = _planner(seed=20 + sid, N=20, s_dot0=9.0 + sid, d0=0.1 * sid - 0.2)
        planners.append(p)
        scns.append(scn)
    batch = ScenarioBatch()
    for p in planners:
        batch.add_scenario(p.vehicle_params, p.coordinate_system, p.collision_checker)
    x0 = np.array([[p.x_0.position[0], p.x_0.position[1], p.x_0.orientation, p.x_0.velocity, 0.0, 0.0] for p in planners])
    lon, lat = batch.reset(x0)                                       # Cartesian -> curvilinear for all scenarios in one launch
    for k, p in enumerate(planners):
        # (the test planners' curvilinear VELOCITIES are set by hand, only the positions derive from the Cartesian state)
        assert abs(p.x_0_cl[0][0] - lon[k][0]) < 1e-8 and abs(p.x_0_cl[1][0] - lat[k][0]) < 1e-8
    batch.reset(x0, states_curv=(np.array([p.x_0_cl[0] for p in planners]), np.array([p.x_0_cl[1] for p in planners])))
    level, steps = 2, 3
    for cycle in range(3):
        outs, cyc = [], []
        for k, p in enumerate(planners):
            p.set_desired_velocity(current_speed=p.x_0.velocity)
            outs.append(p.plan(current_sampling_level=level))
            assert outs[-1] is not None
            t, lon_s, d = p.sampling_space.sample_grid(level, batch.x0_lat[k], "velocity_keeping")
            p._low_vel_mode = bool(batch.x0_cart[k, 3] < p.config.planning.low_vel_mode_threshold)
            inputs = type(p._inputs).from_buffer_copy(p._plan_inputs(batch.x0_lon[k], batch.x0_lat[k], p._device_cost_spec(), False))
            inputs.x0_orientation = float(batch.x0_cart[k, 2])
            cyc.append((inputs, t, lon_s, d))
        res = batch.plan(cyc)
        assert [r.winner for r in res] == [p.last_result.winner for p in planners], cycle
        w = batch.advance(steps)
        for k, (p, out) in enumerate(zip(planners, outs)):
            cart, _, lon_list, lat_list = out
            st = cart.state_list[steps]
            want = [st.position[0], st.position[1], st.orientation, st.velocity, st.acceleration]
            assert w[k, 12] == 1.0 and H.rel_err(want, w[k, 0:5]) < 1e-9, (cycle, k)
            assert H.rel_err(lon_list[steps], w[k, 6:9]) < 1e-9 and H.rel_err(lat_list[steps], w[k, 9:12]) < 1e-9, (cycle, k)
            assert np.array_equal(batch.x0_lon[k], w[k, 6:9])
            nxt = ReactivePlannerState(time_step=st.time_step, position=st.position, orientation=st.orientation, velocity=st.velocity,
                                       acceleration=st.acceleration, yaw_rate=st.yaw_rate, steering_angle=st.steering_angle)
            p.reset(initial_state_cart=nxt, initial_state_curv=(lon_list[steps], lat_list[steps]),
                    collision_checker=p.collision_checker, coordinate_system=p.coordinate_system)
    batch.close()
