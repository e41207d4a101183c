"""CPU tests of the host side: the mirror of the reference's Python API (sampling order, configuration,
coordinate-system tables, containers), the C-ABI library surface, and the multi-GPU merge logic.
No CUDA compute calls are made here."""
import ctypes
import os
import re

import numpy as np
import pytest

from commonroad_rp_b200 import _lib, collision
from commonroad_rp_b200.cost_function import DefaultCostFunction, DefaultCostFunctionFailSafe
from commonroad_rp_b200.sampling import FixedIntervalSampling, PositionSampling, TimeSampling, VelocitySampling
from commonroad_rp_b200.trajectories import CartesianSample, CurviLinearSample, TrajectoryBundle
from commonroad_rp_b200.utility import synthetic
from commonroad_rp_b200.utility.config import ReactivePlannerConfiguration
from commonroad_rp_b200.utility.utils_coordinate_system import CoordinateSystem, interpolate_angle
from oracle import rp_oracle as O
from oracle import third_party as tp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- C-ABI surface ------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    """the .so loads (no GPU needed) and exports exactly what include/rp_b200.h declares"""
    header = open(os.path.join(ROOT, "include", "rp_b200.h")).read()
    declared = set(re.findall(r"\b(rp_[a-z0-9_]+)\s*\(", header))
    lib = _lib.load_library()
    for name in declared:
        assert hasattr(lib, name), "missing export %s" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.rp_version() >= 100


def test_struct_layouts_match_header_sizes():
    # rp_plan_inputs: 7 doubles, 4 int32, double, 2 int32 + uint32 + 3 int32, 4 doubles, 2 int32
    assert ctypes.sizeof(_lib.VehicleParams) == 9 * 8
    assert ctypes.sizeof(_lib.PlanResult) == 6 * 4 + 8 * 4 + 8
    assert ctypes.sizeof(_lib.PlanInputs) % 8 == 0


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.RpError):
        _lib.Engine(0)


# ---- sampling ------------------------------------------------------------------------------------------
def test_sample_sets_follow_reference_expressions():
    ts = TimeSampling(0.2, 2.0, 4, 0.1)
    assert sorted(ts.samples_at_level(0)) == pytest.approx([0.2, 1.2])
    assert len(ts.samples_at_level(1)) == 4 and len(ts.samples_at_level(3)) == 10
    assert 2.1 not in ts.samples_at_level(3)
    for cls in (PositionSampling, VelocitySampling):
        s = cls(-3, 3, 4)
        assert [len(s.samples_at_level(k)) for k in range(4)] == [3, 5, 9, 17]
    with pytest.raises(AssertionError):
        TimeSampling(0.1, 2.0, 4, 0.1)


def test_sample_grid_is_python_set_order():
    cfg = ReactivePlannerConfiguration()
    cfg.planning.time_steps_computation = 20
    fs = FixedIntervalSampling(cfg)
    fs.samples_v = VelocitySampling(10.0, 17.0, 4)
    d0 = 0.1
    t, lon, d = fs.sample_grid(1, [d0, 0, 0], "velocity_keeping")
    assert list(t) == [float(x) for x in fs.samples_t.samples_at_level(1)]
    assert list(d) == [float(x) for x in fs.samples_d.samples_at_level(1).union({d0})]
    assert len(d) == 6 and d0 in d
    # d0 on the grid is de-duplicated
    assert len(fs.sample_grid(1, [0.0, 0, 0], "velocity_keeping")[2]) == 5
    with pytest.raises(AttributeError):
        fs.sample_grid(1, [0.0, 0, 0], "bogus")


@pytest.mark.reference
def test_sampling_order_identical_to_reference():
    from oracle import ref_harness as H
    scn = synthetic.make_scenario(seed=0)
    p = H.build_planner(scn, N=20, t_min=0.2)
    p.set_desired_velocity(desired_velocity=13.0, current_speed=13.0)
    cfg = ReactivePlannerConfiguration()
    cfg.planning.time_steps_computation = 20
    cfg.sampling.t_min = 0.2
    fs = FixedIntervalSampling(cfg)
    ref_v = p.sampling_space.samples_v
    fs.samples_v = VelocitySampling(ref_v.low, ref_v.up, 4)
    for level in range(4):
        for d0 in (0.0, 0.1, -1.2345, 3.0):
            t, lon, d = fs.sample_grid(level, [d0, 0, 0], "velocity_keeping")
            rs = p.sampling_space
            assert list(t) == [float(x) for x in rs.samples_t.samples_at_level(level)]
            assert list(lon) == [float(x) for x in rs.samples_v.samples_at_level(level)]
            assert list(d) == [float(x) for x in rs.samples_d.samples_at_level(level).union({d0})]


# ---- configuration -------------------------------------------------------------------------------------
def test_config_defaults_and_yaml(tmp_path):
    cfg = ReactivePlannerConfiguration()
    assert cfg.planning.dt == 0.1 and cfg.planning.time_steps_computation == 60 and cfg.sampling.num_sampling_levels == 4
    assert cfg.vehicle.kappa_max == pytest.approx(0.7017693147614497, rel=0, abs=0)
    assert cfg.vehicle.wheelbase == pytest.approx(2.5789128)
    y = tmp_path / "c.yaml"
    y.write_text("planning:\n  dt: 0.1\n  time_steps_computation: 20\n  low_vel_mode_threshold: 2\nsampling:\n  t_min: 1.0\n"
                 "debug:\n  draw_traj_set: True\n  multiproc: True\n  num_workers: 6\nvehicle:\n  id_type_vehicle: 2\n")
    c = ReactivePlannerConfiguration.load(str(y), "ZAM_Tjunction-1_42_T-1.xml")
    assert c.planning.time_steps_computation == 20 and c.sampling.t_min == 1.0 and c.debug.draw_traj_set
    assert c.general.path_scenario.endswith("ZAM_Tjunction-1_42_T-1.xml")
    assert c["planning"]["dt"] == 0.1
    with pytest.raises(KeyError):
        c["nope"]


@pytest.mark.reference
def test_reference_yaml_files_load():
    for name in ("ZAM_Over-1_1", "DEU_Test-1_1_T-1", "ZAM_Tjunction-1_42_T-1"):
        c = ReactivePlannerConfiguration.load("/root/reference/configurations/%s.yaml" % name, name + ".xml")
        assert c.planning.time_steps_computation == 20


# ---- coordinate system ---------------------------------------------------------------------------------
def test_coordinate_system_tables_match_oracle_restatement():
    scn = synthetic.make_scenario(seed=1, amplitude=12.0, wavelength=55.0)
    co = CoordinateSystem(scn["ref_path"])
    ref, ccosy, cc = O.reference_tables(scn["ref_path"])
    tb = co.device_tables()
    for mine, theirs in ((tb["ref_pos"], ref["ref_pos"]), (tb["ref_theta"], ref["ref_theta"]), (tb["ref_curv"], ref["ref_curv"]),
                         (tb["ref_curv_d"], ref["ref_curv_d"]), (tb["path_xy"], ccosy["path"]), (tb["path_normals"], ccosy["normals"])):
        assert np.allclose(mine, theirs, rtol=1e-11, atol=1e-11)
    rng = np.random.default_rng(0)
    for _ in range(50):
        s = rng.uniform(1.0, co.ref_pos[-1] - 1.0)
        d = rng.uniform(-4, 4)
        xy = co.convert_to_cartesian_coords(s, d)
        assert np.allclose(xy, cc.convert_to_cartesian_coords(s, d), atol=1e-10)
        sd = co.convert_to_curvilinear_coords(xy[0], xy[1])
        assert np.allclose(sd, [s, d], atol=1e-8)
    assert co.convert_to_cartesian_coords(-5.0, 0.0) is None
    assert co.convert_to_cartesian_coords(10.0, 25.0) is None
    with pytest.raises(ValueError):
        co.convert_to_curvilinear_coords(1e4, 1e4)


def test_interpolate_angle_is_plain_lerp():
    assert interpolate_angle(0.5, 0.0, 1.0, 0.2, 0.4) == pytest.approx(0.3)
    assert interpolate_angle(0.5, 0.0, 1.0, 3.0, -3.0) == pytest.approx(0.0)        # no shortest-arc logic
    assert interpolate_angle(0.0, 0.0, 1.0, 7.0, 7.0) == pytest.approx(7.0 - 2 * np.pi)


# ---- containers ----------------------------------------------------------------------------------------
def _filled_states(n, traj_len, seed):
    rng = np.random.default_rng(seed)
    st = rng.uniform(-1, 1, size=(14, n))
    st[:, traj_len:] = 0.0
    return st


def test_enlarge_matches_oracle_restatement():
    n, tl, dt = 21, 13, 0.1
    st = _filled_states(n, tl, 0)
    a, b = st.copy(), st.copy()
    ca = CartesianSample(a[0], a[1], a[2], a[3], a[4], a[5], a[6], current_time_step=tl)
    cu = CurviLinearSample(a[7], a[8], a[9], current_time_step=tl, ss=a[10], sss=a[11], dd=a[12], ddd=a[13])
    ca.enlarge(dt)
    cu.enlarge(dt)
    O.enlarge_cartesian(b[0], b[1], b[2], b[3], b[4], b[5], b[6], tl, dt)
    O.enlarge_curvilinear(b[7], b[8], b[9], b[10], b[11], b[12], b[13], tl, dt)
    assert np.array_equal(a, b)
    assert ca.current_time_step == n and cu.current_time_step == n


def test_cost_functions_match_oracle_restatement():
    class _S:
        pass
    st = _filled_states(61, 61, 3)
    s = _S()
    s.cartesian = CartesianSample(*st[:7], current_time_step=61)
    s.curvilinear = CurviLinearSample(st[7], st[8], st[9], current_time_step=61, ss=st[10], sss=st[11], dd=st[12], ddd=st[13])
    cf = DefaultCostFunction(desired_speed=0.3, desired_d=0.1, desired_s=None)
    assert cf.evaluate(s) == O.default_cost(st, {"w_a": 5, "desired_speed": 0.3, "desired_s": None, "desired_d": 0.1})
    cf.desired_s, cf.w_a = 2.0, 1
    assert cf.evaluate(s) == O.default_cost(st, {"w_a": 1, "desired_speed": 0.3, "desired_s": 2.0, "desired_d": 0.1})
    assert DefaultCostFunctionFailSafe().evaluate(s) == O.failsafe_cost(st)
    assert cf.device_spec()["cost_kind"] == _lib.COST_DEFAULT


def test_collision_checker_packing():
    scn = synthetic.make_scenario(seed=2, n_dynamic=3)
    cc = collision.checker_from_arrays(**{k: scn[k] for k in ("static_boxes", "dyn_t0", "dyn_states", "dyn_lw", "boundary_boxes", "boundary_tris")})
    cc.add_collision_object(collision.Triangle(0, 0, 1, 0, 0, 1))
    arr = cc.device_arrays()
    assert arr["static_obb"].shape == (2 + len(scn["boundary_boxes"]), 5)
    assert arr["static_obb"][0, 3] == pytest.approx(0.5 * scn["static_boxes"][0, 3])      # half extents on the wire
    assert len(arr["dyn_boxes"]) == 3 and arr["dyn_boxes"][0].shape == (101, 5)
    assert arr["tris"].shape == (1, 6)
    with pytest.raises(TypeError):
        cc.add_collision_object(object()) or cc.device_arrays()


def test_bundle_api_on_plain_samples():
    from commonroad_rp_b200.polynomial_trajectory import QuinticTrajectory
    from commonroad_rp_b200.trajectories import TrajectorySample
    c = np.zeros(6)
    mk = lambda goal: TrajectorySample(2.0, 0.1, QuinticTrajectory(0, 2.0, np.array([1.0, 0, 0]), np.array([goal, 0, 0]), coeffs=c.copy()),
                                       QuinticTrajectory(0, 2.0, np.zeros(3), np.zeros(3), coeffs=c.copy()))
    b = TrajectoryBundle([mk(0.5), mk(2.0), mk(1.0)], cost_function=None)
    assert not b.empty
    b.filter_goals_behind()
    assert len(b.trajectories) == 1 and b.trajectories[0].trajectory_long.x_d[0] == 2.0
    with pytest.raises(AssertionError):
        TrajectoryBundle([1, 2], None)


def test_polynomial_evaluators():
    from commonroad_rp_b200.polynomial_trajectory import QuarticTrajectory
    c = np.array([1.0, 2.0, 0.5, -0.1, 0.01, 0.0])
    q = QuarticTrajectory(0, 3.0, np.array([1.0, 2.0, 1.0]), np.array([4.0, 0.0]), coeffs=c)
    t = np.array([0.0, 0.5, 3.0])
    assert np.allclose(q.calc_position(t, t**2, t**3, t**4, t**5), np.polyval(c[::-1], t))
    assert np.allclose(q.calc_velocity(t, t**2, t**3, t**4), np.polyval(np.polyder(c[::-1]), t))
    assert np.allclose(q.calc_acceleration(t, t**2, t**3), np.polyval(np.polyder(c[::-1], 2), t))
    assert np.allclose(q.evaluate_state_at_tau(10.0), q.evaluate_state_at_tau(3.0))      # clamped
    with pytest.raises(AssertionError):
        q.coeffs = np.zeros(5)


# ---- multi-GPU merge logic on CPU ------------------------------------------------------------------------
def test_shard_ranges_cover_the_bundle():
    from commonroad_rp_b200.parallel import shard_range
    for n in (0, 1, 7, 131072, 131073):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert sum(c for _, c in spans) == n
            pos = 0
            for f, c in spans:
                assert f == min(pos, n) or c == 0
                pos += c


def test_merge_records_lexicographic():
    import torch
    from commonroad_rp_b200.parallel import merge_records
    inf = float("inf")
    g = torch.tensor([[3.0, 40.0, 5, 10], [2.0, 90.0, 1, 4], [2.0, 70.0, 0, 6], [inf, inf, 9, 0]], dtype=torch.float64)
    w, tot = merge_records(g)
    assert w.tolist() == [2.0, 70.0] and tot.tolist() == [15.0, 20.0]
    w, _ = merge_records(torch.tensor([[inf, inf, 1, 0]], dtype=torch.float64))
    assert w[1].item() == inf


def test_collision_checker_version_tracks_content():
    """the planner re-uploads the obstacle tables when the checker's content moved (identity + version): objects added,
    and shapes appended to an object AFTER it was added"""
    from commonroad_rp_b200 import collision
    cc = collision.CollisionChecker()
    v0 = cc.version
    tvo = collision.TimeVariantCollisionObject(0)
    tvo.append_obstacle(collision.RectOBB(1, 1, 0, 0, 0))
    cc.add_collision_object(tvo)
    v1 = cc.version
    tvo.append_obstacle(collision.RectOBB(1, 1, 0, 1, 0))
    v2 = cc.version
    sg = collision.ShapeGroup()
    cc.add_collision_object(sg)
    v3 = cc.version
    sg.add_shape(collision.Triangle(0, 0, 1, 0, 0, 1))
    assert len({v0, v1, v2, v3, cc.version}) == 5
    assert len(cc.device_arrays()["dyn_boxes"][0]) == 2 and len(cc.device_arrays()["tris"]) == 1


def _strip(x0, x1, y0, y1, horizontal, adj_left=None, adj_right=None):
    """a straight lanelet with 1 m border sampling"""
    from types import SimpleNamespace
    if horizontal:
        xs = np.arange(x0, x1 + 0.5, 1.0)
        left = np.stack([xs, np.full_like(xs, y1)], axis=1)
        right = np.stack([xs, np.full_like(xs, y0)], axis=1)
    else:
        ys = np.arange(y0, y1 + 0.5, 1.0)
        left = np.stack([np.full_like(ys, x0), ys], axis=1)       # driving +y: left border at the smaller x
        right = np.stack([np.full_like(ys, x1), ys], axis=1)
    return SimpleNamespace(left_vertices=left, right_vertices=right, adj_left=adj_left, adj_right=adj_right)


def test_road_boundary_of_a_crossing_has_no_walls_inside_the_junction():
    """two roads crossing at right angles, neither lanelet has lateral neighbours: every border lacks an adjacent lanelet,
    but only the parts OUTSIDE the other road are road boundary -- independent expectation: 4 arms x 2 sides x 18 m"""
    from commonroad_rp_b200 import collision
    a = _strip(-20.0, 20.0, -2.0, 2.0, horizontal=True)
    b = _strip(-2.0, 2.0, -20.0, 20.0, horizontal=False)
    segs = collision.road_boundary_segments([a, b])
    length = np.hypot(*(segs[:, 1] - segs[:, 0]).T)
    assert len(segs) == 144 and length.sum() == pytest.approx(144.0)
    mid = 0.5 * (segs[:, 0] + segs[:, 1])
    assert not np.any((np.abs(mid[:, 0]) < 2.0) & (np.abs(mid[:, 1]) < 2.0))            # nothing inside the junction square
    # the naive construction (every neighbour-less border) would have put 4 x 4 m of wall across the junction
    scn = type("S", (), {"lanelet_network": type("LN", (), {"lanelets": [a, b]})()})()
    _, sg = collision.create_road_boundary_obstacle(scn)
    assert len(sg.unpack()) == 144 and all(isinstance(s, collision.RectOBB) for s in sg.unpack())
    _, tg = collision.create_road_boundary_obstacle(scn, method="triangulation", band=1.0)
    tris = np.array([t.row() for t in tg.unpack()]).reshape(-1, 3, 2)
    assert len(tris) > 200 and all(isinstance(t, collision.Triangle) for t in tg.unpack())
    cen = tris.mean(axis=1)
    on_road = ((np.abs(cen[:, 1]) < 2.0) & (np.abs(cen[:, 0]) < 20.0)) | ((np.abs(cen[:, 0]) < 2.0) & (np.abs(cen[:, 1]) < 20.0))
    assert not on_road.any()                                                             # the band lies off-road
    # laterally adjacent lanelets share a border: no boundary between them
    l1 = _strip(0.0, 30.0, 0.0, 3.0, True, adj_left=2)
    l2 = _strip(0.0, 30.0, 3.0, 6.0, True, adj_right=1)
    segs2 = collision.road_boundary_segments([l1, l2])
    assert len(segs2) == 60 and set(np.round(0.5 * (segs2[:, 0, 1] + segs2[:, 1, 1]), 6)) == {0.0, 6.0}


def test_interval_corridor_reach_operations():
    """CorridorSampling on plain (s, d, v) interval arrays (no CommonRoad-Reach): the four reach operations of
    sampling.py:312, :365-371 on node rows (s_lo, s_hi, d_lo, d_hi, v_lo, v_hi)"""
    from commonroad_rp_b200.sampling import CorridorSampling, IntervalCorridor, _IntervalReachOperations as ops
    from commonroad_rp_b200.utility.config import ReactivePlannerConfiguration
    nodes = np.array([[10.0, 14.0, -1.0, 0.5, 4.0, 6.0],      # two nodes overlapping laterally ...
                      [12.0, 16.0, 0.2, 1.5, 5.0, 8.0],
                      [11.0, 15.0, 2.5, 3.0, 3.0, 5.0],       # ... one apart (another lane)
                      [30.0, 34.0, -0.5, 0.5, 9.0, 9.5]])     # not reached longitudinally
    assert ops.lon_velocity_interval_connected_set(nodes) == (3.0, 9.5)
    hit = ops.determine_overlapping_nodes_with_lon_pos(nodes, 13.0)
    assert len(hit) == 3
    comps = ops.determine_connected_components(list(hit))
    assert [len(c) for c in comps] == [2, 1]
    assert ops.lat_interval_connected_set(comps[0]) == (-1.0, 1.5) and ops.lat_interval_connected_set(comps[1]) == (2.5, 3.0)
    assert ops.determine_connected_components([]) == []
    cfg = ReactivePlannerConfiguration()
    cfg.planning.time_steps_computation = 20
    cs = CorridorSampling(cfg)                                   # constructing needs no CommonRoad-Reach ...
    with pytest.raises(AttributeError):
        cs.generate_trajectories_at_level(1, [0, 10, 0], [0, 0, 0], "velocity_keeping", False)
    with pytest.raises(ImportError):
        cs.driving_corridor = {0: object()}                      # ... a commonroad_reach corridor does
    cs.driving_corridor = IntervalCorridor({k: nodes for k in range(0, 25)})
    assert cs._velocity_constraints[3] == [3.0, 9.5]


def test_c_output_packing_equals_python_packing():
    """csrc/rp_pack.c (one C call for plan()'s output packing, reference :514-568) against the Python loop it replaces:
    identical state lists incl. the orientation shift, steering angles, yaw rates, time steps, the lazy curvilinear list"""
    import json
    from commonroad_rp_b200 import build, reactive_planner as RP
    from commonroad_rp_b200.trajectories import DeviceTrajectorySample
    from tests import golden_io, helpers as H
    build.build_pack_module()
    import importlib
    pack = importlib.import_module("commonroad_rp_b200._rp_pack")
    z = np.load(os.path.join(ROOT, "tests", "golden", "plan_free.npz"))
    prob = golden_io.unpack_problem(z)
    meta = json.loads(str(z["plan_meta"]))
    planner = H.planner_from_fixture(prob, z["ref_path_raw"], z["x0"], desired_velocity=meta["desired_velocity"])
    n = z["out_cart"].shape[1]

    def sample():
        blk = np.vstack([z["out_cart"][[0, 1, 2, 3, 4]], np.zeros((2, n)), z["out_lon"].T[0:1], z["out_lat"].T[0:1], np.zeros((1, n)),
                         z["out_lon"].T[1:3], z["out_lat"].T[1:3]])
        blk[5] = np.tan(z["out_cart"][6]) / planner.vehicle_params.wheelbase
        blk[2, 5:9] += 2 * np.pi                                   # exercise the orientation shift
        ts = DeviceTrajectorySample(planner.horizon, planner.dt, lambda: (None, None))
        ts._set_states(np.ascontiguousarray(blk))
        return ts

    saved = RP._rp_pack
    try:
        RP._rp_pack = pack
        a = planner._compute_trajectory_pair(sample())
        RP._rp_pack = None
        b = planner._compute_trajectory_pair(sample())
    finally:
        RP._rp_pack = saved
    ga, gb = H.plan_output_arrays(a), H.plan_output_arrays(b)
    for key in ga:
        assert np.array_equal(ga[key], gb[key]), key
    assert type(a[0].state_list[0]) is type(b[0].state_list[0]) and isinstance(a[2][1], list) and isinstance(a[2][1][0], float)
    assert np.max(np.abs(ga["out_cart"][2] - z["out_cart"][2])) < 1e-9       # the shift folded the extra turn away
    with pytest.raises(ValueError):
        pack.pack(type(a[0].state_list[0]), np.zeros(13), [], 0, 1, 0.1, 2.5, 0.0, -3.0, 3.0)


def test_c_binding_of_the_cycle_call_passes_the_same_arguments():
    """csrc/rp_pack.c plan_levels: the levels' arrays concatenated exactly as Engine.plan_levels' ctypes path stages them"""
    import ctypes as C
    _rp_pack = pytest.importorskip("commonroad_rp_b200._rp_pack")
    seen = {}
    proto = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                        C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_double),
                        C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32))

    def fake(ctx, inp, n_levels, n_t, n_lon, n_d, t_cat, tl_cat, lon_cat, d_cat, out, n_eval, chosen):
        nt, nl, nd = sum(n_t[:n_levels]), sum(n_lon[:n_levels]), sum(n_d[:n_levels])
        seen.update(ctx=ctx, inp=inp, out=out, n=(list(n_t[:n_levels]), list(n_lon[:n_levels]), list(n_d[:n_levels])),
                    t=list(t_cat[:nt]), tl=list(tl_cat[:nt]), lon=list(lon_cat[:nl]), d=list(d_cat[:nd]))
        n_eval[0], chosen[0] = 2, 1
        return 0

    cb = proto(fake)
    levels = [(np.array([1.0, 2.0]), np.array([3.0, 4.0, 5.0]), np.array([0.5]), np.array([11, 21], dtype=np.int32)),
              (np.array([1.0, 1.5, 2.0]), np.array([3.0]), np.array([0.5, -0.5]), np.array([11, 16, 21], dtype=np.int32))]
    rc, chosen, n_eval, counts = _rp_pack.plan_levels(C.cast(cb, C.c_void_p).value, 1234, 5678, levels, 91011)
    assert (rc, chosen, n_eval, counts) == (0, 1, 2, (6, 6))
    assert (seen["ctx"], seen["inp"], seen["out"]) == (1234, 5678, 91011)
    assert seen["n"] == ([2, 3], [3, 1], [1, 2])
    assert seen["t"] == [1.0, 2.0, 1.0, 1.5, 2.0] and seen["tl"] == [11, 21, 11, 16, 21]
    assert seen["lon"] == [3.0, 4.0, 5.0, 3.0] and seen["d"] == [0.5, 0.5, -0.5]
    with pytest.raises(TypeError):          # float32 samples: left to the ctypes path (which converts)
        _rp_pack.plan_levels(C.cast(cb, C.c_void_p).value, 0, 0, [(np.zeros(2, np.float32),) + levels[0][1:]], 0)
    with pytest.raises(ValueError):
        _rp_pack.plan_levels(C.cast(cb, C.c_void_p).value, 0, 0, [(np.zeros(500), levels[0][1], levels[0][2], np.zeros(500, np.int32))], 0)


class _FakeCycleEngine:
    """stands in for _lib.Engine in plan()'s cycle path: records what each rp_plan_levels submission contained and answers
    from a table {level size: has a winner}"""
    plan_generation = 0
    _cyc_chosen = 0
    _cyc_selected = 0

    def __init__(self, winners, states):
        self.res = (_lib.PlanResult * 4)()
        self.winners, self.states, self.submissions = winners, states, []

    def cycle_limits(self):
        return (4, 448, 128, 1 << 20)

    def plan_levels(self, inputs, levels):
        self.plan_generation += 1
        sizes = [len(t) * len(lon) * len(d) for t, lon, d, _ in levels]
        self.submissions.append(sizes)
        chosen = len(levels) - 1
        for j, n in enumerate(sizes):
            ok = self.winners[len(levels[j][2])]            # keyed by the level's number of d samples
            self.res[j].winner = 3 if ok else -1
            self.res[j].n_candidates = n
            self.res[j].n_infeasible_kinematics = 10 * (j + 1)
            self.res[j].winner_cost = 1.0
            if ok:
                chosen = j
                break
        self._cyc_chosen = self._cyc_selected = chosen
        return self.res, chosen

    def select_level(self, j):
        self._cyc_selected = j

    def cycle_winner_states(self):
        return self.states.copy()

    def fetch_states(self, k):
        return self.states.copy()


@pytest.mark.parametrize("policy,winning_level,expected", [
    ("after_failure", 1, [1]), ("after_failure", 2, [1, 2]), ("after_failure", 3, [1, 2]), ("after_failure", None, [1, 2]),
    ("always", 1, [3]), ("always", 3, [3]), ("never", 3, [1, 1, 1]), ("never", 1, [1])])
def test_speculation_policy_of_the_escalation_loop(policy, winning_level, expected):
    """plan()'s level escalation (reference reactive_planner.py:616-636) on the cycle path: how many rp_plan_levels
    submissions a cycle takes and how many sampling levels each carries, per ``ReactivePlanner.speculation``; the result is
    the escalation loop's whatever the policy (lowest level with a winner; counters of the last level reached)"""
    from commonroad_rp_b200.reactive_planner import ReactivePlanner
    from commonroad_rp_b200.state import ReactivePlannerState
    from tests.helpers import EmptyScenario
    cfg = ReactivePlannerConfiguration()
    cfg.planning.time_steps_computation = 20
    cfg.sampling.t_min = 0.4
    cfg.update(scenario=EmptyScenario(), planning_problem=None)
    planner = ReactivePlanner(cfg)
    planner.speculation = policy
    scn = synthetic.make_scenario(seed=0)
    co = CoordinateSystem(scn["ref_path"])
    states = np.zeros((14, 21))
    states[3] = 10.0                                                        # velocity row: no standstill fallback
    n_d = {1: 6, 2: 10, 3: 18}                                              # d samples per level (+ d0)
    fake = _FakeCycleEngine({n_d[lv]: (lv == winning_level) for lv in (1, 2, 3)}, states)
    planner._engine = fake
    planner._sync_device_tables = lambda: None
    x0 = ReactivePlannerState(time_step=0, position=np.array([10.0, 0.0]), orientation=0.0, velocity=10.0, acceleration=0.0,
                              yaw_rate=0.0, steering_angle=0.0)
    planner.reset(initial_state_cart=x0, initial_state_curv=([10.0, 10.0, 0.0], [0.3, 0.0, 0.0]),
                  collision_checker=collision.CollisionChecker(), coordinate_system=co)
    planner.set_desired_velocity(desired_velocity=10.0, current_speed=10.0)
    out = planner.plan()
    assert [len(s) for s in fake.submissions] == expected
    assert (out is not None) == (winning_level is not None)
    reached = winning_level if winning_level is not None else 3
    # counters are those of the last level the loop reached: its position inside ITS submission decides the record
    per_submission = {"after_failure": {1: 0, 2: 0, 3: 1}, "always": {1: 0, 2: 1, 3: 2}, "never": {1: 0, 2: 0, 3: 0}}[policy]
    assert planner.infeasible_count_kinematics == 10 * (per_submission[reached] + 1)
