"""GPU parity: the CUDA path, called through the C-ABI, against the oracle on the same seeded inputs."""
import os

import numpy as np
import pytest

from commonroad_rp_b200.utility import synthetic
from oracle import rp_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _bundle(seed, level, N, low_vel=False, lon_mode="velocity_keeping", draw_all=False, s_dot0=15.0, d0=0.3,
            amplitude=20.0, wavelength=40.0, t_min=0.4, x0_time_step=0, desired_s=None, static_offset=0.0, wall=False,
            **scn_kw):
    scn = synthetic.make_scenario(seed=seed, amplitude=amplitude, wavelength=wavelength, static_offset=static_offset, **scn_kw)
    if wall:        # a wall across the road where the vehicle starts: every candidate collides at its first step
        scn = dict(scn)
        scn["static_boxes"] = np.concatenate([np.asarray(scn["static_boxes"], dtype=np.float64).reshape(-1, 5),
                                              np.array([[scn["ref_path"][10, 0], scn["ref_path"][10, 1], 0.0, 4.0, 200.0]])])
    dt = 0.1
    if lon_mode == "velocity_keeping":
        lo = max(0.0, s_dot0 - 0.125 * N * dt * 11.5)
        hi = max(lo + 5.0, s_dot0 + 2)
    else:
        lo, hi = desired_s - 1.0, desired_s + 1.0
    t, lon, dset = H.level_sets(level, t_min, N * dt, dt, lo, hi)
    d = [float(x) for x in dset.union({d0})]
    tables = O.reference_tables(scn["ref_path"])
    s0 = float(tables[0]["ref_pos"][10])
    return H.make_problem(scn, t, lon, d, [s0, s_dot0, 0.0], [d0, 0.0, 0.0], N=N, dt=dt, lon_mode=lon_mode,
                          low_vel_mode=low_vel, draw_all=draw_all, desired_speed=s_dot0, desired_s=desired_s,
                          x0_time_step=x0_time_step, tables=tables, w_a=5 if lon_mode == "velocity_keeping" else 1)


CASES = [
    dict(seed=0, level=1, N=20),
    dict(seed=0, level=2, N=20),
    dict(seed=1, level=3, N=20, d0=-0.4),
    dict(seed=2, level=2, N=60, s_dot0=12.0),
    dict(seed=3, level=2, N=20, low_vel=True, s_dot0=3.0),
    dict(seed=4, level=2, N=30, s_dot0=1.0, low_vel=False),          # standstill carry branch
    dict(seed=5, level=2, N=20, draw_all=True),
    dict(seed=6, level=2, N=30, lon_mode="stopping", s_dot0=8.0, desired_s=24.0),
    dict(seed=7, level=2, N=20, amplitude=0.0, static_offset=1.0),
    dict(seed=8, level=1, N=20, x0_time_step=30),
    dict(seed=9, level=2, N=7, t_min=0.2),                           # N+1 == 8: pairwise-sum edge
    dict(seed=10, level=2, N=6, t_min=0.2),                          # N+1 < 8: plain-loop sum
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join("%s=%s" % kv for kv in c.items()))
def test_bundle_parity(case):
    prob = _bundle(**case)
    o = O.plan_grid(prob, want_states=True, full_collision=True)
    eng = H.engine_for(prob)
    g = H.run_engine_grid(eng, prob, want_all_states=True)
    H.assert_parity(o, g, prob, tag=str(case))
    eng.close()


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join("%s=%s" % kv for kv in c.items()))
def test_bundle_parity_candidate_major(case):
    """the candidate-major kernel (rp_cand.cuh; select-only mode) against the oracle, and bit-for-bit against
    the step-parallel kernel (draw mode falls back to the step-parallel kernel by design)"""
    from commonroad_rp_b200 import _lib
    prob = _bundle(**case)
    o = O.plan_grid(prob, want_states=True, full_collision=True)
    eng = H.engine_for(prob)
    g = H.run_engine_grid(eng, prob, want_all_states=False, kernel=_lib.KERNEL_CANDIDATE_MAJOR)
    H.assert_parity(o, g, prob, tag="cand " + str(case))
    g2 = H.run_engine_grid(eng, prob, want_all_states=False, kernel=_lib.KERNEL_STEP_PARALLEL)
    assert np.array_equal(g["status"], g2["status"]) and np.array_equal(g["reason"], g2["reason"])
    assert np.array_equal(g["step"], g2["step"])
    assert np.array_equal(g["cost"].view(np.int64), g2["cost"].view(np.int64))       # identical bits, NaNs included
    assert g["winner"] == g2["winner"]
    eng.close()


def test_shared_reciprocal_division_is_ieee_division():
    """div_rcp(a, b, rcp_refined(b)) (rp_device.cuh) == a / b bit for bit: random operands over the full exponent
    range, the operand classes the path produces (zeros, tiny, huge, denormal, inf, nan), and numpy's own quotient"""
    from commonroad_rp_b200._lib import Engine
    rng = np.random.default_rng(7)
    n = 1 << 20
    mant = lambda: rng.uniform(1.0, 2.0, n) * rng.choice([-1.0, 1.0], n)
    a = np.concatenate([mant() * 2.0 ** rng.integers(-60, 60, n), mant() * 2.0 ** rng.integers(-1070, 1023, n),
                        rng.normal(size=n), rng.uniform(-100, 100, n)])
    b = np.concatenate([mant() * 2.0 ** rng.integers(-60, 60, n), mant() * 2.0 ** rng.integers(-1070, 1023, n),
                        rng.choice([0.1, 100000.0, 2.5789, 1.0, 3.0], n), rng.uniform(0.5, 1.0, n)])
    special = np.array([0.0, -0.0, 5e-324, -5e-324, 2.2250738585072014e-308, 1e-300, 1e-292, 1e-291, 1.0, -1.0, 3.0, 1e300,
                        1.7976931348623157e308, 2.0 ** 1017, 2.0 ** 1016, np.inf, -np.inf, np.nan])
    sa, sb = np.meshgrid(special, special)
    a = np.concatenate([a, sa.ravel()])
    b = np.concatenate([b, sb.ravel()])
    eng = Engine(0)
    q1, q2, rej = eng.selftest_divide(a, b)
    # operands of the path's magnitudes never leave the fast form -- zero dividends of either sign included
    az = np.concatenate([np.zeros(8), -np.zeros(8), rng.normal(size=4096), np.rint(rng.normal(size=4096) * 0.4)])
    bz = rng.choice([0.1, 100000.0, 2.5789, -0.7, 15.0], len(az))
    z1, z2, zrej = eng.selftest_divide(az, bz)
    eng.close()
    assert not zrej.any()
    assert np.array_equal(z1.view(np.int64), z2.view(np.int64))
    with np.errstate(all="ignore"):
        assert np.array_equal(z1.view(np.int64), (az / bz).view(np.int64))
    assert rej[:n].sum() == 0                                       # |a|, |b| in [2^-60, 2^60]
    same = (q1.view(np.int64) == q2.view(np.int64)) | (np.isnan(q1) & np.isnan(q2))
    assert same.all(), (a[~same][:5], b[~same][:5], q1[~same][:5], q2[~same][:5])
    with np.errstate(all="ignore"):
        ref = a / b
    same = (q1.view(np.int64) == ref.view(np.int64)) | (np.isnan(q1) & np.isnan(ref))
    assert same.all(), (a[~same][:5], b[~same][:5], q1[~same][:5], ref[~same][:5])


@pytest.mark.parametrize("low_vel", [False, True], ids=["high_vel", "low_vel"])
def test_warp_shared_longitudinal_rows(low_vel):
    """n_d a multiple of 32: the candidate-major kernel computes the (t, lon)-invariant part of a step once per warp
    (cand_kernel<128, true>) -- same bits as the per-lane form and the step-parallel kernel, parity with the oracle"""
    from commonroad_rp_b200 import _lib
    prob = _bundle(seed=2, level=1, N=40, s_dot0=3.0 if low_vel else 12.0, low_vel=low_vel, t_min=0.8)
    d0 = float(prob["x0_lat"][0])
    rng = np.random.default_rng(5)
    d = np.concatenate([[d0], rng.uniform(-3.0, 3.0, 63)])
    prob["d"] = np.array([float(x) for x in set(d)])                  # 64 lateral targets, d0 among them
    assert len(prob["d"]) == 64
    o = O.plan_grid(prob, want_states=True, full_collision=True)
    eng = H.engine_for(prob)
    g = H.run_engine_grid(eng, prob, want_all_states=False, kernel=_lib.KERNEL_CANDIDATE_MAJOR)
    H.assert_parity(o, g, prob, tag="shared-lon")
    g2 = H.run_engine_grid(eng, prob, want_all_states=False, kernel=_lib.KERNEL_STEP_PARALLEL)
    assert np.array_equal(g["status"], g2["status"]) and np.array_equal(g["reason"], g2["reason"])
    assert np.array_equal(g["step"], g2["step"])
    assert np.array_equal(g["cost"].view(np.int64), g2["cost"].view(np.int64))
    eng.close()


@pytest.mark.parametrize("mode", ["velocity_keeping", "stopping"])
def test_list_form_candidate_major(mode):
    """rp_plan_list through the candidate-major kernel (every lane its own longitudinal polynomial and traj_len,
    filter_goals_behind as skip flags): identical, candidate by candidate, to the grid form of the same bundle"""
    from commonroad_rp_b200 import _lib
    kw = dict(seed=6, level=2, N=30, lon_mode="stopping", s_dot0=8.0, desired_s=24.0) if mode == "stopping" \
        else dict(seed=2, level=2, N=40, s_dot0=12.0)
    prob = _bundle(**kw)
    eng = H.engine_for(prob)
    g = H.run_engine_grid(eng, prob, want_all_states=False, kernel=_lib.KERNEL_CANDIDATE_MAJOR)
    n_t, n_lon, n_d = len(prob["t"]), len(prob["lon"]), len(prob["d"])
    cl = np.repeat(g["coeffs_lon"].reshape(n_t, n_lon, 1, 6), n_d, axis=2).reshape(-1, 6) if g["coeffs_lon"].shape[0] != g["n"] \
        else g["coeffs_lon"]
    ct = g["coeffs_lat"]
    tl = np.repeat([_lib.traj_len_of(t, prob["dt"]) for t in prob["t"]], n_lon * n_d).astype(np.int32)
    skip = (g["status"] == 3).astype(np.uint8)
    # shuffle the list so that lanes of a warp get different polynomials AND different traj_len
    perm = np.random.default_rng(1).permutation(g["n"])
    eng.set_kernel_policy(_lib.KERNEL_CANDIDATE_MAJOR)
    res = eng.plan_list(H.inputs_for(prob), cl[perm], ct[perm], tl[perm], skip[perm])
    cost, status, reason, step = eng.fetch_candidates()
    assert np.array_equal(status, g["status"][perm]) and np.array_equal(reason, g["reason"][perm])
    assert np.array_equal(step, g["step"][perm])
    assert np.array_equal(cost.view(np.int64), g["cost"][perm].view(np.int64))
    assert res.n_infeasible_kinematics == g["n_infeasible_kinematics"]
    if g["winner"] >= 0:
        assert res.winner_cost == g["winner_cost"]
    eng.close()


@pytest.mark.parametrize("kernel", ["step_parallel", "candidate_major"])
def test_edge_cases_no_obstacles_no_check_empty_bundle(kernel):
    """no obstacle tables at all, collision check switched off, and an empty bundle -- through both schedules"""
    from commonroad_rp_b200 import _lib
    pol = _lib.KERNEL_STEP_PARALLEL if kernel == "step_parallel" else _lib.KERNEL_CANDIDATE_MAJOR
    prob = _bundle(seed=0, level=2, N=20)
    empty = {"static_boxes": np.zeros((0, 5)), "dyn_t0": [], "dyn_states": [], "dyn_lw": np.zeros((0, 2)),
             "boundary_boxes": np.zeros((0, 5)), "boundary_tris": np.zeros((0, 6))}
    free = dict(prob, obstacles=empty)
    o = O.plan_grid(free, want_states=True, full_collision=True)
    eng = H.engine_for(free)
    g = H.run_engine_grid(eng, free, want_all_states=False, kernel=pol)
    H.assert_parity(o, g, free, tag="no obstacles " + kernel)
    assert g["n_collision_total"] == 0
    eng.close()
    # collision check off on a scenario WITH obstacles: same verdicts as the obstacle-free scenario
    eng = H.engine_for(prob)
    eng.set_kernel_policy(pol)
    res = eng.plan_grid(H.inputs_for(prob, check_collision=0), prob["t"], prob["lon"], prob["d"])
    cost, status, reason, step = eng.fetch_candidates()
    assert np.array_equal(status, g["status"]) and np.array_equal(cost.view(np.int64), g["cost"].view(np.int64))
    assert res.winner == g["winner"]
    # empty bundle
    res = eng.plan_grid(H.inputs_for(prob), [], prob["lon"], prob["d"])
    assert res.winner == -1 and res.n_candidates == 0 and res.n_feasible == 0
    eng.close()


def _shifted(scn, ox, oy):
    out = dict(scn)
    out["ref_path"] = np.asarray(scn["ref_path"], dtype=np.float64) + [ox, oy]
    sb = np.array(scn["static_boxes"], dtype=np.float64).reshape(-1, 5)
    sb[:, 0] += ox; sb[:, 1] += oy
    out["static_boxes"] = sb
    bb = np.array(scn["boundary_boxes"], dtype=np.float64).reshape(-1, 5)
    bb[:, 0] += ox; bb[:, 1] += oy
    out["boundary_boxes"] = bb
    out["dyn_states"] = [np.asarray(s, dtype=np.float64) + [ox, oy, 0.0] for s in scn["dyn_states"]]
    return out


@pytest.mark.parametrize("kernel", ["step_parallel", "candidate_major"])
def test_utm_sized_coordinates(kernel):
    """the single-precision collision pre-rejects work on positions relative to the obstacle table's own origin: a
    scenario moved to UTM-sized coordinates (4e5, 5.2e6) gives the verdicts of the oracle evaluated there"""
    from commonroad_rp_b200 import _lib
    scn = _shifted(synthetic.make_scenario(seed=2), 4.0e5, 5.2e6)
    dt, N, s_dot0, d0 = 0.1, 30, 12.0, 0.3
    lo = max(0.0, s_dot0 - 0.125 * N * dt * 11.5)
    t, lon, dset = H.level_sets(2, 0.4, N * dt, dt, lo, max(lo + 5.0, s_dot0 + 2))
    d = [float(x) for x in dset.union({d0})]
    tables = O.reference_tables(scn["ref_path"])
    s0 = float(tables[0]["ref_pos"][10])
    prob = H.make_problem(scn, t, lon, d, [s0, s_dot0, 0.0], [d0, 0.0, 0.0], N=N, dt=dt, desired_speed=s_dot0, tables=tables)
    o = O.plan_grid(prob, want_states=True, full_collision=True)
    assert (o["status"] == O.ST_COLLISION).any()
    eng = H.engine_for(prob)
    g = H.run_engine_grid(eng, prob, want_all_states=False,
                          kernel=_lib.KERNEL_STEP_PARALLEL if kernel == "step_parallel" else _lib.KERNEL_CANDIDATE_MAJOR)
    # flags, reasons, steps, winner and counters exact; costs 1e-9.  Positions of magnitude 5e6 carry ~1e-9 absolute
    # rounding, so the states are compared with that absolute floor (helpers.rel_err scales by max(|a|, |b|, 1)).
    H.assert_parity(o, g, prob, tag="utm " + kernel)
    eng.close()


def test_schedules_identical_on_randomised_bundles():
    """kernel-vs-kernel identity (status, reason, step, cost bits, winner, counters) on randomised bundles: horizons,
    sampling levels, initial speeds incl. low-velocity and standstill-carry regimes, stopping mode, time offsets,
    straight and tight reference paths -- no oracle involved, so the bundles can be many"""
    from commonroad_rp_b200 import _lib
    rng = np.random.default_rng(2026)
    n_checked = 0
    for trial in range(24):
        N = int(rng.choice([12, 20, 33, 47, 60]))
        level = int(rng.choice([1, 2, 3]))
        regime = trial % 4
        kw = dict(seed=int(rng.integers(0, 50)), level=level, N=N, d0=float(rng.uniform(-1.0, 1.0)),
                  amplitude=float(rng.choice([0.0, 10.0, 20.0])), wavelength=float(rng.choice([25.0, 40.0, 80.0])),
                  x0_time_step=int(rng.integers(0, 40)), t_min=float(rng.choice([0.2, 0.4, 1.0])))
        if regime == 0:
            kw.update(s_dot0=float(rng.uniform(6.0, 22.0)))
        elif regime == 1:
            kw.update(s_dot0=float(rng.uniform(0.5, 3.5)), low_vel=True)
        elif regime == 2:
            kw.update(s_dot0=float(rng.uniform(0.2, 1.5)), low_vel=False)            # standstill carry branch
        else:
            kw.update(lon_mode="stopping", s_dot0=float(rng.uniform(4.0, 10.0)), desired_s=float(rng.uniform(15.0, 40.0)))
        kw["t_min"] = min(kw["t_min"], N * 0.1)
        prob = _bundle(**kw)
        eng = H.engine_for(prob)
        a = H.run_engine_grid(eng, prob, want_all_states=False, kernel=_lib.KERNEL_CANDIDATE_MAJOR)
        b = H.run_engine_grid(eng, prob, want_all_states=False, kernel=_lib.KERNEL_STEP_PARALLEL)
        eng.close()
        tag = str(kw)
        assert np.array_equal(a["status"], b["status"]), tag
        assert np.array_equal(a["reason"], b["reason"]) and np.array_equal(a["step"], b["step"]), tag
        assert np.array_equal(a["cost"].view(np.int64), b["cost"].view(np.int64)), tag
        assert a["winner"] == b["winner"] and a["n_infeasible_collision"] == b["n_infeasible_collision"], tag
        assert a["reason_counts"] == b["reason_counts"] and a["n_collision_total"] == b["n_collision_total"], tag
        n_checked += a["n"]
    assert n_checked > 10000


# ---- one replanning cycle in one launch (rp_plan_levels) ------------------------------------------------------------
def _level_problems(seed, N, s_dot0, levels=(1, 2, 3), **kw):
    return [_bundle(seed=seed, level=lv, N=N, s_dot0=s_dot0, **kw) for lv in levels]


@pytest.mark.parametrize("case", [dict(seed=0, N=20, s_dot0=15.0), dict(seed=3, N=20, s_dot0=3.0, low_vel=True),
                                  dict(seed=2, N=60, s_dot0=12.0, levels=(1, 2)), dict(seed=5, N=20, s_dot0=15.0, draw=True),
                                  dict(seed=6, N=30, s_dot0=8.0, lon_mode="stopping", desired_s=24.0),
                                  dict(seed=4, N=30, s_dot0=1.0)],
                         ids=lambda c: "-".join("%s=%s" % kv for kv in c.items()))
def test_cycle_launch_equals_one_launch_chain_per_level(case):
    """rp_plan_levels evaluates several sampling levels in ONE kernel (coefficients solved in shared memory, selection by
    the last block, result block in mapped host memory): every level's verdicts, costs, counters and the winner's states
    are bit-identical to rp_plan_grid on that level alone, and the chosen level is the lowest one with a winner"""
    from commonroad_rp_b200 import _lib
    from commonroad_rp_b200._lib import traj_len_of
    case = dict(case)
    draw = case.pop("draw", False)
    probs = _level_problems(**case)
    eng = H.engine_for(probs[0])
    singles = []
    for pr in probs:
        pr["draw_all"] = draw
        r = eng.plan_grid(H.inputs_for(pr, want_all_states=draw), pr["t"], pr["lon"], pr["d"])
        cost, status, reason, step = eng.fetch_candidates()
        singles.append({"res": r, "cost": cost, "status": status, "reason": reason, "step": step,
                        "ws": eng.fetch_states(r.winner) if r.winner >= 0 else None,
                        "coeffs": eng.fetch_coeffs()})
    levels = [(pr["t"], pr["lon"], pr["d"], np.array([traj_len_of(x, pr["dt"]) for x in pr["t"]], dtype=np.int32)) for pr in probs]
    for rep in range(2):
        records, chosen = eng.plan_levels(H.inputs_for(probs[0], want_all_states=draw), levels)
        want_chosen = next((j for j, sg in enumerate(singles) if sg["res"].winner >= 0), len(singles) - 1)
        assert chosen == want_chosen
        assert eng.launches_per_plan() == 1
        for j in range(chosen + 1):
            r, sg = records[j], singles[j]
            for f in ("winner", "n_candidates", "n_feasible", "n_infeasible_kinematics", "n_infeasible_collision", "n_collision_total"):
                assert getattr(r, f) == getattr(sg["res"], f), (j, f)
            assert list(r.reason_counts) == list(sg["res"].reason_counts)
            assert r.winner_cost == sg["res"].winner_cost or (np.isnan(r.winner_cost) and np.isnan(sg["res"].winner_cost))
            eng.select_level(j)
            cost, status, reason, step = eng.fetch_candidates()
            assert np.array_equal(status, sg["status"]) and np.array_equal(reason, sg["reason"]) and np.array_equal(step, sg["step"])
            assert np.array_equal(cost.view(np.int64), sg["cost"].view(np.int64))                  # identical bits
            if r.winner >= 0:
                assert np.array_equal(eng.fetch_states(r.winner), sg["ws"])
                other = int(np.flatnonzero(status != 1)[-1])                                       # a non-winner, re-evaluated on demand
                assert eng.fetch_states(other).shape == sg["ws"].shape and np.isfinite(eng.fetch_states(other)).all()
            cl, ct, tau = eng.fetch_coeffs()
            assert np.array_equal(cl, sg["coeffs"][0]) and np.array_equal(ct, sg["coeffs"][1])
    # the grid form still works on the same context afterwards
    r = eng.plan_grid(H.inputs_for(probs[0]), probs[0]["t"], probs[0]["lon"], probs[0]["d"])
    assert r.winner == singles[0]["res"].winner or draw
    eng.close()


def test_cycle_launch_escalates_to_the_level_that_has_a_winner():
    """levels 1 and 2 without a collision-free candidate, level 3 with one: the cycle launch reports three records and
    chooses level 3 (reactive_planner.py:616-636); with nothing feasible anywhere it reports every level, no winner"""
    from commonroad_rp_b200._lib import traj_len_of
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cyc_DEU_Test-1_1_T-1.npz"))
    import json
    from tests import golden_io
    meta = json.loads(str(z["meta"]))
    ci = next(i for i, c in enumerate(meta["cycles"]) if [lv["winner"] >= 0 for lv in c["levels"]] == [False, False, True])
    from commonroad_rp_b200.sampling import FixedIntervalSampling, VelocitySampling
    from commonroad_rp_b200.utility.config import ReactivePlannerConfiguration
    cfg = ReactivePlannerConfiguration()
    cfg.planning.time_steps_computation = meta["N"]
    cfg.planning.planning_horizon = meta["N"] * meta["dt"]
    cfg.sampling.t_min = meta["t_min"]
    fs = FixedIntervalSampling(cfg)
    x = z["c%d_x0" % ci]
    v0 = float(x[3])
    lo = max(0, v0 - 0.125 * meta["N"] * meta["dt"] * O.vehicle_dict()["a_max"])
    fs.samples_v = VelocitySampling(lo, max(lo + 5.0, v0 + 2), 4)
    x0_lon, x0_lat = z["c%d_x0_lon" % ci], z["c%d_x0_lat" % ci]
    grids = [fs.sample_grid(lv, x0_lat, "velocity_keeping") for lv in (1, 2, 3)]
    prob = {"t": grids[0][0], "lon": grids[0][1], "d": grids[0][2], "x0_lon": x0_lon, "x0_lat": x0_lat, "x0_orientation": float(x[2]),
            "x0_time_step": int(x[7]), "lon_mode": "velocity_keeping", "low_vel_mode": False, "dt": meta["dt"], "N": meta["N"],
            "factor": 1, "draw_all": True, "constraints": O.CONSTRAINTS,
            "cost": {"kind": "default", "desired_speed": meta["desired_velocity"], "desired_s": None, "desired_d": 0.0, "w_a": 5},
            "vehicle": O.vehicle_dict(), "ref": {k: z[k] for k in ("ref_pos", "ref_theta", "ref_curv", "ref_curv_d")},
            "ccosy": {"path": z["cc_path"], "S": z["cc_S"], "normals": z["cc_normals"], "limit": 20.0},
            "obstacles": golden_io.unpack_obstacles(z, "ob_")}
    eng = H.engine_for(prob)
    levels = [(t, lon, d, np.array([traj_len_of(q, meta["dt"]) for q in t], dtype=np.int32)) for t, lon, d in grids]
    records, chosen = eng.plan_levels(H.inputs_for(prob, want_all_states=True), levels)
    assert chosen == 2
    for j, lv in enumerate(meta["cycles"][ci]["levels"]):
        assert (records[j].n_candidates, records[j].winner, records[j].n_infeasible_kinematics, records[j].n_infeasible_collision) == \
            (lv["n"], lv["winner"], lv["n_inf_kin"], lv["n_inf_col"])
    ws = z["c%d_l2_winner_states" % ci]
    assert H.rel_err(ws, eng.fetch_states(records[2].winner)) < 1e-9
    # nothing feasible anywhere: all records, last level chosen, no winner
    blocked = dict(prob)
    ob = dict(prob["obstacles"])
    j0 = int(np.argmax(prob["ref"]["ref_pos"] > x0_lon[0]))
    wall = [[prob["ccosy"]["path"][j0][0], prob["ccosy"]["path"][j0][1], prob["ref"]["ref_theta"][j0] + np.pi / 2, 30.0, 6.0]]
    ob["static_boxes"] = np.vstack([np.asarray(ob["static_boxes"]).reshape(-1, 5), wall])
    blocked["obstacles"] = ob
    eng2 = H.engine_for(blocked)
    records, chosen = eng2.plan_levels(H.inputs_for(blocked), levels)
    assert chosen == 2 and all(records[j].winner == -1 for j in range(3))
    assert [records[j].n_candidates for j in range(3)] == [len(t) * len(lon) * len(d) for t, lon, d in grids]
    eng.close()
    eng2.close()


@pytest.mark.parametrize("kernel", ["candidate_major", "step_parallel"])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_lon_interleaved_shards_merge_to_the_unsharded_result(kernel, world):
    """rp_set_candidate_stripe: rank r owns the lon samples r, r + world, ... of every sampled t (shards with the same mix
    of horizons); the shards' verdicts are those of the unsharded bundle and merge to the same winner / counters"""
    from commonroad_rp_b200 import _lib
    prob = _bundle(seed=3, level=3, N=30, s_dot0=11.0)
    eng = H.engine_for(prob)
    eng.set_kernel_policy(_lib.KERNEL_CANDIDATE_MAJOR if kernel == "candidate_major" else _lib.KERNEL_STEP_PARALLEL)
    inputs = H.inputs_for(prob, check_collision=_lib.COLLISION_ALL)
    full = eng.plan_grid(inputs, prob["t"], prob["lon"], prob["d"])
    cost_f, status_f, reason_f, step_f = eng.fetch_candidates()
    n_t, n_lon, n_d = len(prob["t"]), len(prob["lon"]), len(prob["d"])
    seen = np.zeros(full.n_candidates, dtype=int)
    recs = []
    for rank in range(world):
        eng.set_candidate_stripe(rank, world)
        r = eng.plan_grid(inputs, prob["t"], prob["lon"], prob["d"])
        cost, status, reason, step = eng.fetch_candidates()
        own = np.zeros((n_t, n_lon, n_d), dtype=bool)
        own[:, rank::world, :] = True
        own = own.ravel()
        assert r.n_candidates == int(own.sum())
        assert np.array_equal(status[own], status_f[own]) and np.array_equal(step[own], step_f[own])
        assert np.array_equal(cost[own].view(np.int64), cost_f[own].view(np.int64))
        assert r.winner < 0 or own[r.winner]
        seen += own
        recs.append(r)
    assert np.all(seen == 1)
    eng.set_candidate_stripe(0, 1)
    best = min((r for r in recs if r.winner >= 0), key=lambda r: (r.winner_cost, r.winner))
    assert best.winner == full.winner and best.winner_cost == full.winner_cost
    assert sum(r.n_infeasible_kinematics for r in recs) == full.n_infeasible_kinematics
    assert sum(r.n_collision_total for r in recs) == full.n_collision_total
    assert sum(r.n_feasible for r in recs) == full.n_feasible
    again = eng.plan_grid(inputs, prob["t"], prob["lon"], prob["d"])
    assert again.winner == full.winner and again.n_candidates == full.n_candidates
    eng.close()


@pytest.mark.parametrize("shard", ["none", "range", "stripe"])
def test_deferred_lazy_collision_check_equals_full_checking(shard):
    """candidate-major schedule, check_collision = 2: ego boxes stored by the march, collision verdicts by the deferred
    tile-parallel checker.  Against full checking (every flag, checked while marching): same winner, cost and lazy
    collision count; identical verdicts for everything ranked up to the winner; the rest is either unchecked or identical.
    Cases: dynamic + static obstacles, road boundary as triangles, low-velocity mode, stopping mode (filtered goals), a
    bundle where everything collides (no bound: the second pass checks all), one where nothing does; shards by range and
    by lon-interleaved stripes"""
    from commonroad_rp_b200 import _lib
    cases = [dict(seed=2, level=3, N=60, s_dot0=12.0), dict(seed=1, level=3, N=20, d0=-0.4),
             dict(seed=4, level=3, N=30, s_dot0=9.0, boundary_kind="tris", boundary_offset=3.4), dict(seed=3, level=3, N=20, s_dot0=3.0, low_vel=True),
             dict(seed=6, level=3, N=30, s_dot0=8.0, lon_mode="stopping", desired_s=24.0), dict(seed=5, level=2, N=33, s_dot0=14.0)]
    n_unchecked = 0
    for case in cases:
        prob = _bundle(**case)
        variants = [("as is", prob)]
        if shard == "none":
            variants.append(("wall", _bundle(wall=True, **case)))
        for tag, pr in variants:
            eng = H.engine_for(pr)
            eng.set_kernel_policy(_lib.KERNEL_CANDIDATE_MAJOR)
            n_lon = len(pr["lon"])
            shards = {"none": [None], "range": [(0, 0.4), (0.4, 1.0)], "stripe": [(0, 2), (1, 2)] if n_lon >= 2 else [None]}[shard]
            for sh in shards:
                n_all = len(pr["t"]) * n_lon * len(pr["d"])
                own = np.ones(n_all, dtype=bool)
                if shard == "range" and sh is not None:
                    first = int(sh[0] * n_all) // 32 * 32
                    eng.set_candidate_range(first, int(sh[1] * n_all) - first)
                    own[:] = False
                    own[first:int(sh[1] * n_all)] = True
                elif shard == "stripe" and sh is not None:
                    eng.set_candidate_stripe(*sh)
                    own = np.zeros((len(pr["t"]), n_lon, len(pr["d"])), dtype=bool)
                    own[:, sh[0]::sh[1], :] = True
                    own = own.ravel()
                full = eng.plan_grid(H.inputs_for(pr, check_collision=_lib.COLLISION_ALL), pr["t"], pr["lon"], pr["d"])
                cost_f, status_f, reason_f, step_f = eng.fetch_candidates()
                for rep in range(2):
                    lazy = eng.plan_grid(H.inputs_for(pr, check_collision=_lib.COLLISION_LAZY), pr["t"], pr["lon"], pr["d"])
                    assert eng.last_main_kernel() == _lib.KERNEL_CANDIDATE_MAJOR
                    cost_l, status_l, reason_l, step_l = eng.fetch_candidates()
                    what = (case, tag, sh, rep)
                    assert lazy.winner == full.winner and (lazy.winner_cost == full.winner_cost or lazy.winner < 0), what
                    assert lazy.n_infeasible_collision == full.n_infeasible_collision, what
                    assert lazy.n_infeasible_kinematics == full.n_infeasible_kinematics and lazy.n_feasible == full.n_feasible, what
                    assert list(lazy.reason_counts) == list(full.reason_counts), what
                    n = len(cost_f)
                    idx = np.arange(n)
                    before = ((cost_f < full.winner_cost) | ((cost_f == full.winner_cost) & (idx <= full.winner))) if full.winner >= 0 \
                        else np.isin(status_f, (0, 2))
                    assert lazy.n_candidates == int(own.sum()), what
                    unchecked = (status_l == _lib.ST_UNCHECKED) & own
                    assert not (unchecked & before).any(), what
                    assert np.all(np.isin(status_f[unchecked], (0, 2))), what
                    same = ~unchecked & own
                    assert np.array_equal(status_f[same], status_l[same]) and np.array_equal(step_f[same], step_l[same]), what
                    assert np.array_equal(reason_f[same], reason_l[same]), what
                    assert np.array_equal(cost_f[own].view(np.int64), cost_l[own].view(np.int64)), what
                    n_unchecked += int(unchecked.sum())
                    if tag == "wall":
                        assert lazy.winner < 0 and not unchecked.any(), what
            eng.close()
    assert n_unchecked > 0


def test_repeated_sample_arrays_are_read_afresh():
    """Engine.plan_grid keeps the pointer objects of sample arrays it is handed again (host-side saving): the CONTENTS are
    still read at every call, other calls in between do not disturb it, and new array objects are picked up"""
    from commonroad_rp_b200 import _lib
    from commonroad_rp_b200._lib import traj_len_of
    prob = _bundle(seed=1, level=2, N=20)
    other = _bundle(seed=1, level=2, N=20, d0=-0.35)
    eng = H.engine_for(prob)
    inputs = H.inputs_for(prob)
    t, lon = np.array(prob["t"], dtype=np.float64), np.array(prob["lon"], dtype=np.float64)
    d = np.array(prob["d"], dtype=np.float64)
    d_other = np.array(other["d"], dtype=np.float64)
    assert len(d_other) == len(d) and not np.array_equal(d_other, d)
    tl = np.array([traj_len_of(x, prob["dt"]) for x in t], dtype=np.int32)
    a = eng.plan_grid(inputs, t, lon, d, tl)
    cost_a = eng.fetch_candidates()[0]
    again = eng.plan_grid(inputs, t, lon, d, tl)                  # same objects: kept pointers
    assert again.winner == a.winner and np.array_equal(eng.fetch_candidates()[0], cost_a, equal_nan=True)
    fresh = eng.plan_grid(inputs, t, lon, d_other.copy(), tl)     # a new array object
    cost_fresh = eng.fetch_candidates()[0]
    assert not np.array_equal(cost_fresh, cost_a, equal_nan=True)
    eng.plan_grid(inputs, t, lon, d, tl)
    # a cycle launch in between changes the engine's candidate count; the kept arguments restore it
    eng.plan_levels(inputs, [(t[:2], lon[:2], d[:3], tl[:2])])
    back = eng.plan_grid(inputs, t, lon, d, tl)
    assert back.winner == a.winner and len(eng.fetch_candidates()[0]) == len(cost_a)
    d[:] = d_other                                                # the SAME object with new contents
    changed = eng.plan_grid(inputs, t, lon, d, tl)
    assert changed.winner == fresh.winner and np.array_equal(eng.fetch_candidates()[0], cost_fresh, equal_nan=True)
    st = eng.fetch_states(changed.winner)
    st2 = eng.fetch_states(changed.winner)
    assert st is not st2 and np.array_equal(st, st2)               # (fetch_states hands out copies of its kept buffer)
    eng.close()
