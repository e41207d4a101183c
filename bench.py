#!/usr/bin/env python
"""bench.py -- candidate-trajectory throughput of the B200 engine (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host cores

Workload (config.workload): BASELINE.json configs[3], the synthetic dense sampling sweep
64 d x 64 v x 32 t end states over a 60-step horizon (131 072 candidates x 61 time steps per
replanning cycle), SURVEY.md section 8d.  One "step" = one replanning cycle over that bundle:
coefficient solve, fused evaluation (kinematics, projection, cost, collision), feasible arg-min and
the winner's state block.

  value      candidates/s with the sample lists already resident in HBM (K x rp_grid_launch,
             device time by CUDA events on the launching stream, L2 flushed between iterations)
  e2e        the same through the public host-buffer call (rp_plan_grid + rp_fetch_states): H2D of the
             cycle's inputs and D2H of the result + winner states inside the timed region
  roofline   the fused kernel against the MEASURED FP64-FMA peak (this path is FP64-pipe bound, not
             HBM or tensor bound -- SURVEY 8d); algorithmic work = 200 flop per candidate-timestep
  N > 1      weak scaling: the bundle grows with N (64*N velocity samples) and is sharded over the ranks by interleaving
             the lon samples (every rank sees the same mix of horizons); the shard records are exchanged by two kernels
             storing into peer-mapped mailboxes over NVLink (rp_peer_*; `--exchange nccl` keeps the NCCL form).  The same
             line carries `strong_scaling` (the FIXED 64 x 64 x 32 bundle over the N ranks) and `scenario_batch`
             (BASELINE configs[4]: independent scenarios, scenario-major shards, 512 per rank at N = 8)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_CAND_STEP = 200.0      # SURVEY.md section 8d "ALGORITHMIC work per unit"
SELECT_BYTES_PER_CAND = 13 * 8 + 16          # select-only: 13 doubles in, (cost f64, info i32, padding) out per candidate
STATE_BYTES_PER_CAND_STEP = 112.0
METRIC = "candidate_trajectories_per_sec"
UNIT = "candidates/s"
N_HORIZON = 60
DT = 0.1


# --------------------------------------------------------------------------------------------------
def dense_workload(n_gpus=1, seed=0):
    """Scenario + ordered sample lists of the dense sweep.  The enumeration order is the host's
    Python-set order, as the reference iterates it (sampling.py:218-226)."""
    from commonroad_rp_b200.utility import synthetic
    from commonroad_rp_b200.utility.utils_coordinate_system import CoordinateSystem
    scn = synthetic.make_scenario(seed=seed)
    t, v, d, d0 = synthetic.dense_grid(n_v=64 * n_gpus)
    cosy = CoordinateSystem(scn["ref_path"])
    s0 = float(cosy.ref_pos[10])
    j = int(np.argmax(cosy.ref_pos > s0)) - 1
    work = {
        "scn": scn, "cosy": cosy,
        "t": [float(x) for x in set(t)], "lon": [float(x) for x in set(v)],
        "d": [float(x) for x in set(d).union({d0})],
        "x0_lon": [s0, 15.0, 0.0], "x0_lat": [d0, 0.0, 0.0], "x0_orientation": float(cosy.ref_theta[j]),
        "desired_speed": 15.0,
    }
    work["n_cand"] = len(work["t"]) * len(work["lon"]) * len(work["d"])
    return work


def obstacle_arrays(scn):
    sb = np.asarray(scn["static_boxes"], dtype=np.float64).reshape(-1, 5)
    static = np.stack([sb[:, 0], sb[:, 1], sb[:, 2], 0.5 * sb[:, 3], 0.5 * sb[:, 4]], axis=1)
    static = np.concatenate([static, np.asarray(scn["boundary_boxes"], dtype=np.float64).reshape(-1, 5)], axis=0)
    dyn = []
    for st, lw in zip(scn["dyn_states"], scn["dyn_lw"]):
        st = np.asarray(st, dtype=np.float64).reshape(-1, 3)
        dyn.append(np.concatenate([st, np.full((len(st), 1), 0.5 * lw[0]), np.full((len(st), 1), 0.5 * lw[1])], axis=1))
    return static, np.asarray(scn["dyn_t0"], dtype=np.int32), dyn


def make_engine(work, device, stream):
    from commonroad_rp_b200._lib import Engine
    from commonroad_rp_b200.utility.config import VehicleConfiguration
    veh = VehicleConfiguration()
    eng = Engine(device, stream)
    eng.set_vehicle(veh.length, veh.width, veh.wb_rear_axle, veh.wheelbase, veh.a_max, veh.v_switch, veh.delta_max,
                    veh.v_delta_max, veh.kappa_max)
    tb = work["cosy"].device_tables()
    eng.set_reference(tb["ref_pos"], tb["ref_theta"], tb["ref_curv"], tb["ref_curv_d"], tb["path_xy"], tb["path_s"],
                      tb["path_normals"], tb["proj_limit"])
    static, t0, dyn = obstacle_arrays(work["scn"])
    eng.set_obstacles(static, t0, dyn)
    return eng


def make_inputs(work, want_all_states=False, check_collision=2):
    from commonroad_rp_b200._lib import Engine
    return Engine.make_inputs(work["x0_lon"], work["x0_lat"], work["x0_orientation"], 0, False, "velocity_keeping",
                              N_HORIZON, DT, desired_speed=work["desired_speed"], desired_d=0.0, w_a=5.0,
                              want_all_states=want_all_states, check_collision=check_collision)


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for row in self.rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# --------------------------------------------------------------------------------------------------
def _cyclic_fixture(name="ZAM_Over-1_1"):
    """Replanning cycles of a bundled reference scenario (inputs recorded by oracle/make_golden.py)."""
    from tests import golden_io
    z = np.load(os.path.join(ROOT, "tests", "golden", "cyc_%s.npz" % name))
    meta = json.loads(str(z["meta"]))
    return z, meta, golden_io.unpack_obstacles(z, "ob_")


def replanning_latency_b200(name="ZAM_Over-1_1", repeats=5, speculation=None):
    """p50 wall time of ReactivePlanner.plan() (this repo's drop-in API) over the scenario's replanning cycles.
    ``speculation``: the planner's policy for evaluating the following sampling levels in the same launch (None: default)."""
    from commonroad_rp_b200 import collision
    from commonroad_rp_b200.reactive_planner import ReactivePlanner
    from commonroad_rp_b200.state import ReactivePlannerState
    from commonroad_rp_b200.utility.config import ReactivePlannerConfiguration
    from commonroad_rp_b200.utility.utils_coordinate_system import CoordinateSystem
    z, meta, ob = _cyclic_fixture(name)
    cfg = ReactivePlannerConfiguration()
    cfg.planning.time_steps_computation = meta["N"]
    cfg.planning.low_vel_mode_threshold = meta["low_vel_mode_threshold"]
    cfg.sampling.t_min = meta["t_min"]
    cfg.debug.draw_traj_set = meta.get("draw_traj_set", False)       # DEU_Test's YAML keeps every trajectory for plotting
    cfg.debug.save_plots = meta.get("draw_traj_set", False)

    class _Empty:
        static_obstacles, dynamic_obstacles = (), ()
        lanelet_network = type("LN", (), {"lanelets": ()})()

    cfg.update(scenario=_Empty(), planning_problem=None)
    planner = ReactivePlanner(cfg)
    if speculation is not None:
        planner.speculation = speculation
    cc = collision.checker_from_arrays(**ob)
    co = CoordinateSystem(z["ref_path_raw"])
    times = []
    for rep in range(repeats + 1):
        for ci in range(meta["n_cycles"]):
            x = z["c%d_x0" % ci]
            x0 = ReactivePlannerState(time_step=int(x[7]), position=np.array([x[0], x[1]]), orientation=x[2],
                                      velocity=x[3], acceleration=x[4], yaw_rate=x[5], steering_angle=x[6])
            planner.reset(initial_state_cart=x0, initial_state_curv=(list(z["c%d_x0_lon" % ci]), list(z["c%d_x0_lat" % ci])),
                          collision_checker=cc, coordinate_system=co)
            planner.set_desired_velocity(desired_velocity=meta["desired_velocity"] if ci == 0 else None,
                                         current_speed=x0.velocity)
            t0 = time.perf_counter()
            out = planner.plan()
            dt_s = time.perf_counter() - t0
            assert (out is not None) == meta["cycles"][ci]["ok"]
            if rep > 0:
                times.append(dt_s)
    t_ms = np.array(times) * 1e3
    return {"p50_ms": float(np.percentile(t_ms, 50)), "p95_ms": float(np.percentile(t_ms, 95)), "cycles": len(t_ms),
            "scenario": "%s, N=%d, cyclic replanning inputs of the reference run (tests/golden)" % (name, meta["N"]),
            "api": "ReactivePlanner.plan()", "speculation": planner.speculation}


def replanning_latency_port(name="ZAM_Over-1_1", max_cycles=6):
    """The same cycles through the oracle port (level escalation as reactive_planner.py:616-636), single process."""
    from oracle import rp_oracle as O
    from commonroad_rp_b200.sampling import FixedIntervalSampling, VelocitySampling
    from commonroad_rp_b200.utility.config import ReactivePlannerConfiguration
    z, meta, ob = _cyclic_fixture(name)
    cfg = ReactivePlannerConfiguration()
    cfg.planning.time_steps_computation = meta["N"]
    cfg.sampling.t_min = meta["t_min"]
    fs = FixedIntervalSampling(cfg)
    veh = cfg.vehicle
    horizon = meta["N"] * meta["dt"]
    times = []
    for ci in range(min(max_cycles, meta["n_cycles"])):
        x = z["c%d_x0" % ci]
        v0 = float(x[3])
        lo = max(0, v0 - 0.125 * horizon * veh.a_max)
        fs.samples_v = VelocitySampling(lo, max(lo + 5.0, v0 + 2), 4)
        x0_lon, x0_lat = z["c%d_x0_lon" % ci], z["c%d_x0_lat" % ci]
        t0 = time.perf_counter()
        for level in range(1, 4):
            t, lon, d = fs.sample_grid(level, x0_lat, "velocity_keeping")
            prob = {"t": t, "lon": lon, "d": d, "x0_lon": x0_lon, "x0_lat": x0_lat, "x0_orientation": float(x[2]),
                    "x0_time_step": int(x[7]), "lon_mode": "velocity_keeping",
                    "low_vel_mode": bool(v0 < meta["low_vel_mode_threshold"]), "dt": meta["dt"], "N": meta["N"],
                    "factor": 1, "draw_all": False, "constraints": O.CONSTRAINTS,
                    "cost": {"kind": "default", "desired_speed": meta["desired_velocity"], "desired_s": None,
                             "desired_d": 0.0, "w_a": 5},
                    "vehicle": {"length": veh.length, "width": veh.width, "wb_rear_axle": veh.wb_rear_axle,
                                "wheelbase": veh.wheelbase, "a_max": veh.a_max, "v_switch": veh.v_switch,
                                "delta_max": veh.delta_max, "v_delta_max": veh.v_delta_max},
                    "ref": {k: z[k] for k in ("ref_pos", "ref_theta", "ref_curv", "ref_curv_d")},
                    "ccosy": {"path": z["cc_path"], "S": z["cc_S"], "normals": z["cc_normals"], "limit": 20.0},
                    "obstacles": ob}
            if O.plan_grid(prob, want_states=False, full_collision=False)["winner"] >= 0:
                break
        times.append(time.perf_counter() - t0)
    t_ms = np.array(times) * 1e3
    return {"p50_ms": float(np.percentile(t_ms, 50)), "p95_ms": float(np.percentile(t_ms, 95)), "cycles": len(t_ms),
            "scenario": "%s, N=%d" % (name, meta["N"]), "api": "oracle port, single process"}


def scenario_batch_rate(device, stream_handle, n_scenarios=64, cycles=3, rank=0, world=1):
    """BASELINE configs[4] shape on one rank: independent seeded scenarios, default level-3 grid at N = 60
    (29 t x 17 v x 18 d = 8 874 candidates each), one resident device context per scenario, launches
    enqueued back to back (commonroad_rp_b200.parallel.ScenarioBatch).  Returns candidates/s."""
    import torch
    from commonroad_rp_b200 import collision
    from commonroad_rp_b200._lib import Engine
    from commonroad_rp_b200.parallel import ScenarioBatch
    from commonroad_rp_b200.sampling import FixedIntervalSampling, VelocitySampling
    from commonroad_rp_b200.utility import synthetic
    from commonroad_rp_b200.utility.config import ReactivePlannerConfiguration
    from commonroad_rp_b200.utility.utils_coordinate_system import CoordinateSystem
    cfg = ReactivePlannerConfiguration()
    keys = ("static_boxes", "dyn_t0", "dyn_states", "dyn_lw", "boundary_boxes", "boundary_tris")
    batch = ScenarioBatch(device, stream_handle)
    cycle, n_cand = [], 0
    samplers, x0_cart, x0_curv = [], [], []
    for sid in range(rank * n_scenarios, (rank + 1) * n_scenarios):       # scenario-major shards: no exchange on the data path
        scn, s_dot0, d0 = synthetic.scenario_seeded(sid)
        co = CoordinateSystem(scn["ref_path"])
        batch.add_scenario(cfg.vehicle, co, collision.checker_from_arrays(**{k: scn[k] for k in keys}))
        fs = FixedIntervalSampling(cfg)
        samplers.append(fs)
        lo = max(0, s_dot0 - 0.125 * fs.horizon * cfg.vehicle.a_max)
        fs.samples_v = VelocitySampling(lo, max(lo + 5.0, s_dot0 + 2), 4)
        t, lon, d = fs.sample_grid(3, [d0, 0.0, 0.0], "velocity_keeping")
        s0 = float(co.ref_pos[10])
        j = int(np.argmax(co.ref_pos > s0)) - 1
        # (check_collision = 2: the reference's lazy collision pass, as in the headline)
        inputs = Engine.make_inputs([s0, s_dot0, 0.0], [d0, 0.0, 0.0], float(co.ref_theta[j]), 0, s_dot0 < 4.0,
                                    "velocity_keeping", N_HORIZON, DT, desired_speed=s_dot0, check_collision=2)
        xy = co.convert_to_cartesian_coords(s0, d0)
        x0_cart.append([xy[0], xy[1], float(co.ref_theta[j]), s_dot0, 0.0, 0.0])
        x0_curv.append(([s0, s_dot0, 0.0], [d0, 0.0, 0.0]))
        from commonroad_rp_b200._lib import traj_len_of
        cycle.append((inputs, np.asarray(t, dtype=np.float64), np.asarray(lon, dtype=np.float64), np.asarray(d, dtype=np.float64),
                      np.asarray([traj_len_of(x, DT) for x in t], dtype=np.int32)))
        n_cand += len(t) * len(lon) * len(d)
    from commonroad_rp_b200._lib import Batch
    packed = Batch.pack(cycle)                          # the host buffers of a cycle: PlanInputs array + concatenated sample lists
    for _ in range(2):
        batch.plan(packed)                              # warm-up (allocations, geometry)
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    t0 = time.perf_counter()
    dev_ms = []
    for _ in range(cycles):
        res = batch.plan(packed)
        dev_ms.append(batch.batch.last_ms()[0])
    torch.cuda.synchronize()
    dt_s = (time.perf_counter() - t0) / cycles
    n_win = sum(1 for r in res if r.winner >= 0)
    n_local = n_cand
    if world > 1:                                       # whole job: all ranks' scenarios / slowest rank
        agg = torch.tensor([dt_s, float(np.mean(dev_ms))], dtype=torch.float64, device="cuda:%d" % device)
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
        cnt = torch.tensor([float(n_cand), float(n_win)], dtype=torch.float64, device="cuda:%d" % device)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dt_s, dev_ms = float(agg[0].item()), [float(agg[1].item())]
        n_cand, n_win, n_scenarios = int(cnt[0].item()), int(cnt[1].item()), n_scenarios * world
    # ---- closed loop (run_planner.py:61-107 for all scenarios at once): plan, advance every scenario three steps along its
    # winner (rp_batch_winner_states, one launch), re-derive the velocity samples from the new speed (:332-335), next cycle
    closed = None
    if n_win > 0:
        batch.reset(np.array(x0_cart), states_curv=(np.array([c[0] for c in x0_curv]), np.array([c[1] for c in x0_curv])))
        a_max, horizon = cfg.vehicle.a_max, N_HORIZON * DT
        host_s, wall_s, dev2, steps_per_cycle, closed_cycles = 0.0, 0.0, [], 3, 5
        time_step = 0
        cur = list(cycle)
        for c in range(closed_cycles + 1):
            t_a = time.perf_counter()
            if c > 0:
                time_step += steps_per_cycle
                cur = []
                for k, fs in enumerate(samplers):
                    v = float(batch.x0_cart[k, 3])
                    lo = max(0, v - 0.125 * horizon * a_max)
                    fs.samples_v = VelocitySampling(lo, max(lo + 5.0, v + 2), 4)
                    t, lon, d = fs.sample_grid(3, batch.x0_lat[k], "velocity_keeping")
                    inp = cycle[k][0]
                    inp.x0_lon[:] = batch.x0_lon[k].tolist()
                    inp.x0_lat[:] = batch.x0_lat[k].tolist()
                    inp.x0_orientation = float(batch.x0_cart[k, 2])
                    inp.x0_time_step = time_step
                    inp.low_vel_mode = int(v < 4.0)
                    cur.append((inp, t, lon, d, cycle[k][4]))
            pk = Batch.pack(cur)
            t_b = time.perf_counter()
            res_c = batch.plan(pk)
            batch.advance(steps_per_cycle)
            t_c = time.perf_counter()
            if c > 0:
                host_s += t_b - t_a
                wall_s += t_c - t_a
                dev2.append(batch.batch.last_ms()[0])
        closed = {"cycles": closed_cycles, "steps_between_cycles": steps_per_cycle,
                  "ms_per_cycle_of_one_ranks_scenarios": 1e3 * wall_s / closed_cycles,
                  "host_resampling_ms_per_cycle": 1e3 * host_s / closed_cycles, "device_ms_per_cycle": float(np.mean(dev2)),
                  "scenarios_with_winner_last_cycle": sum(1 for r in res_c if r.winner >= 0),
                  "note": "every cycle starts from the state three steps along the previous winner (one launch for all "
                          "scenarios); the host re-iterates each scenario's Python sample sets (set order = enumeration order)"}
    batch.plan_one_by_one(cycle)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(cycles):
        batch.plan_one_by_one(cycle)
    torch.cuda.synchronize()
    one_s = (time.perf_counter() - t1) / cycles
    # ---- stress point of BASELINE configs[4]: the DENSE grid (32 t x 64 v x (64 d + d0)) on a subset of the
    # scenarios -- 64 per rank at --gpus 8 (the 512-scenario subset), 8 per rank otherwise -- in one launch chain
    dense = None
    n_dense = min(n_scenarios // max(world, 1) if world > 1 else n_scenarios, 64 if world >= 8 else 8)
    if n_dense > 0:
        from commonroad_rp_b200._lib import traj_len_of
        td, vd, dd, _ = synthetic.dense_grid(n_v=64)
        t_arr = np.asarray([float(x) for x in set(td)], dtype=np.float64)
        v_arr = np.asarray([float(x) for x in set(vd)], dtype=np.float64)
        tl_arr = np.asarray([traj_len_of(x, DT) for x in t_arr], dtype=np.int32)
        dense_cycle = []
        for k in range(len(cycle)):
            inp = cycle[k][0]
            d_arr = np.asarray([float(x) for x in set(dd).union({float(inp.x0_lat[0])})], dtype=np.float64)
            if k < n_dense:
                dense_cycle.append((inp, t_arr, v_arr, d_arr, tl_arr))
            else:           # the other scenarios of the batch idle on one candidate
                dense_cycle.append((inp, t_arr[:1], v_arr[:1], d_arr[:1], tl_arr[:1]))
        pk = Batch.pack(dense_cycle)
        n_dense_cand = sum(len(c[1]) * len(c[2]) * len(c[3]) for c in dense_cycle)
        for _ in range(2):
            batch.plan(pk)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t2 = time.perf_counter()
        dms = []
        for _ in range(3):
            res_d = batch.plan(pk)
            dms.append(batch.batch.last_ms()[0])
        torch.cuda.synchronize()
        dd_s = (time.perf_counter() - t2) / 3
        n_dw = sum(1 for r in res_d[:n_dense] if r.winner >= 0)
        if world > 1:
            agg = torch.tensor([dd_s, float(np.mean(dms))], dtype=torch.float64, device="cuda:%d" % device)
            dist.all_reduce(agg, op=dist.ReduceOp.MAX)
            cnt = torch.tensor([float(n_dense_cand), float(n_dw)], dtype=torch.float64, device="cuda:%d" % device)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
            dd_s, dms, n_dense_cand, n_dw = float(agg[0].item()), [float(agg[1].item())], int(cnt[0].item()), int(cnt[1].item())
        dense = {"value": n_dense_cand / dd_s, "unit": UNIT, "dense_scenarios": n_dense * max(world, 1),
                 "candidates_per_dense_scenario": int(len(t_arr) * len(v_arr) * len(dense_cycle[0][3])),
                 "ms_per_cycle_of_all_scenarios": 1e3 * dd_s, "device_ms_per_cycle": float(np.mean(dms)),
                 "device_value": n_dense_cand / (float(np.mean(dms)) * 1e-3), "dense_scenarios_with_winner": n_dw,
                 "note": "the dense 32 t x 64 v x (64 d + d0) grid at N = 60 on every one of these scenarios, all in one "
                         "launch chain (rp_batch_*); 3 timed cycles, wall clock incl. host staging and result D2H"}
    batch.close()
    return {"value": n_cand / dt_s, "unit": UNIT, "scenarios": n_scenarios, "candidates_per_scenario": n_cand // n_scenarios,
            "ms_per_cycle_of_all_scenarios": 1e3 * dt_s, "device_ms_per_cycle": float(np.mean(dev_ms)),
            "device_value": n_cand / (float(np.mean(dev_ms)) * 1e-3), "scenarios_with_winner": n_win,
            "one_launch_chain_per_scenario": {"value": n_local / one_s, "ms_per_cycle_of_one_ranks_scenarios": 1e3 * one_s},
            "timed_cycles": cycles, "closed_loop": closed, "dense_grid_stress": dense,
            "workload": "BASELINE configs[4]: %d independent seeded scenarios (%d per rank), default level-3 grid at N = 60"
                        % (n_scenarios, n_scenarios // max(world, 1)),
            "note": "rp_batch_*: one H2D, seven launches (coefficients, obstacle rows, march, deferred collision check x 2 + "
                    "gather, selection), one D2H per cycle of all scenarios; wall clock incl. host "
                    "staging of every scenario's inputs (host buffers) and D2H of every result"}


def ncu_record(kernel_name):
    """Per-launch DRAM traffic and FP64-pipe activity of ``kernel_name`` from the committed ncu capture of this
    same command (profiles/ncu_summary.json, written from `ncu --set full`); {} if there is none."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_summary.json")
    try:
        with open(path) as fh:
            return json.load(fh).get(kernel_name, {})
    except (OSError, ValueError):
        return {}


def cpu_port_rate(work, stride, workers, repeats=1):
    """The oracle port (reference algorithm restated, oracle/rp_oracle.py) on a sub-grid of the same
    workload: every ``stride``-th v and d sample.  Returns (candidates/s, description, seconds)."""
    from oracle import rp_oracle as O
    from commonroad_rp_b200.utility.config import VehicleConfiguration
    veh = VehicleConfiguration()
    tb = work["cosy"].device_tables()
    prob = {
        "t": np.array(work["t"]), "lon": np.array(work["lon"][::stride]), "d": np.array(work["d"][::stride]),
        "x0_lon": np.array(work["x0_lon"]), "x0_lat": np.array(work["x0_lat"]),
        "x0_orientation": work["x0_orientation"], "x0_time_step": 0, "lon_mode": "velocity_keeping",
        "low_vel_mode": False, "dt": DT, "N": N_HORIZON, "factor": 1, "draw_all": False,
        "constraints": O.CONSTRAINTS,
        "cost": {"kind": "default", "desired_speed": work["desired_speed"], "desired_s": None, "desired_d": 0.0, "w_a": 5},
        "vehicle": {"length": veh.length, "width": veh.width, "wb_rear_axle": veh.wb_rear_axle,
                    "wheelbase": veh.wheelbase, "a_max": veh.a_max, "v_switch": veh.v_switch,
                    "delta_max": veh.delta_max, "v_delta_max": veh.v_delta_max},
        "ref": {"ref_pos": tb["ref_pos"], "ref_theta": tb["ref_theta"], "ref_curv": tb["ref_curv"],
                "ref_curv_d": tb["ref_curv_d"]},
        "ccosy": {"path": tb["path_xy"], "S": tb["path_s"], "normals": tb["path_normals"], "limit": tb["proj_limit"]},
        "obstacles": {k: work["scn"][k] for k in ("static_boxes", "dyn_t0", "dyn_states", "dyn_lw", "boundary_boxes",
                                                  "boundary_tris")},
    }
    n = len(prob["t"]) * len(prob["lon"]) * len(prob["d"])
    t0 = time.perf_counter()
    for _ in range(repeats):
        if workers <= 1:
            O.plan_grid(prob, want_states=False, full_collision=False)
        else:
            O.plan_grid_parallel(prob, workers)
    dt_s = (time.perf_counter() - t0) / repeats
    desc = "every %d-th v and d sample of the dense sweep: %d candidates x %d steps per cycle" % (stride, n, N_HORIZON + 1)
    return n / dt_s, desc, dt_s, n


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference itself is Python and
    cannot travel to the GPU box) with all host cores, fork-parallel like reactive_planner.py:1084-1111."""
    if rank != 0:
        return
    work = dense_workload(1)
    cores = os.cpu_count() or 1
    stride = 8
    # calibrate so that steps + warmup stay within a few minutes
    rate, desc, sec, n = cpu_port_rate(work, 16, cores)
    budget = 150.0 / max(1, args.steps + args.warmup)
    while stride > 1 and (len(work["t"]) * (64 // (stride // 2)) ** 2) / rate < budget:
        stride //= 2
    for _ in range(args.warmup):
        cpu_port_rate(work, stride, cores)
    times, n_s = [], 0
    for _ in range(args.steps):
        rate, desc, sec, n_s = cpu_port_rate(work, stride, cores)
        times.append(sec)
    total = float(np.sum(times))
    value = n_s * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "dense sampling sweep 64 d x 64 v x 32 t, 60-step horizon (BASELINE configs[3])",
                   "sample": desc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cand_timesteps_per_sec": value * (N_HORIZON + 1), "gpu_launches": 0,
        "p50_replanning_cycle_ms": replanning_latency_port(),
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: the arg-min exchange of the sharded bundle -- stores into peer-mapped mailboxes over NVLink "
                         "(rp_peer_*) or an NCCL all-gather + all-reduce")
    ap.add_argument("--full-states", action="store_true", help="also time the full-state (HBM-heavy) variant")
    ap.add_argument("--scenarios", type=int, default=None,
                    help="independent scenarios PER RANK of the scenario-batch leg (default: 512 at --gpus 8 = BASELINE "
                         "configs[4]'s 4 096 scenarios, else 64)")
    args = ap.parse_args()
    warmup_requested = args.warmup
    if args.impl == "b200" and args.warmup < 3:
        print("[bench] --warmup %d raised to 3 (timing rules: at least three warm-up steps)" % args.warmup, file=sys.stderr)
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    work = dense_workload(n_gpus)
    # a dedicated (non-default) stream shared by torch and the engine, so that torch's CUDA events
    # bracket exactly the kernels the C-ABI launches
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng = make_engine(work, local_rank, stream.cuda_stream)
    inputs = make_inputs(work)
    n_total = work["n_cand"]
    if args.scenarios is None:
        args.scenarios = 512 if world == 8 else 64
    # shards: the lon samples interleaved over the ranks (rank r: lon indices r, r + N, ...; every sampled t and d)
    n_lon_r = len(range(rank, len(work["lon"]), world))
    count = len(work["t"]) * n_lon_r * len(work["d"])
    if world > 1:
        eng.set_candidate_stripe(rank, world)
    Np1 = N_HORIZON + 1

    from commonroad_rp_b200.parallel import PeerExchange, global_argmin
    rec = torch.zeros(4, dtype=torch.float64, device=dev)
    peer = None

    def exchange():
        return None if peer is not None else global_argmin(eng, rec, world)

    def step_device():
        eng.grid_launch()
        if world > 1:
            return exchange()
        return None

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2
    eng.grid_upload(inputs, work["t"], work["lon"], work["d"])
    nccl_ref = None
    if world > 1 and args.exchange == "peer":
        # one cycle through the NCCL exchange first (the cross-check of the peer-memory exchange below), then open the
        # peer group: from here on rp_grid_launch itself ends with the exchange (two kernels storing into the peers'
        # mailboxes) and every rank's result block is the global one
        eng.grid_launch()
        w_, t_, b_ = global_argmin(eng, rec, world)
        nccl_ref = w_.tolist() + t_.tolist() + b_.tolist()
        # (a box without CUDA IPC / peer access between its GPUs: every rank falls back to the NCCL form together)
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            peer = PeerExchange(eng)
        except Exception as exc:                                      # noqa: BLE001
            print("[bench] peer-memory exchange unavailable on rank %d: %s" % (rank, exc), file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            if peer is not None:
                peer.close()
            peer = None
    fp64_peak = eng.measure_fp64_peak()
    # clocks are sampled from the warm-up to the end of the e2e loop (the same kernels throughout);
    # nvidia-smi needs a few hundred ms to start, so keep the GPU under this load until it reports
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    # nvidia-smi needs a few hundred ms to report: keep every GPU under this load until rank 0 has three samples (all
    # ranks run the same cycles -- the exchange is a lockstep protocol -- rank 0 broadcasts when to stop)
    t_wait = time.perf_counter()
    while True:
        for _ in range(50):
            step_device()
        torch.cuda.synchronize()
        done = torch.tensor([1 if (len(sampler.rows) >= 3 or time.perf_counter() - t_wait > 3.0) else 0], device=dev)
        if world > 1:
            dist.broadcast(done, src=0)
        if int(done.item()):
            break
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.fill_(k & 0xFF)                     # L2 flush between timed iterations (outside the event pair)
        starts[k].record(stream)
        step_device()
        stops[k].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
    # per-kernel durations (stage_ms, roofline.kernel_ms): the same K steps once more with the library's stage events
    # on -- five CUDA event records per launch, which cost the cycle ~15 us and are therefore off (the library's
    # default) in the timed region above
    eng.set_stage_timing(True)
    ev_starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev_stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        ev_starts[k].record(stream)
        step_device()
        ev_stops[k].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    step_ms_events = [s.elapsed_time(e) for s, e in zip(ev_starts, ev_stops)]
    fused_ms = [eng.stage_ms(back)[1] for back in range(min(args.steps, 64))]
    stage_ms = np.mean([eng.stage_ms(back) for back in range(min(args.steps, 64))], axis=0)
    eng.set_stage_timing(False)
    total_ms = torch.tensor([float(np.sum(step_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    res = eng.grid_result()
    main_kernel = eng.last_main_kernel()
    launches_main = eng.launches_per_plan()
    exchange_check = None
    if peer is not None:
        # the peer-memory exchange against the NCCL one on the same bundle: identical winner / totals / count
        got = [res.winner_cost, float(res.winner), float(res.n_infeasible_kinematics), float(res.n_feasible),
               float(res.n_infeasible_collision)]
        exchange_check = got == nccl_ref
        assert exchange_check, "peer-memory exchange differs from the NCCL exchange: %r vs %r" % (got, nccl_ref)

    # ---- end to end through the host-buffer API ----
    t_np, lon_np, d_np = np.array(work["t"]), np.array(work["lon"]), np.array(work["d"])
    from commonroad_rp_b200._lib import traj_len_of
    tl_np = np.array([traj_len_of(x, DT) for x in t_np], dtype=np.int32)        # an input array of rp_plan_grid
    for _ in range(2):
        r = eng.plan_grid(inputs, t_np, lon_np, d_np, tl_np)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e2e_t0 = time.perf_counter()
    for _ in range(args.steps):
        r = eng.plan_grid(inputs, t_np, lon_np, d_np, tl_np)
        if world > 1:
            exchange()
        if r.winner >= 0:
            eng.fetch_states(r.winner)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - e2e_t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    clocks = sampler.stop() if rank == 0 else None
    import ctypes
    from commonroad_rp_b200 import _lib
    h2d = 8 * (len(work["t"]) + len(work["lon"]) + len(work["d"])) + 4 * len(work["t"]) + ctypes.sizeof(_lib.PlanInputs)
    d2h = ctypes.sizeof(_lib.PlanResult) + 4 + 14 * Np1 * 8

    full = None
    if args.full_states and world == 1:
        fin = make_inputs(work, want_all_states=True)
        eng.set_stage_timing(True)
        eng.grid_upload(fin, work["t"], work["lon"], work["d"])
        for _ in range(3):
            eng.grid_launch()
        torch.cuda.synchronize()
        ms = []
        for k in range(args.steps):
            flush.fill_(k & 0xFF)
            eng.grid_launch()
            torch.cuda.synchronize()
            ms.append(eng.stage_ms(0)[1])
        full = float(np.mean(ms))
        eng.set_stage_timing(False)

    # the headline runs the reference's own lazy collision pass (check_collision = 2, reactive_planner.py:1031-1063: only
    # the candidates ranked before the winner get a collision verdict).  Here the same bundle with a collision flag for
    # EVERY feasible candidate (check_collision = 1): same winner and counters
    lazy_ms = None
    if world == 1:
        lin = make_inputs(work, check_collision=1)
        eng.grid_upload(lin, work["t"], work["lon"], work["d"])
        for _ in range(3):
            eng.grid_launch()
        torch.cuda.synchronize()
        l0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        l1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        for k in range(args.steps):
            flush.fill_(k & 0xFF)
            l0[k].record(stream)
            eng.grid_launch()
            l1[k].record(stream)
        torch.cuda.synchronize()
        lazy_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(l0, l1)]))
        lazy_res = eng.grid_result()
        assert lazy_res.winner == res.winner and lazy_res.n_infeasible_collision == res.n_infeasible_collision

    # ---- strong scaling: the FIXED 64 x 64 x 32 bundle (BASELINE configs[3] itself) over the N ranks -----------------
    strong = None
    if world > 1:
        work1 = dense_workload(1)
        n1 = work1["n_cand"]
        inputs1 = make_inputs(work1)
        eng.grid_upload(inputs1, work1["t"], work1["lon"], work1["d"])

        def timed_cycles(k_steps):
            for _ in range(args.warmup):
                step_device()
            torch.cuda.synchronize()
            dist.barrier()
            a0 = [torch.cuda.Event(enable_timing=True) for _ in range(k_steps)]
            a1 = [torch.cuda.Event(enable_timing=True) for _ in range(k_steps)]
            for k in range(k_steps):
                flush.fill_(k & 0xFF)
                a0[k].record(stream)
                step_device()
                a1[k].record(stream)
            torch.cuda.synchronize()
            dist.barrier()
            tot = torch.tensor([float(np.sum([x.elapsed_time(y) for x, y in zip(a0, a1)]))], dtype=torch.float64, device=dev)
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
            return float(tot.item()) / k_steps

        eng.set_candidate_stripe(rank, world)
        sharded_ms = timed_cycles(args.steps)
        res_s = eng.grid_result()
        kernel_s = eng.last_main_kernel()
        # the same bundle unsharded on every rank (no exchange): the single-GPU time of this very run
        eng.set_candidate_stripe(0, 1)
        single_ms = []
        for _ in range(args.warmup):
            eng.grid_launch()
        torch.cuda.synchronize()
        for k in range(args.steps):
            flush.fill_(k & 0xFF)
            e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e_a.record(stream)
            eng.grid_launch()
            e_b.record(stream)
            torch.cuda.synchronize()
            single_ms.append(e_a.elapsed_time(e_b))
        res_1 = eng.grid_result()
        assert (res_s.winner, res_s.n_infeasible_kinematics, res_s.n_infeasible_collision) == \
            (res_1.winner, res_1.n_infeasible_kinematics, res_1.n_infeasible_collision)
        from commonroad_rp_b200 import _lib as _l0
        strong = {"workload": "dense sampling sweep 64 d x 64 v x 32 t, 60-step horizon: the FIXED bundle, lon samples interleaved "
                              "over %d ranks (%d candidates per rank)" % (world, n1 // world),
                  "ms_per_cycle": sharded_ms, "value": n1 / (sharded_ms * 1e-3), "unit": UNIT,
                  "single_gpu_ms_per_cycle_same_run": float(np.mean(single_ms)),
                  "speedup_vs_single_gpu": float(np.mean(single_ms)) / sharded_ms,
                  "kernel": "candidate-major" if kernel_s == _l0.KERNEL_CANDIDATE_MAJOR else "step-parallel",
                  "winner": int(res_s.winner),
                  "note": "a shard of 131 072 / N candidates is one partial wave of warp marches (or one wave of the "
                          "step-parallel kernel): the cycle is bound by one march / the launch chain, not by throughput"}
        eng.set_candidate_stripe(rank, world)

    scen = scenario_batch_rate(local_rank, stream.cuda_stream, n_scenarios=args.scenarios, rank=rank, world=world,
                               cycles=20 if args.scenarios >= 256 else 10)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    from commonroad_rp_b200 import _lib as _l
    kernel_name = "rp::cand_kernel<512>" if main_kernel == _l.KERNEL_CANDIDATE_MAJOR else "rp::fused_kernel<256>"
    ncu = ncu_record(kernel_name)
    value = n_total * args.steps / (total_ms * 1e-3)
    fused_mean_ms = float(np.mean(fused_ms))
    cand_steps_launch = count * Np1
    achieved_tf = cand_steps_launch * FLOP_PER_CAND_STEP / (fused_mean_ms * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        # the cycle's answer (the reference's own run of this bundle, tests/golden/big_dense_full.npz: winner 97084,
        # 2 713 kinematic rejects, 8 179 colliders ranked before the winner -- at N = 1)
        "winner": int(res.winner), "n_feasible": int(res.n_feasible),
        "n_infeasible_kinematics": int(res.n_infeasible_kinematics), "n_infeasible_collision": int(res.n_infeasible_collision),
        "n_collision_total": int(res.n_collision_total),
        "config": {"workload": "dense sampling sweep 64 d x %d v x 32 t, 60-step horizon (BASELINE configs[3]%s)"
                               % (64 * n_gpus, "" if n_gpus == 1 else "; v grid scaled with N, lon samples interleaved over the ranks "
                                  "(rp_set_candidate_stripe), arg-min exchange: " +
                                  ("stores into peer-mapped mailboxes over NVLink (rp_peer_*)" if peer is not None else "NCCL")),
                   "candidates_per_cycle": n_total, "time_steps": Np1, "l2": "flushed between timed iterations (256 MiB fill)",
                   "mode": "select-only (winner states materialised), lazy collision pass as in the reference "
                           "(check_collision = 2), fmad off for parity"},
        "cand_timesteps_per_sec": value * Np1,
        "p50_cycle_ms": float(np.median(step_ms)),
        "ms_per_step_with_stage_events": float(np.mean(step_ms_events)),
        "stage_ms": {"coeff": float(stage_ms[0]), "fused": float(stage_ms[1]), "argmin": float(stage_ms[2]),
                     "winner_states": float(stage_ms[3]),
                     "note": "second pass of the same K steps with the library's stage events on (ms_per_step_with_stage_events)"},
        "warmup_requested": warmup_requested,
        "strong_scaling": strong,
        "exchange": None if world == 1 else {"kind": "peer" if peer is not None else "nccl", "equals_nccl": exchange_check},
        "e2e": {"value": n_total * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s / args.steps,
                "api": "rp_plan_grid + rp_fetch_states (host buffers)"},
        # NCCL exchange: record export, merge, count kernels of this library (+ NCCL's own two)
        "gpu_launches": int((launches_main + (3 if (world > 1 and peer is None) else 0)) * args.steps),
        "clocks": clocks,
        "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved_tf / fp64_peak if fp64_peak else None,
                     "traffic": ncu.get("dram_bytes_per_launch"), "traffic_source": ncu.get("source"),
                     "fp64_pipe_active_pct_ncu": ncu.get("fp64_pipe_active_pct"),
                     "kernel": kernel_name + (" + deferred collision check (deferred_collision_kernel x 2, deferred_gather_kernel)"
                                              if main_kernel == _l.KERNEL_CANDIDATE_MAJOR else ""),
                     "kernel_ms": fused_mean_ms,
                     "kernel_ms_note": "the library's stage events around the main stage: the march (cand_kernel, ~80 % of it) and "
                                       "the deferred lazy collision check of the same candidates; algorithmic work counted as for "
                                       "a full check of every candidate-timestep, so the lazy pass's skipped checks read as speed",
                     "peak_source": "DFMA micro-benchmark measured in this process right before the timed region (rp_measure_fp64_peak: "
                                    "148 x 8 blocks x 256 threads x 8 independent chains x 20 000 FMA, best of 4 timed launches after "
                                    "one warm-up, FMA = 2 flop; SM clock: `clocks`); MEASURED_PEAKS.json has no FP64 entry.  "
                                    "Algorithmic work = 200 flop per candidate-timestep (SURVEY 8d)",
                     "peak_samples": 4,
                     # the path multiplies and adds separately (--fmad=false: the reference's IEEE operation order decides
                     # flags bit-exactly), one flop per FP64 issue slot: the attainable ceiling is half the FMA peak
                     "frac_of_unfused_ceiling": 2.0 * achieved_tf / fp64_peak if fp64_peak else None,
                     "hbm_peak_gbs": peaks.get("hbm_gbs"), "hbm_peak_source": peak_kind},
    }
    # the same kernel against the HBM roofline (the contract's other bound): select-only mode moves 13 doubles in and
    # 16 bytes out per candidate -- two orders of magnitude below the copy bandwidth, the path is FP64-bound
    sel_bytes = count * SELECT_BYTES_PER_CAND
    line["roofline_hbm"] = {"bound": "hbm", "achieved": sel_bytes / (fused_mean_ms * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs"),
                            "unit": "GB/s", "frac": sel_bytes / (fused_mean_ms * 1e-3) / 1e9 / peaks.get("hbm_gbs"),
                            "traffic": ncu.get("dram_bytes_per_launch"), "algorithmic_bytes_per_launch": sel_bytes,
                            "peak_source": peak_kind + " (MEASURED_PEAKS.json)"}
    if lazy_ms is not None:
        line["all_collision_flags"] = {"value": n_total / (lazy_ms * 1e-3), "unit": UNIT, "ms_per_step": lazy_ms,
                                       "note": "check_collision=1: every feasible candidate gets its collision flag (the march "
                                               "checks each step as it goes); same winner / counters as the headline, whose lazy "
                                               "pass (the reference's semantics) checks only what can be ranked before the winner: "
                                               "the march stores the ego boxes, a tile-parallel checker runs afterwards"}
    if full is not None:
        gbs = cand_steps_launch * STATE_BYTES_PER_CAND_STEP / (full * 1e-3) / 1e9
        line["roofline_full_states"] = {"bound": "hbm", "achieved": gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                                        "frac": gbs / peaks.get("hbm_gbs"), "traffic": None, "kernel_ms": full,
                                        "peak_source": peak_kind + " (MEASURED_PEAKS.json)"}
    line["p50_replanning_cycle_ms"] = replanning_latency_b200()
    line["p50_replanning_cycle_ms"]["other_bundled_scenarios"] = {
        name: {k: v for k, v in replanning_latency_b200(name, repeats=3).items() if k in ("p50_ms", "p95_ms", "cycles")}
        for name in ("ZAM_Tjunction-1_42_T-1", "DEU_Test-1_1_T-1")}
    # the other policy: every cycle is ONE submission whatever happens (levels 1..3 always go together)
    line["p50_replanning_cycle_ms"]["speculation_always"] = {
        name: {k: v for k, v in replanning_latency_b200(name, repeats=3, speculation="always").items()
               if k in ("p50_ms", "p95_ms", "cycles")} for name in ("ZAM_Over-1_1", "DEU_Test-1_1_T-1")}
    line["scenario_batch"] = scen
    if not args.no_cpu_baseline:
        line["p50_replanning_cycle_ms"]["cpu_port"] = replanning_latency_port()
        cores = 1
        rate, desc, sec, n_s = cpu_port_rate(dense_workload(1), 3, cores)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                                "seconds": sec}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def capture_stdout():
    """Everything native libraries print on fd 1 (e.g. NCCL's version banner) goes to stderr; the ONE JSON line is
    written to the real stdout by ``emit``."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


if __name__ == "__main__":
    capture_stdout()
    main()
