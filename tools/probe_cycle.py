"""GPU probe (development aid): wall time of one replanning-size cycle through rp_plan_levels (1, 2, 3 levels in one
launch) and through rp_plan_grid + rp_fetch_states, on the recorded cycles of a bundled scenario.

    gpurun -- python tools/probe_cycle.py [scenario] [reps]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from commonroad_rp_b200 import _lib  # noqa: E402
from commonroad_rp_b200.sampling import FixedIntervalSampling, VelocitySampling  # noqa: E402
from commonroad_rp_b200.utility.config import ReactivePlannerConfiguration  # noqa: E402
from tests import golden_io, helpers as H  # noqa: E402
from oracle import rp_oracle as O  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "ZAM_Over-1_1"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    torch.cuda.set_device(0)
    z = np.load(os.path.join(ROOT, "tests", "golden", "cyc_%s.npz" % name))
    meta = json.loads(str(z["meta"]))
    cfg = ReactivePlannerConfiguration()
    cfg.planning.time_steps_computation = meta["N"]
    cfg.planning.planning_horizon = meta["N"] * meta["dt"]
    cfg.sampling.t_min = meta["t_min"]
    fs = FixedIntervalSampling(cfg)
    ci = 1
    x = z["c%d_x0" % ci]
    v0 = float(x[3])
    lo = max(0, v0 - 0.125 * meta["N"] * meta["dt"] * O.vehicle_dict()["a_max"])
    fs.samples_v = VelocitySampling(lo, max(lo + 5.0, v0 + 2), 4)
    x0_lon, x0_lat = z["c%d_x0_lon" % ci], z["c%d_x0_lat" % ci]
    grids = [fs.sample_grid(lv, x0_lat, "velocity_keeping") for lv in (1, 2, 3)]
    prob = {"t": grids[0][0], "lon": grids[0][1], "d": grids[0][2], "x0_lon": x0_lon, "x0_lat": x0_lat, "x0_orientation": float(x[2]),
            "x0_time_step": int(x[7]), "lon_mode": "velocity_keeping", "low_vel_mode": False, "dt": meta["dt"], "N": meta["N"],
            "factor": 1, "draw_all": False, "constraints": O.CONSTRAINTS,
            "cost": {"kind": "default", "desired_speed": meta["desired_velocity"], "desired_s": None, "desired_d": 0.0, "w_a": 5},
            "vehicle": O.vehicle_dict(), "ref": {k: z[k] for k in ("ref_pos", "ref_theta", "ref_curv", "ref_curv_d")},
            "ccosy": {"path": z["cc_path"], "S": z["cc_S"], "normals": z["cc_normals"], "limit": 20.0},
            "obstacles": golden_io.unpack_obstacles(z, "ob_")}
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng = H.engine_for(prob, 0, stream.cuda_stream)
    levels = [(t, lon, d, np.array([_lib.traj_len_of(q, meta["dt"]) for q in t], dtype=np.int32)) for t, lon, d in grids]
    for lazy in (2, 1):
        inputs = H.inputs_for(prob, check_collision=lazy)
        for nl in (1, 2, 3):
            for _ in range(20):
                eng.plan_levels(inputs, levels[:nl])
            ts = []
            for _ in range(reps):
                t0 = time.perf_counter()
                rec, chosen = eng.plan_levels(inputs, levels[:nl])
                st = eng.fetch_states(rec[chosen].winner)
                ts.append(time.perf_counter() - t0)
            ts = np.array(ts) * 1e6
            n = sum(len(a) * len(b) * len(c) for a, b, c, _ in levels[:nl])
            print("check_collision=%d plan_levels %d level(s) %5d cand: p50 %.1f us  p95 %.1f us  min %.1f us" % (
                lazy, nl, n, np.percentile(ts, 50), np.percentile(ts, 95), ts.min()))
        if lazy == 1:
            for gap_us in (0, 50, 100, 200, 1000):
                ts = []
                for _ in range(reps):
                    t_end = time.perf_counter() + gap_us * 1e-6
                    while time.perf_counter() < t_end:
                        pass
                    t0 = time.perf_counter()
                    rec, chosen = eng.plan_levels(inputs, levels[:3])
                    ts.append(time.perf_counter() - t0)
                ts = np.array(ts) * 1e6
                print("  host busy-wait %4d us between cycles: plan_levels(3 levels) p50 %.1f us  p95 %.1f us" % (gap_us, np.percentile(ts, 50), np.percentile(ts, 95)))
        t, lon, d, tl = levels[0]
        for _ in range(20):
            eng.plan_grid(inputs, t, lon, d, tl)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            r = eng.plan_grid(inputs, t, lon, d, tl)
            st = eng.fetch_states(r.winner)
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts) * 1e6
        print("check_collision=%d plan_grid   level 1: p50 %.1f us  p95 %.1f us  min %.1f us" % (lazy, np.percentile(ts, 50), np.percentile(ts, 95), ts.min()))


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "stamps"):
    main()


def stamps():
    """RP_B200_LIB=build/librp_timing.so python tools/probe_cycle.py stamps : cycle stamps of block 0 (RP_CYCLE_TIMING build)"""
    import ctypes as C
    torch.cuda.set_device(0)
    name = "ZAM_Over-1_1"
    z = np.load(os.path.join(ROOT, "tests", "golden", "cyc_%s.npz" % name))
    meta = json.loads(str(z["meta"]))
    cfg = ReactivePlannerConfiguration()
    cfg.planning.time_steps_computation = meta["N"]
    cfg.planning.planning_horizon = meta["N"] * meta["dt"]
    cfg.sampling.t_min = meta["t_min"]
    fs = FixedIntervalSampling(cfg)
    ci = 1
    x = z["c%d_x0" % ci]
    v0 = float(x[3])
    lo = max(0, v0 - 0.125 * meta["N"] * meta["dt"] * O.vehicle_dict()["a_max"])
    fs.samples_v = VelocitySampling(lo, max(lo + 5.0, v0 + 2), 4)
    x0_lon, x0_lat = z["c%d_x0_lon" % ci], z["c%d_x0_lat" % ci]
    grids = [fs.sample_grid(lv, x0_lat, "velocity_keeping") for lv in (1, 2, 3)]
    prob = {"t": grids[0][0], "lon": grids[0][1], "d": grids[0][2], "x0_lon": x0_lon, "x0_lat": x0_lat, "x0_orientation": float(x[2]),
            "x0_time_step": int(x[7]), "lon_mode": "velocity_keeping", "low_vel_mode": False, "dt": meta["dt"], "N": meta["N"],
            "factor": 1, "draw_all": False, "constraints": O.CONSTRAINTS,
            "cost": {"kind": "default", "desired_speed": meta["desired_velocity"], "desired_s": None, "desired_d": 0.0, "w_a": 5},
            "vehicle": O.vehicle_dict(), "ref": {k: z[k] for k in ("ref_pos", "ref_theta", "ref_curv", "ref_curv_d")},
            "ccosy": {"path": z["cc_path"], "S": z["cc_S"], "normals": z["cc_normals"], "limit": 20.0},
            "obstacles": golden_io.unpack_obstacles(z, "ob_")}
    eng = H.engine_for(prob, 0, None)
    levels = [(t, lon, d, np.array([_lib.traj_len_of(q, meta["dt"]) for q in t], dtype=np.int32)) for t, lon, d in grids]
    lib = _lib.load_library()
    lib.rp_debug_stamps.argtypes = [C.POINTER(C.c_longlong)]
    lib.rp_debug_launch_floor.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    us = (C.c_double * 4)()
    lib.rp_debug_launch_floor(eng._ctx, 500, us)
    print("launch floor (null kernel, 9 x 64 threads): small params + flag %.1f us, 7.4 KB params + flag %.1f us, "
          "small + stream sync %.1f us, 7.4 KB + stream sync %.1f us" % tuple(us))
    names = ["entry", "setup", "solve", "poly+prefilter", "orient+curv", "limits+proj", "publish", "extension", "cost", "collision",
             "verdict", "body end", "last-block start", "select+gather", "fence"]
    wall = []
    for nl in (1, 3):
        inputs = H.inputs_for(prob, check_collision=1)
        acc, gl = [], []
        for _ in range(50):
            t0 = time.perf_counter()
            eng.plan_levels(inputs, levels[:nl])
            wall.append(time.perf_counter() - t0)
            buf = (C.c_longlong * 32)()
            lib.rp_debug_stamps(buf)
            acc.append(np.array(buf[:15], dtype=np.int64))
            g = np.array(buf[16:32], dtype=np.int64)
            gl.append([g[11] - g[0], g[15] - g[0], g[12] - g[0], g[13] - g[0], g[14] - g[0]])
        print("globaltimer ns since block 0 entry: block-0 body end %d, LAST block's fence done %d, selection start %d, "
              "select+gather done %d, flag %d" % tuple(np.median(np.array(gl[10:]), axis=0)))
        bt = (C.c_longlong * (3 * 1024))()
        grid, threads = C.c_int(), C.c_int()
        lib.rp_debug_block_times.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.rp_debug_block_times(eng._ctx, bt, 1024, C.byref(grid), C.byref(threads))
        g = min(grid.value, 1024)
        b = np.array(bt[:3 * g], dtype=np.int64).reshape(g, 3)
        t00 = b[:, 0].min()
        print("grid %d x %d threads; last cycle, per block (entry ns, body ns, sm):" % (grid.value, threads.value))
        print("  " + " ".join("%d:%d+%d@%d" % (i, b[i, 0] - t00, b[i, 1] - b[i, 0], b[i, 2]) for i in range(g)))
        a = np.median(np.array(acc[10:]), axis=0)
        print("levels", nl, "wall p50 %.1f us" % (np.median(wall[-40:]) * 1e6), "(cycles since entry; block 0, its LAST group)")
        for k in range(1, 15):
            print("  %-18s %8d  (+%d)" % (names[k], a[k] - a[0], a[k] - a[k - 1]))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "stamps":
    stamps()
