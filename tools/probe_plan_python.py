"""GPU probe (development aid): wall time of the pieces of ReactivePlanner.plan() on ZAM_Over-1_1 (perf_counter
wrappers around the planner's own methods; each wrapper costs ~0.2 us).

    gpurun -- python tools/probe_plan_python.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from commonroad_rp_b200 import _lib, reactive_planner as RP, sampling as S  # noqa: E402

ACC = {}


def wrap(obj, name, label=None):
    orig = getattr(obj, name)
    label = label or name

    def w(*a, **k):
        t0 = time.perf_counter()
        try:
            return orig(*a, **k)
        finally:
            ACC.setdefault(label, []).append(time.perf_counter() - t0)
    setattr(obj, name, w)


def main():
    torch.cuda.set_device(0)
    P = RP.ReactivePlanner
    for n in ("_create_trajectory_bundle", "_get_optimal_trajectory", "_compute_trajectory_pair", "_sync_device_tables",
              "_plan_inputs", "_device_cost_spec", "_view", "_reset_statistics"):
        wrap(P, n)
    wrap(_lib.Engine, "plan_grid", "Engine.plan_grid")
    wrap(_lib.Engine, "plan_levels", "Engine.plan_levels")
    wrap(_lib.Engine, "select_level", "Engine.select_level")
    wrap(P, "_cycle_levels")
    wrap(_lib.Engine, "fetch_states", "Engine.fetch_states")
    wrap(_lib.Engine, "_grid_args", "Engine._grid_args")
    wrap(S.FixedIntervalSampling, "sample_grid", "sample_grid")
    wrap(RP, "shift_orientation", "shift_orientation")
    wrap(_lib.Engine, "cycle_winner_states", "Engine.cycle_winner_states")
    from commonroad_rp_b200 import trajectories as T
    wrap(T.TrajectorySample, "_set_states", "_set_states")
    wrap(P, "plan", "plan")
    if RP._rp_pack is not None:
        class _Shim:
            pack = staticmethod(RP._rp_pack.pack)
        RP._rp_pack = _Shim
        wrap(_Shim, "pack", "_rp_pack.pack")
    if _lib._rp_pack is not None:
        class _Shim2:
            plan_levels = staticmethod(_lib._rp_pack.plan_levels)
        _lib._rp_pack = _Shim2
        wrap(_Shim2, "plan_levels", "C rp_plan_levels (via _rp_pack)")
    res = bench.replanning_latency_b200("ZAM_Over-1_1", repeats=5)
    n = res["cycles"]
    print(res["p50_ms"], res["p95_ms"], n)
    for k, v in sorted(ACC.items(), key=lambda kv: -np.sum(kv[1][-n:])):
        v = np.array(v[-(len(v) * n // (n + n // 5)):]) * 1e6
        print("%-28s calls/cycle %.2f  p50 %7.1f us  mean %7.1f us  per cycle %7.1f us" % (k, len(v) / n, np.percentile(v, 50), v.mean(), v.sum() / n))


if __name__ == "__main__":
    main()
