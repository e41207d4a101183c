"""GPU probe (development aid): where the wall time of ReactivePlanner.plan() goes on the replanning-size bundles of
ZAM_Over-1_1 -- engine call (ctypes + launches + D2H + sync) against the Python around it.

    gpurun -- python tools/probe_plan_latency.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from commonroad_rp_b200 import _lib  # noqa: E402


def main():
    torch.cuda.set_device(0)
    acc = {"plan_grid": [], "fetch_states": [], "grid_result": []}
    for name in ("plan_grid", "fetch_states"):
        orig = getattr(_lib.Engine, name)

        def wrap(self, *a, _orig=orig, _name=name, **k):
            t0 = time.perf_counter()
            r = _orig(self, *a, **k)
            acc[_name].append(time.perf_counter() - t0)
            return r
        setattr(_lib.Engine, name, wrap)
    res = bench.replanning_latency_b200("ZAM_Over-1_1", repeats=5)
    print(res)
    n = res["cycles"]
    for k, v in acc.items():
        if v:
            v = np.array(v[-n:]) * 1e3
            print("%-14s calls/cycle %.2f  p50 %.4f ms  mean %.4f ms" % (k, len(v) / n, np.percentile(v, 50), v.mean()))
    # the engine call alone, back to back, on the last bundle
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    bench.replanning_latency_b200("ZAM_Over-1_1", repeats=3)
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(22)


if __name__ == "__main__":
    main()
