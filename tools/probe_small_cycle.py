"""GPU probe (development aid): the engine call of a replanning-size bundle (120 candidates, N = 20) -- wall time of
rp_plan_grid + rp_fetch_states against the GPU time of its three launches.

    gpurun -- python tools/probe_small_cycle.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from commonroad_rp_b200 import _lib  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.test_gpu_parity import _bundle  # noqa: E402


def main():
    torch.cuda.set_device(0)
    for level in (1, 2, 3):
        prob = _bundle(seed=0, level=level, N=20)
        eng = H.engine_for(prob)
        eng.set_stage_timing(len(sys.argv) > 1 and sys.argv[1] == "stages")
        inp = H.inputs_for(prob, check_collision=_lib.COLLISION_LAZY)
        t, lon, d = np.array(prob["t"]), np.array(prob["lon"]), np.array(prob["d"])
        tl = np.array([_lib.traj_len_of(x, prob["dt"]) for x in t], dtype=np.int32)
        for _ in range(50):
            r = eng.plan_grid(inp, t, lon, d, tl)
        wall, stages = [], []
        for _ in range(300):
            t0 = time.perf_counter()
            r = eng.plan_grid(inp, t, lon, d, tl)
            if r.winner >= 0:
                eng.fetch_states(r.winner)
            wall.append(time.perf_counter() - t0)
            stages.append(eng.stage_ms(0) if len(sys.argv) > 1 and sys.argv[1] == "stages" else [0, 0, 0, 0])
        wall = np.array(wall) * 1e3
        st = np.mean(stages, axis=0)
        print("level %d: %d candidates  wall p50 %.4f ms  mean %.4f | GPU stages ms: prep %.4f main %.4f select %.4f states %.4f (sum %.4f)"
              % (level, r.n_candidates, np.percentile(wall, 50), wall.mean(), st[0], st[1], st[2], st[3], st.sum()), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
