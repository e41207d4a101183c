"""GPU probe (development aid): main-kernel time of the dense sweep when every sampled t has the same traj_len --
calibrates the per-candidate work model of parallel.balanced_shard_range.

    gpurun -- python tools/probe_tl.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    work = bench.dense_workload(1)
    eng = bench.make_engine(work, 0, stream.cuda_stream)
    eng.set_stage_timing(True)
    t_all = np.array(work["t"])
    for mode, t_one in [(m, t) for m in (1, 0) for t in (t_all.min(), 4.5, t_all.max())]:
        inp = bench.make_inputs(work, check_collision=mode)
        # 32 (nearly) equal horizons: same traj_len, distinct sample values
        t = [float(t_one) + 1e-9 * k for k in range(len(t_all))]
        eng.grid_upload(inp, t, work["lon"], work["d"])
        for _ in range(3):
            eng.grid_launch()
        torch.cuda.synchronize()
        ms = []
        for k in range(10):
            eng.grid_launch()
            torch.cuda.synchronize()
            ms.append(eng.stage_ms(0)[1])
        from commonroad_rp_b200._lib import traj_len_of
        res = eng.grid_result()
        print("check_collision %d  n_feasible %d  n_col %d" % (mode, res.n_feasible, res.n_collision_total))
        print("t = %.2f  traj_len %d  main %.4f ms" % (t_one, traj_len_of(t[0], bench.DT), float(np.mean(ms))), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
