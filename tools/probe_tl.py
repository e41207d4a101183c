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
    inp = bench.make_inputs(work)
    t_all = np.array(work["t"])
    for t_one in (t_all.min(), 3.7, 4.5, 5.2, t_all.max()):
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
        print("t = %.2f  traj_len %d  main %.4f ms" % (t_one, traj_len_of(t[0], bench.DT), float(np.mean(ms))), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
