"""GPU probe (development aid, not part of the product or the tests): dense sweep through the candidate-major kernel
with check_collision = 1 / 2 / 0 -- main-kernel time, lazy collision count, how many candidates the lazy pass skipped.

    gpurun -- python tools/probe_lazy.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from commonroad_rp_b200 import _lib  # noqa: E402


def main():
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    work = bench.dense_workload(1)
    eng = bench.make_engine(work, 0, stream.cuda_stream)
    eng.set_stage_timing(True)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    out = {}
    for mode in (1, 2, 0):
        inp = bench.make_inputs(work, check_collision=mode)
        eng.grid_upload(inp, work["t"], work["lon"], work["d"])
        for _ in range(3):
            eng.grid_launch()
        torch.cuda.synchronize()
        ms, tot = [], []
        for k in range(10):
            flush.fill_(k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            eng.grid_launch()
            e1.record(stream)
            torch.cuda.synchronize()
            ms.append(eng.stage_ms(0)[1])
            tot.append(e0.elapsed_time(e1))
        res = eng.grid_result()
        cost, status, _, _ = eng.fetch_candidates()
        out[mode] = (res, cost, status)
        print("check_collision=%d: main %.4f ms, cycle %.4f ms, winner %d cost %.6f, n_feasible %d, n_kin %d, n_col %d, "
              "status histogram %s" % (mode, np.mean(ms), np.mean(tot), res.winner, res.winner_cost, res.n_feasible,
                                       res.n_infeasible_kinematics, res.n_infeasible_collision,
                                       np.bincount(status, minlength=5).tolist()), flush=True)
    res1, cost1, status1 = out[1]
    feas = status1 != _lib.ST_KINEMATIC
    order = np.argsort(cost1[feas], kind="stable")
    rank_of_winner = int(np.nonzero(np.arange(len(cost1))[feas][order] == res1.winner)[0][0])
    print("winner is the %d-th cheapest of %d kinematically feasible candidates" % (rank_of_winner, int(feas.sum())))
    eng.close()


if __name__ == "__main__":
    main()
