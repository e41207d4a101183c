"""GPU probe (development aid): first-collision step distribution of the dense sweep (all flags mode)."""
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
    inputs = bench.make_inputs(work)
    res = eng.plan_grid(inputs, work["t"], work["lon"], work["d"])
    cost, status, reason, step = eng.fetch_candidates()
    print("winner", res.winner, "cost", res.winner_cost, {k: (v.shape if hasattr(v, "shape") else v) for k, v in work.items() if k not in ("t", "lon", "d")}.keys())
    col = status == 2
    before = col & (cost <= res.winner_cost)
    for name, m in (("all colliders", col), ("colliders ranked before the winner", before)):
        s = step[m]
        print(name, m.sum(), "first-collision step percentiles 10/25/50/75/90/99:", np.percentile(s, [10, 25, 50, 75, 90, 99]),
              "share < 8: %.2f  < 16: %.2f  < 32: %.2f" % ((s < 8).mean(), (s < 16).mean(), (s < 32).mean()))
    ob = work.get("obstacles", {})
    print({k: np.asarray(v).shape for k, v in ob.items()} if isinstance(ob, dict) else type(ob))


if __name__ == "__main__":
    main()
