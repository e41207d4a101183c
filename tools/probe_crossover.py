"""GPU probe (development aid): cycle time of one lon-interleaved shard of the dense sweep (1 / world of 131 072 candidates)
under both kernel policies, lazy collision pass.   gpurun -- python tools/probe_crossover.py"""
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
    inputs = bench.make_inputs(work)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for world in (1, 2, 4, 8, 16, 32):
        for name, pol in (("candidate-major", _lib.KERNEL_CANDIDATE_MAJOR), ("step-parallel", _lib.KERNEL_STEP_PARALLEL)):
            eng.set_kernel_policy(pol)
            eng.set_candidate_stripe(0, world)
            eng.grid_upload(inputs, work["t"], work["lon"], work["d"])
            for _ in range(3):
                eng.grid_launch()
            torch.cuda.synchronize()
            e0 = [torch.cuda.Event(enable_timing=True) for _ in range(10)]
            e1 = [torch.cuda.Event(enable_timing=True) for _ in range(10)]
            for k in range(10):
                flush.fill_(k)
                e0[k].record(stream)
                eng.grid_launch()
                e1[k].record(stream)
            torch.cuda.synchronize()
            res = eng.grid_result()
            ms = np.mean([a.elapsed_time(b) for a, b in zip(e0, e1)])
            print("%6d candidates  %-16s cycle %.4f ms  winner %d" % (res.n_candidates, name, ms, res.winner))


if __name__ == "__main__":
    main()
