"""GPU probe (development aid): main-kernel time of the dense sweep (BASELINE configs[3]) for the library named by
RP_B200_LIB, with the reference-fixture check of winner / counters.   gpurun -- python tools/probe_dense.py [tag]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("RP_B200_LIB", "default")
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    work = bench.dense_workload(1)
    eng = bench.make_engine(work, 0, stream.cuda_stream)
    inputs = bench.make_inputs(work)
    if os.environ.get("RP_PROBE_COLLISION"):
        inputs.check_collision = int(os.environ["RP_PROBE_COLLISION"])       # 0: the march alone
    eng.grid_upload(inputs, work["t"], work["lon"], work["d"])
    for _ in range(5):
        eng.grid_launch()
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(20)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(20)]
    for k in range(20):
        flush.fill_(k)
        e0[k].record(stream)
        eng.grid_launch()
        e1[k].record(stream)
    torch.cuda.synchronize()
    cyc = np.array([a.elapsed_time(b) for a, b in zip(e0, e1)])
    eng.set_stage_timing(True)
    ms = []
    for k in range(20):
        flush.fill_(k)
        eng.grid_launch()
        torch.cuda.synchronize()
        ms.append(eng.stage_ms(0)[1])
    res = eng.grid_result()
    ok = (res.winner, res.n_infeasible_kinematics, res.n_infeasible_collision) == (97084, 2713, 8179)
    print("%-28s main kernel %.4f ms (min %.4f)  cycle %.4f ms  winner/counters %s" % (
        tag, np.mean(ms), np.min(ms), cyc.mean(), "OK" if ok else "MISMATCH %r" % ((res.winner, res.n_infeasible_kinematics, res.n_infeasible_collision),)))


if __name__ == "__main__":
    main()
