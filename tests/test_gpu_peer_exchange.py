"""Multi-GPU arg-min exchange over peer-mapped memory (rp_peer_*): sharded bundles give the result of the
unsharded plan.  world = 1 runs in-process; world = 2 runs two processes -- on two GPUs when the box has them, else
both on cuda:0 (CUDA IPC between processes of one device; the kernels of the two ranks time-slice)."""
import os
import socket

import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _problem():
    from tests.test_gpu_parity import _bundle
    return _bundle(seed=3, level=3, N=30, s_dot0=11.0)


def _reference_result(prob, eng):
    from commonroad_rp_b200 import _lib
    r = eng.plan_grid(H.inputs_for(prob, check_collision=_lib.COLLISION_ALL), prob["t"], prob["lon"], prob["d"])
    return _as_tuple(r), eng.fetch_states(r.winner)


def _as_tuple(r):
    return (r.winner, r.winner_cost, r.n_candidates, r.n_feasible, r.n_infeasible_kinematics, r.n_infeasible_collision,
            r.n_collision_total, tuple(r.reason_counts))


def test_peer_exchange_world_1():
    from commonroad_rp_b200 import _lib
    from commonroad_rp_b200.parallel import PeerExchange
    prob = _problem()
    eng = H.engine_for(prob)
    want, want_states = _reference_result(prob, eng)
    n = want[2]
    ex = PeerExchange(eng, rank=0, world=1)
    eng.set_candidate_range(0, n)
    for _ in range(3):                                      # epochs advance, slots alternate
        r = eng.plan_grid(H.inputs_for(prob, check_collision=_lib.COLLISION_ALL), prob["t"], prob["lon"], prob["d"])
        assert _as_tuple(r) == want
        assert np.array_equal(eng.fetch_states(r.winner), want_states)
    ex.close()
    eng.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_dev, kernel, out_q):
    import torch.distributed as dist
    from commonroad_rp_b200 import _lib
    from commonroad_rp_b200.parallel import PeerExchange, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = rank % n_dev
    torch.cuda.set_device(dev)
    prob = _problem()
    eng = H.engine_for(prob, device=dev, stream=torch.cuda.current_stream().cuda_stream)
    eng.set_kernel_policy(kernel)
    n = len(prob["t"]) * len(prob["lon"]) * len(prob["d"])
    ex = PeerExchange(eng)
    got = []
    for cycle in range(3):
        if cycle == 2:
            eng.set_candidate_stripe((rank + cycle) % world, world)          # lon-interleaved shards instead of t-major tiles
        else:
            first, count = shard_range(n, (rank + cycle) % world, world)     # the shards change hands between cycles
            eng.set_candidate_range(first, count)
        r = eng.plan_grid(H.inputs_for(prob, check_collision=_lib.COLLISION_ALL), prob["t"], prob["lon"], prob["d"])
        got.append((_as_tuple(r), eng.fetch_states(r.winner)))
    out_q.put((rank, got))
    dist.barrier()
    ex.close()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("kernel", [1, 2])          # _lib.KERNEL_STEP_PARALLEL, _lib.KERNEL_CANDIDATE_MAJOR
def test_peer_exchange_two_ranks_equal_the_unsharded_plan(kernel):
    import torch.multiprocessing as mp
    prob = _problem()
    eng = H.engine_for(prob)
    want, want_states = _reference_result(prob, eng)
    eng.close()
    world, n_dev = 2, torch.cuda.device_count()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_dev, kernel, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank in range(world):
        for got, states in results[rank]:
            assert got == want, (rank, got, want)
            assert np.array_equal(states, want_states)


def _worker_timeout(rank, world, port, n_dev, out_q):
    import time
    import torch.distributed as dist
    from commonroad_rp_b200 import _lib
    from commonroad_rp_b200.parallel import PeerExchange
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = rank % n_dev
    torch.cuda.set_device(dev)
    prob = _problem()
    eng = H.engine_for(prob, device=dev, stream=torch.cuda.current_stream().cuda_stream)
    n = len(prob["t"]) * len(prob["lon"]) * len(prob["d"])
    ex = PeerExchange(eng)
    if rank == 0:                                            # rank 1 never launches its shard
        eng.set_candidate_range(0, n // 2)
        t0 = time.perf_counter()
        try:
            eng.plan_grid(H.inputs_for(prob, check_collision=_lib.COLLISION_ALL), prob["t"], prob["lon"], prob["d"])
            outcome = "no error"
        except _lib.RpError as exc:
            outcome = str(exc)
        waited = time.perf_counter() - t0
        ex.close()
        eng.set_candidate_range(0, -1)                       # the context keeps working afterwards
        r = eng.plan_grid(H.inputs_for(prob, check_collision=_lib.COLLISION_ALL), prob["t"], prob["lon"], prob["d"])
        out_q.put((outcome, waited, _as_tuple(r)))
    dist.barrier()
    if rank != 0:
        ex.close()
    eng.close()
    dist.destroy_process_group()


def test_missing_rank_times_out_instead_of_hanging():
    """A peer group of two where the second rank never launches: the first rank's result call fails with an error after
    the bounded wait (~8 s) -- the kernel does not spin forever -- and the context keeps working afterwards."""
    import torch.multiprocessing as mp
    prob = _problem()
    eng = H.engine_for(prob)
    want, _ = _reference_result(prob, eng)
    eng.close()
    world, n_dev = 2, torch.cuda.device_count()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_timeout, args=(r, world, port, n_dev, q)) for r in range(world)]
    for p in procs:
        p.start()
    outcome, waited, after = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert "timed out" in outcome, outcome
    assert waited < 60.0
    assert after == want
