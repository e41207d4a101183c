"""world_size-2 CPU test (gloo) of the multi-GPU arg-min protocol: every rank contributes its shard's
(cost, index, counters) record, one all-gather, lexicographic merge (commonroad_rp_b200/parallel.py).
The device kernels are replaced by numpy on a synthetic verdict table; the collective logic is the
same code the NCCL path runs."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from commonroad_rp_b200.parallel import merge_records, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, seed, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    cost = np.round(rng.uniform(0, 5, n), 1)            # many exact ties
    status = rng.integers(0, 3, n)                      # 0 feasible, 1 kinematic, 2 collision
    first, count = shard_range(n, rank, world)
    sl = slice(first, first + count)
    ok = np.flatnonzero(status[sl] == 0)
    if len(ok):
        k = ok[np.lexsort((ok, cost[sl][ok]))[0]]
        rec = torch.tensor([cost[sl][k], float(first + k), float((status[sl] == 1).sum()), float((status[sl] != 1).sum())],
                           dtype=torch.float64)
    else:
        rec = torch.tensor([float("inf"), float("inf"), float((status[sl] == 1).sum()), float((status[sl] != 1).sum())],
                           dtype=torch.float64)
    gathered = [torch.zeros(4, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, rec)
    winner, totals = merge_records(torch.stack(gathered))
    wc, wi = winner.tolist()
    idx = np.arange(first, first + count)
    before = torch.tensor([float(((status[sl] == 2) & ((cost[sl] < wc) | ((cost[sl] == wc) & (idx < wi)))).sum())],
                          dtype=torch.float64)
    dist.all_reduce(before)
    if rank == 0:
        out_q.put((wc, wi, totals.tolist(), before.item()))
    dist.destroy_process_group()


def test_two_rank_argmin_equals_single_rank():
    n, seed, world = 5000, 7, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    wc, wi, totals, before = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(seed)
    cost = np.round(rng.uniform(0, 5, n), 1)
    status = rng.integers(0, 3, n)
    ok = np.flatnonzero(status == 0)
    k = ok[np.lexsort((ok, cost[ok]))[0]]
    assert (wc, int(wi)) == (cost[k], k)
    assert totals == [float((status == 1).sum()), float((status != 1).sum())]
    idx = np.arange(n)
    assert before == float(((status == 2) & ((cost < cost[k]) | ((cost == cost[k]) & (idx < k)))).sum())


# ---- PeerExchange set-up is collective-safe: the ranks either all open the group or all raise ------------------
class _FakeEngine:
    """Stands in for _lib.Engine on a box without a GPU: records the calls, fails where told to."""

    def __init__(self, fail_create=False, fail_open=False):
        self.fail_create, self.fail_open = fail_create, fail_open
        self.opened = None
        self.closed = 0

    def peer_create(self):
        from commonroad_rp_b200._lib import RpError
        if self.fail_create:
            raise RpError("no mailbox")
        return bytes(range(64))

    def peer_open(self, rank, world, handles):
        from commonroad_rp_b200._lib import RpError
        if self.fail_open:
            raise RpError("cannot map")
        self.opened = (rank, world, len(handles))

    def peer_close(self):
        self.closed += 1


def _peer_worker(rank, world, port, mode, out_q):
    from commonroad_rp_b200._lib import RpError
    from commonroad_rp_b200.parallel import PeerExchange
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = _FakeEngine(fail_create=(mode == "create" and rank == 1), fail_open=(mode == "open" and rank == 0))
    try:
        PeerExchange(eng)
        outcome = "opened"
    except RpError as exc:
        outcome = "raised: %s" % exc
    out_q.put((rank, outcome, eng.opened, eng.closed))
    dist.barrier()
    dist.destroy_process_group()


def _run_peer_setup(mode):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return got


def test_peer_exchange_setup_opens_on_every_rank():
    got = _run_peer_setup("ok")
    assert [g[1] for g in got] == ["opened", "opened"]
    assert [g[2] for g in got] == [(0, 2, 128), (1, 2, 128)]


def test_peer_exchange_setup_fails_on_every_rank_together():
    for mode, word in (("create", "created"), ("open", "mapped")):
        got = _run_peer_setup(mode)
        assert all(g[1].startswith("raised") and word in g[1] for g in got), got
        assert all(g[3] >= 1 for g in got)                 # every rank released what it had
