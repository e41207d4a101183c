"""Multi-GPU plumbing: one process per GPU, bundles sharded t-major over the ranks, one small
NCCL exchange for the global arg-min.  The reference's only parallelism is
fork + pickle of candidate chunks (reactive_planner.py:1084-1111); candidates are independent up to
the final arg-min, so the only data that ever crosses NVLink is a 32-byte record per rank.
"""
import torch
import torch.distributed as dist


def shard_range(n_candidates: int, rank: int, world: int):
    """Contiguous (t-major) tile of the enumeration space owned by ``rank``: (first, count)."""
    per = (n_candidates + world - 1) // world
    first = min(rank * per, n_candidates)
    return first, max(0, min(per, n_candidates - first))


def merge_records(gathered: torch.Tensor):
    """Lexicographic min on (cost, enumeration index) over the ranks' records [world, 4]; returns
    (winner[2] = [cost, index] (+inf = none), totals[2] = [n_infeasible_kinematics, n_feasible])."""
    cost = gathered[:, 0]
    idx = gathered[:, 1]
    cmin = cost.min()
    imin = torch.where(cost == cmin, idx, torch.full_like(idx, float("inf"))).min()
    return torch.stack([cmin, imin]), gathered[:, 2:4].sum(dim=0)


def global_argmin(engine, rec: torch.Tensor, world: int, group=None):
    """After ``engine.grid_launch()`` on every rank: all-gather the shard records, pick the global
    winner and count the colliders ranked before it.  Everything stays on the device and on the
    current stream; returns device tensors (winner[2], totals[2], n_collision_before[1])."""
    engine.export_record_dev(rec.data_ptr())
    gathered = torch.empty(world * rec.numel(), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(gathered, rec, group=group)
    winner, totals = merge_records(gathered.view(world, rec.numel()))
    before = torch.zeros(1, dtype=torch.float64, device=rec.device)
    engine.count_colliders_before_dev(winner.data_ptr(), before.data_ptr())
    dist.all_reduce(before, group=group)
    return winner, totals, before
