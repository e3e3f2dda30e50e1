"""Multi-GPU plumbing: one process per GPU, instances sharded, one all-gather.

MPC instances are independent (the reference solves them one by one,
``qpmpc/solve_mpc.py:42-44``), so rank r of W owns a contiguous block of the
batch and no data-path collective is needed while solving.  The only exchange
is the all-gather of the stacked input trajectories U (+ status) so that every
rank ends with the full result -- NCCL over NVLink on GPUs, gloo in CPU tests.
"""

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of rank's contiguous block; the first batch % world ranks get
    one extra instance."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_plans(
    U_local: torch.Tensor, status_local: torch.Tensor, batch: int, group=None
) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather per-rank results into U [batch, n] and status [batch].

    Equal shards use a single ``all_gather_into_tensor`` per tensor; uneven
    shards are padded to the largest one and trimmed afterwards.
    """
    world = dist.get_world_size(group)
    n = U_local.shape[1]
    sizes = [shard_bounds(batch, r, world) for r in range(world)]
    counts = [hi - lo for lo, hi in sizes]
    cmax = max(counts)
    if U_local.shape[0] != cmax:
        pad = cmax - U_local.shape[0]
        U_local = torch.cat([U_local, U_local.new_zeros((pad, n))])
        status_local = torch.cat([status_local, status_local.new_zeros(pad)])
    U_all = U_local.new_empty((world * cmax, n))
    st_all = status_local.new_empty(world * cmax)
    dist.all_gather_into_tensor(U_all, U_local.contiguous(), group=group)
    dist.all_gather_into_tensor(st_all, status_local.contiguous(), group=group)
    if all(c == cmax for c in counts):
        return U_all, st_all
    keep = torch.cat([torch.arange(r * cmax, r * cmax + c) for r, c in enumerate(counts)])
    keep = keep.to(U_all.device)
    return U_all[keep], st_all[keep]


def solve_mpc_sharded(workload: dict, rank: Optional[int] = None, world: Optional[int] = None,
                      group=None, dtype=torch.float64):
    """Solve this rank's block of a host-side workload dict on the current CUDA
    device and all-gather U / status.  Returns (U [batch, n], status [batch])."""
    from .batched import solve_mpc_batch
    from .workloads import slice_workload, to_batched

    rank = dist.get_rank(group) if rank is None else rank
    world = dist.get_world_size(group) if world is None else world
    lo, hi = shard_bounds(workload["batch"], rank, world)
    plan = solve_mpc_batch(to_batched(slice_workload(workload, lo, hi), dtype=dtype))
    U_local = plan.inputs.reshape(hi - lo, -1)
    return gather_plans(U_local, plan.status, workload["batch"], group)
