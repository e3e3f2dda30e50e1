"""Multi-GPU plumbing: one process per GPU, instances sharded, one all-gather.

MPC instances are independent (the reference solves them one by one,
``qpmpc/solve_mpc.py:42-44``), so rank r of W owns a contiguous block of the
batch and no data-path collective is needed while solving.  The only exchange
is the all-gather of the stacked input trajectories U (+ status) so that every
rank ends with the full result -- NCCL over NVLink on GPUs, gloo in CPU tests.
"""

import ctypes
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


class PeerGather:
    """Fused solve + all-gather over NVLink: every rank's kernel stores its U
    rows (and status) directly into the symmetric buffers of ALL ranks
    (``qpmpc_b200_solve_scatter``), so no collective moves data afterwards; one
    barrier makes the rows visible.  Replaces ``solve_mpc_batch`` followed by
    ``all_gather_into_tensor`` for equal shards of ``rows_per_rank`` instances.

    Needs ``torch.distributed._symmetric_memory`` (CUDA VMM handles shared
    across the processes of one node); raises ``RuntimeError`` if the buffers
    cannot be set up, in which case callers fall back to :func:`gather_plans`.
    """

    def __init__(self, rows_per_rank: int, nb_vars: int, dtype=torch.float64, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 8:
            raise RuntimeError("the fused gather addresses at most 8 peers")
        self.rows, self.n = int(rows_per_rank), int(nb_vars)
        dev = torch.device("cuda", torch.cuda.current_device())
        total = self.world * self.rows
        self.dtype = dtype
        # Two sets of buffers used alternately: a rank that is one call ahead stores into the set
        # its peers are NOT reading (they can only be one call behind: the barrier below), so a
        # result stays valid, in stream order, until the call after the next one.
        self._sets = []
        for _ in range(2):
            U_all = symm_mem.empty((total, self.n), dtype=dtype, device=dev)
            status_all = symm_mem.empty((total,), dtype=torch.int32, device=dev)
            hU = symm_mem.rendezvous(U_all, self.group)
            hS = symm_mem.rendezvous(status_all, self.group)
            ptrU = [int(v) for v in hU.buffer_ptrs]
            ptrS = [int(v) for v in hS.buffer_ptrs]
            if len(ptrU) != self.world or not all(ptrU):
                raise RuntimeError("symmetric memory rendezvous returned no peer pointers")
            self._sets.append((U_all, status_all, hU, ptrU, ptrS))
        self._calls = 0
        self.U_all, self.status_all = self._sets[0][0], self._sets[0][1]

    def solve(self, problem, max_iter: int = 0):
        """Solve this rank's ``rows_per_rank`` instances; on return (stream
        order) the returned ``U_all`` / ``status_all`` hold the rows of every
        rank.  They stay valid until the call after the next one (two buffer
        sets alternate): consume them in stream order before that."""
        from . import _capi
        from .batched import _ptr

        if problem.batch_size != self.rows or problem.nb_vars != self.n:
            raise ValueError("problem does not match the gather buffers")
        if problem.dtype != self.dtype:
            raise ValueError(f"problem is {problem.dtype}, the gather buffers are {self.dtype}")
        U_all, status_all, hU, ptrU, ptrS = self._sets[self._calls % 2]
        self._calls += 1
        self.U_all, self.status_all = U_all, status_all
        lib = _capi.load()
        desc = problem.desc(_capi.ACTIVE_SET, max_iter, 0.0)
        peers = _capi.Peers()
        peers.count = self.world
        peers.row_offset = self.rank * self.rows
        for r in range(self.world):
            peers.U[r] = ptrU[r]
            peers.status[r] = ptrS[r]
        iters = torch.empty(self.rows, dtype=torch.int32, device=problem.device)
        outs = _capi.Outputs(None, None, _ptr(iters), None)
        ops = problem.operands()
        stream = ctypes.c_void_p(torch.cuda.current_stream(problem.device).cuda_stream)
        rc = lib.qpmpc_b200_solve_scatter(ctypes.byref(desc), ctypes.byref(ops), ctypes.byref(outs),
                                          ctypes.byref(peers), stream)
        _capi.check(rc, "qpmpc_b200_solve_scatter")
        hU.barrier()  # every rank's stores have landed everywhere
        return U_all, status_all, iters
