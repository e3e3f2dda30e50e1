"""Two-GPU check of the fused gather (kernel epilogue stores over NVLink):
every rank ends with the U rows of all ranks, bit-identical to a one-GPU
solve of the whole batch.  Skipped on boxes with a single GPU."""

import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from qpmpc_b200 import solve_mpc_batch
    from qpmpc_b200.distributed import PeerGather, gather_plans
    from qpmpc_b200.workloads import slice_workload, to_batched, triple_integrator_batch

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    B = 4096
    w = triple_integrator_batch(world * B, seed=7)
    mine = to_batched(slice_workload(w, rank * B, (rank + 1) * B))
    gather = PeerGather(B, 16)
    for _ in range(3):  # repeated use of the same symmetric buffers
        U_all, st_all, _ = gather.solve(mine)
    torch.cuda.synchronize()
    plan = solve_mpc_batch(mine)
    U_nccl, st_nccl = gather_plans(plan.inputs.reshape(B, -1), plan.status, world * B)
    torch.cuda.synchronize()
    full = solve_mpc_batch(to_batched(w))
    torch.cuda.synchronize()
    ok = (torch.equal(U_all, full.inputs.reshape(world * B, -1)) and torch.equal(U_all, U_nccl)
          and torch.equal(st_all, st_nccl) and bool((st_all == 0).all()))
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([ok]))
    dist.barrier()
    dist.destroy_process_group()


def test_fused_gather_two_gpus(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(2, 29533, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert bool(np.load(tmp_path / f"ok{r}.npy")[0]), f"rank {r}"
