#!/usr/bin/env python3
"""Aggregate host<->device copy bandwidth of the box, all ranks at once (the ceiling of the
end-to-end leg of bench.py at N GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29517 tools/pcie_probe.py [--mb 256] [--reps 20]

Every rank copies --mb MiB host->device and (208:132, the byte ratio of bench config 2) device->host
from / to page-locked memory on two streams at the same time; the ranks start together
(barrier) and the slowest one sets the time.  Rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=256)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n_in = a.mb << 20
n_out = n_in * 132 // 208
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
d_out = torch.ones(n_out, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h):
    for timed in (False, True):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        for _ in range(a.reps if timed else 2):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    nbytes = a.reps * ((n_in if h2d else 0) + (n_out if d2h else 0))
    return {"per_rank_gbs": nbytes / ms / 1e6, "aggregate_gbs": world * nbytes / ms_max / 1e6}


out = {"n_gpus": world, "mb_h2d": a.mb, "h2d_only": run(True, False), "d2h_only": run(False, True),
       "both_208_132": run(True, True)}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
