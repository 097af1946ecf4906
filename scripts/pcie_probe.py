#!/usr/bin/env python
"""Host <-> device copy bandwidth per rank, alone and all ranks together, with and without the
NUMA placement of pyemir_b200.distributed.bind_to_gpu_numa (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29512 scripts/pcie_probe.py

What limits the end-to-end rate of N ranks on one box: every rank moves 16.8 MB in and 28.4 MB
out per frame through page-locked host memory.  Rank 0 prints one JSON line.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyemir_b200.distributed import bind_to_gpu_numa, gpu_numa_node, init_from_env  # noqa: E402

NBYTES = 1 << 30


def measure(dev, h_in, h_out, d_in, d_out, both, h2d=True):
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 4
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if both or not h2d:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return reps * NBYTES / dt / 1e9


def main():
    placement = {"before_bind": None}
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    rank, world, local_rank = init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    d_in = torch.empty(NBYTES, dtype=torch.uint8, device=dev)
    d_out = torch.empty(NBYTES, dtype=torch.uint8, device=dev)
    results = {}
    for label in ("default placement", "numa-bound placement"):
        if label.startswith("numa"):
            placement = bind_to_gpu_numa(local_rank)
        h_in = torch.empty(NBYTES, dtype=torch.uint8, pin_memory=True)
        h_out = torch.empty(NBYTES, dtype=torch.uint8, pin_memory=True)
        h_in.fill_(1)
        h_out.fill_(1)
        measure(dev, h_in, h_out, d_in, d_out, True)            # warm-up
        rows = {}
        # every rank alone, one after the other
        alone = torch.zeros(world, dtype=torch.float64, device=dev)
        for r in range(world):
            dist.barrier()
            if r == rank:
                alone[r] = measure(dev, h_in, h_out, d_in, d_out, True)
        dist.all_reduce(alone)
        rows["alone_GBs_each_direction"] = [round(v, 1) for v in alone.tolist()]
        # all ranks at once: both directions, then host -> device only
        for name, both, h2d in (("together_GBs_each_direction", True, True),
                                ("together_h2d_only_GBs", False, True),
                                ("together_d2h_only_GBs", False, False)):
            dist.barrier()
            val = torch.zeros(world, dtype=torch.float64, device=dev)
            val[rank] = measure(dev, h_in, h_out, d_in, d_out, both, h2d)
            dist.all_reduce(val)
            rows[name] = [round(v, 1) for v in val.tolist()]
            rows[name + "_sum"] = round(float(val.sum()), 1)
        results[label] = rows
        del h_in, h_out
    nodes = torch.zeros(world, dtype=torch.int64, device=dev)
    node = gpu_numa_node(local_rank)
    nodes[rank] = -1 if node is None else node
    dist.all_reduce(nodes)
    if rank == 0:
        print(json.dumps({"world": world, "gpu_numa_nodes": nodes.tolist(), "cpus": os.cpu_count(),
                          "rank0_placement": placement, "copy_GBs": results,
                          "bytes_per_copy": NBYTES}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
