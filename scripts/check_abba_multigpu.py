#!/usr/bin/env python
"""ABBA combinations over NCCL, checked against numpy (run under torchrun, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/check_abba_multigpu.py

Every rank holds a contiguous shard of a seeded stack of "rectified products"; the mean, the
median and the sigma-clip combination (partial sums + reduce; row-slab exchange +
emirwv_abba_median / emirwv_abba_sigmaclip + slab gather) must reproduce the single-process
numpy result bit for bit on rank 0.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.pipeline import combine_abba, combine_abba_mean, combine_abba_median  # noqa: E402  (checker)
from pyemir_b200 import abba  # noqa: E402
from pyemir_b200.distributed import init_from_env, shard_range  # noqa: E402


def main():
    rank, world, local_rank = init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    results = []
    for nseq, shape in ((3, (77, 516)), (6, (2090, 1270)), (1, (5, 33))):
        full_set = abba.build_full_set("ABBA", 1, nseq * world)
        n = len(full_set)
        labels = abba.labels_of(full_set)
        rng = np.random.default_rng(1234 + nseq)
        stack = rng.normal(300.0, 100.0, size=(n,) + shape).astype(np.float32)
        hits = rng.integers(0, [n, shape[0], shape[1]], size=(200, 3))
        stack[hits[:, 0], hits[:, 1], hits[:, 2]] += 5000.0        # cosmic rays
        counts = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
        start, stop = shard_range(n, rank, world)
        local = torch.from_numpy(stack[start:stop]).cuda()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        med = abba.combine_abba_median(local, labels, counts, dst=0)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        mean = abba.combine_abba_mean(local, labels[start:stop], full_set.count("A"),
                                      full_set.count("B"), dst=0)
        clip = abba.combine_abba_sigmaclip(local, labels, counts, dst=0, low=3.0, high=2.5)
        # the mean again, exchange fused into the summing kernel over NVLink peer memory; twice
        # (buffer parities), the second time in two passes with carried sums
        from pyemir_b200.peer import PeerWorkspace, combine_abba_mean_fused
        ws = PeerWorkspace(shape[0] * shape[1])
        fused = combine_abba_mean_fused(local, labels[start:stop], full_set.count("A"),
                                        full_set.count("B"), ws, dst=0)
        fused1 = fused.clone() if fused is not None else None
        half = (stop - start) // 2
        carried = abba.partial_sums(local[:half].contiguous(), labels[start:start + half]) if half else None
        fused = combine_abba_mean_fused(local[half:].contiguous(), labels[start + half:stop],
                                        full_set.count("A"), full_set.count("B"), ws, dst=0,
                                        sums=carried)
        fused2 = fused.clone() if fused is not None else None
        torch.cuda.synchronize()
        timed_out = ws.timed_out()
        ws.close()
        if rank == 0:
            ok_clip = bool(np.array_equal(
                clip.cpu().numpy(), combine_abba(stack, full_set, None, "sigmaclip", 3.0, 2.5)))
            want_mean = combine_abba_mean(stack, full_set)
            ok_fused = bool(not timed_out and np.array_equal(fused1.cpu().numpy(), want_mean)
                            and np.array_equal(fused2.cpu().numpy(), want_mean))
            ok_med = bool(np.array_equal(med.cpu().numpy(), combine_abba_median(stack, full_set)))
            ok_mean = bool(np.array_equal(mean.cpu().numpy(), combine_abba_mean(stack, full_set)))
            results.append({"frames": n, "shape": list(shape), "world": world,
                            "median_bit_exact": ok_med, "mean_bit_exact": ok_mean,
                            "sigmaclip_bit_exact": ok_clip, "fused_mean_bit_exact": ok_fused,
                            "median_first_call_ms": 1e3 * (t1 - t0)})
        else:
            assert med is None and mean is None and clip is None and fused is None
            assert not timed_out
    if rank == 0:
        print(json.dumps(results), flush=True)
        if not all(r["median_bit_exact"] and r["mean_bit_exact"] and r["sigmaclip_bit_exact"]
                   and r["fused_mean_bit_exact"] for r in results):
            raise SystemExit("ABBA combination over NCCL differs from the single-process result")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
