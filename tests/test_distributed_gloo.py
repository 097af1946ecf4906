"""Host-side sharding / combination logic on CPU: world_size 2, gloo backend."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyemir_b200.distributed import gather_frames, reduce_mean, shard_range


def test_shard_range_partitions_in_order():
    for nframes in (0, 1, 3, 4, 63, 64, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(nframes, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == nframes
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            for a0, a1 in spans:
                assert a0 % 4 == 0 or a0 == a1    # ABBA cycles are never split
    assert shard_range(4096, 3, 8) == (1536, 2048)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nframes, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        counts = []
        for r in range(world):
            a, b = shard_range(nframes, r, world)
            counts.append(b - a)
        start, stop = shard_range(nframes, rank, world)
        # stand-in for rectified products: frame i is filled with i + row/10
        full = (torch.arange(nframes, dtype=torch.float32)[:, None, None]
                + torch.arange(6, dtype=torch.float32)[None, :, None] / 10
                + torch.zeros(1, 1, 5))
        local = full[start:stop].clone()
        gathered = gather_frames(local, counts, dst=0)
        mean = reduce_mean(local, nframes, dst=0)
        if rank == 0:
            assert torch.equal(gathered, full)
            assert torch.allclose(mean, full.mean(dim=0), rtol=1e-6)
            np.save(os.path.join(result_dir, "ok.npy"), np.array([1]))
        else:
            assert gathered is None and mean is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nframes", [10, 3])
def test_gather_and_mean_world2(tmp_path, nframes):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, nframes, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok.npy")


def _abba_worker(rank, world, port, result_dir):
    """ABBA "mean" combination over two ranks: the partial sums are formed here with torch
    (on a GPU box emirwv_abba_partial does it), the reduction is the code under test."""
    from pyemir_b200 import abba
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full_set = abba.build_full_set("ABBA", 1, 3)                 # 12 frames
        labels = abba.labels_of(full_set)
        start, stop = shard_range(len(full_set), rank, world)
        rng = np.random.default_rng(5)
        stack = torch.from_numpy(rng.normal(size=(len(full_set), 4, 7)))
        sums = torch.zeros((2, 4, 7), dtype=torch.float64)
        for f in range(start, stop):
            sums[0 if labels[f] > 0 else 1] += stack[f]
        reduced = abba.reduce_sums(sums, dst=0)
        if rank == 0:
            want_a = stack[[i for i, c in enumerate(full_set) if c == "A"]].sum(dim=0)
            want_b = stack[[i for i, c in enumerate(full_set) if c == "B"]].sum(dim=0)
            assert torch.allclose(reduced[0], want_a, rtol=0, atol=1e-12)
            assert torch.allclose(reduced[1], want_b, rtol=0, atol=1e-12)
            np.save(os.path.join(result_dir, "abba_ok.npy"), np.array([1]))
        else:
            assert reduced is None
    finally:
        dist.destroy_process_group()


def test_abba_reduce_world2(tmp_path):
    port = _free_port()
    mp.spawn(_abba_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "abba_ok.npy")


def _abba_median_worker(rank, world, port, nframes, height, result_dir):
    """ABBA "median" combination over two ranks: row-slab exchange, per-slab medians (a numpy
    stand-in here; emirwv_abba_median on a GPU box), slab gather."""
    from pyemir_b200 import abba
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full_set = abba.build_full_set("ABBA", 1, -(-nframes // 4))[:nframes]
        labels = abba.labels_of(full_set)
        counts = [shard_range(nframes, r, world)[1] - shard_range(nframes, r, world)[0]
                  for r in range(world)]
        start, stop = shard_range(nframes, rank, world)
        rng = np.random.default_rng(9)
        stack = torch.from_numpy(rng.normal(size=(nframes, height, 7)).astype(np.float32))
        local = stack[start:stop].clone()

        slabs = abba.exchange_row_slabs(local, counts)
        r0, r1 = abba.slab_range(height, rank, world)
        assert torch.equal(slabs, stack[:, r0:r1])

        def median_difference(frames, index_a, index_b):
            def med(index):
                return np.median(frames[index].numpy().astype(np.float64), axis=0)
            return torch.from_numpy(med(index_a).astype(np.float32) -
                                    med(index_b).astype(np.float32))

        got = abba.combine_abba_median(local, labels, counts, dst=0,
                                       median_difference=median_difference)
        if rank == 0:
            from oracle.pipeline import combine_abba_median as oracle_abba
            want = oracle_abba(stack.numpy(), full_set)
            assert np.array_equal(got.numpy(), want)
            np.save(os.path.join(result_dir, "median_ok.npy"), np.array([1]))
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nframes,height", [(12, 9), (8, 1), (3, 6)])
def test_abba_median_world2(tmp_path, nframes, height):
    """(8, 1): one rank gets an empty slab; (3, 6): one rank owns no frame."""
    port = _free_port()
    mp.spawn(_abba_median_worker, args=(2, port, nframes, height, str(tmp_path)), nprocs=2,
             join=True)
    assert os.path.exists(tmp_path / "median_ok.npy")


def _abba_sigmaclip_worker(rank, world, port, result_dir):
    """ABBA "sigmaclip" combination over two ranks: the row-slab exchange of the median path
    with the sigma-clip difference per slab (the oracle's restatement stands in for
    emirwv_abba_sigmaclip on a CPU box)."""
    from oracle.pipeline import combine_abba, sigmaclip_stack
    from pyemir_b200 import abba
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nframes, height = 16, 7
        full_set = abba.build_full_set("ABBA", 2, 2)
        labels = abba.labels_of(full_set)
        counts = [shard_range(nframes, r, world)[1] - shard_range(nframes, r, world)[0]
                  for r in range(world)]
        start, stop = shard_range(nframes, rank, world)
        rng = np.random.default_rng(19)
        data = rng.normal(100.0, 2.0, size=(nframes, height, 9)).astype(np.float32)
        data[5, 3, 4] += 900.0                                    # a cosmic ray in an A image
        stack = torch.from_numpy(data)

        def difference(frames, index_a, index_b):
            x = frames.numpy()
            return torch.from_numpy(sigmaclip_stack(x[index_a], 2.0, 2.0).astype(np.float32)
                                    - sigmaclip_stack(x[index_b], 2.0, 2.0).astype(np.float32))

        got = abba.combine_abba_median(stack[start:stop].clone(), labels, counts, dst=0,
                                       median_difference=difference)
        if rank == 0:
            want = combine_abba(data, full_set, None, "sigmaclip", 2.0, 2.0)
            assert np.array_equal(got.numpy(), want)
            np.save(os.path.join(result_dir, "sigmaclip_ok.npy"), np.array([1]))
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


def test_abba_sigmaclip_world2(tmp_path):
    port = _free_port()
    mp.spawn(_abba_sigmaclip_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "sigmaclip_ok.npy")
