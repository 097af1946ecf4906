"""Sharding of independent exposures over one process per GPU.

The reference has nothing distributed; its natural batch is the ABBA recipe's
per-frame loop (abba.py:351-381): N frames, one shared RectWaveCoeff, independent
until ``combine_imgs`` (abba.py:579-601).  Frames are therefore split in contiguous
chunks over ranks with no traffic while rectifying.  A collective runs only when a
combined product is asked for:

* ``mean``  : per-rank partial sums + one ``reduce`` (28 MB per rank, not N x 28 MB);
* ``gather``: rectified frames gathered on one rank (median / sigma-clip need all
  samples of a pixel on one device).

``torch.distributed`` is the plumbing (NCCL on GPUs, gloo in the CPU tests); the
functions are backend-agnostic and never touch pixel arithmetic themselves.
"""

import os
import subprocess

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Join the torchrun rendezvous if there is one; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local_rank


def shard_range(nframes, rank, world, multiple=4):
    """Contiguous [start, stop) of the frames owned by ``rank``.

    Chunk boundaries fall on multiples of ``multiple`` frames (4 keeps every ABBA
    cycle on one device so that A and B partial sums can be formed locally); the
    last ranks may own fewer frames, or none.
    """
    if world < 1 or not 0 <= rank < world:
        raise ValueError("invalid rank / world size")
    groups = -(-nframes // multiple)
    per_rank = -(-groups // world)
    start = min(nframes, rank * per_rank * multiple)
    stop = min(nframes, (rank + 1) * per_rank * multiple)
    return start, stop


def reduce_mean(local_frames, total_frames, dst=0):
    """Mean over all frames of all ranks from per-rank partial sums.

    local_frames: (n_local, H, W) tensor (may be empty).  Returns the (H, W) mean on
    ``dst`` (None elsewhere).  One reduce of one image per rank.
    """
    partial = local_frames.sum(dim=0, dtype=torch.float64 if local_frames.dtype ==
                               torch.float64 else torch.float32)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(partial, dst=dst, op=dist.ReduceOp.SUM)
        if dist.get_rank() != dst:
            return None
    return partial / float(total_frames)


def gather_frames(local_frames, counts, dst=0):
    """Gather rectified frames of all ranks on ``dst`` in frame order.

    counts: frames owned by every rank (from :func:`shard_range`).  Returns the
    (sum(counts), H, W) tensor on ``dst`` and None elsewhere.  Ranks own different
    numbers of frames, so the exchange is a grouped send/recv (NCCL fuses it).
    """
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_frames
    rank, world = dist.get_rank(), dist.get_world_size()
    shape = tuple(local_frames.shape[1:])
    if rank == dst:
        total = sum(counts)
        out = torch.empty((total,) + shape, dtype=local_frames.dtype,
                          device=local_frames.device)
        ops, offset = [], 0
        for src in range(world):
            n = counts[src]
            if n:
                if src == dst:
                    out[offset:offset + n].copy_(local_frames)
                else:
                    ops.append(dist.P2POp(dist.irecv, out[offset:offset + n], src))
            offset += n
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return out
    if counts[rank]:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, local_frames.contiguous(), dst)]):
            req.wait()
    return None


def rectify_sharded(frames_local, rectwv_coeff, args_resampling=2, dtype="float32",
                    max_order=4, combine=None, total_frames=None, counts=None, dst=0):
    """Rectify this rank's frames on its GPU; optionally combine across ranks.

    frames_local: CUDA tensor (n_local, 2048, 2048) of this rank's shard.
    combine: None (products stay sharded), "mean" or "gather".
    Returns (out_local, jminmax_local, combined or None).
    """
    from .plan import get_plan
    device = frames_local.device.index
    plan = get_plan(rectwv_coeff, args_resampling, dtype, max_order, device=device)
    if frames_local.shape[0]:
        out, jmm = plan.run_torch(frames_local)
    else:
        out = frames_local.new_empty((0,) + plan.out_shape)
        jmm = torch.empty((0, 55, 2), dtype=torch.int32, device=frames_local.device)
    combined = None
    if combine == "mean":
        combined = reduce_mean(out, total_frames, dst)
    elif combine == "gather":
        combined = gather_frames(out, counts, dst)
    elif combine is not None:
        raise ValueError("combine must be None, 'mean' or 'gather'")
    return out, jmm, combined


# --------------------------------------------------------------------------- host placement
def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(index):
    """NUMA node the GPU ``index`` hangs off (sysfs, then ``nvidia-smi topo``), or None."""
    try:
        props = torch.cuda.get_device_properties(index)
        bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as fd:
            node = int(fd.read().strip())
        if node >= 0:
            return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        pass
    try:
        text = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True,
                              timeout=20).stdout
        header = None
        for line in text.splitlines():
            cells = [c.strip() for c in line.split("\t") if c.strip()]
            if not cells:
                continue
            if header is None and any("NUMA Affinity" in c for c in cells):
                header = cells
                continue
            if header is not None and cells[0] == f"GPU{index}":
                # the header has no cell above the row labels
                col = [i for i, c in enumerate(header) if "NUMA Affinity" in c][0] + 1
                if col < len(cells):
                    return int(cells[col].split("-")[0].split(",")[0])
    except (OSError, ValueError, IndexError, subprocess.SubprocessError):
        pass
    return None


def bind_to_gpu_numa(index):
    """Run this process on the CPUs of the NUMA node of GPU ``index`` and prefer that node's
    memory, so that the page-locked buffers it allocates AFTERWARDS sit next to the PCIe root
    the GPU hangs off (one process per GPU: without it every rank's buffers may land on the node
    the launcher started on, and all copies of all GPUs go through one node's memory and the
    socket interconnect).  Returns what was done, for the record.  ``EMIRWV_NUMA_BIND=0`` skips it.
    """
    info = {"numa_node": None, "cpus": None, "mempolicy": None}
    if os.environ.get("EMIRWV_NUMA_BIND", "1") == "0":
        info["skipped"] = "EMIRWV_NUMA_BIND=0"
        return info
    node = gpu_numa_node(index)
    info["numa_node"] = node
    if node is None:
        return info
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fd:
            cpus = _parse_cpulist(fd.read())
        allowed = os.sched_getaffinity(0)
        cpus = (cpus & allowed) or allowed
        os.sched_setaffinity(0, cpus)
        info["cpus"] = len(cpus)
    except (OSError, ValueError):
        pass
    try:
        import ctypes
        libc = ctypes.CDLL(None, use_errno=True)
        mpol_preferred, nr_set_mempolicy = 1, 238            # x86-64
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(nr_set_mempolicy, mpol_preferred, ctypes.byref(mask),
                          ctypes.c_ulong(8 * ctypes.sizeof(mask)))
        info["mempolicy"] = "preferred" if rc == 0 else f"errno {ctypes.get_errno()}"
    except (OSError, AttributeError):
        pass
    return info
