"""ABBA bookkeeping and the "mean" combination of rectified frames on the GPU.

Host-side mirror of the pieces of ``emirdrp/recipes/spec/abba.py`` that sit right behind the
hot path in the ABBA recipe:

* ``build_full_set`` (``abba.py:151-160,236``) and ``get_isky`` (``abba.py:46-75``): which
  frame is an A, which a B, and which frame serves as sky for which (pure bookkeeping,
  pinned by golden vectors produced with the reference's own ``get_isky``);
* ``combine_abba_mean``: ``combine_imgs(list_a, mean)`` and ``combine_imgs(list_b, mean)``
  followed by ``data_a.astype("float32") - data_b.astype("float32")`` (``abba.py:579-606``)
  for frames that live sharded on the GPUs: each rank adds its A and its B frames in
  float64 (``emirwv_abba_partial``), one reduce brings the two sums to the destination rank
  (NCCL; gloo in the CPU tests), ``emirwv_abba_finish`` forms the difference of the means.

* ``shift_image2d`` (``abba.py:556``, numina [recalled]): the per-frame vertical offset is a
  flux-preserving resampling with a pure translation map, i.e. ``rectify2d`` with
  ``aij = [-xoffset, 1, 0]``, ``bij = [-yoffset, 0, 1]`` (GPU, float64).

* ``combine_abba_median``: the same with method "median" (``abba.py:132-137``).  A median
  needs every sample of a pixel on one device, so sharded frames are first exchanged by row
  slabs (rank r receives rows ``slab_range(H, r, world)`` of every frame: one grouped
  send/recv, 1/world of each frame per peer), each rank takes the per-pixel medians of the A
  and of the B images of its slab and their float32 difference in one pass
  (``emirwv_abba_median``: samples in registers, pruned odd-even merge network), and the slabs
  are gathered on the destination rank.

* ``combine_abba_sigmaclip``: the same with method "sigmaclip" (``emirwv_abba_sigmaclip``).

* ``shift_frames``: the vertical offsets of all frames of a device-resident stack in one launch
  (``emirwv_shift_rows``), and ``combine_abba``: offsets + combination + A - B, i.e.
  ``abba.py:545-610`` for a stack that is already on the GPU.

The mean, the median and the sigma clipping of numina's ``combine`` and ``shift_image2d`` are
restated (parity unpinned).
"""

import ctypes

from . import _cabi


def build_full_set(basic_pattern, repeat, nsequences):
    """``"ABBA", 2, 3 -> "AABBBBAA" * 3`` (``abba.py:151-160,236``)."""
    if basic_pattern not in ("AB", "ABBA"):
        raise ValueError(f"Unexpected pattern: {basic_pattern}")
    pattern = ""
    for char in basic_pattern:
        pattern += char * repeat
    return pattern * nsequences


def get_isky(i, basic_pattern, repeat):
    """Index of the image to be employed as sky for image ``i`` (``abba.py:46-75``)."""
    len_pattern = len(basic_pattern * repeat)
    n_previous_sequences = (i + 1) // len_pattern
    if (i + 1) % len_pattern == 0:
        n_previous_sequences -= 1
    ieff = i - n_previous_sequences * len_pattern
    if basic_pattern == "AB":
        first_half = ieff < repeat
    elif basic_pattern == "ABBA":
        first_half = ieff < repeat or 2 * repeat <= ieff < 3 * repeat
    else:
        raise ValueError(f"Unexpected pattern: {basic_pattern}")
    return i + repeat if first_half else i - repeat


def shift_image2d(image2d, xoffset=0.0, yoffset=0.0, resampling=2):
    """Shift an image by fractional (or integer) offsets, flux preserving by default.

    ``reduced_mos_image[0].data = shift_image2d(data, yoffset=-offset)`` in the recipe
    (``abba.py:556``); the pixel at (x, y) of the result comes from (x - xoffset,
    y - yoffset) of the input.
    """
    from .slitlet2d import rectify2d
    return rectify2d(image2d, [-float(xoffset), 1.0, 0.0], [-float(yoffset), 0.0, 1.0],
                     resampling)


def labels_of(full_set):
    """int8 labels of ``emirwv_abba_partial``: +1 for A, -1 for B."""
    out = []
    for char in full_set:
        if char == "A":
            out.append(1)
        elif char == "B":
            out.append(-1)
        else:
            raise ValueError("Unexpected char value: {}".format(char))
    return out


_LABEL_CACHE = {}


def _labels_on(device, labels):
    """int8 label vector on the device (cached: the upload would synchronise every call)."""
    import torch
    key = (str(device), labels)
    if key not in _LABEL_CACHE:
        if len(_LABEL_CACHE) > 64:
            _LABEL_CACHE.clear()
        _LABEL_CACHE[key] = torch.tensor(labels, dtype=torch.int8, device=device)
    return _LABEL_CACHE[key]


def partial_sums(frames, labels, sums=None):
    """float64 sums of this rank's A frames and B frames: a (2, H, W) CUDA tensor.

    frames: (n, H, W) contiguous float32/float64 CUDA tensor; labels: n ints (+1 A, -1 B,
    0 unused).  ``sums`` accumulates over several calls (chunked stacks).
    """
    import torch
    if frames.dim() != 3 or not frames.is_cuda or not frames.is_contiguous() or \
            frames.dtype not in (torch.float32, torch.float64):
        raise ValueError("frames must be a contiguous (n, H, W) float32/float64 CUDA tensor")
    n, height, width = frames.shape
    if len(labels) != n:
        raise ValueError("one label per frame is needed")
    accumulate = sums is not None
    if sums is None:
        sums = torch.empty((2, height, width), dtype=torch.float64, device=frames.device)
    lab = _labels_on(frames.device, tuple(int(v) for v in labels))
    stream = torch.cuda.current_stream(frames.device).cuda_stream
    vp = ctypes.c_void_p
    _cabi.check(_cabi.load().emirwv_abba_partial(
        vp(frames.data_ptr()), vp(lab.data_ptr()), n, height * width,
        _cabi.F32 if frames.dtype == torch.float32 else _cabi.F64,
        vp(sums[0].data_ptr()), vp(sums[1].data_ptr()), int(accumulate),
        vp(stream) if stream else None))
    return sums


def reduce_sums(sums, dst=0):
    """Add the (2, H, W) partial sums of every rank on ``dst`` (None elsewhere)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(sums, dst=dst, op=dist.ReduceOp.SUM)
        if dist.get_rank() != dst:
            return None
    return sums


def finish(sums, count_a, count_b):
    """float32(mean A) - float32(mean B) from the reduced sums (``abba.py:604-606``)."""
    import torch
    if count_a < 1 or count_b < 1:
        raise ValueError("both A and B images are needed")
    out = torch.empty(sums.shape[1:], dtype=torch.float32, device=sums.device)
    stream = torch.cuda.current_stream(sums.device).cuda_stream
    vp = ctypes.c_void_p
    _cabi.check(_cabi.load().emirwv_abba_finish(
        vp(sums[0].data_ptr()), vp(sums[1].data_ptr()), int(count_a), int(count_b),
        sums.shape[1] * sums.shape[2], vp(out.data_ptr()), vp(stream) if stream else None))
    return out


def combine_abba_mean(frames_local, labels_local, count_a, count_b, dst=0):
    """A - B image of an ABBA sequence whose rectified frames are sharded over the ranks.

    frames_local / labels_local: this rank's frames and their labels; count_a / count_b:
    number of A and B frames of the whole sequence.  Returns the (H, W) float32 image on
    ``dst`` and None on the other ranks.
    """
    sums = reduce_sums(partial_sums(frames_local, labels_local), dst)
    if sums is None:
        return None
    return finish(sums, count_a, count_b)


# ------------------------------------------------------------------ method "median"
STACK_MAX = 64      # images per nodding position in one emirwv_stack_median call


def stack_median(frames, index):
    """Per-pixel median over ``frames[index]``: a float64 CUDA tensor of the image shape.

    frames: (n, H, W) contiguous float32/float64 CUDA tensor; index: the images that take
    part (at most 64).  numpy.median semantics (mean of the two middle values, NaN if any
    sample is NaN); float64 result like numina's ``combine``.
    """
    import numpy as np
    import torch
    if frames.dim() != 3 or not frames.is_cuda or not frames.is_contiguous() or \
            frames.dtype not in (torch.float32, torch.float64):
        raise ValueError("frames must be a contiguous (n, H, W) float32/float64 CUDA tensor")
    n, height, width = frames.shape
    idx = np.ascontiguousarray(index, dtype=np.int32)
    out = torch.empty((height, width), dtype=torch.float64, device=frames.device)
    stream = torch.cuda.current_stream(frames.device).cuda_stream
    vp = ctypes.c_void_p
    _cabi.check(_cabi.load().emirwv_stack_median(
        vp(frames.data_ptr()), n, idx.ctypes.data_as(vp), int(idx.size), height * width,
        _cabi.F32 if frames.dtype == torch.float32 else _cabi.F64, vp(out.data_ptr()),
        vp(stream) if stream else None))
    return out


def slab_range(height, rank, world):
    """Rows [start, stop) of every image whose samples end up on ``rank``."""
    per_rank = -(-height // world)
    return min(height, rank * per_rank), min(height, (rank + 1) * per_rank)


def exchange_row_slabs(frames_local, counts):
    """(sum(counts), slab rows, W): this rank's row slab of EVERY frame, in frame order.

    frames_local: (counts[rank], H, W); counts: frames owned by every rank (shard_range).
    One grouped send/recv: each rank sends rows slab_range(H, r, world) of its frames to
    rank r.  No arithmetic.
    """
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return frames_local
    rank, world = dist.get_rank(), dist.get_world_size()
    height, width = frames_local.shape[1:]
    r0, r1 = slab_range(height, rank, world)
    out = torch.empty((sum(counts), r1 - r0, width), dtype=frames_local.dtype,
                      device=frames_local.device)
    ops, keep, offset = [], [], 0
    for peer in range(world):
        n = counts[peer]
        if n and r1 > r0:
            if peer == rank:
                out[offset:offset + n].copy_(frames_local[:, r0:r1])
            else:
                ops.append(dist.P2POp(dist.irecv, out[offset:offset + n], peer))
        offset += n
        p0, p1 = slab_range(height, peer, world)
        if peer != rank and counts[rank] and p1 > p0:
            keep.append(frames_local[:, p0:p1].contiguous())
            ops.append(dist.P2POp(dist.isend, keep[-1], peer))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out


def abba_median_difference(frames, index_a, index_b):
    """float32(median of frames[index_a]) - float32(median of frames[index_b]), one pass.

    ``emirwv_abba_median``: both medians of a pixel are taken in registers and only the
    float32 difference (``abba.py:604-606``) is written.
    """
    import numpy as np
    import torch
    if frames.dim() != 3 or not frames.is_cuda or not frames.is_contiguous() or \
            frames.dtype not in (torch.float32, torch.float64):
        raise ValueError("frames must be a contiguous (n, H, W) float32/float64 CUDA tensor")
    n, height, width = frames.shape
    ia = np.ascontiguousarray(index_a, dtype=np.int32)
    ib = np.ascontiguousarray(index_b, dtype=np.int32)
    out = torch.empty((height, width), dtype=torch.float32, device=frames.device)
    stream = torch.cuda.current_stream(frames.device).cuda_stream
    vp = ctypes.c_void_p
    _cabi.check(_cabi.load().emirwv_abba_median(
        vp(frames.data_ptr()), n, ia.ctypes.data_as(vp), int(ia.size), ib.ctypes.data_as(vp),
        int(ib.size), height * width,
        _cabi.F32 if frames.dtype == torch.float32 else _cabi.F64, vp(out.data_ptr()),
        vp(stream) if stream else None))
    return out


def abba_sigmaclip_difference(frames, index_a, index_b, low=3.0, high=3.0):
    """float32(clipped mean of frames[index_a]) - float32(clipped mean of frames[index_b])
    (``emirwv_abba_sigmaclip``; numina ``combine.sigmaclip`` restated, at most 64 images per
    nodding position)."""
    import numpy as np
    import torch
    if frames.dim() != 3 or not frames.is_cuda or not frames.is_contiguous() or \
            frames.dtype not in (torch.float32, torch.float64):
        raise ValueError("frames must be a contiguous (n, H, W) float32/float64 CUDA tensor")
    n, height, width = frames.shape
    ia = np.ascontiguousarray(index_a, dtype=np.int32)
    ib = np.ascontiguousarray(index_b, dtype=np.int32)
    out = torch.empty((height, width), dtype=torch.float32, device=frames.device)
    stream = torch.cuda.current_stream(frames.device).cuda_stream
    vp = ctypes.c_void_p
    _cabi.check(_cabi.load().emirwv_abba_sigmaclip(
        vp(frames.data_ptr()), n, ia.ctypes.data_as(vp), int(ia.size), ib.ctypes.data_as(vp),
        int(ib.size), height * width,
        _cabi.F32 if frames.dtype == torch.float32 else _cabi.F64, float(low), float(high),
        vp(out.data_ptr()), vp(stream) if stream else None))
    return out


def stack_sigmaclip(frames, index, low=3.0, high=3.0):
    """Per-pixel sigma-clipped mean over ``frames[index]``: a float64 CUDA tensor."""
    import numpy as np
    import torch
    if frames.dim() != 3 or not frames.is_cuda or not frames.is_contiguous() or \
            frames.dtype not in (torch.float32, torch.float64):
        raise ValueError("frames must be a contiguous (n, H, W) float32/float64 CUDA tensor")
    n, height, width = frames.shape
    idx = np.ascontiguousarray(index, dtype=np.int32)
    out = torch.empty((height, width), dtype=torch.float64, device=frames.device)
    stream = torch.cuda.current_stream(frames.device).cuda_stream
    vp = ctypes.c_void_p
    _cabi.check(_cabi.load().emirwv_stack_sigmaclip(
        vp(frames.data_ptr()), n, idx.ctypes.data_as(vp), int(idx.size), height * width,
        _cabi.F32 if frames.dtype == torch.float32 else _cabi.F64, float(low), float(high),
        vp(out.data_ptr()), vp(stream) if stream else None))
    return out


def shift_frames(frames, yoffsets, out=None, out_dtype=None):
    """``shift_image2d(frame, yoffset=yoffsets[i])`` for every frame of a CUDA stack, one launch.

    frames: (n, H, W) contiguous float32/float64 CUDA tensor; yoffsets: n floats (the recipe
    passes ``-offset``, ``abba.py:556``); frames whose offset is 0 are copied.  The result is
    float32 by default, like the recipe's ``.astype("float32")``.
    """
    import numpy as np
    import torch
    if frames.dim() != 3 or not frames.is_cuda or not frames.is_contiguous() or \
            frames.dtype not in (torch.float32, torch.float64):
        raise ValueError("frames must be a contiguous (n, H, W) float32/float64 CUDA tensor")
    n, height, width = frames.shape
    off = np.ascontiguousarray(yoffsets, dtype=np.float64)
    if off.shape != (n,):
        raise ValueError("one offset per frame is needed")
    out_dtype = out_dtype or torch.float32
    if out is None:
        out = torch.empty((n, height, width), dtype=out_dtype, device=frames.device)
    code = {torch.float32: _cabi.F32, torch.float64: _cabi.F64}
    stream = torch.cuda.current_stream(frames.device).cuda_stream
    vp = ctypes.c_void_p
    _cabi.check(_cabi.load().emirwv_shift_rows(
        vp(frames.data_ptr()), code[frames.dtype], vp(out.data_ptr()), code[out.dtype], n,
        height, width, off.ctypes.data_as(vp), vp(stream) if stream else None))
    return out


def combine_abba(frames, full_set, offsets=None, method="mean", method_kwargs=None):
    """``abba.py:545-610`` for a device-resident stack of rectified frames: the per-frame
    vertical offsets (``shift_image2d(data, yoffset=-offset).astype("float32")`` when the offset
    is not zero), ``combine_imgs`` of the A and of the B images with ``method`` ("mean",
    "median" or "sigmaclip", ``abba.py:132-137``), float32(A) - float32(B).

    frames: (n, H, W) CUDA tensor; full_set: "ABBA..." (``build_full_set``); offsets: n floats
    or None.  Returns the (H, W) float32 CUDA tensor.  Median and sigma clipping take at most
    64 images per nodding position.
    """
    labels = labels_of(full_set)
    if len(labels) != frames.shape[0]:
        raise ValueError("one label per frame is needed")
    kwargs = dict(method_kwargs or {})
    if offsets is not None and any(float(o) != 0.0 for o in offsets):
        frames = shift_frames(frames, [-float(o) for o in offsets])
    index_a = [i for i, lab in enumerate(labels) if lab > 0]
    index_b = [i for i, lab in enumerate(labels) if lab < 0]
    if not index_a or not index_b:
        raise ValueError("both A and B images are needed")
    if method == "mean":
        return finish(partial_sums(frames, labels), len(index_a), len(index_b))
    if method == "median":
        return abba_median_difference(frames, index_a, index_b)
    if method == "sigmaclip":
        return abba_sigmaclip_difference(frames, index_a, index_b, kwargs.get("low", 3.0),
                                         kwargs.get("high", 3.0))
    raise ValueError(f"Unexpected method: {method}")


def combine_abba_sigmaclip(frames_local, labels, counts, dst=0, low=3.0, high=3.0):
    """sigmaclip(A images) - sigmaclip(B images) of a sequence sharded over the ranks: the
    row-slab exchange of :func:`combine_abba_median` with ``emirwv_abba_sigmaclip`` per slab."""
    def difference(stack, index_a, index_b):
        return abba_sigmaclip_difference(stack, index_a, index_b, low, high)
    return combine_abba_median(frames_local, labels, counts, dst, median_difference=difference)


def combine_abba_median(frames_local, labels, counts, dst=0,
                        median_difference=abba_median_difference):
    """median(A images) - median(B images) of a sequence sharded over the ranks.

    frames_local: this rank's rectified frames; labels: the labels of ALL frames of the
    sequence (labels_of(full_set)); counts: frames owned by every rank.  Returns the (H, W)
    float32 image on ``dst`` and None elsewhere.
    ``median_difference(stack, index_a, index_b) -> float32 image`` is the CUDA entry point
    (emirwv_abba_median); the CPU (gloo) tests of the exchange pass a torch stand-in.
    """
    import torch
    from .distributed import gather_frames
    import torch.distributed as dist
    index_a = [i for i, lab in enumerate(labels) if lab > 0]
    index_b = [i for i, lab in enumerate(labels) if lab < 0]
    if not index_a or not index_b:
        raise ValueError("both A and B images are needed")
    if len(labels) != sum(counts):
        raise ValueError("one label per frame of the sequence is needed")
    world = dist.get_world_size() if dist.is_initialized() else 1
    slabs = exchange_row_slabs(frames_local, counts)
    if slabs.shape[1]:
        diff = median_difference(slabs, index_a, index_b)
    else:
        diff = torch.empty((0, slabs.shape[2]), dtype=torch.float32, device=slabs.device)
    if world == 1:
        return diff
    height = frames_local.shape[1]
    rows = [slab_range(height, r, world)[1] - slab_range(height, r, world)[0]
            for r in range(world)]
    return gather_frames(diff, rows, dst)
