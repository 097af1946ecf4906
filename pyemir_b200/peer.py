"""The ABBA "mean" combination fused with its exchange over NVLink peer memory.

``recipes/spec/abba.py:579-606`` for a sequence whose rectified frames are sharded over one
process per GPU of one box: instead of per-rank partial sums followed by an NCCL ``reduce`` and
a finishing kernel on one rank (``abba.combine_abba_mean``), the kernel that adds a rank's frames
pushes every float64 sum straight into the buffer of the rank that owns the pixel (stores over
NVLink, opened through CUDA IPC), the owners finish their slabs and store them into the image on
the destination rank.  ``torch.distributed`` only carries the 64-byte IPC handles, once.
"""

import ctypes

from . import _cabi


class PeerWorkspace:
    """Peer-memory workspace of :func:`combine_abba_mean_fused` (one per rank and image size)."""

    def __init__(self, npix, device=None):
        import torch
        import torch.distributed as dist
        self._lib = _cabi.load()
        self._handle = ctypes.c_void_p()
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.npix = int(npix)
        self.device = torch.cuda.current_device() if device is None else int(device)
        ipc = ctypes.create_string_buffer(64)
        # every rank goes through the same exchanges whatever happens locally: a rank that failed
        # must not leave the others waiting in a collective
        error = None
        with torch.cuda.device(self.device):
            try:
                _cabi.check(self._lib.emirwv_peer_create(self.npix, self.rank, self.world,
                                                         ctypes.byref(self._handle), ipc))
            except Exception as exc:                              # noqa: BLE001
                error = exc
            handles = [None] * self.world
            if self.world > 1:
                dist.all_gather_object(handles, ipc.raw if error is None else None)
            else:
                handles[0] = ipc.raw
            if error is None and all(h is not None for h in handles):
                try:
                    blob = ctypes.create_string_buffer(b"".join(handles), 64 * self.world)
                    _cabi.check(self._lib.emirwv_peer_connect(self._handle, blob))
                except Exception as exc:                          # noqa: BLE001
                    error = exc
            elif error is None:
                error = RuntimeError("another rank could not create its peer workspace")
            ok = torch.tensor([0 if error is not None else 1], dtype=torch.int32,
                              device=torch.device("cuda", self.device))
            if self.world > 1:
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)     # (also: every rank has connected)
        if int(ok.item()) == 0:
            if self._handle.value:
                self._lib.emirwv_peer_destroy(self._handle)
                self._handle = ctypes.c_void_p()
            raise error if error is not None else RuntimeError(
                "another rank could not connect its peer workspace")

    def close(self):
        if getattr(self, "_handle", None) and self._handle.value:
            import torch.distributed as dist
            # nobody still writes into a workspace, nobody keeps one open that is about to be freed
            if self.world > 1 and dist.is_initialized():
                dist.barrier()
            self._lib.emirwv_peer_disconnect(self._handle)
            if self.world > 1 and dist.is_initialized():
                dist.barrier()
            self._lib.emirwv_peer_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def timed_out(self):
        return bool(_cabi.check(self._lib.emirwv_peer_timed_out(self._handle)))


def combine_abba_mean_fused(frames_local, labels_local, count_a, count_b, workspace, dst=0,
                            sums=None):
    """A - B image (method "mean") of a sequence sharded over the ranks, exchange fused into the
    summing kernel.  Same arguments and result as ``abba.combine_abba_mean`` plus the
    workspace; ``sums``: the (2, H, W) float64 sums of earlier passes of this rank
    (``abba.partial_sums``).  Returns the (H, W) float32 image on ``dst`` (a view into the
    workspace, overwritten by the next call) and None elsewhere."""
    import torch
    from .abba import _labels_on
    if frames_local.dim() != 3 or not frames_local.is_cuda or not frames_local.is_contiguous() or \
            frames_local.dtype not in (torch.float32, torch.float64):
        raise ValueError("frames must be a contiguous (n, H, W) float32/float64 CUDA tensor")
    n, height, width = frames_local.shape
    if height * width != workspace.npix:
        raise ValueError("the workspace was made for another image size")
    if len(labels_local) != n:
        raise ValueError("one label per frame is needed")
    lab = _labels_on(frames_local.device, tuple(int(v) for v in labels_local)) if n else None
    stream = torch.cuda.current_stream(frames_local.device).cuda_stream
    vp = ctypes.c_void_p
    image = vp()
    _cabi.check(workspace._lib.emirwv_abba_mean_push(
        workspace._handle, vp(frames_local.data_ptr()) if n else None,
        vp(lab.data_ptr()) if n else None, n,
        _cabi.F32 if frames_local.dtype == torch.float32 else _cabi.F64,
        vp(sums[0].data_ptr()) if sums is not None else None,
        vp(sums[1].data_ptr()) if sums is not None else None,
        int(count_a), int(count_b), int(dst), ctypes.byref(image), vp(stream) if stream else None))
    if not image.value:
        return None
    return _as_tensor(image.value, (height, width), frames_local.device)


def _as_tensor(ptr, shape, device):
    """float32 CUDA tensor over memory the library owns (no copy)."""
    import torch

    class _Holder:
        pass
    holder = _Holder()
    n = shape[0] * shape[1]
    holder.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(ptr), False),
                                       "version": 2}
    return torch.as_tensor(holder, device=device).view(shape)
