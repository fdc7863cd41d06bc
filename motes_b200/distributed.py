"""Multi-GPU plumbing for the extraction path: frames are independent end to end, so the
only exchange is ONE all-gather of the per-frame outputs (spectra, limits, bin parameters; ~64 B/column).

One process per GPU (``torchrun``); ``torch.distributed`` with NCCL on GPUs, gloo in CPU tests.

``GatherPlan`` lays the selected outputs of a rank's shard out in one flat buffer: the C ABI writes each output
straight into its slice (``plan.local(name)`` is a contiguous ``(frames, ...)`` view, so ``motes_run_frames_dev``
needs no packing copy), ``plan.gather()`` is a single ``all_gather_into_tensor`` over NVLink, and
``plan.gathered(name)`` returns the ``(world, frames, ...)`` view of every rank's block.
"""

from __future__ import annotations

import numpy as np

from .pipeline import shard_range

_ALIGN = 256


def frame_shard(nframes: int, rank: int | None = None, world: int | None = None):
    """(start, stop) of this rank's contiguous frame range."""
    import torch.distributed as dist
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    return shard_range(nframes, rank, world)


def _collective_device(group=None):
    """Where tensors handed to a collective must live: the rank's CUDA device under NCCL, the host under gloo."""
    import torch
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


class GatherPlan:
    """One flat buffer per rank holding ``fields`` for ``frames`` frames, and its all-gathered image.

    ``fields``: ``{name: (per_frame_shape, torch dtype)}``.  ``frames`` must be the same on every rank (pad the
    last shard; ``all_gather_frames`` does that for uneven splits)."""

    def __init__(self, fields: dict, frames: int, device=None, group=None):
        import torch
        import torch.distributed as dist
        self.group = group
        self.frames = int(frames)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device(device) if device is not None else _collective_device(group)
        self.layout = {}
        off = 0
        for name, (shape, dtype) in fields.items():
            shape = tuple(int(s) for s in shape)
            nbytes = self.frames * int(np.prod(shape, dtype=np.int64)) * torch.empty((), dtype=dtype).element_size()
            self.layout[name] = (off, nbytes, shape, dtype)
            off = (off + nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
        self.nbytes = max(off, _ALIGN)
        self.flat = torch.zeros(self.nbytes, dtype=torch.uint8, device=self.device)
        self.all = self.flat.view(1, -1) if self.world == 1 else torch.zeros((self.world, self.nbytes), dtype=torch.uint8,
                                                                              device=self.device)

    def local(self, name):
        """This rank's ``(frames, *shape)`` block of ``name`` (contiguous; hand ``.data_ptr()`` to the C ABI)."""
        off, nbytes, shape, dtype = self.layout[name]
        return self.flat[off:off + nbytes].view(dtype).view((self.frames,) + shape)

    def gather(self):
        """The path's only collective: one all-gather of the flat buffers."""
        import torch.distributed as dist
        if self.world > 1:
            dist.all_gather_into_tensor(self.all.view(-1), self.flat, group=self.group)
        return self

    def gathered(self, name):
        """``(world, frames, *shape)`` view of every rank's block (valid after ``gather()``)."""
        off, nbytes, shape, dtype = self.layout[name]
        return self.all[:, off:off + nbytes].contiguous().view(dtype).view((self.world, self.frames) + shape) \
            if self.world > 1 else self.local(name).unsqueeze(0)

    def bytes_per_rank(self) -> int:
        return self.nbytes


def all_gather_frames(local, nframes: int, group=None):
    """All-gather a per-frame tensor whose leading axis is this rank's shard into the full (nframes, ...) tensor.

    Shards may be uneven (nframes not divisible by the world size): every rank pads to the largest
    shard, one ``all_gather_into_tensor`` moves the data, and the padding is dropped on assembly.
    The tensor is staged on the collective's device (the rank's GPU under NCCL) and returned where it came from.
    """
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    home = local.device
    where = _collective_device(group)
    local = local.to(where)
    sizes = [shard_range(nframes, r, world) for r in range(world)]
    biggest = max(b - a for a, b in sizes)
    padded = local
    if local.shape[0] < biggest:
        pad = torch.zeros((biggest - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], dim=0)
    out = torch.empty((world * biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if not all(b - a == biggest for a, b in sizes):
        out = torch.cat([out[r * biggest: r * biggest + (b - a)] for r, (a, b) in enumerate(sizes)], dim=0)
    return out.to(home)


def gather_results(result: dict, nframes: int, keys=("spectra", "ext_limits", "ext_params", "n_ext_bins", "frame_status"),
                   group=None) -> dict:
    """All-gather the numpy outputs of ``pipeline.run_frames`` (host arrays of this rank's shard) across ranks:
    the selected arrays are packed into one ``GatherPlan`` buffer on the collective's device (the GPU under NCCL),
    exchanged with one collective and returned as full ``(nframes, ...)`` numpy arrays."""
    import torch
    import torch.distributed as dist
    present = [k for k in keys if k in result]
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return {k: np.asarray(result[k]) for k in present}
    world = dist.get_world_size(group)
    sizes = [shard_range(nframes, r, world) for r in range(world)]
    biggest = max(b - a for a, b in sizes)
    fields = {}
    for k in present:
        t = torch.from_numpy(np.ascontiguousarray(result[k]))
        fields[k] = (tuple(t.shape[1:]), t.dtype)
    plan = GatherPlan(fields, biggest, group=group)
    for k in present:
        t = torch.from_numpy(np.ascontiguousarray(result[k]))
        plan.local(k)[: t.shape[0]].copy_(t)
    plan.gather()
    out = {}
    for k in present:
        g = plan.gathered(k).cpu()
        out[k] = torch.cat([g[r, : b - a] for r, (a, b) in enumerate(sizes)], dim=0).numpy()
    return out
