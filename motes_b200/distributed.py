"""Multi-GPU plumbing for the extraction path: frames are independent end to end, so the
only exchange is one all-gather of the per-frame outputs (spectra, limits, bin parameters; ~64 B/column).

One process per GPU (``torchrun``); ``torch.distributed`` with NCCL on GPUs, gloo in CPU tests.
"""

from __future__ import annotations

import numpy as np

from .pipeline import shard_range


def frame_shard(nframes: int, rank: int | None = None, world: int | None = None):
    """(start, stop) of this rank's contiguous frame range."""
    import torch.distributed as dist
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    return shard_range(nframes, rank, world)


def all_gather_frames(local, nframes: int, group=None):
    """All-gather a per-frame tensor whose leading axis is this rank's shard into the full (nframes, ...) tensor.

    Shards may be uneven (nframes not divisible by the world size): every rank pads to the largest
    shard, one ``all_gather_into_tensor`` moves the data, and the padding is dropped on assembly.
    """
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(nframes, r, world) for r in range(world)]
    biggest = max(b - a for a, b in sizes)
    padded = local
    if local.shape[0] < biggest:
        pad = torch.zeros((biggest - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], dim=0)
    out = torch.empty((world * biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if all(b - a == biggest for a, b in sizes):
        return out
    return torch.cat([out[r * biggest: r * biggest + (b - a)] for r, (a, b) in enumerate(sizes)], dim=0)


def gather_results(result: dict, nframes: int, keys=("spectra", "ext_limits", "ext_params", "n_ext_bins", "frame_status"),
                   group=None) -> dict:
    """All-gather the numpy outputs of ``pipeline.run_frames`` (host arrays) across ranks."""
    import torch
    gathered = {}
    for key in keys:
        if key in result:
            t = torch.from_numpy(np.ascontiguousarray(result[key]))
            gathered[key] = all_gather_frames(t, nframes, group).numpy()
    return gathered
