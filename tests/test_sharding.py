"""Host-side multi-process logic (gloo, world size 2, CPU): frame sharding and the output all-gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from motes_b200 import pipeline


def test_shard_ranges_partition_the_batch():
    for n in (1, 7, 128, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [pipeline.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nframes, ndisp, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from motes_b200 import distributed
    a, b = distributed.frame_shard(nframes)
    frames = np.arange(a, b, dtype=np.float64)
    # stand-in for this rank's run_frames() outputs: values that encode the global frame index
    local = {"spectra": frames[:, None, None] * 1000.0 + np.arange(4)[None, :, None] * 100.0 + np.arange(ndisp)[None, None, :],
             "n_ext_bins": (frames.astype(np.int32) % 7),
             "frame_status": np.zeros(b - a, dtype=np.int32)}
    full = distributed.gather_results(local, nframes)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nframes", [6, 7])
def test_all_gather_assembles_frames_in_order_on_every_rank(tmp_path, nframes):
    world, ndisp = 2, 5
    port = _free_port()
    mp.spawn(_worker, args=(world, port, nframes, ndisp, str(tmp_path)), nprocs=world, join=True)
    frames = np.arange(nframes, dtype=np.float64)
    want = frames[:, None, None] * 1000.0 + np.arange(4)[None, :, None] * 100.0 + np.arange(ndisp)[None, None, :]
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(got["spectra"], want)
        assert np.array_equal(got["n_ext_bins"], frames.astype(np.int32) % 7)
        assert got["frame_status"].shape == (nframes,)
