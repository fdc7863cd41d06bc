"""The path's one collective under NCCL (needs two GPUs; skipped on a single-GPU box): `distributed.gather_results`
stages the host outputs of each rank's shard on its GPU, all-gathers once, and every rank ends with the frames of a
single-GPU run in order; `GatherPlan` does the same for device tensors written in place."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from motes_b200 import _lib, distributed, pipeline, synth
    nframes = 5                                           # uneven split: 3 + 2
    frames = [synth.make_frame(synth.config_spec("T2", seed=60 + i)) for i in range(nframes)]
    axes, header = frames[0][1], frames[0][2]
    args = synth.default_args()
    a, b = distributed.frame_shard(nframes)
    data = np.stack([f[0]["data"] for f in frames[a:b]])
    errs = np.stack([f[0]["errs"] for f in frames[a:b]])
    qual = np.stack([f[0]["qual"] for f in frames[a:b]])
    res = pipeline.run_frames(data, errs, qual, args, header["seeing"], header["pixel_resolution"],
                              seeds=[100 + i for i in range(a, b)], axes_dict=axes, handle=_lib.Handle(rank))
    full = distributed.gather_results(res, nframes)
    # GatherPlan on device tensors
    plan = distributed.GatherPlan({"x": ((3,), torch.float64), "k": ((), torch.int32)}, 2, device=torch.device("cuda", rank))
    plan.local("x").copy_(torch.full((2, 3), float(rank + 1), dtype=torch.float64))
    plan.local("k").copy_(torch.tensor([10 * rank, 10 * rank + 1], dtype=torch.int32))
    plan.gather()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), x=plan.gathered("x").cpu().numpy(), k=plan.gathered("k").cpu().numpy(), **full)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_results_over_nccl_equals_single_gpu_run(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from motes_b200 import pipeline, synth
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    frames = [synth.make_frame(synth.config_spec("T2", seed=60 + i)) for i in range(5)]
    axes, header = frames[0][1], frames[0][2]
    want = pipeline.run_frames(np.stack([f[0]["data"] for f in frames]), np.stack([f[0]["errs"] for f in frames]),
                               np.stack([f[0]["qual"] for f in frames]), synth.default_args(), header["seeing"],
                               header["pixel_resolution"], seeds=[100 + i for i in range(5)], axes_dict=axes)
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npz")
        for key in ("spectra", "ext_limits", "n_ext_bins", "frame_status"):
            assert np.array_equal(got[key], want[key], equal_nan=True), (r, key)
        n = want["n_ext_bins"]
        for i in range(5):
            assert np.array_equal(got["ext_params"][i, :n[i]], want["ext_params"][i, :n[i]])
        assert np.array_equal(got["x"], np.stack([np.full((2, 3), 1.0), np.full((2, 3), 2.0)]))
        assert np.array_equal(got["k"], np.array([[0, 1], [10, 11]], dtype=np.int32))
