#!/usr/bin/env python3
"""Benchmark of the MOTES extraction hot path on B200: spectral columns extracted per second.

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference algorithm on host cores

Workload (BASELINE.json configs[1] / configs[3]): FORS2-like 4096 x 400 float64 spectrograms, default
MOTES flags (-k -ko 0 -m 15 -xs 10 -ks 100 -kf 3.0).  A step is one pass of the whole path
(median fit -> S/N bins -> per-bin Moffat fits -> sky bootstrap + subtraction -> bins -> fits -> limits
-> optimal + aperture extraction) over `frames_per_gpu` frames on every GPU, followed, for N > 1, by the
all-gather of the extracted spectra and limits.  frames_per_gpu defaults to 128 = 1024 / 8, so N = 8
is exactly the 1024-frame batch of configs[3] (weak scaling: per-GPU work is fixed).

One JSON line is printed by rank 0; see the README / DESIGN.md for the meaning of every key.
"""

from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NSPAT, NDISP = 400, 4096            # configs[1]: FORS2-like 4096x400 spectrogram
METRIC = "spectral columns extracted/sec (fit+sky+optimal)"
UNIT = "columns/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-gpu", type=int, default=128)
    ap.add_argument("--nspat", type=int, default=NSPAT)
    ap.add_argument("--ndisp", type=int, default=NDISP)
    ap.add_argument("--sky-order", type=int, default=0, help="MOTES -ko: 0 = bootstrap median sky (the headline workload), 1..5 = polynomial sky")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-columns", type=int, default=0, help="columns of one frame the CPU baseline runs (0 = auto)")
    ap.add_argument("--profile-steps", type=int, default=1, help="extra untimed steps with per-stage CUDA events")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ synthetic frames
def make_frames_torch(nframes, nspat, ndisp, seed0, device):
    """FORS2-like frames on the device: Moffat trace (beta 3, curving centre, amplitude x2 along dispersion),
    sky with ripple and sparse sky lines, Gaussian noise sqrt(signal + RN^2); errs = sigma, qual = 1."""
    import torch
    data = torch.empty((nframes, nspat, ndisp), dtype=torch.float64, device=device)
    errs = torch.empty_like(data)
    r = torch.arange(nspat, dtype=torch.float64, device=device)[:, None]
    x = torch.arange(ndisp, dtype=torch.float64, device=device)[None, :]
    t = x / ndisp
    tc = t - 0.5
    saw = (torch.arange(ndisp, device=device) % 257).to(torch.float64) / 257.0
    line = (((torch.arange(ndisp, device=device) * 7919) % 97) < 3).to(torch.float64)
    gen = torch.Generator(device=device)
    for i in range(nframes):
        gen.manual_seed(seed0 + i)
        k = (seed0 + i) % 17
        alpha = 3.2 + 0.05 * k
        centre = nspat * 0.5 + 0.37 + 0.11 * k + 9.0 * (4.0 * tc * tc - 0.5)
        b = 1.0 + ((r - centre) ** 2) / (alpha * alpha)
        trace = (900.0 + 20.0 * k) * (1.0 + t) / (b * b * b)
        sky = (120.0 + k) * (1.0 + 0.3 * saw) + 400.0 * line
        signal = trace + sky[None, :]
        sigma = torch.sqrt(signal + 25.0)
        noise = torch.randn((nspat, ndisp), dtype=torch.float64, device=device, generator=gen)
        data[i] = signal + sigma * noise
        errs[i] = sigma
    qual = torch.ones((nframes, nspat, ndisp), dtype=torch.uint8, device=device)
    return data, errs, qual


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self.stop_flag = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            dev = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(dev, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            while not self.stop_flag.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(dev, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(dev)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(dev)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as exc:           # NVML missing: fall back to one nvidia-smi query
            self.reasons.add(f"nvml_unavailable:{type(exc).__name__}")

    def result(self):
        self.stop_flag.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------ CPU arm (oracle port)
def _cpu_worker(job):
    """One process: the reference algorithm (oracle port) on one cropped frame; returns (columns, seconds)."""
    seed, nspat, ndisp, crop = job
    import numpy as np
    from motes_b200 import synth
    from oracle import motes_oracle as mo
    import warnings
    spec = synth.config_spec("C2", seed=seed, nspat=nspat, ndisp=crop)
    frame, axes, header = synth.make_frame(spec)
    args = synth.default_args()
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mo.run_frame(frame, axes, header, args, seed=seed)
    return crop, time.perf_counter() - t0


def cpu_baseline_single(nspat, ndisp, crop):
    """Oracle port, 1 Python thread (the reference as shipped has no parallelism), one cropped frame."""
    cols, secs = _cpu_worker((4242, nspat, ndisp, crop))
    return {"value": cols / secs, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"1 frame {nspat}x{crop} (dispersion crop of the {ndisp}-column workload), whole path, "
                      f"oracle/motes_oracle.py (numpy + scipy least_squares as the reference calls it), {secs:.1f} s"}


def run_reference_arm(a):
    """--impl reference: the reference algorithm on all host cores (one process per core over independent
    frames; the reference itself is single-threaded).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    total_steps = a.steps + a.warmup
    # size the per-step sample so the run takes about two minutes: calibrate on a 256-column crop
    cal_cols, cal_secs = _cpu_worker((1, a.nspat, a.ndisp, 256))
    rate = cal_cols / cal_secs                       # columns/s of one core
    budget = 120.0 / max(1, total_steps)
    crop = int(min(a.ndisp, max(256, (budget * rate * 0.7) // 256 * 256)))
    if a.cpu_sample_columns:
        crop = int(min(crop, max(64, a.cpu_sample_columns)))
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        def step(k):
            jobs = [(1000 * k + p, a.nspat, a.ndisp, crop) for p in range(procs)]
            t0 = time.perf_counter()
            done = pool.map(_cpu_worker, jobs)
            return sum(c for c, _ in done), time.perf_counter() - t0
        for k in range(a.warmup):
            step(k)
        cols = secs = 0.0
        for k in range(a.steps):
            c, s = step(a.warmup + k)
            cols += c
            secs += s
    value = cols / secs
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * secs / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": {"workload": f"C2 FORS2-like {a.ndisp}x{a.nspat} float64 frames, default MOTES flags, sky order 0",
                   "frames_per_step": procs, "columns_per_frame_sample": crop},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": f"{procs} processes x 1 frame {a.nspat}x{crop} per step (dispersion crop), "
                                   f"oracle/motes_oracle.py"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------ GPU arm
def run_b200_arm(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from motes_b200 import _lib, pipeline, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    F, ns, nd = a.frames_per_gpu, a.nspat, a.ndisp
    args = synth.default_args(sky_order=a.sky_order)
    h = _lib.Handle(local)
    stream = torch.cuda.current_stream(device)
    _lib.check(_lib.lib.motes_set_stream(h.raw, ctypes.c_void_p(stream.cuda_stream)))
    p = pipeline.make_params(args)
    cap = pipeline.bin_capacity(nd, p.minimum_column_limit)

    # pristine inputs (device) + pinned host copies for the end-to-end leg
    data0, errs0, qual = make_frames_torch(F, ns, nd, 100000 + rank * F, device)
    data = torch.empty_like(data0)
    errs = torch.empty_like(errs0)
    seeing = torch.full((F,), 0.8, dtype=torch.float64, device=device)
    pixres = torch.full((F,), 0.16, dtype=torch.float64, device=device)
    seeds = (np.arange(F) + 7 + rank * F).astype(np.uint32)
    states0 = torch.from_numpy(pipeline.seed_states(seeds).view(np.int32)).to(device)
    states = torch.empty_like(states0)

    out = {
        "spectra": torch.zeros((F, 4, nd), dtype=torch.float64, device=device),
        "ext_limits": torch.zeros((F, 2, nd), dtype=torch.float64, device=device),
        "ext_params": torch.zeros((F, cap, 8), dtype=torch.float64, device=device),
        "n_ext_bins": torch.zeros((F,), dtype=torch.int32, device=device),
        "n_sky_bins": torch.zeros((F,), dtype=torch.int32, device=device),
        "sky_model": torch.zeros((F, nd), dtype=torch.float64, device=device),
        "frame_status": torch.zeros((F,), dtype=torch.int32, device=device),
    }
    if a.sky_order > 0:
        out.pop("sky_model")            # (F, nspat, ndisp) for a polynomial sky: not part of the timed outputs
    ostruct = _lib.MotesOutputs()
    for k, v in out.items():
        setattr(ostruct, k, v.data_ptr())
    gathered = None
    if world > 1:
        gathered = {k: torch.empty((world * F,) + tuple(out[k].shape[1:]), dtype=out[k].dtype, device=device)
                    for k in ("spectra", "ext_limits", "ext_params")}

    def device_step():
        data.copy_(data0)
        errs.copy_(errs0)
        states.copy_(states0)
        _lib.check(_lib.lib.motes_run_frames_dev(h.raw, data.data_ptr(), errs.data_ptr(), qual.data_ptr(), F, ns, nd, cap,
                                                 ctypes.byref(p), seeing.data_ptr(), pixres.data_ptr(),
                                                 states.data_ptr(), ctypes.byref(ostruct)))
        if world > 1:
            for k in gathered:
                dist.all_gather_into_tensor(gathered[k], out[k])

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for _ in range(a.warmup):
        device_step()
    fence()
    status = out["frame_status"].cpu().numpy()
    if np.any(status != 0):
        raise SystemExit(f"rank {rank}: frames failed on the device: {np.unique(status)}")
    launches0 = h.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    ev0.record(stream)
    for _ in range(a.steps):
        device_step()
    ev1.record(stream)
    fence()
    clocks = sampler.result()
    ms = ev0.elapsed_time(ev1)
    launches = h.launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    columns_per_step = world * F * nd
    value = columns_per_step * a.steps / (ms * 1e-3)

    # ---- per-stage CUDA-event timing (separate, untimed steps) for the roofline of the dominant kernel
    stage_ms = {}
    if a.profile_steps > 0:
        _lib.check(_lib.lib.motes_profile_enable(h.raw, 1))
        for _ in range(a.profile_steps):
            device_step()
        n_st = _lib.lib.motes_profile_stage_count()
        buf = (ctypes.c_double * n_st)()
        cnt = (ctypes.c_longlong * n_st)()
        _lib.check(_lib.lib.motes_profile_read(h.raw, buf, cnt, 1))
        _lib.check(_lib.lib.motes_profile_enable(h.raw, 0))
        for i in range(n_st):
            name = _lib.lib.motes_profile_stage_name(i).decode()
            stage_ms[name] = {"ms_per_step": buf[i] / a.profile_steps, "launches_per_step": cnt[i] / a.profile_steps}
    fence()

    # ---- BASELINE.json configs[1] as stated: ONE 4096 x 400 frame on one B200 (latency of a single-frame call)
    def single_step():
        data[:1].copy_(data0[:1])
        errs[:1].copy_(errs0[:1])
        states[:1].copy_(states0[:1])
        _lib.check(_lib.lib.motes_run_frames_dev(h.raw, data.data_ptr(), errs.data_ptr(), qual.data_ptr(), 1, ns, nd, cap,
                                                 ctypes.byref(p), seeing.data_ptr(), pixres.data_ptr(),
                                                 states.data_ptr(), ctypes.byref(ostruct)))
    for _ in range(3):
        single_step()
    torch.cuda.synchronize(device)
    # a one-frame call is ~90 short launches: its time is sensitive to whatever else the host is doing, so it is
    # taken as the best of four groups of five calls (the median group is reported beside it)
    groups = []
    for _ in range(4):
        sv0, sv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sv0.record(stream)
        for _ in range(5):
            single_step()
        sv1.record(stream)
        torch.cuda.synchronize(device)
        groups.append(sv0.elapsed_time(sv1) / 5.0)
    single_ms = min(groups)
    single_ms_median = sorted(groups)[len(groups) // 2]
    fence()

    # ---- end-to-end leg: pinned host buffers -> H2D -> path -> D2H of the spectra, through the host C ABI
    hd = torch.empty((F, ns, nd), dtype=torch.float64).pin_memory()
    he = torch.empty((F, ns, nd), dtype=torch.float64).pin_memory()
    hq = torch.ones((F, ns, nd), dtype=torch.uint8).pin_memory()
    hd.copy_(data0)
    he.copy_(errs0)
    hsee = np.full(F, 0.8)
    hpix = np.full(F, 0.16)
    hstates0 = pipeline.seed_states(seeds)
    hout = {"spectra": torch.zeros((F, 4, nd), dtype=torch.float64).pin_memory(),
            "frame_status": torch.zeros((F,), dtype=torch.int32).pin_memory()}
    hstruct = _lib.MotesOutputs()
    for k, v in hout.items():
        setattr(hstruct, k, v.data_ptr())
    h2d = F * ns * nd * (8 + 8 + 1) + F * (625 * 4 + 16)
    d2h = hout["spectra"].numel() * 8 + F * 4 + F * 625 * 4

    def e2e_step():
        hst = hstates0.copy()
        _lib.check(_lib.lib.motes_run_frames_host(h.raw, hd.data_ptr(), he.data_ptr(), hq.data_ptr(), F, ns, nd, cap,
                                                  ctypes.byref(p), hsee.ctypes.data, hpix.ctypes.data, hst.ctypes.data, 0,
                                                  ctypes.byref(hstruct)))
    e2e_steps = max(1, min(a.steps, 3))
    e2e_step()
    fence()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    fence()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_sync_value = columns_per_step * e2e_steps / float(te.item())

    # the same with two batches in flight: motes_run_frames_host_async on two handles, so that the upload of
    # step k+1 runs under the kernels of step k (what a caller with a stream of batches does); every step still
    # uploads its own inputs and downloads its own spectra inside the timed region
    del data, errs, states
    torch.cuda.empty_cache()
    h2 = _lib.Handle(local)
    handles = [h, h2]
    houts = [hout, {"spectra": torch.zeros((F, 4, nd), dtype=torch.float64).pin_memory(),
                    "frame_status": torch.zeros((F,), dtype=torch.int32).pin_memory()}]
    hstructs = [hstruct, _lib.MotesOutputs()]
    for k, v in houts[1].items():
        setattr(hstructs[1], k, v.data_ptr())
    hsts = [torch.from_numpy(hstates0.view(np.int32).copy()).pin_memory() for _ in range(2)]
    hstates0_t = torch.from_numpy(hstates0.view(np.int32))

    def e2e_pipelined(nsteps):
        for k in range(nsteps):
            i = k % 2
            if k >= 2:
                _lib.check(_lib.lib.motes_synchronize(handles[i].raw))
            hsts[i].copy_(hstates0_t)
            _lib.check(_lib.lib.motes_run_frames_host_async(handles[i].raw, hd.data_ptr(), he.data_ptr(), hq.data_ptr(), F, ns,
                                                            nd, cap, ctypes.byref(p), hsee.ctypes.data, hpix.ctypes.data,
                                                            hsts[i].data_ptr(), 0, ctypes.byref(hstructs[i])))
        for hh in handles:
            _lib.check(_lib.lib.motes_synchronize(hh.raw))

    e2e_pipelined(2)
    fence()
    pipe_steps = max(2, min(2 * a.steps, 6))
    t0 = time.perf_counter()
    e2e_pipelined(pipe_steps)
    fence()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = columns_per_step * pipe_steps / float(te.item())
    if not np.array_equal(houts[0]["spectra"].numpy(), houts[1]["spectra"].numpy()):
        raise SystemExit("end-to-end leg: the two in-flight batches (same inputs) returned different spectra")

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        alg_bytes_per_col = 17 * ns + 80
        dominant = max(stage_ms, key=lambda k: stage_ms[k]["ms_per_step"]) if stage_ms else None
        roofline = None
        if dominant:
            dur = stage_ms[dominant]["ms_per_step"] * 1e-3
            launches_dom = max(1.0, stage_ms[dominant]["launches_per_step"])
            achieved = alg_bytes_per_col * F * nd / dur / 1e9
            traffic = None                  # DRAM bytes per launch of that kernel from the committed ncu capture
            ncu_util = None                 # issue-slot / FP64-pipe utilisation of that kernel, same capture
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
                w = tj["workload"]
                if (w["frames_per_gpu"], w["nspat"], w["ndisp"]) == (F, ns, nd):
                    traffic = tj["dram_bytes_per_launch"].get(dominant)
                    ncu_util = tj.get("ncu_utilisation", {}).get(dominant)
            except Exception:
                pass
            # the kernel closest to the memory roofline: measured DRAM bytes per launch (same ncu capture) over its
            # CUDA-event time in this run
            nearest = None
            try:
                if (w["frames_per_gpu"], w["nspat"], w["ndisp"]) == (F, ns, nd):
                    for name, nbytes in tj["dram_bytes_per_launch"].items():
                        if name in stage_ms and stage_ms[name]["ms_per_step"] > 0 and stage_ms[name]["launches_per_step"] > 0:
                            per_launch_s = stage_ms[name]["ms_per_step"] * 1e-3 / stage_ms[name]["launches_per_step"]
                            gbs = nbytes / per_launch_s / 1e9
                            if nearest is None or gbs > nearest["achieved"]:
                                nearest = {"kernel": name, "achieved": gbs, "unit": "GB/s", "frac": gbs / hbm_peak,
                                           "dram_bytes_per_launch": nbytes, "kernel_ms_per_launch": per_launch_s * 1e3}
            except Exception:
                nearest = None
            roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                        "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                        "algorithmic_bytes_per_column": alg_bytes_per_col,
                        "kernel_ms_per_launch": 1e3 * dur / launches_dom, "ncu_utilisation": ncu_util,
                        "nearest_to_hbm_roofline": nearest,
                        "note": "whole-path algorithmic bytes (17*nspat+80 per column) over the dominant stage's "
                                "CUDA-event time (stage times are taken in separate profiling steps, each kernel alone: the "
                                "timed steps overlap the generator with the localisation stages); the path is "
                                "FP64/integer-latency bound, not HBM bound (DESIGN.md)"}
        total_stage = sum(v["ms_per_step"] for v in stage_ms.values()) or 1.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C2 FORS2-like {nd}x{ns} float64 frames (C4 batch slice), default MOTES flags, sky order {a.sky_order}, "
                                   f"bit-exact numpy-MT19937 sky bootstrap",
                       "frames_per_gpu": F, "frames_per_step": world * F, "columns_per_step": columns_per_step,
                       "l2": f"inputs {2 * F * ns * nd * 8 / 1e9:.2f} GB per GPU per step, larger than the 126 MB L2",
                       "parallelism": f"frames sharded over {world} GPU(s), all-gather of spectra/limits/params" if world > 1
                                      else "1 GPU"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": pipe_steps,
                    "api": "motes_run_frames_host_async on two handles (pinned host buffers; every step uploads its inputs, "
                           "runs the path and downloads its spectra; two steps in flight so the next upload overlaps the "
                           "kernels)",
                    "one_call_at_a_time": {"value": e2e_sync_value, "steps": e2e_steps,
                                           "api": "motes_run_frames_host (H2D + path + D2H, then return)"}},
            "gpu_launches": int(launches),
            "single_frame": {"ms": single_ms, "ms_median_group": single_ms_median, "columns_per_s": nd / (single_ms * 1e-3),
                             "note": "BASELINE.json configs[1] as stated: one 4096x400 frame per call on one GPU, device-resident"},
            "roofline": roofline,
            "stage_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in stage_ms.items()},
            "stage_share": {k: round(v["ms_per_step"] / total_stage, 4) for k, v in stage_ms.items()},
        }
        if not a.no_cpu_baseline and world == 1:
            crop = a.cpu_sample_columns or 2048
            line["cpu_baseline"] = cpu_baseline_single(ns, nd, min(crop, nd))
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


RESULT_FD = None


def emit(line: dict) -> None:
    """The one JSON line of the run, on the process's original stdout."""
    text = (json.dumps(line) + "\n").encode()
    if RESULT_FD is None:
        sys.stdout.write(text.decode())
        sys.stdout.flush()
    else:
        os.write(RESULT_FD, text)


def main():
    global RESULT_FD
    a = parse()
    # Native libraries write to file descriptor 1 as well (NCCL prints its version banner there): point fd 1 at
    # stderr for the run and keep the original for the result line, so stdout carries exactly one JSON line.
    sys.stdout.flush()
    RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_b200_arm(a)


if __name__ == "__main__":
    main()
