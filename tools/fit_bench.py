#!/usr/bin/env python3
"""Micro-benchmark of the batched Moffat fit stage alone (motes_moffat_fit_batch_dev) on synthetic bin profiles.

    python tools/fit_bench.py [--nfit 32896] [--nspat 400] [--variants group,rounds64,rounds128,rounds32]

Each variant is a handle created under the matching MOTES_B200_FIT* switches; the first one is the reference the others
are compared with (parameters bit for bit, evaluation counts, stop reasons).  Times are CUDA events over `--reps` calls.
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {
    "group": {"MOTES_B200_FIT": "group"},
    "rounds128": {"MOTES_B200_FIT": "auto", "MOTES_B200_FIT_THREADS": "128", "MOTES_B200_FIT_MIN": "0"},
    "rounds64": {"MOTES_B200_FIT": "auto", "MOTES_B200_FIT_THREADS": "64", "MOTES_B200_FIT_MIN": "0"},
    "rounds32": {"MOTES_B200_FIT": "auto", "MOTES_B200_FIT_THREADS": "32", "MOTES_B200_FIT_MIN": "0"},
}


def make_profiles(nfit, nspat, seed=3):
    """Sky-subtracted bin-median profiles like the extraction bins of the bench frames (16-column medians)."""
    rng = np.random.default_rng(seed)
    r = np.arange(nspat, dtype=np.float64)[None, :]
    k = rng.integers(0, 17, size=(nfit, 1))
    alpha = 3.2 + 0.05 * k
    centre = nspat * 0.5 + 0.37 + 0.11 * k + 9.0 * (4.0 * (rng.random((nfit, 1)) - 0.5) ** 2 - 0.5)
    amp = (900.0 + 20.0 * k) * (1.0 + rng.random((nfit, 1)))
    b = 1.0 + ((r - centre) ** 2) / (alpha * alpha)
    trace = amp / (b * b * b)
    sigma = np.sqrt(trace + 120.0 + 25.0) * (1.2533 / 4.0)          # median of 16 columns
    return trace + sigma * rng.standard_normal((nfit, nspat))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nfit", type=int, default=32896)
    ap.add_argument("--nspat", type=int, default=400)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--variants", default="group,rounds128,rounds64")
    ap.add_argument("--extra-env", default="", help="K=V,K=V applied to every variant")
    a = ap.parse_args()
    import torch
    from motes_b200 import _lib
    dev = torch.device("cuda", 0)
    prof = torch.from_numpy(make_profiles(a.nfit, a.nspat)).to(dev)
    alpha0 = torch.full((a.nfit,), 5.0, dtype=torch.float64, device=dev)
    results = {}
    first = None
    for name in a.variants.split(","):
        env = dict(VARIANTS.get(name.split("@")[0], {}))
        for kv in (name.split("@")[1:] + [x for x in a.extra_env.split(",") if x]):
            key, val = kv.split("=")
            env[key] = val
        old = {key: os.environ.get(key) for key in env}
        os.environ.update(env)
        h = _lib.Handle(0)
        for key, val in old.items():
            if val is None:
                os.environ.pop(key, None)
            else:
                os.environ[key] = val
        stream = torch.cuda.current_stream(dev)
        _lib.check(_lib.lib.motes_set_stream(h.raw, ctypes.c_void_p(stream.cuda_stream)))
        params = torch.zeros((a.nfit, 6), dtype=torch.float64, device=dev)
        cost = torch.zeros((a.nfit,), dtype=torch.float64, device=dev)
        info = torch.zeros((3, a.nfit), dtype=torch.int32, device=dev)

        def run():
            _lib.check(_lib.lib.motes_moffat_fit_batch_dev(h.raw, prof.data_ptr(), a.nfit, a.nspat, alpha0.data_ptr(), params.data_ptr(),
                                                           cost.data_ptr(), info[0].data_ptr(), info[1].data_ptr(), info[2].data_ptr()))
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.reps):
            run()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        out = {"params": params.cpu().numpy(), "nfev": info[0].cpu().numpy(), "njev": info[1].cpu().numpy(),
               "status": info[2].cpu().numpy(), "cost": cost.cpu().numpy()}
        rep = {"ms": round(ms, 3), "fits_per_s": round(a.nfit / ms * 1e3), "nfev_mean": float(out["nfev"].mean()),
               "nfev_max": int(out["nfev"].max()), "status_hist": np.bincount(out["status"].clip(0, 5), minlength=6).tolist()}
        if first is None:
            first = out
        else:
            rel = np.abs(out["params"] - first["params"]) / np.maximum(np.abs(first["params"]), 1e-300)
            rep.update({"params_bit_equal": bool(np.array_equal(out["params"], first["params"])),
                        "nfev_equal_frac": float((out["nfev"] == first["nfev"]).mean()),
                        "status_equal_frac": float((out["status"] == first["status"]).mean()),
                        "params_rel_p99": [float(v) for v in np.percentile(rel[:, :4], 99, axis=0)],
                        "params_rel_max": [float(v) for v in rel[:, :4].max(axis=0)]})
        results[name] = rep
        print(name, json.dumps(rep), flush=True)
        h.close()


if __name__ == "__main__":
    main()
