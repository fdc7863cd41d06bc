#!/usr/bin/env python3
"""Generate the committed golden vectors by running the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference modules are imported from ``/root/reference/src`` (matplotlib is
stubbed in ``sys.modules`` because ``motes/sky.py:10`` imports the plotting module;
no plot function is reached while every ``diag_plot_*`` flag is False) and driven
stage by stage in the order of ``motes/core.py:80-275``.  ``motes.core`` itself
cannot be imported here (astropy is absent).  Inputs are the seeded, bit-reproducible
frames of ``motes_b200.synth``; each ``.npz`` stores the generator spec, a sha256 of
the inputs and every stage output.  A second run on inputs perturbed by +-1 ulp is
stored as the reference's own self-noise control.
"""

from __future__ import annotations

import copy
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

REFERENCE_SRC = "/root/reference/src"


def import_reference():
    for name in ("matplotlib", "matplotlib.gridspec", "matplotlib.pyplot", "matplotlib.widgets"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.widgets"].Slider = object
    sys.modules["matplotlib"].gridspec = sys.modules["matplotlib.gridspec"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import motes.tracing as tracing
    import motes.sky as sky
    import motes.extraction as extraction
    import motes.common as common
    return tracing, sky, extraction, common


class FitRecorder:
    """Wraps scipy's least_squares inside the reference's tracing module to record nfev/njev/status."""

    def __init__(self, tracing):
        self.inner = tracing.least_squares
        self.log = []
        tracing.least_squares = self

    def __call__(self, *a, **k):
        res = self.inner(*a, **k)
        self.log.append((res.nfev, res.njev, res.status, res.cost))
        return res

    def take(self):
        out, self.log = np.array(self.log, dtype=np.float64).reshape(-1, 4), []
        return out


def run_reference(tracing, sky, extraction, recorder, frame, axes, header, args, seed):
    """core.motes stage order (core.py:80-275) on in-memory dicts; returns every stage output."""
    np.random.seed(seed)
    out = {}
    ndisp = axes["dispersion_axis_length"]
    profile = np.nanmedian(frame["data"], axis=1)
    out["median_profile"] = profile
    params, scale = tracing.median_moffat_fitting(profile, header, axes["spatial_axis"])
    out["median_params"] = np.array(params)
    out["median_fitinfo"] = recorder.take()
    out["scale"] = float(scale)
    out["seeing_px"] = float(header["seeing"])
    lims = tracing.set_extraction_limits(params)
    row_lo, row_hi = int(np.floor(lims[0])), int(np.ceil(lims[1]))
    out["bin_rows"] = np.array([row_lo, row_hi], dtype=np.int64)

    if args.subtract_sky:
        bins, frame = tracing.get_bins(frame, row_lo, row_hi, ndisp, args, has_sky=True)
        out["sky_bins"] = np.array(bins, dtype=np.float64).reshape(-1, 3)
        frame, sky_params, sky_limits = sky.sky_locator(frame, axes, scale, header, bins, args, "synthetic")
        out["sky_fitinfo"] = recorder.take()
        out["sky_params"] = np.array(sky_params)
        out["sky_limits"] = np.array(sky_limits)
        if int(args.sky_order) == 0:
            sm, su = frame["sky_model"], frame["sky_uncertainty"]
            out["sky_model"] = np.array(sm[0]) if sm.size else np.zeros(0)
            out["sky_uncertainty"] = np.array(su[0]) if su.size else np.zeros(0)
        else:
            out["sky_model"] = np.array(frame["sky_model"])
            out["sky_uncertainty"] = np.array(frame["sky_uncertainty"])

    bins, frame = tracing.get_bins(frame, row_lo, row_hi, ndisp, args)
    out["ext_bins"] = np.array(bins, dtype=np.float64).reshape(-1, 3)
    all_params, rows, profiles = [], [], []
    for each_bin in bins:
        p, rows, prof = tracing.moffat_fitting(each_bin, frame["data"], axes, scale, header,
                                               args.sky_fwhm_multiplier, rows)
        all_params.append(p)
        profiles.append(prof)
    out["ext_fitinfo"] = recorder.take()
    all_params = np.array(all_params)
    table = np.array(rows).T
    out["ext_params"] = all_params
    out["ext_limit_table"] = table
    out["ext_profiles"] = np.array(profiles)
    final = tracing.interpolate_extraction_lims(table, ndisp)
    out["ext_limits"] = np.array(final)
    out["data_after"] = np.ascontiguousarray(frame["data"])
    out["errs_after"] = np.ascontiguousarray(frame["errs"])

    frame["data"] = frame["data"].T
    frame["errs"] = frame["errs"].T
    frame["qual"] = frame["qual"].T
    spectra = extraction.optimal_extraction(frame["data"], frame["errs"], final, all_params, axes)
    out["optimal"], out["optimal_err"], out["aperture"], out["aperture_err"] = [np.array(s) for s in spectra]
    return out


def perturb_one_ulp(frame, seed):
    rng = np.random.default_rng(seed)
    for key in ("data", "errs"):
        a = frame[key]
        up = rng.random(a.shape) < 0.5
        gap = a == 1.0                      # keep the literal chip-gap marker
        moved = np.where(up, np.nextafter(a, np.inf), np.nextafter(a, -np.inf))
        frame[key] = np.where(gap, a, moved)
    return frame


CASES = {
    # name: (config, spec overrides, args overrides, rng seed, store frames after sky)
    "t1_sky0": ("T1", {}, {}, 11, True),
    "t1_nosky": ("T1", {}, {"subtract_sky": False}, 11, False),
    "t1_sky2": ("T1", {"sky_slope": 0.2}, {"sky_order": 2}, 12, False),
    "t2_cr_sky0": ("T2", {}, {}, 13, True),
    "t2_unbinned": ("T2", {"amplitude": 4000.0, "cosmic_fraction": 0.0},
                    {"minimum_column_limit": 0, "extraction_snr_limit": 0, "sky_snr_limit": 0}, 14, False),
    "t1_faint": ("T1", {"amplitude": 25.0, "chip_gaps": ()}, {}, 15, False),
    "c1_sky0": ("C1", {}, {}, 1234, False),
    # GMOS dtypes as harvested (gmosio.py:40-52): data / errs big-endian float32, qual float64
    "t1_f32": ("T1", {"input_dtype": ">f4"}, {}, 16, False),
}


def main():
    from motes_b200 import synth
    tracing, sky, extraction, _ = import_reference()
    recorder = FitRecorder(tracing)
    only = set(sys.argv[1:])
    for name, (config, spec_over, args_over, seed, keep_frames) in CASES.items():
        if only and name not in only:
            continue
        spec_over = dict(spec_over)
        input_dtype = spec_over.pop("input_dtype", None)
        spec = synth.config_spec(config, seed=seed, **spec_over)
        args = synth.default_args(**args_over)
        frame, axes, header = synth.make_frame(spec)
        digest = synth.frame_digest(frame)
        if input_dtype:
            frame["data"] = frame["data"].astype(input_dtype)
            frame["errs"] = frame["errs"].astype(input_dtype)
        out = run_reference(tracing, sky, extraction, recorder, copy.deepcopy(frame), axes,
                            dict(header), args, seed)
        # self-noise control: same pipeline, inputs moved by one ulp
        noisy = perturb_one_ulp(copy.deepcopy(frame), seed + 1000)
        ctl = run_reference(tracing, sky, extraction, recorder, noisy, axes, dict(header), args, seed)
        for key in ("median_params", "sky_params", "ext_params", "optimal", "optimal_err",
                    "aperture", "aperture_err", "sky_bins", "ext_bins", "sky_model"):
            if key in ctl:
                out["ctl_" + key] = ctl[key]
        if not keep_frames:
            out.pop("data_after")
            out.pop("errs_after")
        meta = dict(case=name, config=config, spec=synth.spec_to_dict(spec), args=vars(args),
                    rng_seed=seed, input_sha256=digest, input_dtype=input_dtype,
                    versions=dict(numpy=np.__version__, scipy=__import__("scipy").__version__))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, meta=json.dumps(meta), **out)
        print(f"{name}: {os.path.getsize(path)/1024:.0f} KiB  sky_bins={len(out.get('sky_bins', []))} "
              f"ext_bins={len(out['ext_bins'])}")


if __name__ == "__main__":
    main()
