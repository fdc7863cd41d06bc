"""Loader and stage driver for the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

``load()`` imports the reference's hot-path modules either from the compiled staging area ``oracle/_ref``
(marshalled module code objects made by ``oracle/build_ref.py``; it travels to the GPU box) or, in the build
container, straight from ``/root/reference/src``.  ``matplotlib`` is absent from the image and ``motes/sky.py:10`` imports the plotting
module, so empty stand-in modules are registered first; no plot function is reached while every ``diag_plot_*``
flag is False.  ``motes.core`` needs astropy (absent), so ``run_frame`` drives the stage functions in the order of
``core.motes`` (``core.py:80-275``) on in-memory dictionaries - the same driver the golden files were made with.

Used by ``tests/`` (checker), ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs (the timed CPU baseline,
``cpu_baseline.kind = "reference"``).  The product never imports this file.
"""

from __future__ import annotations

import marshal
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")
REFERENCE_SRC = "/root/reference/src"

_loaded = None


STAGED_MARK = os.path.join(STAGED, "motes", "tracing.code")
# import order of the staged modules (each only imports the ones before it, numpy and scipy)
STAGED_ORDER = ("common", "logs", "diagnostics", "tracing", "extraction", "sky")


def available() -> bool:
    return os.path.exists(STAGED_MARK) or os.path.isdir(os.path.join(REFERENCE_SRC, "motes"))


def _import_staged():
    """Execute the staged code objects into a ``motes`` package (sys.modules), dependencies first."""
    pkg_dir = os.path.join(STAGED, "motes")
    pkg = types.ModuleType("motes")
    pkg.__path__ = [pkg_dir]
    pkg.__file__ = os.path.join(pkg_dir, "__init__.code")
    pkg.__package__ = "motes"
    sys.modules["motes"] = pkg
    with open(pkg.__file__, "rb") as fh:
        exec(marshal.load(fh), pkg.__dict__)
    for name in STAGED_ORDER:
        mod = types.ModuleType(f"motes.{name}")
        mod.__file__ = os.path.join(pkg_dir, name + ".code")
        mod.__package__ = "motes"
        sys.modules[mod.__name__] = mod
        setattr(pkg, name, mod)
        with open(mod.__file__, "rb") as fh:
            exec(marshal.load(fh), mod.__dict__)


def _stub_matplotlib():
    for name in ("matplotlib", "matplotlib.gridspec", "matplotlib.pyplot", "matplotlib.widgets"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if not hasattr(sys.modules["matplotlib.widgets"], "Slider"):
        sys.modules["matplotlib.widgets"].Slider = object
    if not hasattr(sys.modules["matplotlib"], "gridspec"):
        sys.modules["matplotlib"].gridspec = sys.modules["matplotlib.gridspec"]
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def load(prefer_staged: bool = True):
    """Returns a namespace with ``tracing, sky, extraction, common, origin``; raises ImportError when neither
    the staged copy nor /root/reference exists."""
    global _loaded
    if _loaded is not None:
        return _loaded
    staged_ok = os.path.exists(STAGED_MARK)
    source_ok = os.path.isdir(os.path.join(REFERENCE_SRC, "motes"))
    if not (staged_ok or source_ok):
        raise ImportError("the reference is neither staged under oracle/_ref (python oracle/build_ref.py in the build "
                          "container) nor present at /root/reference")
    root = STAGED if (staged_ok and (prefer_staged or not source_ok)) else REFERENCE_SRC
    _stub_matplotlib()
    already = getattr(sys.modules.get("motes"), "__file__", None) or ""
    if already:
        # the process has the reference imported already (a test that drives /root/reference directly): use that copy
        root = STAGED if already.startswith(STAGED) else os.path.dirname(os.path.dirname(already))
    if root == STAGED and not already:
        _import_staged()
    sys.path.insert(0, root)
    try:
        import motes.common as common
        import motes.extraction as extraction
        import motes.sky as sky
        import motes.tracing as tracing
    finally:
        sys.path.remove(root)
    _loaded = types.SimpleNamespace(tracing=tracing, sky=sky, extraction=extraction, common=common,
                                    origin="oracle/_ref (compiled code objects of the reference modules)" if root == STAGED else REFERENCE_SRC)
    return _loaded


class FitRecorder:
    """Wraps scipy's least_squares inside the reference's tracing module to record nfev / njev / status / cost."""

    def __init__(self, tracing):
        self.tracing = tracing
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

    def close(self):
        self.tracing.least_squares = self.inner


def run_stages(tracing, sky, extraction, frame, axes, header, args, seed=None, recorder=None, keep_frames=True):
    """core.motes stage order (core.py:80-275) over any module triple with the reference's function API - the
    reference itself or the CUDA-backed mirrors ``motes_b200.{tracing,sky,extraction}``."""
    if seed is not None:
        np.random.seed(seed)
    out = {}
    ndisp = axes["dispersion_axis_length"]
    profile = np.nanmedian(frame["data"], axis=1)                                          # core.py:80
    out["median_profile"] = profile
    params, scale = tracing.median_moffat_fitting(profile, header, axes["spatial_axis"])   # core.py:83-88
    out["median_params"] = np.array(params)
    if recorder:
        out["median_fitinfo"] = recorder.take()
    out["scale"] = float(scale)
    out["seeing_px"] = float(header["seeing"])
    lims = tracing.set_extraction_limits(params)                                           # core.py:93
    row_lo, row_hi = int(np.floor(lims[0])), int(np.ceil(lims[1]))
    out["bin_rows"] = np.array([row_lo, row_hi], dtype=np.int64)

    if args.subtract_sky:                                                                  # core.py:128-168
        bins, frame = tracing.get_bins(frame, row_lo, row_hi, ndisp, args, has_sky=True)
        out["sky_bins"] = np.array(bins, dtype=np.float64).reshape(-1, 3)
        frame, sky_params, sky_limits = sky.sky_locator(frame, axes, scale, header, bins, args, "synthetic")
        if recorder:
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

    bins, frame = tracing.get_bins(frame, row_lo, row_hi, ndisp, args)                     # core.py:183-190
    out["ext_bins"] = np.array(bins, dtype=np.float64).reshape(-1, 3)
    all_params, rows, profiles = [], [], []
    for each_bin in bins:                                                                  # core.py:200-212
        p, rows, prof = tracing.moffat_fitting(each_bin, frame["data"], axes, scale, header,
                                               args.sky_fwhm_multiplier, rows)
        all_params.append(p)
        profiles.append(prof)
    if recorder:
        out["ext_fitinfo"] = recorder.take()
    all_params = np.array(all_params)
    table = np.array(rows).T
    out["ext_params"] = all_params
    out["ext_limit_table"] = table
    out["ext_profiles"] = np.array(profiles)
    final = tracing.interpolate_extraction_lims(table, ndisp)                              # core.py:232
    out["ext_limits"] = np.array(final)
    if keep_frames:
        out["data_after"] = np.ascontiguousarray(frame["data"])
        out["errs_after"] = np.ascontiguousarray(frame["errs"])

    frame["data"] = frame["data"].T                                                        # core.py:265-267
    frame["errs"] = frame["errs"].T
    frame["qual"] = frame["qual"].T
    spectra = extraction.optimal_extraction(frame["data"], frame["errs"], final, all_params, axes)
    out["optimal"], out["optimal_err"], out["aperture"], out["aperture_err"] = [np.array(s) for s in spectra]
    return out


def run_frame(frame, axes, header, args, seed=None, keep_frames=False):
    """The reference on one frame: same call shape as ``oracle.motes_oracle.run_frame``."""
    ref = load()
    return run_stages(ref.tracing, ref.sky, ref.extraction, frame, axes, header, args, seed=seed, keep_frames=keep_frames)
