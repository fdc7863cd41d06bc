"""Parity metrics between a device run and the reference / oracle on the same frame.  TEST INFRASTRUCTURE ONLY.

One function, ``compare``, produces the numbers BASELINE.md §3 asks for beside every timed configuration:
integer outputs bit-exact (bin edges, integer extraction limits with the flip count), the bootstrap sky model
bit-exact, Moffat parameters as median / p99 / max per parameter with the count of fits beyond 1e-6, spectra as the
largest relative difference - optionally next to the reference's own +-1-ulp self-noise control.  Used by
``tests/``, ``bench.py``'s ``parity`` key and ``tools/parity_table.py``.
"""

from __future__ import annotations

import numpy as np

PARAM_NAMES = ("A", "c", "alpha", "beta", "B", "m")
SPECTRA = ("optimal", "optimal_err", "aperture", "aperture_err")


def param_errors(got, ref, nspat):
    """Per-parameter error of fitted Moffat parameters (rows = fits, first 6 columns).  A, c, alpha, beta: relative.
    B and m sit near zero after sky subtraction, where a relative error is ill-posed: they are measured against
    max(|x|, A) and max(|x|, A / nspat)."""
    got, ref = np.atleast_2d(np.asarray(got, dtype=float))[:, :6], np.atleast_2d(np.asarray(ref, dtype=float))[:, :6]
    amp = np.abs(ref[:, 0])
    scale = np.abs(ref).copy()
    scale[:, 4] = np.maximum(scale[:, 4], amp)
    scale[:, 5] = np.maximum(scale[:, 5], amp / nspat)
    return np.abs(got - ref) / np.maximum(scale, 1e-300)


def _param_block(got, ref, nspat):
    err = param_errors(got, ref, nspat)
    return {"fits": int(err.shape[0]),
            "median": [float(v) for v in np.median(err, axis=0)],
            "p99": [float(v) for v in np.percentile(err, 99, axis=0)],
            "max": [float(v) for v in err.max(axis=0)],
            "fits_beyond_1e-6": int((err.max(axis=1) > 1e-6).sum())}


def integer_limits(limits):
    return np.floor(np.asarray(limits)[0]), np.ceil(np.asarray(limits)[1])


def compare(got: dict, want: dict, nspat: int, control: dict | None = None) -> dict:
    """``got`` / ``want``: stage dictionaries as returned by ``pipeline.run_frame`` / ``run_stages``.
    ``control``: the reference's outputs on inputs moved by +-1 ulp (golden ``ctl_*`` keys without the prefix)."""
    rep = {}
    rep["ext_bins_exact"] = bool(np.array_equal(np.asarray(got["ext_bins"])[:, :2], np.asarray(want["ext_bins"])[:, :2]))
    rep["n_ext_bins"] = int(len(want["ext_bins"]))
    if "sky_bins" in want and len(want["sky_bins"]):
        rep["sky_bins_exact"] = bool(np.array_equal(np.asarray(got["sky_bins"])[:, :2], np.asarray(want["sky_bins"])[:, :2]))
        rep["n_sky_bins"] = int(len(want["sky_bins"]))
    rep["bins_exact"] = rep["ext_bins_exact"] and rep.get("sky_bins_exact", True)
    glo, ghi = integer_limits(got["ext_limits"])
    wlo, whi = integer_limits(want["ext_limits"])
    same = (glo == wlo) & (ghi == whi)
    rep["limit_flips"] = int((glo != wlo).sum() + (ghi != whi).sum())
    rep["columns"] = int(len(wlo))
    if "sky_limits" in want and "sky_limits" in got:
        slo, shi = integer_limits(got["sky_limits"])
        tlo, thi = integer_limits(want["sky_limits"])
        rep["sky_limit_flips"] = int((slo != tlo).sum() + (shi != thi).sum())
    if "sky_model" in want and "sky_model" in got:
        a, b = np.asarray(got["sky_model"]), np.asarray(want["sky_model"])
        rep["sky_exact"] = bool(a.shape == b.shape and np.array_equal(a, b, equal_nan=True)
                                and np.array_equal(got["sky_uncertainty"], want["sky_uncertainty"], equal_nan=True))
        if a.shape == b.shape and a.size:
            with np.errstate(invalid="ignore", divide="ignore"):
                rep["sky_max_rel"] = float(np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
    rep["scale_exact"] = bool(float(got["scale"]) == float(want["scale"]))
    rep["median_params_max"] = float(param_errors(got["median_params"], want["median_params"], nspat).max())
    for key in ("sky_params", "ext_params"):
        if key in want and len(want[key]) and key in got and np.shape(got[key])[0] == np.shape(want[key])[0]:
            rep[key] = _param_block(got[key], want[key], nspat)
            if control and key in control and np.shape(control[key]) == np.shape(want[key]):
                rep[key]["reference_self_noise"] = _param_block(control[key], want[key], nspat)
    flux_max = 0.0
    for key in SPECTRA:
        ref = np.asarray(want[key], dtype=float)
        val = np.asarray(got[key], dtype=float)
        nz = ref[np.isfinite(ref) & (ref != 0)]
        floor = 1e-5 * np.median(np.abs(nz)) if nz.size else 1.0
        err = np.abs(val - ref) / np.maximum(np.abs(ref), floor)
        sel = same & np.isfinite(ref)
        rep[key + "_max_rel"] = float(err[sel].max()) if sel.any() else 0.0
        rep[key + "_nan_pattern_equal"] = bool(np.array_equal(np.isnan(val), np.isnan(ref)))
        flux_max = max(flux_max, rep[key + "_max_rel"])
    rep["flux_max_rel"] = flux_max
    both = [rep[k] for k in ("sky_params", "ext_params") if k in rep]
    if both:
        rep["param_p99"] = float(max(max(b["p99"]) for b in both))
        rep["param_max"] = float(max(max(b["max"]) for b in both))
        rep["fits_beyond_1e-6"] = int(sum(b["fits_beyond_1e-6"] for b in both))
        rep["fits"] = int(sum(b["fits"] for b in both))
    return rep


def control_of(gold: dict) -> dict:
    """The ``ctl_*`` entries of a golden file as a stage dictionary."""
    return {k[4:]: v for k, v in gold.items() if k.startswith("ctl_")}
