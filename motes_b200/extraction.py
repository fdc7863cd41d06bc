"""Drop-in mirror of ``motes.extraction`` (/root/reference/src/motes/extraction.py) backed by CUDA kernels."""

from __future__ import annotations

import logging

import numpy as np

from . import _lib

logger = logging.getLogger("motes")


def optimal_extraction(data_2d, errs_2d, extraction_limits, bin_parameters, axes_dict,
                       handle=None, return_masks=False):
    """extraction.optimal_extraction (extraction.py:39-210).

    ``data_2d`` / ``errs_2d`` are the transposed ``(ndisp, nspat)`` views ``core.motes`` passes
    (core.py:265-275).  As in the reference, non-finite pixels of ``data_2d`` are set to 0
    in place (extraction.py:31-34, 86) and ``errs_2d`` is left untouched.  Returns
    ``[optimal, optimal_err, aperture, aperture_err]``; with ``return_masks`` also a dict with the
    integer slice bounds and the 5-sigma mask of every column.
    """
    logger.info("Performing optimal extraction of 1D spectrum.")
    h = handle or _lib.default_handle()
    ndisp, nspat = np.shape(data_2d)
    frame = _lib.f64(np.asarray(data_2d).T)          # (nspat, ndisp), dispersion fastest
    errs = _lib.f64(np.asarray(errs_2d).T)
    lo = _lib.f64(extraction_limits[0])
    hi = _lib.f64(extraction_limits[1])
    bins = _lib.f64(np.atleast_2d(bin_parameters))
    spectra = np.zeros((4, ndisp))
    lo_idx = np.zeros(ndisp, dtype=np.int32)
    hi_idx = np.zeros(ndisp, dtype=np.int32)
    mask = np.zeros((ndisp, nspat), dtype=np.uint8) if return_masks else None
    with h.lock:
        _lib.check(_lib.lib.motes_optimal_extract_host(
            h.raw, _lib.ptr(frame), _lib.ptr(errs), nspat, ndisp, _lib.ptr(lo), _lib.ptr(hi), _lib.ptr(bins),
            bins.shape[0], int(axes_dict["wavelength_start"]), _lib.ptr(spectra),
            _lib.ptr(lo_idx, _lib.c_ip), _lib.ptr(hi_idx, _lib.c_ip),
            _lib.ptr(mask, _lib.c_u8p) if mask is not None else None))
    # reproduce filter_data's in-place scrub of the caller's array
    bad = ~np.isfinite(data_2d)
    if bad.any():
        data_2d[bad] = 0.0
    out = [spectra[0].copy(), spectra[1].copy(), spectra[2].copy(), spectra[3].copy()]
    if return_masks:
        return out, dict(lo_idx=lo_idx, hi_idx=hi_idx, crmask=mask)
    return out
