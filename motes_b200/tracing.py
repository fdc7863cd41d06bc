"""Drop-in mirror of ``motes.tracing`` (/root/reference/src/motes/tracing.py) backed by CUDA kernels.

Same function names, argument meaning, return shapes and error behaviour as the
reference; the work is done by ``libmotes_b200.so`` through ``motes_b200._lib``.
"""

from __future__ import annotations

import logging

import numpy as np

from . import _lib, common

logger = logging.getLogger("motes")

_FIT_ERRORS = {
    -1: "Initial guess is outside of provided bounds",
    -2: "Residuals are not finite in the initial point.",
    -3: "Each lower bound must be strictly less than each upper bound.",
}


def moffat_fit_batch(profiles, alpha0, handle=None, full_output=False):
    """Fit ``nfit`` profiles at once (rows of ``profiles``); the batched form of ``moffat_least_squares``.

    Raises ``ValueError`` like scipy does when a fit cannot start (tracing.py:591-598 via
    scipy ``least_squares`` checks).
    """
    h = handle or _lib.default_handle()
    prof = _lib.f64(np.atleast_2d(profiles))
    nfit, nspat = prof.shape
    a0 = _lib.f64(np.broadcast_to(np.asarray(alpha0, dtype=np.float64), (nfit,)))
    params = np.empty((nfit, 6))
    cost = np.empty(nfit)
    nfev = np.empty(nfit, dtype=np.int32)
    njev = np.empty(nfit, dtype=np.int32)
    status = np.empty(nfit, dtype=np.int32)
    with h.lock:
        _lib.check(_lib.lib.motes_moffat_fit_batch_host(
            h.raw, _lib.ptr(prof), nfit, nspat, _lib.ptr(a0), _lib.ptr(params), _lib.ptr(cost),
            _lib.ptr(nfev, _lib.c_ip), _lib.ptr(njev, _lib.c_ip), _lib.ptr(status, _lib.c_ip)))
    bad = np.nonzero(status < 0)[0]
    if bad.size:
        raise ValueError(_FIT_ERRORS.get(int(status[bad[0]]), "Moffat fit failed") + f" (fit {int(bad[0])})")
    if full_output:
        return params, dict(cost=cost, nfev=nfev, njev=njev, status=status)
    return params


def moffat_least_squares(r, col, seeing, pixel_resolution):
    """tracing.moffat_least_squares (tracing.py:550-607): list of the six best-fit parameters.

    ``r`` must be the pixel axis 0..n-1 that ``data_harvest`` builds (io/motesio.py:94); the
    device kernel generates it instead of reading it.
    """
    r = np.asarray(r, dtype=np.float64)
    if r.shape != np.shape(col) or not np.array_equal(r, np.arange(len(r), dtype=np.float64)):
        raise ValueError("motes_b200 fits on the pixel axis 0..n-1 produced by motesio.data_harvest")
    params = moffat_fit_batch(np.asarray(col, dtype=np.float64)[None, :], seeing / pixel_resolution)
    return [params[0, k] for k in range(6)]


def set_extraction_limits(moffat_parameters, width_multiplier=3.0):
    """tracing.set_extraction_limits (tracing.py:642-674)."""
    fwhm = common.moffat_fwhm(moffat_parameters[2], moffat_parameters[3])
    lower = moffat_parameters[1] - (width_multiplier * fwhm)
    upper = moffat_parameters[1] + (width_multiplier * fwhm)
    return [lower, upper, moffat_parameters[1]]


def median_moffat_fitting(median_profile, header_parameters, spatial_axis):
    """tracing.median_moffat_fitting (tracing.py:403-456); overwrites header_parameters['seeing'] (tracing.py:449)."""
    centre_value = np.nanmedian(median_profile)
    data_scaling_factor = 10 ** np.abs(np.floor(np.log10(np.abs(centre_value))))
    params = moffat_least_squares(spatial_axis, median_profile * data_scaling_factor,
                                  header_parameters["seeing"], header_parameters["pixel_resolution"])
    header_parameters["seeing"] = common.moffat_fwhm(params[2], params[3])
    for k in (0, 4, 5):
        params[k] /= data_scaling_factor
    return params, data_scaling_factor
