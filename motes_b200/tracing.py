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


def moffat_least_squares(r, col, seeing, pixel_resolution, handle=None):
    """tracing.moffat_least_squares (tracing.py:550-607): list of the six best-fit parameters.

    ``r`` must be the pixel axis 0..n-1 that ``data_harvest`` builds (io/motesio.py:94); the
    device kernel generates it instead of reading it.
    """
    r = np.asarray(r, dtype=np.float64)
    if r.shape != np.shape(col) or not np.array_equal(r, np.arange(len(r), dtype=np.float64)):
        raise ValueError("motes_b200 fits on the pixel axis 0..n-1 produced by motesio.data_harvest")
    params = moffat_fit_batch(np.asarray(col, dtype=np.float64)[None, :], seeing / pixel_resolution, handle=handle)
    return [params[0, k] for k in range(6)]


def set_extraction_limits(moffat_parameters, width_multiplier=3.0):
    """tracing.set_extraction_limits (tracing.py:642-674)."""
    fwhm = common.moffat_fwhm(moffat_parameters[2], moffat_parameters[3])
    lower = moffat_parameters[1] - (width_multiplier * fwhm)
    upper = moffat_parameters[1] + (width_multiplier * fwhm)
    return [lower, upper, moffat_parameters[1]]


def median_moffat_fitting(median_profile, header_parameters, spatial_axis, handle=None):
    """tracing.median_moffat_fitting (tracing.py:403-456); overwrites header_parameters['seeing'] (tracing.py:449)."""
    centre_value = np.nanmedian(median_profile)
    data_scaling_factor = 10 ** np.abs(np.floor(np.log10(np.abs(centre_value))))
    params = moffat_least_squares(spatial_axis, median_profile * data_scaling_factor,
                                  header_parameters["seeing"], header_parameters["pixel_resolution"], handle=handle)
    header_parameters["seeing"] = common.moffat_fwhm(params[2], params[3])
    for k in (0, 4, 5):
        params[k] /= data_scaling_factor
    return params, data_scaling_factor


def _qual_u8(qual):
    """Quality frames are float64 {0,1} in the reference (io/gmosio.py:50-52); the device reads uint8."""
    return np.ascontiguousarray(np.asarray(qual) != 0, dtype=np.uint8)


def get_bins(frame_dict, spatial_lo_limit, spatial_hi_limit, dispersion_axis_length, parameter_dict,
             has_sky=False, handle=None):
    """tracing.get_bins (tracing.py:114-254): ``([[first, last, snr], ...], frame_dict)``."""
    if parameter_dict.subtract_sky and has_sky:
        minimum_snr = parameter_dict.sky_snr_limit
    else:
        minimum_snr = parameter_dict.extraction_snr_limit
    logger.info("Determining spectrum localisation bins on dispersion axis.")
    logger.info("User defined S/N threshold = %s.", minimum_snr)
    h = handle or _lib.default_handle()
    data, errs = _lib.f64(frame_dict["data"]), _lib.f64(frame_dict["errs"])
    qual = _qual_u8(frame_dict["qual"])
    nspat, ndisp = data.shape
    if ndisp != int(dispersion_axis_length):
        raise ValueError("dispersion_axis_length does not match the frame")
    cap = _lib.lib.motes_bin_capacity(ndisp, float(parameter_dict.minimum_column_limit))
    edges = np.zeros((cap, 2), dtype=np.int32)
    snr = np.zeros(cap)
    nb = np.zeros(1, dtype=np.int32)
    with h.lock:
        _lib.check(_lib.lib.motes_get_bins_host(
            h.raw, _lib.ptr(data), _lib.ptr(errs), _lib.ptr(qual, _lib.c_u8p), nspat, ndisp,
            int(spatial_lo_limit), int(spatial_hi_limit), float(minimum_snr),
            float(parameter_dict.minimum_column_limit), cap, _lib.ptr(edges, _lib.c_ip), _lib.ptr(snr),
            _lib.ptr(nb, _lib.c_ip), None))
    n = int(nb[0])
    return [[int(edges[i, 0]), int(edges[i, 1]), float(snr[i])] for i in range(n)], frame_dict


def bin_profiles(data, bins, handle=None):
    """Median spatial profile of every bin, chip-gap columns left out (tracing.py:505-507), batched."""
    h = handle or _lib.default_handle()
    data = _lib.f64(data)
    nspat, ndisp = data.shape
    edges = np.ascontiguousarray([[int(b[0]), int(b[1])] for b in bins], dtype=np.int32)
    out = np.empty((len(edges), nspat))
    with h.lock:
        _lib.check(_lib.lib.motes_bin_profiles_host(h.raw, _lib.ptr(data), nspat, ndisp, _lib.ptr(edges, _lib.c_ip),
                                                    len(edges), _lib.ptr(out)))
    return out


def moffat_fitting(each_bin, data, axes_dict, data_scaling_factor, header_parameters, fwhm_multiplier,
                   extraction_limits, handle=None):
    """tracing.moffat_fitting (tracing.py:459-547) for one bin: ``(list[8], extraction_limits, bin_data)``."""
    bin_data = bin_profiles(data, [each_bin], handle=handle)[0]
    params = moffat_least_squares(axes_dict["spatial_axis"], bin_data * data_scaling_factor,
                                  header_parameters["seeing"], header_parameters["pixel_resolution"], handle=handle)
    for k in (0, 4, 5):
        params[k] /= data_scaling_factor
    lower, upper, centre = set_extraction_limits(params, width_multiplier=fwhm_multiplier)
    extraction_limits.append([(each_bin[0] + each_bin[1]) * 0.5, lower, upper, centre])
    params.append(each_bin[0] + axes_dict["wavelength_start"])
    params.append(each_bin[1] + axes_dict["wavelength_start"])
    return params, extraction_limits, bin_data


def moffat_fitting_all(bins, data, axes_dict, data_scaling_factor, header_parameters, fwhm_multiplier, handle=None):
    """All bins of a frame in two launches (profiles, fits): what the per-bin loops at core.py:200-212
    and sky.py:96-109 compute.  Returns ``(ndarray(nbins, 8), limit table (4, nbins), profiles)``."""
    profiles = bin_profiles(data, bins, handle=handle)
    alpha0 = header_parameters["seeing"] / header_parameters["pixel_resolution"]
    fitted = moffat_fit_batch(profiles * data_scaling_factor, alpha0, handle=handle)
    rows, table = [], []
    for b, p in zip(bins, fitted):
        p = [p[k] for k in range(6)]
        for k in (0, 4, 5):
            p[k] /= data_scaling_factor
        lower, upper, centre = set_extraction_limits(p, width_multiplier=fwhm_multiplier)
        table.append([(b[0] + b[1]) * 0.5, lower, upper, centre])
        rows.append(p + [b[0] + axes_dict["wavelength_start"], b[1] + axes_dict["wavelength_start"]])
    return np.array(rows), np.array(table).T, profiles


def interpolate_extraction_lims(extraction_limits, dispersion_axis_length, handle=None):
    """tracing.interpolate_extraction_lims (tracing.py:315-375): ``[lower(ndisp,), upper(ndisp,)]``."""
    h = handle or _lib.default_handle()
    table = _lib.f64(np.asarray(extraction_limits))
    nbins = table.shape[1]
    ndisp = int(dispersion_axis_length)
    lo, hi = np.empty(ndisp), np.empty(ndisp)
    with h.lock:
        _lib.check(_lib.lib.motes_interp_limits_host(h.raw, _lib.ptr(table), nbins, ndisp, _lib.ptr(lo), _lib.ptr(hi)))
    return [lo, hi]
