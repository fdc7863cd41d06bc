"""Drop-in mirror of ``motes.sky`` (/root/reference/src/motes/sky.py) backed by CUDA kernels.

The bootstrap draws come from numpy's legacy global ``RandomState`` exactly as in the
reference: the mirror hands the current ``np.random.get_state()`` to the device, the kernel
consumes the draws ``np.random.choice`` would consume (device MT19937, masked rejection,
row-major fill) and the advanced state is installed with ``np.random.set_state`` afterwards.
Seed ``np.random.seed(k)`` before a frame and the sky model is bit-identical to the reference's.
"""

from __future__ import annotations

import logging

import numpy as np

from . import _lib, tracing

logger = logging.getLogger("motes")


def _take_global_rng():
    kind, key, pos, has_gauss, cached = np.random.get_state()
    state = np.empty(625, dtype=np.uint32)
    state[:624] = key
    state[624] = pos
    return state, (kind, has_gauss, cached)


def _put_global_rng(state, rest):
    kind, has_gauss, cached = rest
    np.random.set_state((kind, state[:624].copy(), int(state[624]), has_gauss, cached))


def subtract_sky(background_spatial_lo_limit, background_spatial_hi_limit, frame_dict, axes_dict, parameters,
                 header_parameters, input_file_name, handle=None, trace=None):
    """sky.subtract_sky (sky.py:257-399).  ``frame_dict['data']`` / ``['errs']`` and the two limit arrays are
    updated in place; ``frame_dict['sky_model']`` / ``['sky_uncertainty']`` are added.  Raises RuntimeError
    (after the reference's five critical log lines) when a column has no sky pixel."""
    h = handle or _lib.default_handle()
    order = int(parameters.sky_order)
    data = _lib.f64(frame_dict["data"])
    errs = _lib.f64(frame_dict["errs"])
    in_place = data is frame_dict["data"] and errs is frame_dict["errs"]
    if not in_place:
        data, errs = data.copy(), errs.copy()
    nspat, ndisp = data.shape
    lo = _lib.f64(background_spatial_lo_limit).copy()
    hi = _lib.f64(background_spatial_hi_limit).copy()
    sky_shape = (ndisp,) if order == 0 else (nspat, ndisp)
    sky, sky_err = np.zeros(sky_shape), np.zeros(sky_shape)
    flag, n_sky, n_good = (np.zeros(ndisp, dtype=np.int32) for _ in range(3))
    state, rest = _take_global_rng()
    if order > 0 and rest[1]:
        raise ValueError("np.random holds a cached Gaussian deviate (has_gauss=1); call np.random.seed() or draw one "
                         "np.random.normal() before a polynomial sky fit so the device can replay the stream")
    with h.lock:
        rc = _lib.lib.motes_sky_subtract_host(
            h.raw, _lib.ptr(data), _lib.ptr(errs), nspat, ndisp, _lib.ptr(lo), _lib.ptr(hi),
            int(axes_dict["data_spatial_floor"]), order, _lib.ptr(state, _lib.c_u32p), _lib.ptr(sky),
            _lib.ptr(sky_err), _lib.ptr(flag, _lib.c_ip), _lib.ptr(n_sky, _lib.c_ip), _lib.ptr(n_good, _lib.c_ip))
    if rc == _lib.E_NOSKY:
        limit = round(min(nspat, ndisp) / (2 * header_parameters["seeing"]), 1)
        logger.critical("No pixels contained inside sky region.")
        logger.critical("SKY_FWHM_MULTIPLIER is probably too large.")
        logger.critical("SKY_FWHM_MULTIPLIER < %s is recommended in this case.", limit)
        logger.critical("Enlarging the 2D spectrum region in reg.txt is also a viable solution")
        logger.critical("Raising RuntimeError")
        raise RuntimeError
    _lib.check(rc)
    _put_global_rng(state, rest)
    # in-place semantics of the reference: the caller's limit arrays are clamped (sky.py:303-307) ...
    np.copyto(np.asarray(background_spatial_lo_limit), lo)
    np.copyto(np.asarray(background_spatial_hi_limit), hi)
    # ... and data / errs hold the sky-subtracted frame
    if not in_place:
        frame_dict["data"] = data
        frame_dict["errs"] = errs
    modelled = flag == 0
    if order == 0:
        frame_dict["sky_model"] = np.tile(sky[modelled], (nspat, 1))          # sky.py:383-388
        frame_dict["sky_uncertainty"] = np.tile(sky_err[modelled], (nspat, 1))
    else:
        frame_dict["sky_model"] = np.ascontiguousarray(sky[:, modelled])      # sky.py:390-391
        frame_dict["sky_uncertainty"] = np.ascontiguousarray(sky_err[:, modelled])
    if trace is not None:
        trace.update(n_sky=n_sky, n_good=n_good, flag=flag, sky=sky, sky_err=sky_err)
    return frame_dict


def sky_locator(frame_dict, axes_dict, data_scaling_factor, header_parameters, bin_parameters, motes_parameters,
                input_file_name, handle=None):
    """sky.sky_locator (sky.py:46-174): ``(frame_dict, ndarray(nbins, 8), [lower, upper])``."""
    logger.info("Fitting Moffat Functions to each bin to localise sky.")
    params, table, _ = tracing.moffat_fitting_all(bin_parameters, frame_dict["data"], axes_dict, data_scaling_factor,
                                                  header_parameters, motes_parameters.sky_fwhm_multiplier, handle=handle)
    logger.info("Fitting complete. Drawing target/sky boundaries.")
    limits = tracing.interpolate_extraction_lims(table, axes_dict["dispersion_axis_length"], handle=handle)
    logger.info("Subtracting sky.")
    frame_dict = subtract_sky(limits[0], limits[1], frame_dict, axes_dict, motes_parameters, header_parameters,
                              input_file_name, handle=handle)
    return frame_dict, params, limits
