"""CPU oracle for the MOTES per-column extraction hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference algorithm.  It is imported by
``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s CPU-baseline legs and
by nothing else: the product (``motes_b200``) never imports it and has no CPU path.

Parity status: **pinned against the reference itself**.  The reference has no tests
or golden vectors of its own, so ``tests/golden/make_golden.py``
imports the unmodified reference from ``/root/reference/src`` in the build
container, runs it stage by stage on seeded synthetic frames and commits the
outputs; ``tests/test_oracle_golden.py`` requires this file to reproduce them
bit for bit.

Third-party arithmetic on the path (not under ``/root/reference``):

* ``scipy.optimize.least_squares(method="trf")`` (reference lock file pins scipy
  1.14.1, this image has 1.18.1) -- called here exactly as the reference calls it
  (``tracing.py:591-598``); ``oracle/trf_restated.py`` restates its algorithm.
* numpy's legacy global RNG (``np.random.choice`` / ``np.random.normal``,
  ``sky.py:34-38,219-223``) -- called here as the reference does;
  ``oracle/legacy_rng.py`` restates the MT19937 stream and its consumption.
* ``scipy.interpolate.interp1d`` (``tracing.py:397``) -- restated here with the two
  formulas it dispatches to (``np.interp`` and the two-weight form).

Citations are to ``/root/reference/src/motes/<file>:<lines>``.
"""

from __future__ import annotations

import numpy as np
from scipy.optimize import least_squares

BETA_START = 4.765        # tracing.py:581
N_BOOTSTRAP = 100         # sky.py:36,216


# --------------------------------------------------------------------------- model
def moffat_model(params, axis):
    """Moffat profile on a linear background (common.py:65-94)."""
    amp, cen, alpha, beta, level, grad = params
    off = axis - cen
    unit = (1 + ((off * off) / (alpha * alpha))) ** -beta
    return amp * unit + level + (axis * grad)


def moffat_fwhm(alpha, beta):
    """FWHM = 2 alpha sqrt(2^(1/beta) - 1) (common.py:97-114)."""
    return 2 * alpha * np.sqrt((2 ** (1 / beta)) - 1)


def _residual(params, axis, profile):
    return moffat_model(params, axis) - profile       # tracing.py:610-639


# --------------------------------------------------------------------------- fit
def fit_start_and_bounds(profile, seeing, pixel_resolution):
    """x0, lower and upper bounds of the bounded fit (tracing.py:578-588)."""
    peak = int(np.argmax(profile))
    alpha0 = seeing / pixel_resolution
    x0 = [
        np.nanmedian(np.sort(profile)[-3:]),
        peak,
        alpha0,
        BETA_START,
        np.median(np.concatenate((profile[:5], profile[-5:]))),
        0.0,
    ]
    lower = [0.0, peak - 1.0, 0.0, 0.0, -np.inf, -np.inf]
    upper = [np.inf, peak + 1, 5 * alpha0, 5.0, np.inf, np.inf]
    return x0, lower, upper


def fit_moffat(axis, profile, seeing, pixel_resolution, solver=None, info=None):
    """Bounded TRF fit of the six Moffat parameters (tracing.py:550-607).

    ``solver`` defaults to scipy's ``least_squares`` as in the reference; tests pass
    ``oracle.trf_restated.trf_bounded`` to check the restatement.  ``info`` (a dict)
    receives nfev / njev / status when given.
    """
    x0, lower, upper = fit_start_and_bounds(profile, seeing, pixel_resolution)
    if solver is None:
        res = least_squares(_residual, x0, bounds=(lower, upper), args=(axis, profile),
                            method="trf", ftol=1e-12)
    else:
        res = solver(_residual, x0, lower, upper, args=(axis, profile), ftol=1e-12)
    if info is not None:
        info.update(nfev=int(res.nfev), njev=int(res.njev), status=int(res.status),
                    cost=float(res.cost))
    return [res.x[k] for k in range(6)]


def scale_factor(median_profile):
    """10**|floor(log10|median|)| (tracing.py:435-437)."""
    mid = np.nanmedian(median_profile)
    return 10 ** np.abs(np.floor(np.log10(np.abs(mid))))


def fit_median_profile(median_profile, header, axis, solver=None):
    """tracing.median_moffat_fitting (tracing.py:403-456); overwrites header['seeing'] with FWHM in pixels."""
    scale = scale_factor(median_profile)
    p = fit_moffat(axis, median_profile * scale, header["seeing"],
                   header["pixel_resolution"], solver)
    header["seeing"] = moffat_fwhm(p[2], p[3])
    for k in (0, 4, 5):
        p[k] /= scale
    return p, scale


def limits_from_fit(params, width_multiplier=3.0):
    """[c - k FWHM, c + k FWHM, c] (tracing.py:642-674)."""
    reach = width_multiplier * moffat_fwhm(params[2], params[3])
    return [params[1] - reach, params[1] + reach, params[1]]


# --------------------------------------------------------------------------- bins
def _window_snr(data, errs):
    """tracing.get_bin_snr (tracing.py:257-278)."""
    return np.nansum(data) / np.sqrt(np.nansum(errs * errs))


def _has_short_row(qual, first, last, min_cols):
    """True when any spatial row holds <= min_cols good pixels in [first,last) (tracing.py:281-312)."""
    return bool(np.any(np.sum(qual[:, first:last], axis=1) <= min_cols))


def snr_bins(frame, row_lo, row_hi, ndisp, args, has_sky=False):
    """Greedy centre-outward S/N binning of the dispersion axis (tracing.py:114-254).

    Returns ``[[first, last, snr], ...]`` sorted by ``first``.
    """
    threshold = args.sky_snr_limit if (args.subtract_sky and has_sky) else args.extraction_snr_limit
    min_cols = args.minimum_column_limit
    data, errs, qual = frame["data"], frame["errs"], frame["qual"]
    centre = int(ndisp / 2)
    bins = []

    # towards column 0 (tracing.py:177-210)
    edge, width = centre, 0
    while edge - width > 0:
        snr = 0.0
        while snr <= threshold:
            width += 1
            if edge - width < 0:
                width = int(edge)
                break
            if _has_short_row(qual, edge - width, edge, min_cols):
                continue
            snr = _window_snr(data[row_lo:row_hi, edge - width:edge],
                              errs[row_lo:row_hi, edge - width:edge])
        bins.append([int(edge - width), int(edge), snr])
        edge -= width
        width = 0

    # towards the last column (tracing.py:212-249)
    edge, width = centre, 0
    while edge + width < ndisp:
        snr = 0.0
        while snr <= threshold:
            width += 1
            if edge + width > ndisp:
                width = int(ndisp - edge)
                break
            if _has_short_row(qual, edge, edge + width, min_cols):
                continue
            snr = _window_snr(data[row_lo:row_hi, edge:edge + width],
                              errs[row_lo:row_hi, edge:edge + width])
        bins.append([int(edge), int(edge + width), snr])
        edge += width
        width = 0

    bins.sort(key=lambda b: b[0])
    return bins


def bin_profile(data, first, last):
    """Per-row nanmedian over a bin, leaving out chip-gap columns (tracing.py:505-507)."""
    block = data[:, first:last]
    keep = np.where(np.median(block, axis=0) != 1)[0]
    return np.nanmedian(block[:, keep], axis=1)


def fit_bin(one_bin, data, axes, scale, header, fwhm_multiplier, limit_rows, solver=None, info=None):
    """tracing.moffat_fitting (tracing.py:459-547): 8-vector, appended limits row, bin profile."""
    prof = bin_profile(data, one_bin[0], one_bin[1])
    p = fit_moffat(axes["spatial_axis"], prof * scale, header["seeing"],
                   header["pixel_resolution"], solver, info)
    for k in (0, 4, 5):
        p[k] /= scale
    lo, hi, cen = limits_from_fit(p, width_multiplier=fwhm_multiplier)
    limit_rows.append([(one_bin[0] + one_bin[1]) * 0.5, lo, hi, cen])
    p.append(one_bin[0] + axes["wavelength_start"])
    p.append(one_bin[1] + axes["wavelength_start"])
    return p, limit_rows, prof


# --------------------------------------------------------------------------- limits
def _interp_inside(xq, xp, fp):
    """interp1d(kind='linear', fill_value=nan) -> np.interp (scipy interpolate/_interpolate.py:397-403,487-489)."""
    return np.interp(xq, xp, fp)


def _interp_two_weight(xq, xp, fp):
    """interp1d(kind='linear', fill_value='extrapolate') (scipy interpolate/_interpolate.py:491-518)."""
    idx = np.searchsorted(xp, xq).clip(1, len(xp) - 1).astype(int)
    x0, x1 = xp[idx - 1], xp[idx]
    return ((xq - x0) / (x1 - x0)) * fp[idx] + ((x1 - xq) / (x1 - x0)) * fp[idx - 1]


def _end_gradient(values, marks):
    """tracing.extrapolate_gradient (tracing.py:85-111)."""
    y_a = np.median(values[marks[0]:marks[1]])
    y_b = np.median(values[marks[1]:marks[2]])
    x_a = np.median([marks[0], marks[1]])
    x_b = np.median([marks[1], marks[2]])
    return (y_a - y_b) / (x_a - x_b)


def interpolate_limits(limit_table, ndisp):
    """tracing.interpolate_extraction_lims (tracing.py:315-375).

    ``limit_table`` is ``(4, nbins)``: bin centre, lower, upper, Moffat centre.
    Returns ``[lower(ndisp,), upper(ndisp,)]``.
    """
    centres = limit_table[0]
    if len(centres) == 1:
        return [np.repeat(limit_table[1][0], ndisp), np.repeat(limit_table[2][0], ndisp)]

    inner_x = np.arange(centres[0], centres[-1])
    inner = [np.zeros(len(inner_x)) + _interp_inside(inner_x, centres, limit_table[k])
             for k in (1, 2)]

    # tracing.py:36-82: gradients from medians of fixed 150-sample windows
    first, last = centres[0], centres[-1]
    capped = []
    for k in range(2):
        g_short = _end_gradient(inner[k], [0, 150, 300])
        g_long = _end_gradient(inner[k], [-300, -150, -1])
        head = inner[k][0] - (g_short * first)
        tail = inner[k][-1] + (g_long * (ndisp - last))
        capped.append(np.append(np.insert(inner[k], 0, head), tail))
    full_x = np.append(np.insert(inner_x, 0, 0), ndisp)

    pixels = np.array(range(ndisp))
    return [np.zeros(ndisp) + _interp_two_weight(pixels, full_x, capped[k]) for k in range(2)]


# --------------------------------------------------------------------------- sky
def bootstrap_median(good):
    """sky.sky_median (sky.py:17-43): 100 resampled medians -> mean, std/sqrt(99)."""
    draws = np.random.choice(good, (len(good), N_BOOTSTRAP), replace=True)
    meds = np.nanmedian(draws, axis=0)
    return np.nanmean(meds), np.std(meds) / (99 ** 0.5)


def bootstrap_polynomial(good, good_err, good_axis, column_axis, order):
    """sky.sky_polynomial (sky.py:177-254): 100 Gaussian-perturbed polyfits -> per-pixel median & std."""
    cloud = np.array([np.random.normal(loc=good[j], scale=good_err[j], size=N_BOOTSTRAP)
                      for j in range(len(good))]).T
    coeff = np.zeros((N_BOOTSTRAP, order + 1))
    for j in range(N_BOOTSTRAP):
        coeff[j] += np.flip(np.polyfit(good_axis, cloud[j], order))
    powers = np.array([column_axis ** k for k in np.arange(order + 1)]).T     # (nspat, order+1)
    all_powers = np.array([powers] * N_BOOTSTRAP)
    all_coeff = np.array([[c] * len(column_axis) for c in coeff])
    samples = np.sum(all_powers * all_coeff, axis=2).T                         # (nspat, 100)
    return np.median(samples, axis=1), np.std(samples, axis=1)


class NoSkyPixels(RuntimeError):
    """A column has no pixel outside the sky limits (sky.py:323-335 raises RuntimeError)."""


def subtract_sky(lim_lo, lim_hi, frame, axes, args, trace=None):
    """sky.subtract_sky (sky.py:257-399).  Modifies ``frame`` and the two limit arrays in place.

    ``trace`` (a dict) receives per-column integer decisions: number of sky pixels,
    number kept by the 10-sigma clip, and whether the column was skipped.
    """
    data = frame["data"].T
    errs = frame["errs"].T
    floor = axes["data_spatial_floor"]
    ncol = len(lim_lo)
    nspat = data.shape[1]
    top = (axes["spatial_axis"] + floor)[-1]
    order = int(args.sky_order)
    model, model_err = [], []
    n_sky = np.zeros(ncol, dtype=np.int32)
    n_good = np.zeros(ncol, dtype=np.int32)
    skipped = np.zeros(ncol, dtype=np.uint8)

    for c in range(ncol):
        if lim_lo[c] < 0:
            lim_lo[c] = 0
        if lim_hi[c] > top:
            lim_hi[c] = top
        column = data[c]
        column_err = errs[c]
        column_axis = np.array(range(nspat)) + floor
        outside = np.logical_or(column_axis < lim_lo[c] + floor, column_axis > lim_hi[c] + floor)
        pix = column[np.where(outside)]
        pix_err = column_err[np.where(outside)]
        n_sky[c] = len(pix)
        if len(pix) == 0:
            raise NoSkyPixels(c)
        pix_axis = column_axis[outside]
        if len(set(pix)) == 1:                      # sky.py:338-339
            skipped[c] = 1
            continue
        mid = np.nanmedian(pix)
        spread = np.nanstd(pix)
        keep = np.where(np.logical_and(pix > mid - (10 * spread), pix < mid + (10 * spread)))
        good, good_axis, good_err = pix[keep], pix_axis[keep], pix_err[keep]
        n_good[c] = len(good)
        if order == 0:
            sky, sky_err = bootstrap_median(good)
        else:
            sky, sky_err = bootstrap_polynomial(good, good_err, good_axis, column_axis, order)
        model.append(sky)
        model_err.append(sky_err)
        data[c] -= sky
        errs[c] = ((errs[c] ** 2) + (sky_err ** 2)) ** 0.5

    model = np.array(model)
    model_err = np.array(model_err)
    if args.sky_order == 0:
        frame["sky_model"] = np.tile(model, (nspat, 1))
        frame["sky_uncertainty"] = np.tile(model_err, (nspat, 1))
    else:
        frame["sky_model"] = model.T
        frame["sky_uncertainty"] = model_err.T

    frame["data"] = data.T
    frame["errs"] = errs.T
    gap = np.where(np.nanmedian(frame["data"], axis=0) == 0)       # sky.py:396-397
    frame["data"][:, gap[0]] += 1
    if trace is not None:
        trace.update(n_sky=n_sky, n_good=n_good, skipped=skipped, regapped=gap[0].astype(np.int32))
    return frame


def locate_and_subtract_sky(frame, axes, scale, header, bins, args, solver=None, trace=None):
    """sky.sky_locator (sky.py:46-174)."""
    rows, params = [], []
    for b in bins:
        p, rows, _ = fit_bin(b, frame["data"], axes, scale, header, args.sky_fwhm_multiplier,
                             rows, solver)
        params.append(p)
    params = np.array(params)
    table = np.array(rows).T
    limits = interpolate_limits(table, axes["dispersion_axis_length"])
    if trace is not None:
        trace["sky_limit_table"] = table.copy()
        trace["sky_limits_unclamped"] = [limits[0].copy(), limits[1].copy()]
    frame = subtract_sky(limits[0], limits[1], frame, axes, args, trace)
    return frame, params, limits


# --------------------------------------------------------------------------- extraction
def _scrub(data2d, errs2d):
    """extraction.filter_data (extraction.py:10-36), same statement order."""
    errs2d[~np.isfinite(errs2d)] = 0.0
    data2d[~np.isfinite(errs2d)] = 0.0
    data2d[~np.isfinite(data2d)] = 0.0
    errs2d[~np.isfinite(data2d)] = 0.0
    return data2d, errs2d


def optimal_extract(data2d, errs2d, limits, bin_params, axes, trace=None):
    """extraction.optimal_extraction (extraction.py:39-210).

    ``data2d`` / ``errs2d`` are ``(ndisp, nspat)`` (transposed views as in core.py:265-267).
    Returns ``[optimal, optimal_err, aperture, aperture_err]``.  ``trace`` receives the
    integer slice bounds and the per-pixel cosmic-ray mask of every column.
    """
    data2d, errs2d = _scrub(data2d, np.abs(errs2d))
    ncol = data2d.shape[0]
    nspat = data2d.shape[1]
    opt, opt_err = np.zeros(ncol), np.zeros(ncol)
    ape, ape_err = np.zeros(ncol), np.zeros(ncol)
    lo_idx = np.zeros(ncol, dtype=np.int32)
    hi_idx = np.zeros(ncol, dtype=np.int32)
    crmask = np.zeros((ncol, nspat), dtype=np.uint8)     # 1 = pixel kept, indexed by spatial row
    which = -1
    b = np.zeros(8)
    last_bin = len(bin_params) - 1
    axis = axes["spatial_axis"]

    for i in range(ncol):
        col = data2d[i]
        if which < last_bin and (i + axes["wavelength_start"]) == bin_params[which + 1][-2]:
            which += 1
            b = bin_params[which]
        fine_lo, fine_hi = limits[0][i], limits[1][i]
        lo = int(np.floor(limits[0][i]))
        hi = int(np.ceil(limits[1][i])) + 1
        lo_idx[i], hi_idx[i] = lo, hi
        ax = axis[lo:hi]
        col = col[lo:hi]
        err = errs2d[i]
        if all(e == 0.0 for e in err):
            continue
        err[np.where(err == 0)] = np.median(err[np.where(err != 0)])
        err = err[lo:hi]
        var = err * err
        flux = np.sum(col)
        ape[i] += flux
        ape_err[i] += np.sqrt(np.sum(var))
        cen = (fine_hi + fine_lo) * 0.5
        prof = moffat_model([b[0], cen, b[2], b[3], 0.0, 0.0], ax)
        prof = prof / np.sum(prof)
        keep = np.invert([np.abs((col - (flux * prof)) ** 2) > var * 5 ** 2])[0]
        if not np.any(keep):
            keep = np.full(shape=keep.shape, fill_value=True)
        rows = np.arange(nspat)[lo:hi]
        crmask[i, rows] = keep
        if all(v == 0.0 for v in col) or all(v == 1.0 for v in col):
            f_opt = 0.0
            v_opt = 0.0
        else:
            weight = np.sum(((prof ** 2) / var)[keep])
            f_opt = np.sum(((prof * col) / var)[keep]) / weight
            v_opt = np.sum(prof[keep]) / weight
        opt[i] += f_opt
        opt_err[i] += np.sqrt(v_opt)
    if trace is not None:
        trace.update(lo_idx=lo_idx, hi_idx=hi_idx, crmask=crmask)
    return [opt, opt_err, ape, ape_err]


# --------------------------------------------------------------------------- whole frame
def run_frame(frame, axes, header, args, seed=None, solver=None):
    """Drive the stages in the order of core.motes (core.py:80-275) on in-memory arrays.

    ``frame`` / ``header`` are modified in place exactly as the reference modifies
    them.  ``seed`` (int) seeds numpy's legacy global RNG before the frame, which is
    what the parity harness does (the reference itself never seeds).  Returns a dict
    of every stage output.
    """
    if seed is not None:
        np.random.seed(seed)
    out = {}
    ndisp = axes["dispersion_axis_length"]
    median_profile = np.nanmedian(frame["data"], axis=1)
    out["median_profile"] = median_profile
    params, scale = fit_median_profile(median_profile, header, axes["spatial_axis"], solver)
    out["median_params"] = np.array(params)
    out["scale"] = float(scale)
    out["seeing_px"] = float(header["seeing"])
    region = limits_from_fit(params)
    row_lo, row_hi = int(np.floor(region[0])), int(np.ceil(region[1]))
    out["bin_rows"] = np.array([row_lo, row_hi], dtype=np.int64)

    if args.subtract_sky:
        sky_bins = snr_bins(frame, row_lo, row_hi, ndisp, args, has_sky=True)
        out["sky_bins"] = np.array(sky_bins, dtype=np.float64).reshape(-1, 3)
        trace = {}
        frame, sky_params, sky_limits = locate_and_subtract_sky(
            frame, axes, scale, header, sky_bins, args, solver, trace)
        out["sky_params"] = sky_params
        out["sky_limits"] = np.array(sky_limits)
        out["sky_limit_table"] = trace["sky_limit_table"]
        out["sky_limits_unclamped"] = np.array(trace["sky_limits_unclamped"])
        out["sky_n_sky"], out["sky_n_good"] = trace["n_sky"], trace["n_good"]
        out["sky_skipped"], out["sky_regapped"] = trace["skipped"], trace["regapped"]
        if int(args.sky_order) == 0:
            out["sky_model"] = np.array(frame["sky_model"][0]) if frame["sky_model"].size else np.zeros(0)
            out["sky_uncertainty"] = (np.array(frame["sky_uncertainty"][0])
                                      if frame["sky_uncertainty"].size else np.zeros(0))
        else:
            out["sky_model"] = np.array(frame["sky_model"])
            out["sky_uncertainty"] = np.array(frame["sky_uncertainty"])

    ext_bins = snr_bins(frame, row_lo, row_hi, ndisp, args)
    out["ext_bins"] = np.array(ext_bins, dtype=np.float64).reshape(-1, 3)
    rows, all_params, profiles = [], [], []
    for b in ext_bins:
        p, rows, prof = fit_bin(b, frame["data"], axes, scale, header,
                                args.sky_fwhm_multiplier, rows, solver)   # core.py:208 quirk
        all_params.append(p)
        profiles.append(prof)
    all_params = np.array(all_params)
    table = np.array(rows).T
    out["ext_params"] = all_params
    out["ext_limit_table"] = table
    out["ext_profiles"] = np.array(profiles)
    limits = interpolate_limits(table, ndisp)
    out["ext_limits"] = np.array(limits)

    frame["data"] = frame["data"].T
    frame["errs"] = frame["errs"].T
    frame["qual"] = frame["qual"].T
    trace = {}
    spectra = optimal_extract(frame["data"], frame["errs"], limits, all_params, axes, trace)
    out["optimal"], out["optimal_err"], out["aperture"], out["aperture_err"] = spectra
    out["lo_idx"], out["hi_idx"], out["crmask"] = trace["lo_idx"], trace["hi_idx"], trace["crmask"]
    # hand the frames back in (nspat, ndisp) orientation for inspection
    out["data_after"] = np.ascontiguousarray(frame["data"].T)
    out["errs_after"] = np.ascontiguousarray(frame["errs"].T)
    return out
