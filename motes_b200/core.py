"""``core.motes`` (/root/reference/src/motes/core.py:31-306) on the GPU: read ``./inputs/reg.txt``, harvest every
listed FITS file, run the whole extraction path on the device and, with ``save`` set, write the same product file
the reference writes (``motes_b200.io.motesio.save_fits``).

Two ways to run the frames:

* ``seeds=None`` (default) - the reference's semantics: one file after the other, the bootstrap sky drawing from
  numpy's global legacy generator exactly where the reference would (the device consumes the draws and the advanced
  state is put back with ``np.random.set_state``).  ``np.random.seed(k); motes(args)`` therefore gives the reference's
  sky model bit for bit.
* ``seeds=k`` or a list - batch mode: all files of one frame shape go to the device in ONE ``run_frames`` call, file
  ``i`` with the generator state of ``np.random.seed(seeds[i])`` (``k + i`` for an integer).  Frames are independent,
  so each product equals what a sequential run with that seed gives.

Diagnostics plots, per-file log switching and the command line stay in the reference (out of scope).
"""

from __future__ import annotations

import copy
import logging
import os

import numpy as np

from . import common, pipeline, sky as _sky, tracing
from .io import motesio

logger = logging.getLogger("motes")


def _frame_products(res, i, frame_dict, header_parameters, motes_args):
    """Per-file view of batch result ``res[.][i]`` in the shapes ``core.motes`` hands to ``save_fits``; leaves
    ``frame_dict`` / ``header_parameters`` as the reference leaves them (core.py:257-259, tracing.py:449, sky.py:383-399)."""
    pipeline.raise_for_status(int(res["frame_status"][i]))
    ne, ns = int(res["n_ext_bins"][i]), int(res["n_sky_bins"][i])
    nspat = res["data"].shape[1]
    header_parameters["seeing"] = float(res["seeing_px"][i])
    out = {
        "moffat_profile_parameters": list(res["median_params"][i]),
        "data_scaling_factor": float(res["scale"][i]),
        "moffat_parameters_all_bins": np.array(res["ext_params"][i, :ne]),
        "final_extraction_limits": np.array(res["ext_limits"][i]),
        "extracted": [np.array(res["spectra"][i, k]) for k in range(4)],
        "moffat_parameters_all_sky_bins": None,
        "sky_extraction_limits": None,
    }
    if motes_args.subtract_sky:
        modelled = res["sky_flag"][i] == 0
        out["moffat_parameters_all_sky_bins"] = np.array(res["sky_params"][i, :ns])
        out["sky_extraction_limits"] = [np.array(res["sky_limits"][i, 0]), np.array(res["sky_limits"][i, 1])]
        if int(motes_args.sky_order) == 0:
            frame_dict["sky_model"] = np.tile(res["sky_model"][i][modelled], (nspat, 1))
            frame_dict["sky_uncertainty"] = np.tile(res["sky_uncertainty"][i][modelled], (nspat, 1))
        else:
            frame_dict["sky_model"] = np.ascontiguousarray(res["sky_model"][i][:, modelled])
            frame_dict["sky_uncertainty"] = np.ascontiguousarray(res["sky_uncertainty"][i][:, modelled])
    frame_dict["data"] = res["data"][i].T          # the reference ends with the frames transposed (core.py:257-259)
    frame_dict["errs"] = res["errs"][i].T
    frame_dict["qual"] = np.asarray(frame_dict["qual"]).T
    out["frame_dict"], out["header_parameters"] = frame_dict, header_parameters
    return out


def fit_diagnostics(data, axes_dict, bin_parameters, handle=None):
    """The numbers behind the reference's ``-dm`` plots (core.py:214-229, diagnostics.plot_fitted_spatial_profile), without
    the plotting: per bin the median spatial profile the fit saw (device, ``tracing.bin_profiles``) and the fitted
    Moffat curve on the 5x oversampled spatial axis.  ``data`` is the ``(nspat, ndisp)`` frame the bins were fitted on
    (the harvested frame for sky bins, the sky-subtracted one for extraction bins); ``bin_parameters`` is the
    ``(nbins, 8)`` table with the bin edges in columns 6-7."""
    bin_parameters = np.asarray(bin_parameters)
    edges = (bin_parameters[:, 6:8] - int(axes_dict["wavelength_start"])).astype(np.int32)
    return {
        "bin_edges": edges,
        "spatial_axis": axes_dict["spatial_axis"] + axes_dict["data_spatial_floor"],
        "bin_profiles": tracing.bin_profiles(data, edges, handle=handle),
        "hi_resolution_spatial_axis": axes_dict["hi_resolution_spatial_axis"] + axes_dict["data_spatial_floor"],
        "fitted_curves": np.array([common.moffat(row[:6], axes_dict["hi_resolution_spatial_axis"]) for row in bin_parameters]),
    }


def localisation_lines(axes_dict, boundaries):
    """The lines the reference overplots for ``-dl`` / ``-dk`` (diagnostics.get_draw_lines, diagnostics.py:75-109):
    ``boundaries`` is either the ``(4, nbins)`` limit table (bin centres, lower, upper, Moffat centre) or the
    ``[lower(ndisp,), upper(ndisp,)]`` pair of interpolated limits; returns ``[x, lower, x, upper]`` in frame
    coordinates (wavelength start and spatial floor added)."""
    boundaries = [np.asarray(b) for b in boundaries]
    if len(boundaries) > 2:
        x = boundaries[0] + axes_dict["wavelength_start"]
        lower, upper = boundaries[1], boundaries[2]
    else:
        x = np.arange(int(axes_dict["dispersion_axis_length"])) + axes_dict["wavelength_start"]
        lower, upper = boundaries[0], boundaries[1]
    floor = axes_dict["data_spatial_floor"]
    return [x, lower + floor, x, upper + floor]


def sky_fit_diagnostics(frame_before, sky_limits, sky_model, sky_uncertainty, axes_dict, columns):
    """The numbers behind the reference's ``-df`` plots (sky.py:236-252, diagnostics.plot_sky_fit), without the plotting,
    for the dispersion ``columns`` asked for: per column the good sky pixels the model was bootstrapped from (the sky
    region outside the clamped limits, single-pass 10 sigma clip, sky.py:303-354), their errors and spatial positions,
    the column's spatial axis, and the sky model and uncertainty along it as the device computed them.
    ``frame_before`` holds the ``(nspat, ndisp)`` data / errs planes BEFORE sky subtraction; ``sky_limits`` the clamped
    ``[lower, upper]`` pair ``sky.sky_locator`` returns; ``sky_model`` / ``sky_uncertainty`` the planes it left in
    ``frame_dict`` (order 0: one value per modelled column tiled over the rows; order > 0: full planes)."""
    data = np.asarray(frame_before["data"], dtype=np.float64)
    errs = np.asarray(frame_before["errs"], dtype=np.float64)
    nspat, ndisp = data.shape
    floor = axes_dict["data_spatial_floor"]
    column_axis = np.arange(nspat) + floor
    # sky_model / sky_uncertainty only hold the modelled columns (constant chip-gap columns are skipped, sky.py:338-339)
    modelled = np.array([len(set(data[(column_axis < sky_limits[0][c] + floor) | (column_axis > sky_limits[1][c] + floor), c])) != 1
                         for c in range(ndisp)])
    where = np.cumsum(modelled) - 1
    out = {}
    for c in columns:
        in_sky = (column_axis < sky_limits[0][c] + floor) | (column_axis > sky_limits[1][c] + floor)
        sky_pixels, sky_errors, sky_axis = data[in_sky, c], errs[in_sky, c], column_axis[in_sky]
        entry = {"column": int(c), "column_axis": column_axis, "modelled": bool(modelled[c])}
        if modelled[c]:
            med, sig = np.nanmedian(sky_pixels), np.nanstd(sky_pixels)
            good = (sky_pixels > med - 10 * sig) & (sky_pixels < med + 10 * sig)
            entry.update({"good_sky_axis": sky_axis[good], "good_sky_pixels": sky_pixels[good], "good_sky_errs": sky_errors[good],
                          "column_sky_model": np.asarray(sky_model)[:, where[c]],
                          "column_sky_model_err": np.asarray(sky_uncertainty)[:, where[c]]})
        out[int(c)] = entry
    return out


def extract_frames(frames, motes_args, rng_states=None, seeds=None, handle=None):
    """``frames``: list of ``(frame_dict, axes_dict, header_parameters)`` of ONE shape and ONE region offset.
    Returns one product dictionary per frame (see ``_frame_products``)."""
    data = np.ascontiguousarray(np.stack([np.asarray(f[0]["data"], dtype=np.float64) for f in frames]))
    errs = np.ascontiguousarray(np.stack([np.asarray(f[0]["errs"], dtype=np.float64) for f in frames]))
    qual = np.stack([np.asarray(f[0]["qual"]) for f in frames])
    res = pipeline.run_frames(data, errs, qual, motes_args, [f[2]["seeing"] for f in frames],
                              [f[2]["pixel_resolution"] for f in frames], seeds=seeds, rng_states=rng_states,
                              axes_dict=frames[0][1], copy_back=True, handle=handle)
    products = [_frame_products(res, i, f[0], f[2], motes_args) for i, f in enumerate(frames)]
    for p, f in zip(products, frames):
        p["axes_dict"] = f[1]
    if "rng_states" in res:
        for i, p in enumerate(products):
            p["rng_state"] = res["rng_states"][i]       # numpy's RandomState after the frame
    return products


def _save(entry_name, harvested, products, motes_args):
    header_parameters, frame_dict, axes_dict, original_hdu_list = harvested
    motesio.save_fits(original_hdu_list, axes_dict, header_parameters, products["extracted"], motes_args, entry_name,
                      products["moffat_profile_parameters"], frame_dict, products["moffat_parameters_all_bins"],
                      products["final_extraction_limits"], products["moffat_parameters_all_sky_bins"],
                      products["sky_extraction_limits"])


def motes(motes_args, seeds=None, handle=None):
    """Process every file of ``./inputs/reg.txt``.  Returns ``{input_file_name: products}``."""
    cwd = os.getcwd()
    logger.info("Current working directory is " + cwd)
    files_and_regions = motesio.read_regions(cwd)
    harvested = []
    for entry in files_and_regions:
        logger.info("Begin MOTES processing for ./inputs/" + entry[-1])
        h = motesio.data_harvest(cwd + "/inputs/" + entry[-1], entry[:4])
        h[1]["original_data"] = copy.deepcopy(h[1]["data"])
        h[1]["original_errs"] = copy.deepcopy(h[1]["errs"])
        harvested.append(h)
    results = {}

    if seeds is None:                                  # reference semantics: sequential, global generator
        for entry, h in zip(files_and_regions, harvested):
            header_parameters, frame_dict, axes_dict, _ = h
            state = rest = None
            if motes_args.subtract_sky:
                state, rest = _sky._take_global_rng()
                if int(motes_args.sky_order) > 0 and rest[1]:
                    raise ValueError("np.random holds a cached Gaussian deviate; seed the generator before a polynomial sky fit")
            products = extract_frames([(frame_dict, axes_dict, header_parameters)], motes_args,
                                      rng_states=None if state is None else state[None].copy(), handle=handle)[0]
            if state is not None:
                _sky._put_global_rng(products["rng_state"], rest)
            results[entry[-1][:-5]] = products
            if motes_args.save:
                _save(entry[-1][:-5], h, products, motes_args)
        logger.info("MOTES Processing completed.")
        return results

    seeds = [int(seeds) + i for i in range(len(harvested))] if np.isscalar(seeds) else [int(s) for s in seeds]
    if len(seeds) != len(harvested):
        raise ValueError("one seed per line of reg.txt is required")
    groups = {}
    for i, h in enumerate(harvested):
        key = (h[1]["data"].shape, h[2]["data_spatial_floor"], int(h[2]["wavelength_start"]))
        groups.setdefault(key, []).append(i)
    for members in groups.values():
        frames = [(harvested[i][1], harvested[i][2], harvested[i][0]) for i in members]
        outs = extract_frames(frames, motes_args, seeds=[seeds[i] for i in members] if motes_args.subtract_sky else None,
                              handle=handle)
        for i, products in zip(members, outs):
            name = files_and_regions[i][-1][:-5]
            results[name] = products
            if motes_args.save:
                _save(name, harvested[i], products, motes_args)
    logger.info("MOTES Processing completed.")
    return results
