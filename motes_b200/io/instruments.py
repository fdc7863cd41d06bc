"""FORS2, X-shooter and FLOYDS harvesters in the interface ``data_harvest`` calls today:
``harvest(hdul) -> data, errs, qual, header_dict, wavelength_axis`` (/root/reference/src/motes/io/motesio.py:67-70).

Upstream's ``forsio.py`` / ``xshooio.py`` / ``floydsio.py`` still have the pre-1.0 signature
``harvest(hdul, primary_header) -> (..., original_qual, ...)`` and cannot be reached through ``data_harvest``
(SURVEY.md §8f rank 4).  These functions compute the same frames, header dictionary and wavelength axis as that code
(forsio.py:9-96, xshooio.py:9-76, floydsio.py:9-75; checked side by side in ``tests/test_fits_io.py``) and drop the
``original_qual`` return the current interface no longer has.  Failures raise instead of calling ``sys.exit``.
The save hooks append nothing, as upstream's do.
"""

from __future__ import annotations

import numpy as np


def make_wav_axis(start, increment, length):
    """common.make_wav_axis (/root/reference/src/motes/common.py:40-62): ``length`` values from ``start`` in steps of
    ``increment`` (``np.arange`` with a computed stop, so rounding can add a value, as in the reference)."""
    return np.arange(start=start, step=increment, stop=(length * increment) + start)


# (spatial binning, collimator) -> arcsec per pixel (forsio.py:42-70)
FORS2_PIXEL_SCALE = {(1, "COLL_SR"): 0.125, (1, "COLL_HR"): 0.0632, (2, "COLL_SR"): 0.25, (2, "COLL_HR"): 0.125}


def _mean_ambient_fwhm(header):
    return 0.5 * (header["HIERARCH ESO TEL AMBI FWHM START"] + header["HIERARCH ESO TEL AMBI FWHM END"])


def harvest_fors2(input_fits_hdu):
    primary = input_fits_hdu[0].header
    data = input_fits_hdu[0].data
    errs = input_fits_hdu[1].data ** 0.5
    qual = np.ones(np.shape(data))                 # FORS2 products carry no quality plane: every pixel is good
    key = (primary["HIERARCH ESO DET WIN1 BINY"], primary["HIERARCH ESO INS COLL NAME"])
    if key not in FORS2_PIXEL_SCALE:
        raise ValueError(f"FORS2: no spatial pixel scale for binning / collimator {key}")
    header_dict = {
        "object": primary["OBJECT"].replace(" ", "_"),
        "pixel_resolution": FORS2_PIXEL_SCALE[key],
        "exptime": primary["HIERARCH ESO INS SHUT EXPTIME"],
        "instrument": primary["INSTRUME"],
        "seeing": _mean_ambient_fwhm(primary),
        "flux_unit": primary["BUNIT"],
        "wavelength_unit": "Angstroms",
    }
    return data, errs, qual, header_dict, make_wav_axis(primary["CRVAL1"], primary["CD1_1"], primary["NAXIS1"])


# X-shooter pipeline flags that still mean "usable pixel": interpolated, and the two b-spline sky-subtraction outlier
# flags (xshooio.py:36-43)
XSHOOTER_BENIGN_FLAGS = (4194304, 8388608, 16777216)


def harvest_xshoo(input_fits_hdu):
    primary = input_fits_hdu[0].header
    data = input_fits_hdu[0].data
    errs = input_fits_hdu[1].data                  # already a sigma plane
    flags = input_fits_hdu[2].data
    good = (flags == 0) | np.isin(flags, XSHOOTER_BENIGN_FLAGS)      # exact values only: sums of flags stay bad
    qual = flags                                   # converted in place, like the reference (keeps the plane's dtype)
    qual[...] = np.where(good & np.isfinite(data), 1, 0)
    header_dict = {
        "object": primary["OBJECT"].replace(" ", "_"),
        "pixel_resolution": primary["CDELT2"],
        "exptime": primary["EXPTIME"],
        "seeing": _mean_ambient_fwhm(primary),
        "flux_unit": primary["BUNIT"],
        "wavelength_unit": "nm",
    }
    return data, errs, qual, header_dict, make_wav_axis(primary["CRVAL1"], primary["CDELT1"], primary["NAXIS1"])


def harvest_floyds(input_fits_hdu):
    """FLOYDS frames have no variance plane: sigma = sqrt(counts + read noise^2 + dark current^2), so the frame must not
    be flux calibrated (floydsio.py:42-48)."""
    primary = input_fits_hdu[0].header
    data = input_fits_hdu[0].data
    errs = np.sqrt(data + primary["RDNOISE"] * primary["RDNOISE"] + primary["DARKCURR"] * primary["DARKCURR"])
    qual = np.ones(np.shape(data))
    header_dict = {
        "object": primary["OBJECT"].replace(" ", "_"),
        "pixel_resolution": primary["PIXSCALE"],
        "exptime": primary["EXPTIME"],
        "instrument": primary["INSTRUME"],
        "seeing": primary["AGFWHM"],               # autoguider FWHM estimate
        "flux_unit": "electrons",
        "wavelength_unit": primary["WAT2_001"].split(" ")[2].split("=")[1],
    }
    return data, errs, qual, header_dict, make_wav_axis(primary["CRVAL1"], primary["CD1_1"], primary["NAXIS1"])


def save_passthrough(hdu_list, original_hdu_list):
    """forsio.save_fors / xshooio.save_xshoo / floydsio.save_floyds: nothing of the input file is appended."""
    return hdu_list
