"""GMOS harvester and save hook with the reference's contract (/root/reference/src/motes/io/gmosio.py).

``harvest_gmos(hdul) -> data, errs, qual, header_dict, wavelength_axis`` (gmosio.py:10-128): SCI, sqrt(VAR) and DQ
planes, chip-gap pixels (DQ == 16) forced to 1.0 in data and errs, DQ turned into good = 1 / bad = 0, the linear
wavelength axis from CRPIX1 / CRVAL1 / CD1_1 (flipped for DRAGONS reductions, CD1_1 < 0) and a first seeing guess from
the Gemini image-quality percentile.  Host-side only; the arrays go to the device through ``motes_b200.pipeline``.
"""

from __future__ import annotations

import numpy as np

# Gemini image-quality bins: percentile -> column, shortest wavelength [nm] -> row, value = FWHM in arcsec
# (the table gmosio.py:79-111 takes from the Gemini observing-condition constraints).
IQ_COLUMN = {"20-percentile": 0, "70-percentile": 1, "85-percentile": 2, "100-percentile": 3, "Any": 3, "UNKNOWN": 3}
IQ_BAND_EDGES = (0.0, 400.0, 550.0, 700.0, 850.0, 975.0, 1100.0)
IQ_ARCSEC = ((0.6, 0.90, 1.20, 2.00), (0.6, 0.85, 1.10, 1.90), (0.5, 0.75, 1.05, 1.80),
             (0.5, 0.75, 1.05, 1.70), (0.5, 0.70, 0.95, 1.70), (0.4, 0.70, 0.95, 1.65))


def gmos_wavelength_axis(science_header) -> np.ndarray:
    """Linear dispersion solution evaluated at every pixel (gmosio.py:55-67)."""
    crpix = science_header["CRPIX1"]
    if crpix > 0.5:
        ref_pixel = np.ceil(crpix)
        shift = ref_pixel - crpix
    else:
        ref_pixel = np.floor(crpix)
        shift = crpix - ref_pixel
    step = science_header["CD1_1"]
    ref_wavelength = shift * step + science_header["CRVAL1"]
    pixels = np.arange(-ref_pixel + 1, science_header["NAXIS1"] - ref_pixel)
    return pixels * step + ref_wavelength


def gmos_seeing_guess(raw_iq: str, shortest_wavelength: float) -> float:
    """First FWHM estimate [arcsec] for the median Moffat fit (gmosio.py:69-116).  A wavelength that sits on a band
    edge or outside 0-1100 nm has no entry - the reference fails there with UnboundLocalError; here it is a KeyError."""
    for row, (lo, hi) in enumerate(zip(IQ_BAND_EDGES[:-1], IQ_BAND_EDGES[1:])):
        if lo < shortest_wavelength < hi:
            return float(IQ_ARCSEC[row][IQ_COLUMN[raw_iq]])
    raise KeyError(f"no Gemini IQ band holds the wavelength {shortest_wavelength}")


def harvest_gmos(input_fits_hdu):
    primary = input_fits_hdu[0].header
    science = input_fits_hdu["SCI"].header
    data = input_fits_hdu["SCI"].data
    errs = input_fits_hdu["VAR"].data ** 0.5
    dq = input_fits_hdu["DQ"].data

    in_gap = dq == 16
    data[in_gap] = 1.0
    errs[in_gap] = 1.0
    qual = np.ones(np.shape(dq))
    qual[dq > 0] = 0

    wavelength_axis = gmos_wavelength_axis(science)
    if science["CD1_1"] < 0:                       # DRAGONS stores red to blue
        wavelength_axis = np.flip(wavelength_axis)
        data, errs, qual = (np.flip(a, axis=1) for a in (data, errs, qual))

    header_dict = {
        "object": primary["OBJECT"].replace(" ", "_"),
        "exptime": primary["EXPTIME"],
        "seeing": gmos_seeing_guess(primary["RAWIQ"], np.min(wavelength_axis)),
        "instrument": primary["INSTRUME"],
        "flux_unit": science["BUNIT"],
        "wavelength_unit": science["CUNIT1"],
        "pixel_resolution": float(science["PIXSCALE"]),
    }
    return data, errs, qual, header_dict, wavelength_axis


def save_gmos(hdu_list, original_hdu_list):
    """Append the input file's MDF / SCI / VAR / DQ planes, renamed ORIG_* (gmosio.py:131-156)."""
    for name in ("MDF", "SCI", "VAR", "DQ"):
        original_hdu_list[name].header["EXTNAME"] = "ORIG_" + name
        hdu_list.append(original_hdu_list["ORIG_" + name])
    return hdu_list
