"""``motes.io.motesio`` with the same functions, arguments and products
(/root/reference/src/motes/io/motesio.py): ``read_regions``, ``data_harvest``, ``save_fits`` and the two table
builders.  The FITS layer is astropy when it is installed, else ``motes_b200.io.fits``
(``MOTES_B200_FITS=shim`` forces the latter).  Deliberate difference: ``read_regions`` keeps the order of reg.txt
(the reference iterates a ``set`` of lines, so its file order changes from run to run).
"""

from __future__ import annotations

import copy
import datetime
import logging
import os

import numpy as np

from . import gmosio, instruments

if os.environ.get("MOTES_B200_FITS", "") == "shim":
    from . import fits
    from .fits import Table
else:
    try:
        import astropy.io.fits as fits
        from astropy.table import Table
    except ImportError:
        from . import fits
        from .fits import Table

logger = logging.getLogger("motes")

# INSTRUME value -> (harvest, save hook), the two dispatch tables of the reference (motesio.py:52-59, 509-516).
INSTRUMENTS = {
    "en06": (instruments.harvest_floyds, instruments.save_passthrough),
    "en12": (instruments.harvest_floyds, instruments.save_passthrough),
    "FORS2": (instruments.harvest_fors2, instruments.save_passthrough),
    "GMOS-N": (gmosio.harvest_gmos, gmosio.save_gmos),
    "GMOS-S": (gmosio.harvest_gmos, gmosio.save_gmos),
    "XSHOOTER": (instruments.harvest_xshoo, instruments.save_passthrough),
}

BIN_COLUMNS = ("amplitude", "center", "alpha", "beta", "background_level", "background_grad", "bin_limit_lo", "bin_limit_hi")
BIN_DTYPES = ("f8",) * 6 + ("i4", "i4")


def read_regions(cwd):
    """``./inputs/reg.txt``: one line ``rows_off_bottom,rows_off_top,wavelength_lo,wavelength_hi,file.fits`` per
    input file -> ``[[int, int, int, int, "file.fits"], ...]`` (motesio.py:257-325, same checks and exceptions)."""
    reg_path = "./inputs/reg.txt"
    if not os.path.isfile(reg_path):
        logger.critical("reg.txt region file not found at " + cwd + "/inputs/")
        logger.critical("reg.txt is needed to define which part(s) of each input 2D spectrum file MOTES will process.")
        logger.critical("Raising FileNotFoundError.")
        raise FileNotFoundError
    logger.info("./inputs/reg.txt file found.")
    with open(reg_path, "r", encoding="utf-8") as region_file:
        lines = [line.strip() for line in region_file.read().splitlines()]
    regions, seen = [], set()
    for number, line in enumerate(lines):
        if line in seen:
            logger.warning("Ignoring duplicate region %s on line %s of reg.txt", line, number)
            continue
        seen.add(line)
        fields = line.split(",")
        if line[-5:] != ".fits":
            logger.critical("File name string on line %s of reg.txt is not '.fits'.", number + 1)
            logger.critical(line)
            logger.critical("Raising RuntimeError.")
            raise RuntimeError
        if len(fields) != 5:
            logger.critical("Number of comma-separated variables on line %s of reg.txt is not five.", number + 1)
            logger.critical(line)
            logger.critical("Raising RuntimeError.")
            raise RuntimeError
        if any(len(f) == 0 for f in fields):
            logger.critical("Variable(s) on line %s of reg.txt have zero length.", number + 1)
            logger.critical(line)
            logger.critical("Raising RuntimeError.")
            raise RuntimeError
        try:
            limits = [int(f) for f in fields[:4]]
        except ValueError:
            logger.critical("Non-numeric character found in numeric region variables on line %s of reg.txt.", number + 1)
            logger.critical(line)
            logger.critical("Raising ValueError.")
            raise ValueError from None
        regions.append(limits + [fields[4]])
    logger.info("%s regions read from ./inputs/reg.txt. Basic format checks passed.", len(regions))
    return regions


def data_harvest(input_file_path, data_regions):
    """Open the file, call the instrument's harvester, cut the frames to the region of reg.txt and build the three
    dictionaries the path works on (motesio.py:24-150).  Returns ``header_dict, frame_dict, axes_dict, original_hdu_list``."""
    with fits.open(input_file_path) as hdul:
        instrument = hdul[0].header["INSTRUME"]
        original_hdu_list = copy.deepcopy(hdul)
        data, errs, qual, header_dict, wavelength_axis = INSTRUMENTS[instrument][0](hdul)
    for key, val in header_dict.items():
        logger.info("header_dict[%s] = %s", key, val)

    spatial_floor = int(data_regions[0])
    spatial_ceiling = int(np.shape(data)[0] - data_regions[1])
    rows = slice(spatial_floor, spatial_ceiling + 1)

    if data_regions[2] < wavelength_axis[0] or data_regions[3] > wavelength_axis[-1]:
        logger.critical("Wavelength limit(s) in 'reg.txt' are beyond the wavelength range of the input data.")
        logger.critical("Raising ValueError")
        raise ValueError
    columns = np.where((wavelength_axis >= data_regions[2]) & (wavelength_axis <= data_regions[3]))[0]
    wavelength_axis = wavelength_axis[columns]
    frame_dict = {name: np.squeeze(plane[rows][:, columns]) for name, plane in (("data", data), ("errs", errs), ("qual", qual))}

    nspat = frame_dict["data"].shape[0]
    spatial_axis = np.linspace(0.0, float(nspat - 1), num=nspat)
    axes_dict = {
        "spataxislen": nspat,
        "spatial_axis": spatial_axis,
        "hi_resolution_spatial_axis": np.linspace(spatial_axis[0], spatial_axis[-1], num=nspat * 5),
        "data_spatial_floor": spatial_floor,
        "data_spatial_ceiling": spatial_ceiling,
        "dispersion_axis_length": len(wavelength_axis),
        "wavelength_axis": wavelength_axis,
        "wavelength_start": columns[0],
        "wavelength_end": columns[-1],
    }
    logger.info("2D data frame sliced to [%s:%s,%s:%s]", columns[0], columns[-1] + 1, spatial_floor, spatial_ceiling + 1)
    logger.info("Returned wavelength range is %s-%s %s.", wavelength_axis[0], wavelength_axis[-1], header_dict["wavelength_unit"])
    return header_dict, frame_dict, axes_dict, original_hdu_list


def datasec_string(floor, ceiling):
    return "[" + str(floor) + ":" + str(ceiling) + "]"


def make_bin_table(bin_data, flux_unit, sky=False):
    """(nbins, 8) Moffat / bin-edge array -> XBMPARAM (or SBMPARAM) table HDU; like the reference it makes the upper
    bin edge inclusive in the caller's array (motesio.py:173-214)."""
    bin_data[:, 7] -= 1
    hdu = fits.BinTableHDU(Table(bin_data, names=BIN_COLUMNS, dtype=BIN_DTYPES), name="SBMPARAM" if sky else "XBMPARAM")
    units = (flux_unit, "spatial_pixel", None, None, flux_unit + " / spatial_pixel", flux_unit, "spectral_pixel", "spectral_pixel")
    for i, unit in enumerate(units, 1):
        hdu.header[f"TABUNIT{i}"] = unit
    return hdu


def make_ext_table(extraction_data, wavelength_unit, sky=False):
    """(ndisp, 3) [wavelength, lower, upper] -> XTLMTAB (or SKYLMTAB) table HDU (motesio.py:217-254)."""
    table = Table(extraction_data, names=("wavelength", "ext_lim_lo", "ext_lim_hi"), dtype=("f8", "f8", "f8"))
    hdu = fits.BinTableHDU(table, name="SKYLMTAB" if sky else "XTLMTAB")
    for i, unit in enumerate((wavelength_unit, "pixel", "pixel"), 1):
        hdu.header[f"TABUNIT{i}"] = unit
    return hdu


def output_file_name(input_file_name, axes_dict):
    lo = int(np.floor(axes_dict["wavelength_axis"][0]))
    hi = int(np.ceil(axes_dict["wavelength_axis"][-1]))
    return f"m{input_file_name}_{lo}:{hi}_{axes_dict['data_spatial_floor']}:{axes_dict['data_spatial_ceiling']}.fits"


def save_fits(original_hdu_list, axes_dict, header_parameters, extracted, motes_parameters, input_file_name,
              moffat_parameters, frame_dict, moffat_parameters_all_bins, extraction_limits,
              moffat_parameters_all_sky_bins, sky_extraction_limits):
    """Write the products of one frame (motesio.py:328-528): OPTI_1D_SPEC (primary: wavelength, flux, error),
    APER_1D_SPEC, XBMPARAM, XTLMTAB, with sky subtraction also 2D_SKY, 2D_SKY_UNC, SBMPARAM, SKYLMTAB, then the
    instrument's original planes.  Same keywords, comments, roundings and file name as the reference."""
    floor, ceiling = axes_dict["data_spatial_floor"], axes_dict["data_spatial_ceiling"]
    wave_axis = axes_dict["wavelength_axis"]
    wave_unit, flux_unit = header_parameters["wavelength_unit"], header_parameters["flux_unit"]
    save_file_name = output_file_name(input_file_name, axes_dict)
    logger.info("Saving extracted spectrum and metadata to %s", save_file_name)

    cards = [
        ("MOTES", "motes.py", "Extraction script"),
        ("MOTESV", "v1.0.0", "MOTES Version"),
        ("MOTESDOI", "UNKNOWN", "MOTES DOI"),
        ("UTXTIME", datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%M:%S"), "UT timestamp for MOTES"),
        ("SPATPIXL", floor, "lower limit of spatial axis, pix"),
        ("SPATPIXH", ceiling, "upper limit of spatial axis, pix"),
        ("DISPPIXL", axes_dict["wavelength_start"], "lower limit of dispersion axis, pix"),
        ("DISPPIXH", axes_dict["wavelength_end"], "upper limit of dispersion axis, pix"),
        ("WAVL", np.floor(wave_axis[0]), "lower limit of wav range, " + wave_unit),
        ("WAVH", np.ceil(wave_axis[-1]), "upper limit of wav range, " + wave_unit),
        ("WAVU", wave_unit, "wavelength unit"),
        ("MOFFA", round(moffat_parameters[0], 5), "median Moffat profile amplitude"),
        ("MOFFC", round(moffat_parameters[1] + floor, 5), "median Moffat profile center"),
        ("MOFFALPH", round(moffat_parameters[2], 5), "median Moffat profile alpha value"),
        ("MOFFBETA", round(moffat_parameters[3], 5), "median Moffat profile beta value"),
        ("MOFFBGLV", round(moffat_parameters[4], 5), "median Moffat profile background level"),
        ("MOFFBGSL", round(moffat_parameters[5], 5), "median Moffat profile background slope"),
        ("PLATESCL", round(header_parameters["pixel_resolution"], 2), "spatial pixel plate scale, arcsec/pix"),
        ("PIXELIQ", round(header_parameters["seeing"], 2), "IQ measured from median profile, pix"),
        ("ARCSECIQ", round(header_parameters["seeing"] * header_parameters["pixel_resolution"], 2),
         "IQ measured from median profile, arcsec"),
        ("SNRBNLIM", motes_parameters.extraction_snr_limit, "maximum SNR per bin"),
        ("COLBNLIM", int(motes_parameters.minimum_column_limit), "minimum number of columns per bin"),
        ("FWHMMULT", motes_parameters.extraction_fwhm_multiplier, "FWHM used to define the extraction limits"),
    ]
    if motes_parameters.subtract_sky:
        cards += [
            ("SFWHMMLT", motes_parameters.sky_fwhm_multiplier, "FWHM multiplier to define sky region"),
            ("SSNRBNLM", motes_parameters.sky_snr_limit, "max SNR per bin for sky subtraction"),
            ("SKYORDER", motes_parameters.sky_order, "polynomial order of spatial sky model"),
        ]
    cards += [
        ("HDUROW0", "Wavelength Axis, " + wave_unit, ""),
        ("HDUROW1", "Spectrum data, " + flux_unit, ""),
        ("HDUROW2", "Spectrum uncertainty, " + flux_unit, ""),
        ("EXTNAME", "OPTI_1D_SPEC", ""),
    ]
    primary_header = original_hdu_list[0].header
    for key, value, comment in cards:
        primary_header[key] = (value, comment)

    optimal_hdu = fits.PrimaryHDU([wave_axis, extracted[0], extracted[1]], header=primary_header)
    aperture_hdu = fits.ImageHDU([wave_axis, extracted[2], extracted[3]], header=primary_header)
    aperture_hdu.header["EXTNAME"] = "APER_1D_SPEC"
    hdu_list = [optimal_hdu, aperture_hdu, make_bin_table(moffat_parameters_all_bins, flux_unit),
                make_ext_table(np.vstack([wave_axis, extraction_limits]).T, wave_unit)]

    if motes_parameters.subtract_sky:
        sky_cards = [
            ("DATATYPE", "Model Intensity", "Type of data"),
            ("DATASECS", datasec_string(floor, ceiling), "Data section; spatial axis, pixels"),
            ("DATASECW", datasec_string(axes_dict["wavelength_start"], axes_dict["wavelength_end"]),
             "Data section; wavelength axis, pixels"),
            ("BUNIT", flux_unit, ""),
            ("SKYORDER", int(motes_parameters.sky_order), "MOTES sky characterization order."),
            ("COMMENT", "2D_SKY data has same wavelength orientation as OPTI_1D_SPEC.", ""),
        ]
        sky_hdus = [fits.ImageHDU(frame_dict["sky_model"]), fits.ImageHDU(frame_dict["sky_uncertainty"])]
        for hdu, name in zip(sky_hdus, ("2D_SKY", "2D_SKY_UNC")):
            for key, value, comment in sky_cards:
                hdu.header[key] = (value, comment)
            hdu.header["EXTNAME"] = (name, "")
        hdu_list += sky_hdus
        hdu_list.append(make_bin_table(moffat_parameters_all_sky_bins, flux_unit, sky=True))
        hdu_list.append(make_ext_table(np.vstack([wave_axis, sky_extraction_limits]).T, wave_unit, sky=True))

    hdu_list = INSTRUMENTS[primary_header["INSTRUME"]][1](hdu_list, original_hdu_list)
    out = fits.HDUList(hdu_list)
    out.writeto(save_file_name)
    out.close()
    logger.info("Data saved.")
    return None
