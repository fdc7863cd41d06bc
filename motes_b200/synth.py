"""Bit-reproducible synthetic long-slit spectrograms.

Every arithmetic step is +, -, *, / or sqrt on float64 (all correctly rounded by
IEEE-754), the Moffat exponent is a half-integer so ``(1+u)**-beta`` is built from
multiplications and one sqrt, and the noise is an Irwin-Hall (sum of 12 uniforms)
deviate drawn from numpy's PCG64 ``Generator.random`` (integer arithmetic only).
The same seed therefore yields the same bytes on any host, which lets
``tests/golden`` store only the seed and a digest of the inputs.

Frames follow the harvester contract of the reference
(``/root/reference/src/motes/io/motesio.py:132-148``): ``data``, ``errs``, ``qual``
are C-contiguous ``(nspat, ndisp)`` float64 arrays, dispersion fastest; chip gaps
hold the literal value 1.0 in ``data`` and ``errs`` and 0 in ``qual``
(``/root/reference/src/motes/io/gmosio.py:48-52``).
"""

from __future__ import annotations

import hashlib
from dataclasses import dataclass, field, asdict
from types import SimpleNamespace

import numpy as np


@dataclass
class FrameSpec:
    """Shape and content of one synthetic 2D spectrogram."""

    nspat: int = 200
    ndisp: int = 3000
    seed: int = 1
    beta2: int = 6            # 2*beta of the Moffat trace (5..8 -> beta 2.5..4)
    alpha: float = 3.2        # Moffat alpha in pixels
    amplitude: float = 900.0  # peak counts at the blue end (x2 towards the red end)
    center_offset: float = 0.37
    curvature: float = 6.0    # peak-to-centre wander of the trace in pixels
    sky_level: float = 120.0
    sky_ripple: float = 0.3   # fractional sawtooth modulation of the sky along dispersion
    sky_line: float = 400.0   # extra counts on sparse "sky line" columns
    sky_slope: float = 0.0    # fractional spatial gradient of the sky across the slit
    read_noise: float = 5.0
    chip_gaps: tuple = ()     # ((first_column, width), ...)
    cosmic_fraction: float = 0.0  # fraction of pixels hit by a cosmic ray
    nan_fraction: float = 0.0     # fraction of pixels set to NaN with qual = 0
    seeing: float = 0.8       # header seeing [arcsec] handed to the first fit
    pixel_resolution: float = 0.16


# The five BASELINE.json configurations.
def config_spec(name: str, seed: int = 1, **over) -> FrameSpec:
    base = {
        "C1": dict(nspat=200, ndisp=3000, chip_gaps=((980, 40), (2010, 40)),
                   seeing=0.75, pixel_resolution=0.16),
        "C2": dict(nspat=400, ndisp=4096, alpha=3.6, beta2=6, curvature=9.0),
        "C3": dict(nspat=80, ndisp=24000, alpha=2.4, beta2=7, curvature=3.0,
                   cosmic_fraction=0.01, amplitude=1500.0),
        "C5a": dict(nspat=256, ndisp=2048, alpha=3.0, amplitude=60.0),
        "C5b": dict(nspat=512, ndisp=16384, alpha=3.0, amplitude=60.0),
        # small shapes used by fast tests and committed golden vectors
        "T1": dict(nspat=96, ndisp=640, alpha=2.6, curvature=4.0, chip_gaps=((200, 24),)),
        "T2": dict(nspat=64, ndisp=400, alpha=2.2, beta2=7, curvature=2.0,
                   cosmic_fraction=0.01),
    }[name]
    base.update(over)
    return FrameSpec(seed=seed, **base)


def _unit_moffat(u: np.ndarray, beta2: int) -> np.ndarray:
    """(1+u)**(-beta2/2) from multiplications and a sqrt only."""
    b = 1.0 + u
    whole = beta2 // 2
    p = np.ones_like(b)
    for _ in range(whole):
        p = p * b
    if beta2 % 2:
        p = p * np.sqrt(b)
    return 1.0 / p


def irwin_hall(rng: np.random.Generator, shape) -> np.ndarray:
    """Zero-mean unit-variance deviate: sum of 12 U(0,1) minus 6 (additions only)."""
    acc = np.zeros(shape)
    for _ in range(12):
        acc += rng.random(shape)
    return acc - 6.0


def make_frame(spec: FrameSpec):
    """Return ``(frame_dict, axes_dict, header_parameters)`` for ``spec``.

    The three dicts carry the keys the reference's ``data_harvest`` produces
    (``/root/reference/src/motes/io/motesio.py:132-148``,
    ``/root/reference/src/motes/io/gmosio.py:118-126``).
    """
    ns, nd = spec.nspat, spec.ndisp
    rng = np.random.default_rng(spec.seed)
    r = np.arange(ns, dtype=np.float64)[:, None]
    x = np.arange(nd, dtype=np.float64)[None, :]
    t = x / float(nd)                       # 0..1 along dispersion
    tc = t - 0.5
    center = ns * 0.5 + spec.center_offset + spec.curvature * (4.0 * tc * tc - 0.5)
    amp = spec.amplitude * (1.0 + t)
    d = r - center
    u = (d * d) / (spec.alpha * spec.alpha)
    trace = amp * _unit_moffat(u, spec.beta2)

    saw = (np.arange(nd) % 257).astype(np.float64) / 257.0
    sky = spec.sky_level * (1.0 + spec.sky_ripple * saw)
    line = ((np.arange(nd) * 7919) % 97) < 3
    sky = sky + spec.sky_line * line.astype(np.float64)
    sky2d = sky[None, :] * (1.0 + spec.sky_slope * (r / float(ns)))

    signal = trace + sky2d
    sigma = np.sqrt(signal + spec.read_noise * spec.read_noise)
    data = signal + sigma * irwin_hall(rng, (ns, nd))
    errs = sigma.copy()
    qual = np.ones((ns, nd))

    if spec.cosmic_fraction > 0.0:
        hit = rng.random((ns, nd)) < spec.cosmic_fraction
        boost = 500.0 + 4500.0 * rng.random((ns, nd))
        data = np.where(hit, data + boost, data)
    if spec.nan_fraction > 0.0:
        bad = rng.random((ns, nd)) < spec.nan_fraction
        data = np.where(bad, np.nan, data)
        qual = np.where(bad, 0.0, qual)
    for first, width in spec.chip_gaps:
        data[:, first:first + width] = 1.0
        errs[:, first:first + width] = 1.0
        qual[:, first:first + width] = 0.0

    frame_dict = {"data": np.ascontiguousarray(data),
                  "errs": np.ascontiguousarray(errs),
                  "qual": np.ascontiguousarray(qual)}
    axes_dict = make_axes(ns, nd)
    header_parameters = {
        "object": "synthetic", "exptime": 1.0, "seeing": float(spec.seeing),
        "instrument": "SYNTH", "flux_unit": "count", "wavelength_unit": "pixel",
        "pixel_resolution": float(spec.pixel_resolution),
    }
    return frame_dict, axes_dict, header_parameters


def make_axes(nspat: int, ndisp: int, wavelength_start: int = 0, spatial_floor: int = 0) -> dict:
    """axes_dict with the keys of ``motesio.data_harvest`` (motesio.py:138-148)."""
    spatial_axis = np.linspace(0.0, float(nspat - 1), num=nspat)
    return {
        "spataxislen": nspat,
        "spatial_axis": spatial_axis,
        "hi_resolution_spatial_axis": np.linspace(spatial_axis[0], spatial_axis[-1], num=nspat * 5),
        "data_spatial_floor": int(spatial_floor),
        "data_spatial_ceiling": int(spatial_floor + nspat - 1),
        "dispersion_axis_length": int(ndisp),
        "wavelength_axis": np.arange(wavelength_start, wavelength_start + ndisp, dtype=np.float64),
        "wavelength_start": int(wavelength_start),
        "wavelength_end": int(wavelength_start + ndisp - 1),
    }


def default_args(**over) -> SimpleNamespace:
    """The argparse Namespace of ``xmotes.py`` (/root/reference/src/xmotes.py:33-138) with its defaults."""
    ns = SimpleNamespace(
        save=False, minimum_column_limit=15, extraction_snr_limit=10,
        extraction_fwhm_multiplier=2.0, subtract_sky=True, sky_order=0,
        sky_snr_limit=100, sky_fwhm_multiplier=3.0,
        diag_plot_spectrum=False, diag_plot_collapsed=False, diag_plot_moffat=False,
        diag_plot_bins=False, diag_plot_localisation=False, diag_plot_skyfit=False,
        diag_plot_skysub=False, verbose=False,
    )
    for k, v in over.items():
        setattr(ns, k, v)
    return ns


def write_gmos_fits(path, spec: FrameSpec, dtype=">f8", crval=500.0, step=0.1, raw_iq="70-percentile", flipped=False):
    """Write ``make_frame(spec)`` as a GMOS-style multi-extension FITS file (primary + MDF + SCI + VAR + DQ, the layout
    ``harvest_gmos`` reads, /root/reference/src/motes/io/gmosio.py:38-42): SCI = data, VAR = errs**2, DQ = 16 in chip
    gaps, 1 on other bad pixels.  ``flipped`` stores the planes red to blue with CD1_1 < 0 (DRAGONS).  Real GMOS
    planes are float32 (``dtype=">f4"``); float64 keeps the frame bit for bit.  Returns the frame dictionaries."""
    from .io import fits
    frame, axes, header = make_frame(spec)
    gap = np.zeros(frame["data"].shape, dtype=bool)
    for first, width in spec.chip_gaps:
        gap[:, first:first + width] = True
    dq = np.where(gap, 16, np.where(frame["qual"] == 0, 1, 0)).astype(np.int16)
    sci, var = frame["data"].astype(dtype), (frame["errs"] * frame["errs"]).astype(dtype)
    if flipped:
        sci, var, dq = (np.ascontiguousarray(np.flip(a, axis=1)) for a in (sci, var, dq))
    primary = fits.PrimaryHDU()
    for key, value in (("INSTRUME", "GMOS-S"), ("OBJECT", "synthetic target"), ("EXPTIME", 300.0), ("RAWIQ", raw_iq)):
        primary.header[key] = value
    mdf = fits.BinTableHDU(fits.Table(np.array([[1.0, 2.0]]), names=("slitpos_mx", "slitpos_my"), dtype=("f4", "f4")), name="MDF")
    planes = [fits.ImageHDU(sci, name="SCI"), fits.ImageHDU(var, name="VAR"), fits.ImageHDU(dq, name="DQ")]
    wcs = (("CRPIX1", float(spec.ndisp) if flipped else 1.0), ("CRVAL1", float(crval)), ("CD1_1", -float(step) if flipped else float(step)),
           ("CUNIT1", "nm"), ("BUNIT", "electron"), ("PIXSCALE", float(spec.pixel_resolution)))
    for key, value in wcs:
        planes[0].header[key] = value
    fits.HDUList([primary, mdf] + planes).writeto(path, overwrite=True)
    return frame, axes, header


def frame_digest(frame_dict: dict) -> str:
    """sha256 over the bytes of data, errs and qual (little-endian float64)."""
    h = hashlib.sha256()
    for key in ("data", "errs", "qual"):
        h.update(np.ascontiguousarray(frame_dict[key], dtype="<f8").tobytes())
    return h.hexdigest()


def spec_to_dict(spec: FrameSpec) -> dict:
    d = asdict(spec)
    d["chip_gaps"] = [list(g) for g in spec.chip_gaps]
    return d
