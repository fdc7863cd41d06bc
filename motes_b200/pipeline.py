"""Batched whole-path API: every stage of ``core.motes`` (core.py:80-275) for many frames in one call.

``run_frames`` is the entry point bench.py times and the one a multi-GPU driver shards over
(frames are independent; see ``shard_range``).  ``run_frame`` wraps it for one frame and returns the
same per-stage dictionary as the oracle's ``run_frame`` so the parity tests can compare key by key.
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


def make_params(args, axes_dict=None) -> _lib.MotesParams:
    p = _lib.MotesParams()
    p.minimum_column_limit = float(args.minimum_column_limit)
    p.extraction_snr_limit = float(args.extraction_snr_limit)
    p.sky_snr_limit = float(args.sky_snr_limit)
    p.sky_fwhm_multiplier = float(args.sky_fwhm_multiplier)
    p.subtract_sky = int(bool(args.subtract_sky))
    p.sky_order = int(args.sky_order)
    p.spatial_floor = int(axes_dict["data_spatial_floor"]) if axes_dict else 0
    p.wavelength_start = int(axes_dict["wavelength_start"]) if axes_dict else 0
    return p


def seed_states(seeds) -> np.ndarray:
    """numpy legacy RandomState of ``np.random.seed(s)`` for every s: array (n, 625) uint32."""
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    states = np.empty((len(seeds), 625), dtype=np.uint32)
    _lib.check(_lib.lib.motes_seed_states(_lib.ptr(seeds, _lib.c_u32p), len(seeds), _lib.ptr(states, _lib.c_u32p)))
    return states


def shard_range(nframes: int, rank: int, world: int):
    """Contiguous frame range of ``rank`` (frames are independent end to end)."""
    per, extra = divmod(nframes, world)
    start = rank * per + min(rank, extra)
    return start, start + per + (1 if rank < extra else 0)


OUTPUT_SHAPES = {
    "spectra": lambda F, ns, nd, cap: ((F, 4, nd), np.float64),
    "ext_limits": lambda F, ns, nd, cap: ((F, 2, nd), np.float64),
    "sky_limits": lambda F, ns, nd, cap: ((F, 2, nd), np.float64),
    "ext_params": lambda F, ns, nd, cap: ((F, cap, 8), np.float64),
    "sky_params": lambda F, ns, nd, cap: ((F, cap, 8), np.float64),
    "ext_bins": lambda F, ns, nd, cap: ((F, cap, 2), np.int32),
    "sky_bins": lambda F, ns, nd, cap: ((F, cap, 2), np.int32),
    "ext_bin_snr": lambda F, ns, nd, cap: ((F, cap), np.float64),
    "sky_bin_snr": lambda F, ns, nd, cap: ((F, cap), np.float64),
    "n_ext_bins": lambda F, ns, nd, cap: ((F,), np.int32),
    "n_sky_bins": lambda F, ns, nd, cap: ((F,), np.int32),
    "median_params": lambda F, ns, nd, cap: ((F, 6), np.float64),
    "scale": lambda F, ns, nd, cap: ((F,), np.float64),
    "seeing_px": lambda F, ns, nd, cap: ((F,), np.float64),
    "sky_model": lambda F, ns, nd, cap: ((F, nd), np.float64),
    "sky_uncertainty": lambda F, ns, nd, cap: ((F, nd), np.float64),
    "sky_flag": lambda F, ns, nd, cap: ((F, nd), np.int32),
    "frame_status": lambda F, ns, nd, cap: ((F,), np.int32),
}
SLIM_OUTPUTS = ("spectra", "ext_limits", "n_ext_bins", "frame_status")


def bin_capacity(ndisp, minimum_column_limit) -> int:
    return int(_lib.lib.motes_bin_capacity(int(ndisp), float(minimum_column_limit)))


def run_frames(data, errs, qual, args, seeing, pixel_resolution, seeds=None, rng_states=None, axes_dict=None,
               outputs=None, copy_back=False, handle=None):
    """Run the whole path on ``F`` frames held in host arrays of shape ``(F, nspat, ndisp)``.

    qual may be float64 {0,1} (as harvested) or uint8.  ``seeds`` (one uint32 per frame) stands for
    ``np.random.seed(seed)`` before each frame; ``rng_states`` (F, 625) passes explicit generator states and
    is advanced in place.  ``outputs`` selects which result arrays are produced (default: all).  With
    ``copy_back`` the sky-subtracted frames are written back into ``data`` / ``errs`` (must then be
    C-contiguous float64).  Returns a dict of numpy arrays; per-frame failures are reported in
    ``frame_status`` (0, or the code of the exception the reference would have raised) and do not stop
    the other frames.
    """
    h = handle or _lib.default_handle()
    # float32 planes (what harvest_gmos hands over, big-endian, io/gmosio.py:40-52) travel as they are and are widened on
    # the device; everything else is converted to little-endian float64 here
    as_f32 = (not copy_back and np.asarray(data).dtype.kind == "f" and np.asarray(data).dtype.itemsize == 4
              and np.asarray(errs).dtype == np.asarray(data).dtype)
    if as_f32:
        d, e = np.ascontiguousarray(data), np.ascontiguousarray(errs)
    else:
        d, e = _lib.f64(data), _lib.f64(errs)
    if d.ndim == 2:
        d, e = d[None], e[None]
        qual = np.asarray(qual)[None]
    F, nspat, ndisp = d.shape
    q = np.asarray(qual)
    q = np.ascontiguousarray(q if q.dtype == np.uint8 else (q != 0), dtype=np.uint8).reshape(F, nspat, ndisp)
    if copy_back and (d is not data or e is not errs) and np.ndim(data) == 3:
        raise ValueError("copy_back needs C-contiguous float64 data/errs")
    p = make_params(args, axes_dict)
    cap = bin_capacity(ndisp, p.minimum_column_limit)
    see = _lib.f64(np.broadcast_to(np.asarray(seeing, dtype=np.float64), (F,)))
    pix = _lib.f64(np.broadcast_to(np.asarray(pixel_resolution, dtype=np.float64), (F,)))
    if rng_states is None and p.subtract_sky:
        if seeds is None:
            seeds = np.random.randint(0, 2 ** 32 - 1, size=F, dtype=np.uint64)
        rng_states = seed_states(np.broadcast_to(np.asarray(seeds, dtype=np.uint64), (F,)).astype(np.uint32))
    names = list(OUTPUT_SHAPES) if outputs is None else list(outputs)
    out_struct = _lib.MotesOutputs()
    result = {}
    for name in names:
        shape, dtype = OUTPUT_SHAPES[name](F, nspat, ndisp, cap)
        if name in ("sky_model", "sky_uncertainty") and p.sky_order > 0:
            shape = (F, nspat, ndisp)
        arr = np.zeros(shape, dtype=dtype)
        result[name] = arr
        setattr(out_struct, name, arr.ctypes.data)
    rng_ptr = rng_states.ctypes.data if rng_states is not None else None
    with h.lock:
        if as_f32:
            big_endian = int(d.dtype.byteorder == ">" or (d.dtype.byteorder == "=" and np.little_endian is False))
            _lib.check(_lib.lib.motes_run_frames_host_f32(
                h.raw, d.ctypes.data, e.ctypes.data, q.ctypes.data, F, nspat, ndisp, cap, ctypes.byref(p),
                see.ctypes.data, pix.ctypes.data, rng_ptr, big_endian, 1, ctypes.byref(out_struct)))
        else:
            _lib.check(_lib.lib.motes_run_frames_host(
                h.raw, d.ctypes.data, e.ctypes.data, q.ctypes.data, F, nspat, ndisp, cap, ctypes.byref(p),
                see.ctypes.data, pix.ctypes.data, rng_ptr, int(bool(copy_back)), ctypes.byref(out_struct)))
    result["cap"] = cap
    if rng_states is not None:
        result["rng_states"] = rng_states
    if copy_back:
        result["data"], result["errs"] = d, e
    return result


def raise_for_status(status: int):
    """Raise what the reference raises for a frame that ended with ``status``."""
    if status == _lib.E_NOSKY:
        raise RuntimeError("No pixels contained inside sky region.")
    if status == _lib.E_FIT:
        raise ValueError("A Moffat fit could not start (x0 outside bounds or non-finite residuals).")
    if status != 0:
        raise RuntimeError(f"motes_b200 frame status {status}")


def run_frame(frame_dict, axes_dict, header_parameters, args, seed=None):
    """One frame through ``run_frames``; returns the oracle-style dictionary of stage outputs and
    updates ``frame_dict`` / ``header_parameters`` the way ``core.motes`` leaves them."""
    data = np.ascontiguousarray(frame_dict["data"], dtype=np.float64).copy()
    errs = np.ascontiguousarray(frame_dict["errs"], dtype=np.float64).copy()
    res = run_frames(data[None], errs[None], np.asarray(frame_dict["qual"])[None], args,
                     header_parameters["seeing"], header_parameters["pixel_resolution"],
                     seeds=None if seed is None else [seed], axes_dict=axes_dict, copy_back=True)
    raise_for_status(int(res["frame_status"][0]))
    ne, ns = int(res["n_ext_bins"][0]), int(res["n_sky_bins"][0])
    out = {
        "median_params": res["median_params"][0], "scale": float(res["scale"][0]),
        "seeing_px": float(res["seeing_px"][0]),
        "ext_bins": np.column_stack([res["ext_bins"][0, :ne].astype(np.float64), res["ext_bin_snr"][0, :ne]]),
        "ext_params": res["ext_params"][0, :ne], "ext_limits": res["ext_limits"][0],
        "optimal": res["spectra"][0, 0], "optimal_err": res["spectra"][0, 1],
        "aperture": res["spectra"][0, 2], "aperture_err": res["spectra"][0, 3],
        "data_after": res["data"][0], "errs_after": res["errs"][0],
    }
    if args.subtract_sky:
        modelled = res["sky_flag"][0] == 0
        out.update({
            "sky_bins": np.column_stack([res["sky_bins"][0, :ns].astype(np.float64), res["sky_bin_snr"][0, :ns]]),
            "sky_params": res["sky_params"][0, :ns], "sky_limits": res["sky_limits"][0],
            "sky_model": res["sky_model"][0][..., modelled], "sky_uncertainty": res["sky_uncertainty"][0][..., modelled],
            "sky_flag": res["sky_flag"][0],
        })
    header_parameters["seeing"] = out["seeing_px"]
    frame_dict["data"], frame_dict["errs"] = out["data_after"], out["errs_after"]
    return out
