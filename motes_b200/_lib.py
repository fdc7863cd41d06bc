"""ctypes binding of libmotes_b200.so (the C ABI in include/motes_b200.h).

There is no CPU implementation behind this module: if the shared library has not
been built (``python -c 'import __graft_entry__ as g; g.build()'`` or
``make -C motes_b200/csrc``) importing it raises, and if no CUDA device is usable
every compute call raises ``MotesCudaError``.
"""

from __future__ import annotations

import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MOTES_B200_LIB") or os.path.join(_HERE, "libmotes_b200.so")   # override: A/B builds (tools/)

MOTES_OK, E_CUDA, E_ARG, E_NOMEM, E_NOSKY, E_FIT, E_INTERNAL = 0, -1, -2, -3, -4, -5, -6


class MotesCudaError(RuntimeError):
    """CUDA runtime failure inside libmotes_b200 (includes 'no device')."""


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build the CUDA extension first (make -C motes_b200/csrc). "
        "motes_b200 has no CPU fallback.")

lib = ctypes.CDLL(LIB_PATH)

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int32)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_u32p = ctypes.POINTER(ctypes.c_uint32)
c_handle = ctypes.c_void_p



class MotesParams(ctypes.Structure):
    """struct motes_params (include/motes_b200.h)."""
    _fields_ = [("minimum_column_limit", ctypes.c_double), ("extraction_snr_limit", ctypes.c_double),
                ("sky_snr_limit", ctypes.c_double), ("sky_fwhm_multiplier", ctypes.c_double),
                ("subtract_sky", ctypes.c_int32), ("sky_order", ctypes.c_int32),
                ("spatial_floor", ctypes.c_int32), ("wavelength_start", ctypes.c_int32)]


OUTPUT_FIELDS = ["spectra", "ext_limits", "sky_limits", "ext_params", "sky_params", "ext_bins", "sky_bins",
                 "ext_bin_snr", "sky_bin_snr", "n_ext_bins", "n_sky_bins", "median_params", "scale", "seeing_px",
                 "sky_model", "sky_uncertainty", "sky_flag", "frame_status"]


class MotesOutputs(ctypes.Structure):
    """struct motes_outputs (include/motes_b200.h): one void* per output, NULL = not wanted."""
    _fields_ = [(name, ctypes.c_void_p) for name in OUTPUT_FIELDS]


c_vp = ctypes.c_void_p
c_int = ctypes.c_int

# name -> (restype, argtypes); one entry per declaration in include/motes_b200.h
SIGNATURES = {
    "motes_profile_enable": (ctypes.c_int, [c_handle, c_int]),
    "motes_profile_stage_count": (ctypes.c_int, []),
    "motes_profile_stage_name": (ctypes.c_char_p, [c_int]),
    "motes_profile_read": (ctypes.c_int, [c_handle, c_dp, ctypes.POINTER(ctypes.c_longlong), c_int]),
    "motes_fp64_peak": (ctypes.c_int, [c_handle, c_dp]),
    "motes_bin_capacity": (ctypes.c_int, [c_int, ctypes.c_double]),
    "motes_seed_states": (ctypes.c_int, [c_u32p, c_int, c_u32p]),
    "motes_run_frames_host": (ctypes.c_int, [c_handle, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int,
                                             ctypes.POINTER(MotesParams), c_vp, c_vp, c_vp, c_int,
                                             ctypes.POINTER(MotesOutputs)]),
    "motes_run_frames_host_async": (ctypes.c_int, [c_handle, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int,
                                                   ctypes.POINTER(MotesParams), c_vp, c_vp, c_vp, c_int,
                                                   ctypes.POINTER(MotesOutputs)]),
    "motes_run_frames_host_f32": (ctypes.c_int, [c_handle, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int,
                                                 ctypes.POINTER(MotesParams), c_vp, c_vp, c_vp, c_int, c_int,
                                                 ctypes.POINTER(MotesOutputs)]),
    "motes_run_frames_dev": (ctypes.c_int, [c_handle, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int,
                                            ctypes.POINTER(MotesParams), c_vp, c_vp, c_vp,
                                            ctypes.POINTER(MotesOutputs)]),
    "motes_qual_pack_dev": (ctypes.c_int, [c_handle, c_vp, ctypes.c_size_t, c_vp]),
    "motes_row_median_host": (ctypes.c_int, [c_handle, c_dp, c_int, c_int, c_dp]),
    "motes_get_bins_host": (ctypes.c_int, [c_handle, c_dp, c_dp, c_u8p, c_int, c_int, c_int, c_int, ctypes.c_double,
                                           ctypes.c_double, c_int, c_ip, c_dp, c_ip, c_u8p]),
    "motes_bin_profiles_host": (ctypes.c_int, [c_handle, c_dp, c_int, c_int, c_ip, c_int, c_dp]),
    "motes_interp_limits_host": (ctypes.c_int, [c_handle, c_dp, c_int, c_int, c_dp, c_dp]),
    "motes_sky_subtract_host": (ctypes.c_int, [c_handle, c_dp, c_dp, c_int, c_int, c_dp, c_dp, c_int, c_int, c_u32p,
                                               c_dp, c_dp, c_ip, c_ip, c_ip]),
    "motes_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(c_handle)]),
    "motes_destroy": (ctypes.c_int, [c_handle]),
    "motes_last_error": (ctypes.c_char_p, []),
    "motes_version": (ctypes.c_char_p, []),
    "motes_stream": (ctypes.c_void_p, [c_handle]),
    "motes_set_stream": (ctypes.c_int, [c_handle, ctypes.c_void_p]),
    "motes_synchronize": (ctypes.c_int, [c_handle]),
    "motes_launch_count": (ctypes.c_longlong, [c_handle]),
    "motes_moffat_fit_batch_host": (ctypes.c_int, [c_handle, c_dp, ctypes.c_int, ctypes.c_int, c_dp, c_dp,
                                                   c_dp, c_ip, c_ip, c_ip]),
    "motes_moffat_fit_batch_dev": (ctypes.c_int, [c_handle, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "motes_optimal_extract_host": (ctypes.c_int, [c_handle, c_dp, c_dp, ctypes.c_int, ctypes.c_int, c_dp, c_dp,
                                                  c_dp, ctypes.c_int, ctypes.c_int, c_dp, c_ip, c_ip, c_u8p]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return (lib.motes_last_error() or b"").decode()


def check(rc: int) -> None:
    """Map a C status to the exception the reference would raise at that point."""
    if rc == MOTES_OK:
        return
    msg = last_error()
    if rc == E_CUDA:
        raise MotesCudaError(msg or "CUDA error")
    if rc == E_NOMEM:
        raise MemoryError(msg or "device allocation failed")
    if rc == E_NOSKY:
        raise RuntimeError(msg or "No pixels contained inside sky region.")
    if rc == E_FIT:
        raise ValueError(msg or "Moffat fit could not start")
    raise ValueError(msg or f"motes_b200 error {rc}")


def f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype="<f8")


def ptr(a: np.ndarray, ctype=c_dp):
    return a.ctypes.data_as(ctype) if a is not None else None


class Handle:
    """One library handle (device + stream + scratch).  Calls on a handle are serialised."""

    def __init__(self, device: int = 0):
        self._h = c_handle()
        check(lib.motes_create(int(device), ctypes.byref(self._h)))
        self.device = int(device)
        self.lock = threading.Lock()

    @property
    def raw(self):
        return self._h

    def synchronize(self):
        check(lib.motes_synchronize(self._h))

    def launch_count(self) -> int:
        return int(lib.motes_launch_count(self._h))

    def close(self):
        if self._h:
            lib.motes_destroy(self._h)
            self._h = c_handle()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default = {}
_default_lock = threading.Lock()


def default_handle(device: int | None = None) -> Handle:
    """Process-wide handle per device (MOTES_B200_DEVICE or LOCAL_RANK picks the default device)."""
    if device is None:
        device = int(os.environ.get("MOTES_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    with _default_lock:
        if device not in _default:
            _default[device] = Handle(device)
        return _default[device]
