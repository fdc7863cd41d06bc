"""CPU check of the CUDA solver's control flow: the solver headers compiled by g++ (one-lane context) vs the reference's fits."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import load_golden, ROOT
from parity import param_errors, assert_within_reference_noise

HOSTSIM_DIR = os.path.join(ROOT, "tests", "hostsim")


@pytest.fixture(scope="module")
def hostsim():
    subprocess.check_call(["make", "-s", "-C", HOSTSIM_DIR])
    return ctypes.CDLL(os.path.join(HOSTSIM_DIR, "libmotes_hostsim.so"))


def _fit(lib, profiles, alpha0, entry="hostsim_fit_batch"):
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    profiles = np.ascontiguousarray(profiles, dtype=np.float64)
    nfit, m = profiles.shape
    a0 = np.full(nfit, alpha0)
    params, cost = np.zeros((nfit, 6)), np.zeros(nfit)
    nfev, njev, status = (np.zeros(nfit, dtype=np.int32) for _ in range(3))
    getattr(lib, entry)(profiles.ctypes.data_as(dp), nfit, m, a0.ctypes.data_as(dp), params.ctypes.data_as(dp),
                          cost.ctypes.data_as(dp), nfev.ctypes.data_as(ip), njev.ctypes.data_as(ip),
                          status.ctypes.data_as(ip))
    return params, cost, nfev, njev, status


@pytest.mark.parametrize("name", ["t1_sky0", "t2_cr_sky0", "t1_faint", "t2_unbinned", "c1_sky0"])
def test_solver_tracks_reference_fits(hostsim, name):
    meta, g = load_golden(name)
    scale = g["scale"]
    profiles = g["ext_profiles"] * scale
    alpha0 = g["seeing_px"] / meta["spec"]["pixel_resolution"]
    params, cost, nfev, njev, status = _fit(hostsim, profiles, alpha0)
    ref = g["ext_params"][:, :6].copy()
    ref[:, [0, 4, 5]] *= scale
    err = param_errors(params, ref, profiles.shape[1])
    assert np.all(status >= 1) and np.all(status <= 4)
    assert_within_reference_noise(err, g, ref, scale, profiles.shape[1])
    # cost reached is the reference's cost
    assert np.allclose(cost, g["ext_fitinfo"][:, 3], rtol=1e-9)
    # iteration counts agree for most fits (they may differ by one at the ftol/xtol edge)
    assert np.mean(np.abs(nfev - g["ext_fitinfo"][:, 0]) <= 1) > 0.8


@pytest.mark.parametrize("name", ["t1_sky0", "t1_faint", "c1_sky0"])
def test_phased_solver_tracks_reference_fits(hostsim, name):
    """The phase-split form the GPU group kernel runs (forward differences from the base pow instead of
    a second pow) against the reference's fits, and against the one-loop form."""
    meta, g = load_golden(name)
    scale = g["scale"]
    profiles = g["ext_profiles"] * scale
    alpha0 = g["seeing_px"] / meta["spec"]["pixel_resolution"]
    a = _fit(hostsim, profiles, alpha0)
    b = _fit(hostsim, profiles, alpha0, entry="hostsim_fit_batch_phased")
    ref = g["ext_params"][:, :6].copy()
    ref[:, [0, 4, 5]] *= scale
    err = param_errors(b[0], ref, profiles.shape[1])
    assert np.all(b[4] >= 1) and np.all(b[4] <= 4)
    assert_within_reference_noise(err, g, ref, scale, profiles.shape[1])
    assert np.allclose(b[1], g["ext_fitinfo"][:, 3], rtol=1e-9)
    assert np.mean(np.abs(b[2] - g["ext_fitinfo"][:, 0]) <= 1) > 0.8
    assert np.mean(np.abs(b[2] - a[2]) <= 1) > 0.8


def test_solver_error_codes(hostsim):
    m = 64
    flat = np.linspace(-5.0, -1.0, m)[None, :]          # negative top-3 median -> x0 below the lower bound
    _, _, _, _, status = _fit(hostsim, flat, 4.0)
    assert status[0] == -1
    nanny = np.ones((1, m))
    nanny[0, 10] = np.nan
    _, _, _, _, status = _fit(hostsim, nanny, 4.0)
    assert status[0] in (-1, -2)
