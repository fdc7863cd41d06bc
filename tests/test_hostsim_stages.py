"""CPU checks of the non-solver device routines (compiled by g++ with the one-thread scope) against the oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import load_golden, golden_inputs, ROOT
from oracle import motes_oracle as mo

HOSTSIM_DIR = os.path.join(ROOT, "tests", "hostsim")
dp = ctypes.POINTER(ctypes.c_double)
ip = ctypes.POINTER(ctypes.c_int32)
bp = ctypes.POINTER(ctypes.c_uint8)


@pytest.fixture(scope="module")
def hostsim():
    subprocess.check_call(["make", "-s", "-C", HOSTSIM_DIR])
    lib = ctypes.CDLL(os.path.join(HOSTSIM_DIR, "libmotes_hostsim.so"))
    lib.hostsim_median.restype = ctypes.c_double
    return lib


def test_radix_median_matches_numpy(hostsim):
    rng = np.random.default_rng(0)
    for n in [1, 2, 3, 4, 7, 16, 33, 150, 401, 1000]:
        for kind in range(3):
            v = rng.normal(0, 10.0 ** rng.integers(-3, 4), n)
            if kind == 1:
                v = np.round(v)              # many ties, +-0
            if kind == 2:
                v[rng.integers(0, n)] = -0.0
            got = hostsim.hostsim_median(np.ascontiguousarray(v).ctypes.data_as(dp), n)
            assert got == np.median(v), (n, kind)


def _bins(hostsim, frame, row_lo, row_hi, limit, min_cols):
    data, errs = np.ascontiguousarray(frame["data"]), np.ascontiguousarray(frame["errs"])
    qual = np.ascontiguousarray(frame["qual"] != 0, dtype=np.uint8)
    nspat, ndisp = data.shape
    bins_i = np.zeros(2 * ndisp, dtype=np.int32)
    snr = np.zeros(ndisp)
    gap = np.zeros(ndisp, dtype=np.uint8)
    nb = hostsim.hostsim_snr_bins(data.ctypes.data_as(dp), errs.ctypes.data_as(dp), qual.ctypes.data_as(bp), nspat, ndisp,
                                  int(row_lo), int(row_hi), ctypes.c_double(limit), ctypes.c_double(min_cols),
                                  bins_i.ctypes.data_as(ip), snr.ctypes.data_as(dp), gap.ctypes.data_as(bp))
    return bins_i[:2 * nb].reshape(nb, 2), snr[:nb], gap


@pytest.mark.parametrize("name", ["t1_sky0", "t2_cr_sky0", "t1_faint", "t2_unbinned", "c1_sky0"])
def test_bins_are_bit_exact(hostsim, name):
    meta, g = load_golden(name)
    frame, axes, header, args = golden_inputs(meta)
    row_lo, row_hi = g["bin_rows"]
    edges, snr, gap = _bins(hostsim, frame, row_lo, row_hi, args.sky_snr_limit, args.minimum_column_limit)
    assert np.array_equal(edges, g["sky_bins"][:, :2].astype(np.int32))
    assert np.allclose(snr, g["sky_bins"][:, 2], rtol=1e-12, equal_nan=True)
    want_gap = (np.median(frame["data"], axis=0) == 1).astype(np.uint8)
    assert np.array_equal(gap, want_gap)


def test_bins_with_bad_pixels_and_nans(hostsim):
    from motes_b200 import synth
    frame, axes, header = synth.make_frame(synth.config_spec("T2", seed=21, nan_fraction=0.02))
    args = synth.default_args()
    for row_lo, row_hi in [(20, 44), (-3, 70), (0, 64)]:
        want = mo.snr_bins(frame, row_lo, row_hi, axes["dispersion_axis_length"], args, has_sky=False)
        edges, snr, _ = _bins(hostsim, frame, row_lo, row_hi, args.extraction_snr_limit, args.minimum_column_limit)
        assert np.array_equal(edges, np.array(want)[:, :2].astype(np.int32))
        assert np.allclose(snr, np.array(want)[:, 2], rtol=1e-12, equal_nan=True)


def _interp(hostsim, table, ndisp):
    table = np.ascontiguousarray(table, dtype=np.float64)
    lo, hi = np.zeros(ndisp), np.zeros(ndisp)
    hostsim.hostsim_interp_limits(table.ctypes.data_as(dp), table.shape[1], ndisp, lo.ctypes.data_as(dp), hi.ctypes.data_as(dp))
    return lo, hi


@pytest.mark.parametrize("name", ["t1_sky0", "t2_cr_sky0", "t1_faint", "t2_unbinned", "c1_sky0"])
def test_limit_interpolation_is_bit_exact(hostsim, name):
    meta, g = load_golden(name)
    ndisp = meta["spec"]["ndisp"]
    lo, hi = _interp(hostsim, g["ext_limit_table"], ndisp)
    assert np.array_equal(lo, g["ext_limits"][0])
    assert np.array_equal(hi, g["ext_limits"][1])


def test_limit_interpolation_single_bin_and_short_axis(hostsim):
    table = np.array([[100.0], [10.5], [30.25], [20.0]])
    lo, hi = _interp(hostsim, table, 200)
    assert np.all(lo == 10.5) and np.all(hi == 30.25)
    # fewer than 300 interior samples: the fixed 150-sample windows shrink (python slices)
    table = np.array([[8.0, 40.0, 120.0, 190.5], [10.0, 11.0, 12.5, 12.0], [30.0, 31.0, 33.0, 32.0], [20.0, 21.0, 22.0, 22.0]])
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = mo.interpolate_limits(table, 200)
    lo, hi = _interp(hostsim, table, 200)
    assert np.array_equal(lo, want[0], equal_nan=True) and np.array_equal(hi, want[1], equal_nan=True)


def _sky(hostsim, data, errs, lo, hi, floor, state):
    nspat, ndisp = data.shape
    u32p = ctypes.POINTER(ctypes.c_uint32)
    mt = np.ascontiguousarray(state[1], dtype=np.uint32).copy()
    pos = ctypes.c_int(int(state[2]))
    sky, sky_err = np.zeros(ndisp), np.zeros(ndisp)
    n_sky, n_good, flag = (np.zeros(ndisp, dtype=np.int32) for _ in range(3))
    rc = hostsim.hostsim_sky_median(data.ctypes.data_as(dp), errs.ctypes.data_as(dp), nspat, ndisp,
                                    lo.ctypes.data_as(dp), hi.ctypes.data_as(dp), int(floor), mt.ctypes.data_as(u32p),
                                    ctypes.byref(pos), sky.ctypes.data_as(dp), sky_err.ctypes.data_as(dp),
                                    n_sky.ctypes.data_as(ip), n_good.ctypes.data_as(ip), flag.ctypes.data_as(ip))
    return rc, sky, sky_err, n_sky, n_good, flag, mt, pos.value


@pytest.mark.parametrize("name", ["t2_cr_sky0", "t1_sky0"])
def test_sky_bootstrap_replays_numpy_stream_bit_exactly(hostsim, name):
    import copy
    meta, g = load_golden(name)
    frame, axes, header, args = golden_inputs(meta)
    out = mo.run_frame(copy.deepcopy(frame), axes, dict(header), args, seed=meta["rng_seed"])
    lo, hi = (np.ascontiguousarray(a).copy() for a in out["sky_limits_unclamped"])
    data, errs = np.ascontiguousarray(frame["data"]).copy(), np.ascontiguousarray(frame["errs"]).copy()
    np.random.seed(meta["rng_seed"])
    state = np.random.get_state()
    rc, sky, sky_err, n_sky, n_good, flag, mt, pos = _sky(hostsim, data, errs, lo, hi, axes["data_spatial_floor"], state)
    assert rc == 0
    assert np.array_equal(n_sky, out["sky_n_sky"]) and np.array_equal(n_good, out["sky_n_good"])
    assert np.array_equal(flag == 1, out["sky_skipped"] == 1)
    modelled = flag == 0
    assert np.array_equal(sky[modelled], g["sky_model"])                 # bit-exact: same draws, same medians
    assert np.array_equal(sky_err[modelled], g["sky_uncertainty"])
    assert np.array_equal(lo, g["sky_limits"][0]) and np.array_equal(hi, g["sky_limits"][1])
    assert np.array_equal(data, g["data_after"]) and np.array_equal(errs, g["errs_after"])
    # the generator ends where numpy's does after the same frame
    np.random.seed(meta["rng_seed"])
    mo.run_frame(copy.deepcopy(frame), axes, dict(header), args, seed=meta["rng_seed"])
    after = np.random.get_state()
    assert pos == after[2] and np.array_equal(mt, after[1])


def test_mt_seed_matches_numpy(hostsim):
    mt = np.zeros(624, dtype=np.uint32)
    hostsim.hostsim_mt_seed(mt.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), ctypes.c_uint32(1234))
    np.random.seed(1234)
    assert np.array_equal(mt, np.random.get_state()[1])


def test_polynomial_sky_matches_reference_within_tolerance(hostsim):
    """sky_order = 2: same Gaussian draws (numpy legacy polar method on the MT19937 stream), QR polyfit
    instead of LAPACK gelsd -> agreement far inside the flux tolerance, and the same RNG end state."""
    import copy
    meta, g = load_golden("t1_sky2")
    frame, axes, header, args = golden_inputs(meta)
    out = mo.run_frame(copy.deepcopy(frame), axes, dict(header), args, seed=meta["rng_seed"])
    after = np.random.get_state()
    lo, hi = (np.ascontiguousarray(a).copy() for a in out["sky_limits_unclamped"])
    data, errs = np.ascontiguousarray(frame["data"]).copy(), np.ascontiguousarray(frame["errs"]).copy()
    nspat, ndisp = data.shape
    np.random.seed(meta["rng_seed"])
    state = np.random.get_state()
    u32p = ctypes.POINTER(ctypes.c_uint32)
    mt = np.ascontiguousarray(state[1], dtype=np.uint32).copy()
    pos = ctypes.c_int(int(state[2]))
    model, model_err = np.zeros((nspat, ndisp)), np.zeros((nspat, ndisp))
    flag, n_good = np.zeros(ndisp, dtype=np.int32), np.zeros(ndisp, dtype=np.int32)
    rc = hostsim.hostsim_sky_poly(data.ctypes.data_as(dp), errs.ctypes.data_as(dp), nspat, ndisp, lo.ctypes.data_as(dp),
                                  hi.ctypes.data_as(dp), int(axes["data_spatial_floor"]), int(args.sky_order),
                                  mt.ctypes.data_as(u32p), ctypes.byref(pos), model.ctypes.data_as(dp),
                                  model_err.ctypes.data_as(dp), flag.ctypes.data_as(ip), n_good.ctypes.data_as(ip))
    assert rc == 0
    modelled = flag == 0
    assert np.array_equal(n_good, out["sky_n_good"])
    ref_model, ref_err = g["sky_model"], g["sky_uncertainty"]          # (nspat, n_modelled)
    assert ref_model.shape == (nspat, modelled.sum())
    assert np.allclose(model[:, modelled], ref_model, rtol=1e-10, atol=0)
    assert np.allclose(model_err[:, modelled], ref_err, rtol=1e-8, atol=0)
    assert np.allclose(data, out["data_after"], rtol=0, atol=1e-9 * np.abs(ref_model).max())
    assert np.allclose(errs, out["errs_after"], rtol=1e-10)
    assert pos.value == after[2] and np.array_equal(mt, after[1])
