"""GPU parity of the stage entry points (C ABI through the drop-in mirrors) against the reference's golden outputs."""
import copy

import numpy as np
import pytest

from conftest import load_golden, golden_inputs

pytestmark = pytest.mark.gpu

CASES = ["t1_sky0", "t2_cr_sky0", "t1_faint", "t2_unbinned", "c1_sky0"]


def test_row_median_matches_numpy_nanmedian():
    from motes_b200 import _lib, synth
    frame, axes, header = synth.make_frame(synth.config_spec("T2", seed=8, nan_fraction=0.05))
    data = np.ascontiguousarray(frame["data"])
    data[3, :] = np.nan
    out = np.zeros(data.shape[0])
    h = _lib.default_handle()
    _lib.check(_lib.lib.motes_row_median_host(h.raw, _lib.ptr(data), data.shape[0], data.shape[1], _lib.ptr(out)))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = np.nanmedian(data, axis=1)
    assert np.array_equal(out, want, equal_nan=True)


def test_row_median_of_a_dispersion_axis_longer_than_shared_memory():
    """40 000 columns do not fit the shared-memory staging of row_median_kernel (227 KB): the long-axis kernel selects
    straight from global memory and still equals np.nanmedian bit for bit (NaNs, ties, even and odd counts)."""
    from motes_b200 import _lib
    rng = np.random.default_rng(5)
    data = rng.normal(100.0, 7.0, size=(6, 40001))
    data[1, ::7] = np.nan
    data[2, :] = np.round(data[2, :])                 # many ties
    data[3, :20001] = np.nan                          # even number of valid values
    data[4, :] = np.nan
    out = np.zeros(data.shape[0])
    h = _lib.default_handle()
    _lib.check(_lib.lib.motes_row_median_host(h.raw, _lib.ptr(data), data.shape[0], data.shape[1], _lib.ptr(out)))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = np.nanmedian(data, axis=1)
    assert np.array_equal(out, want, equal_nan=True)


@pytest.mark.parametrize("name", CASES)
def test_get_bins_bit_exact(name):
    from motes_b200 import tracing
    meta, g = load_golden(name)
    frame, axes, header, args = golden_inputs(meta)
    row_lo, row_hi = (int(v) for v in g["bin_rows"])
    bins, _ = tracing.get_bins(frame, row_lo, row_hi, axes["dispersion_axis_length"], args, has_sky=True)
    got = np.array(bins)
    assert np.array_equal(got[:, :2], g["sky_bins"][:, :2])
    assert np.allclose(got[:, 2], g["sky_bins"][:, 2], rtol=1e-12, equal_nan=True)


@pytest.mark.parametrize("name", CASES)
def test_bin_profiles_bit_exact(name):
    """Per-bin nanmedian profiles incl. chip-gap column exclusion; run on the raw frame with the sky bins."""
    from motes_b200 import tracing
    from oracle import motes_oracle as mo
    meta, g = load_golden(name)
    frame, axes, header, args = golden_inputs(meta)
    bins = g["sky_bins"][:, :2].astype(int)
    got = tracing.bin_profiles(frame["data"], bins)
    want = np.array([mo.bin_profile(frame["data"], b[0], b[1]) for b in bins])
    assert np.array_equal(got, want, equal_nan=True)


@pytest.mark.parametrize("name", CASES)
def test_interpolate_limits_bit_exact(name):
    from motes_b200 import tracing
    meta, g = load_golden(name)
    lo, hi = tracing.interpolate_extraction_lims(g["ext_limit_table"], meta["spec"]["ndisp"])
    assert np.array_equal(lo, g["ext_limits"][0]) and np.array_equal(hi, g["ext_limits"][1])


@pytest.mark.parametrize("name", ["t1_sky0", "t2_cr_sky0", "c1_sky0"])
def test_sky_subtraction_replays_numpy_rng_bit_exactly(name):
    """Device MT19937 + masked rejection consume np.random's stream exactly: sky model, uncertainty,
    sky-subtracted frame and the generator state afterwards are all identical to the reference's."""
    from motes_b200 import sky
    from oracle import motes_oracle as mo
    meta, g = load_golden(name)
    frame, axes, header, args = golden_inputs(meta)
    ref = mo.run_frame(copy.deepcopy(frame), axes, dict(header), args, seed=meta["rng_seed"])
    state_after_reference = np.random.get_state()
    lo, hi = (a.copy() for a in ref["sky_limits_unclamped"])
    fd = {"data": frame["data"].copy(), "errs": frame["errs"].copy(), "qual": frame["qual"]}
    np.random.seed(meta["rng_seed"])
    trace = {}
    fd = sky.subtract_sky(lo, hi, fd, axes, args, header, "synthetic", trace=trace)
    assert np.array_equal(trace["n_sky"], ref["sky_n_sky"]) and np.array_equal(trace["n_good"], ref["sky_n_good"])
    assert np.array_equal(fd["sky_model"][0], g["sky_model"])
    assert np.array_equal(fd["sky_uncertainty"][0], g["sky_uncertainty"])
    assert np.array_equal(lo, g["sky_limits"][0]) and np.array_equal(hi, g["sky_limits"][1])
    assert np.array_equal(fd["data"], ref["data_after"] if "data_after" not in g else g["data_after"])
    mine = np.random.get_state()
    assert mine[2] == state_after_reference[2] and np.array_equal(mine[1], state_after_reference[1])


def test_sky_without_sky_pixels_raises_runtime_error():
    from motes_b200 import sky, synth
    frame, axes, header = synth.make_frame(synth.config_spec("T2", seed=3))
    n = axes["dispersion_axis_length"]
    with pytest.raises(RuntimeError):
        sky.subtract_sky(np.full(n, -5.0), np.full(n, 1e9), frame, axes, synth.default_args(), header, "x")


def test_polynomial_sky_subtraction_matches_reference():
    """sky_order = 2 (sky.sky_polynomial): same Gaussian draws from np.random's stream, QR polyfit."""
    from motes_b200 import sky
    from oracle import motes_oracle as mo
    meta, g = load_golden("t1_sky2")
    frame, axes, header, args = golden_inputs(meta)
    ref = mo.run_frame(copy.deepcopy(frame), axes, dict(header), args, seed=meta["rng_seed"])
    after = np.random.get_state()
    lo, hi = (a.copy() for a in ref["sky_limits_unclamped"])
    fd = {"data": frame["data"].copy(), "errs": frame["errs"].copy(), "qual": frame["qual"]}
    np.random.seed(meta["rng_seed"])
    fd = sky.subtract_sky(lo, hi, fd, axes, args, header, "synthetic")
    assert fd["sky_model"].shape == g["sky_model"].shape
    assert np.allclose(fd["sky_model"], g["sky_model"], rtol=1e-10, atol=0)
    assert np.allclose(fd["sky_uncertainty"], g["sky_uncertainty"], rtol=1e-8, atol=0)
    assert np.allclose(fd["data"], ref["data_after"], rtol=0, atol=1e-9 * np.abs(g["sky_model"]).max())
    mine = np.random.get_state()
    assert mine[2] == after[2] and np.array_equal(mine[1], after[1])
