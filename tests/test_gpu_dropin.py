"""The literal drop-in of INTEGRATION.md §1: ``core.motes``'s stage order (core.py:80-275) with the three module names
rebound to the CUDA-backed mirrors,

    tracing, sky, extraction = motes_b200.tracing, motes_b200.sky, motes_b200.extraction

driven by the same stage driver the golden files were made with (oracle/ref_runner.run_stages), compared with what the
unmodified reference produced.  Covers the per-function mirrors the fused path does not go through:
``median_moffat_fitting`` (header["seeing"] overwritten, tracing.py:449), ``set_extraction_limits``, ``get_bins``,
``sky_locator`` (in-place clamping of the returned limits, sky.py:303-307; frame_dict keys added), per-bin
``moffat_fitting`` (limit rows appended to the caller's list), ``interpolate_extraction_lims``, ``optimal_extraction``.
"""
import copy

import numpy as np
import pytest

from conftest import load_golden, golden_inputs
from parity import param_errors, assert_within_reference_noise

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["t1_sky0", "t2_cr_sky0", "t1_sky2", "t1_nosky"])
def test_four_line_swap_reproduces_reference(name):
    from motes_b200 import extraction, sky, tracing
    from oracle import ref_runner
    meta, g = load_golden(name)
    frame, axes, header, args = golden_inputs(meta)
    frame, header = copy.deepcopy(frame), dict(header)
    seeing_in = header["seeing"]
    out = ref_runner.run_stages(tracing, sky, extraction, frame, axes, header, args, seed=meta["rng_seed"])
    nspat = g["median_profile"].shape[0]
    # header mutation of median_moffat_fitting
    assert header["seeing"] != seeing_in and header["seeing"] == pytest.approx(float(g["seeing_px"]), rel=1e-6)
    assert out["scale"] == float(g["scale"])
    assert np.array_equal(out["bin_rows"], g["bin_rows"])
    assert np.array_equal(out["ext_bins"][:, :2], g["ext_bins"][:, :2])
    assert np.allclose(out["ext_bins"][:, 2], g["ext_bins"][:, 2], rtol=1e-12)
    assert out["ext_limit_table"].shape == g["ext_limit_table"].shape
    if not args.subtract_sky or int(args.sky_order) == 0:
        assert np.array_equal(out["ext_profiles"], g["ext_profiles"], equal_nan=True)
    else:       # polynomial sky: tolerance parity (QR instead of LAPACK gelsd), so the sky-subtracted medians agree to rounding
        assert np.allclose(out["ext_profiles"], g["ext_profiles"], rtol=1e-6, atol=1e-6 * np.nanmax(np.abs(g["ext_profiles"])),
                           equal_nan=True)
    if args.subtract_sky:
        assert np.array_equal(out["sky_bins"][:, :2], g["sky_bins"][:, :2])
        assert out["sky_params"].shape == g["sky_params"].shape
        err = param_errors(out["sky_params"], g["sky_params"], nspat)
        assert_within_reference_noise(err, g, g["sky_params"][:, :6], 1.0, nspat, key="ctl_sky_params")
        # the limits sky_locator returns are the clamped ones (sky.py:303-307)
        assert out["sky_limits"].shape == g["sky_limits"].shape
        assert np.all(out["sky_limits"][0] >= 0) and np.all(out["sky_limits"][1] <= nspat)
        flips = int((np.floor(out["sky_limits"][0]) != np.floor(g["sky_limits"][0])).sum()
                    + (np.ceil(out["sky_limits"][1]) != np.ceil(g["sky_limits"][1])).sum())
        if int(args.sky_order) == 0:
            if flips == 0:
                assert np.array_equal(out["sky_model"], g["sky_model"])
                assert np.array_equal(out["sky_uncertainty"], g["sky_uncertainty"])
                if "data_after" in g:
                    assert np.array_equal(out["data_after"], g["data_after"], equal_nan=True)
                    assert np.array_equal(out["errs_after"], g["errs_after"], equal_nan=True)
        else:
            assert np.allclose(out["sky_model"], g["sky_model"], rtol=1e-6)
        for key in ("sky_model", "sky_uncertainty"):
            assert key in frame, key
    err = param_errors(out["ext_params"], g["ext_params"], nspat)
    assert_within_reference_noise(err, g, g["ext_params"][:, :6], 1.0, nspat, key="ctl_ext_params")
    assert np.array_equal(out["ext_params"][:, 6:], g["ext_params"][:, 6:])          # bin edges + wavelength start
    same = (np.floor(out["ext_limits"][0]) == np.floor(g["ext_limits"][0])) & \
           (np.ceil(out["ext_limits"][1]) == np.ceil(g["ext_limits"][1]))
    assert (~same).sum() <= 2
    for key in ("optimal", "optimal_err", "aperture", "aperture_err"):
        ref = g[key]
        floor = 1e-5 * np.median(np.abs(ref[ref != 0]))
        rel = np.abs(out[key] - ref) / np.maximum(np.abs(ref), floor)
        assert rel[same].max() < 1e-5, key
    # core.py:265-267 leaves the frames transposed; optimal_extraction scrubbed non-finite pixels in place
    assert frame["data"].shape == (axes["dispersion_axis_length"], nspat)
    assert np.isfinite(frame["data"]).all()


def test_numpy_generator_advances_like_the_reference():
    """The four-line swap consumes numpy's global legacy stream exactly as the reference does: the next draw after a
    frame is the same number in both worlds."""
    from motes_b200 import extraction, sky, tracing
    from oracle import motes_oracle as mo, ref_runner
    import warnings
    meta, _ = load_golden("t2_cr_sky0")
    frame, axes, header, args = golden_inputs(meta)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mo.run_frame(copy.deepcopy(frame), axes, dict(header), args, seed=77)
    want = np.random.randint(0, 2 ** 31 - 1, size=4)
    ref_runner.run_stages(tracing, sky, extraction, copy.deepcopy(frame), axes, dict(header), args, seed=77)
    got = np.random.randint(0, 2 ** 31 - 1, size=4)
    assert np.array_equal(got, want)
