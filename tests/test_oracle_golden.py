"""The oracle must reproduce the reference's own outputs (tests/golden/*.npz) bit for bit."""
import copy

import numpy as np
import pytest

from conftest import load_golden, golden_inputs
from oracle import motes_oracle as mo

FAST = ["t1_sky0", "t1_nosky", "t2_cr_sky0", "t1_faint", "t2_unbinned", "t1_sky2", "t1_f32"]
STAGES = ["median_profile", "median_params", "scale", "seeing_px", "bin_rows", "sky_bins", "sky_params",
          "sky_limits", "sky_model", "sky_uncertainty", "ext_bins", "ext_params", "ext_limit_table",
          "ext_profiles", "ext_limits", "data_after", "errs_after", "optimal", "optimal_err",
          "aperture", "aperture_err"]


def _run(name):
    meta, gold = load_golden(name)
    frame, axes, header, args = golden_inputs(meta)
    out = mo.run_frame(copy.deepcopy(frame), axes, dict(header), args, seed=meta["rng_seed"])
    return gold, out


@pytest.mark.parametrize("name", FAST)
def test_oracle_matches_reference_bitwise(name):
    gold, out = _run(name)
    checked = 0
    for key in STAGES:
        if key not in gold:
            continue
        a, b = np.asarray(gold[key]), np.asarray(out[key])
        assert a.shape == b.shape, key
        assert np.array_equal(a, b, equal_nan=True), key
        checked += 1
    assert checked >= 12


def test_oracle_matches_reference_c1_full_size():
    """BASELINE.json configs[0]: 3000x200 GMOS-like frame with two chip gaps, sky order 0."""
    gold, out = _run("c1_sky0")
    for key in ("sky_bins", "ext_bins", "sky_params", "ext_params", "sky_limits", "ext_limits",
                "sky_model", "sky_uncertainty", "optimal", "optimal_err", "aperture", "aperture_err"):
        assert np.array_equal(gold[key], out[key], equal_nan=True), key
    # chip-gap columns come out as exact zeros in the optimal spectrum (extraction.py:193-195)
    assert np.all(out["optimal"][980:1020] == 0.0)


def test_integer_traces_are_consistent():
    gold, out = _run("t2_cr_sky0")
    lo, hi = out["lo_idx"], out["hi_idx"]
    assert np.array_equal(lo, np.floor(out["ext_limits"][0]).astype(np.int32))
    assert np.array_equal(hi, np.ceil(out["ext_limits"][1]).astype(np.int32) + 1)
    # the 5-sigma mask rejects something on a frame with 1 % cosmic rays
    inside = np.zeros_like(out["crmask"], dtype=bool)
    for i, (a, b) in enumerate(zip(lo, hi)):
        inside[i, a:b] = True
    assert (inside & (out["crmask"] == 0)).sum() > 0


def test_no_sky_pixels_raises_runtime_error():
    """sky.py:323-335: a sky window that swallows the slit aborts with RuntimeError."""
    from motes_b200 import synth
    frame, axes, header = synth.make_frame(synth.config_spec("T2", seed=3))
    args = synth.default_args()
    n = axes["dispersion_axis_length"]
    with pytest.raises(RuntimeError):
        mo.subtract_sky(np.full(n, -5.0), np.full(n, 1e9), frame, axes, args)
