"""oracle/trf_restated.py must walk the same iterates as scipy's least_squares(method='trf')."""
import numpy as np
import pytest
from scipy.optimize import least_squares

from conftest import load_golden, golden_inputs
from oracle import motes_oracle as mo
from oracle.trf_restated import trf_bounded, fd_steps


def _profiles(name, count):
    meta, gold = load_golden(name)
    return gold["ext_profiles"][:count], gold["scale"], gold["seeing_px"], meta


@pytest.mark.parametrize("name", ["t1_sky0", "t2_cr_sky0", "t1_faint"])
def test_restated_trf_is_bit_identical_to_scipy(name):
    profiles, scale, seeing_px, meta = _profiles(name, 12)
    axis = np.linspace(0.0, profiles.shape[1] - 1.0, profiles.shape[1])
    pixres = meta["spec"]["pixel_resolution"]
    for prof in profiles:
        col = prof * scale
        x0, lo, hi = mo.fit_start_and_bounds(col, seeing_px, pixres)
        ref = least_squares(mo._residual, x0, bounds=(lo, hi), args=(axis, col), method="trf", ftol=1e-12)
        got = trf_bounded(mo._residual, x0, lo, hi, args=(axis, col), ftol=1e-12)
        assert np.array_equal(ref.x, got.x)
        assert (ref.nfev, ref.njev, ref.status) == (got.nfev, got.njev, got.status)
        assert ref.cost == got.cost


def test_whole_pipeline_with_restated_solver_matches_golden():
    import copy
    meta, gold = load_golden("t1_nosky")
    frame, axes, header, args = golden_inputs(meta)
    out = mo.run_frame(copy.deepcopy(frame), axes, dict(header), args, seed=meta["rng_seed"], solver=trf_bounded)
    for key in ("median_params", "ext_params", "ext_limits", "optimal", "aperture"):
        assert np.array_equal(gold[key], out[key]), key


def test_fd_steps_respect_bounds():
    x = np.array([1e-10, 49.0, 5.0, 4.999999999999999, -3.0, 0.0])
    lb = np.array([0.0, 48.0, 0.0, 0.0, -np.inf, -np.inf])
    ub = np.array([np.inf, 50.0, 25.0, 5.0, np.inf, np.inf])
    h = fd_steps(x, lb, ub)
    assert np.all((x + h >= lb) & (x + h <= ub))
    assert h[3] < 0 and h[4] < 0 and h[5] > 0


def test_infeasible_start_raises_value_error():
    with pytest.raises(ValueError):
        trf_bounded(lambda x: x, [-1.0], [0.0], [1.0])
