"""GPU parity: optimal + aperture extraction (C ABI motes_optimal_extract_host) vs the reference's spectra."""
import numpy as np
import pytest

from conftest import load_golden, golden_inputs
from parity import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["t1_sky0", "t2_cr_sky0"])
def test_extraction_matches_reference_on_reference_inputs(name):
    """Feed the kernel the reference's own sky-subtracted frame, limits and bin parameters."""
    from motes_b200 import extraction
    from oracle import motes_oracle as mo
    meta, g = load_golden(name)
    _, axes, _, _ = golden_inputs(meta)
    data_t = np.ascontiguousarray(g["data_after"].T)
    errs_t = np.ascontiguousarray(g["errs_after"].T)
    limits = [g["ext_limits"][0], g["ext_limits"][1]]
    got, masks = extraction.optimal_extraction(data_t.copy(), errs_t.copy(), limits, g["ext_params"], axes,
                                               return_masks=True)
    for k, key in enumerate(["optimal", "optimal_err", "aperture", "aperture_err"]):
        err = rel_err(got[k], g[key], floor=1e-9 * np.abs(g[key]).max())
        print(name, key, "max rel", err.max())
        assert err.max() < 1e-9, key
    # integer outputs bit-exact against the oracle's trace
    trace = {}
    mo.optimal_extract(data_t.copy(), errs_t.copy(), limits, g["ext_params"], axes, trace)
    assert np.array_equal(masks["lo_idx"], trace["lo_idx"])
    assert np.array_equal(masks["hi_idx"], trace["hi_idx"])
    flips = int((masks["crmask"] != trace["crmask"]).sum())
    print(name, "cosmic-ray mask flips:", flips, "rejected pixels:", int((trace["crmask"] == 0).sum()))
    assert flips == 0


def test_extraction_edge_cases():
    """All-zero error column, zero errors filled by the column median, non-finite pixels, chip gap."""
    from motes_b200 import extraction, synth
    from oracle import motes_oracle as mo
    rng = np.random.default_rng(3)
    nspat, ndisp = 48, 40
    axes = synth.make_axes(nspat, ndisp)
    r = np.arange(nspat)[None, :]
    data = 50.0 * (1 + ((r - 23.3) ** 2) / 9.0) ** -3.0 + rng.normal(0, 1.0, (ndisp, nspat))
    errs = np.full((ndisp, nspat), 1.0) + 0.1 * rng.random((ndisp, nspat))
    errs[3, :] = 0.0                      # whole column bad -> zeros out (extraction.py:133-137)
    errs[5, [2, 20, 30]] = 0.0            # filled with the median of the non-zero errors
    errs[6, 22] = -errs[6, 22]            # |errs|
    data[7, 21] = np.nan
    data[8, 25] = np.inf
    errs[9, 24] = np.nan
    data[11, :] = 1.0                     # chip gap marker -> optimal 0
    data[12, 23] += 500.0                 # cosmic ray
    limits = [np.full(ndisp, 14.2), np.full(ndisp, 32.7)]
    limits[0][15] = -3.5                  # negative lower limit: python slice wraps (extraction.py:122)
    bins = np.array([[50.0, 23.3, 3.0, 3.0, 0.0, 0.0, 0.0, 20.0], [50.0, 23.3, 3.0, 3.0, 0.0, 0.0, 20.0, 40.0]])
    d_ref, e_ref = data.copy(), errs.copy()
    trace = {}
    want = mo.optimal_extract(d_ref, e_ref, limits, bins, axes, trace)
    d_got = data.copy()
    got, masks = extraction.optimal_extraction(d_got, errs.copy(), limits, bins, axes, return_masks=True)
    for k in range(4):
        assert np.allclose(got[k], want[k], rtol=1e-10, atol=1e-12), k
    assert got[0][3] == 0.0 and got[2][3] == 0.0 and got[0][11] == 0.0
    assert np.array_equal(d_got, d_ref)           # the in-place scrub of non-finite pixels
    assert np.array_equal(masks["lo_idx"], trace["lo_idx"]) and np.array_equal(masks["hi_idx"], trace["hi_idx"])
    assert np.array_equal(masks["crmask"], trace["crmask"])
    assert trace["crmask"][12, 23] == 0
