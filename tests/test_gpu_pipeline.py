"""GPU parity of the whole path (motes_run_frames_host) against the reference's golden outputs, plus batch properties."""
import copy

import numpy as np
import pytest

from conftest import load_golden, golden_inputs
from parity import param_errors, rel_err, assert_within_reference_noise

pytestmark = pytest.mark.gpu


def _check_frame(name, out, g, nspat, sky_exact=True):
    report = {}
    # integer outputs: bit-exact
    assert np.array_equal(out["ext_bins"][:, :2], g["ext_bins"][:, :2]), "extraction bin edges"
    if "sky_bins" in g and len(g["sky_bins"]):
        assert np.array_equal(out["sky_bins"][:, :2], g["sky_bins"][:, :2]), "sky bin edges"
    lo_flips = int((np.floor(out["ext_limits"][0]) != np.floor(g["ext_limits"][0])).sum())
    hi_flips = int((np.ceil(out["ext_limits"][1]) != np.ceil(g["ext_limits"][1])).sum())
    report["limit index flips"] = (lo_flips, hi_flips)
    assert lo_flips + hi_flips <= max(2, len(g["optimal"]) // 500), report
    # Moffat parameters
    for key in ("sky_params", "ext_params"):
        if key in g and len(g[key]):
            err = param_errors(out[key], g[key], nspat)
            report[key + " max"] = err.max(axis=0)
            scaled_ref = g[key][:, :6]
            assert_within_reference_noise(err, g, scaled_ref, 1.0, nspat, key="ctl_" + key)
    err = param_errors(out["median_params"][None], g["median_params"][None], nspat)
    assert err.max() < 1e-6
    assert out["scale"] == g["scale"]
    # sky model: same draws as the reference (device MT19937); the limits feeding it are floats from
    # fits, so exactness needs the per-column sky-pixel sets to agree, which they do unless a limit
    # crosses an integer
    if "sky_model" in g and sky_exact:
        same = out["sky_model"].shape == g["sky_model"].shape and np.array_equal(out["sky_model"], g["sky_model"])
        report["sky model bit-exact"] = same
        if not same:
            assert out["sky_model"].shape == g["sky_model"].shape
            assert rel_err(out["sky_model"], g["sky_model"]).max() < 1e-5
    # spectra: north_star tolerance rel 1e-5 (floor: 1e-5 of the typical flux for near-zero columns)
    for key in ("optimal", "optimal_err", "aperture", "aperture_err"):
        floor = 1e-5 * np.median(np.abs(g[key][g[key] != 0])) if np.any(g[key] != 0) else 1.0
        err = np.abs(out[key] - g[key]) / np.maximum(np.abs(g[key]), floor)
        unaffected = np.ones(len(err), dtype=bool)
        report[key + " max rel"] = err.max()
        # columns whose integer aperture flipped by one pixel are counted, not compared
        if lo_flips + hi_flips:
            unaffected = (np.floor(out["ext_limits"][0]) == np.floor(g["ext_limits"][0])) & \
                         (np.ceil(out["ext_limits"][1]) == np.ceil(g["ext_limits"][1]))
        assert err[unaffected].max() < 1e-5, (key, report)
    print(name, report)


@pytest.mark.parametrize("name", ["t1_sky0", "t1_nosky", "t2_cr_sky0", "t1_faint", "t2_unbinned", "c1_sky0", "t1_sky2", "t1_f32"])
def test_whole_path_matches_reference(name):
    from motes_b200 import pipeline
    meta, g = load_golden(name)
    frame, axes, header, args = golden_inputs(meta)
    out = pipeline.run_frame(copy.deepcopy(frame), axes, dict(header), args, seed=meta["rng_seed"])
    _check_frame(name, out, g, frame["data"].shape[0], sky_exact=(int(args.sky_order) == 0))
    if int(args.sky_order) > 0:
        assert np.allclose(out["sky_model"], g["sky_model"], rtol=1e-6)


def test_batch_equals_single_frames_and_reports_per_frame_status():
    """Frames are independent: a batch gives exactly what each frame gives alone; a frame the reference
    would abort on (no sky pixels) is flagged without disturbing its neighbours."""
    from motes_b200 import pipeline, synth
    args = synth.default_args()
    specs = [synth.config_spec("T2", seed=s) for s in (31, 32, 33)]
    frames = [synth.make_frame(s) for s in specs]
    data = np.stack([f[0]["data"] for f in frames])
    errs = np.stack([f[0]["errs"] for f in frames])
    qual = np.stack([f[0]["qual"] for f in frames])
    # frame 1: a trace so wide that the sky window swallows the slit
    wide = synth.make_frame(synth.config_spec("T2", seed=32, alpha=14.0))
    data[1], errs[1] = wide[0]["data"], wide[0]["errs"]
    axes, header = frames[0][1], frames[0][2]
    seeds = [5, 6, 7]
    batch = pipeline.run_frames(data.copy(), errs.copy(), qual, args, header["seeing"], header["pixel_resolution"],
                                seeds=seeds, axes_dict=axes)
    assert batch["frame_status"][0] == 0 and batch["frame_status"][2] == 0
    assert batch["frame_status"][1] == -4
    for i in (0, 2):
        single = pipeline.run_frames(data[i:i + 1].copy(), errs[i:i + 1].copy(), qual[i:i + 1], args,
                                     header["seeing"], header["pixel_resolution"], seeds=seeds[i:i + 1], axes_dict=axes)
        n = int(batch["n_ext_bins"][i])
        assert n == int(single["n_ext_bins"][0])
        for key in ("spectra", "ext_limits", "sky_model"):
            assert np.array_equal(batch[key][i], single[key][0], equal_nan=True), key
        assert np.array_equal(batch["ext_params"][i, :n], single["ext_params"][0, :n])


def test_linearity_property_full_size():
    """Size-independent property at BASELINE.json's configs[1] shape (4096x400, no oracle run needed):
    without sky subtraction, scaling data and errs by 4 leaves the bins identical, moves the limits only
    at rounding level (the decimal scale factor of tracing.py:437 changes, so the fits are not bitwise
    scale-free), scales the aperture flux by exactly 4 and the optimal flux by 4 within 1e-6."""
    from motes_b200 import pipeline, synth
    frame, axes, header = synth.make_frame(synth.config_spec("C2", seed=77))
    args = synth.default_args(subtract_sky=False)
    a = pipeline.run_frames(frame["data"][None], frame["errs"][None], frame["qual"][None], args, header["seeing"],
                            header["pixel_resolution"], axes_dict=axes)
    b = pipeline.run_frames(4.0 * frame["data"][None], 4.0 * frame["errs"][None], frame["qual"][None], args,
                            header["seeing"], header["pixel_resolution"], axes_dict=axes)
    assert a["frame_status"][0] == 0 and b["frame_status"][0] == 0
    n = int(a["n_ext_bins"][0])
    assert n == int(b["n_ext_bins"][0]) and n > 200
    assert np.array_equal(a["ext_bins"][0, :n], b["ext_bins"][0, :n])
    assert np.allclose(a["ext_limits"], b["ext_limits"], rtol=1e-7, atol=0)
    same = (np.floor(a["ext_limits"][0, 0]) == np.floor(b["ext_limits"][0, 0])) & \
           (np.ceil(a["ext_limits"][0, 1]) == np.ceil(b["ext_limits"][0, 1]))
    assert same.mean() > 0.999
    assert np.array_equal(4.0 * a["spectra"][0, 2:, same], b["spectra"][0, 2:, same])          # aperture flux / error
    assert np.allclose(4.0 * a["spectra"][0, :2, same], b["spectra"][0, :2, same], rtol=1e-6)   # optimal flux / error
    assert np.allclose(a["ext_params"][0, :n, 1:4], b["ext_params"][0, :n, 1:4], rtol=1e-5)
    # aperture flux is the plain sum of the pixels inside the integer limits
    lo = np.floor(a["ext_limits"][0, 0]).astype(int)
    hi = np.ceil(a["ext_limits"][0, 1]).astype(int) + 1
    cols = [100, 2048, 4000]
    for c in cols:
        assert a["spectra"][0, 2, c] == pytest.approx(frame["data"][lo[c]:hi[c], c].sum(), rel=1e-13)


@pytest.mark.parametrize("dtype,config,over", [(">f4", "T1", {}), ("<f4", "T2", {}), (">f4", "T1", {"nspat": 97, "ndisp": 1003, "chip_gaps": ((333, 21),)})])
def test_float32_planes_are_widened_on_the_device(dtype, config, over):
    """GMOS planes arrive as big-endian float32 (gmosio.py:40-52).  motes_run_frames_host_f32 uploads them as they are
    (4 bytes per pixel) and widens them on the device: every output equals, bit for bit, the float64 path on the same
    values converted on the host - including frames whose pixel count is not a multiple of four."""
    from motes_b200 import pipeline, synth
    frame, axes, header = synth.make_frame(synth.config_spec(config, seed=41, **over))
    args = synth.default_args()
    d32, e32 = frame["data"].astype(dtype)[None], frame["errs"].astype(dtype)[None]
    a = pipeline.run_frames(d32, e32, frame["qual"][None], args, header["seeing"], header["pixel_resolution"], seeds=[9], axes_dict=axes)
    b = pipeline.run_frames(d32.astype(np.float64), e32.astype(np.float64), frame["qual"][None], args, header["seeing"],
                            header["pixel_resolution"], seeds=[9], axes_dict=axes)
    assert a["frame_status"][0] == 0
    for key in pipeline.OUTPUT_SHAPES:
        assert np.array_equal(a[key], b[key], equal_nan=True), key
    assert np.array_equal(a["rng_states"], b["rng_states"])


def test_frame_whose_first_fit_cannot_start_is_left_untouched():
    """A frame on which the reference raises ValueError in median_moffat_fitting (x0 outside the bounds: scipy's check,
    tracing.py:591-598) is reported with status MOTES_E_FIT, runs no further stage - its pixels come back exactly as
    they went in, its spectra are zero - and does not disturb its neighbours."""
    from motes_b200 import pipeline, synth
    args = synth.default_args()
    frames = [synth.make_frame(synth.config_spec("T2", seed=s)) for s in (51, 52)]
    data = np.stack([f[0]["data"] for f in frames])
    errs = np.stack([f[0]["errs"] for f in frames])
    qual = np.stack([f[0]["qual"] for f in frames])
    data[0] = -np.abs(data[0]) - 10.0                 # negative everywhere: the amplitude start value is below its bound
    axes, header = frames[0][1], frames[0][2]
    before, errs_before = data.copy(), errs.copy()
    batch = pipeline.run_frames(data, errs, qual, args, header["seeing"], header["pixel_resolution"], seeds=[3, 4],
                                axes_dict=axes, copy_back=True)
    assert batch["frame_status"][0] == -5 and batch["frame_status"][1] == 0
    assert np.array_equal(batch["data"][0], before[0]), "the failed frame was modified"
    assert np.all(batch["spectra"][0] == 0.0) and batch["n_ext_bins"][0] == 0
    assert np.array_equal(batch["errs"][0], errs_before[0]), "the failed frame's errors were modified"
    single = pipeline.run_frames(before[1:2].copy(), errs_before[1:2].copy(), qual[1:2], args, header["seeing"],
                                 header["pixel_resolution"], seeds=[4], axes_dict=axes)
    assert np.array_equal(batch["spectra"][1], single["spectra"][0], equal_nan=True)
    with pytest.raises(ValueError):
        pipeline.run_frame({"data": before[0].copy(), "errs": errs_before[0].copy(), "qual": qual[0]}, axes, dict(header), args, seed=3)
