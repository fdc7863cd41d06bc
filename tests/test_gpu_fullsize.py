"""Full-size runs of BASELINE.json's configurations (no oracle run: the CPU path needs 8-50 s per frame there).

What is checked instead are properties that do not depend on the size:
  * the segmented generator replay (jump-ahead / stream / walk / histogram kernels; generator started early on the
    second stream, or where it is consumed) and the literal one-CTA replay of ``np.random.choice`` (validated bit
    for bit against numpy on the golden cases) give the same sky model, the same spectra and leave numpy's
    RandomState in the same state;
  * a frame gives the same result alone and inside a batch;
  * the aperture flux is the plain sum of the sky-subtracted pixels between the integer limits
    (extraction.py:87-88), the optimal flux equals it within the noise where the profile is well sampled,
    and every column of a cosmic-ray frame stays finite.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = {
    # name: (synth config, sky expected to be modelled)
    "C1_gmos_3000x200": ("C1", {}),
    "C3_xshooter_24000x80_cr": ("C3", {}),
    "C5b_faint_16384x512": ("C5b", {}),
    # a wide slit with cosmic rays: outliers are clipped, so the good-pixel count changes from column to column and
    # the table walk's runs are short (C3's 80 rows keep a single outlier inside 10 sigma: long runs)
    "C2_fors2_4096x400_cr": ("C2", {"cosmic_fraction": 0.01}),
}


def _fresh_handle(env):
    """A handle created under the given MOTES_B200_* switches (they are read by motes_create)."""
    from motes_b200 import _lib
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return _lib.Handle(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("name", list(CASES))
def test_segmented_replay_equals_literal_replay_at_full_size(name):
    from motes_b200 import pipeline, synth
    cfg, over = CASES[name]
    frame, axes, header = synth.make_frame(synth.config_spec(cfg, seed=11, **over))
    args = synth.default_args()
    runs = {}
    for label, env in (("segmented", {"MOTES_B200_BOOTSTRAP": "segmented"}), ("literal", {"MOTES_B200_BOOTSTRAP": "classic"}),
                       ("slots", {"MOTES_B200_BOOTSTRAP": "segmented", "MOTES_B200_COLUMN": "0"}),
                       # a histogram window of 24 ranks: many samples' medians fall outside, their columns go through the
                       # retry list to the exact kernel
                       ("narrow_window", {"MOTES_B200_BOOTSTRAP": "segmented", "MOTES_B200_COLUMN": "24"}),
                       ("unhoisted", {"MOTES_B200_BOOTSTRAP": "segmented", "MOTES_B200_HOIST": "0"}),
                       ("table_walk", {"MOTES_B200_BOOTSTRAP": "segmented", "MOTES_B200_WALK": "table"}),
                       ("serial_walk", {"MOTES_B200_BOOTSTRAP": "segmented", "MOTES_B200_WALK": "serial"}),
                       ("stream_walk", {"MOTES_B200_BOOTSTRAP": "segmented", "MOTES_B200_WALK": "stream"})):
        h = _fresh_handle(env)
        data, errs = frame["data"][None].copy(), frame["errs"][None].copy()
        runs[label] = pipeline.run_frames(data, errs, frame["qual"][None], args, header["seeing"], header["pixel_resolution"],
                                          seeds=[2024], axes_dict=axes, copy_back=True, handle=h)
        runs[label]["data"], runs[label]["errs"] = data, errs
        h.close()
    a = runs["segmented"]
    if a["frame_status"][0] != 0:          # the reference would abort on this frame: every variant must say so
        assert all(r["frame_status"][0] == a["frame_status"][0] for r in runs.values())
        pytest.skip(f"frame status {a['frame_status'][0]} (what the reference raises) in every variant")
    assert int((a["sky_flag"][0] == 0).sum()) > 0.9 * a["sky_flag"].shape[1]          # the sky was modelled
    for other in ("literal", "slots", "narrow_window", "unhoisted", "table_walk", "serial_walk", "stream_walk"):
        b = runs[other]
        assert np.array_equal(a["rng_states"], b["rng_states"]), other
        for key in ("sky_model", "sky_uncertainty", "spectra", "ext_limits", "sky_limits", "data", "errs"):
            assert np.array_equal(a[key], b[key], equal_nan=True), (other, key)

    # aperture flux = plain sum of the sky-subtracted pixels inside the integer limits (extraction.py:87-88)
    lo = np.floor(a["ext_limits"][0, 0]).astype(int)
    hi = np.ceil(a["ext_limits"][0, 1]).astype(int) + 1
    nd = frame["data"].shape[1]
    for c in (3, nd // 3, nd // 2, nd - 5):
        want = np.nansum(a["data"][0, max(lo[c], 0):hi[c], c])
        assert a["spectra"][0, 2, c] == pytest.approx(want, rel=1e-12, abs=1e-9)
    assert np.all(np.isfinite(a["spectra"][0, 2:]))


def test_frame_alone_equals_frame_in_batch_at_full_size():
    from motes_b200 import pipeline, synth
    frames = [synth.make_frame(synth.config_spec("C3", seed=s)) for s in (21, 22)]
    data = np.stack([f[0]["data"] for f in frames])
    errs = np.stack([f[0]["errs"] for f in frames])
    qual = np.stack([f[0]["qual"] for f in frames])
    axes, header = frames[0][1], frames[0][2]
    args = synth.default_args()
    batch = pipeline.run_frames(data.copy(), errs.copy(), qual, args, header["seeing"], header["pixel_resolution"],
                                seeds=[5, 6], axes_dict=axes)
    assert np.all(batch["frame_status"] == 0)
    single = pipeline.run_frames(data[1:2].copy(), errs[1:2].copy(), qual[1:2], args, header["seeing"],
                                 header["pixel_resolution"], seeds=[6], axes_dict=axes)
    for key in ("spectra", "ext_limits", "sky_model", "sky_uncertainty"):
        assert np.array_equal(batch[key][1], single[key][0], equal_nan=True), key
    assert np.array_equal(batch["rng_states"][1], single["rng_states"][0])


@pytest.mark.parametrize("cfg,order", [("C1", 2), ("T1", 1), ("T2", 3)])
def test_polynomial_sky_segmented_equals_literal(cfg, order):
    """sky_order > 0: the parallel form (32-bit stream, attempt bitmap + prefix, one CTA per column) and the literal
    one-CTA-per-frame replay of np.random.normal draw the same deviates, fit the same polynomials and leave the
    generator in the same state - bit for bit."""
    from motes_b200 import pipeline, synth
    frame, axes, header = synth.make_frame(synth.config_spec(cfg, seed=17))
    args = synth.default_args(sky_order=order)
    runs = {}
    for label, env in (("segmented", {"MOTES_B200_BOOTSTRAP": "segmented"}), ("literal", {"MOTES_B200_BOOTSTRAP": "classic"})):
        h = _fresh_handle(env)
        runs[label] = pipeline.run_frames(frame["data"][None].copy(), frame["errs"][None].copy(), frame["qual"][None], args,
                                          header["seeing"], header["pixel_resolution"], seeds=[99], axes_dict=axes, handle=h)
        h.close()
    a, b = runs["segmented"], runs["literal"]
    assert a["frame_status"][0] == 0 and b["frame_status"][0] == 0
    assert np.array_equal(a["rng_states"], b["rng_states"])
    for key in ("sky_model", "sky_uncertainty", "spectra", "ext_limits"):
        assert np.array_equal(a[key], b[key], equal_nan=True), key
    assert np.isfinite(a["sky_model"]).mean() > 0.9
