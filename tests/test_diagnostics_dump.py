"""Host-side diagnostics data (the numbers behind the reference's -df / -dl / -dk plots) against the reference's own
intermediate values, captured by wrapping its plotting hooks (the plot functions themselves need matplotlib)."""
import copy
import warnings

import numpy as np
import pytest

from oracle import ref_runner

pytestmark = pytest.mark.skipif(not ref_runner.available(), reason="reference not staged")


def _import_core_helpers():
    # motes_b200.core imports the CUDA-backed modules; the helpers tested here are numpy only, so load them without it
    import ast, os, types
    from conftest import ROOT
    src = open(os.path.join(ROOT, "motes_b200", "core.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("localisation_lines", "sky_fit_diagnostics")]
    mod = types.ModuleType("core_helpers")
    mod.np = np
    exec(compile(ast.Module(body=keep, type_ignores=[]), "core_helpers", "exec"), mod.__dict__)
    return mod


@pytest.mark.parametrize("order", [0, 2])
def test_sky_fit_dump_matches_what_the_reference_would_plot(order):
    from motes_b200 import synth
    core = _import_core_helpers()
    ref = ref_runner.load()
    frame, axes, header = synth.make_frame(synth.config_spec("T2", seed=21, sky_slope=0.1 if order else 0.0))
    args = synth.default_args(sky_order=order, diag_plot_skyfit=True)
    seen = {}

    def capture(good_axis, good_pixels, good_errs, column_axis, model, model_err, col, file_name, flux_unit):
        seen[col] = (np.array(good_axis), np.array(good_pixels), np.array(good_errs), np.array(column_axis), np.array(model), np.array(model_err))
    old = ref.sky.diagnostics.plot_sky_fit
    ref.sky.diagnostics.plot_sky_fit = capture
    try:
        before = copy.deepcopy(frame)
        if order == 0:
            args.diag_plot_skyfit = False          # the reference only plots polynomial fits (sky.py:361-371)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = ref_runner.run_stages(ref.tracing, ref.sky, ref.extraction, frame, axes, dict(header), args, seed=3)
    finally:
        ref.sky.diagnostics.plot_sky_fit = old
    cols = [3, 100, 399] if order == 0 else sorted(seen)[:3] + sorted(seen)[-2:]
    nspat = before["data"].shape[0]
    model = np.tile(out["sky_model"], (nspat, 1)) if order == 0 else out["sky_model"]
    unc = np.tile(out["sky_uncertainty"], (nspat, 1)) if order == 0 else out["sky_uncertainty"]
    got = core.sky_fit_diagnostics(before, out["sky_limits"], model, unc, axes, cols)
    for c in cols:
        g = got[c]
        assert g["modelled"]
        if order > 0:
            want = seen[c]
            assert np.array_equal(g["good_sky_axis"], want[0]) and np.array_equal(g["good_sky_pixels"], want[1])
            assert np.array_equal(g["good_sky_errs"], want[2]) and np.array_equal(g["column_axis"], want[3])
            assert np.array_equal(g["column_sky_model"], want[4]) and np.array_equal(g["column_sky_model_err"], want[5])
        else:
            assert np.all(g["column_sky_model"] == out["sky_model"][c])
            assert len(g["good_sky_pixels"]) > 10


def test_localisation_lines_match_get_draw_lines():
    from motes_b200 import synth
    core = _import_core_helpers()
    ref = ref_runner.load()
    axes = synth.make_axes(64, 400, wavelength_start=17, spatial_floor=5)
    table = np.array([[8.0, 24.0, 40.0], [20.0, 21.0, 22.0], [40.0, 41.0, 42.0], [30.0, 31.0, 32.0]])
    pair = [np.linspace(20, 22, 400), np.linspace(40, 42, 400)]
    import motes.diagnostics as diagnostics
    for b in (table, pair):
        want = diagnostics.get_draw_lines(axes, b)
        got = core.localisation_lines(axes, b)
        assert len(got) == len(want) == 4
        for a, w in zip(got, want):
            assert np.array_equal(np.asarray(a), np.asarray(w))
