"""The reference staged under oracle/_ref (byte-compiled by oracle/build_ref.py) and the oracle port are the same
function: bit-identical outputs on seeded frames - which is what lets the GPU box use either as the checker - and the
oracle port reproduces the reference's full-size golden outputs of BASELINE.json's configurations."""
import copy
import os
import warnings

import numpy as np
import pytest

from conftest import ROOT, load_golden, golden_inputs
from oracle import motes_oracle as mo, ref_runner

KEYS = ("median_params", "scale", "sky_bins", "sky_params", "sky_limits", "sky_model", "sky_uncertainty", "ext_bins",
        "ext_params", "ext_limit_table", "ext_limits", "optimal", "optimal_err", "aperture", "aperture_err")


@pytest.mark.skipif(not ref_runner.available(), reason="reference neither staged under oracle/_ref nor present at /root/reference")
@pytest.mark.parametrize("config,over,seed", [("T2", {}, 5), ("T1", {"nan_fraction": 0.01}, 6)])
def test_staged_reference_equals_oracle_port(config, over, seed):
    from motes_b200 import synth
    frame, axes, header = synth.make_frame(synth.config_spec(config, seed=seed, **over))
    args = synth.default_args()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = ref_runner.run_frame(copy.deepcopy(frame), axes, dict(header), args, seed=seed)
        b = mo.run_frame(copy.deepcopy(frame), axes, dict(header), args, seed=seed)
    for key in KEYS:
        assert np.array_equal(a[key], b[key], equal_nan=True), key


def test_build_recipe_stages_only_bytecode():
    """oracle/_ref holds compiled modules, never reference source text (and is git-ignored)."""
    staged = os.path.join(ROOT, "oracle", "_ref", "motes")
    if os.path.isdir(staged):
        assert not [f for f in os.listdir(staged) if f.endswith((".py", ".pyc"))]
    assert "oracle/_ref/" in open(os.path.join(ROOT, ".gitignore")).read()


@pytest.mark.parametrize("name", ["c2_fors2_4096x400", "c5a_2048x256_binned"])
def test_oracle_port_reproduces_full_size_reference_outputs(name):
    """configs[1] (the shape the headline metric is quoted on) and configs[4]: bit for bit."""
    meta, gold = load_golden(name)
    frame, axes, header, args = golden_inputs(meta)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = mo.run_frame(frame, axes, dict(header), args, seed=meta["rng_seed"])
    for key in KEYS:
        if key in gold:
            assert np.array_equal(gold[key], out[key], equal_nan=True), key
