import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_golden(name):
    import json
    import numpy as np
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    out = {k: z[k] for k in z.files if k != "meta"}
    meta = json.loads(str(z["meta"]))
    return meta, out


def golden_inputs(meta):
    """Regenerate the synthetic frame a golden file was made from and check its digest."""
    from types import SimpleNamespace
    from motes_b200 import synth
    spec = dict(meta["spec"])
    spec["chip_gaps"] = tuple(tuple(g) for g in spec["chip_gaps"])
    frame, axes, header = synth.make_frame(synth.FrameSpec(**spec))
    assert synth.frame_digest(frame) == meta["input_sha256"], "synthetic generator is not bit-reproducible here"
    if meta.get("input_dtype"):          # e.g. ">f4": the dtype the GMOS harvester hands over (gmosio.py:40-52)
        frame["data"] = frame["data"].astype(meta["input_dtype"])
        frame["errs"] = frame["errs"].astype(meta["input_dtype"])
    return frame, axes, header, SimpleNamespace(**meta["args"])


def pytest_collection_modifyitems(config, items):
    """Tests marked gpu are skipped (not failed) on a machine without a CUDA device."""
    if not any("gpu" in item.keywords for item in items):
        return
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device (run with -m gpu on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
