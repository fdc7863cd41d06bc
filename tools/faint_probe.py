"""Per-bin deviation of the fitted Moffat parameters from the golden reference on the faint-source case."""
import copy, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, golden_inputs
from parity import param_errors
from motes_b200 import pipeline
name = sys.argv[1] if len(sys.argv) > 1 else "t1_faint"
meta, g = load_golden(name)
frame, axes, header, args = golden_inputs(meta)
out = pipeline.run_frame(copy.deepcopy(frame), axes, dict(header), args, seed=meta["rng_seed"])
np.set_printoptions(linewidth=200, precision=3)
for key in ("sky_params", "ext_params"):
    err = param_errors(out[key], g[key], frame["data"].shape[0])
    ctl = param_errors(g["ctl_" + key][:, :6], g[key][:, :6], frame["data"].shape[0])
    info = g[key.replace("params", "fitinfo")]
    print(key, "max", err.max(axis=0), "ctl max", ctl.max(axis=0))
    worst = np.argsort(-err[:, 3])[:4]
    for b in worst:
        print("  bin", b, "err", err[b], "ctl", ctl[b], "ref nfev/njev/status/cost", info[b])
