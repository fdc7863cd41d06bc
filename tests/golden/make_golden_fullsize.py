#!/usr/bin/env python3
"""Golden vectors at the FULL sizes of BASELINE.json's configurations, from the UNMODIFIED reference.

    python tests/golden/make_golden_fullsize.py [case ...]        # build container only

Same recipe as ``make_golden.py`` (reference imported from /root/reference/src, stages driven in ``core.motes``
order by ``oracle/ref_runner.run_stages``, seeded bit-reproducible ``motes_b200.synth`` inputs, a second run on
inputs moved by +-1 ulp as the reference's self-noise control), but only the small per-column / per-bin outputs are
stored - the frames are regenerated from the spec and checked against the stored sha256:

    c2_fors2_4096x400        configs[1]: the shape the headline metric is quoted on, default flags
    c3_xshooter_24000x80_cr  configs[2]: 1 % of the pixels hit by cosmic rays
    c5a_2048x256_binned      configs[4]: faint source, default binning (-m 15 -xs 10 -ks 100)
    c5a_2048x256_unbinned    configs[4]: binning off (-m 0 -xs 0 -ks 0): every column its own bin
    c5b_16384x512_binned     configs[4]: the large end of the sweep

Each case runs in its own process (they are independent and the reference is single-threaded).
"""

from __future__ import annotations

import copy
import json
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

UNBINNED = {"minimum_column_limit": 0, "extraction_snr_limit": 0, "sky_snr_limit": 0}
CASES = {
    # name: (config, spec overrides, args overrides, rng seed)
    "c2_fors2_4096x400": ("C2", {}, {}, 2025),
    "c3_xshooter_24000x80_cr": ("C3", {}, {}, 2026),
    "c5a_2048x256_binned": ("C5a", {}, {}, 2027),
    "c5a_2048x256_unbinned": ("C5a", {"amplitude": 900.0}, UNBINNED, 2028),
    "c5b_16384x512_binned": ("C5b", {}, {}, 2029),
}
KEEP = ("median_params", "median_fitinfo", "scale", "seeing_px", "bin_rows", "sky_bins", "sky_fitinfo", "sky_params",
        "sky_limits", "sky_model", "sky_uncertainty", "ext_bins", "ext_fitinfo", "ext_params", "ext_limit_table",
        "ext_limits", "optimal", "optimal_err", "aperture", "aperture_err")
CONTROL = ("median_params", "sky_params", "ext_params", "optimal", "optimal_err", "sky_bins", "ext_bins")


def make_case(name):
    from motes_b200 import synth
    from oracle import ref_runner
    from make_golden import perturb_one_ulp
    config, spec_over, args_over, seed = CASES[name]
    ref = ref_runner.load(prefer_staged=False)
    recorder = ref_runner.FitRecorder(ref.tracing)
    spec = synth.config_spec(config, seed=seed, **spec_over)
    args = synth.default_args(**args_over)
    frame, axes, header = synth.make_frame(spec)
    digest = synth.frame_digest(frame)
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = ref_runner.run_stages(ref.tracing, ref.sky, ref.extraction, copy.deepcopy(frame), axes, dict(header), args,
                                    seed=seed, recorder=recorder, keep_frames=False)
        seconds = time.perf_counter() - t0
        noisy = perturb_one_ulp(copy.deepcopy(frame), seed + 1000)
        ctl = ref_runner.run_stages(ref.tracing, ref.sky, ref.extraction, noisy, axes, dict(header), args, seed=seed,
                                    recorder=recorder, keep_frames=False)
    keep = {k: out[k] for k in KEEP if k in out}
    for key in CONTROL:
        if key in ctl:
            keep["ctl_" + key] = ctl[key]
    meta = dict(case=name, config=config, spec=synth.spec_to_dict(spec), args=vars(args), rng_seed=seed,
                input_sha256=digest, input_dtype=None, reference_seconds=seconds, reference_origin=ref.origin,
                versions=dict(numpy=np.__version__, scipy=__import__("scipy").__version__))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, meta=json.dumps(meta), **keep)
    return (f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, reference {seconds:.1f} s, sky_bins={len(out.get('sky_bins', []))} "
            f"ext_bins={len(out['ext_bins'])}")


def main():
    import multiprocessing as mp
    names = [n for n in CASES if not sys.argv[1:] or n in sys.argv[1:]]
    with mp.get_context("spawn").Pool(min(len(names), os.cpu_count() or 1)) as pool:
        for line in pool.imap_unordered(make_case, names):
            print(line, flush=True)


if __name__ == "__main__":
    main()
