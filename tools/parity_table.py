#!/usr/bin/env python3
"""Collect the parity reports of tests/test_gpu_configs.py (gpurun_out/parity_*.json) into profiles/r02_parity.json:
per configuration the Moffat-parameter errors (median / p99 / max per parameter, count of fits beyond 1e-6) beside the
reference's own +-1-ulp self-noise, the integer flips and the flux error - once for the shipped build (forward
differences from the base pow, trfb.cuh: shifted_unit) and once for the -DMOTES_FIT_REAL_POW build (prefix realpow_).

    bash tools/parity_runs.sh          # under gpurun: produces both sets of gpurun_out/parity_*.json
    python tools/parity_table.py       # here: writes profiles/r02_parity.json and prints the table
"""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ("A", "c", "alpha", "beta", "B", "m")


def load(prefix):
    out = {}
    for path in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"parity_{prefix}c*.json"))):
        case = os.path.basename(path)[len("parity_" + prefix):-5]
        out[case] = json.load(open(path))
    return out


def slim(rep):
    keep = {k: rep[k] for k in ("bins_exact", "sky_exact", "limit_flips", "sky_limit_flips", "flux_max_rel", "median_params_max",
                                "columns", "fits", "fits_beyond_1e-6", "param_p99", "param_max") if k in rep}
    for key in ("sky_params", "ext_params"):
        if key in rep:
            block = {k: rep[key][k] for k in ("fits", "median", "p99", "max", "fits_beyond_1e-6")}
            if "reference_self_noise" in rep[key]:
                block["reference_self_noise"] = {k: rep[key]["reference_self_noise"][k] for k in ("median", "p99", "max", "fits_beyond_1e-6")}
            keep[key] = block
    return keep


def main():
    shipped, realpow = load(""), load("realpow_")
    shipped = {k: v for k, v in shipped.items() if not k.startswith("realpow_")}
    table = {"parameters": NAMES,
             "how": "tests/test_gpu_configs.py on B200 against tests/golden/*.npz (outputs of the unmodified reference); "
                    "errors as oracle/parity_report.py:param_errors; reference_self_noise = the reference against itself on "
                    "inputs moved by +-1 ulp (golden ctl_* arrays)",
             "shipped_build": {k: slim(v) for k, v in shipped.items()},
             "real_pow_build": {k: slim(v) for k, v in realpow.items()}}
    traj = os.path.join(ROOT, "gpurun_out", "parity_fit_trajectories.json")
    if os.path.exists(traj):
        table["fit_trajectories_c2"] = json.load(open(traj))
    with open(os.path.join(ROOT, "profiles", "r02_parity.json"), "w") as fh:
        json.dump(table, fh, indent=1)
    print(f"{'case':28s} {'build':8s} {'fits':>5s} {'>1e-6':>6s} {'ref>1e-6':>8s} {'p99(alpha,beta)':>22s} {'max(alpha,beta)':>22s} {'ref max':>20s} flips sky flux")
    for case in shipped:
        for label, src in (("shipped", shipped), ("realpow", realpow)):
            if case not in src:
                continue
            r = src[case]
            for key in ("sky_params", "ext_params"):
                if key not in r:
                    continue
                b = r[key]
                c = b.get("reference_self_noise", {})
                print(f"{case[:22] + ' ' + key[:3]:28s} {label:8s} {b['fits']:5d} {b['fits_beyond_1e-6']:6d} {c.get('fits_beyond_1e-6', -1):8d} "
                      f"{b['p99'][2]:10.1e} {b['p99'][3]:10.1e}  {b['max'][2]:10.1e} {b['max'][3]:10.1e}  "
                      f"{c.get('max', [0] * 6)[2]:9.1e} {c.get('max', [0] * 6)[3]:9.1e}  {r['limit_flips']:4d} {str(r.get('sky_exact'))[:1]} {r['flux_max_rel']:.1e}")


if __name__ == "__main__":
    main()
