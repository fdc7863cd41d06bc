#!/usr/bin/env python3
"""profiles/ncu_traffic.json from the launch list of one bench step (tools/round_artifacts.sh):
per bench stage the DRAM bytes and the FP64 operations (2 x DFMA + DADD + DMUL, thread level, predicated on) of ONE
stage launch - a stage is what bench.py times with CUDA events (e.g. bin_fits = every kernel of one stage_fit call).

    python tools/ncu_traffic.py gpurun_out/r02_launches.csv [--util gpurun_out/r02_full_fit.ncu-rep ...]
"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE_OF = [  # kernel-name pattern -> bench stage
    (r"fit_init|fit_sweep|fit_step|fit_finish|moffat_fit", "fits"),
    (r"sky_prepare", "sky_prepare"), (r"mt_stream", "boot_stream"), (r"mt_jump|seg_plan", "boot_jump"),
    (r"sky_walk|walk_", "boot_walk"), (r"sky_column", "boot_column"), (r"row_median", "row_median"),
    (r"optimal_extract", "extract"), (r"bin_profile", "bin_profiles"), (r"bins_kernel|qual_columns", "snr_bins"),
    (r"column_stats", "column_stats"), (r"sky_apply", "sky_apply"), (r"limits_kernel|bin_post", "limits"),
]


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mn, mv, idc = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        try:
            val = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        launches.setdefault(r[idc], {"name": r[kn]})[r[mn]] = val
    per_stage = collections.defaultdict(lambda: collections.Counter())
    fit_calls = 0
    for rec in launches.values():
        name = rec["name"]
        stage = next((s for pat, s in STAGE_OF if re.search(pat, name)), None)
        if stage is None:
            continue
        if "fit_init" in name or "moffat_fit_group" in name:
            fit_calls += 1
        c = per_stage[stage]
        c["ns"] += rec.get("gpu__time_duration.sum", 0.0)
        c["dram"] += rec.get("dram__bytes_read.sum", 0.0) + rec.get("dram__bytes_write.sum", 0.0)
        c["flop"] += 2 * rec.get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0.0) + \
            rec.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0.0) + \
            rec.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0.0)
        c["launches"] += 1
    # the fit kernels of one step: the median fit (one small call) + the sky bins + the extraction bins; bench.py's
    # bin_fits stage is the two big calls - attribute everything but the first (median) call's share by time
    fits = per_stage.pop("fits", None)
    out = {"workload": {"frames_per_gpu": 128, "nspat": 400, "ndisp": 4096},
           "source": f"{os.path.basename(path)}: ncu --metrics gpu__time_duration.sum,dram__bytes_*.sum,"
                     "smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul}_pred_on.sum over one bench step (MOTES_B200_HOIST=0), summed per stage",
           "dram_bytes_per_launch": {}, "fp64_flop_per_launch": {}, "kernel_ms_under_ncu": {}, "kernels_per_step": {}}
    for stage, c in per_stage.items():
        out["dram_bytes_per_launch"][stage] = c["dram"]
        out["kernel_ms_under_ncu"][stage] = c["ns"] / 1e6
        out["kernels_per_step"][stage] = c["launches"]
    if fits:
        # bin_fits is launched twice per step (sky bins, extraction bins); the median fit is < 1 % of the fit work
        out["dram_bytes_per_launch"]["bin_fits"] = fits["dram"] / 2
        out["fp64_flop_per_launch"]["bin_fits"] = fits["flop"] / 2
        out["kernel_ms_under_ncu"]["bin_fits"] = fits["ns"] / 1e6 / 2
        out["kernels_per_step"]["bin_fits"] = fits["launches"]
        out["fp64_flop_source"] = "2 x DFMA + DADD + DMUL (thread level, predicated on) of every fit kernel of one step / 2 stage launches"
    util = {}
    for rep in sys.argv[2:]:
        if rep.startswith("--"):
            continue
        import subprocess
        raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
        if len(raw) < 3:
            continue
        hh = raw[0]
        for r in raw[2:]:
            name = r[hh.index("Kernel Name")]
            stage = next((s for pat, s in STAGE_OF if re.search(pat, name)), None)
            key = "bin_fits" if stage == "fits" and "sweep" in name else ("fit_step" if stage == "fits" else stage)
            if key is None or key in util:
                continue
            def g(m):
                return float(r[hh.index(m)]) if m in hh else None
            util[key] = {"kernel": name.split("(")[0][-60:],
                         "issue_slot_pct": g("sm__inst_issued.avg.pct_of_peak_sustained_active"),
                         "fp64_pipe_pct": g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                         "warps_per_sm": g("sm__warps_active.avg.per_cycle_active")}
    if util:
        out["ncu_utilisation"] = util
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1)[:3000])


if __name__ == "__main__":
    main()
