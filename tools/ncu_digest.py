#!/usr/bin/env python3
"""Digest of an .ncu-rep: per kernel the headline raw metrics, the SASS opcode mix and the hottest source lines.

usage: python tools/ncu_digest.py rep.ncu-rep [kernel-regex] [top-lines]
"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 12
sel = ["--kernel-name", "regex:" + kre] if kre else []

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        # which pipe: ALU (logic / shift / integer add: one warp instruction per two cycles), FMA (IMAD), XU (POPC),
        # LSU instructions, and the shared-memory / L1 data pipe (wavefronts)
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"


def run(args):
    return subprocess.run(["ncu", "-i", rep] + args, capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"] + sel))))
hdr, units = raw[0], raw[1]
for r in raw[2:]:
    print("==", r[hdr.index("Kernel Name")][:70], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for w in WANT:
        if w in hdr:
            print(f"   {w:70s} {r[hdr.index(w)][:24]:>24s} {units[hdr.index(w)]}")
    st = []
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(r[hdr.index(h)]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    st.sort(reverse=True)
    print("   stalls per issue:", ", ".join(f"{n} {v:.2f}" for v, n in st[:7]))

sass = list(csv.reader(io.StringIO(run(["--page", "source", "--csv", "--print-source", "sass"] + sel))))
ops, tot, h = collections.Counter(), 0, None
for r in sass:
    if not r:
        continue
    if "Instructions Executed" in r and "Source" in r:
        h = r
        continue
    if h is None:
        continue
    try:
        ex = int(r[h.index("Instructions Executed")]); src = r[h.index("Source")]
    except (ValueError, IndexError):
        continue
    parts = src.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    ops[op.split(".")[0]] += ex
    tot += ex
print("-- SASS mix (all selected launches):", ", ".join(f"{k} {100*v/max(tot,1):.1f}%" for k, v in ops.most_common(14)))

src = list(csv.reader(io.StringIO(run(["--page", "source", "--csv", "--print-source", "cuda,sass"] + sel))))
lines, fname, h = collections.defaultdict(lambda: [0, 0]), None, None
for r in src:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        h = r
        continue
    if h is None or r[0] in ("Function Name", "Kernel Name", ""):
        continue
    try:
        lines[(fname, r[0], r[1].strip()[:80])][0] += int(r[6]); lines[(fname, r[0], r[1].strip()[:80])][1] += int(r[7])
    except (ValueError, IndexError):
        pass
ts = sum(v[0] for v in lines.values()) or 1
ti = sum(v[1] for v in lines.values()) or 1
print("-- hottest source lines (share of samples / of instructions)")
for key, (s_, i_) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"   {100*s_/ts:5.1f}% {100*i_/ti:5.1f}%  {key[0]}:{key[1]}  {key[2]}")
