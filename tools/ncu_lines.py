#!/usr/bin/env python3
"""Summarise `ncu --page source --csv --print-source cuda,sass` output: hottest source lines per kernel instance.

usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass [--kernel-name regex:...] > src.csv
       python tools/ncu_lines.py src.csv [top]
"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
instances = []          # list of dict[(file, line, text)] -> [samples, instr]
cur = None
fname = None
first_file = None
with open(path, newline="") as fh:
    for row in csv.reader(fh):
        if not row:
            continue
        if row[0] == "File Path":
            fname = row[1].split("/")[-1]
            if first_file is None:
                first_file = fname
            if cur is None or fname == first_file and seen_files and fname in seen_files:
                cur = defaultdict(lambda: [0, 0])
                instances.append(cur)
                seen_files = set()
            seen_files.add(fname)
            continue
        if row[0] in ("Function Name", "Line No", "Kernel Name"):
            if row[0] == "Function Name":
                kname = row[1]
            continue
        if row[0] == "":
            continue
        try:
            samples = int(row[6])
            instr = int(row[7])
        except (ValueError, IndexError):
            continue
        key = (fname, row[0], row[1].strip()[:90])
        cur[key][0] += samples
        cur[key][1] += instr

for k, inst in enumerate(instances):
    tot_s = sum(v[0] for v in inst.values()) or 1
    tot_i = sum(v[1] for v in inst.values()) or 1
    print(f"=== instance {k}: samples {tot_s} instructions {tot_i}")
    for key, (s, i) in sorted(inst.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {100*s/tot_s:5.1f}% smp {100*i/tot_i:5.1f}% ins  {key[0]}:{key[1]}  {key[2]}")
