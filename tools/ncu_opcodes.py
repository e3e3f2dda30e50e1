#!/usr/bin/env python3
"""Executed-instruction opcode histogram per kernel phase from an ncu report
(`--import-source on`, built with -lineinfo).  Phases as in ncu_summary.py.

    python tools/ncu_opcodes.py gpurun_out/prof.ncu-rep [warps]
"""
import csv
import io
import os
import re
import sys
from collections import Counter, defaultdict

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_summary import ROOT, ncu, phase_of, phase_table

rep = sys.argv[1]
warps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
table = phase_table(ROOT)
cur_file, cur_phase = None, None
per = defaultdict(Counter)
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1]
    elif len(r) > 7 and r[0].isdigit() and r[2] == "-":
        cur_phase = phase_of(table, cur_file, int(r[0]))
    elif len(r) > 7 and r[0] == "" and r[2] not in ("", "-"):
        try:
            e = int(r[7])
        except ValueError:
            continue
        src = re.sub(r"^@!?U?P\d+\s+", "", r[3].strip())
        op = src.split()[0] if src else "?"
        if op.startswith("IMAD.MOV") or op.startswith("MOV") or op.startswith("CS2R") or op.startswith("UMOV"):
            op = "(move)"
        else:
            op = op.split(".")[0]
        per[cur_phase][op] += e
tot = sum(sum(c.values()) for c in per.values())
print(f"total executed warp instructions {tot} ({tot / warps:.0f} per warp)")
for ph, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values())):
    t = sum(c.values())
    top = ", ".join(f"{op} {v / warps:.0f}" for op, v in c.most_common(9))
    print(f"{ph:34s} {100 * t / tot:5.1f}%  {t / warps:7.0f}/warp   {top}")
