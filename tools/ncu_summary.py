#!/usr/bin/env python3
"""Summarise an ncu report (one kernel, `--set full --import-source on`) as text.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt

Prints the headline raw metrics and the share of warp-stall samples and executed
instructions per kernel phase.  Phases are delimited in the CUDA sources by
comment markers of the form `// @phase <name>`: a source line belongs to the
last marker above it in the same file.  Needs only the `ncu` CLI (no GPU).
"""
import csv
import io
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RAW_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]
STALLS = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio")


def ncu(*args):
    return subprocess.run(["ncu", *args], check=True, capture_output=True, text=True).stdout


def phase_table(path):
    """{file basename: sorted [(line, phase)]} from `// @phase` markers."""
    out = {}
    for dirpath, _, names in os.walk(os.path.join(ROOT, "qpmpc_b200", "csrc")):
        for name in names:
            marks = []
            with open(os.path.join(dirpath, name)) as f:
                for no, line in enumerate(f, 1):
                    m = re.search(r"//\s*@phase\s+(.+)", line)
                    if m:
                        marks.append((no, m.group(1).strip()))
            out[name] = marks
    return out


def phase_of(table, fname, line):
    name = os.path.basename(fname)
    cur = None
    for no, ph in table.get(name, []):
        if no <= line:
            cur = ph
    return cur or f"({name})"


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 25
    for a in sys.argv[2:]:
        if a.startswith("--batch="):  # instances solved by the profiled launch (for per-solve figures)
            print(f"instances_per_launch {int(a.split('=')[1])}")
    raw = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        d = dict(zip(hdr, row))
        print(f"== kernel: {d.get('Kernel Name')}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for h, u, v in zip(hdr, units, row):
            if h in RAW_KEYS:
                print(f"  {h:72s} {v} {u}")
        print("  warp stall reasons (average warps stalled per issue-active cycle):")
        st = sorted(((float(v), STALLS.match(h).group(1)) for h, v in zip(hdr, row)
                     if STALLS.match(h) and v), reverse=True)
        print("    " + ", ".join(f"{n} {x:.2f}" for x, n in st if x >= 0.01))
    src = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "source", "--csv",
                                          "--print-source", "cuda,sass"))))
    table = phase_table(ROOT)
    cur, lines = None, []
    for r in src:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1]
        elif len(r) > 7 and r[0].isdigit() and r[2] == "-":
            lines.append((cur, int(r[0]), int(r[6] or 0), int(r[7] or 0), r[1].strip()))
    # executed floating-point instructions of the launch, by opcode (SASS rows of the source page):
    # the flops the kernel really executes, counted by the profiler instead of a model
    ops = defaultdict(int)
    for r in src:
        if len(r) > 7 and r[0] == "" and r[2] not in ("", "-"):
            try:
                e = int(r[7])
            except ValueError:
                continue
            op = re.sub(r"^@!?U?P\d+\s+", "", r[3].strip()).split()[0].split(".")[0] if r[3].strip() else "?"
            ops[op] += e
    f64 = {k: ops.get(k, 0) for k in ("DFMA", "DMUL", "DADD")}
    f32 = {k: ops.get(k, 0) for k in ("FFMA", "FMUL", "FADD")}
    print(f"  fp64 warp instructions executed: " + ", ".join(f"{k} {v}" for k, v in f64.items())
          + f"  => flops_fp64_per_launch {32 * (2 * f64['DFMA'] + f64['DMUL'] + f64['DADD'])}")
    print(f"  fp32 warp instructions executed: " + ", ".join(f"{k} {v}" for k, v in f32.items())
          + f"  => flops_fp32_per_launch {32 * (2 * f32['FFMA'] + f32['FMUL'] + f32['FADD'])}")
    ts = sum(x[2] for x in lines) or 1
    ti = sum(x[3] for x in lines) or 1
    agg = defaultdict(lambda: [0, 0])
    for f, ln, s, i, _ in lines:
        a = agg[phase_of(table, f, ln)]
        a[0] += s
        a[1] += i
    print(f"\n== per phase (first launch in the report): {ts} stall samples, {ti} warp instructions")
    for ph, (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"  {ph:34s} samples {100 * s / ts:5.1f}%   instructions {100 * i / ti:5.1f}%")
    print(f"\n== top {top} source lines by stall samples")
    for f, ln, s, i, text in sorted(lines, key=lambda x: -x[2])[:top]:
        print(f"  {os.path.basename(f)}:{ln:<4d} s {100 * s / ts:4.1f}%  i {100 * i / ti:4.1f}%  {text[:88]}")


if __name__ == "__main__":
    main()
