#!/usr/bin/env python
"""Warp instructions per source line of an `ncu --import-source on` report, normalised by the execution count of one
reference line (e.g. the line that runs once per scan step).

    python tools/ncu_steps.py report.ncu-rep <reference line> [min per-step]
"""
import csv, io, subprocess, sys
rep, ref = sys.argv[1], int(sys.argv[2])
floor = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
agg = {}
for r in rows:
    if len(r) > 10 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    d = dict(zip(hdr[4:], r[4:]))
    try:
        inst, smp, ln = int(d["Instructions Executed"]), int(d["# Samples"]), int(r[0])
    except (ValueError, KeyError):
        continue
    a = agg.setdefault(ln, [0, 0, r[1].strip()[:100]])
    a[0] += inst
    a[1] += smp
steps = agg[ref][0]
print("reference line %d executed %d warp instructions" % (ref, steps))
tot = 0
for ln in sorted(agg):
    i, s_, src = agg[ln]
    tot += i
    if i / steps >= floor:
        print("%4d %7.1f %7d  %s" % (ln, i / steps, s_, src))
print("total per reference execution: %.1f" % (tot / steps))
