#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).

    python tools/summarize_launches.py gpurun_out/r01b_launches.csv "command line" > profiles/r01b_launches_summary.md
"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    m = re.match(r"([A-Za-z0-9_:]+)", name)
    base = m.group(1) if m else name
    return base.split("::")[-1] if not base.startswith("cub") else base.split("::")[-1]


def main():
    path = sys.argv[1]
    cmd = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        rows.append((short(r["Kernel Name"]), v * scale))
    tot = sum(v for _, v in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for k, v in rows:
        agg[k][0] += 1
        agg[k][1] += v
    print("# ncu launch list (gpu__time_duration.sum, --clock-control none)\n")
    if cmd:
        print("Command: `%s`\n" % cmd)
    print("%d launches, %.3f ms in total (serialised, cold cache: compare SHARES, not absolute times)\n" % (len(rows), tot))
    print("| kernel | launches | total ms | share |")
    print("|---|---:|---:|---:|")
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("| %s | %d | %.3f | %.1f%% |" % (k, n, v, 100 * v / tot))


if __name__ == "__main__":
    main()
