#!/usr/bin/env python
"""Per-source-line hot spots of an `ncu --set full --import-source on` report (kernels built with -lineinfo).

    python tools/ncu_lines.py gpurun_out/x.ncu-rep [top_n] [kernel substring]

Prints, for the source lines with the most stall samples: samples, warp instructions executed, average active
threads and the dominant stall reasons.
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    want = sys.argv[3] if len(sys.argv) > 3 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    func = ""
    lines = []
    tot_s = tot_i = 0
    for r in rows:
        if len(r) >= 2 and r[0] == "Function Name":
            func = r[1]
            continue
        if len(r) > 10 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[0] == "" or (want and want not in func):
            continue
        d = dict(zip(hdr[4:], r[4:]))
        try:
            samples = int(d["# Samples"]); inst = int(d["Instructions Executed"])
        except (ValueError, KeyError):
            continue
        stalls = sorted(((int(v), k) for k, v in d.items() if k.startswith("stall_") and "(Not Issued)" not in k and v.isdigit() and int(v) > 0), reverse=True)[:3]
        lines.append((samples, inst, d.get("Avg. Threads Executed", ""), r[0], r[1].strip()[:110], " ".join("%s=%d" % (k[6:], v) for v, k in stalls), func))
        tot_s += samples; tot_i += inst
    print("total samples %d, warp instructions %d" % (tot_s, tot_i))
    print("%7s %6s %12s %5s  %5s  %s" % ("samples", "%", "warp-inst", "thr", "line", "source | top stalls"))
    for s, i, thr, ln, src, st, fn in sorted(lines, reverse=True)[:top]:
        print("%7d %5.1f%% %12d %5s  %5s  %s | %s" % (s, 100.0 * s / max(tot_s, 1), i, thr, ln, src, st))


if __name__ == "__main__":
    main()
