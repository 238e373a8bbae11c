#!/usr/bin/env python
"""Dynamic SASS opcode mix of the kernels in an ncu report (needs --import-source on / --set full).
    python tools/ncu_opmix.py gpurun_out/x.ncu-rep [top_n]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
cur = None; hdr = None
agg = collections.defaultdict(collections.Counter)
for r in csv.reader(io.StringIO(out)):
    if len(r) >= 2 and r[0] == "Kernel Name": cur = r[1]; continue
    if len(r) > 5 and r[0] == "Address": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    try: n = int(d["Instructions Executed"])
    except (KeyError, ValueError): continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", d["Source"].strip())
    op = ".".join((m.group(2) if m else d["Source"]).split(".")[:3])
    agg[cur][op] += n
for k, c in agg.items():
    tot = sum(c.values())
    print("%s: %d warp instructions" % (k.split("(")[0], tot))
    for op, n in c.most_common(top): print("   %-28s %6.2f%%" % (op, 100.0 * n / tot))
