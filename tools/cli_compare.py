#!/usr/bin/env python
"""The CLI against the reference CLI on the SAME files, wall clock including the database load (not the bench: a record
kept under profiles/).

    python tools/cli_compare.py [--queries 256] [--db-seqs 1000000] [--threads N] [--cards 0]

Writes q.fa / d.fa with bench.py's host generator, packs d.s4gdb, then runs
    oracle/_ref/sift4g_ref      -q q.fa -d d.fa     -t <threads> --out ref/   (the unmodified reference CPU build)
    sift4g_b200/bin/sift4g_b200 -q q.fa -d d.fa     -t <threads> --out fa/    (FASTA parsed every run)
    sift4g_b200/bin/sift4g_b200 -q q.fa -d d.s4gdb  -t <threads> --out pk/    (packed database)
and prints one JSON line: seconds of each run, the stage banners' split, and whether every output file has the
reference's bytes.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import oracle as O  # noqa: E402


def run(cmd, env=None):
    """-> (seconds, {banner: seconds since start}) ; the CLI prints its stage banners ("** ... **") on stderr"""
    t0 = time.time()
    p = subprocess.Popen(cmd, stderr=subprocess.PIPE, stdout=subprocess.DEVNULL, text=True, env=dict(os.environ, **(env or {})))
    marks = []
    for line in p.stderr:
        if line.startswith("** "):
            marks.append((line.strip(" *\n"), time.time() - t0))
    p.wait()
    dt = time.time() - t0
    if p.returncode != 0:
        raise SystemExit("%s failed (%d)" % (cmd[0], p.returncode))
    stages = {}
    for i, (name, t) in enumerate(marks):
        stages[name] = round((marks[i + 1][1] if i + 1 < len(marks) else dt) - t, 3)
    return round(dt, 3), stages


def digest(d):
    return {f: hashlib.sha256(open(os.path.join(d, f), "rb").read()).hexdigest() for f in sorted(os.listdir(d))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=256)
    ap.add_argument("--db-seqs", type=int, default=1_000_000)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--cards", default="")
    ap.add_argument("--skip-reference", action="store_true")
    args = ap.parse_args()
    tmp = tempfile.mkdtemp()
    qf, df = bench.reference_sample(tmp, args.queries, args.db_seqs)
    ours = os.path.join(ROOT, "sift4g_b200", "bin", "sift4g_b200")
    packed = os.path.join(tmp, "d.s4gdb")
    t0 = time.time()
    subprocess.run([os.path.join(ROOT, "sift4g_b200", "bin", "s4g_pack"), df, packed], check=True)
    pack_s = round(time.time() - t0, 3)
    out = {"queries": args.queries, "db_seqs": args.db_seqs, "fasta_mb": round(os.path.getsize(df) / 1e6, 1), "threads": args.threads,
           "host_cores": os.cpu_count(), "pack_s": pack_s}
    flags = ["-t", str(args.threads)] + (["--cards", args.cards] if args.cards else [])
    dirs = {}
    for key, db in (("ours_fasta", df), ("ours_packed", packed)):
        dirs[key] = os.path.join(tmp, key)
        os.mkdir(dirs[key])
        run([ours, "-q", qf, "-d", db, "--out", dirs[key]] + flags)                       # warm the page cache / CUDA context caches once
        for f in os.listdir(dirs[key]):
            os.remove(os.path.join(dirs[key], f))
        out[key + "_s"], out[key + "_stages_s"] = run([ours, "-q", qf, "-d", db, "--out", dirs[key]] + flags)
    if not args.skip_reference and os.path.exists(O.REF_SIFT4G):
        dirs["reference"] = os.path.join(tmp, "reference")
        os.mkdir(dirs["reference"])
        out["reference_s"], out["reference_stages_s"] = run([O.REF_SIFT4G, "-q", qf, "-d", df, "--out", dirs["reference"], "-t", str(args.threads)])
        ref = digest(dirs["reference"])
        for key in ("ours_fasta", "ours_packed"):
            got = digest(dirs[key])
            out[key + "_files_equal_reference"] = sum(1 for f in ref if got.get(f) == ref[f])
        out["files"] = len(ref)
        out["speedup_whole_cli"] = round(out["reference_s"] / out["ours_packed_s"], 2)
        hot = lambda st: sum(v for k, v in st.items() if k.startswith("Searching database") or k.startswith("Aligning queries"))
        out["speedup_search_plus_align"] = round(hot(out["reference_stages_s"]) / max(hot(out["ours_packed_stages_s"]), 1e-9), 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
