"""Generates the committed golden fixtures from the UNMODIFIED reference (oracle/_ref, built by
oracle/Makefile from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

Outputs (all under tests/golden/):
  test_files/                 the reference's own end-to-end fixture (data files, copied verbatim)
  expected_test_files/        what the reference CPU build writes for it (--subst + --sub-results)
  expected_hashes.json        sha256 of every expected output, incl. the no---subst run
  synth_e2e/{q.fa,d.fa}       seeded synthetic end-to-end case with planted homologs
  seams.json                  seam dumps of the reference on synth_e2e: BLOSUM62 table, candidate ids
                              (-t 1), swimd scores, E-values/selection and SSW / swAlign paths
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O          # noqa: E402
from sift4g_b200 import synth           # noqa: E402

REF = "/root/reference"


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def run_ref(args, out_dir):
    os.makedirs(out_dir, exist_ok=True)
    subprocess.run([O.REF_SIFT4G] + args + ["--out", out_dir], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return {f: sha(os.path.join(out_dir, f)) for f in sorted(os.listdir(out_dir))}


def main():
    O.build()
    assert O.have_ref(), "oracle/_ref missing"
    hashes = {}
    # ---- 1. the reference's own fixture ----
    tf = os.path.join(HERE, "test_files")
    os.makedirs(tf, exist_ok=True)
    for f in ("query.fasta", "sample_protein_database.fa", "LACI_ECOLI.subst", "PURR_SALTY.subst"):
        shutil.copy(os.path.join(REF, "test_files", f), os.path.join(tf, f))
    exp = os.path.join(HERE, "expected_test_files")
    shutil.rmtree(exp, ignore_errors=True)
    hashes["test_files_subst"] = run_ref(["-q", tf + "/query.fasta", "-d", tf + "/sample_protein_database.fa", "--subst", tf + "/", "--sub-results"], exp)
    tmp = tempfile.mkdtemp()
    hashes["test_files_nosubst"] = run_ref(["-q", tf + "/query.fasta", "-d", tf + "/sample_protein_database.fa"], tmp + "/a")

    # ---- 2. synthetic end-to-end case ----
    sd = os.path.join(HERE, "synth_e2e")
    os.makedirs(sd, exist_ok=True)
    queries, db = synth.make_dataset(20240001, 6, 1500, q_len=(60, 420), homologs=(8, 30), rare_fraction=0.004)
    synth.write_fasta(sd + "/q.fa", queries, "QRY")
    synth.write_fasta(sd + "/d.fa", db, "DBS")
    hashes["synth_default"] = run_ref(["-q", sd + "/q.fa", "-d", sd + "/d.fa", "--sub-results"], tmp + "/b")
    hashes["synth_C200_M50"] = run_ref(["-q", sd + "/q.fa", "-d", sd + "/d.fa", "--sub-results", "--max-candidates", "200", "--max-aligns", "50"], tmp + "/c")
    json.dump(hashes, open(os.path.join(HERE, "expected_hashes.json"), "w"), indent=1, sort_keys=True)

    # ---- 3. seam dumps on the synthetic case ----
    seams = {}
    seams["blosum62"] = None
    N = 120
    out = subprocess.run([O.REF_DUMP, "candidates", sd + "/q.fa", sd + "/d.fa", "5", str(N), "1"], capture_output=True, text=True, check=True).stdout.split("\n")
    seams["cells"] = int(out[0].split()[1])
    seams["max_candidates"] = N
    seams["candidates_t1"] = [list(map(int, l.split()[1:])) for l in out[1:1 + len(queries)]]
    with open(tmp + "/c.txt", "w") as f:
        for c in seams["candidates_t1"]:
            f.write("%d %s\n" % (len(c), " ".join(map(str, c))))
    out = subprocess.run([O.REF_DUMP, "scores", sd + "/q.fa", sd + "/d.fa", tmp + "/c.txt"], capture_output=True, text=True, check=True).stdout.strip().split("\n")
    seams["scores"] = [list(map(int, l.split()[1:])) for l in out]
    pairs = [(q, seams["candidates_t1"][q][i], s) for q in range(len(queries)) for i, s in enumerate(seams["scores"][q]) if s >= 40]
    with open(tmp + "/p.txt", "w") as f:
        for p in pairs:
            f.write("%d %d %d\n" % p)
    out = subprocess.run([O.REF_DUMP, "align", sd + "/q.fa", sd + "/d.fa", tmp + "/p.txt"], capture_output=True, text=True, check=True).stdout.strip().split("\n")
    seams["alignments"] = [{"q": p[0], "t": p[1], "score": p[2], "coords": list(map(int, l.split()[:4])), "path": l.split()[6]} for p, l in zip(pairs, out)]
    # swAlign fallback forced by gap open 128 (sse_module.c:215)
    sub = pairs[:30]
    with open(tmp + "/p2.txt", "w") as f:
        for p in sub:
            f.write("%d %d -1\n" % (p[0], p[1]))
    out = subprocess.run([O.REF_DUMP, "align", sd + "/q.fa", sd + "/d.fa", tmp + "/p2.txt", "128", "1"], capture_output=True, text=True, check=True).stdout.strip().split("\n")
    seams["alignments_swalign_go128"] = [{"q": p[0], "t": p[1], "score": int(l.split()[4]), "coords": list(map(int, l.split()[:4])), "path": l.split()[6]} for p, l in zip(sub, out)]
    # full pipeline: hits with E-values (hex doubles) in the reference's order
    out = subprocess.run([O.REF_DUMP, "pipeline", sd + "/q.fa", sd + "/d.fa", "5", "5000", "1", "0.0001", "400"], capture_output=True, text=True, check=True).stdout.split("\n")
    hits, cur = [], None
    for l in out:
        w = l.split()
        if not w:
            continue
        if w[0] == "query":
            cur = []
            hits.append(cur)
        elif w[0] == "hit":
            cur.append({"name": w[1], "score": int(w[2]), "evalue_hex": w[3], "coords": list(map(int, w[4:8])), "path": w[9] if len(w) > 9 else ""})
    seams["pipeline_hits"] = hits
    # the reference's BLOSUM_62 table through its own scorer
    out = subprocess.run([O.REF_DUMP, "matrix"], capture_output=True, text=True, check=True).stdout.split()
    seams["blosum62"] = list(map(int, out[1:]))
    json.dump(seams, open(os.path.join(HERE, "seams.json"), "w"))
    print("golden fixtures written:", {k: len(v) for k, v in hashes.items()}, "alignments", len(seams["alignments"]), "pipeline hits", sum(len(h) for h in hits))


if __name__ == "__main__":
    main()
