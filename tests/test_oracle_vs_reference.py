"""CPU: the oracle restatement against the reference itself (oracle/_ref binaries), on fresh seeded data.
Skipped where oracle/_ref is absent; in the build container it always runs."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import oracle as O
from sift4g_b200 import synth
from tests import util

pytestmark = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")


def _dump(args):
    return subprocess.run([O.REF_DUMP] + args, capture_output=True, text=True, check=True).stdout.strip().split("\n")


@pytest.mark.parametrize("seed", [101, 102])
def test_all_seams_on_random_homolog_sets(seed):
    tmp = tempfile.mkdtemp()
    queries, db = synth.make_dataset(seed, 5, 1200, q_len=(50, 350), homologs=(5, 15), rare_fraction=0.02)
    qf, df = tmp + "/q.fa", tmp + "/d.fa"
    synth.write_fasta(qf, queries, "Q"); synth.write_fasta(df, db, "D")
    qc, qo = synth.pack(queries); dc, do = synth.pack(db)
    mat = O.blosum62()
    N = 40
    out = _dump(["candidates", qf, df, "5", str(N), "1"])
    cells, ids, sc, dense = O.prefilter(dc, do, qc, qo, 5, N, dense=True)
    assert cells == int(out[0].split()[1])
    for q in range(len(queries)):
        ref = set(map(int, out[1 + q].split()[1:]))
        mine = set(ids[q].tolist())
        assert len(ref) == len(mine)
        if ref != mine:
            cut = min(dense[q, list(ref)])
            assert all(dense[q, x] == cut for x in ref ^ mine)
    with open(tmp + "/c.txt", "w") as f:
        for c in ids:
            f.write("%d %s\n" % (len(c), " ".join(map(str, c))))
    out = _dump(["scores", qf, df, tmp + "/c.txt"])
    pairs = []
    for q in range(len(queries)):
        ref = list(map(int, out[q].split()[1:]))
        for t, r in zip(ids[q], ref):
            assert O.sw_score(queries[q], db[t], mat) == r
            if r >= 35:
                pairs.append((q, int(t), r))
    with open(tmp + "/p.txt", "w") as f:
        for p in pairs:
            f.write("%d %d %d\n" % p)
    out = _dump(["align", qf, df, tmp + "/p.txt"])
    assert len(pairs) > 20
    for p, l in zip(pairs, out):
        w = l.split()
        coords, path = O.align(queries[p[0]], db[p[1]], p[2], mat)
        assert list(coords) == list(map(int, w[:4]))
        assert util.path_str(path) == w[6]


def test_reference_binary_reproduces_committed_hashes():
    import hashlib, json
    exp = json.load(open(os.path.join(util.GOLDEN, "expected_hashes.json")))["test_files_subst"]
    tmp = tempfile.mkdtemp()
    tf = os.path.join(util.GOLDEN, "test_files")
    subprocess.run([O.REF_SIFT4G, "-q", tf + "/query.fasta", "-d", tf + "/sample_protein_database.fa", "--subst", tf + "/",
                    "--sub-results", "--out", tmp], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for f, h in exp.items():
        assert hashlib.sha256(open(os.path.join(tmp, f), "rb").read()).hexdigest() == h, f
    # and the digests published in BASELINE.md section 5
    assert exp["LACI_ECOLI.SIFTprediction"].startswith("9941c9d6") and exp["alignments.txt"].startswith("69098a8f")
