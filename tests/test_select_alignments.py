"""alignmentsExtract + alignmentsSelect (sift4g/src/select_alignments.cpp:127-242).

CPU: the oracle's restatement reproduces the reference CLI's own `.aligned.fasta` files (hashes committed by
tests/golden/make_golden.py) from the reference's alignments of the same run (tests/golden/seams.json, pipeline_hits).
GPU: s4g_alignment_strings / s4g_alignments_select (C ABI) against the oracle on seeded inputs, and the same golden files.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import util


def _aligned_fasta(query, names, strings):
    """outputSelectedAlignments, select_alignments.cpp:76-110"""
    def block(s):
        out = []
        for j in range(1, len(s) + 1):
            out.append(s[j - 1:j])
            if j % 60 == 0:
                out.append(b"\n")
        return b"".join(out) + b"\n"
    txt = b">QUERY\n" + block(bytes(query + 65))
    for n, s in zip(names, strings):
        txt += b">" + n.encode() + b"\n" + block(s)
    return txt


def _golden_case():
    s = util.seams()
    qn, queries, dn, db = util.synth_e2e()
    index = {n: i for i, n in enumerate(dn)}
    return s["pipeline_hits"], qn, queries, dn, db, index


def test_oracle_reproduces_the_reference_aligned_fasta_files():
    hits, qn, queries, dn, db, index = _golden_case()
    want = json.load(open(os.path.join(util.GOLDEN, "expected_hashes.json")))["synth_default"]
    kept_total = 0
    for q, qh in enumerate(hits):
        strings = [O.alignment_string(db[index[h["name"]]], len(queries[q]), h["coords"], np.array(list(map(int, h["path"])), dtype=np.uint8)) for h in qh]
        k = O.alignments_select(strings, len(queries[q]), 2.75)
        kept_total += k
        txt = _aligned_fasta(queries[q], [h["name"] for h in qh[:k]], strings[:k])
        assert hashlib.sha256(txt).hexdigest() == want["%s.aligned.fasta" % qn[q]], "query %d: %d of %d alignments kept" % (q, k, len(qh))
    assert 0 < kept_total <= sum(len(h) for h in hits)


def test_alignment_table_of_the_reference_from_result_buffers(tmp_path):
    """Row F4: the oracle's per-hit counts + the library's writer (s4g_write_blast_tab: host only, no GPU) reproduce the
    reference CLI's alignments.txt (bm9) byte for byte from the reference's own alignments."""
    from sift4g_b200 import capi
    hits, qn, queries, dn, db, index = _golden_case()
    want = json.load(open(os.path.join(util.GOLDEN, "expected_hashes.json")))["synth_default"]["alignments.txt"]
    flat = [x for h in hits for x in h]
    hoff = np.zeros(len(hits) + 1, dtype=np.int64); hoff[1:] = np.cumsum([len(h) for h in hits])
    stats = np.array([O.alignment_stats(queries[q], db[index[x["name"]]], x["coords"], np.array(list(map(int, x["path"])), dtype=np.uint8))
                      for q, h in enumerate(hits) for x in h], dtype=np.int32)
    out = str(tmp_path / "alignments.txt")
    capi.write_blast_tab(out, True, hoff, qn, [x["name"] for x in flat], stats, np.array([x["coords"] for x in flat], dtype=np.int32),
                         np.array([float.fromhex(x["evalue_hex"]) for x in flat]), np.array([x["score"] for x in flat], dtype=np.int32))
    assert hashlib.sha256(open(out, "rb").read()).hexdigest() == want
    assert (stats[:, 0] + stats[:, 1] <= stats[:, 3]).all() and stats[:, 2].max() > 0


def test_median_quirk_and_edge_cases_of_the_oracle():
    # one string: kept unless the threshold is not below log2(20) to begin with
    assert O.alignments_select([b"ACDE"], 4, 2.75) == 1
    assert O.alignments_select([b"ACDE"], 4, 5.0) == 0
    # fully conserved columns never drop the median: everything is kept
    assert O.alignments_select([b"ACDEFGHIKL"] * 7, 10, 2.75) == 7
    # diverse columns stop the loop early
    rng = np.random.default_rng(5)
    strings = [bytes(rng.integers(65, 85, size=40).astype(np.uint8)) for _ in range(30)]
    k = O.alignments_select(strings, 40, 2.75)
    assert 1 <= k < 30


@pytest.mark.gpu
def test_gpu_strings_and_selection_match_the_oracle(ctx, blosum):
    from sift4g_b200 import capi, pipeline, synth
    queries, db = synth.make_dataset(61, 14, 3000, q_len=(40, 700), homologs=(10, 60), rare_fraction=0.005)
    queries.append(synth.random_codes(np.random.default_rng(1), 1))            # a query that will keep nothing
    qc, qo = synth.pack(queries); dc, do = synth.pack(db)
    D = ctx.database(dc, do)
    Q = ctx.queries(qc, qo)
    out = capi.search(ctx, D, Q, blosum, max_candidates=400, max_alignments=60)
    pq, pt, coords, paths, poff, hoff = (out.pair_q.copy(), out.pair_t.copy(), out.coords.copy(), out.paths.copy(), out.path_off.copy(), out.hit_off.copy())
    assert len(pq) > 200
    strings, soff = capi.alignment_strings(ctx, D, Q, pq, pt, coords, paths, poff)
    for h in range(len(pq)):
        exp = O.alignment_string(db[pt[h]], len(queries[pq[h]]), coords[h], paths[poff[h]:poff[h + 1]])
        assert bytes(strings[soff[h]:soff[h + 1]]) == exp, "string of hit %d" % h
    stats = capi.alignment_stats(ctx, D, Q, pq, pt, coords, paths, poff)
    for h in range(len(pq)):
        assert np.array_equal(stats[h], O.alignment_stats(queries[pq[h]], db[pt[h]], coords[h], paths[poff[h]:poff[h + 1]])), "table counts of hit %d" % h
    for thr in (2.75, 3.25, 1.0):
        sel = capi.alignments_select(ctx, np.diff(qo), hoff, strings, thr)
        for q in range(len(queries)):
            a, b = int(hoff[q]), int(hoff[q + 1])
            exp = O.alignments_select([bytes(strings[soff[h]:soff[h + 1]]) for h in range(a, b)], len(queries[q]), thr)
            assert sel[q] == exp, "query %d (%d hits) at threshold %.2f" % (q, b - a, thr)
    assert sel.sum() == len(pq)                     # at threshold 1.0 no median gets that low: every hit is kept
    Q.close(); D.close()


@pytest.mark.gpu
def test_gpu_selection_reproduces_the_reference_aligned_fasta_files(ctx):
    from sift4g_b200 import capi, synth
    hits, qn, queries, dn, db, index = _golden_case()
    want = json.load(open(os.path.join(util.GOLDEN, "expected_hashes.json")))["synth_default"]
    qc, qo = synth.pack(queries); dc, do = synth.pack(db)
    D = ctx.database(dc, do); Q = ctx.queries(qc, qo)
    pq = np.concatenate([np.full(len(h), q, dtype=np.uint32) for q, h in enumerate(hits)])
    pt = np.array([index[x["name"]] for h in hits for x in h], dtype=np.uint32)
    coords = np.array([x["coords"] for h in hits for x in h], dtype=np.int32)
    plist = [np.array(list(map(int, x["path"])), dtype=np.uint8) for h in hits for x in h]
    poff = np.zeros(len(plist) + 1, dtype=np.int64); poff[1:] = np.cumsum([len(p) for p in plist])
    hoff = np.zeros(len(hits) + 1, dtype=np.int64); hoff[1:] = np.cumsum([len(h) for h in hits])
    strings, soff = capi.alignment_strings(ctx, D, Q, pq, pt, coords, np.concatenate(plist), poff)
    sel = capi.alignments_select(ctx, np.diff(qo), hoff, strings, 2.75)
    for q, qh in enumerate(hits):
        a = int(hoff[q])
        kept = [bytes(strings[soff[h]:soff[h + 1]]) for h in range(a, a + int(sel[q]))]
        txt = _aligned_fasta(queries[q], [x["name"] for x in qh[:int(sel[q])]], kept)
        assert hashlib.sha256(txt).hexdigest() == want["%s.aligned.fasta" % qn[q]]
    Q.close(); D.close()
