"""CPU: the oracle restatement against the committed golden vectors, which are outputs of the
UNMODIFIED reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np

from oracle import oracle as O
from sift4g_b200 import synth
from tests import util


def test_blosum62_table_equals_reference_scorer():
    s = util.seams()
    assert np.array_equal(O.blosum62(), np.array(s["blosum62"], dtype=np.int32))


def test_prefilter_candidates_match_reference_up_to_cutoff_ties():
    s = util.seams()
    _, queries, _, db = util.synth_e2e()
    qc, qo = synth.pack(queries); dc, do = synth.pack(db)
    N = s["max_candidates"]
    cells, ids, sc, dense = O.prefilter(dc, do, qc, qo, 5, N, dense=True)
    assert cells == s["cells"]
    for q, ref in enumerate(s["candidates_t1"]):
        a, b = set(ids[q].tolist()), set(ref)
        assert len(a) == len(b)
        if a != b:          # SURVEY.md section 8c tie rule: differences only at the cut-off score
            cut = min(dense[q, list(b)])
            assert cut == min(dense[q, list(a)])
            assert all(dense[q, x] == cut for x in a ^ b)


def test_sw_scores_equal_swimd():
    s = util.seams()
    _, queries, _, db = util.synth_e2e()
    mat = O.blosum62()
    for q, (cands, scores) in enumerate(zip(s["candidates_t1"], s["scores"])):
        for t, ref in zip(cands, scores):
            assert O.sw_score(queries[q], db[t], mat) == ref


def test_ssw_rule_paths_equal_reference():
    s = util.seams()
    _, queries, _, db = util.synth_e2e()
    mat = O.blosum62()
    assert len(s["alignments"]) > 100
    for a in s["alignments"]:
        coords, path = O.align(queries[a["q"]], db[a["t"]], a["score"], mat)
        assert list(coords) == a["coords"], a
        assert util.path_str(path) == a["path"], a


def test_swalign_rule_paths_equal_reference_fallback():
    s = util.seams()
    _, queries, _, db = util.synth_e2e()
    mat = O.blosum62()
    for a in s["alignments_swalign_go128"]:
        score = O.sw_score(queries[a["q"]], db[a["t"]], mat, 128, 1)
        assert score == a["score"]
        coords, path = O.align(queries[a["q"]], db[a["t"]], score, mat, 128, 1)
        assert list(coords) == a["coords"], a
        assert util.path_str(path) == (a["path"] if a["path"] != "-" else ""), a


def test_evalues_and_selection_order_equal_reference_pipeline():
    s = util.seams()
    qn, queries, dn, db = util.synth_e2e()
    mat = O.blosum62()
    name_to_idx = {n: i for i, n in enumerate(dn)}
    qc, qo = synth.pack(queries); dc, do = synth.pack(db)
    cells, ids, _, _ = O.prefilter(dc, do, qc, qo, 5, 5000)
    for q, hits in enumerate(s["pipeline_hits"]):
        scores = np.array([O.sw_score(queries[q], db[t], mat) for t in ids[q]], dtype=np.int32)
        values = np.array([O.evalue(sc, len(queries[q]), len(db[t]), cells) for sc, t in zip(scores, ids[q])])
        order = O.select(values, scores, [dn[t] for t in ids[q]], 1e-4, 400)
        assert len(order) == len(hits)
        for k, h in zip(order, hits):
            t = ids[q][k]
            assert name_to_idx[h["name"]] == t
            assert scores[k] == h["score"]
            assert values[k] == float.fromhex(h["evalue_hex"]), (values[k].hex(), h["evalue_hex"])
            coords, path = O.align(queries[q], db[t], scores[k], mat)
            assert list(coords) == h["coords"]
            assert util.path_str(path) == h["path"]


def test_lis_is_strict_and_counts_same_position_hits():
    assert O.lis([1, 2, 3]) == 3
    assert O.lis([3, 3, 3]) == 1
    assert O.lis([5, 1, 2, 2, 3]) == 3
    assert O.lis([]) == 0
    assert O.lis([10, 20, 5, 6, 7, 1]) == 3


def test_encode_drops_non_letters_and_folds_case():
    assert O.encode("aC-d*Z 1\n").tolist() == [0, 2, 3, 25]
