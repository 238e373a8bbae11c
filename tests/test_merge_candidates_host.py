"""CPU: s4g_merge_candidates_host (pure host code) -- the merge of per-shard best-first candidate rows the CLI uses when it
drives several GPUs: per query the global max_candidates best by (score desc, id asc), written in ascending id, i.e. the
reference's merge of its per-thread lists (sift4g/src/database_search.cpp:132-154,173-180) under the deterministic tie rule.
Checked against the definition (one sort of the union) and against what a single shard holding everything would return."""
import numpy as np
import pytest

from sift4g_b200 import capi


def _shard_rows(rng, nq, n_total, N, n_shards, tie_heavy):
    """Scores of n_total sequences per query; every shard returns its own best-first top N."""
    if tie_heavy:
        scores = (rng.integers(1, 4, size=(nq, n_total)) / rng.integers(30, 40, size=(1, n_total))).astype(np.float32)
    else:
        scores = rng.random((nq, n_total)).astype(np.float32)
    present = rng.random((nq, n_total)) < 0.7                 # a query shares a k-mer with only some sequences
    bounds = np.linspace(0, n_total, n_shards + 1).astype(int)
    bounds[1:-1] += rng.integers(-3, 4, size=n_shards - 1)    # uneven shards
    ids, scs, cnts = [], [], []
    for r in range(n_shards):
        lo, hi = bounds[r], bounds[r + 1]
        I = np.zeros((nq, N), dtype=np.uint32); S = np.zeros((nq, N), dtype=np.float32); Cn = np.zeros(nq, dtype=np.uint32)
        for q in range(nq):
            cand = np.nonzero(present[q, lo:hi])[0] + lo
            order = np.lexsort((cand, -scores[q, cand].astype(np.float64)))[:N]
            I[q, :len(order)] = cand[order]; S[q, :len(order)] = scores[q, cand[order]]; Cn[q] = len(order)
        ids.append(I); scs.append(S); cnts.append(Cn)
    return scores, present, ids, scs, cnts


@pytest.mark.parametrize("n_shards,tie_heavy", [(1, False), (2, True), (3, False), (8, True)])
def test_host_merge_equals_one_sort_of_the_union(n_shards, tie_heavy):
    rng = np.random.default_rng(40 + n_shards)
    nq, n_total, N = 9, 400, 50
    scores, present, ids, scs, cnts = _shard_rows(rng, nq, n_total, N, n_shards, tie_heavy)
    for threads in (1, 4):
        out, cnt = capi.merge_candidates_host(ids, scs, cnts, N, n_threads=threads)
        for q in range(nq):
            cand = np.nonzero(present[q])[0]
            best = cand[np.lexsort((cand, -scores[q, cand].astype(np.float64)))[:N]]        # what one shard holding everything keeps
            assert int(cnt[q]) == len(best)
            assert np.array_equal(out[q, :cnt[q]], np.sort(best).astype(np.uint32))


def test_host_merge_with_short_and_empty_rows():
    ids = [np.array([[5, 3, 0], [0, 0, 0]], dtype=np.uint32), np.array([[9, 0, 0], [0, 0, 0]], dtype=np.uint32)]
    scs = [np.array([[.5, .25, 0], [0, 0, 0]], dtype=np.float32), np.array([[.25, 0, 0], [0, 0, 0]], dtype=np.float32)]
    cnts = [np.array([2, 0], dtype=np.uint32), np.array([1, 0], dtype=np.uint32)]
    out, cnt = capi.merge_candidates_host(ids, scs, cnts, 3)
    assert cnt.tolist() == [3, 0] and out[0].tolist() == [3, 5, 9]
    # rows of 2: shard 0 keeps (5, 3), shard 1 keeps (9); ids 3 and 9 tie at 0.25 -- the smaller id stays
    out, cnt = capi.merge_candidates_host([a[:, :2] for a in ids], [a[:, :2] for a in scs], cnts, 2)
    assert cnt.tolist() == [2, 0] and out[0, :2].tolist() == [3, 5]
