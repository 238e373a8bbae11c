"""GPU parity: s4g_sw_align (C ABI) against the oracle's SSW-rule restatement -- identical end/begin
cells and identical path bytes."""
import numpy as np
import pytest

from oracle import oracle as O
from sift4g_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _align_and_check(ctx, blosum, queries, db, pairs, go=10, ge=1):
    qc, qo = synth.pack(queries)
    dc, do = synth.pack(db)
    D = ctx.database(dc, do)
    Q = ctx.queries(qc, qo)
    pq = np.array([p[0] for p in pairs], dtype=np.uint32)
    pt = np.array([p[1] for p in pairs], dtype=np.uint32)
    ps = np.array([O.sw_score(queries[a], db[b], blosum, go, ge) for a, b in pairs], dtype=np.int32)
    keep = ps > 0
    pq, pt, ps = pq[keep], pt[keep], ps[keep]
    qlens = np.diff(qo); tlens = np.diff(do)
    coords, paths = capi.sw_align(ctx, D, Q, pq, pt, ps, blosum, go, ge, q_lens=qlens, t_lens=tlens)
    Q.close(); D.close()
    bad = 0
    for i in range(len(pq)):
        ec, ep = O.align(queries[pq[i]], db[pt[i]], ps[i], blosum, go, ge)
        if not (np.array_equal(ec, coords[i]) and np.array_equal(ep, paths[i])):
            bad += 1
            assert bad < 4, "pair %d (q%d,t%d,score %d): gpu %s len %d, oracle %s len %d" % (
                i, pq[i], pt[i], ps[i], coords[i], len(paths[i]), ec, len(ep))
    assert bad == 0
    return len(pq)


def test_planted_homologs(ctx, blosum):
    queries, db = synth.make_dataset(31, 10, 400, q_len=(40, 500), homologs=(6, 14), rare_fraction=0.01)
    pairs = []
    for q in range(len(queries)):
        sc = np.array([O.sw_score(queries[q], t, blosum) for t in db])
        top = np.argsort(-sc)[:25]
        pairs += [(q, int(t)) for t in top]
    n = _align_and_check(ctx, blosum, queries, db, pairs)
    assert n > 100


def test_indel_rich_pairs_need_band_doubling(ctx, blosum):
    rng = np.random.default_rng(32)
    queries, db, pairs = [], [], []
    for i in range(40):
        q = synth.random_codes(rng, int(rng.integers(80, 400)))
        t = synth.mutate(rng, q, identity=0.75, indel_rate=0.08, max_indel=25)
        # a long insertion in the middle forces wide bands
        cut = len(t) // 2
        t = np.concatenate([synth.random_codes(rng, 15), t[:cut], synth.random_codes(rng, int(rng.integers(0, 60))), t[cut:], synth.random_codes(rng, 9)])
        queries.append(q); db.append(t.astype(np.uint8)); pairs.append((i, i))
    _align_and_check(ctx, blosum, queries, db, pairs)


def test_repeats_and_ties(ctx, blosum):
    rng = np.random.default_rng(33)
    unit = synth.random_codes(rng, 12)
    queries = [np.tile(unit, 8), np.concatenate([unit, unit[::-1], unit]), synth.random_codes(rng, 60)]
    db = [np.tile(unit, 5), np.tile(unit, 13), np.concatenate([unit[:6], unit, unit]), queries[2][10:50].copy(),
          np.concatenate([queries[2][:30], queries[2][:30]])]
    pairs = [(q, t) for q in range(3) for t in range(5)]
    _align_and_check(ctx, blosum, queries, db, pairs)


def test_other_gap_penalties_and_long_query(ctx, blosum):
    rng = np.random.default_rng(34)
    q = synth.random_codes(rng, 1500)          # > 256 rows: multi-pass end/begin sweeps
    queries = [q, synth.random_codes(rng, 200)]
    db = [synth.mutate(rng, q[300:1200], 0.8), synth.mutate(rng, q, 0.6), synth.mutate(rng, queries[1], 0.7), q[:100].copy()]
    pairs = [(0, 0), (0, 1), (0, 3), (1, 2)]
    for go, ge in ((10, 1), (11, 1), (5, 2)):
        _align_and_check(ctx, blosum, queries, db, pairs, go, ge)


def test_identical_and_single_residue_hits(ctx, blosum):
    rng = np.random.default_rng(35)
    q = synth.random_codes(rng, 120)
    w = np.array([22], dtype=np.uint8)             # a single tryptophan scores 11
    queries = [q, w]
    db = [q.copy(), w.copy(), np.concatenate([synth.random_codes(rng, 5), w])]
    _align_and_check(ctx, blosum, queries, db, [(0, 0), (1, 1), (1, 2), (0, 2)])
