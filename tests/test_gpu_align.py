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


def test_swalign_rule_golden_paths_from_the_reference(ctx, blosum):
    """Gap open 128 makes the reference leave SSW for swAlign (sse_module.c:215); tests/golden/seams.json holds the
    reference's own coords/paths for 30 such pairs."""
    from tests import util
    s = util.seams()
    _, queries, _, db = util.synth_e2e()
    qc, qo = synth.pack(queries); dc, do = synth.pack(db)
    D = ctx.database(dc, do); Q = ctx.queries(qc, qo)
    al = [a for a in s["alignments_swalign_go128"] if a["score"] > 0]
    pq = np.array([a["q"] for a in al], dtype=np.uint32); pt = np.array([a["t"] for a in al], dtype=np.uint32)
    ps = np.array([a["score"] for a in al], dtype=np.int32)
    coords, paths = capi.sw_align(ctx, D, Q, pq, pt, ps, blosum, 128, 1, q_lens=np.diff(qo), t_lens=np.diff(do))
    Q.close(); D.close()
    assert len(al) >= 20
    for i, a in enumerate(al):
        assert list(coords[i]) == a["coords"], a
        assert util.path_str(paths[i]) == a["path"], a


def test_swalign_rule_forced_by_large_gap_penalties(ctx, blosum):
    queries, db = synth.make_dataset(36, 6, 120, q_len=(60, 700), homologs=(5, 9), rare_fraction=0.01)
    pairs = []
    for q in range(len(queries)):
        sc = np.array([O.sw_score(queries[q], t, blosum, 128, 2) for t in db])
        pairs += [(q, int(t)) for t in np.argsort(-sc)[:12]]
    assert _align_and_check(ctx, blosum, queries, db, pairs, 128, 2) > 50
    assert _align_and_check(ctx, blosum, queries, db, pairs[:30], 20, 130) > 10


def test_scores_above_32767_take_the_swalign_rule(ctx, blosum):
    """A near-identical copy of a 7 000-residue query scores > 32767: the reference switches to swAlign
    (sse_module.c:181-185); mixed with an ordinary hit of the same query in one call."""
    rng = np.random.default_rng(37)
    q = synth.random_codes(rng, 7000)
    t_big = synth.mutate(rng, q, identity=0.93, indel_rate=0.004)
    t_big = np.concatenate([synth.random_codes(rng, 40), t_big, synth.random_codes(rng, 25)]).astype(np.uint8)
    t_small = synth.mutate(rng, q[1000:1800], identity=0.7)
    t_shift = np.concatenate([q[3000:], q[:3000]]).astype(np.uint8)            # two competing diagonals
    queries = [q, synth.random_codes(rng, 300)]
    db = [t_big, t_small, t_shift, synth.mutate(rng, queries[1], 0.8)]
    assert O.sw_score(q, t_big, blosum) > 32767
    _align_and_check(ctx, blosum, queries, db, [(0, 0), (0, 1), (0, 2), (1, 3)])
