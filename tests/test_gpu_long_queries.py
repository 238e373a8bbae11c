"""GPU parity for BASELINE.json configs[3] (queries of 5 000 - 35 000 aa): SW scores, end/begin cells and path bytes
against the oracle for the cases the short-query tests do not reach --
  * queries far beyond one stripe of the intra-sequence kernel (34 stripes at 35 000 aa),
  * targets longer than the striped kernel's column buffer (4 096): the 32-bit kernel,
  * a planted 90 %-identical homolog whose score exceeds 32 767: exact 32-bit re-run of the score and the swAlign
    traceback rule (sw/sse_module.c:181-185, sw/cpu_module.c:1185-1413),
  * ordinary hits (score <= 32 767) of very long queries: SSW end/begin cells + banded_sw path (sw/ssw/ssw.c).
"""
import numpy as np
import pytest

from oracle import oracle as O
from sift4g_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _dataset():
    rng = np.random.default_rng(4004)
    qlens = (5000, 9000, 20011, 35000)
    queries = [synth.random_codes(rng, n) for n in qlens]
    db = []
    # 0: 90 % identical homolog of the 5 000-aa query: target > 4 096, score < 32 767 (SSW rule)
    db.append(np.concatenate([synth.random_codes(rng, 60), synth.mutate(rng, queries[0], identity=0.90, indel_rate=0.004), synth.random_codes(rng, 41)]).astype(np.uint8))
    # 1: 90 % identical homolog of the 9 000-aa query: score > 32 767 (swAlign rule), target > 4 096
    db.append(np.concatenate([synth.random_codes(rng, 33), synth.mutate(rng, queries[1], identity=0.90, indel_rate=0.003), synth.random_codes(rng, 20)]).astype(np.uint8))
    # 2, 3: domain-level homologs of the 20 011- and the 35 000-aa query (windows deep inside the query)
    db.append(synth.mutate(rng, queries[2][15000:17600], identity=0.75, indel_rate=0.02))
    db.append(np.concatenate([synth.random_codes(rng, 120), synth.mutate(rng, queries[3][31000:34000], identity=0.8, indel_rate=0.01)]).astype(np.uint8))
    # 4: two windows of the 35 000-aa query far apart, joined: two competing local alignments
    db.append(np.concatenate([queries[3][2000:2900], synth.random_codes(rng, 50), queries[3][30000:30800]]).astype(np.uint8))
    # 5..: unrelated sequences of ordinary lengths, one of them longer than 4 096
    db += [synth.random_codes(rng, n) for n in (35, 180, 333, 612, 1500, 4500)]
    return queries, db


def test_long_query_scores_cells_and_paths(ctx, blosum):
    queries, db = _dataset()
    qc, qo = synth.pack(queries)
    dc, do = synth.pack(db)
    D = ctx.database(dc, do)
    Q = ctx.queries(qc, qo)
    nq, nd = len(queries), len(db)
    # every query against every target
    ids = np.tile(np.arange(nd, dtype=np.uint32), nq)
    off = np.arange(nq + 1, dtype=np.int64) * nd
    out = capi.sw_score(ctx, D, Q, ids, off, blosum, 10, 1)
    exp = np.array([O.sw_score(queries[q], db[t], blosum) for q in range(nq) for t in range(nd)], dtype=np.int32)
    assert np.array_equal(out, exp), "scores differ at pairs %s" % np.nonzero(out != exp)[0][:8]
    assert exp[1 * nd + 1] > 32767                    # the planted homolog of the 9 000-aa query leaves the 16-bit range
    assert exp[0 * nd + 0] < 32767 and len(db[0]) > 4096 and len(db[1]) > 4096
    # alignments of the planted pairs, the two-window target and one chance hit per query
    pairs = [(0, 0), (1, 1), (2, 2), (3, 3), (3, 4), (0, 10), (2, 9), (3, 8)]
    pq = np.array([p[0] for p in pairs], dtype=np.uint32)
    pt = np.array([p[1] for p in pairs], dtype=np.uint32)
    ps = np.array([exp[q * nd + t] for q, t in pairs], dtype=np.int32)
    assert (ps > 0).all()
    coords, paths = capi.sw_align(ctx, D, Q, pq, pt, ps, blosum, 10, 1, q_lens=np.diff(qo), t_lens=np.diff(do))
    Q.close(); D.close()
    for i, (q, t) in enumerate(pairs):
        ec, ep = O.align(queries[q], db[t], int(ps[i]), blosum)
        assert np.array_equal(ec, coords[i]), "pair (q%d, t%d) score %d: cells %s, oracle %s" % (q, t, ps[i], coords[i], ec)
        assert np.array_equal(ep, paths[i]), "pair (q%d, t%d) score %d: path of %d moves, oracle %d" % (q, t, ps[i], len(paths[i]), len(ep))
