"""CPU: s4g_select_hits (threaded host C++, no GPU involved) against the oracle -- E-values bit-identical to the libm
evaluation in the reference's operation order (sw/evalue.cu:436-489) although the query-side factors are memoised per
(query, score), and the kept rows in dbAlignmentDataCmp order (sw/database.c:1043-1059)."""
import numpy as np
import pytest

from oracle import oracle as O
from sift4g_b200 import capi


@pytest.mark.parametrize("go,ge", [(10, 1), (11, 1), (9, 2)])
def test_select_hits_matches_the_oracle(go, ge):
    rng = np.random.default_rng(100 + go)
    nq, per_q, db_res = 7, 900, 3_231_000_000
    qlens = rng.integers(40, 2500, size=nq).astype(np.int32)
    off = np.arange(nq + 1, dtype=np.int64) * per_q
    ids = rng.permutation(nq * per_q * 3)[:nq * per_q].astype(np.uint32)
    # few distinct scores per query (the memo is hit hard), some far above and below the E-value threshold
    scores = rng.choice(np.array([25, 60, 118, 119, 120, 121, 150, 151, 300, 1200, 40000], dtype=np.int32), size=nq * per_q)
    tlens = rng.integers(30, 3000, size=nq * per_q).astype(np.int32)
    for threads in (1, 3):
        pq, pt, ps, ev, hoff = capi.select_hits(None, qlens, ids, off, scores, tlens, db_res, go, ge, 1e-4, 400, n_threads=threads)
        for q in range(nq):
            rows = []
            for i in range(off[q], off[q + 1]):
                e = O.evalue(int(scores[i]), int(qlens[q]), int(tlens[i]), db_res, go, ge)
                if e <= 1e-4:
                    rows.append((e, -int(scores[i]), int(ids[i])))
            rows.sort()
            rows = rows[:400]
            got = [(float(ev[h]), -int(ps[h]), int(pt[h])) for h in range(hoff[q], hoff[q + 1])]
            assert got == rows
            assert all(int(x) == q for x in pq[hoff[q]:hoff[q + 1]])
