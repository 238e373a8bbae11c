"""GPU: every kernel that has an older, independently written form still in the library gives that form's results bit for bit on a
workload large enough to meet all their code paths (begin cells: packed reverse sweep vs the 32-bit sweep; paths: 8/16-lane group
kernels vs one warp per hit; long queries: streaming stripes vs pair by pair, striped end cells vs the 32-bit sweep; no speculative
traceback; score kernel with one stream column per step vs two; tiles in query order vs by descending query length; flagged sequences
of the prefilter re-walked vs replayed from the rank queue).  The forms are selected per call through the environment switches the kernels document."""
import os

import numpy as np
import pytest

from sift4g_b200 import pipeline, synth

pytestmark = pytest.mark.gpu

SWITCHES = [("S4G_BEGINS", "sweep32"), ("S4G_BAND_GROUPS", "0"), ("S4G_BAND_GROUPS", "16"), ("S4G_STRIPED", "pairs"), ("S4G_STRIPED", "stream1"), ("S4G_ENDS", "sweep32"),
            ("S4G_NO_SPECULATE", "1"), ("S4G_SCORE", "1col"), ("S4G_TILE_ORDER", "query"), ("S4G_PF_REPLAY", "0")]


def _run(ctx, D, qc, qo, blosum, N, M):
    o = pipeline.search_host(ctx, D, qc, qo, blosum, max_candidates=N, max_alignments=M)
    n = int(o.path_off[-1])
    return (o.cand_ids.copy(), o.pair_q.copy(), o.pair_t.copy(), o.pair_score.copy(), o.evalue.copy(), o.coords.copy(), o.path_off.copy(), o.paths[:n].copy())


def test_older_kernel_forms_agree_with_the_current_ones(ctx, blosum):
    rng = np.random.default_rng(5)
    queries, db = synth.make_dataset(91, 40, 60000, q_len=(60, 1000), homologs=(20, 60), rare_fraction=0.003)
    # a few long queries with long, partly very similar homologs (stripes, boundary rows, end rows beyond 1024, score > 32767)
    for L in (1500, 2600, 9000):
        q = synth.random_codes(rng, L, 0.001)
        queries.append(q)
        for ident in (0.5, 0.8, 0.97):
            a = int(rng.integers(0, L // 10)); b = int(rng.integers(9 * L // 10, L))
            db[int(rng.integers(0, len(db)))] = np.concatenate([synth.random_codes(rng, 40, 0.001), synth.mutate(rng, q[a:b], identity=ident), synth.random_codes(rng, 25, 0.001)])
    qc, qo = synth.pack(queries); dc, do = synth.pack(db)
    D = ctx.database(dc, do)
    want = _run(ctx, D, qc, qo, blosum, 800, 100)
    assert len(want[1]) > 800, len(want[1])
    assert (want[3] > 32767).any(), int(want[3].max())
    assert (want[5][:, 1] >= 1024).any()
    for name, value in SWITCHES:
        old = os.environ.get(name)
        os.environ[name] = value
        try:
            got = _run(ctx, D, qc, qo, blosum, 800, 100)
        finally:
            if old is None:
                del os.environ[name]
            else:
                os.environ[name] = old
        for i, (g, w) in enumerate(zip(got, want)):
            assert np.array_equal(g, w), "%s=%s: result array %d differs" % (name, value, i)
    D.close()
