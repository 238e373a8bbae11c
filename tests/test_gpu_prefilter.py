"""GPU parity: s4g_prefilter (C ABI) against the oracle restatement of searchDatabase -- bit exact
candidate sets and float32 scores under the deterministic tie rule (score desc, id asc)."""
import numpy as np
import pytest

from oracle import oracle as O
from sift4g_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _compare(ctx, queries, db, k, N, id_base=0):
    qc, qo = synth.pack(queries)
    dc, do = synth.pack(db)
    D = ctx.database(dc, do, id_base=id_base)
    Q = ctx.queries(qc, qo)
    ids, sc, cnt = capi.prefilter(ctx, D, Q, k, N, sorted_by_id=True)
    ids_b, sc_b, cnt_b = capi.prefilter(ctx, D, Q, k, N, sorted_by_id=False)
    Q.close(); D.close()
    cells, oids, osc, _ = O.prefilter(dc, do, qc, qo, k, N)
    assert cells == int(do[-1])
    for q in range(len(queries)):
        n = int(cnt[q])
        assert n == len(oids[q]), "query %d: %d candidates, oracle %d" % (q, n, len(oids[q]))
        assert np.array_equal(ids[q, :n], oids[q] + id_base), "query %d candidate ids differ" % q
        assert np.array_equal(sc[q, :n].view(np.uint32), osc[q].view(np.uint32)), "query %d scores differ" % q
        # best-first layout: same set, ordered by (score desc, id asc)
        assert int(cnt_b[q]) == n
        order = np.lexsort((oids[q], -osc[q].astype(np.float64)))
        assert np.array_equal(ids_b[q, :n], (oids[q] + id_base)[order])
        assert np.array_equal(sc_b[q, :n], osc[q][order])


def test_planted_homologs_k5(ctx):
    queries, db = synth.make_dataset(21, 12, 4000, q_len=(40, 600), homologs=(5, 25), rare_fraction=0.01)
    _compare(ctx, queries, db, 5, 100)


def test_cutoff_ties_are_broken_by_id(ctx):
    # purely random data: almost every score is 1/len or 2/len, massive ties at the cut-off
    rng = np.random.default_rng(22)
    queries = [synth.random_codes(rng, n) for n in (300, 700, 1000)]
    db = [synth.random_codes(rng, n) for n in rng.integers(30, 200, size=6000)]
    _compare(ctx, queries, db, 5, 50)
    _compare(ctx, queries, db, 5, 5000)


def test_kmer_lengths_3_and_4(ctx):
    queries, db = synth.make_dataset(23, 4, 600, q_len=(30, 200), homologs=(2, 6))
    _compare(ctx, queries, db, 4, 40)
    _compare(ctx, queries[:2], db[:300], 3, 25)


def test_short_and_degenerate_sequences(ctx):
    rng = np.random.default_rng(24)
    q0 = synth.random_codes(rng, 120)
    queries = [q0, synth.random_codes(rng, 4), synth.random_codes(rng, 5), np.full(50, 0, dtype=np.uint8)]
    db = [q0[:4], q0[:5], q0[10:90], np.full(300, 0, dtype=np.uint8), np.full(7, 0, dtype=np.uint8), q0, queries[2].copy(),
          np.tile(q0[:10], 30), synth.random_codes(rng, 3)]
    _compare(ctx, queries, db, 5, 5)
    _compare(ctx, queries, db, 5, 3)


def test_many_hits_take_the_deferred_path(ctx):
    # repeats make > 1024 hits per sequence (beyond the shared-memory buffer)
    rng = np.random.default_rng(25)
    unit = synth.random_codes(rng, 40)
    queries = [np.tile(unit, 25), synth.random_codes(rng, 500), np.concatenate([unit, synth.random_codes(rng, 300), unit])]
    db = [np.tile(unit, 10), np.tile(unit, 60), synth.random_codes(rng, 200), np.concatenate([synth.random_codes(rng, 100), unit, unit]),
          queries[1].copy(), synth.mutate(rng, queries[1], 0.9)] + [synth.random_codes(rng, 150) for _ in range(200)]
    _compare(ctx, queries, db, 5, 20)


def test_id_base_offsets_shard_ids(ctx):
    queries, db = synth.make_dataset(26, 3, 500, q_len=(50, 300), homologs=(2, 5))
    _compare(ctx, queries, db, 5, 30, id_base=1000000)


def test_more_candidates_than_scored_sequences(ctx):
    queries, db = synth.make_dataset(27, 3, 50, q_len=(50, 120), homologs=(1, 3))
    _compare(ctx, queries, db, 5, 5000)


def test_multi_chunk_database(ctx):
    # > one scan chunk (131072 sequences) so the cut-off / compaction logic runs between chunks
    rng = np.random.default_rng(28)
    n = 300000
    lens = rng.integers(30, 60, size=n)
    off = np.zeros(n + 1, dtype=np.int64); off[1:] = np.cumsum(lens)
    codes = rng.choice(26, size=int(off[-1]), p=synth.letter_table(0.001)).astype(np.uint8)
    queries = [synth.random_codes(rng, 400), synth.random_codes(rng, 900)]
    db = [codes[off[i]:off[i + 1]] for i in range(n)]
    db[5] = queries[0][:59].copy(); db[250000] = queries[0][100:150].copy(); db[299999] = queries[1][:40].copy()
    _compare(ctx, queries, db, 5, 200)


def _reduced(rng, n, letters):
    return rng.integers(0, letters, size=n).astype(np.uint8)


def test_large_batch_takes_the_dense_scan(ctx):
    # > 8192 queries: 16-warp CTAs around the shared cut-off table, four hits per lane in flight.  A 12-letter alphabet
    # makes the steps dense (~2 index entries per k-mer position, a few hundred hits per 128 positions).
    rng = np.random.default_rng(29)
    queries = [_reduced(rng, n, 12) for n in rng.integers(40, 90, size=9000)]
    db = [_reduced(rng, n, 12) for n in rng.integers(30, 700, size=1200)]
    for i in range(0, 300, 7):                      # planted: mutated copies of queries inside random flanks
        db[i] = np.concatenate([_reduced(rng, 20, 12), synth.mutate(rng, queries[i * 13], 0.85), _reduced(rng, 15, 12)])
    _compare(ctx, queries, db, 5, 40)


def test_large_batch_with_thousands_of_hits_per_step(ctx):
    # 6-letter alphabet: thousands of hits per step and several thousand per sequence (counted first, re-walked for the survivors)
    rng = np.random.default_rng(30)
    queries = [_reduced(rng, n, 6) for n in rng.integers(40, 70, size=8500)]
    db = [_reduced(rng, n, 6) for n in rng.integers(30, 110, size=160)]
    db[3] = queries[17].copy(); db[90] = np.tile(queries[4000][:25], 3)
    _compare(ctx, queries, db, 5, 12)


def test_batch_too_large_for_the_shared_cutoff_table(ctx):
    # 27 000 queries: the 2-byte cut-off table no longer fits beside the per-warp buffers -> cut-offs read from global memory
    rng = np.random.default_rng(31)
    queries = [synth.random_codes(rng, n) for n in rng.integers(6, 14, size=27000)]
    db = [synth.random_codes(rng, n) for n in rng.integers(30, 300, size=800)]
    db[11] = np.concatenate([queries[5], queries[26999], queries[13000]])
    _compare(ctx, queries, db, 5, 6)


def test_mid_size_batch_takes_the_four_in_flight_build_with_8_warp_ctas(ctx):
    # 5000 queries x ~200 residues: the hit buffer grows to 512 entries, only two 8-warp CTAs fit an SM and the scan runs
    # the 4-hits-in-flight build with 8-warp CTAs and one hash -- the configuration of the 8-GPU weak-scaling bench
    # (8000 queries per rank)
    rng = np.random.default_rng(32)
    queries = [synth.random_codes(rng, n) for n in rng.integers(150, 250, size=5000)]
    db = [synth.random_codes(rng, n) for n in rng.integers(100, 500, size=500)]
    for i in range(0, 120, 3):
        db[i] = np.concatenate([synth.random_codes(rng, 30), synth.mutate(rng, queries[i * 40], 0.8), synth.random_codes(rng, 25)])
    _compare(ctx, queries, db, 5, 25)
