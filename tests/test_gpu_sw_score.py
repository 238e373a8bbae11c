"""GPU parity: s4g_sw_score (C ABI) against the oracle's scalar Gotoh restatement -- bit exact."""
import numpy as np
import pytest

from oracle import oracle as O
from sift4g_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _run(ctx, blosum, queries, db, cands, go=10, ge=1):
    qc, qo = synth.pack(queries)
    dc, do = synth.pack(db)
    D = ctx.database(dc, do)
    Q = ctx.queries(qc, qo)
    ids = np.concatenate(cands).astype(np.uint32) if cands else np.zeros(0, np.uint32)
    off = np.zeros(len(queries) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(c) for c in cands])
    out = capi.sw_score(ctx, D, Q, ids, off, blosum, go, ge)
    Q.close(); D.close()
    return ids, off, out


def _check(queries, db, ids, off, out, blosum, go=10, ge=1):
    bad = 0
    for q in range(len(queries)):
        for i in range(off[q], off[q + 1]):
            exp = O.sw_score(queries[q], db[ids[i]], blosum, go, ge)
            if exp != out[i]:
                bad += 1
                assert bad < 5, "score mismatch q=%d t=%d gpu=%d oracle=%d" % (q, ids[i], out[i], exp)
    assert bad == 0


def test_random_queries_all_length_classes(ctx, blosum):
    rng = np.random.default_rng(7)
    # one query per row-class K = 2..32 (length <= 1024) incl. boundaries
    qlens = [1, 2, 5, 31, 32, 33, 64, 65, 100, 127, 128, 129, 191, 200, 256, 257, 300, 333, 384, 400, 449, 512, 513, 600,
             640, 700, 768, 800, 896, 960, 1000, 1023, 1024]
    queries = [synth.random_codes(rng, n, 0.02) for n in qlens]
    db = [synth.random_codes(rng, n, 0.02) for n in rng.integers(1, 700, size=400)]
    # plant relatives so scores are not all tiny
    for i, q in enumerate(queries):
        db[(7 * i) % len(db)] = synth.mutate(rng, q, identity=0.8)
        db[(7 * i + 3) % len(db)] = np.concatenate([synth.random_codes(rng, 20), synth.mutate(rng, q, identity=0.5), synth.random_codes(rng, 33)])
    cands = [np.sort(rng.choice(len(db), size=int(rng.integers(1, 40)), replace=False)) for _ in queries]
    for i in range(len(queries)):
        cands[i] = np.unique(np.concatenate([cands[i], [(7 * i) % len(db), (7 * i + 3) % len(db)]]))
    ids, off, out = _run(ctx, blosum, queries, db, cands)
    _check(queries, db, ids, off, out, blosum)


def test_odd_candidate_counts_and_single_pair(ctx, blosum):
    rng = np.random.default_rng(11)
    queries = [synth.random_codes(rng, 150), synth.random_codes(rng, 77)]
    db = [synth.random_codes(rng, n) for n in (5, 1, 300, 1200, 64)]
    cands = [np.array([3]), np.array([0, 1, 2, 3, 4])]
    ids, off, out = _run(ctx, blosum, queries, db, cands)
    _check(queries, db, ids, off, out, blosum)


def test_empty_candidate_list_for_some_queries(ctx, blosum):
    rng = np.random.default_rng(12)
    queries = [synth.random_codes(rng, 90) for _ in range(4)]
    db = [synth.random_codes(rng, 120) for _ in range(10)]
    cands = [np.array([], dtype=np.int64), np.array([1, 2]), np.array([], dtype=np.int64), np.array([9])]
    ids, off, out = _run(ctx, blosum, queries, db, cands)
    _check(queries, db, ids, off, out, blosum)


def test_rare_letters_and_other_gap_penalties(ctx, blosum):
    rng = np.random.default_rng(13)
    queries = [synth.random_codes(rng, 200, 0.2) for _ in range(3)]
    db = [synth.random_codes(rng, n, 0.2) for n in rng.integers(20, 400, size=60)]
    db[5] = synth.mutate(rng, queries[0], 0.7)
    cands = [np.arange(60) for _ in queries]
    for go, ge in ((10, 1), (11, 1), (5, 2), (1, 1), (20, 3)):
        ids, off, out = _run(ctx, blosum, queries, db, cands, go, ge)
        _check(queries, db, ids, off, out, blosum, go, ge)


def test_sixteen_bit_overflow_is_rerun_exactly(ctx, blosum):
    # a 7000-residue tryptophan-rich self hit scores far above 32767: the packed kernel must flag it and
    # the 32-bit kernel must return the exact value (swimd's 8->16->32 escalation, Swimd.cpp:412-449)
    rng = np.random.default_rng(14)
    big = synth.random_codes(rng, 1000)
    longt = np.concatenate([big] * 7)
    queries = [big, synth.random_codes(rng, 300)]
    db = [longt, synth.random_codes(rng, 400), big.copy()]
    cands = [np.array([0, 1, 2]), np.array([0, 1, 2])]
    ids, off, out = _run(ctx, blosum, queries, db, cands)
    _check(queries, db, ids, off, out, blosum)
    # and a genuinely > 32767 score: long query (32-bit multi-pass path)
    q2 = np.concatenate([big] * 7)
    ids, off, out = _run(ctx, blosum, [q2], [q2.copy(), longt[:5000]], [np.array([0, 1])])
    assert out[0] > 32767
    _check([q2], [q2.copy(), longt[:5000]], ids, off, out, blosum)


def test_long_queries_use_the_multipass_kernel(ctx, blosum):
    rng = np.random.default_rng(15)
    queries = [synth.random_codes(rng, n) for n in (1025, 2000, 3100)]
    db = [synth.random_codes(rng, n) for n in rng.integers(30, 900, size=20)]
    db[3] = synth.mutate(rng, queries[1][200:1500], 0.6)
    cands = [np.arange(20), np.arange(20), np.arange(0, 20, 3)]
    ids, off, out = _run(ctx, blosum, queries, db, cands)
    _check(queries, db, ids, off, out, blosum)


def test_device_pointer_path_matches_host_path(ctx, blosum):
    import torch
    rng = np.random.default_rng(16)
    queries = [synth.random_codes(rng, n) for n in (120, 480)]
    db = [synth.random_codes(rng, n) for n in rng.integers(30, 500, size=50)]
    cands = [np.arange(50), np.arange(0, 50, 2)]
    ids, off, out = _run(ctx, blosum, queries, db, cands)
    qc, qo = synth.pack(queries); dc, do = synth.pack(db)
    D = ctx.database(dc, do); Q = ctx.queries(qc, qo)
    dev = torch.device("cuda:0")
    t_ids = torch.from_numpy(ids.astype(np.int64)).to(dev).to(torch.int32)   # same bits as uint32
    t_off = torch.from_numpy(off).to(dev)
    t_out = torch.zeros(len(ids), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    capi.sw_score(ctx, D, Q, t_ids, t_off, blosum, out=t_out, where=capi.S4G_DEVICE)
    ctx.sync()
    assert np.array_equal(t_out.cpu().numpy(), out)
    Q.close(); D.close()
