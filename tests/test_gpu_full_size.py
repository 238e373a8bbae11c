"""GPU, BASELINE.json configs[1] at FULL size (1,000 queries x 10 M sequences / 3.2 B residues, bench.py's workload) and
a configs[2]-SHAPED batch (20,000 human-proteome-shaped queries -- log-normal lengths, 8 % beyond 1,024 aa: the striped
kernel; the prefilter scans it in groups of queries -- against 1 M sequences): size-independent properties of every stage
of the hot path, plus spot checks against the oracle on samples the CPU restatement finishes in seconds.

  stage 1  every list has max_candidates strictly ascending ids; two half shards merged = the single shard;
           sampled candidate scores equal the oracle's float bit for bit and sampled non-candidates lose to the cut-off
           under (score desc, id asc)
  stage 2  sampled SW scores equal the oracle's
  select   a sampled query's kept hits are exactly the oracle's best 400 with E <= 1e-4 under dbAlignmentDataCmp
           (sw/database.c:1043-1059)
  stage 3  EVERY kept hit: its path, re-scored cell by cell under BLOSUM62 10/1, gives the pair's SW score, and it consumes
           exactly the query / target spans its cells name
"""
import numpy as np
import pytest

from oracle import oracle as O
from sift4g_b200 import capi, pipeline

pytestmark = pytest.mark.gpu

SHAPES = {"configs1": (1000, 10_000_000, 5000, "uniform"), "configs2_shaped": (20000, 1_000_000, 5000, "human")}


@pytest.fixture(scope="module", params=list(SHAPES))
def c2(request, ctx, blosum):
    import torch
    import bench
    N_QUERIES, N_DB, N_CAND, shape = SHAPES[request.param]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    q_codes, q_off = bench.make_queries(N_QUERIES, shape=shape)
    codes, loc_off, lens, total_res = bench.build_db_device(torch, dev, N_DB, 0, N_DB, q_codes, q_off)
    db = ctx.database(codes, loc_off, id_base=0, where=capi.S4G_DEVICE)
    pipe = pipeline.DevicePipeline(ctx, db, q_codes, q_off, blosum, lens, total_res, max_candidates=N_CAND)
    r = pipe.step()
    torch.cuda.synchronize()
    # best-first rows with their scores (the step itself asks for id order)
    ids = torch.zeros((N_QUERIES, N_CAND), dtype=torch.int32, device=dev)
    sc = torch.zeros((N_QUERIES, N_CAND), dtype=torch.float32, device=dev)
    cnt = torch.zeros(N_QUERIES, dtype=torch.int32, device=dev)
    capi.prefilter(ctx, db, pipe.Q, 5, N_CAND, False, out=(ids, sc, cnt), where=capi.S4G_DEVICE)
    ctx.sync()
    yield dict(nq=N_QUERIES, n_db=N_DB, n_cand=N_CAND, torch=torch, dev=dev, q_codes=q_codes, q_off=q_off, codes=codes, off=loc_off, lens=lens, total=total_res, db=db,
               pipe=pipe, r=r, best_ids=ids.cpu().numpy().view(np.uint32), best_sc=sc.cpu().numpy(), best_cnt=cnt.cpu().numpy())
    pipe.close()
    db.close()
    del codes, loc_off, ids, sc, cnt, r
    torch.cuda.empty_cache()


def _seq(c2, i):
    a, b = int(c2["off"][i].item()), int(c2["off"][i + 1].item())
    return c2["codes"][a:b].cpu().numpy()


def _query(c2, q):
    return c2["q_codes"][c2["q_off"][q]:c2["q_off"][q + 1]]


def test_candidate_lists_are_full_ascending_and_in_range(c2):
    N_QUERIES, N_DB, N_CAND = c2["nq"], c2["n_db"], c2["n_cand"]
    r = c2["r"]
    off = r.cand_off.cpu().numpy()
    ids = r.cand_ids.cpu().numpy().view(np.uint32).astype(np.int64)
    cnt = np.diff(off)
    # a list is full unless fewer sequences than max_candidates share a k-mer with the query (very short queries of the
    # human-shaped batch); configs[1]'s queries all fill theirs
    assert (cnt <= N_CAND).all() and (cnt > 0).all() and np.array_equal(cnt, c2["best_cnt"])
    assert (cnt == N_CAND).mean() > (0.999 if N_QUERIES == 1000 else 0.98)
    assert ids.min() >= 0 and ids.max() < N_DB
    d = np.diff(ids)
    d[off[1:-1] - 1] = 1                                   # boundaries between queries
    assert (d > 0).all(), "candidate ids must be strictly ascending per query (database_search.cpp:173-180)"
    # the id-ordered lists and the best-first rows hold the same sets
    inside = np.arange(N_CAND)[None, :] < cnt[:, None]
    rows = np.where(inside, c2["best_ids"].astype(np.int64), np.int64(1) << 40)
    assert np.array_equal(np.sort(rows, axis=1)[inside], ids)
    assert (np.diff(c2["best_sc"], axis=1)[inside[:, 1:]] <= 0).all()


def test_two_half_shards_merge_to_the_single_shard_lists(ctx, c2):
    N_QUERIES, N_DB, N_CAND = c2["nq"], c2["n_db"], c2["n_cand"]
    torch, dev = c2["torch"], c2["dev"]
    W, half = 2, N_DB // 2
    g_ids = torch.zeros((W, N_QUERIES, N_CAND), dtype=torch.int32, device=dev)
    g_sc = torch.zeros((W, N_QUERIES, N_CAND), dtype=torch.float32, device=dev)
    g_cnt = torch.zeros((W, N_QUERIES), dtype=torch.int32, device=dev)
    for rk in range(W):
        lo, hi = rk * half, (rk + 1) * half
        a, b = int(c2["off"][lo].item()), int(c2["off"][hi].item())
        shard = ctx.database(c2["codes"][a:b + 64].contiguous(), (c2["off"][lo:hi + 1] - a).contiguous(), id_base=lo, where=capi.S4G_DEVICE)
        capi.prefilter(ctx, shard, c2["pipe"].Q, 5, N_CAND, False, out=(g_ids[rk], g_sc[rk], g_cnt[rk]), where=capi.S4G_DEVICE)
        ctx.sync()
        shard.close()
    o_ids = torch.zeros((N_QUERIES, N_CAND), dtype=torch.int32, device=dev)
    o_sc = torch.zeros((N_QUERIES, N_CAND), dtype=torch.float32, device=dev)
    o_cnt = torch.zeros(N_QUERIES, dtype=torch.int32, device=dev)
    ctx.check(ctx.lib.s4g_merge_candidates(ctx.h, W, N_QUERIES, N_CAND, g_ids.data_ptr(), g_sc.data_ptr(), g_cnt.data_ptr(),
                                           o_ids.data_ptr(), o_sc.data_ptr(), o_cnt.data_ptr()))
    ctx.sync()
    cnt = c2["best_cnt"].astype(np.int64)
    assert np.array_equal(o_cnt.cpu().numpy().astype(np.int64), cnt)
    inside = np.arange(N_CAND)[None, :] < cnt[:, None]
    assert np.array_equal(o_ids.cpu().numpy().view(np.uint32)[inside], c2["r"].cand_ids.cpu().numpy().view(np.uint32))


def test_one_call_search_returns_the_pipelines_results(ctx, c2, blosum):
    """s4g_search (the host-buffer entry point bench.py times as e2e) against the stage-by-stage device pipeline whose
    outputs the other tests of this module check against the oracle: same candidate lists, kept hits, E-value bits,
    cells and path bytes."""
    N_QUERIES, N_DB, N_CAND = c2["nq"], c2["n_db"], c2["n_cand"]
    r = c2["r"]
    out = capi.search(ctx, c2["db"], c2["pipe"].Q, blosum, max_candidates=N_CAND)
    assert out.n_pairs == r.n_pairs and out.sw_cells == r.sw_cells
    assert np.array_equal(out.cand_off, r.cand_off.cpu().numpy())
    assert np.array_equal(out.cand_ids, r.cand_ids.cpu().numpy().view(np.uint32))
    assert np.array_equal(out.pair_q, r.pair_q) and np.array_equal(out.pair_t, r.pair_t) and np.array_equal(out.pair_score, r.pair_score)
    assert np.array_equal(out.evalue, r.evalue) and np.array_equal(out.hit_off, r.hit_off)
    assert np.array_equal(out.coords, r.coords.cpu().numpy())
    assert np.array_equal(out.path_off, r.path_off.cpu().numpy())
    assert np.array_equal(out.paths, r.paths[:int(out.path_off[-1])].cpu().numpy())


def test_sampled_prefilter_scores_match_the_oracle(c2):
    N_QUERIES, N_DB, N_CAND = c2["nq"], c2["n_db"], c2["n_cand"]
    rng = np.random.default_rng(7)
    full = np.nonzero(c2["best_cnt"] == N_CAND)[0]
    for q in rng.choice(full, size=4, replace=False):
        row_ids, row_sc = c2["best_ids"][q], c2["best_sc"][q]
        kept = set(row_ids.tolist())
        pick_in = rng.choice(N_CAND, size=60, replace=False)
        pick_in[0], pick_in[1] = 0, N_CAND - 1                          # the best and the cut-off row
        outside = [int(i) for i in rng.choice(N_DB, size=80, replace=False) if int(i) not in kept][:60]
        sample = [int(row_ids[i]) for i in pick_in] + outside
        seqs = [_seq(c2, i) for i in sample]
        off = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum([len(s) for s in seqs], out=off[1:])
        qc = _query(c2, q)
        _, _, _, dense = O.prefilter(np.concatenate(seqs), off, qc, np.array([0, len(qc)], dtype=np.int64), 5, len(seqs), dense=True)
        dense = np.asarray(dense, dtype=np.float32).reshape(-1)
        for n, i in enumerate(pick_in):
            assert dense[n] == row_sc[i], "prefilter score of (query %d, sequence %d)" % (q, sample[n])
        cut_sc, cut_id = float(row_sc[N_CAND - 1]), int(row_ids[N_CAND - 1])
        for n, sid in enumerate(outside):
            s = float(dense[len(pick_in) + n])
            assert s < cut_sc or (s == cut_sc and sid > cut_id), "sequence %d should have displaced the cut-off row of query %d" % (sid, q)


def test_sampled_sw_scores_match_the_oracle(c2, blosum):
    N_QUERIES, N_DB, N_CAND = c2["nq"], c2["n_db"], c2["n_cand"]
    r = c2["r"]
    rng = np.random.default_rng(8)
    ids = r.cand_ids.cpu().numpy().view(np.uint32)
    off = r.cand_off.cpu().numpy()
    scores = r.scores.cpu().numpy()[:len(ids)]
    qlen = np.diff(c2["q_off"])
    for k in rng.choice(len(ids), size=1500, replace=False):
        q = int(np.searchsorted(off, k, side="right")) - 1
        assert scores[k] == O.sw_score(_query(c2, q), _seq(c2, int(ids[k])), blosum), "SW score of (query %d, sequence %d)" % (q, ids[k])
    # the strongest pairs too (planted homologs, long alignments)
    for k in np.argsort(scores)[-40:]:
        q = int(np.searchsorted(off, k, side="right")) - 1
        assert scores[k] == O.sw_score(_query(c2, q), _seq(c2, int(ids[k])), blosum)
    # and pairs of the longest queries (beyond 1024 aa: the intra-sequence striped kernel)
    for q in np.argsort(qlen)[-6:]:
        for k in rng.choice(np.arange(off[q], off[q + 1]), size=12, replace=False):
            assert scores[k] == O.sw_score(_query(c2, int(q)), _seq(c2, int(ids[k])), blosum), "SW score of (query %d of %d aa, sequence %d)" % (q, qlen[q], ids[k])


def test_kept_hits_of_sampled_queries_are_the_oracles_selection(c2):
    N_QUERIES, N_DB, N_CAND = c2["nq"], c2["n_db"], c2["n_cand"]
    r = c2["r"]
    ids = r.cand_ids.cpu().numpy().view(np.uint32)
    off = r.cand_off.cpu().numpy()
    scores = r.scores.cpu().numpy()
    lens, total = c2["lens"], c2["total"]
    hoff = r.hit_off
    assert (np.diff(hoff) <= 400).all() and (r.evalue <= 1e-4).all()
    rng = np.random.default_rng(9)
    busiest = int(np.argmax(np.diff(hoff)))
    for q in [busiest] + [int(x) for x in rng.choice(N_QUERIES, size=5, replace=False)]:
        qlen = int(c2["q_off"][q + 1] - c2["q_off"][q])
        rows = []
        for k in range(int(off[q]), int(off[q + 1])):
            if scores[k] < 60:
                continue                                   # E(60) ~ 1e4 for these lengths: far above 1e-4; keeps the libm loop short
            e = O.evalue(int(scores[k]), qlen, int(lens[ids[k]]), total)
            if e <= 1e-4:
                rows.append((e, -int(scores[k]), int(ids[k])))
        rows.sort()                                         # names are ">D%08d": name order = id order
        rows = rows[:400]
        got = [(float(r.evalue[h]), -int(r.pair_score[h]), int(r.pair_t[h])) for h in range(hoff[q], hoff[q + 1])]
        assert got == rows, "kept hits of query %d" % q


def test_every_path_rescored_gives_the_sw_score(c2, blosum):
    N_QUERIES, N_DB, N_CAND = c2["nq"], c2["n_db"], c2["n_cand"]
    torch, dev, r = c2["torch"], c2["dev"], c2["r"]
    n = len(r.pair_q)
    assert n > 50_000
    poff = r.path_off
    total = int(poff[-1].item())
    op = r.paths[:total].to(torch.int64)
    plen = poff[1:] - poff[:-1]
    hid = torch.repeat_interleave(torch.arange(n, device=dev), plen)
    coords = r.coords.to(torch.int64)
    qs, qe, ts, te = coords[:, 0], coords[:, 1], coords[:, 2], coords[:, 3]
    qadv = ((op == 1) | (op == 3)).to(torch.int64)
    tadv = ((op == 1) | (op == 2)).to(torch.int64)
    assert int(((op < 1) | (op > 3)).sum().item()) == 0

    def seg_sum(x):
        cs = torch.zeros(total + 1, dtype=torch.int64, device=dev)
        cs[1:] = torch.cumsum(x, 0)
        return cs[poff[1:]] - cs[poff[:-1]], cs

    nq_used, cq = seg_sum(qadv)
    nt_used, ct = seg_sum(tadv)
    # banded_sw walks back from the end cell `while (i > 0)` (ssw.c:634-706): every path consumes its whole query span;
    # when co-optimal alignments exist it may reach row 0 right of the begin column the reverse sweep reported, so the
    # target span is only bounded.  Paths are therefore anchored at their END cell for re-scoring.
    assert torch.equal(nq_used, qe - qs + 1), "paths must consume exactly their query span"
    assert bool((nt_used <= te - ts + 1).all()), "paths must stay inside their target span"
    n_short = int((nt_used != te - ts + 1).sum().item())
    assert n_short <= n // 50, "%d of %d paths stop right of their begin column" % (n_short, n)
    first = poff[:-1][hid]
    qi = (qe + 1 - nq_used)[hid] + (cq[:-1] - cq[first])   # residue this op consumes (valid where it advances)
    ti = (te + 1 - nt_used)[hid] + (ct[:-1] - ct[first])
    pq = torch.from_numpy(r.pair_q.astype(np.int64)).to(dev)
    pt = torch.from_numpy(r.pair_t.astype(np.int64)).to(dev)
    q_off = torch.from_numpy(c2["q_off"]).to(dev)
    q_codes = torch.from_numpy(c2["q_codes"]).to(dev).to(torch.int64)
    qa = q_codes[(q_off[pq][hid] + qi).clamp_(max=q_codes.numel() - 1)]
    ta = c2["codes"][(c2["off"][pt][hid] + ti).clamp_(max=c2["codes"].numel() - 1)].to(torch.int64)
    mat = torch.from_numpy(np.ascontiguousarray(blosum, dtype=np.int64).reshape(-1)).to(dev)
    sub = mat[qa * 26 + ta]
    prev = torch.roll(op, 1)
    idx = torch.arange(total, device=dev)
    opens = (idx == first) | (prev != op)
    cell = torch.where(op == 1, sub, torch.where(opens, torch.full_like(sub, -10), torch.full_like(sub, -1)))
    path_score, _ = seg_sum(cell)
    want = torch.from_numpy(r.pair_score.astype(np.int64)).to(dev)
    bad = int((path_score != want).sum().item())
    short = torch.nonzero(nt_used != te - ts + 1).reshape(-1)
    odd = torch.nonzero(path_score != want).reshape(-1)
    print("full-size paths: %d hits, %d stop right of the begin column, %d do not re-score to the SW score" % (n, n_short, int(odd.numel())))
    # The few paths that do not re-score are the reference's own: banded_sw closes whatever is left at row 0 as a match
    # (ssw.c:700-706).  They, and a sample of the short ones, must be byte-identical to the oracle's restatement of SSW.
    assert odd.numel() <= n // 1000, "%d of %d paths do not re-score to their SW score" % (int(odd.numel()), n)
    check = odd.cpu().numpy().tolist() + short.cpu().numpy().tolist()[:40]
    h_coords, h_poff = r.coords.cpu().numpy(), poff.cpu().numpy()
    for h in check:
        q, t = int(r.pair_q[h]), int(r.pair_t[h])
        o_coords, o_path = O.align(_query(c2, q), _seq(c2, t), int(r.pair_score[h]), blosum)
        assert np.array_equal(o_coords, h_coords[h]), "cells of hit %d (query %d, sequence %d)" % (h, q, t)
        assert np.array_equal(o_path, r.paths[h_poff[h]:h_poff[h + 1]].cpu().numpy()), "path of hit %d (query %d, sequence %d)" % (h, q, t)
    # alignments begin and end on a match (local alignment; ssw.c:634-706 closes the path with M)
    assert int((op[poff[:-1]] != 1).sum().item()) == 0 and int((op[poff[1:] - 1] != 1).sum().item()) == 0
