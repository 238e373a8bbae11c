"""GPU: the device-resident pipeline (GPU E-value screen + host exact selection) returns exactly what the
host-buffer path returns, and the multi-shard candidate merge equals the single-shard prefilter."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from sift4g_b200 import capi, pipeline, synth
from tests.util import GOLDEN

pytestmark = pytest.mark.gpu


def test_device_pipeline_equals_host_pipeline(ctx, blosum):
    queries, db = synth.make_dataset(41, 8, 3000, q_len=(60, 500), homologs=(8, 25), rare_fraction=0.005)
    qc, qo = synth.pack(queries); dc, do = synth.pack(db)
    lens = np.diff(do)
    D = ctx.database(dc, do)
    rh = pipeline.run_host(ctx, D, qc, qo, blosum, lens, max_candidates=300)
    pipe = pipeline.DevicePipeline(ctx, D, qc, qo, blosum, lens, int(do[-1]), max_candidates=300)
    rd = pipe.step()
    assert rd.sw_cells == rh.sw_cells and rd.n_pairs == rh.n_pairs
    assert np.array_equal(rd.pair_q, rh.pair_q) and np.array_equal(rd.pair_t, rh.pair_t)
    assert np.array_equal(rd.pair_score, rh.pair_score)
    assert np.array_equal(rd.evalue, rh.evalue)             # bit-identical doubles
    assert len(rh.pair_q) > 50
    assert np.array_equal(rd.coords.cpu().numpy(), rh.coords)
    n = int(rd.path_off[-1].item())
    assert np.array_equal(rd.path_off.cpu().numpy(), rh.path_off)
    assert np.array_equal(rd.paths[:n].cpu().numpy(), rh.paths[:n])
    # e2e flavour returns the same thing
    re_ = pipe.step(e2e=True)
    assert np.array_equal(re_.pair_t, rh.pair_t) and re_.d2h_bytes > 0 and re_.h2d_bytes > 0
    # every kept E-value equals the oracle's libm evaluation
    for h in range(0, len(rh.pair_q), 7):
        q, t = int(rh.pair_q[h]), int(rh.pair_t[h])
        assert rh.evalue[h] == O.evalue(int(rh.pair_score[h]), len(queries[q]), len(db[t]), int(do[-1]))
    pipe.close(); D.close()


def test_one_call_search_matches_the_oracle_stage_by_stage(ctx, blosum):
    """s4g_search (prefilter -> scores -> selection -> traceback in one C-ABI call, host buffers out) against the oracle:
    candidate sets, the kept hits of EVERY query (libm E-values, order of dbAlignmentDataCmp), cells and path bytes."""
    queries, db = synth.make_dataset(43, 9, 2500, q_len=(50, 600), homologs=(8, 30), rare_fraction=0.005)
    qc, qo = synth.pack(queries); dc, do = synth.pack(db)
    lens = np.diff(do)
    total = int(do[-1])
    D = ctx.database(dc, do)
    N, M = 250, 12                                  # M below the homolog count: the top-M truncation is exercised
    out = pipeline.search_host(ctx, D, qc, qo, blosum, max_candidates=N, max_alignments=M)
    cells, oids, osc, _ = O.prefilter(dc, do, qc, qo, 5, N)
    assert cells == total == out.db_residues
    n_hits = 0
    for q in range(len(queries)):
        got = out.cand_ids[out.cand_off[q]:out.cand_off[q + 1]]
        assert np.array_equal(got, oids[q]), "candidate set of query %d" % q
        # oracle selection over the oracle's scores of these candidates
        sc = np.array([O.sw_score(queries[q], db[t], blosum) for t in oids[q]], dtype=np.int32)
        ev = np.array([O.evalue(int(s), len(queries[q]), int(lens[t]), total) for s, t in zip(sc, oids[q])])
        keep = O.select(ev, sc, ["D%08d" % t for t in oids[q]], 1e-4, M)
        a, b = int(out.hit_off[q]), int(out.hit_off[q + 1])
        assert np.array_equal(out.pair_t[a:b], oids[q][keep]), "kept hits of query %d" % q
        assert np.array_equal(out.pair_score[a:b], sc[keep]) and np.array_equal(out.evalue[a:b], ev[keep])
        assert (out.pair_q[a:b] == q).all()
        for h in range(a, b):
            coords, path = O.align(queries[q], db[int(out.pair_t[h])], int(out.pair_score[h]), blosum)
            assert np.array_equal(coords, out.coords[h])
            assert np.array_equal(path, out.paths[out.path_off[h]:out.path_off[h + 1]])
        n_hits += b - a
    assert n_hits == out.n_hits and n_hits > 40
    assert out.sw_cells == sum(len(queries[q]) * int(lens[oids[q]].sum()) for q in range(len(queries)))
    assert out.d2h_bytes > 0 and out.h2d_bytes > 0
    # the stage-by-stage host path and the torch-resident pipeline return the same thing
    hits = (out.pair_q.copy(), out.pair_t.copy(), out.pair_score.copy(), out.evalue.copy(), out.coords.copy(), out.paths.copy(), out.path_off.copy())
    rh = pipeline.run_host(ctx, D, qc, qo, blosum, lens, max_candidates=N, max_alignments=M)
    assert np.array_equal(rh.pair_q, hits[0]) and np.array_equal(rh.pair_t, hits[1]) and np.array_equal(rh.pair_score, hits[2])
    assert np.array_equal(rh.evalue, hits[3]) and np.array_equal(rh.coords, hits[4]) and np.array_equal(rh.path_off, hits[6])
    assert np.array_equal(rh.paths[:int(rh.path_off[-1])], hits[5])
    D.close()


def test_score_screen_returns_every_pair_that_can_pass(ctx, blosum):
    """s4g_score_screen: the survivors are a superset of the pairs with E <= max_evalue (oracle doubles), in candidate order,
    with exact scores and target lengths; sw_cells is the algorithmic count."""
    queries, db = synth.make_dataset(44, 6, 1200, q_len=(50, 500), homologs=(5, 20), rare_fraction=0.005)
    qc, qo = synth.pack(queries); dc, do = synth.pack(db)
    lens = np.diff(do)
    total = int(do[-1])
    D = ctx.database(dc, do); Q = ctx.queries(qc, qo)
    rng = np.random.default_rng(3)
    cands = [np.sort(rng.choice(len(db), size=400, replace=False)).astype(np.uint32) for _ in queries]
    ids = np.concatenate(cands); off = np.arange(len(queries) + 1, dtype=np.int64) * 400
    s_q, s_id, s_sc, s_tl, cells = capi.score_screen(ctx, D, Q, ids, off, blosum)
    assert cells == sum(len(queries[q]) * int(lens[cands[q]].sum()) for q in range(len(queries)))
    want = []
    for q in range(len(queries)):
        for t in cands[q]:
            s = O.sw_score(queries[q], db[t], blosum)
            if O.evalue(s, len(queries[q]), int(lens[t]), total) <= 1e-4:
                want.append((q, int(t), s, int(lens[t])))
    got = list(zip(s_q.tolist(), s_id.tolist(), s_sc.tolist(), s_tl.tolist()))
    assert got == sorted(got, key=lambda r: (r[0], r[1])) and set(want) <= set(got) and len(want) > 10
    for q, t, s, tl in got:
        assert s == O.sw_score(queries[q], db[t], blosum) and tl == lens[t]
    assert len(got) <= len(want) + 5               # the screen's margin is 1e-6 relative: practically the same set
    Q.close(); D.close()


def test_sharded_prefilter_merge_equals_single_shard(ctx):
    import torch
    queries, db = synth.make_dataset(42, 6, 4000, q_len=(50, 400), homologs=(5, 20))
    qc, qo = synth.pack(queries); dc, do = synth.pack(db)
    N, nq, W = 150, len(queries), 3
    D = ctx.database(dc, do); Q = ctx.queries(qc, qo)
    ids1, sc1, cnt1 = capi.prefilter(ctx, D, Q, 5, N, sorted_by_id=True)
    D.close()
    dev = torch.device("cuda:0")
    g_ids = torch.zeros((W, nq, N), dtype=torch.int32, device=dev)
    g_sc = torch.zeros((W, nq, N), dtype=torch.float32, device=dev)
    g_cnt = torch.zeros((W, nq), dtype=torch.int32, device=dev)
    bounds = [0, 1300, 2500, 4000]
    for r in range(W):
        lo, hi = bounds[r], bounds[r + 1]
        Dr = ctx.database(dc[do[lo]:do[hi]], do[lo:hi + 1] - do[lo], id_base=lo)
        i, s, c = capi.prefilter(ctx, Dr, Q, 5, N, sorted_by_id=False)
        g_ids[r] = torch.from_numpy(i.view(np.int32)).to(dev); g_sc[r] = torch.from_numpy(s).to(dev); g_cnt[r] = torch.from_numpy(c.view(np.int32)).to(dev)
        Dr.close()
    o_ids = torch.zeros((nq, N), dtype=torch.int32, device=dev)
    o_sc = torch.zeros((nq, N), dtype=torch.float32, device=dev)
    o_cnt = torch.zeros(nq, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ctx.check(ctx.lib.s4g_merge_candidates(ctx.h, W, nq, N, g_ids.data_ptr(), g_sc.data_ptr(), g_cnt.data_ptr(), o_ids.data_ptr(), o_sc.data_ptr(), o_cnt.data_ptr()))
    ctx.sync()
    assert np.array_equal(o_cnt.cpu().numpy().view(np.uint32), cnt1)
    for q in range(nq):
        n = int(cnt1[q])
        assert np.array_equal(o_ids[q, :n].cpu().numpy().view(np.uint32), ids1[q, :n])
        assert np.array_equal(o_sc[q, :n].cpu().numpy(), sc1[q, :n])
    Q.close()


def test_query_owner_cutoff_protocol_equals_single_shard(ctx):
    """Multi-GPU stage-1 exchange on one device: rows of 3 shards -> (simulated all-to-all) -> s4g_topn_cutoff per owner ->
    (simulated all-gather) -> s4g_cutoff_counts per shard; the union of the surviving prefixes is the single-shard set."""
    import torch
    queries, db = synth.make_dataset(43, 7, 5000, q_len=(50, 400), homologs=(5, 20))
    qc, qo = synth.pack(queries); dc, do = synth.pack(db)
    nq, W = len(queries), 3
    S = (nq + W - 1) // W
    rows = S * W
    dev = torch.device("cuda:0")
    Q = ctx.queries(qc, qo)
    for N in (40, 200, 6000):                              # heavy cut-off ties / typical / union smaller than N
        D = ctx.database(dc, do)
        ids1, sc1, cnt1 = capi.prefilter(ctx, D, Q, 5, N, sorted_by_id=True)
        D.close()
        bounds = [0, 1700, 3100, 5000]
        loc = []
        for r in range(W):
            lo, hi = bounds[r], bounds[r + 1]
            Dr = ctx.database(dc[do[lo]:do[hi]], do[lo:hi + 1] - do[lo], id_base=lo)
            t_ids = torch.zeros((rows, N), dtype=torch.int32, device=dev)
            t_sc = torch.zeros((rows, N), dtype=torch.float32, device=dev)
            t_cnt = torch.zeros(rows, dtype=torch.int32, device=dev)
            capi.prefilter(ctx, Dr, Q, 5, N, False, out=(t_ids, t_sc, t_cnt), where=capi.S4G_DEVICE)
            ctx.sync()
            loc.append((t_ids, t_sc, t_cnt))
            Dr.close()
        cut_all = torch.zeros(rows, dtype=torch.int64, device=dev)
        for owner in range(W):                             # what all_to_all_single delivers to `owner`
            g_ids = torch.stack([loc[r][0][owner * S:(owner + 1) * S] for r in range(W)]).contiguous()
            g_sc = torch.stack([loc[r][1][owner * S:(owner + 1) * S] for r in range(W)]).contiguous()
            g_cnt = torch.stack([loc[r][2][owner * S:(owner + 1) * S] for r in range(W)]).contiguous()
            cut = torch.zeros(S, dtype=torch.int64, device=dev)
            ctx.check(ctx.lib.s4g_topn_cutoff(ctx.h, W, S, N, g_ids.data_ptr(), g_sc.data_ptr(), g_cnt.data_ptr(), cut.data_ptr()))
            ctx.sync()
            cut_all[owner * S:(owner + 1) * S] = cut
        got = [[] for _ in range(nq)]
        for r in range(W):
            t_ids, t_sc, t_cnt = loc[r]
            own = torch.zeros(nq, dtype=torch.int32, device=dev)
            ctx.check(ctx.lib.s4g_cutoff_counts(ctx.h, nq, N, t_ids.data_ptr(), t_sc.data_ptr(), t_cnt.data_ptr(), cut_all.data_ptr(), own.data_ptr()))
            ctx.sync()
            own = own.cpu().numpy(); h_ids = t_ids.cpu().numpy().view(np.uint32)
            assert (own <= t_cnt[:nq].cpu().numpy()).all()
            for q in range(nq):
                got[q].append(h_ids[q, :own[q]])
        for q in range(nq):
            g = np.sort(np.concatenate(got[q]))
            assert np.array_equal(g, ids1[q, :cnt1[q]]), "query %d, N %d" % (q, N)
    Q.close()


def test_fasta_reader_quirks(ctx, tmp_path):
    # sw/pre_proc.c:437-538: names trimmed, non-letters dropped, case folded, last byte of a file without trailing
    # newline is consumed as terminator
    p = tmp_path / "x.fa"
    p.write_bytes(b">  first seq description  \r\nACD-E*fg\nHI 12\n>second\nKLMNPQ")
    D = ctx.database_from_fasta(str(p))
    assert D.n_seqs == 2
    assert D.name(0) == "first seq description" and D.name(1) == "second"
    off = D.host_offsets(); codes = D.host_codes()
    assert "".join(chr(65 + c) for c in codes[off[0]:off[1]]) == "ACDEFGHI"
    assert "".join(chr(65 + c) for c in codes[off[1]:off[2]]) == "KLMNP"       # trailing Q consumed
    D.close()
    # sharded open keeps FASTA-order ids
    p2 = tmp_path / "y.fa"
    p2.write_text("".join(">s%d\n%s\n" % (i, "ACDEFGHIKL"[: 3 + i % 5]) for i in range(10)))
    A = ctx.database_from_fasta(str(p2), 0, 2); B = ctx.database_from_fasta(str(p2), 1, 2)
    assert A.n_seqs == 5 and B.n_seqs == 5 and A.id_base == 0 and B.id_base == 5 and B.name(0) == "s5"
    A.close(); B.close()


def test_packed_database_opens_like_the_fasta(ctx, tmp_path):
    # s4g_db_open_packed reads only the shard's byte ranges of the packed file; the resident shard must equal the one
    # s4g_db_open_fasta builds (ids, names, offsets, codes), for one shard and for three
    src = os.path.join(GOLDEN, "synth_e2e", "d.fa")
    packed = str(tmp_path / "d.s4gdb")
    capi.pack_fasta(src, packed)
    for n_shards in (1, 3):
        for s in range(n_shards):
            A = ctx.database_from_fasta(src, s, n_shards)
            B = ctx.database_from_file(packed, s, n_shards)
            assert (A.n_seqs, A.n_residues, A.id_base) == (B.n_seqs, B.n_residues, B.id_base)
            assert A.total_seqs == B.total_seqs and A.total_residues == B.total_residues
            assert np.array_equal(A.host_offsets(), B.host_offsets()) and np.array_equal(A.host_codes(), B.host_codes())
            assert all(A.name(i) == B.name(i) for i in range(A.n_seqs))
            A.close(); B.close()
    D = ctx.database_from_file(src)          # s4g_db_open on a FASTA falls through to the FASTA reader
    assert D.n_seqs == D.total_seqs > 0
    D.close()


def test_one_call_search_without_hits_and_with_lonely_queries(ctx, blosum):
    """s4g_search edge cases: a batch none of whose queries has a hit (no survivors of the screen: nothing to trace back), and a
    batch where only one query has homologs (the others contribute empty hit lists)."""
    rng = np.random.default_rng(77)
    db = [synth.random_codes(rng, int(l), 0.001) for l in rng.integers(40, 400, size=1500)]
    queries = [synth.random_codes(rng, int(l), 0.001) for l in (120, 333, 64)]
    qc, qo = synth.pack(queries); dc, do = synth.pack(db)
    D = ctx.database(dc, do)
    out = pipeline.search_host(ctx, D, qc, qo, blosum, max_candidates=100, max_alignments=10)
    _, ids_o, _, _ = O.prefilter(dc, do, qc, qo, 5, 100)      # random sequences: fewer than max_candidates share a k-mer chain with a query
    assert out.n_hits == 0 and out.n_pairs == sum(len(x) for x in ids_o) and int(out.hit_off[-1]) == 0 and int(out.path_off[-1]) == 0
    D.close()
    # plant homologs of query 1 only
    db2 = list(db)
    for i in range(6):
        db2[100 + 7 * i] = synth.mutate(rng, queries[1], identity=0.8)
    dc2, do2 = synth.pack(db2)
    D2 = ctx.database(dc2, do2)
    out = pipeline.search_host(ctx, D2, qc, qo, blosum, max_candidates=100, max_alignments=10)
    assert out.n_hits == 6 and list(np.diff(out.hit_off)) == [0, 6, 0]
    assert sorted(out.pair_t.tolist()) == [100 + 7 * i for i in range(6)]
    for h in range(6):
        coords, path = O.align(queries[1], db2[int(out.pair_t[h])], int(out.pair_score[h]), blosum)
        assert np.array_equal(coords, out.coords[h]) and np.array_equal(path, out.paths[out.path_off[h]:out.path_off[h + 1]])
    D2.close()
