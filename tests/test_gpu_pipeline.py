"""GPU: the device-resident pipeline (GPU E-value screen + host exact selection) returns exactly what the
host-buffer path returns, and the multi-shard candidate merge equals the single-shard prefilter."""
import numpy as np
import pytest

from oracle import oracle as O
from sift4g_b200 import capi, pipeline, synth

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
