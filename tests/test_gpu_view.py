"""GPU: a database seen through a striped view (csrc/view.cu: several physical stripes mapped into one virtual range, one
of them imported through its file descriptor like a peer's) gives exactly the results of the plain resident database --
the kernels never see where a page lives.  On one GPU all stripes are on the same device; tools/check_multi_gpu.py runs the
same comparison with one stripe per GPU under torchrun."""
import numpy as np
import pytest

from sift4g_b200 import capi, pipeline, stripes, synth

pytestmark = pytest.mark.gpu


def _same(a, b):
    assert np.array_equal(a.cand_ids, b.cand_ids) and np.array_equal(a.cand_off, b.cand_off)
    assert np.array_equal(a.pair_q, b.pair_q) and np.array_equal(a.pair_t, b.pair_t) and np.array_equal(a.pair_score, b.pair_score)
    assert np.array_equal(a.evalue, b.evalue) and np.array_equal(a.coords, b.coords) and np.array_equal(a.path_off, b.path_off)
    assert np.array_equal(a.paths[:int(a.path_off[-1])], b.paths[:int(b.path_off[-1])])
    assert a.sw_cells == b.sw_cells and a.n_hits == b.n_hits


def test_view_database_equals_resident_database(ctx, blosum):
    import torch
    # ~7.5 MB of residues: three 2 MiB-granular stripes, sequences straddle the stripe boundaries
    queries, db = synth.make_dataset(61, 12, 26000, q_len=(60, 500), homologs=(8, 25), rare_fraction=0.005)
    qc, qo = synth.pack(queries); dc, do = synth.pack(db)
    D = ctx.database(dc, do)
    want = pipeline.search_host(ctx, D, qc, qo, blosum, max_candidates=400, max_alignments=50)
    D.close()
    d_codes = torch.from_numpy(dc).cuda()
    S = stripes.StripedDatabase(ctx, d_codes, do, 0, len(db), emulate=3)
    assert len(S.stripes) == 3 and S.view.nbytes == S.bounds[-1] >= int(do[-1]) + 256
    assert S.db.n_seqs == len(db) and S.db.n_residues == int(do[-1]) and S.db.id_base == 0
    got = pipeline.search_host(ctx, S.db, qc, qo, blosum, max_candidates=400, max_alignments=50)
    _same(got, want)
    assert want.n_hits > 100
    # the bytes really sit in the view (read back through a plain device copy of a range that crosses a stripe boundary)
    lo = S.bounds[1] - 1000
    back = torch.empty(2000, dtype=torch.uint8, device="cuda")
    import ctypes
    rt = ctypes.CDLL("libcudart.so")
    assert rt.cudaMemcpy(ctypes.c_void_p(back.data_ptr()), ctypes.c_void_p(S.view.ptr + lo), ctypes.c_size_t(2000), 3) == 0
    assert np.array_equal(back.cpu().numpy(), dc[lo:lo + 2000])
    S.close()


def test_view_rejects_bad_arguments(ctx):
    g = int(ctx.lib.s4g_stripe_granularity(ctx.h))
    assert g > 0 and g % 4096 == 0
    with pytest.raises(capi.S4GError):
        capi.Stripe(ctx, g + 1)                     # not a multiple of the granularity
    s = capi.Stripe(ctx, g)
    v = capi.View(ctx, [s], [g])
    with pytest.raises(capi.S4GError):
        v.database(np.array([0, g], dtype=np.int64))   # no room for the readable pad
    with pytest.raises(capi.S4GError):
        v.write(g - 10, 0, 100)
    v.close(); s.free()
