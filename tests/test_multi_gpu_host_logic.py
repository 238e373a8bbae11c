"""CPU, world_size 2, gloo: the host-side exchange of the multi-GPU path (global top-M hit merge) gives every
rank exactly the hits it owns out of the single-process answer."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sift4g_b200 import pipeline


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _make_hits(nq, M, seed):
    rng = np.random.default_rng(seed)
    pq, pt, ps, ev, off = [], [], [], [], [0]
    for q in range(nq):
        n = int(rng.integers(0, 2 * M))
        ids = rng.choice(10000, size=n, replace=False)
        sc = rng.integers(40, 200, size=n)
        e = np.exp(-0.25 * sc) * rng.choice([1.0, 1.0, 2.0], size=n)     # ties in E on purpose
        pq += [q] * n; pt += ids.tolist(); ps += sc.tolist(); ev += e.tolist(); off.append(off[-1] + n)
    return (np.array(pq, np.uint32), np.array(pt, np.uint32), np.array(ps, np.int32), np.array(ev, np.float64), np.array(off, np.int64))


def _worker(rank, world, port, nq, M, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pq, pt, ps, ev, off = _make_hits(nq, M, 5)
    lo, hi = (0, 5000) if rank == 0 else (5000, 10000)
    # local = the hits this rank owns, already cut to its local top M (what select_hits returns per rank)
    lq, lt, ls, le, loff = [], [], [], [], [0]
    for qi in range(nq):
        a, b = off[qi], off[qi + 1]
        m = (pt[a:b] >= lo) & (pt[a:b] < hi)
        rows = np.stack([ev[a:b][m], ps[a:b][m].astype(np.float64), pt[a:b][m].astype(np.float64)], 1)
        order = np.lexsort((rows[:, 2], -rows[:, 1], rows[:, 0]))[:M]
        rows = rows[order]
        lq += [qi] * len(rows); lt += rows[:, 2].astype(np.uint32).tolist(); ls += rows[:, 1].astype(np.int32).tolist(); le += rows[:, 0].tolist()
        loff.append(loff[-1] + len(rows))
    out = pipeline.merge_hits(torch, dist, torch.device("cpu"), nq, M, np.array(lq, np.uint32), np.array(lt, np.uint32), np.array(ls, np.int32),
                              np.array(le, np.float64), np.array(loff, np.int64), lo, hi)
    q.put((rank, [x.tolist() for x in out]))
    dist.barrier()
    dist.destroy_process_group()


def test_merge_hits_world_size_2_gloo():
    nq, M, world = 7, 20, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nq, M, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pq, pt, ps, ev, off = _make_hits(nq, M, 5)
    for qi in range(nq):
        a, b = off[qi], off[qi + 1]
        rows = np.stack([ev[a:b], ps[a:b].astype(np.float64), pt[a:b].astype(np.float64)], 1)
        order = np.lexsort((rows[:, 2], -rows[:, 1], rows[:, 0]))[:M]
        want = rows[order]
        got = []
        for r in range(world):
            oq, ot, osc, oev, ooff = res[r]
            got += [(oev[i], osc[i], ot[i]) for i in range(ooff[qi], ooff[qi + 1])]
            lo, hi = (0, 5000) if r == 0 else (5000, 10000)
            assert all(lo <= ot[i] < hi for i in range(ooff[qi], ooff[qi + 1]))
        got.sort(key=lambda x: (x[0], -x[1], x[2]))
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert g[0] == w[0] and g[1] == w[1] and g[2] == w[2]


def test_native_merge_equals_numpy_statement():
    from sift4g_b200 import capi
    rng = np.random.default_rng(9)
    W, nq, M = 3, 11, 16
    counts = rng.integers(0, M + 1, size=(W, nq)).astype(np.int64)
    counts[1, 4] = 0
    stride = int(counts.sum(axis=1).max()) + 2
    allh = np.zeros((W, stride, 3), dtype=np.float64)
    for r in range(W):
        pos = 0
        for q in range(nq):
            n = int(counts[r, q])
            sc = rng.integers(40, 60, size=n)
            e = np.exp(-0.25 * sc)                      # equal scores -> equal E: exercises the tie keys
            ids = rng.choice(3000, size=n, replace=False) + 3000 * r
            order = np.lexsort((ids, -sc, e))
            allh[r, pos:pos + n, 0] = e[order]; allh[r, pos:pos + n, 1] = sc[order]; allh[r, pos:pos + n, 2] = ids[order]
            pos += n
    for lo, hi in ((0, 0xffffffff), (3000, 6000)):
        got = capi.merge_hits(None, allh, counts, nq, M, lo, hi, n_threads=3)
        want = pipeline.merge_hits_numpy(allh, counts, nq, M, lo, hi)
        for g, w in zip(got, want):
            assert np.array_equal(g, w)
