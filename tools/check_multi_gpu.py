#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU): the sharded pipeline (one database shard per rank,
query-owner cut-off exchange, overflow-only hit merge) must give exactly the single-GPU answer: same candidate sets,
same kept hits with bit-identical E-values, same alignment cells and paths.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_multi_gpu.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def collect(r, nq):
    """host view of one rank's step result: per query candidate ids, hits as tuples, paths"""
    cand_ids = r.cand_ids.cpu().numpy().view(np.uint32)
    cand_off = r.cand_off.cpu().numpy()
    cands = [cand_ids[cand_off[q]:cand_off[q + 1]] for q in range(nq)]
    hits = []
    if r.coords is not None:
        coords = r.coords.cpu().numpy(); poff = r.path_off.cpu().numpy(); paths = r.paths[:int(poff[-1])].cpu().numpy()
    for h in range(len(r.pair_q)):
        hits.append((int(r.pair_q[h]), float(r.evalue[h]).hex(), int(r.pair_score[h]), int(r.pair_t[h]), tuple(coords[h].tolist()),
                     paths[poff[h]:poff[h + 1]].tobytes()))
    return cands, hits


def main():
    import torch
    import torch.distributed as dist
    from sift4g_b200 import capi, pipeline, synth
    import bench

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = capi.Context(local)
    mat = np.array(bench.BLOSUM62_A_TO_Z, dtype=np.int32)
    ok = True
    for seed, nq, n_db, N, M in ((51, 13, 6000, 200, 400), (52, 9, 4000, 300, 6), (53, 5, 900, 2000, 400)):
        queries, db = synth.make_dataset(seed, nq, n_db, q_len=(50, 400), homologs=(8, 25), rare_fraction=0.005)
        qc, qo = synth.pack(queries); dc, do = synth.pack(db)
        lens = np.diff(do)
        lo, hi = n_db * rank // world, n_db * (rank + 1) // world
        D = ctx.database(dc[do[lo]:do[hi]], do[lo:hi + 1] - do[lo], id_base=lo)
        pipe = pipeline.DevicePipeline(ctx, D, qc, qo, mat, lens[lo:hi], int(do[-1]), max_candidates=N, max_alignments=M, dist=dist)
        r = pipe.step()
        r = pipe.step(e2e=True)
        mine = collect(r, nq)
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        pipe.close(); D.close()
        if rank == 0:
            D1 = ctx.database(dc, do)
            p1 = pipeline.DevicePipeline(ctx, D1, qc, qo, mat, lens, int(do[-1]), max_candidates=N, max_alignments=M)
            c1, h1 = collect(p1.step(), nq)
            p1.close(); D1.close()
            for q in range(nq):
                got = np.sort(np.concatenate([p[0][q] for p in parts]))
                if not np.array_equal(got, np.sort(c1[q])):
                    ok = False
                    print("seed %d query %d: candidate sets differ (%d vs %d)" % (seed, q, len(got), len(c1[q])))
            got_hits = sorted(h for p in parts for h in p[1])
            if got_hits != sorted(h1):
                ok = False
                print("seed %d: kept hits / alignments differ (%d vs %d)" % (seed, len(got_hits), len(h1)))
            # global order inside a query: ranks keep the global (E, score desc, id) order restricted to their targets
            print("seed %d: %d queries, %d candidates, %d hits compared over %d shards" % (seed, nq, sum(len(c) for c in c1), len(h1), world))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()
    if rank == 0:
        print("multi-GPU parity: %s" % ("OK" if int(flag.item()) else "FAILED"))
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
