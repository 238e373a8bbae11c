"""Per-stage timing of the striped (NVLink) multi-GPU form at a bench.py shape, under torchrun (not the bench):

    S4G_TRACE=1 python -m torch.distributed.run --nproc-per-node 2 ... tools/perf_striped.py [--stage prefilter|all]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--queries", type=int, default=1000, help="per rank")
ap.add_argument("--db-seqs", type=int, default=10_000_000)
ap.add_argument("--max-candidates", type=int, default=5000)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--stage", default="all")
args = ap.parse_args()

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from sift4g_b200 import capi, pipeline, stripes  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = capi.Context(local)
mat = np.array(bench.BLOSUM62_A_TO_Z, dtype=np.int32)
nq = args.queries * world
q_codes, q_off = bench.make_queries(nq)
n_db = args.db_seqs
lo, hi = n_db * rank // world, n_db * (rank + 1) // world
codes, loc_off, lens, total_res = bench.build_db_device(torch, dev, n_db, lo, hi, q_codes, q_off)
all_lens = bench.db_lengths(n_db)
all_off = np.zeros(n_db + 1, dtype=np.int64)
np.cumsum(all_lens, out=all_off[1:])
S = stripes.StripedDatabase(ctx, codes, all_off, lo, hi, dist=dist)
del codes
qa, qb = nq * rank // world, nq * (rank + 1) // world
pipe = pipeline.DevicePipeline(ctx, S.db, q_codes[q_off[qa]:q_off[qb]], q_off[qa:qb + 1] - q_off[qa], mat, all_lens, total_res, max_candidates=args.max_candidates)
for it in range(args.iters):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.time()
    if args.stage == "prefilter":
        capi.prefilter(ctx, S.db, pipe.Q, pipe.k, pipe.N, True, out=(pipe.t_ids, pipe.t_sc, pipe.t_cnt), where=capi.S4G_DEVICE)
    else:
        pipe.step()
    torch.cuda.synchronize()
    print("rank %d iter %d: %.3f ms" % (rank, it, (time.time() - t0) * 1e3), file=sys.stderr)
pipe.close()
S.close()
dist.barrier()
dist.destroy_process_group()
