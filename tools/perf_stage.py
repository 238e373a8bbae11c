"""Per-phase timing of the hot path at a bench.py shape (not the bench): run with S4G_TRACE=1 to get the
library's own phase split on stderr.

    S4G_TRACE=1 python tools/perf_stage.py [--queries 1000] [--db-seqs 10000000] [--iters 3] [--stage prefilter|all]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--queries", type=int, default=1000)
ap.add_argument("--db-seqs", type=int, default=10_000_000)
ap.add_argument("--max-candidates", type=int, default=5000)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--stage", default="all")
ap.add_argument("--query-shape", default="uniform", choices=["uniform", "human"])
args = ap.parse_args()

import torch  # noqa: E402
from sift4g_b200 import capi, pipeline  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
ctx = capi.Context(0)
mat = np.array(bench.BLOSUM62_A_TO_Z, dtype=np.int32)
q_codes, q_off = bench.make_queries(args.queries, shape=args.query_shape)
codes, loc_off, lens, total_res = bench.build_db_device(torch, dev, args.db_seqs, 0, args.db_seqs, q_codes, q_off)
db = ctx.database(codes, loc_off, id_base=0, where=capi.S4G_DEVICE)
del codes
pipe = pipeline.DevicePipeline(ctx, db, q_codes, q_off, mat, lens, total_res, max_candidates=args.max_candidates)
for it in range(args.iters):
    torch.cuda.synchronize()
    t0 = time.time()
    if args.stage == "prefilter":
        capi.prefilter(ctx, db, pipe.Q, pipe.k, pipe.N, True, out=(pipe.t_ids, pipe.t_sc, pipe.t_cnt), where=capi.S4G_DEVICE)
    else:
        pipe.step()
    torch.cuda.synchronize()
    print("iter %d: %.3f ms" % (it, (time.time() - t0) * 1e3), file=sys.stderr)
