"""One GPU: what the prefilter of ONE rank costs in bench.py's weak-scaling family (W x 1000 queries against shard 0 of a
W-way split of the 10 M-sequence database), without the other W - 1 GPUs.  Prints one JSON line per W.

    python tools/shard_emulation.py [--worlds 1,2,4,8]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worlds", default="1,2,4,8")
    ap.add_argument("--db-seqs", type=int, default=10_000_000)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import torch
    from sift4g_b200 import capi
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ctx = capi.Context(0)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    for W in [int(x) for x in args.worlds.split(",")]:
        nq, n_db = 1000 * W, args.db_seqs
        q_codes, q_off = bench.make_queries(nq)
        hi = n_db // W
        codes, loc_off, lens, total_res = bench.build_db_device(torch, dev, n_db, 0, hi, q_codes, q_off)
        db = ctx.database(codes, loc_off, id_base=0, where=capi.S4G_DEVICE)
        del codes
        Q = ctx.queries(q_codes, q_off)
        N = 5000
        out = (torch.zeros((nq, N), dtype=torch.int32, device=dev), torch.zeros((nq, N), dtype=torch.float32, device=dev),
               torch.zeros(nq, dtype=torch.int32, device=dev))
        for _ in range(2):
            capi.prefilter(ctx, db, Q, 5, N, False, out=out, where=capi.S4G_DEVICE)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            capi.prefilter(ctx, db, Q, 5, N, False, out=out, where=capi.S4G_DEVICE)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps({"world": W, "queries": nq, "shard_seqs": hi, "shard_residues": int(lens.sum()), "prefilter_ms": round(ms, 3),
                          "candidates": int(out[2].sum().item())}), flush=True)
        Q.close(); db.close()
        del out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
