#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU): both multi-GPU forms of the pipeline -- one database shard
per rank with the query-owner cut-off exchange and the overflow-only hit merge, and the NVLink-striped database where
every rank runs its own queries against all stripes -- must give exactly the single-GPU answer: same candidate sets,
same kept hits with bit-identical E-values, same alignment cells and paths.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_multi_gpu.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from sift4g_b200 import capi
    import bench

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = capi.Context(local)
    mat = np.array(bench.BLOSUM62_A_TO_Z, dtype=np.int32)
    # the checks bench.py runs in its warm-up, in both multi-GPU modes: sharded database + NCCL exchange of candidate cut-offs and
    # hits, and NVLink-striped database + split queries
    res = bench.parity_digest(ctx, mat, dist, log=lambda m: print("[exchange] " + m, flush=True), mode="exchange")
    res2 = bench.parity_digest(ctx, mat, dist, log=lambda m: print("[striped] " + m, flush=True), mode="striped")
    ok = res["sharded_equals_single"] and res2["sharded_equals_single"] and (rank != 0 or res["digest"] == res2["digest"])
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()
    if rank == 0:
        print("digest %s (exchange) %s (striped)" % (res["digest"], res2["digest"]))
        print("multi-GPU parity: %s" % ("OK" if int(flag.item()) else "FAILED"))
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
