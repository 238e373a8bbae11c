"""Experiment: two query batches in flight on ONE GPU (two contexts = two streams, one host thread each, s4g_search on each).

The scan (bound by L2 request rate, ALU pipe ~half busy) and the score kernel (ALU-pipe bound, no memory traffic) want
different parts of an SM; with both kernels shaped to half an SM (S4G_PF_WARPS=16 + S4G_PF_BUILD=1024: 32 K registers;
S4G_SW_CTAS=1: 32 K registers) they can be co-resident.  Prints steps/s for: one batch at a time (shipped shape), two in
flight (shipped shape: the persistent grids just queue), two in flight (half-SM shapes), one at a time (half-SM shapes).

    python tools/corun_experiment.py [--queries 1000] [--db-seqs 10000000] [--steps 4]
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--queries", type=int, default=1000)
ap.add_argument("--db-seqs", type=int, default=10_000_000)
ap.add_argument("--max-candidates", type=int, default=5000)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--in-flight", type=int, default=2)
args = ap.parse_args()

import torch  # noqa: E402
from sift4g_b200 import capi  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
mat = np.array(bench.BLOSUM62_A_TO_Z, dtype=np.int32)
q_codes, q_off = bench.make_queries(args.queries)
ctxs = [capi.Context(0) for _ in range(args.in_flight)]
codes, loc_off, lens, total_res = bench.build_db_device(torch, dev, args.db_seqs, 0, args.db_seqs, q_codes, q_off)
db = ctxs[0].database(codes, loc_off, id_base=0, where=capi.S4G_DEVICE)
del codes
Qs = [c.queries(q_codes, q_off) for c in ctxs]
host_threads = max(1, min(16, (os.cpu_count() or 1)))


def step(i, n_threads):
    return capi.search(ctxs[i], db, Qs[i], mat, 5, args.max_candidates, n_threads=n_threads, want_candidates=False, device_results=True)


def run(n_flight, steps_each, n_threads):
    """steps_each searches on each of n_flight contexts, concurrently; returns (ms per search, last result of ctx 0)"""
    out = [None] * n_flight
    err = []

    def work(i):
        try:
            torch.cuda.set_device(0)
            for _ in range(steps_each):
                out[i] = step(i, n_threads)
        except Exception as exc:     # noqa: BLE001
            err.append(exc)
    torch.cuda.synchronize()
    t0 = time.time()
    th = [threading.Thread(target=work, args=(i,)) for i in range(n_flight)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    if err:
        raise err[0]
    ms = (time.time() - t0) * 1e3
    return ms / (steps_each * n_flight), out[0]


def set_env(shape):
    for k in ("S4G_PF_WARPS", "S4G_PF_BUILD", "S4G_SW_CTAS"):
        os.environ.pop(k, None)
    os.environ.update(shape)


shapes = {
    "shipped": {},
    "half_sm": {"S4G_PF_WARPS": "16", "S4G_PF_BUILD": "1024", "S4G_SW_CTAS": "1"},
    "pf16_only": {"S4G_PF_WARPS": "16", "S4G_PF_BUILD": "1024"},
}
# warm every context (buffers, the database's length order) one after the other
set_env({})
for i in range(args.in_flight):
    for _ in range(2):
        r = step(i, host_threads)
ref_hits = (r.n_pairs, len(r.pair_q), r.sw_cells)
for name, n_flight in (("shipped", 1), ("shipped", args.in_flight), ("half_sm", args.in_flight), ("half_sm", 1), ("pf16_only", args.in_flight)):
    set_env(shapes[name])
    run(n_flight, 1, max(1, host_threads // n_flight))
    ms, r = run(n_flight, args.steps, max(1, host_threads // n_flight))
    assert (r.n_pairs, len(r.pair_q), r.sw_cells) == ref_hits, "results changed"
    sm = r.stage_ms
    print(json.dumps({"shape": name, "in_flight": n_flight, "ms_per_search": round(ms, 2), "env": shapes[name],
                      "last_call_stage_ms": {k: round(float(v), 2) for k, v in sm.items()}, "sw_kernel_ms": round(float(r.sw_kernel_ms), 2)}), flush=True)
