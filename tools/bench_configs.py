"""Measurements for the BASELINE.json configurations that are not the bench.py line (one GPU each):

    python tools/bench_configs.py --config c4    # long-query stress: 500 queries of 5 000-35 000 aa, whole hot path,
                                                 # scoring through the intra-sequence striped kernel
    python tools/bench_configs.py --config c5    # prefilter-only sweep, top-N in {1000, 5000, 20000}, on a
                                                 # UniRef90-shaped database (40 M sequences / ~14 B residues)

Prints one JSON line per measurement (device-resident inputs, CUDA events on the library's stream, 1 warm-up pass).
Not the bench: bench.py measures configs[1]; these lines are kept under profiles/.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["c4", "c5"])
    ap.add_argument("--queries", type=int, default=0)
    ap.add_argument("--db-seqs", type=int, default=0)
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()

    import torch
    from sift4g_b200 import capi, pipeline

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ctx = capi.Context(0)
    mat = np.array(bench.BLOSUM62_A_TO_Z, dtype=np.int32)
    hbm = bench.hbm_peak()

    def timed(fn, steps):
        fn()                                    # warm-up (allocations, cut-off free first pass)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, out

    if args.config == "c4":
        nq = args.queries or 500
        n_db = args.db_seqs or 10_000_000
        t0 = time.time()
        q_codes, q_off = bench.make_queries(nq, 5000, 35000, seed=bench.SEED + 2)
        codes, loc_off, lens, total_res = bench.build_db_device(torch, dev, n_db, 0, n_db, q_codes, q_off, seed=bench.SEED + 2)
        torch.cuda.synchronize()
        gen_s = time.time() - t0
        db = ctx.database(codes, loc_off, id_base=0, where=capi.S4G_DEVICE)
        del codes
        pipe = pipeline.DevicePipeline(ctx, db, q_codes, q_off, mat, lens, total_res)
        peak = ctx.dpx_peak(300)
        ms, r = timed(pipe.step, args.steps)
        split = bench.stage_split(torch, ctx, pipe)
        sw_ms = ctx.last_sw_kernel_ms()
        roof = peak * 2 / 6 / 1e9
        print(json.dumps({
            "config": "configs[3]: %d queries of 5000-35000 aa (%.1f M residues) vs %d-sequence / %.2f B-residue synthetic database, whole hot path" % (
                nq, q_off[-1] / 1e6, n_db, total_res / 1e9),
            "ms_per_step": round(ms, 2), "queries_per_sec": round(nq / (ms * 1e-3), 2), "sw_cells_per_step": r.sw_cells,
            "sw_gcups_whole_path": round(r.sw_cells / (ms * 1e-3) / 1e9, 1), "pairs_per_step": r.n_pairs, "kept_hits_per_step": int(len(r.pair_q)),
            "stages_ms": split,
            "roofline": {"bound": "int_dpx", "kernel": "sw_score_striped_stream2_kernel", "kernel_ms": round(sw_ms, 2),
                         "achieved": round(r.sw_cells / (sw_ms * 1e-3) / 1e9, 1), "peak": round(roof, 1), "unit": "GCUPS",
                         "frac": round(r.sw_cells / (sw_ms * 1e-3) / 1e9 / roof, 4)},
            "db_generation_s": round(gen_s, 1)}))
        pipe.close()
        db.close()
        return

    # c5
    nq = args.queries or 1000
    n_db = args.db_seqs or 40_000_000
    t0 = time.time()
    q_codes, q_off = bench.make_queries(nq)
    codes, loc_off, lens, total_res = bench.build_db_device(torch, dev, n_db, 0, n_db, q_codes, q_off)
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    db = ctx.database(codes, loc_off, id_base=0, where=capi.S4G_DEVICE)
    del codes
    Q = ctx.queries(q_codes, q_off)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    n_kmers = int(np.maximum(np.diff(q_off) - 4, 0).sum())
    for N in (1000, 5000, 20000):
        ids = torch.zeros((nq, N), dtype=torch.int32, device=dev)
        sc = torch.zeros((nq, N), dtype=torch.float32, device=dev)
        cnt = torch.zeros(nq, dtype=torch.int32, device=dev)
        ms, _ = timed(lambda: capi.prefilter(ctx, db, Q, 5, N, True, out=(ids, sc, cnt), where=capi.S4G_DEVICE), args.steps)
        alg_bytes = total_res + 8 * n_db + 8 * int(cnt.sum().item())
        print(json.dumps({
            "config": "configs[4]: prefilter only, k=5, top-%d, %d queries (%d indexed k-mers) vs %d-sequence / %.2f B-residue synthetic UniRef90-shaped database" % (
                N, nq, n_kmers, n_db, total_res / 1e9),
            "top_n": N, "ms": round(ms, 2), "residues_per_s": round(total_res / (ms * 1e-3), 0), "sequences_per_s": round(n_db / (ms * 1e-3), 0),
            "roofline": {"bound": "hbm", "achieved": round(alg_bytes / (ms * 1e-3) / 1e9, 2), "peak": hbm, "unit": "GB/s",
                         "frac": round(alg_bytes / (ms * 1e-3) / 1e9 / hbm, 4),
                         "bytes": "1 B/residue + 8 B/sequence + 8 B/retained candidate, read once per query batch"},
            "db_generation_s": round(gen_s, 1)}))
        del ids, sc, cnt
    Q.close()
    db.close()


if __name__ == "__main__":
    main()
