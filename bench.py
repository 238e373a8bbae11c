#!/usr/bin/env python
"""bench.py -- SIFT4G database-search hot path on B200 (BASELINE.json: "SW GCUPS and queries/sec/box").

A step = one pass of the whole hot path (k-mer prefilter -> Smith-Waterman scores -> E-value selection ->
traceback of kept hits) over one batch of synthetic queries against a synthetic, HBM-resident database of
BASELINE.json configs[1]'s shape: 1,000 queries (len 100-1,000) vs 10 M sequences / ~3.5 B residues.
At N GPUs the database is sharded (one resident shard per rank) and the query batch grows with N (1,000 x N queries per
step): every GPU then scans 1/N of the database for N times the queries and scores/aligns 1/N of every candidate list,
i.e. per-GPU work is fixed -- "scaling": "weak" (towards configs[2]'s 20,000-query batch).  --scaling strong keeps the
1,000-query batch and only shards the database.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

value  = GCUPS = algorithmic SW cells (sum_q len(q) * sum_{t in cand(q)} len(t)) / step time, device timed
         (CUDA events on the launching stream, max over ranks), inputs resident in HBM.
e2e    = the same metric through the host-buffer API (queries uploaded, candidates / scores / alignments read
         back every step).
--impl reference times the UNMODIFIED reference CPU build (oracle/_ref) on a bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20240002           # 20240001 + config id


# ------------------------------------------------------------------------------------------------------------
# synthetic workload (same recipe on host and device; see sift4g_b200/synth.py)

def make_queries(n, lo=100, hi=1000, seed=SEED, shape="uniform"):
    """uniform: lengths U[lo, hi] (configs[1]).  human: human-proteome-shaped log-normal lengths (median ~415, mean ~560,
    clipped to [50, 35000]) for the configs[2]-shaped batch -- queries beyond 1024 aa take the
    striped kernel."""
    from sift4g_b200 import synth
    rng = np.random.default_rng(seed)
    if shape == "human":
        lens = np.clip(np.exp(rng.normal(6.03, 0.775, size=n)), 50, 35000).astype(np.int64)
        qs = [synth.random_codes(rng, int(l), 0.001) for l in lens]
    else:
        qs = [synth.random_codes(rng, rng.integers(lo, hi + 1), 0.001) for _ in range(n)]
    return synth.pack(qs)


def db_lengths(n_db, seed=SEED):
    rng = np.random.default_rng(seed + 1)
    return np.clip(np.exp(rng.normal(5.6, 0.6, size=n_db)), 30, 35000).astype(np.int64)


def plant_plan(n_db, lens, q_off, seed=SEED, homologs=(50, 400), scale=1.0):
    """Which database sequences carry a mutated copy of which query window (host, global, seeded)."""
    rng = np.random.default_rng(seed + 2)
    nq = len(q_off) - 1
    qlen = np.diff(q_off)
    per_q = np.maximum(1, (rng.integers(homologs[0], homologs[1] + 1, size=nq) * scale).astype(np.int64))
    per_q = np.minimum(per_q, max(1, n_db // (2 * nq)))
    P = int(per_q.sum())
    qid = np.repeat(np.arange(nq), per_q)
    seq = rng.choice(n_db, size=P, replace=False)
    L = lens[seq]
    ql = qlen[qid]
    partial = rng.random(P) < 0.5
    a = np.where(partial, (rng.random(P) * ql * 0.5).astype(np.int64), 0)
    span = np.where(partial, np.maximum(30, ((ql - a) * rng.uniform(0.3, 1.0, size=P)).astype(np.int64)), ql - a)
    w = np.minimum(span, L)
    s = ((L - w) * rng.random(P)).astype(np.int64)
    ident = rng.uniform(0.3, 0.95, size=P)
    ev_pos = (rng.random((P, 3)) * np.maximum(w, 1)[:, None]).astype(np.int64)
    ev_delta = rng.integers(-5, 6, size=(P, 3))
    return dict(qid=qid, seq=seq, a=a, w=w, s=s, ident=ident, ev_pos=ev_pos, ev_delta=ev_delta)


def build_db_device(torch, dev, n_db, lo, hi, q_codes, q_off, seed=SEED, plant_scale=1.0):
    """Database shard [lo, hi) generated on the device; identical content whatever the shard split."""
    from sift4g_b200 import synth
    lens = db_lengths(n_db, seed)
    off = np.zeros(n_db + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    r0, r1 = int(off[lo]), int(off[hi])
    codes = torch.empty(r1 - r0 + 64, dtype=torch.uint8, device=dev)
    cdf = torch.tensor(np.cumsum(synth.letter_table(0.001)), dtype=torch.float32, device=dev)
    chunk = 1 << 26
    g = torch.Generator(device=dev)
    for c in range(r0 // chunk, (r1 + chunk - 1) // chunk if r1 > r0 else 0):
        g.manual_seed(seed * 1000003 + c)
        u = torch.rand(chunk, generator=g, device=dev)
        v = torch.searchsorted(cdf, u).clamp_(max=25).to(torch.uint8)
        a, b = max(c * chunk, r0), min((c + 1) * chunk, r1)
        codes[a - r0:b - r0] = v[a - c * chunk:b - c * chunk]
        del u, v
    # planted homologs
    plan = plant_plan(n_db, lens, q_off, seed, scale=plant_scale)
    m = (plan["seq"] >= lo) & (plan["seq"] < hi)
    if m.any():
        t = lambda x, dt=torch.int64: torch.from_numpy(np.ascontiguousarray(x[m])).to(dev).to(dt)
        w = t(plan["w"])
        P = int(w.numel())
        pid = torch.repeat_interleave(torch.arange(P, device=dev), w)
        start = torch.cumsum(w, 0) - w
        x = torch.arange(int(w.sum()), device=dev) - start[pid]
        shift = torch.zeros_like(x)
        ev_pos, ev_delta = t(plan["ev_pos"]), t(plan["ev_delta"])
        for e in range(3):
            shift += ev_delta[pid, e] * (x >= ev_pos[pid, e])
        src = t(plan["a"])[pid] + x + shift
        qid = t(plan["qid"])[pid]
        qo = torch.from_numpy(q_off).to(dev)
        qlen = (qo[1:] - qo[:-1])[qid]
        g.manual_seed(seed * 7919 + lo)
        keep = (torch.rand(x.numel(), generator=g, device=dev) < t(plan["ident"], torch.float32)[pid]) & (src >= 0) & (src < qlen)
        dst = torch.from_numpy(off).to(dev)[t(plan["seq"])[pid]] + t(plan["s"])[pid] + x - r0
        qc = torch.from_numpy(q_codes).to(dev)
        codes[dst[keep]] = qc[(qo[qid] + src)[keep]]
    loc_off = torch.from_numpy(off[lo:hi + 1] - r0).to(dev)
    return codes, loc_off, lens[lo:hi], int(off[-1])


def build_db_host(n_db, q_codes, q_off, seed=SEED, plant_scale=1.0):
    """Host twin of build_db_device for the bounded CPU-baseline sample (numpy)."""
    from sift4g_b200 import synth
    lens = db_lengths(n_db, seed)
    off = np.zeros(n_db + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    rng = np.random.default_rng(seed + 3)
    codes = rng.choice(26, size=int(off[-1]), p=synth.letter_table(0.001)).astype(np.uint8)
    plan = plant_plan(n_db, lens, q_off, seed, scale=plant_scale)
    qlen = np.diff(q_off)
    for p in range(len(plan["seq"])):
        w = int(plan["w"][p])
        x = np.arange(w)
        shift = np.zeros(w, dtype=np.int64)
        for e in range(3):
            shift += plan["ev_delta"][p, e] * (x >= plan["ev_pos"][p, e])
        src = plan["a"][p] + x + shift
        q = plan["qid"][p]
        keep = (rng.random(w) < plan["ident"][p]) & (src >= 0) & (src < qlen[q])
        dst = off[plan["seq"][p]] + plan["s"][p] + x
        codes[dst[keep]] = q_codes[q_off[q] + src[keep]]
    return codes, off


def write_fasta(path, codes, off, prefix):
    txt = (codes + 65).astype(np.uint8).tobytes()
    with open(path, "wb") as f:
        for i in range(len(off) - 1):
            f.write(b">%s%08d\n" % (prefix, i))
            f.write(txt[off[i]:off[i + 1]])
            f.write(b"\n")


# ------------------------------------------------------------------------------------------------------------
# clocks

class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML is queried in-process (pynvml, ~50 us per
    sample, no fork): spawning nvidia-smi from a process with a large address space stalled the timed steps by tens of
    milliseconds.  nvidia-smi is only the fallback when pynvml is unusable."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, period=0.05):
        self.index, self.samples, self.stop, self.th, self.period = index, [], False, None, period
        self.nvml = self.handle = None
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                ent = vis.split(",")[index].strip()
                idx = int(ent) if ent.isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        mhz = int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        r = int(get(self.handle))
        bits = [0x8, 0x40, 0x20, 0x4]      # HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap (nvml.h)
        return [str(mhz), str(self.max_mhz)] + ["Active" if r & b else "Not Active" for b in bits]

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        return [x.strip() for x in out.split(",")] if out else None

    def _run(self):
        while not self.stop:
            try:
                s = self._sample_nvml() if self.nvml else self._sample_smi()
                if s:
                    self.samples.append(s)
            except Exception:
                pass
            time.sleep(self.period if self.nvml else 0.5)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML and nvidia-smi unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        reasons = [n for i, n in enumerate(self.NAMES) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples), "source": "nvml" if self.nvml else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline (rank 0 only): the unmodified reference on a bounded sample

def reference_sample(tmp, n_queries, n_db, seed=SEED):
    q_codes, q_off = make_queries(n_queries, seed=seed)
    codes, off = build_db_host(n_db, q_codes, q_off, seed, plant_scale=0.25)
    write_fasta(tmp + "/q.fa", q_codes, q_off, b"Q")
    write_fasta(tmp + "/d.fa", codes, off, b"D")
    return tmp + "/q.fa", tmp + "/d.fa"


def run_reference_once(qf, df, threads):
    from oracle import oracle as O
    out = subprocess.run([O.REF_DUMP, "pipeline", qf, df, "5", "5000", str(threads), "0.0001", "400", "quiet"], capture_output=True, text=True, check=True).stdout
    for l in out.split("\n"):
        if l.startswith("timing"):
            w = l.split()
            d = {w[i]: w[i + 1] for i in range(1, len(w) - 1, 2)}
            return float(d["search_s"]), float(d["align_s"]), int(d["sw_cells"]), int(d["pairs"])
    raise RuntimeError("reference produced no timing line:\n" + out[-500:])


def reference_arm(args, out=sys.stdout):
    from oracle import oracle as O
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not O.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref binaries missing (built by oracle/Makefile in the build container)"}), file=out, flush=True)
        return
    cores = os.cpu_count() or 1
    tmp = tempfile.mkdtemp()
    qf, df = reference_sample(tmp, args.ref_queries, args.ref_db_seqs)
    times, cells = [], 0
    for i in range(args.warmup + args.steps):
        s, a, cells, pairs = run_reference_once(qf, df, cores)
        if i >= args.warmup:
            times.append((s, a))
    tot = sum(s + a for s, a in times)
    gcups = cells * len(times) / tot / 1e9
    sample = "%d queries (len 100-1000) x %d-sequence database (same generator as the GPU arm), max_candidates 5000, %d threads; whole reference path (searchDatabase + alignDatabase) per step" % (args.ref_queries, args.ref_db_seqs, cores)
    line = {"impl": "reference", "metric": "sw_gcups", "value": round(gcups, 4), "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(tot / len(times) * 1e3, 3), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "int8/int16/int32 SIMD (swimd AVX2)",
            "data": "synthetic", "config": {"workload": "configs[1] bounded sample: " + sample},
            "queries_per_sec": round(args.ref_queries * len(times) / tot, 4),
            "stages_s": {"search": round(sum(s for s, _ in times) / len(times), 4), "align": round(sum(a for _, a in times) / len(times), 4)},
            "cpu_baseline": {"value": round(gcups, 4), "unit": "GCUPS", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": round(gcups, 4), "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), file=out, flush=True)


def cpu_baseline(args):
    """Bounded reference run beside the GPU numbers (rank 0, N=1)."""
    from oracle import oracle as O
    if not O.have_ref():
        return {"value": None, "unit": "GCUPS", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
    cores = os.cpu_count() or 1
    tmp = tempfile.mkdtemp()
    qf, df = reference_sample(tmp, args.ref_queries, args.ref_db_seqs)
    s, a, cells, pairs = run_reference_once(qf, df, cores)
    return {"value": round(cells / (s + a) / 1e9, 4), "unit": "GCUPS", "cores": cores, "kind": "reference",
            "sample": "%d queries x %d-sequence database sample of the same generator, %d (query,candidate) pairs, %.3e SW cells; reference searchDatabase %.2f s + alignDatabase %.2f s" % (
                args.ref_queries, args.ref_db_seqs, pairs, cells, s, a),
            "sw_stage_gcups": round(cells / a / 1e9, 4), "queries_per_sec": round(args.ref_queries / (s + a), 4)}


# ------------------------------------------------------------------------------------------------------------

def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner under NCCL_DEBUG), so the
    real stdout is kept aside for the JSON line and file descriptor 1 is pointed at stderr for everybody else."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--queries", type=int, default=1000, help="queries per step and GPU (weak) / per step (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--query-shape", default="uniform", choices=["uniform", "human"], help="query length distribution (human: configs[2])")
    ap.add_argument("--db-seqs", type=int, default=10_000_000)
    ap.add_argument("--max-candidates", type=int, default=5000)
    ap.add_argument("--ref-queries", type=int, default=16)
    ap.add_argument("--ref-db-seqs", type=int, default=100_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "ours" else max(args.warmup, 1)

    if args.impl == "reference":
        reference_arm(args, out)
        return

    import torch
    import torch.distributed as dist
    from sift4g_b200 import capi, pipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    use_dist = world > 1
    if use_dist:
        dist.init_process_group("nccl", device_id=dev)
    ctx = capi.Context(local)
    mat = np.array(BLOSUM62_A_TO_Z, dtype=np.int32)

    n_queries = args.queries * world if args.scaling == "weak" else args.queries
    q_codes, q_off = make_queries(n_queries, shape=args.query_shape)
    n_db = args.db_seqs
    lo, hi = n_db * rank // world, n_db * (rank + 1) // world
    t0 = time.time()
    codes, loc_off, lens, total_res = build_db_device(torch, dev, n_db, lo, hi, q_codes, q_off)
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    db = ctx.database(codes, loc_off, id_base=lo, where=capi.S4G_DEVICE)
    del codes
    pipe = pipeline.DevicePipeline(ctx, db, q_codes, q_off, mat, lens, total_res, max_candidates=args.max_candidates, dist=dist if use_dist else None)

    peak = ctx.dpx_peak(300)                  # sustained VIADDMNMX.S16x2 lane-ops/s on this device
    roof_gcups = peak * 2 / 6 / 1e9

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        r = pipe.step()
    barrier()
    ctx.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage = {"prefilter": 0.0, "score": 0.0, "select_host": 0.0, "align": 0.0}
    sw_ms = []
    cells_local = 0
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for _ in range(args.steps):
            r = pipe.step()
            cells_local = r.sw_cells
            sw_ms.append(r.sw_kernel_ms)
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count()
    sw_kernel_ms = sw_ms[-1]                 # SW score kernel launches of the last step (one per half of the query batch), CUDA events inside the library
    t = torch.tensor([ms, float(cells_local), float(sw_kernel_ms), float(r.n_pairs), float(len(r.pair_q))], dtype=torch.float64, device=dev)
    if use_dist:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, cells, sw_kernel_ms_max = float(tmax[0]), float(tsum[1]), float(tmax[2])
        pairs, hits = float(tsum[3]), float(tsum[4])
    else:
        cells, sw_kernel_ms_max, pairs, hits = float(cells_local), float(sw_kernel_ms), float(r.n_pairs), float(len(r.pair_q))
    ms_per_step = ms / args.steps
    gcups = cells / (ms_per_step * 1e-3) / 1e9

    # stage split (one extra, untimed step with synchronisation between stages; rank-local)
    split = stage_split(torch, ctx, pipe)

    # e2e through the host-buffer API
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(torch, ctx, db, pipe, q_codes, q_off, mat, lens, total_res, args, use_dist, dist, dev, n_queries)

    base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base = cpu_baseline(args)

    if rank == 0:
        kern_gcups = cells_local / (sw_kernel_ms * 1e-3) / 1e9 if sw_kernel_ms > 0 else None
        line = {
            "metric": "sw_gcups", "value": round(gcups, 2), "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "s16x2 (DPX), s32 re-run on overflow",
            "data": "synthetic",
            "config": {"workload": "%s: %d queries (%s) vs %d-sequence / %.2f B-residue synthetic database, whole hot path per step (prefilter k=5, top %d; SW BLOSUM62 10/1; E<=1e-4, top 400; traceback)" % (
                "configs[1]" if n_db == 10_000_000 and args.query_shape == "uniform" else ("configs[2]-shaped" if n_db >= 40_000_000 else "custom"),
                n_queries, "len 100-1000" if args.query_shape == "uniform" else "human-proteome-shaped log-normal lengths, median %d, max %d" % (
                    int(np.median(np.diff(q_off))), int(np.diff(q_off).max())), n_db, total_res / 1e9, args.max_candidates),
                "sharding": "database split in %d contiguous shards, one resident per GPU; %d queries per step (%s scaling: %s)" % (
                    world, n_queries, args.scaling, "1000 queries per GPU and step" if args.scaling == "weak" else "same batch at every N"),
                "l2": "inputs (%.2f GB database shard per GPU) exceed the 126 MB L2; no explicit flush" % ((hi - lo) / n_db * total_res / 1e9),
                "db_generation_s": round(gen_s, 2)},
            "queries_per_sec": round(n_queries / (ms_per_step * 1e-3), 2),
            "sw_cells_per_step": cells, "pairs_per_step": pairs, "kept_hits_per_step": hits,
            "stages_ms": split,
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "roofline": {"bound": "int_dpx", "kernel": "sw_score_packed_kernel", "achieved": round(kern_gcups, 2) if kern_gcups else None, "peak": round(roof_gcups, 2),
                         "unit": "GCUPS", "frac": round(kern_gcups / roof_gcups, 4) if kern_gcups else None,
                         "traffic": NCU_TRAFFIC_C2 if (world == 1 and n_queries == 1000 and n_db == 10_000_000 and args.max_candidates == 5000) else None,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full capture of this command (profiles/r01i_sw_digest.md; re-captured as two half-batch launches, 0.728 GB each, in profiles/r01t_sw_digest.md)",
                         "kernel_ms": round(sw_kernel_ms, 3),
                         "peak_source": "measured live on this GPU (no DPX figure in MEASURED_PEAKS.json): %.4e VIADDMNMX.S16x2 lane-ops/s sustained over 300 ms x 2 cells per op / 6 instructions per cell (BASELINE.md)" % peak},
            "roofline_prefilter": {"bound": "hbm", "achieved": round(((hi - lo) / n_db * total_res + 8 * (hi - lo)) / (split["prefilter"] * 1e-3) / 1e9, 2) if split["prefilter"] else None,
                                   "peak": hbm_peak(), "unit": "GB/s", "note": "database bytes (1 B/residue + 8 B/sequence) / prefilter stage time"},
        }
        if line["roofline_prefilter"]["achieved"]:
            line["roofline_prefilter"]["frac"] = round(line["roofline_prefilter"]["achieved"] / line["roofline_prefilter"]["peak"], 4)
        if e2e is not None:
            line["e2e"] = e2e
        if base is not None:
            line["cpu_baseline"] = base
        print(json.dumps(line), file=out, flush=True)
    pipe.close()
    db.close()
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()


# DRAM bytes of one sw_score_packed_kernel launch at the default C2 workload (1.404 GB read + 28.3 MB written), from the
# ncu --set full capture summarised in profiles/r01i_sw_digest.md.  The kernel is integer-issue bound; its algorithmic
# DRAM traffic is the candidates' residues (5 M targets, ~107 residues each, fetched in 32-byte sectors).
NCU_TRAFFIC_C2 = 1404079000 + 28333056


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0   # fallback stated in B200_PROFILING.md


def stage_split(torch, ctx, pipe, reps=3):
    """Per-stage device+host time of a step (stream synchronised after every stage, median of `reps` untimed steps;
    explains `value`).  score_kernel is the dominant kernel alone (CUDA events inside the library);
    score_select_other = candidate gathering, sort of the pairs, re-runs, E-value screen, D2H of the survivors,
    exact host selection and (N > 1) the hit merge."""
    runs = []
    for _ in range(reps):
        st = {}
        pipe.step(stages=st)
        runs.append(st)
    med = lambda k: sorted(r.get(k, 0.0) for r in runs)[len(runs) // 2]
    pre = med("prefilter") + med("exchange_rows") + med("cutoff")
    kern = med("sw_kernel")
    other = med("own_candidates") + med("sw_score") - kern + med("screen_d2h") + med("select_hits") + med("merge_hits")
    out = {"prefilter": round(pre, 3), "score_kernel": round(kern, 3), "score_select_other": round(other, 3), "align": round(med("align"), 3)}
    if "exchange_rows" in runs[0]:
        out["prefilter_scan_only"] = round(med("prefilter"), 3)
        out["exchange"] = round(med("exchange_rows") + med("cutoff") + med("merge_hits"), 3)
    return out


def run_e2e(torch, ctx, db, pipe, q_codes, q_off, mat, lens, total_res, args, use_dist, dist, dev, n_queries):
    """Same step through the public pipeline API with HOST buffers: the query batch is uploaded every step and the
    candidate lists, survivor scores and alignments are copied back to the host inside the timed region."""
    pipe.step(e2e=True)
    if use_dist:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    cells = h2d = d2h = 0
    n = max(1, min(args.steps, 3))
    for _ in range(n):
        r = pipe.step(e2e=True)
        cells, h2d, d2h = r.sw_cells, r.h2d_bytes, r.d2h_bytes
    torch.cuda.synchronize()
    dt = (time.time() - t0) / n
    t = torch.tensor([dt, float(cells), float(h2d), float(d2h)], dtype=torch.float64, device=dev)
    if use_dist:
        tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        dt, cells, h2d, d2h = float(tm[0]), float(ts[1]), float(ts[2]), float(ts[3])
    return {"value": round(cells / dt / 1e9, 2), "unit": "GCUPS", "ms_per_step": round(dt * 1e3, 3), "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "queries_per_sec": round(n_queries / dt, 2),
            "timed": "host wall clock around pipeline.DevicePipeline.step(e2e=True): queries H2D, candidate lists + survivor scores + alignments D2H every step; max over ranks"}


# BLOSUM62 over 'A'..'Z' exactly as the reference's scorer hands it to the GPU seam (sw/constants.c:87-114);
# tests/test_oracle_golden.py pins the same 676 numbers against the reference's scorerCreateMatrix().
def _blosum():
    order = "ARNDCQEGHILKMFPSTWYVBZX"
    rows = """4 -1 -2 -2 0 -1 -1 0 -2 -1 -1 -1 -1 -2 -1 1 0 -3 -2 0 -2 -1 0
-1 5 0 -2 -3 1 0 -2 0 -3 -2 2 -1 -3 -2 -1 -1 -3 -2 -3 -1 0 -1
-2 0 6 1 -3 0 0 0 1 -3 -3 0 -2 -3 -2 1 0 -4 -2 -3 3 0 -1
-2 -2 1 6 -3 0 2 -1 -1 -3 -4 -1 -3 -3 -1 0 -1 -4 -3 -3 4 1 -1
0 -3 -3 -3 9 -3 -4 -3 -3 -1 -1 -3 -1 -2 -3 -1 -1 -2 -2 -1 -3 -3 -2
-1 1 0 0 -3 5 2 -2 0 -3 -2 1 0 -3 -1 0 -1 -2 -1 -2 0 3 -1
-1 0 0 2 -4 2 5 -2 0 -3 -3 1 -2 -3 -1 0 -1 -3 -2 -2 1 4 -1
0 -2 0 -1 -3 -2 -2 6 -2 -4 -4 -2 -3 -3 -2 0 -2 -2 -3 -3 -1 -2 -1
-2 0 1 -1 -3 0 0 -2 8 -3 -3 -1 -2 -1 -2 -1 -2 -2 2 -3 0 0 -1
-1 -3 -3 -3 -1 -3 -3 -4 -3 4 2 -3 1 0 -3 -2 -1 -3 -1 3 -3 -3 -1
-1 -2 -3 -4 -1 -2 -3 -4 -3 2 4 -2 2 0 -3 -2 -1 -2 -1 1 -4 -3 -1
-1 2 0 -1 -3 1 1 -2 -1 -3 -2 5 -1 -3 -1 0 -1 -3 -2 -2 0 1 -1
-1 -1 -2 -3 -1 0 -2 -3 -2 1 2 -1 5 0 -2 -1 -1 -1 -1 1 -3 -1 -1
-2 -3 -3 -3 -2 -3 -3 -3 -1 0 0 -3 0 6 -4 -2 -2 1 3 -1 -3 -3 -1
-1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4 7 -1 -1 -4 -3 -2 -2 -1 -2
1 -1 1 0 -1 0 0 0 -1 -2 -2 0 -1 -2 -1 4 1 -3 -2 -2 0 0 0
0 -1 0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1 1 5 -2 -2 0 -1 -1 0
-3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1 1 -4 -3 -2 11 2 -3 -4 -3 -2
-2 -2 -2 -3 -2 -1 -2 -3 2 -1 -1 -2 -1 3 -3 -2 -2 2 7 -1 -3 -2 -1
0 -3 -3 -3 -1 -2 -2 -3 -3 3 1 -2 1 -1 -2 -2 0 -3 -1 4 -3 -2 -1
-2 -1 3 4 -3 0 1 -1 0 -3 -4 0 -3 -3 -2 0 -1 -4 -3 -3 4 1 -1
-1 0 0 1 -3 3 4 -2 0 -3 -3 1 -1 -3 -1 0 -1 -3 -2 -2 1 4 -1
0 -1 -1 -1 -2 -1 -1 -1 -1 -1 -1 -1 -1 -1 -2 0 0 -2 -1 -1 -1 -1 -1""".split("\n")
    t = [[int(x) for x in r.split()] for r in rows]
    pos = {c: i for i, c in enumerate(order)}
    out = []
    for a in range(26):
        for b in range(26):
            ca, cb = chr(65 + a), chr(65 + b)
            if ca in pos and cb in pos:
                out.append(t[pos[ca]][pos[cb]])
            elif ca not in pos and cb not in pos:
                out.append(1)
            else:
                out.append(-4)
    return out


BLOSUM62_A_TO_Z = _blosum()

if __name__ == "__main__":
    main()
