#!/usr/bin/env python
"""bench.py -- SIFT4G database-search hot path on B200 (BASELINE.json: "SW GCUPS and queries/sec/box").

A step = one pass of the whole hot path (k-mer prefilter -> Smith-Waterman scores -> E-value selection ->
traceback of kept hits) over one batch of synthetic queries against a synthetic, HBM-resident database of
BASELINE.json configs[1]'s shape: 1,000 queries (len 100-1,000) vs 10 M sequences / ~3.5 B residues.
At N GPUs (weak scaling, the default: 1,000 queries per GPU and step) the database is STRIPED over the GPUs' HBM -- one
resident stripe each, all stripes mapped into every GPU's address space (CUDA VMM, peers read over NVLink) -- and every rank
runs the whole path for its own 1,000 queries against all of it: per-GPU work is exactly the one-GPU step and nothing is
merged.  --multi-gpu exchange selects the other form (one database shard per rank scanned for all queries, candidate cut-offs
and hits exchanged over NCCL), which is the default for --scaling strong (the same 1,000-query batch at every N).

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
    # letters through a 65 536-entry quantile table of the background distribution (rng.choice takes ~15 s for 3e8 residues)
    table = np.searchsorted(np.cumsum(synth.letter_table(0.001)), (np.arange(65536) + 0.5) / 65536.0).clip(max=25).astype(np.uint8)
    codes = table[rng.integers(0, 65536, size=int(off[-1]), dtype=np.uint16)]
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
    """>P%08d records, one line per sequence; assembled with numpy in blocks of sequences (a Python loop over a million
    records would take longer than the reference run the file is for)."""
    n = len(off) - 1
    with open(path, "wb") as f:
        for a in range(0, n, 200_000):
            b = min(n, a + 200_000)
            m = b - a
            lens = np.diff(off[a:b + 1])
            hdr = np.frombuffer(b"".join(b">%s%08d\n" % (prefix, i) for i in range(a, b)), dtype=np.uint8).reshape(m, -1)
            h = hdr.shape[1]
            start = (off[a:b] - off[a]) + (h + 1) * np.arange(m)                  # first byte of every record in this block
            out = np.empty(int(off[b] - off[a]) + (h + 1) * m, dtype=np.uint8)
            out[(start[:, None] + np.arange(h)[None, :]).reshape(-1)] = hdr.reshape(-1)
            out[start + h + lens] = 10
            res = np.ones(len(out), dtype=bool)
            res[(start[:, None] + np.arange(h)[None, :]).reshape(-1)] = False
            res[start + h + lens] = False
            out[res] = codes[off[a]:off[b]] + 65
            f.write(out.tobytes())


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
    """`--impl reference`: the UNMODIFIED reference (oracle/_ref: swimd AVX2 scoring, SSW traceback, its own prefilter) on all
    host cores, quoted on our arm's config; every step is one run of its whole database-search path (searchDatabase +
    alignDatabase through the seam harness, FASTA parses included as in the reference) over a bounded sample of the workload."""
    from oracle import oracle as O
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    if not O.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref binaries missing (built by oracle/Makefile in the build container)"}), file=out, flush=True)
        return
    cores = os.cpu_count() or 1
    tmp = tempfile.mkdtemp()
    qf, df = reference_sample(tmp, args.ref_queries, args.ref_db_seqs)
    times, cells, pairs = [], 0, 0
    warm = min(args.warmup, 1)                      # a warm-up run only warms the page cache
    for i in range(warm + args.steps):
        s, a, cells, pairs = run_reference_once(qf, df, cores)
        if i >= warm:
            times.append((s, a))
    tot = sum(s + a for s, a in times)
    gcups = cells * len(times) / tot / 1e9
    sample = "%d queries (len 100-1000) x %d-sequence database of the same generator (%d pairs, %.3e SW cells per step), max_candidates 5000, %d threads; whole reference path (searchDatabase + alignDatabase, both FASTA parses) per step" % (
        args.ref_queries, args.ref_db_seqs, pairs, cells, cores)
    line = {"impl": "reference", "metric": "sw_gcups", "value": round(gcups, 4), "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(tot / len(times) * 1e3, 3), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "int8/int16/int32 SIMD (swimd AVX2)",
            "data": "synthetic", "config": workload_config(args, world),
            "queries_per_sec": round(args.ref_queries * len(times) / tot, 4),
            "stages_s": {"search": round(sum(s for s, _ in times) / len(times), 4), "align": round(sum(a for _, a in times) / len(times), 4)},
            "cpu_baseline": {"value": round(gcups, 4), "unit": "GCUPS", "cores": cores, "kind": "reference", "sample": sample,
                             "sw_stage_gcups": round(cells * len(times) / sum(a for _, a in times) / 1e9, 4)},
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
# result digest of a FIXED verification workload, computed in the warm-up of every run (any N): the database is sharded
# over the ranks like the bench workload; rank 0 also runs it unsharded.  Equal digests across the N = 1, 2, 4, 8 lines of a
# scaling run mean the sharded path returns the single-GPU results bit for bit (candidate sets, kept hits with their
# E-values as hex doubles, alignment cells, path bytes).

PARITY_CASES = ((51, 13, 6000, 200, 400), (52, 9, 4000, 300, 6), (53, 5, 900, 2000, 400), (54, 48, 40000, 500, 400))


def _collect(r, nq):
    """host view of one rank's step result: per query candidate ids, hits as tuples incl. cells and path bytes"""
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


def parity_digest(ctx, mat, dist=None, log=None, mode="exchange"):
    """-> {"digest": sha256 of the order-normalised results of PARITY_CASES through the (sharded) device pipeline,
           "sharded_equals_single": bool (rank 0 re-runs every case on one unsharded database), "cases": n}"""
    import hashlib
    from sift4g_b200 import pipeline, synth
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    sha = hashlib.sha256()
    same = True
    for seed, nq, n_db, N, M in PARITY_CASES:
        queries, db = synth.make_dataset(seed, nq, n_db, q_len=(50, 400), homologs=(8, 25), rare_fraction=0.005)
        qc, qo = synth.pack(queries); dc, do = synth.pack(db)
        lens = np.diff(do)
        lo, hi = n_db * rank // world, n_db * (rank + 1) // world
        if mode == "striped" and dist is not None:
            # NVLink-striped database: this rank contributes the residues of [lo, hi), sees the whole database and runs ITS
            # slice of the queries against all of it -- nothing is merged, the ranks' results are just put side by side
            import torch
            from sift4g_b200 import stripes
            S = stripes.StripedDatabase(ctx, torch.from_numpy(dc[do[lo]:do[hi]].copy()).cuda(), do, lo, hi, dist=dist)
            qa, qb = nq * rank // world, nq * (rank + 1) // world
            mine = ([], [])
            if qb > qa:
                pipe = pipeline.DevicePipeline(ctx, S.db, qc[qo[qa]:qo[qb]], qo[qa:qb + 1] - qo[qa], mat, lens, int(do[-1]), max_candidates=N, max_alignments=M)
                c, h = _collect(pipe.step(), qb - qa)
                mine = (c, [(x[0] + qa,) + x[1:] for x in h])
                pipe.close()
            torch.cuda.synchronize()
            S.close()
            parts = [None] * world
            dist.all_gather_object(parts, mine)
            if rank == 0:
                cands = [np.sort(c) for p in parts for c in p[0]]
                hits = sorted(h for p in parts for h in p[1])
        else:
            D = ctx.database(dc[do[lo]:do[hi]], do[lo:hi + 1] - do[lo], id_base=lo)
            pipe = pipeline.DevicePipeline(ctx, D, qc, qo, mat, lens[lo:hi], int(do[-1]), max_candidates=N, max_alignments=M, dist=dist)
            mine = _collect(pipe.step(), nq)
            pipe.close(); D.close()
            parts = [mine]
            if dist is not None:
                parts = [None] * world
                dist.all_gather_object(parts, mine)
            if rank == 0:
                cands = [np.sort(np.concatenate([p[0][q] for p in parts])) for q in range(nq)]
                hits = sorted(h for p in parts for h in p[1])
        if rank == 0:
            for q in range(nq):
                sha.update(cands[q].astype("<u4").tobytes())
            for h in hits:
                sha.update(repr(h[:5]).encode()); sha.update(h[5])
            if world > 1:
                D1 = ctx.database(dc, do)
                p1 = pipeline.DevicePipeline(ctx, D1, qc, qo, mat, lens, int(do[-1]), max_candidates=N, max_alignments=M)
                c1, h1 = _collect(p1.step(), nq)
                p1.close(); D1.close()
                ok = all(np.array_equal(cands[q], np.sort(c1[q])) for q in range(nq)) and hits == sorted(h1)
                same = same and ok
                if log:
                    log("parity case seed %d: %d queries, %d candidates, %d hits over %d shards: %s" % (
                        seed, nq, sum(len(c) for c in c1), len(h1), world, "equal to one GPU" if ok else "DIFFERENT from one GPU"))
    return {"digest": sha.hexdigest() if rank == 0 else None, "sharded_equals_single": bool(same), "cases": len(PARITY_CASES),
            "what": "sha256 over candidate id sets, kept hits (query, E as hex double, score, target, cells) and path bytes of %d fixed small workloads run through the %d-GPU pipeline (%s) in the warm-up; the same at every N" % (
                len(PARITY_CASES), world, "database striped over the GPUs' HBM, queries split" if (mode == "striped" and world > 1) else "database sharded, candidate/hit exchange" if world > 1 else "one resident database")}


def workload_config(args, world):
    """`config` of the JSON line: the workload both arms are quoted on (the reference arm times bounded samples of it)."""
    n_queries = args.queries * world if args.scaling == "weak" else args.queries
    n_db = args.db_seqs
    total_res = int(db_lengths(n_db).sum())
    qdesc = "len 100-1000"
    if args.query_shape == "human":
        ql = np.diff(make_queries(n_queries, shape="human")[1])
        qdesc = "human-proteome-shaped log-normal lengths, median %d, max %d" % (int(np.median(ql)), int(ql.max()))
    name = "configs[1]" if n_db == 10_000_000 and args.query_shape == "uniform" else ("configs[2]-shaped" if n_db >= 40_000_000 else "custom")
    return {"workload": "%s: %d queries (%s) vs %d-sequence / %.2f B-residue synthetic database, whole hot path per step (prefilter k=5, top %d; SW BLOSUM62 10/1; E<=1e-4, top 400; traceback)" % (
                name, n_queries, qdesc, n_db, total_res / 1e9, args.max_candidates),
            "sharding": ("database striped over the HBM of %d GPUs (one resident stripe each, all stripes mapped into every GPU's address space, peers read over NVLink); "
                         "every GPU runs the whole path for its own slice of the queries, no merge; %d queries per step (%s scaling: %s)" if multi_gpu_mode(args, world) == "striped" else
                         "database split in %d contiguous shards, one resident per GPU, candidate lists and hits exchanged over NCCL; %d queries per step (%s scaling: %s)") % (
                world, n_queries, args.scaling, "%d queries per GPU and step" % args.queries if args.scaling == "weak" else "same batch at every N"),
            "l2": "inputs (%.2f GB database %s per GPU) exceed the 126 MB L2; no explicit flush" % (total_res / world / 1e9, "stripe" if multi_gpu_mode(args, world) == "striped" else "shard")}


def multi_gpu_mode(args, world):
    """striped: NVLink-striped database, queries split over the ranks (default for weak scaling: per-GPU work is exactly the
    one-GPU step).  exchange: one database shard per rank scanned for ALL queries, candidate cut-offs and hits exchanged
    over NCCL (default for strong scaling: the prefilter's cost follows the residues scanned, not the queries)."""
    if world == 1:
        return "single"
    return args.multi_gpu or ("striped" if args.scaling == "weak" else "exchange")


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
    ap.add_argument("--ref-queries", type=int, default=256, help="queries of the bounded sample the reference CPU build is timed on")
    ap.add_argument("--ref-db-seqs", type=int, default=1_000_000)
    ap.add_argument("--multi-gpu", default=None, choices=["striped", "exchange"], help="N > 1: NVLink-striped database + split queries, or sharded database + NCCL exchange (default: striped for weak, exchange for strong scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the result digest of the fixed verification workload in the warm-up")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

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
    mode = multi_gpu_mode(args, world)
    striped = None
    if mode == "striped":
        from sift4g_b200 import stripes
        all_lens = db_lengths(n_db)
        all_off = np.zeros(n_db + 1, dtype=np.int64)
        np.cumsum(all_lens, out=all_off[1:])
        try:
            striped = stripes.StripedDatabase(ctx, codes, all_off, lo, hi, dist=dist)
            ok = 1
        except Exception as exc:            # no peer mapping on this box (no NVLink / no VMM export): the exchange form needs neither
            print("bench: striped database unavailable (%s); falling back to --multi-gpu exchange" % exc, file=sys.stderr)
            ok = 0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if striped is not None:
                striped.close()
                striped = None
            args.multi_gpu = mode = "exchange"
    if mode == "striped":
        db = striped.db
        del codes
        # this rank's queries: an equal slice of the batch
        qa, qb = n_queries * rank // world, n_queries * (rank + 1) // world
        my_q_codes, my_q_off = q_codes[q_off[qa]:q_off[qb]], q_off[qa:qb + 1] - q_off[qa]
        lens = all_lens
    else:
        db = ctx.database(codes, loc_off, id_base=lo, where=capi.S4G_DEVICE)
        del codes
        my_q_codes, my_q_off = q_codes, q_off
    # One GPU and the striped form run the product call itself: s4g_search on a resident query batch, cells and paths left in
    # HBM (`value`: inputs resident, no bulk D2H; the hit lists cross to the host because the exact selection runs there).
    # The exchange form sequences the stages from Python around its NCCL exchanges (pipeline.DevicePipeline).
    use_search = mode in ("single", "striped")
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)) or 1)
    host_threads = max(1, min(16, (os.cpu_count() or 1) // max(1, local_world)))
    pipe = None
    if use_search:
        ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        Q = ctx.queries(my_q_codes, my_q_off)

        class _Step:
            def step(self):
                return capi.search(ctx, db, Q, mat, 5, args.max_candidates, n_threads=host_threads, want_candidates=False, device_results=True)
        runner = _Step()
    else:
        pipe = pipeline.DevicePipeline(ctx, db, q_codes, q_off, mat, lens, total_res, max_candidates=args.max_candidates, dist=dist if use_dist else None)

        class _Step:
            def step(self):
                return pipe.step()
        runner = _Step()

    peak = ctx.dpx_peak(300)                  # sustained VIADDMNMX.S16x2 lane-ops/s on this device
    roof_gcups = peak * 2 / 6 / 1e9
    gather_peak = ctx.gather_peak(100)        # random 8-byte lookups/s into an L2-resident table: the prefilter's request-rate ceiling

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        r = runner.step()
    barrier()
    parity = parity_digest(ctx, mat, dist if use_dist else None, log=lambda m: print(m, file=sys.stderr), mode=mode) if not args.no_parity else None
    if parity is not None and not parity["sharded_equals_single"]:
        raise SystemExit("bench: the sharded pipeline does not reproduce the single-GPU results (see stderr)")
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
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
            r = runner.step()
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
    # traceback phases of the last timed step on this rank (CUDA events inside the library) against the same DPX roofline
    roofline_align = None
    try:
        al_ms, al_cells = ctx.last_align_profile()
        roofline_align = {"bound": "int_dpx", "unit": "GCUPS", "what": "stage 3 of the last timed step (rank 0): DP cells each phase has to cover / its device time; "
                          "ends = al_forward_packed_kernel (s16x2, columns up to the end cell), begins = al_reverse_packed_kernel (s16x2, reversed prefixes up to the "
                          "begin column), paths = al_band_persistent_kernel (32-bit banded_sw with direction bytes and traceback; cells of the first band SSW tries, "
                          "doublings not counted; peak = half the s16x2 figure)", "phases": {}}
        for name in ("ends", "begins", "paths"):
            pk = roof_gcups if name != "paths" else roof_gcups / 2
            ach = al_cells[name] / (al_ms[name] * 1e-3) / 1e9 if al_ms[name] > 0 else None
            roofline_align["phases"][name] = {"ms": round(al_ms[name], 3), "cells": al_cells[name], "achieved": round(ach, 2) if ach else None,
                                              "peak": round(pk, 2), "frac": round(ach / pk, 4) if ach else None}
    except capi.S4GError:
        pass

    # stage split (one extra, untimed step with synchronisation between stages; rank-local)
    if use_search:
        # the call's own stage clocks (host clock at the stream synchronisations the path has anyway), last timed step of this rank
        sm = r.stage_ms
        split = {"prefilter": round(float(sm["prefilter"]), 3), "score_kernel": round(float(r.sw_kernel_ms), 3),
                 "score_select_other": round(float(sm["score"]) - float(r.sw_kernel_ms), 3), "align": round(float(sm["align"]), 3),
                 "select_host_beside_align": round(float(sm["select"]), 3)}
    else:
        split = stage_split(torch, ctx, pipe)

    # e2e through the host-buffer API
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(torch, ctx, db, pipe, my_q_codes, my_q_off, mat, lens, total_res, args, use_dist, dist, dev, n_queries, mode)

    base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base = cpu_baseline(args)

    if rank == 0:
        kern_gcups = cells_local / (sw_kernel_ms * 1e-3) / 1e9 if sw_kernel_ms > 0 else None
        line = {
            "metric": "sw_gcups", "value": round(gcups, 2), "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "s16x2 (DPX), s32 re-run on overflow",
            "data": "synthetic",
            "config": workload_config(args, world), "db_generation_s": round(gen_s, 2),
            "queries_per_sec": round(n_queries / (ms_per_step * 1e-3), 2),
            "sw_cells_per_step": cells, "pairs_per_step": pairs, "kept_hits_per_step": hits,
            "stages_ms": split,
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "roofline": {"bound": "int_dpx", "kernel": "sw_score_packed2_kernel", "achieved": round(kern_gcups, 2) if kern_gcups else None, "peak": round(roof_gcups, 2),
                         "unit": "GCUPS", "frac": round(kern_gcups / roof_gcups, 4) if kern_gcups else None,
                         "traffic": NCU_TRAFFIC_C2 if (world == 1 and n_queries == 1000 and n_db == 10_000_000 and args.max_candidates == 5000) else None,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch of this workload, ncu --set full capture (profiles/r04r_kernels_digest.md)",
                         "frac_of_kernel_floor": round(kern_gcups / (roof_gcups * 6 / 4.5), 4) if kern_gcups else None,
                         "kernel_floor": "the kernel's own count: 4.5 ALU-pipe instructions per 2 cells (VIADD.16x2 issues on the FMA pipe), against BASELINE.md's 6",
                         "kernel_ms": round(sw_kernel_ms, 3),
                         "peak_source": "measured live on this GPU (no DPX figure in MEASURED_PEAKS.json): %.4e VIADDMNMX.S16x2 lane-ops/s sustained over 300 ms x 2 cells per op / 6 instructions per cell (BASELINE.md)" % peak},
            "roofline_prefilter": {"bound": "hbm", "achieved": round(((total_res + 8 * n_db) if mode == "striped" else ((hi - lo) / n_db * total_res + 8 * (hi - lo))) / (split["prefilter"] * 1e-3) / 1e9, 2) if split["prefilter"] else None,
                                   "peak": hbm_peak(), "unit": "GB/s", "note": "database bytes (1 B/residue + 8 B/sequence) / prefilter stage time"},
        }
        if split.get("prefilter"):
            # the scan issues one presence/rank probe per k-mer position (k = 5): its algorithmic request count against the
            # measured random-lookup rate; the HBM figure above is kept because north_star asks for it
            scanned = float(total_res if mode == "striped" else (hi - lo) / n_db * total_res)
            ach = scanned / (split["prefilter"] * 1e-3)
            line["roofline_prefilter_l2"] = {"bound": "l2_request_rate", "unit": "G lookups/s", "achieved": round(ach / 1e9, 2), "peak": round(gather_peak / 1e9, 2),
                                             "frac": round(ach / gather_peak, 4),
                                             "note": "k-mer positions scanned per second (one random 8-byte probe of the 8 MB presence/rank table each; entry and hit loads not "
                                                     "counted) / random-lookup rate measured live on this GPU (s4g_measure_gather_peak, 100 ms); ncu: profiles/r04r_kernels_digest.md (one row per chunk), profiles/r02i_pf_scan_digest.md"}
            if world == 1 and n_queries == 1000 and n_db == 10_000_000 and args.max_candidates == 5000 and args.query_shape == "uniform":
                # all L2 sectors of the scan kernels of one step (probes + entries + hits + residues), from the ncu capture of this
                # workload, over the measured prefilter stage time: what the request-rate ceiling is actually spent on
                line["roofline_prefilter_l2"]["l2_sectors_per_step"] = NCU_PF_SECTORS_C2
                line["roofline_prefilter_l2"]["sectors_frac"] = round(NCU_PF_SECTORS_C2 / (split["prefilter"] * 1e-3) / gather_peak, 4)
                line["roofline_prefilter_l2"]["sectors_source"] = ("lts__t_sectors.sum over the 15 pf_scan_kernel launches of a step (profiles/r04r_kernels_digest.md: chunks of "
                                                                   "sequences >= 235 aa run at 0.67-0.87 of the ceiling, the short-sequence chunks at 0.15-0.39)")
        if roofline_align is not None:
            line["roofline_align"] = roofline_align
        if line["roofline_prefilter"]["achieved"]:
            line["roofline_prefilter"]["frac"] = round(line["roofline_prefilter"]["achieved"] / line["roofline_prefilter"]["peak"], 4)
        if parity is not None:
            line["parity"] = parity
        if e2e is not None:
            line["e2e"] = e2e
        if base is not None:
            line["cpu_baseline"] = base
        print(json.dumps(line), file=out, flush=True)
    if pipe is not None:
        pipe.close()
    else:
        Q.close()
    if striped is not None:
        striped.close()
    else:
        db.close()
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()


# DRAM bytes of one sw_score_packed2_kernel launch at the default C2 workload (1.391 GB read + 24.5 MB written), from the
# ncu --set full capture summarised in profiles/r04r_kernels_digest.md.  The kernel is integer-issue bound; its algorithmic
# DRAM traffic is the candidates' residues (5 M targets, ~107 residues each, fetched in 32-byte sectors).
NCU_TRAFFIC_C2 = 1390900000 + 24490496
# L2 sectors of the 15 pf_scan_kernel launches of one step at the same workload (profiles/r04r_kernels_digest.md)
NCU_PF_SECTORS_C2 = 4766630000


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


def run_e2e(torch, ctx, db, pipe, q_codes, q_off, mat, lens, total_res, args, use_dist, dist, dev, n_queries, mode="single"):
    """Same step end to end with HOST buffers, host<->device copies inside the timed region.
    One GPU: the product boundary itself -- the query batch is uploaded (s4g_queries_create) and ONE C-ABI call, s4g_search,
    returns candidate lists, kept hits with E-values and alignments in host memory (pipeline.search_host; no torch, no
    Python between the stages).  N > 1 (one process per GPU under torchrun): the sharded pipeline with its NCCL exchanges,
    queries uploaded and every result copied back each step (DevicePipeline.step(e2e=True))."""
    from sift4g_b200 import pipeline
    n = max(1, min(args.steps, 3))
    if not use_dist or mode == "striped":
        # q_codes / q_off: this rank's queries (the whole batch on one GPU)
        ctx.sync()
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
        run = lambda: pipeline.search_host(ctx, db, q_codes, q_off, mat, max_candidates=args.max_candidates, want_candidates=True, align=True,
                                           n_threads=max(1, min(16, (os.cpu_count() or 1) // max(1, local_world))))
        out = run()
        if use_dist:
            dist.barrier()
        t0 = time.time()
        for _ in range(n):
            out = run()
            checksum = int(out.path_off[-1]) + int(out.hit_off[-1]) + int(out.cand_off[-1])       # results are on the host
        dt = (time.time() - t0) / n
        ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        cells, h2d, d2h, hits = float(out.sw_cells), float(out.h2d_bytes), float(out.d2h_bytes), float(out.n_hits)
        if use_dist:
            t = torch.tensor([dt, cells, h2d, d2h, hits, float(checksum)], dtype=torch.float64, device=dev)
            tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
            dt, cells, h2d, d2h, hits, checksum = float(tm[0]), float(ts[1]), float(ts[2]), float(ts[3]), float(ts[4]), int(ts[5])
        return {"value": round(cells / dt / 1e9, 2), "unit": "GCUPS", "ms_per_step": round(dt * 1e3, 3), "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "queries_per_sec": round(n_queries / dt, 2),
                "stages_ms": {k: round(float(v), 3) for k, v in out.stage_ms.items()}, "kept_hits": int(hits), "result_checksum": checksum,
                "timed": "host wall clock around s4g_queries_create + s4g_search (C ABI, host buffers): queries H2D; candidate lists, kept hits, E-values, alignment cells and paths D2H every step"
                         + ("; every rank for its own queries against the striped database, max over ranks (stages_ms: rank 0)" if use_dist else "")}
    pipe.step(e2e=True)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    cells = h2d = d2h = 0
    for _ in range(n):
        r = pipe.step(e2e=True)
        cells, h2d, d2h = r.sw_cells, r.h2d_bytes, r.d2h_bytes
    torch.cuda.synchronize()
    dt = (time.time() - t0) / n
    t = torch.tensor([dt, float(cells), float(h2d), float(d2h)], dtype=torch.float64, device=dev)
    tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
    dt, cells, h2d, d2h = float(tm[0]), float(ts[1]), float(ts[2]), float(ts[3])
    return {"value": round(cells / dt / 1e9, 2), "unit": "GCUPS", "ms_per_step": round(dt * 1e3, 3), "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "queries_per_sec": round(n_queries / dt, 2),
            "timed": "host wall clock around pipeline.DevicePipeline.step(e2e=True): queries H2D, this rank's share of the candidate lists + survivor scores + alignments D2H every step; max over ranks"}


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
