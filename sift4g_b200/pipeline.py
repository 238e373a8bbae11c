"""Host-side orchestration of the hot path over the C ABI (Python mirror of what the C++ shims in
sift4g_b200/host do for the CLI): prefilter -> SW scores -> E-value selection -> traceback.

Three entry points, same results:
  * search_host(...)   -- ONE C-ABI call (s4g_search) with host buffers in and out; candidate lists, scores and survivors
                          stay in HBM between the stages, the exact selection runs on C++ host threads.  This is the call a
                          user of the library makes and what bench.py times as `e2e` on one GPU.
  * run_host(...)      -- the same path stage by stage through the host-buffer form of every stage call (s4g_prefilter,
                          s4g_sw_score, s4g_select_hits, s4g_sw_align): every intermediate result crosses the bus; kept for
                          the parity tests, which look at every stage's output.
  * DevicePipeline     -- inputs and intermediate results stay in HBM as torch tensors (bench.py's `value`); with
                          torch.distributed initialised the database is sharded, one resident shard per rank, and the
                          per-query cut-offs / kept hits are exchanged over NCCL (the only exchanges on the path).
torch is plumbing here (device buffers, streams, NCCL); all compute is in libsift4g_b200.so.
"""
import os
import sys
import threading
import time

import numpy as np

from . import capi


class Result:
    __slots__ = ("cand_ids", "cand_off", "scores", "pair_q", "pair_t", "pair_score", "evalue", "hit_off", "coords", "paths",
                 "path_off", "sw_cells", "n_pairs", "timings", "h2d_bytes", "d2h_bytes", "sw_kernel_ms")

    def __init__(self):
        for s in self.__slots__:
            setattr(self, s, None)


def _segment_sums(values, off):
    cs = np.zeros(len(values) + 1, dtype=np.int64)
    np.cumsum(values, out=cs[1:])
    return cs[off[1:]] - cs[off[:-1]]


def _ragged(ids2d, cnt):
    nq, n = ids2d.shape
    if int(cnt.min()) == n:
        return ids2d.reshape(-1), np.arange(nq + 1, dtype=np.int64) * n
    mask = np.arange(n)[None, :] < cnt[:, None]
    off = np.zeros(nq + 1, dtype=np.int64)
    off[1:] = np.cumsum(cnt)
    return np.ascontiguousarray(ids2d[mask]), off


def search_host(ctx, db, q_codes, q_off, matrix, k=5, max_candidates=5000, gap_open=10, gap_extend=1, max_evalue=1e-4, max_alignments=400,
                n_threads=0, want_candidates=True, align=True):
    """Whole hot path through s4g_search: query batch uploaded, one call, results as numpy views of the library's pinned
    buffers (capi.SearchOutput; valid until the next call on the context)."""
    Q = ctx.queries(q_codes, q_off)
    try:
        out = capi.search(ctx, db, Q, matrix, k, max_candidates, gap_open, gap_extend, max_evalue, max_alignments, n_threads,
                          want_candidates=want_candidates, want_alignments=align)
    finally:
        Q.close()
    out.h2d_bytes += q_codes.nbytes + q_off.nbytes
    return out


def run_host(ctx, db, q_codes, q_off, matrix, db_lens, k=5, max_candidates=5000, gap_open=10, gap_extend=1, max_evalue=1e-4,
             max_alignments=400, names=None, align=True):
    """Whole hot path with host buffers.  db: capi.Database (resident shard covering the whole database)."""
    r = Result()
    q_lens = np.diff(q_off).astype(np.int32)
    Q = ctx.queries(q_codes, q_off)
    ids2d, _, cnt = capi.prefilter(ctx, db, Q, k, max_candidates, sorted_by_id=True)
    r.cand_ids, r.cand_off = _ragged(ids2d, cnt)
    r.scores = capi.sw_score(ctx, db, Q, r.cand_ids, r.cand_off, matrix, gap_open, gap_extend)
    cand_lens = db_lens[r.cand_ids - db.id_base].astype(np.int32)
    r.sw_cells = int(np.dot(q_lens.astype(np.int64), _segment_sums(cand_lens, r.cand_off)))
    r.n_pairs = len(r.cand_ids)
    cand_names = [names[i] for i in r.cand_ids] if names is not None else None
    r.pair_q, r.pair_t, r.pair_score, r.evalue, r.hit_off = capi.select_hits(
        ctx, q_lens, r.cand_ids, r.cand_off, r.scores, cand_lens, db.n_residues, gap_open, gap_extend, max_evalue, max_alignments, cand_names)
    r.h2d_bytes = q_codes.nbytes + q_off.nbytes + r.cand_ids.nbytes + r.cand_off.nbytes
    r.d2h_bytes = ids2d.nbytes + cnt.nbytes + r.scores.nbytes
    if align and len(r.pair_q):
        cap = int(q_lens[r.pair_q].astype(np.int64).sum() + db_lens[r.pair_t - db.id_base].astype(np.int64).sum()) + 16
        r.coords, paths = _align_host(ctx, db, Q, r, matrix, gap_open, gap_extend, cap)
        r.paths, r.path_off = paths
        r.h2d_bytes += 3 * r.pair_q.nbytes
        r.d2h_bytes += r.coords.nbytes + int(r.path_off[-1]) + r.path_off.nbytes
    Q.close()
    return r


def _align_host(ctx, db, Q, r, matrix, go, ge, cap):
    import ctypes as C
    n = len(r.pair_q)
    coords = np.zeros((n, 4), dtype=np.int32)
    paths = np.zeros(cap, dtype=np.uint8)
    off = np.zeros(n + 1, dtype=np.int64)
    m = np.ascontiguousarray(matrix, dtype=np.int32)
    ctx.check(ctx.lib.s4g_sw_align(ctx.h, db.h, Q.h, n, r.pair_q.ctypes.data, r.pair_t.ctypes.data, r.pair_score.ctypes.data, m.ctypes.data,
                                   go, ge, coords.ctypes.data, paths.ctypes.data, cap, off.ctypes.data, capi.S4G_HOST))
    return coords, (paths[:off[-1]], off)


class DevicePipeline:
    """Device-resident hot path for one rank.  With `dist` (torch.distributed, NCCL) every rank holds one
    database shard; candidates and kept hits are merged across ranks."""

    def __init__(self, ctx, db, q_codes, q_off, matrix, db_lens, total_residues, k=5, max_candidates=5000, gap_open=10, gap_extend=1,
                 max_evalue=1e-4, max_alignments=400, dist=None):
        import torch
        self.torch = torch
        self.ctx, self.db, self.matrix = ctx, db, np.ascontiguousarray(matrix, dtype=np.int32)
        self.k, self.N, self.go, self.ge = k, max_candidates, gap_open, gap_extend
        self.max_evalue, self.max_alignments = max_evalue, max_alignments
        self.dist = dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.dev = torch.device("cuda", ctx.device)
        self.q_lens = np.diff(q_off).astype(np.int32)
        self.nq = len(self.q_lens)
        self.db_lens = db_lens                      # host, this shard
        self.total_residues = int(total_residues)   # whole database (E-value length)
        self.Q = ctx.queries(q_codes, q_off)
        nq, N = self.nq, self.N
        # multi-GPU: rank r owns the queries [r * S, (r + 1) * S) for the cut-off exchange (rows padded to W * S)
        self.slice = (nq + self.world - 1) // self.world
        rows = self.slice * self.world
        self.t_ids = torch.zeros((rows, N), dtype=torch.int32, device=self.dev)
        self.t_sc = torch.zeros((rows, N), dtype=torch.float32, device=self.dev)
        self.t_cnt = torch.zeros(rows, dtype=torch.int32, device=self.dev)
        if self.world > 1:
            self.g_ids = torch.zeros((self.world, self.slice, N), dtype=torch.int32, device=self.dev)
            self.g_sc = torch.zeros((self.world, self.slice, N), dtype=torch.float32, device=self.dev)
            self.g_cnt = torch.zeros((self.world, self.slice), dtype=torch.int32, device=self.dev)
            self.cut_own = torch.zeros(self.slice, dtype=torch.int64, device=self.dev)
            self.cut_all = torch.zeros(rows, dtype=torch.int64, device=self.dev)
            self.own_cnt = torch.zeros(nq, dtype=torch.int32, device=self.dev)
        self.col = torch.arange(N, device=self.dev, dtype=torch.int32)[None, :]
        self.t_db_lens = torch.from_numpy(np.ascontiguousarray(db_lens, dtype=np.int64)).to(self.dev)
        self.t_q_lens = torch.from_numpy(self.q_lens.astype(np.int64)).to(self.dev)
        self.q_host = (q_codes, q_off)
        # host threads of the exact selection: this rank's share of the cores (one process per GPU on the box), at most
        # 16 -- spawning a thread per core of a large host costs more than the selection itself
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(self.world)) or 1)
        self.host_threads = max(1, min(16, (os.cpu_count() or 1) // max(1, local_world)))
        # 1: score, select and align the whole batch in turn; 2: in two halves of the query batch, the exact host selection of
        # one half running beside the GPU work of the other.  Two launches per stage cost ~10 ms of GPU time at configs[1]
        # (tails, per-call sorts; 133.3 vs 136.5 ms per step on one GPU), so halves only pay where the ranks of a box leave
        # each other so few host cores that the selection takes longer than that (8 ranks on 16 cores: ~12 ms, 154 -> 151 ms).
        self.parts = int(os.environ.get("S4G_PARTS", "0")) or (2 if self.host_threads <= 2 else 1)
        if self.world > 1:
            # the number of halves decides how many hit-merge collectives a rank issues per step: every rank must use the same
            # one, whatever its own core count or environment says (the smallest wins)
            t = torch.tensor([self.parts], dtype=torch.int32, device=self.dev if dist.get_backend() == "nccl" else "cpu")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            self.parts = int(t.item())
        ctx.set_stream(torch.cuda.current_stream(self.dev).cuda_stream)

    def close(self):
        self.Q.close()

    def step(self, align=True, e2e=False, stages=None):
        """One pass of the hot path.  e2e=True: the query batch is uploaded from host memory first and every
        result (candidate lists, survivor scores, alignments) is copied back to the host at the end.
        stages: optional dict; when given, the stream is synchronised after every stage and the stage's wall time
        (ms) is added under its name (a diagnostic mode: the timed bench steps run without it)."""
        torch, ctx, db = self.torch, self.ctx, self.db
        r = Result()
        trace = os.environ.get("S4G_TRACE", "") not in ("", "0")
        timed = trace or stages is not None
        marks = []
        if timed:
            torch.cuda.synchronize()
            t_last = [time.time()]

        def mark(name):
            if timed:
                torch.cuda.synchronize()
                now = time.time()
                marks.append("%s=%.3fms" % (name, (now - t_last[0]) * 1e3))
                if stages is not None:
                    stages[name] = stages.get(name, 0.0) + (now - t_last[0]) * 1e3
                    if name == "sw_score":
                        try:
                            stages["sw_kernel"] = stages.get("sw_kernel", 0.0) + self.ctx.last_sw_kernel_ms()
                        except capi.S4GError:
                            pass                      # no pair was scored on this rank
                t_last[0] = now
        if e2e:
            self.Q.close()
            self.Q = ctx.queries(*self.q_host)
            r.h2d_bytes = self.q_host[0].nbytes + self.q_host[1].nbytes
        nq, N = self.nq, self.N
        lo, hi = db.id_base, db.id_base + db.n_seqs
        # ---- stage 1 ----
        if self.world == 1:
            capi.prefilter(ctx, db, self.Q, self.k, N, True, out=(self.t_ids, self.t_sc, self.t_cnt), where=capi.S4G_DEVICE)
            ids, cnt = self.t_ids, self.t_cnt
            mark("prefilter")
        else:
            # best-first rows of this shard; the owners of the queries find the global cut-off keys, every shard keeps
            # the prefix of its own rows inside the cut-off (no row leaves its GPU except for the owner's search)
            W, S = self.world, self.slice
            capi.prefilter(ctx, db, self.Q, self.k, N, False, out=(self.t_ids, self.t_sc, self.t_cnt), where=capi.S4G_DEVICE)
            mark("prefilter")
            self.dist.all_to_all_single(self.g_ids.view(-1), self.t_ids.view(-1))
            self.dist.all_to_all_single(self.g_sc.view(-1), self.t_sc.view(-1))
            self.dist.all_to_all_single(self.g_cnt.view(-1), self.t_cnt)
            mark("exchange_rows")
            ctx.check(ctx.lib.s4g_topn_cutoff(ctx.h, W, S, N, self.g_ids.data_ptr(), self.g_sc.data_ptr(), self.g_cnt.data_ptr(), self.cut_own.data_ptr()))
            self.dist.all_gather_into_tensor(self.cut_all, self.cut_own)
            ctx.check(ctx.lib.s4g_cutoff_counts(ctx.h, nq, N, self.t_ids.data_ptr(), self.t_sc.data_ptr(), self.t_cnt.data_ptr(), self.cut_all.data_ptr(),
                                                self.own_cnt.data_ptr()))
            ids, cnt = self.t_ids[:nq], self.own_cnt
            mark("cutoff")
        # this rank's candidates (ids are uint32 bit patterns in int32 tensors)
        own = self.col < cnt[:, None]
        cand_ids = ids[own].contiguous()
        cand_off = torch.zeros(nq + 1, dtype=torch.int64, device=self.dev)
        cand_off[1:] = torch.cumsum(cnt, 0)
        n_pairs = int(cand_ids.numel())
        scores = torch.empty(max(n_pairs, 1), dtype=torch.int32, device=self.dev)
        mark("own_candidates")
        # ---- stages 2 + 3, in two halves of the query batch: while the host selects the hits of one half (exact libm
        # E-values, a few ms on a few cores) the GPU scores or aligns the other half
        r.n_pairs = n_pairs
        r.cand_ids, r.cand_off, r.scores = cand_ids, cand_off, scores     # device tensors (all candidates)
        n_parts = self.parts
        overlap = n_parts > 1 and nq >= 2 and stages is None and os.environ.get("S4G_NO_OVERLAP", "") in ("", "0")
        mid = nq // 2 if (nq >= 2 and n_parts > 1) else nq
        p_mid = int(cand_off[mid].item()) if 0 < mid < nq else n_pairs
        halves = [(0, mid, 0, p_mid), (mid, nq, p_mid, n_pairs)] if mid < nq else [(0, nq, 0, n_pairs)]
        cap = max(n_pairs, 1)
        if getattr(self, "_scr_cap", 0) < cap:
            self._scr = [torch.empty(cap, dtype=torch.int32, device=self.dev) for _ in range(4)]
            self._scr_cnt = torch.zeros(1, dtype=torch.int32, device=self.dev)
            self._scr_cap = cap
        cells_dev = None
        if n_pairs:
            # algorithmic SW cells of this step (statistic): sum_q len(q) * sum_{t in cand(q)} len(t)
            lens_dev = self.t_db_lens[(cand_ids.to(torch.int64) & 0xffffffff) - lo]
            cs = torch.zeros(n_pairs + 1, dtype=torch.int64, device=self.dev)
            cs[1:] = torch.cumsum(lens_dev, 0)
            seg = cs[cand_off[1:]] - cs[cand_off[:-1]]
            cells_dev = (seg * self.t_q_lens).sum()
        sw_ms = [0.0]
        n_surv = [0]

        def score_half(hx, qa, qb, pa, pb):
            """SW scores + E-value screen of the queries [qa, qb); survivors to pinned host arrays"""
            n = pb - pa
            off = (cand_off.clamp(min=pa, max=pb) - pa).contiguous()              # other half's queries: empty lists
            ids_h, sc_h = cand_ids[pa:pb], scores[pa:pb]
            if n:
                capi.sw_score(ctx, db, self.Q, ids_h, off, self.matrix, self.go, self.ge, out=sc_h, where=capi.S4G_DEVICE)
                try:
                    sw_ms[0] += ctx.last_sw_kernel_ms()
                except capi.S4GError:
                    pass
            mark("sw_score")
            s_q, s_id, s_sc, s_tl = self._scr
            ctx.check(ctx.lib.s4g_evalue_screen(ctx.h, db.h, self.Q.h, ids_h.data_ptr(), off.data_ptr(), n, sc_h.data_ptr(),
                                                b"BLOSUM_62", self.total_residues, self.go, self.ge, self.max_evalue, s_q.data_ptr(), s_id.data_ptr(),
                                                s_sc.data_ptr(), s_tl.data_ptr(), self._scr_cnt.data_ptr()))
            n_s = int(self._scr_cnt.item())
            n_surv[0] += n_s
            surv = [self._to_host("%s%d" % (nm, hx), t[:n_s]) for nm, t in (("s_q", s_q), ("s_id", s_id), ("s_sc", s_sc), ("s_tl", s_tl))]
            torch.cuda.current_stream(self.dev).synchronize()
            mark("screen_d2h")
            return surv

        def select_half(surv, box):
            h_q = surv[0].numpy().view(np.uint32)
            h_off = np.searchsorted(h_q, np.arange(nq + 1), side="left").astype(np.int64)
            box.append(capi.select_hits(ctx, self.q_lens, surv[1].numpy().view(np.uint32), h_off, surv[2].numpy(), surv[3].numpy(),
                                        self.total_residues, self.go, self.ge, self.max_evalue, self.max_alignments, n_threads=self.host_threads))
            # target residues of all survivors: an upper bound for those of the kept hits (a subset), which sizes the path
            # buffer without a 170 k-element random gather from the shard's length array on the host (2.5-5 ms per step)
            box.append(int(surv[3].numpy().sum(dtype=np.int64)))

        def finish_half(box):
            pq, pt, ps, ev, hoff = box[0]
            mark("select_hits")
            if self.world > 1:
                pq, pt, ps, ev, hoff = self._merge_hits(pq, pt, ps, ev, hoff, lo, hi)
                mark("merge_hits")
            return pq, pt, ps, ev, hoff, box[1]

        def start_select(surv):
            box = []
            if not overlap:
                select_half(surv, box)
                return None, box
            th = threading.Thread(target=select_half, args=(surv, box))
            th.start()
            return th, box

        parts = []            # per half: (pq, pt, ps, ev, hoff)
        aligned = []          # per half: (coords, poff, n_path)
        path_buf = [None, 0]  # shared path buffer, bytes used

        def align_half(hits):
            pq, pt, ps = hits[0], hits[1], hits[2]
            if not (align and len(pq)):
                mark("align")
                return
            d_pq = torch.from_numpy(pq.view(np.int32)).to(self.dev)
            d_pt = torch.from_numpy(pt.view(np.int32)).to(self.dev)
            d_ps = torch.from_numpy(ps).to(self.dev)
            cap_h = int(self.q_lens[pq].sum(dtype=np.int64)) + hits[5] + 16
            if path_buf[0] is None:
                path_buf[0] = torch.empty(cap_h * (2 if len(halves) > 1 else 1) + 1024, dtype=torch.uint8, device=self.dev)
            elif path_buf[0].numel() < path_buf[1] + cap_h:
                grown = torch.empty(path_buf[1] + cap_h, dtype=torch.uint8, device=self.dev)
                grown[:path_buf[1]] = path_buf[0][:path_buf[1]]
                path_buf[0] = grown
            d_coords = torch.empty((len(pq), 4), dtype=torch.int32, device=self.dev)
            d_poff = torch.empty(len(pq) + 1, dtype=torch.int64, device=self.dev)
            out = path_buf[0][path_buf[1]:]
            ctx.check(ctx.lib.s4g_sw_align(ctx.h, db.h, self.Q.h, len(pq), d_pq.data_ptr(), d_pt.data_ptr(), d_ps.data_ptr(), self.matrix.ctypes.data,
                                           self.go, self.ge, d_coords.data_ptr(), out.data_ptr(), cap_h, d_poff.data_ptr(), capi.S4G_DEVICE))
            n_path = int(d_poff[-1].item())
            aligned.append((d_coords, d_poff[:-1] + path_buf[1], n_path))
            path_buf[1] += n_path
            mark("align")

        pending = None        # (thread, box) of the half whose hits are being selected
        for hx, (qa, qb, pa, pb) in enumerate(halves):
            surv = score_half(hx, qa, qb, pa, pb)
            if pending is not None:
                # the previous half was selected while this half was scored; align it while this half is selected
                if pending[0] is not None:
                    pending[0].join()
                prev = finish_half(pending[1])
                pending = start_select(surv)
                parts.append(prev)
                align_half(prev)
            else:
                pending = start_select(surv)
        if pending[0] is not None:
            pending[0].join()
        last = finish_half(pending[1])
        parts.append(last)
        align_half(last)
        r.sw_cells = int(cells_dev.item()) if cells_dev is not None else 0
        r.sw_kernel_ms = sw_ms[0]
        n_s = n_surv[0]
        pq = np.concatenate([x[0] for x in parts]); pt = np.concatenate([x[1] for x in parts])
        ps = np.concatenate([x[2] for x in parts]); ev = np.concatenate([x[3] for x in parts])
        hoff = parts[0][4].copy()
        for x in parts[1:]:
            hoff += x[4]
        r.pair_q, r.pair_t, r.pair_score, r.evalue, r.hit_off = pq, pt, ps, ev, hoff
        if aligned:
            r.coords = torch.cat([a_[0] for a_ in aligned]) if len(aligned) > 1 else aligned[0][0]
            r.path_off = torch.cat([a_[1] for a_ in aligned] + [torch.tensor([path_buf[1]], dtype=torch.int64, device=self.dev)])
            r.paths = path_buf[0]
        mark("align")
        if trace and self.rank == 0:
            print("[s4g trace] step: " + " ".join(marks), file=sys.stderr)
        if e2e:
            # results to the host through grow-only pinned buffers: async copies on the stream, one synchronise
            out_ids = ids if self.world == 1 else cand_ids          # sharded: this rank's share of every candidate list
            host = [self._to_host("ids", out_ids), self._to_host("cnt", cnt)]
            r.d2h_bytes = out_ids.numel() * 4 + cnt.numel() * 4 + n_s * 16 + 4
            r.h2d_bytes += 3 * 4 * len(pq)
            if r.coords is not None:
                host += [self._to_host("coords", r.coords), self._to_host("poff", r.path_off)]
                # every path byte the capacity bound allows is at most 2x the real total; copy only the real ones
                torch.cuda.current_stream(self.dev).synchronize()
                n_path = int(host[-1][-1])
                host.append(self._to_host("paths", r.paths[:n_path]))
                r.d2h_bytes += r.coords.numel() * 4 + n_path + r.path_off.numel() * 8
            torch.cuda.current_stream(self.dev).synchronize()
            r.timings = host
        return r

    def _to_host(self, name, t):
        """Asynchronous device-to-host copy of `t` into a pinned buffer owned by the pipeline (valid until the next
        step); the caller synchronises the stream."""
        torch = self.torch
        t = t.reshape(-1)
        pool = self.__dict__.setdefault("_pinned", {})
        buf = pool.get(name)
        if buf is None or buf.dtype != t.dtype or buf.numel() < t.numel():
            buf = torch.empty(max(int(t.numel() * 1.25), 1), dtype=t.dtype, pin_memory=True)
            pool[name] = buf
        out = buf[:t.numel()]
        out.copy_(t, non_blocking=True)
        return out

    def _merge_hits(self, pq, pt, ps, ev, hoff, lo, hi):
        return merge_hits(self.torch, self.dist, self.dev, self.nq, self.max_alignments, pq, pt, ps, ev, hoff, lo, hi, ctx=self.ctx)


def merge_hits(torch, dist, dev, nq, M, pq, pt, ps, ev, hoff, lo, hi, ctx=None):
    """Global top M hits per query over all ranks under (E asc, score desc, id asc) -- dbAlignmentsMerge's order and
    truncation (sw/post_proc.c:299-339,432-456); every rank keeps the hits whose targets it owns ([lo, hi)), so the
    traceback needs no further exchange.  One small all-gather of the per-query counts tells every rank which queries
    have more than M hits over all shards; only their {E, score, id} rows are exchanged (a second all-gather, padded to
    the largest rank) and cut by s4g_merge_hits.  Queries with at most M hits in total keep their local lists as is."""
    W = dist.get_world_size()
    cuda = dev.type == "cuda"

    def gather(t):
        if cuda:
            out = torch.empty((W,) + tuple(t.shape), dtype=t.dtype, device=dev)
            dist.all_gather_into_tensor(out.view(-1), t.to(dev).view(-1))
            return out.cpu().numpy()
        parts = [torch.empty_like(t) for _ in range(W)]
        dist.all_gather(parts, t)
        return torch.stack(parts).numpy()

    cnt = np.diff(hoff).astype(np.int64)
    all_cnt = gather(torch.from_numpy(cnt))                                   # [W, nq]
    over = np.nonzero(all_cnt.sum(axis=0) > M)[0]
    if len(over) == 0:
        return pq, pt, ps, ev, hoff
    is_over = np.zeros(nq, dtype=bool)
    is_over[over] = True
    mine = is_over[pq]
    sub_cnt = np.ascontiguousarray(all_cnt[:, over])                          # [W, n_over]
    stride = max(int(sub_cnt.sum(axis=1).max()), 1)
    loc = np.zeros((stride, 3), dtype=np.float64)
    n_mine = int(mine.sum())
    loc[:n_mine, 0] = ev[mine]; loc[:n_mine, 1] = ps[mine]; loc[:n_mine, 2] = pt[mine]
    allh = gather(torch.from_numpy(loc))                                      # [W, stride, 3]
    oq, ot, osc, oev, _ = capi.merge_hits(ctx, allh, sub_cnt, len(over), M, lo, hi)
    keep = ~mine
    nq_ = np.concatenate([pq[keep], over[oq].astype(np.uint32)])
    order = np.argsort(nq_, kind="stable")
    pq2 = nq_[order]
    pt2 = np.concatenate([pt[keep], ot])[order]
    ps2 = np.concatenate([ps[keep], osc])[order]
    ev2 = np.concatenate([ev[keep], oev])[order]
    hoff2 = np.searchsorted(pq2, np.arange(nq + 1), side="left").astype(np.int64)
    return pq2, pt2, ps2, ev2, hoff2


def merge_hits_numpy(allh, counts, nq, M, lo, hi):
    """numpy statement of s4g_merge_hits (tests compare the two)."""
    W = allh.shape[0]
    start = np.zeros((W, nq + 1), dtype=np.int64)
    start[:, 1:] = np.cumsum(counts, axis=1)
    oq, ot, osc, oev, off = [], [], [], [], [0]
    for q in range(nq):
        rows = np.concatenate([allh[r, start[r, q]:start[r, q + 1]] for r in range(W)])
        order = np.lexsort((rows[:, 2], -rows[:, 1], rows[:, 0]))[:M]
        rows = rows[order]
        mine = rows[(rows[:, 2] >= lo) & (rows[:, 2] < hi)]
        oq.append(np.full(len(mine), q, dtype=np.uint32)); ot.append(mine[:, 2].astype(np.uint32))
        osc.append(mine[:, 1].astype(np.int32)); oev.append(mine[:, 0]); off.append(off[-1] + len(mine))
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    return cat(oq, np.uint32), cat(ot, np.uint32), cat(osc, np.int32), cat(oev, np.float64), np.array(off, dtype=np.int64)
