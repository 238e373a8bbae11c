"""Quick SW-score throughput probe (not the bench): random DB, random candidate lists."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from sift4g_b200 import capi, synth
from oracle import oracle as O

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ncand = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
qlo = int(sys.argv[3]) if len(sys.argv) > 3 else 100
qhi = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
ndb = 200000
rng = np.random.default_rng(1)
lens = np.clip(np.exp(rng.normal(5.6, 0.6, size=ndb)), 30, 35000).astype(np.int64)
off = np.zeros(ndb + 1, dtype=np.int64); off[1:] = np.cumsum(lens)
codes = rng.integers(0, 20, size=int(off[-1]), dtype=np.uint8)
qlens = rng.integers(qlo, qhi + 1, size=nq)
qoff = np.zeros(nq + 1, dtype=np.int64); qoff[1:] = np.cumsum(qlens)
qcodes = rng.integers(0, 20, size=int(qoff[-1]), dtype=np.uint8)
ctx = capi.Context(0)
D = ctx.database(codes, off); Q = ctx.queries(qcodes, qoff)
ids = np.concatenate([np.sort(rng.choice(ndb, size=ncand, replace=False)) for _ in range(nq)]).astype(np.uint32)
coff = np.arange(nq + 1, dtype=np.int64) * ncand
cells = float(sum(int(qlens[q]) * int(lens[ids[q * ncand:(q + 1) * ncand]].sum()) for q in range(nq)))
mat = O.blosum62()
peak = ctx.dpx_peak(200)
print("dpx peak lane-ops/s %.4e -> roofline %.1f GCUPS" % (peak, peak * 2 / 6 / 1e9))
for it in range(4):
    t0 = time.time()
    out = capi.sw_score(ctx, D, Q, ids, coff, mat)
    t1 = time.time()
    ms = ctx.last_sw_kernel_ms()
    print("iter %d: e2e %.1f ms, kernel %.3f ms, cells %.3e, kernel GCUPS %.1f, e2e GCUPS %.1f, frac of roofline %.3f" % (
        it, (t1 - t0) * 1e3, ms, cells, cells / ms / 1e6, cells / (t1 - t0) / 1e9, cells / ms / 1e6 / (peak * 2 / 6 / 1e9)))
print("checksum", int(out.sum()))
