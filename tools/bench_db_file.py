"""Database ingest on the GPU box: FASTA parse (1 thread vs all threads) vs packed .s4gdb open.
Writes one JSON line per measurement.  python tools/bench_db_file.py [n_seqs]"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sift4g_b200 import capi  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
    tmp = tempfile.mkdtemp()
    rng = np.random.default_rng(11)
    lens = bench.db_lengths(n)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    codes = rng.integers(0, 20, int(off[-1]), dtype=np.uint8)
    txt = (np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)[codes]).tobytes()
    fa = tmp + "/d.fa"
    with open(fa, "wb") as f:                     # 60-column FASTA like UniRef
        for i in range(n):
            s = txt[off[i]:off[i + 1]]
            f.write(b">UniRef90_D%08d synthetic n=1\n" % i)
            f.write(b"\n".join(s[j:j + 60] for j in range(0, len(s), 60)))
            f.write(b"\n")
    size = os.path.getsize(fa)
    out = {"fasta_bytes": size, "n_seqs": n, "n_residues": int(off[-1]), "host_cores": os.cpu_count()}
    packed = tmp + "/d.s4gdb"
    for th in (1, os.cpu_count()):
        os.environ["S4G_HOST_THREADS"] = str(th)
        t = time.time()
        capi.pack_fasta(fa, packed)
        out["pack_fasta_s_%d_threads" % th] = round(time.time() - t, 3)
    os.environ.pop("S4G_HOST_THREADS")
    ctx = capi.Context(0)
    for name, fn, path in (("open_fasta_s", ctx.database_from_fasta, fa), ("open_packed_s", ctx.database_from_packed, packed)):
        for rep in range(2):
            t = time.time()
            D = fn(path)
            dt = time.time() - t
            assert D.n_seqs == n and D.n_residues == int(off[-1])
            D.close()
        out[name] = round(dt, 3)
    out["open_fasta_MBps"] = round(size / out["open_fasta_s"] / 1e6, 1)
    out["open_packed_MBps"] = round(os.path.getsize(packed) / out["open_packed_s"] / 1e6, 1)
    ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
