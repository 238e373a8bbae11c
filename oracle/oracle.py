"""TEST INFRASTRUCTURE -- ctypes binding of oracle/libs4g_oracle.so (the plain-C CPU restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package (sift4g_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile the C restatement (and, when /root/reference exists, oracle/_ref)."""
    subprocess.run(["make", "-s", "-C", _HERE, "oracle"], check=True)
    if os.path.isdir("/root/reference/sift4g/src"):
        subprocess.run(["make", "-s", "-j8", "-C", _HERE, "ref"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libs4g_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        u8p, i32p, i64p, u32p, f32p, f64p = (C.POINTER(t) for t in (C.c_uint8, C.c_int32, C.c_int64, C.c_uint32, C.c_float, C.c_double))
        L.s4g_oracle_blosum62.argtypes = [i32p]
        L.s4g_oracle_encode.argtypes = [C.c_char_p, C.c_int64, u8p]
        L.s4g_oracle_encode.restype = C.c_int64
        L.s4g_oracle_prefilter.argtypes = [u8p, i64p, C.c_int64, u8p, i64p, C.c_int32, C.c_int32, C.c_int32, u32p, f32p, u32p, f32p]
        L.s4g_oracle_prefilter.restype = C.c_uint64
        L.s4g_oracle_lis.argtypes = [i32p, C.c_int32]
        L.s4g_oracle_lis.restype = C.c_int32
        L.s4g_oracle_sw_score.argtypes = [u8p, C.c_int32, u8p, C.c_int32, i32p, C.c_int32, C.c_int32]
        L.s4g_oracle_sw_score.restype = C.c_int32
        L.s4g_oracle_evalue.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32]
        L.s4g_oracle_evalue.restype = C.c_double
        L.s4g_oracle_select.argtypes = [f64p, i32p, C.POINTER(C.c_char_p), C.c_int32, C.c_double, C.c_int32, i32p]
        L.s4g_oracle_select.restype = C.c_int32
        L.s4g_oracle_align.argtypes = [u8p, C.c_int32, u8p, C.c_int32, i32p, C.c_int32, C.c_int32, C.c_int32, i32p, u8p, C.c_int32]
        L.s4g_oracle_align.restype = C.c_int32
        L.s4g_oracle_ssw_endpoints.argtypes = [u8p, C.c_int32, u8p, C.c_int32, i32p, C.c_int32, C.c_int32, i32p, i32p]
        L.s4g_oracle_ssw_banded.argtypes = [u8p, C.c_int32, u8p, C.c_int32, i32p, C.c_int32, C.c_int32, C.c_int32, u8p, C.c_int32]
        L.s4g_oracle_ssw_banded.restype = C.c_int32
        L.s4g_oracle_alignment_string.argtypes = [u8p, C.c_int32, C.c_int32, C.c_int32, u8p, C.c_int32, C.c_char_p]
        L.s4g_oracle_alignments_select.argtypes = [C.POINTER(C.c_char_p), C.c_int32, C.c_int32, C.c_float]
        L.s4g_oracle_alignments_select.restype = C.c_int32
        L.s4g_oracle_alignment_stats.argtypes = [u8p, u8p, C.c_int32, C.c_int32, u8p, C.c_int32, i32p]
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def blosum62():
    m = np.zeros(676, dtype=np.int32)
    lib().s4g_oracle_blosum62(_p(m, C.c_int32))
    return m


def encode(s):
    b = s.encode() if isinstance(s, str) else bytes(s)
    out = np.zeros(max(len(b), 1), dtype=np.uint8)
    n = lib().s4g_oracle_encode(b, len(b), _p(out, C.c_uint8))
    return out[:n].copy()


def prefilter(db_codes, db_off, q_codes, q_off, k=5, max_candidates=5000, dense=False):
    """-> (cells, [ids per query], [scores per query], dense matrix or None)"""
    n_db, nq = len(db_off) - 1, len(q_off) - 1
    db_codes = np.ascontiguousarray(db_codes, dtype=np.uint8); q_codes = np.ascontiguousarray(q_codes, dtype=np.uint8)
    db_off = np.ascontiguousarray(db_off, dtype=np.int64); q_off = np.ascontiguousarray(q_off, dtype=np.int64)
    ids = np.zeros((nq, max_candidates), dtype=np.uint32)
    sc = np.zeros((nq, max_candidates), dtype=np.float32)
    cnt = np.zeros(nq, dtype=np.uint32)
    allsc = np.zeros((nq, n_db), dtype=np.float32) if dense else None
    cells = lib().s4g_oracle_prefilter(_p(db_codes, C.c_uint8), _p(db_off, C.c_int64), n_db, _p(q_codes, C.c_uint8),
                                       _p(q_off, C.c_int64), nq, k, max_candidates, _p(ids, C.c_uint32),
                                       _p(sc, C.c_float), _p(cnt, C.c_uint32),
                                       _p(allsc, C.c_float) if dense else None)
    return int(cells), [ids[q, :cnt[q]].copy() for q in range(nq)], [sc[q, :cnt[q]].copy() for q in range(nq)], allsc


def lis(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return lib().s4g_oracle_lis(_p(a, C.c_int32), len(a))


def sw_score(q, t, mat=None, go=10, ge=1):
    mat = blosum62() if mat is None else np.ascontiguousarray(mat, dtype=np.int32)
    q = np.ascontiguousarray(q, dtype=np.uint8); t = np.ascontiguousarray(t, dtype=np.uint8)
    return lib().s4g_oracle_sw_score(_p(q, C.c_uint8), len(q), _p(t, C.c_uint8), len(t), _p(mat, C.c_int32), go, ge)


def evalue(score, qlen, tlen, db_len, go=10, ge=1):
    return lib().s4g_oracle_evalue(int(score), int(qlen), int(tlen), int(db_len), go, ge)


def select(values, scores, names, threshold=1e-4, max_alignments=400):
    values = np.ascontiguousarray(values, dtype=np.float64); scores = np.ascontiguousarray(scores, dtype=np.int32)
    n = len(values)
    arr = (C.c_char_p * max(n, 1))(*[s.encode() if isinstance(s, str) else s for s in names])
    out = np.zeros(max(n, 1), dtype=np.int32)
    k = lib().s4g_oracle_select(_p(values, C.c_double), _p(scores, C.c_int32), arr, n, threshold, max_alignments, _p(out, C.c_int32))
    return out[:k].copy()


def align(q, t, score, mat=None, go=10, ge=1):
    """-> (coords[4], path bytes) ; raises on oracle error codes"""
    mat = blosum62() if mat is None else np.ascontiguousarray(mat, dtype=np.int32)
    q = np.ascontiguousarray(q, dtype=np.uint8); t = np.ascontiguousarray(t, dtype=np.uint8)
    coords = np.zeros(4, dtype=np.int32)
    path = np.zeros(len(q) + len(t) + 2, dtype=np.uint8)
    n = lib().s4g_oracle_align(_p(q, C.c_uint8), len(q), _p(t, C.c_uint8), len(t), _p(mat, C.c_int32), go, ge, int(score),
                               _p(coords, C.c_int32), _p(path, C.c_uint8), len(path))
    if n < 0:
        raise RuntimeError("oracle align error %d" % n)
    return coords, path[:n].copy()


def ssw_endpoints(q, t, mat=None, go=10, ge=1):
    mat = blosum62() if mat is None else np.ascontiguousarray(mat, dtype=np.int32)
    q = np.ascontiguousarray(q, dtype=np.uint8); t = np.ascontiguousarray(t, dtype=np.uint8)
    coords = np.zeros(4, dtype=np.int32); s = C.c_int32(0)
    lib().s4g_oracle_ssw_endpoints(_p(q, C.c_uint8), len(q), _p(t, C.c_uint8), len(t), _p(mat, C.c_int32), go, ge, C.byref(s), _p(coords, C.c_int32))
    return s.value, coords


def alignment_string(t, qlen, coords, path):
    """alignmentsExtract: the hit as bytes over the query positions"""
    t = np.ascontiguousarray(t, dtype=np.uint8); path = np.ascontiguousarray(path, dtype=np.uint8)
    out = C.create_string_buffer(int(qlen) + 1)
    lib().s4g_oracle_alignment_string(_p(t, C.c_uint8), int(qlen), int(coords[0]), int(coords[2]), _p(path, C.c_uint8), len(path), out)
    return out.raw[:int(qlen)]


def alignment_stats(q, t, coords, path):
    """-> [identities, mismatches, gap openings, length] of the --sub-results table"""
    q = np.ascontiguousarray(q, dtype=np.uint8); t = np.ascontiguousarray(t, dtype=np.uint8); path = np.ascontiguousarray(path, dtype=np.uint8)
    out = np.zeros(4, dtype=np.int32)
    lib().s4g_oracle_alignment_stats(_p(q, C.c_uint8), _p(t, C.c_uint8), int(coords[0]), int(coords[2]), _p(path, C.c_uint8), len(path), _p(out, C.c_int32))
    return out


def alignments_select(strings, qlen, threshold=2.75):
    """alignmentsSelect: how many of the strings (in order) are kept"""
    arr = (C.c_char_p * max(len(strings), 1))(*strings)
    return lib().s4g_oracle_alignments_select(arr, len(strings), int(qlen), threshold)


# ---- the real reference, when oracle/_ref was built (this container, or travelled to the GPU box) ----
REF_DUMP = os.path.join(_HERE, "_ref", "ref_dump")
REF_SIFT4G = os.path.join(_HERE, "_ref", "sift4g_ref")


def have_ref():
    return os.path.exists(REF_DUMP) and os.access(REF_DUMP, os.X_OK)
