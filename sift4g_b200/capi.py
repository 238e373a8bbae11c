"""ctypes binding of the C ABI (include/sift4g_b200.h -> sift4g_b200/libsift4g_b200.so).

This is the only way Python reaches the CUDA path: there is no Python/CPU fallback.  Loading fails
loudly when the shared library is missing or was not built.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsift4g_b200.so")

S4G_HOST, S4G_DEVICE = 0, 1

_u8p, _i32p, _i64p, _u32p, _f32p, _f64p = (C.POINTER(t) for t in (C.c_uint8, C.c_int32, C.c_int64, C.c_uint32, C.c_float, C.c_double))
_vp = C.c_void_p

# name -> (restype, argtypes); must list every symbol include/sift4g_b200.h declares
SIGNATURES = {
    "s4g_version": (C.c_int, []),
    "s4g_device_count": (C.c_int, []),
    "s4g_init": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "s4g_shutdown": (None, [_vp]),
    "s4g_last_error": (C.c_char_p, [_vp]),
    "s4g_set_stream": (C.c_int, [_vp, _vp]),
    "s4g_sync": (C.c_int, [_vp]),
    "s4g_launch_count": (C.c_int64, [_vp]),
    "s4g_launch_count_reset": (None, [_vp]),
    "s4g_db_create": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_uint32, C.c_int, C.POINTER(_vp)]),
    "s4g_db_open_fasta": (C.c_int, [_vp, C.c_char_p, C.c_int, C.c_int, C.POINTER(_vp)]),
    "s4g_db_pack_fasta": (C.c_int, [C.c_char_p, C.c_char_p]),
    "s4g_db_file_info": (C.c_int, [C.c_char_p, _i64p, C.POINTER(C.c_uint64)]),
    "s4g_db_open_packed": (C.c_int, [_vp, C.c_char_p, C.c_int, C.c_int, C.POINTER(_vp)]),
    "s4g_db_open": (C.c_int, [_vp, C.c_char_p, C.c_int, C.c_int, C.POINTER(_vp)]),
    "s4g_db_open_sharded": (C.c_int, [_vp, C.c_int, C.c_char_p, C.c_int, _vp]),
    "s4g_db_total_seqs": (C.c_int64, [_vp]),
    "s4g_db_total_residues": (C.c_uint64, [_vp]),
    "s4g_db_close": (None, [_vp]),
    "s4g_db_num_seqs": (C.c_int64, [_vp]),
    "s4g_db_num_residues": (C.c_uint64, [_vp]),
    "s4g_db_id_base": (C.c_uint32, [_vp]),
    "s4g_db_host_offsets": (_i64p, [_vp]),
    "s4g_db_host_codes": (_u8p, [_vp]),
    "s4g_db_name": (C.c_char_p, [_vp, C.c_int64]),
    "s4g_stripe_granularity": (C.c_uint64, [_vp]),
    "s4g_stripe_create": (C.c_int, [_vp, C.c_uint64, C.POINTER(_vp)]),
    "s4g_stripe_bytes": (C.c_uint64, [_vp]),
    "s4g_stripe_export_fd": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "s4g_stripe_free": (None, [_vp]),
    "s4g_view_open": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.POINTER(_vp)]),
    "s4g_view_ptr": (_vp, [_vp]),
    "s4g_view_bytes": (C.c_uint64, [_vp]),
    "s4g_view_write": (C.c_int, [_vp, C.c_uint64, _vp, C.c_uint64]),
    "s4g_view_fill": (C.c_int, [_vp, C.c_uint64, C.c_int, C.c_uint64]),
    "s4g_view_close": (None, [_vp]),
    "s4g_db_create_view": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int, C.POINTER(_vp)]),
    "s4g_queries_create": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int, C.POINTER(_vp)]),
    "s4g_queries_free": (None, [_vp]),
    "s4g_prefilter": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int]),
    "s4g_merge_candidates": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "s4g_topn_cutoff": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "s4g_cutoff_counts": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "s4g_sw_score": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int64, _vp, C.c_int, C.c_int, _vp, C.c_int]),
    "s4g_sw_align": (C.c_int, [_vp, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int64, _vp, C.c_int]),
    "s4g_merge_hits": (C.c_int, [_vp, C.c_int, C.c_int32, C.c_int, _vp, C.c_int64, _vp, C.c_uint32, C.c_uint32, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "s4g_merge_candidates_host": (C.c_int, [C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp]),
    "s4g_select_hits": (C.c_int, [_vp, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "s4g_evalue_screen": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int64, _vp, C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.c_double, _vp, _vp, _vp, _vp, _vp]),
    "s4g_score_screen": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int, _vp, C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.c_double, _vp]),
    "s4g_search": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "s4g_alignment_strings": (C.c_int, [_vp, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "s4g_alignments_select": (C.c_int, [_vp, C.c_int32, _vp, _vp, _vp, C.c_float, _vp]),
    "s4g_alignment_stats": (C.c_int, [_vp, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "s4g_write_blast_tab": (C.c_int, [C.c_char_p, C.c_int, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "s4g_measure_dpx_peak": (C.c_int, [_vp, C.c_int, _f64p]),
    "s4g_measure_gather_peak": (C.c_int, [_vp, C.c_int, _f64p]),
    "s4g_last_sw_kernel_ms": (C.c_int, [_vp, _f32p]),
    "s4g_last_align_profile": (C.c_int, [_vp, _f32p, C.POINTER(C.c_uint64)]),
}



class Survivors(C.Structure):
    """s4g_survivors (include/sift4g_b200.h)"""
    _fields_ = [("n", C.c_int64), ("query", _vp), ("id", _vp), ("score", _vp), ("tlen", _vp), ("n_pairs", C.c_int64),
                ("sw_cells", C.c_uint64), ("sw_kernel_ms", C.c_float)]


class SearchParams(C.Structure):
    """s4g_search_params"""
    _fields_ = [("kmer_length", C.c_int), ("max_candidates", C.c_int), ("matrix", _vp), ("matrix_name", C.c_char_p),
                ("gap_open", C.c_int), ("gap_extend", C.c_int), ("max_evalue", C.c_double), ("max_alignments", C.c_int),
                ("n_threads", C.c_int), ("want_candidates", C.c_int), ("want_alignments", C.c_int), ("device_results", C.c_int)]


class SearchResult(C.Structure):
    """s4g_search_result"""
    _fields_ = [("n_queries", C.c_int32), ("n_pairs", C.c_int64), ("n_survivors", C.c_int64), ("sw_cells", C.c_uint64),
                ("db_residues", C.c_uint64), ("cand_ids", _vp), ("cand_offsets", _vp), ("n_hits", C.c_int64), ("hit_query", _vp),
                ("hit_target", _vp), ("hit_score", _vp), ("hit_evalue", _vp), ("hit_offsets", _vp), ("coords", _vp), ("paths", _vp),
                ("path_offsets", _vp), ("ms_prefilter", C.c_float), ("ms_score", C.c_float), ("ms_select", C.c_float),
                ("ms_align", C.c_float), ("sw_kernel_ms", C.c_float), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]


_LIB = None


def load():
    """Load the shared library (raises if it has not been built: no fallback)."""
    global _LIB
    if _LIB is None:
        path = os.environ.get("S4G_LIB_PATH", LIB_PATH)       # experiments: another build of the same library
        if not os.path.exists(path):
            raise RuntimeError("sift4g_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)" % path)
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


class S4GError(RuntimeError):
    pass


def pack_fasta(fasta_path, out_path):
    """FASTA -> packed database file (host only: needs neither a GPU nor a context)."""
    lib = load()
    rc = lib.s4g_db_pack_fasta(fasta_path.encode(), out_path.encode())
    if rc != 0:
        raise S4GError("s4g_db_pack_fasta failed (%d): %s" % (rc, lib.s4g_last_error(None).decode()))


def packed_info(path):
    lib = load()
    n, r = C.c_int64(0), C.c_uint64(0)
    rc = lib.s4g_db_file_info(path.encode(), C.byref(n), C.byref(r))
    if rc != 0:
        raise S4GError("s4g_db_file_info failed (%d): %s" % (rc, lib.s4g_last_error(None).decode()))
    return n.value, r.value


def read_packed(path):
    """numpy reader of the packed layout documented in include/sift4g_b200.h (tests; independent of the library)."""
    raw = np.fromfile(path, dtype=np.uint8)
    assert raw[:8].tobytes() == b"S4GDB\0\0\1", "bad magic"
    n, n_res, names_bytes, codes_pos = (int(x) for x in raw[8:40].view(np.uint64))
    p = 64
    off = raw[p:p + 8 * (n + 1)].view(np.int64).copy(); p += 8 * (n + 1)
    name_off = raw[p:p + 8 * (n + 1)].view(np.int64).copy(); p += 8 * (n + 1)
    blob = raw[p:p + names_bytes].tobytes()
    names = [blob[name_off[i]:name_off[i + 1] - 1].decode() for i in range(n)]
    codes = raw[codes_pos:codes_pos + n_res].copy()
    return names, off, codes


def _ptr(a):
    """host numpy array / torch tensor / int / None -> void pointer value"""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    return a.data_ptr()      # torch tensor


class Context:
    def __init__(self, device=0):
        self.lib = load()
        h = _vp()
        rc = self.lib.s4g_init(device, C.byref(h))
        if rc != 0:
            raise S4GError("s4g_init failed (%d): %s" % (rc, self.lib.s4g_last_error(None).decode()))
        self.h = h
        self.device = device

    def check(self, rc):
        if rc != 0:
            raise S4GError("sift4g_b200 error %d: %s" % (rc, self.lib.s4g_last_error(self.h).decode()))

    def set_stream(self, stream_ptr):
        self.check(self.lib.s4g_set_stream(self.h, stream_ptr))

    def sync(self):
        self.check(self.lib.s4g_sync(self.h))

    def launch_count(self):
        return int(self.lib.s4g_launch_count(self.h))

    def reset_launch_count(self):
        self.lib.s4g_launch_count_reset(self.h)

    def dpx_peak(self, millis=50):
        v = C.c_double(0)
        self.check(self.lib.s4g_measure_dpx_peak(self.h, millis, C.byref(v)))
        return v.value

    def gather_peak(self, millis=50):
        v = C.c_double(0)
        self.check(self.lib.s4g_measure_gather_peak(self.h, millis, C.byref(v)))
        return v.value

    def last_sw_kernel_ms(self):
        v = C.c_float(0)
        self.check(self.lib.s4g_last_sw_kernel_ms(self.h, C.byref(v)))
        return v.value

    def last_align_profile(self):
        """-> ({"ends": ms, "begins": ms, "paths": ms}, {"ends": cells, ...}) of the last traceback on this context"""
        ms = (C.c_float * 3)()
        cells = (C.c_uint64 * 3)()
        self.check(self.lib.s4g_last_align_profile(self.h, ms, cells))
        names = ("ends", "begins", "paths")
        return {n: float(ms[i]) for i, n in enumerate(names)}, {n: int(cells[i]) for i, n in enumerate(names)}

    def close(self):
        if self.h:
            self.lib.s4g_shutdown(self.h)
            self.h = None

    # ---- objects -----------------------------------------------------------------------------
    def database(self, codes, offsets, id_base=0, where=S4G_HOST):
        return Database(self, codes, offsets, id_base, where)

    def database_from_fasta(self, path, shard=0, n_shards=1):
        return self._open(self.lib.s4g_db_open_fasta, path, shard, n_shards)

    def database_from_packed(self, path, shard=0, n_shards=1):
        return self._open(self.lib.s4g_db_open_packed, path, shard, n_shards)

    def database_from_file(self, path, shard=0, n_shards=1):
        """FASTA or packed (.s4gdb) file, told apart by the magic."""
        return self._open(self.lib.s4g_db_open, path, shard, n_shards)

    def _open(self, fn, path, shard, n_shards):
        db = Database.__new__(Database)
        db.ctx = self
        h = _vp()
        self.check(fn(self.h, path.encode(), shard, n_shards, C.byref(h)))
        db.h = h
        return db

    def queries(self, codes, offsets, where=S4G_HOST):
        return Queries(self, codes, offsets, where)


class Stripe:
    """One GPU's resident part of an NVLink-striped database (s4g_stripe_*)."""

    def __init__(self, ctx, nbytes):
        self.ctx = ctx
        h = _vp()
        ctx.check(ctx.lib.s4g_stripe_create(ctx.h, nbytes, C.byref(h)))
        self.h = h
        self.nbytes = int(nbytes)

    def export_fd(self):
        fd = C.c_int(-1)
        self.ctx.check(self.ctx.lib.s4g_stripe_export_fd(self.h, C.byref(fd)))
        return fd.value

    def free(self):
        if self.h:
            self.ctx.lib.s4g_stripe_free(self.h)
            self.h = None


class View:
    """All stripes of a striped database mapped into one virtual range of ctx's device (s4g_view_*).
    stripes[i] is a Stripe of this process or an int file descriptor received from the owning process."""

    def __init__(self, ctx, stripes, nbytes):
        self.ctx = ctx
        n = len(stripes)
        loc = (_vp * n)(*[s.h if isinstance(s, Stripe) else None for s in stripes])
        fds = (C.c_int * n)(*[-1 if isinstance(s, Stripe) else int(s) for s in stripes])
        by = (C.c_uint64 * n)(*[int(b) for b in nbytes])
        h = _vp()
        ctx.check(ctx.lib.s4g_view_open(ctx.h, n, loc, fds, by, C.byref(h)))
        self.h = h
        self.nbytes = int(ctx.lib.s4g_view_bytes(h))

    @property
    def ptr(self):
        return int(self.ctx.lib.s4g_view_ptr(self.h))

    def write(self, at, src, nbytes):
        """copy `nbytes` of device memory `src` (tensor / pointer) to byte `at` of the view"""
        self.ctx.check(self.ctx.lib.s4g_view_write(self.h, int(at), _ptr(src), int(nbytes)))

    def fill(self, at, value, nbytes):
        self.ctx.check(self.ctx.lib.s4g_view_fill(self.h, int(at), int(value), int(nbytes)))

    def database(self, offsets, where=S4G_HOST):
        """Database over the view's bytes (global ids, codes not copied); close it before the view."""
        db = Database.__new__(Database)
        db.ctx = self.ctx
        if where == S4G_HOST:
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = (len(offsets) if where == S4G_HOST else offsets.numel()) - 1
        h = _vp()
        self.ctx.check(self.ctx.lib.s4g_db_create_view(self.ctx.h, self.h, _ptr(offsets), n, where, C.byref(h)))
        db.h = h
        return db

    def close(self):
        if self.h:
            self.ctx.lib.s4g_view_close(self.h)
            self.h = None


class Database:
    def __init__(self, ctx, codes, offsets, id_base=0, where=S4G_HOST):
        self.ctx = ctx
        if where == S4G_HOST:
            codes = np.ascontiguousarray(codes, dtype=np.uint8)
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = (len(offsets) if where == S4G_HOST else offsets.numel()) - 1
        h = _vp()
        ctx.check(ctx.lib.s4g_db_create(ctx.h, _ptr(codes), _ptr(offsets), n, id_base, where, C.byref(h)))
        self.h = h

    @property
    def n_seqs(self):
        return int(self.ctx.lib.s4g_db_num_seqs(self.h))

    @property
    def n_residues(self):
        return int(self.ctx.lib.s4g_db_num_residues(self.h))

    @property
    def id_base(self):
        return int(self.ctx.lib.s4g_db_id_base(self.h))

    @property
    def total_seqs(self):
        return int(self.ctx.lib.s4g_db_total_seqs(self.h))

    @property
    def total_residues(self):
        return int(self.ctx.lib.s4g_db_total_residues(self.h))

    def host_offsets(self):
        p = self.ctx.lib.s4g_db_host_offsets(self.h)
        return np.ctypeslib.as_array(p, shape=(self.n_seqs + 1,)).copy()

    def host_codes(self):
        p = self.ctx.lib.s4g_db_host_codes(self.h)
        if not p:
            return None
        return np.ctypeslib.as_array(p, shape=(self.n_residues,)).copy()

    def name(self, i):
        s = self.ctx.lib.s4g_db_name(self.h, i)
        return s.decode() if s is not None else None

    def close(self):
        if self.h:
            self.ctx.lib.s4g_db_close(self.h)
            self.h = None


class Queries:
    def __init__(self, ctx, codes, offsets, where=S4G_HOST):
        self.ctx = ctx
        if where == S4G_HOST:
            codes = np.ascontiguousarray(codes, dtype=np.uint8)
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.n = (len(offsets) if where == S4G_HOST else offsets.numel()) - 1
        self.lens = np.diff(offsets).astype(np.int64) if where == S4G_HOST else None
        h = _vp()
        ctx.check(ctx.lib.s4g_queries_create(ctx.h, _ptr(codes), _ptr(offsets), self.n, where, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.ctx.lib.s4g_queries_free(self.h)
            self.h = None


# ---- stage calls (host numpy in / out: the e2e path; device tensors: pass where=S4G_DEVICE) ----------

def prefilter(ctx, db, q, k=5, max_candidates=5000, sorted_by_id=True, out=None, where=S4G_HOST):
    """-> (ids [nq, max_candidates] uint32, scores float32, counts uint32)"""
    if out is None:
        ids = np.zeros((q.n, max_candidates), dtype=np.uint32)
        sc = np.zeros((q.n, max_candidates), dtype=np.float32)
        cnt = np.zeros(q.n, dtype=np.uint32)
    else:
        ids, sc, cnt = out
    ctx.check(ctx.lib.s4g_prefilter(ctx.h, db.h, q.h, k, max_candidates, 1 if sorted_by_id else 0, _ptr(ids), _ptr(sc), _ptr(cnt), where))
    return ids, sc, cnt


def sw_score(ctx, db, q, cand_ids, cand_offsets, matrix, gap_open=10, gap_extend=1, out=None, where=S4G_HOST):
    matrix = np.ascontiguousarray(matrix, dtype=np.int32)
    if where == S4G_HOST:
        cand_ids = np.ascontiguousarray(cand_ids, dtype=np.uint32)
        cand_offsets = np.ascontiguousarray(cand_offsets, dtype=np.int64)
        n = len(cand_ids)
        if out is None:
            out = np.zeros(n, dtype=np.int32)
    else:
        n = cand_ids.numel()
    ctx.check(ctx.lib.s4g_sw_score(ctx.h, db.h, q.h, _ptr(cand_ids), _ptr(cand_offsets), n, _ptr(matrix), gap_open, gap_extend, _ptr(out), where))
    return out


def sw_align(ctx, db, q, pair_q, pair_t, pair_score, matrix, gap_open=10, gap_extend=1, path_capacity=None, q_lens=None, t_lens=None):
    """host path: -> (coords [n,4] int32, list of path arrays)"""
    matrix = np.ascontiguousarray(matrix, dtype=np.int32)
    pair_q = np.ascontiguousarray(pair_q, dtype=np.uint32)
    pair_t = np.ascontiguousarray(pair_t, dtype=np.uint32)
    pair_score = np.ascontiguousarray(pair_score, dtype=np.int32)
    n = len(pair_q)
    if path_capacity is None:
        path_capacity = int(np.sum(q_lens[pair_q]) + np.sum(t_lens[pair_t - db.id_base])) + 16
    coords = np.zeros((n, 4), dtype=np.int32)
    paths = np.zeros(path_capacity, dtype=np.uint8)
    off = np.zeros(n + 1, dtype=np.int64)
    ctx.check(ctx.lib.s4g_sw_align(ctx.h, db.h, q.h, n, _ptr(pair_q), _ptr(pair_t), _ptr(pair_score), _ptr(matrix), gap_open, gap_extend,
                                   _ptr(coords), _ptr(paths), path_capacity, _ptr(off), S4G_HOST))
    return coords, [paths[off[i]:off[i + 1]].copy() for i in range(n)]


def select_hits(ctx, query_lens, cand_ids, cand_offsets, cand_scores, cand_lens, db_residues, gap_open=10, gap_extend=1,
                max_evalue=1e-4, max_alignments=400, names=None, n_threads=0, matrix_name=b"BLOSUM_62"):
    """host: -> (pair_q, pair_t, pair_score, evalues, offsets[nq+1])"""
    query_lens = np.ascontiguousarray(query_lens, dtype=np.int32)
    cand_ids = np.ascontiguousarray(cand_ids, dtype=np.uint32)
    cand_offsets = np.ascontiguousarray(cand_offsets, dtype=np.int64)
    cand_scores = np.ascontiguousarray(cand_scores, dtype=np.int32)
    cand_lens = np.ascontiguousarray(cand_lens, dtype=np.int32)
    nq = len(query_lens)
    cap = max(nq * max_alignments, 1)
    oq = np.zeros(cap, dtype=np.uint32); ot = np.zeros(cap, dtype=np.uint32)
    osc = np.zeros(cap, dtype=np.int32); oe = np.zeros(cap, dtype=np.float64)
    off = np.zeros(nq + 1, dtype=np.int64)
    name_arr = None
    if names is not None:
        name_arr = (C.c_char_p * max(len(names), 1))(*[n.encode() if isinstance(n, str) else n for n in names])
    lib = ctx.lib if ctx is not None else load()        # pure host code: usable without a context (CPU tests)
    rc = lib.s4g_select_hits(ctx.h if ctx is not None else None, nq, _ptr(query_lens), _ptr(cand_ids), _ptr(cand_offsets), _ptr(cand_scores),
                             _ptr(cand_lens), C.cast(name_arr, _vp) if name_arr is not None else None, matrix_name, int(db_residues), gap_open, gap_extend,
                             max_evalue, max_alignments, n_threads, _ptr(oq), _ptr(ot), _ptr(osc), _ptr(oe), _ptr(off))
    if rc != 0:
        raise S4GError("s4g_select_hits failed (%d): %s" % (rc, lib.s4g_last_error(ctx.h if ctx is not None else None).decode()))
    n = int(off[-1])
    return oq[:n], ot[:n], osc[:n], oe[:n], off


def merge_hits(ctx, gathered, counts, nq, max_alignments, own_lo=0, own_hi=0xffffffff, n_threads=0):
    """host: gathered float64 [n_ranks][stride][3] rows {E, score, id} (rank r's hits grouped by query), counts int64
    [n_ranks][nq] -> (pair_q, pair_t, pair_score, evalues, offsets[nq+1]) of the global top hits whose target lies in
    [own_lo, own_hi)"""
    gathered = np.ascontiguousarray(gathered, dtype=np.float64)
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    n_ranks, stride = gathered.shape[0], gathered.shape[1]
    cap = max(nq * max_alignments, 1)
    oq = np.zeros(cap, dtype=np.uint32); ot = np.zeros(cap, dtype=np.uint32)
    osc = np.zeros(cap, dtype=np.int32); oe = np.zeros(cap, dtype=np.float64)
    off = np.zeros(nq + 1, dtype=np.int64)
    lib = ctx.lib if ctx is not None else load()        # pure host code: usable without a context (CPU tests)
    rc = lib.s4g_merge_hits(ctx.h if ctx is not None else None, n_ranks, nq, max_alignments, _ptr(gathered), stride, _ptr(counts),
                            int(own_lo), int(own_hi), n_threads, _ptr(oq), _ptr(ot), _ptr(osc), _ptr(oe), _ptr(off))
    if rc != 0:
        raise S4GError("s4g_merge_hits failed (%d)" % rc)
    n = int(off[-1])
    return oq[:n], ot[:n], osc[:n], oe[:n], off


def merge_candidates_host(ids, scores, counts, max_candidates, n_threads=0):
    """Host merge of per-shard best-first prefilter rows (lists of (nq x N) uint32 / float32 arrays and (nq,) uint32 counts):
    global top max_candidates per query by (score desc, id asc), returned in ascending id -- the CLI's multi-GPU path.
    Pure host code: no context."""
    lib = load()
    n = len(ids)
    nq = int(counts[0].shape[0])
    ids = [np.ascontiguousarray(a, dtype=np.uint32) for a in ids]
    scores = [np.ascontiguousarray(a, dtype=np.float32) for a in scores]
    counts = [np.ascontiguousarray(a, dtype=np.uint32) for a in counts]
    arr = lambda xs: (C.c_void_p * n)(*[x.ctypes.data for x in xs])
    out = np.zeros((nq, max_candidates), dtype=np.uint32)
    out_cnt = np.zeros(nq, dtype=np.uint32)
    rc = lib.s4g_merge_candidates_host(n, nq, max_candidates, arr(ids), arr(scores), arr(counts), n_threads, _ptr(out), _ptr(out_cnt))
    if rc != 0:
        raise S4GError("s4g_merge_candidates_host failed (%d)" % rc)
    return out, out_cnt


def _view(ptr, n, dtype):
    """numpy view of `n` items of library-owned host memory (valid until the next call on the context)"""
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (int(n) * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=int(n))


def score_screen(ctx, db, q, cand_ids, cand_offsets, matrix, gap_open=10, gap_extend=1, max_evalue=1e-4, db_residues=0,
                 matrix_name=b"BLOSUM_62", where=S4G_HOST):
    """s4g_score_screen -> (query, id, score, tlen) views of the survivors, sw_cells"""
    matrix = np.ascontiguousarray(matrix, dtype=np.int32)
    if where == S4G_HOST:
        cand_ids = np.ascontiguousarray(cand_ids, dtype=np.uint32)
        cand_offsets = np.ascontiguousarray(cand_offsets, dtype=np.int64)
        n = len(cand_ids)
    else:
        n = cand_ids.numel()
    sv = Survivors()
    ctx.check(ctx.lib.s4g_score_screen(ctx.h, db.h, q.h, _ptr(cand_ids), _ptr(cand_offsets), n, where, _ptr(matrix), matrix_name,
                                       int(db_residues), gap_open, gap_extend, max_evalue, C.addressof(sv)))
    return _view(sv.query, sv.n, np.uint32), _view(sv.id, sv.n, np.uint32), _view(sv.score, sv.n, np.int32), _view(sv.tlen, sv.n, np.int32), int(sv.sw_cells)


class SearchOutput:
    """numpy views of an s4g_search_result (library-owned pinned memory: copy what must outlive the next call)"""

    def __init__(self, r, device_results=False):
        nq = r.n_queries
        self.device_results = device_results
        self.n_queries, self.n_pairs, self.n_survivors, self.n_hits = nq, int(r.n_pairs), int(r.n_survivors), int(r.n_hits)
        self.sw_cells, self.db_residues = int(r.sw_cells), int(r.db_residues)
        self.cand_off = _view(r.cand_offsets, nq + 1, np.int64) if r.cand_offsets else None
        self.cand_ids = _view(r.cand_ids, self.n_pairs, np.uint32) if r.cand_ids else None
        self.pair_q = _view(r.hit_query, self.n_hits, np.uint32)
        self.pair_t = _view(r.hit_target, self.n_hits, np.uint32)
        self.pair_score = _view(r.hit_score, self.n_hits, np.int32)
        self.evalue = _view(r.hit_evalue, self.n_hits, np.float64)
        self.hit_off = _view(r.hit_offsets, nq + 1, np.int64)
        if device_results:          # device pointers (int), valid until the next call on the context
            self.coords, self.path_off, self.paths = r.coords, r.path_offsets, r.paths
        else:
            self.coords = _view(r.coords, 4 * self.n_hits, np.int32).reshape(-1, 4) if r.coords else None
            self.path_off = _view(r.path_offsets, self.n_hits + 1, np.int64) if r.path_offsets else None
            self.paths = _view(r.paths, int(self.path_off[-1]) if self.path_off is not None and self.n_hits else 0, np.uint8) if r.paths else None
        self.stage_ms = {"prefilter": r.ms_prefilter, "score": r.ms_score, "select": r.ms_select, "align": r.ms_align}
        self.sw_kernel_ms = r.sw_kernel_ms
        self.h2d_bytes, self.d2h_bytes = int(r.h2d_bytes), int(r.d2h_bytes)


def search(ctx, db, q, matrix, k=5, max_candidates=5000, gap_open=10, gap_extend=1, max_evalue=1e-4, max_alignments=400, n_threads=0,
           want_candidates=True, want_alignments=True, matrix_name=b"BLOSUM_62", device_results=False):
    """s4g_search: the whole hot path for one resident shard (or striped view), host buffers out; device_results=True leaves
    cells and paths in HBM (the hit lists always come to the host: the selection runs there)."""
    matrix = np.ascontiguousarray(matrix, dtype=np.int32)
    prm = SearchParams(k, max_candidates, matrix.ctypes.data, matrix_name, gap_open, gap_extend, max_evalue, max_alignments, n_threads,
                       1 if want_candidates else 0, 1 if want_alignments else 0, 1 if device_results else 0)
    res = SearchResult()
    ctx.check(ctx.lib.s4g_search(ctx.h, db.h, q.h, C.addressof(prm), C.addressof(res)))
    return SearchOutput(res, device_results)


def alignment_strings(ctx, db, q, pair_q, pair_t, coords, paths, path_off):
    """s4g_alignment_strings -> (strings uint8 back to back, offsets[n_hits + 1]); needs the query lengths via q"""
    pair_q = np.ascontiguousarray(pair_q, dtype=np.uint32); pair_t = np.ascontiguousarray(pair_t, dtype=np.uint32)
    coords = np.ascontiguousarray(coords, dtype=np.int32); paths = np.ascontiguousarray(paths, dtype=np.uint8)
    path_off = np.ascontiguousarray(path_off, dtype=np.int64)
    n = len(pair_q)
    qlens = q.lens if hasattr(q, "lens") else None
    cap = int(qlens[pair_q].sum()) if qlens is not None else 0
    out = np.zeros(max(cap, 1), dtype=np.uint8)
    off = np.zeros(n + 1, dtype=np.int64)
    ctx.check(ctx.lib.s4g_alignment_strings(ctx.h, db.h, q.h, n, _ptr(pair_q), _ptr(pair_t), _ptr(coords), _ptr(paths), _ptr(path_off), _ptr(out), _ptr(off)))
    return out[:int(off[-1])], off


def alignments_select(ctx, query_lens, hit_off, strings, threshold=2.75):
    query_lens = np.ascontiguousarray(query_lens, dtype=np.int32); hit_off = np.ascontiguousarray(hit_off, dtype=np.int64)
    strings = np.ascontiguousarray(strings, dtype=np.uint8)
    out = np.zeros(len(query_lens), dtype=np.int32)
    ctx.check(ctx.lib.s4g_alignments_select(ctx.h, len(query_lens), _ptr(query_lens), _ptr(hit_off), _ptr(strings), threshold, _ptr(out)))
    return out


def alignment_stats(ctx, db, q, pair_q, pair_t, coords, paths, path_off):
    """s4g_alignment_stats -> int32 [n_hits, 4]: identities, mismatches, gap openings, alignment length"""
    pair_q = np.ascontiguousarray(pair_q, dtype=np.uint32); pair_t = np.ascontiguousarray(pair_t, dtype=np.uint32)
    coords = np.ascontiguousarray(coords, dtype=np.int32); paths = np.ascontiguousarray(paths, dtype=np.uint8)
    path_off = np.ascontiguousarray(path_off, dtype=np.int64)
    out = np.zeros((len(pair_q), 4), dtype=np.int32)
    ctx.check(ctx.lib.s4g_alignment_stats(ctx.h, db.h, q.h, len(pair_q), _ptr(pair_q), _ptr(pair_t), _ptr(coords), _ptr(paths), _ptr(path_off), _ptr(out)))
    return out


def write_blast_tab(path, with_header, hit_off, query_names, target_names, stats, coords, evalues, scores):
    lib = load()
    hit_off = np.ascontiguousarray(hit_off, dtype=np.int64)
    qn = (C.c_char_p * max(len(query_names), 1))(*[n.encode() for n in query_names])
    tn = (C.c_char_p * max(len(target_names), 1))(*[n.encode() for n in target_names])
    stats = np.ascontiguousarray(stats, dtype=np.int32); coords = np.ascontiguousarray(coords, dtype=np.int32)
    evalues = np.ascontiguousarray(evalues, dtype=np.float64); scores = np.ascontiguousarray(scores, dtype=np.int32)
    rc = lib.s4g_write_blast_tab(path.encode(), 1 if with_header else 0, len(hit_off) - 1, _ptr(hit_off), C.cast(qn, _vp), C.cast(tn, _vp), _ptr(stats),
                                 _ptr(coords), _ptr(evalues), _ptr(scores))
    if rc != 0:
        raise S4GError("s4g_write_blast_tab failed (%d): %s" % (rc, lib.s4g_last_error(None).decode()))
