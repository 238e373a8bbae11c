"""Packed database file and the multi-threaded FASTA reader behind it, host side only (no GPU):
s4g_db_pack_fasta must keep exactly what the reference's reader keeps (sw/pre_proc.c:437-538, sw/chain.c:59-105).
Checked against the live reference binary (oracle/_ref/ref_dump fasta -> readFastaChains) and against the
reference's own test database."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from sift4g_b200 import capi
from tests.util import GOLDEN


def ref_records(path):
    out = subprocess.run([O.REF_DUMP, "fasta", str(path)], capture_output=True, check=True).stdout.decode()
    lines = out.split("\n")
    n = int(lines[0])
    recs = [tuple(l.split("\t")) for l in lines[1:1 + n]]
    return recs


def packed_records(path, tmp_path, threads=None):
    out = str(tmp_path / "db.s4gdb")
    if threads is not None:
        os.environ["S4G_HOST_THREADS"] = str(threads)
    try:
        capi.pack_fasta(str(path), out)
    finally:
        os.environ.pop("S4G_HOST_THREADS", None)
    names, off, codes = capi.read_packed(out)
    n, res = capi.packed_info(out)
    assert n == len(names) and res == len(codes) == off[-1]
    txt = (codes + 65).astype(np.uint8).tobytes().decode()
    return [(names[i], txt[off[i]:off[i + 1]]) for i in range(len(names))]


QUIRKS = [
    b">  first seq description  \r\nACD-E*fg\nHI 12\n>second\nKLMNPQ",          # no trailing newline: last byte consumed
    b">a\nACD\n>b|x y\t\nEFG\nHIK\n\n\n>c\nLMN\n",                               # blank lines, tab-trimmed name
    b">a\nAC>b\nDEF\n",                                                           # '>' in the middle of a sequence line
    b">a >b\nACDEF\n>c\nGH\n",                                                    # '>' inside a header
    b">a\nacdefGHIK\n>b\nXBZUOJ\n",                                               # case folding, rare letters
    b">a\r\nACD\r\n>b\r\nEFG\r\n",                                                # CRLF
]


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", range(len(QUIRKS)))
def test_reader_quirks_match_reference(tmp_path, case):
    p = tmp_path / "x.fa"
    p.write_bytes(QUIRKS[case])
    assert packed_records(p, tmp_path) == ref_records(p)


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_parallel_reader_matches_reference_on_a_multi_megabyte_file(tmp_path):
    # > 1 MiB so that the file is cut into pieces parsed on several threads; irregular line lengths, some '>'
    # characters inside header lines, CRLF here and there
    rng = np.random.default_rng(5)
    letters = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYXBZUacdefghiklmnpqrstvwy", dtype=np.uint8)
    parts = []
    for i in range(6000):
        L = int(rng.integers(1, 900))
        seq = letters[rng.integers(0, len(letters), L)].tobytes()
        w = int(rng.integers(20, 90))
        eol = b"\r\n" if i % 7 == 0 else b"\n"
        hdr = b">sp|P%05d|NAME_%d some text > more" % (i, i) if i % 5 == 0 else b">S%d" % i
        parts.append(hdr + eol + eol.join(seq[j:j + w] for j in range(0, L, w)) + eol)
    p = tmp_path / "big.fa"
    p.write_bytes(b"".join(parts))
    assert p.stat().st_size > 2 << 20
    ref = ref_records(p)
    assert len(ref) == 6000
    for threads in (1, 3, 16):
        assert packed_records(p, tmp_path, threads) == ref, "threads=%d" % threads


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_file_size_multiple_of_the_reader_buffer(tmp_path):
    # the reference only notices the end of the file on a short fread of its 1 MiB buffer: a file of exactly 2 MiB
    # never closes its last record (sw/pre_proc.c:465-488) -- same here
    body = b">a\n" + b"ACDEFGHIKL\n" * 1000
    rec = b">b\nMNPQ\n"
    pad_len = (2 << 20) - len(body) - len(rec)
    data = body + b">c\n" + b"A" * (pad_len - 4) + b"\n" + rec
    assert len(data) == 2 << 20
    p = tmp_path / "m.fa"
    p.write_bytes(data)
    ref = ref_records(p)
    assert [r[0] for r in ref] == ["a", "c"]
    assert packed_records(p, tmp_path) == ref


def test_pack_reference_test_database(tmp_path):
    # the reference's own sample database (a copy travels as a golden fixture): same records as the plain reader
    from tests.util import read_fasta_codes
    src = os.path.join(GOLDEN, "test_files", "sample_protein_database.fa")
    names, seqs = read_fasta_codes(src)
    got = packed_records(src, tmp_path)
    assert len(got) == len(names)
    for i in range(len(names)):
        assert got[i][0] == names[i]
        assert got[i][1] == "".join(chr(65 + c) for c in seqs[i])


def test_errors(tmp_path):
    with pytest.raises(capi.S4GError):
        capi.pack_fasta(str(tmp_path / "missing.fa"), str(tmp_path / "o"))
    p = tmp_path / "e.fa"
    p.write_bytes(b">a\n>b\nACD\n")           # record without residues: the reference aborts, we return an error
    with pytest.raises(capi.S4GError):
        capi.pack_fasta(str(p), str(tmp_path / "o"))
    q = tmp_path / "notpacked"
    q.write_bytes(b">a\nACD\n")
    with pytest.raises(capi.S4GError):
        capi.packed_info(str(q))


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("n_last", [1, 2, 7, 8, 12, 13, 24, 25, 40, 56, 57, 120])
def test_last_byte_is_a_residue_letter(tmp_path, n_last):
    # no trailing newline and the file ends in a residue letter: the reader consumes that byte as the record terminator
    # (sw/pre_proc.c:488).  The EMIT pass must not store it either: the code vector was sized without it (round-1 advisor
    # finding: a one-byte heap overflow that aborted in free() when the residue count filled its malloc chunk exactly).
    p = tmp_path / "t.fa"
    p.write_bytes(b">a\nAAAAAAAAAAAA\n>b\n" + b"C" * n_last + b"D")
    got = packed_records(p, tmp_path)
    assert got == ref_records(p)
    assert got[-1] == ("b", "C" * n_last)
