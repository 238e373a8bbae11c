"""GPU end to end: the sift4g CLI with the B200 hot path (sift4g_b200/bin/sift4g_b200 = the reference's
main/select_alignments/sift_prediction compiled unchanged around our searchDatabase/alignDatabase) must write
byte-identical .SIFTprediction / alignments / aligned.fasta files (hashes of the reference CPU build, committed
under tests/golden by make_golden.py)."""
import hashlib
import json
import os
import subprocess
import tempfile

import pytest

from tests import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "sift4g_b200", "bin", "sift4g_b200")


def _run(args, env=None):
    out = tempfile.mkdtemp()
    r = subprocess.run([BIN] + args + ["--out", out], capture_output=True, text=True, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stderr[-2000:]
    return {f: hashlib.sha256(open(os.path.join(out, f), "rb").read()).hexdigest() for f in sorted(os.listdir(out))}


def _expected(key):
    return json.load(open(os.path.join(util.GOLDEN, "expected_hashes.json")))[key]


@pytest.fixture(scope="module", autouse=True)
def _binary():
    assert os.path.exists(BIN), "sift4g_b200/bin/sift4g_b200 missing: build it with make -C sift4g_b200/host (build container)"


def test_reference_fixture_with_subst_files():
    tf = os.path.join(util.GOLDEN, "test_files")
    got = _run(["-q", tf + "/query.fasta", "-d", tf + "/sample_protein_database.fa", "--subst", tf + "/", "--sub-results"])
    assert got == _expected("test_files_subst")
    # byte-for-byte against the committed files too
    for f in got:
        assert got[f] == hashlib.sha256(open(os.path.join(util.GOLDEN, "expected_test_files", f), "rb").read()).hexdigest()


def test_reference_fixture_full_prediction_matrix():
    tf = os.path.join(util.GOLDEN, "test_files")
    assert _run(["-q", tf + "/query.fasta", "-d", tf + "/sample_protein_database.fa"]) == _expected("test_files_nosubst")


def test_synthetic_database_default_flags():
    sd = os.path.join(util.GOLDEN, "synth_e2e")
    assert _run(["-q", sd + "/q.fa", "-d", sd + "/d.fa", "--sub-results"]) == _expected("synth_default")


def test_synthetic_database_with_candidate_cutoff_and_fewer_alignments():
    sd = os.path.join(util.GOLDEN, "synth_e2e")
    got = _run(["-q", sd + "/q.fa", "-d", sd + "/d.fa", "--sub-results", "--max-candidates", "200", "--max-aligns", "50"])
    assert got == _expected("synth_C200_M50")


def test_packed_database_gives_the_same_files(tmp_path):
    # the CLI takes a packed .s4gdb database in place of the FASTA (no parse at all) -- same bytes out
    from sift4g_b200 import capi
    sd = os.path.join(util.GOLDEN, "synth_e2e")
    packed = str(tmp_path / "d.s4gdb")
    capi.pack_fasta(sd + "/d.fa", packed)
    assert _run(["-q", sd + "/q.fa", "-d", packed, "--sub-results"]) == _expected("synth_default")
    pack_tool = os.path.join(ROOT, "sift4g_b200", "bin", "s4g_pack")
    packed2 = str(tmp_path / "d2.s4gdb")
    subprocess.run([pack_tool, sd + "/d.fa", packed2], check=True)
    assert open(packed, "rb").read() == open(packed2, "rb").read()


def test_cli_with_one_shard_per_device_writes_the_same_files(tmp_path):
    # S4G_DEVICES: one resident database shard per listed device inside the CLI (one host thread each, host merge of the
    # candidate lists) -- every output file equals the single-shard / reference bytes, for the FASTA and the packed
    # database.  Listing a device twice gives it two shards, so the 2- and 3-shard (uneven) splits also run on one GPU.
    import torch
    n = torch.cuda.device_count()
    from sift4g_b200 import capi
    sd = os.path.join(util.GOLDEN, "synth_e2e")
    tf = os.path.join(util.GOLDEN, "test_files")
    packed = str(tmp_path / "d.s4gdb")
    capi.pack_fasta(sd + "/d.fa", packed)
    for devs in ["0,0", "0,0,0"] + (["0,1"] if n >= 2 else []) + ([",".join(str(i) for i in range(n))] if n >= 3 else []):
        env = {"S4G_DEVICES": devs}
        assert _run(["-q", sd + "/q.fa", "-d", sd + "/d.fa", "--sub-results"], env) == _expected("synth_default"), devs
        assert _run(["-q", sd + "/q.fa", "-d", packed, "--sub-results", "--max-candidates", "200", "--max-aligns", "50"], env) == _expected("synth_C200_M50"), devs
        assert _run(["-q", tf + "/query.fasta", "-d", tf + "/sample_protein_database.fa", "--subst", tf + "/", "--sub-results"], env) == _expected("test_files_subst"), devs


def test_cli_flags_cards_and_threads():
    # --cards (sift4g/src/main.cpp:123-125: one card index per character) names the GPUs; -t the host threads of the FASTA
    # parse and of the exact hit selection.  Neither changes a byte of the output; an unknown card is refused like the
    # reference refuses it (main.cpp:186 "invalid cuda cards").
    import torch
    n = torch.cuda.device_count()
    sd = os.path.join(util.GOLDEN, "synth_e2e")
    for cards in ["0"] + (["01", "10"] if n >= 2 else []):
        for t in ("1", "5"):
            assert _run(["-q", sd + "/q.fa", "-d", sd + "/d.fa", "--sub-results", "--cards", cards, "-t", t]) == _expected("synth_default"), (cards, t)
    out = tempfile.mkdtemp()
    r = subprocess.run([BIN, "-q", sd + "/q.fa", "-d", sd + "/d.fa", "--out", out, "--cards", "9"], capture_output=True, text=True)
    assert r.returncode != 0 and "invalid cuda cards" in r.stderr


def test_cli_defaults_to_every_visible_gpu():
    # help text of the reference: "--cards ... default: all available CUDA cards" (main.cpp:329-332)
    sd = os.path.join(util.GOLDEN, "synth_e2e")
    out = tempfile.mkdtemp()
    env = {k: v for k, v in os.environ.items() if k not in ("S4G_DEVICES", "S4G_DEVICE")}
    r = subprocess.run([BIN, "-q", sd + "/q.fa", "-d", sd + "/d.fa", "--out", out], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    import torch
    assert "%d GPU" % torch.cuda.device_count() in r.stderr


def test_selection_step_on_the_gpu_and_in_reference_code_write_the_same_files():
    # selectAlignments runs through s4g_alignment_strings / s4g_alignments_select by default (every other
    # test of this module); S4G_SELECT=reference sends main.cpp's call to the reference's own code, still in the binary.
    sd = os.path.join(util.GOLDEN, "synth_e2e")
    tf = os.path.join(util.GOLDEN, "test_files")
    for env in ({"S4G_SELECT": "reference"}, {"S4G_SELECT": "reference", "S4G_DEVICES": "0,0"}):
        assert _run(["-q", sd + "/q.fa", "-d", sd + "/d.fa", "--sub-results"], env) == _expected("synth_default")
        assert _run(["-q", tf + "/query.fasta", "-d", tf + "/sample_protein_database.fa", "--subst", tf + "/", "--sub-results"], env) == _expected("test_files_subst")
    # another median threshold changes the selection: both paths must still agree with each other
    a = _run(["-q", sd + "/q.fa", "-d", sd + "/d.fa", "--sub-results", "--median-threshold", "3.4"])
    b = _run(["-q", sd + "/q.fa", "-d", sd + "/d.fa", "--sub-results", "--median-threshold", "3.4"], {"S4G_SELECT": "reference"})
    assert a == b and a != _expected("synth_default")


def test_alignment_table_from_result_buffers_and_from_reference_code_are_the_same_file():
    # alignments.txt (--sub-results) is written from the result buffers (s4g_alignment_stats on the GPU +
    # s4g_write_blast_tab) for the tabular formats; S4G_WRITER=reference sends main.cpp's call to the reference writer.
    sd = os.path.join(util.GOLDEN, "synth_e2e")
    assert _run(["-q", sd + "/q.fa", "-d", sd + "/d.fa", "--sub-results"], {"S4G_WRITER": "reference"}) == _expected("synth_default")
    for fmt in ("bm8", "bm9", "bm0", "light"):
        a = _run(["-q", sd + "/q.fa", "-d", sd + "/d.fa", "--sub-results", "--outfmt", fmt])
        b = _run(["-q", sd + "/q.fa", "-d", sd + "/d.fa", "--sub-results", "--outfmt", fmt], {"S4G_WRITER": "reference", "S4G_DEVICES": "0,0"})
        assert a == b and "alignments.txt" in a, fmt
