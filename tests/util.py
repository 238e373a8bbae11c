"""Shared helpers for the tests (not collected)."""
import json
import os

import numpy as np

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def read_fasta_codes(path):
    """Plain FASTA -> (names, list of code arrays).  Only for well-formed fixture files."""
    names, seqs, cur = [], [], []
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith(">"):
            if names:
                seqs.append(O.encode("".join(cur)))
            names.append(line[1:].strip())
            cur = []
        else:
            cur.append(line)
    if names:
        seqs.append(O.encode("".join(cur)))
    return names, seqs


def seams():
    return json.load(open(os.path.join(GOLDEN, "seams.json")))


def synth_e2e():
    qn, q = read_fasta_codes(os.path.join(GOLDEN, "synth_e2e", "q.fa"))
    dn, d = read_fasta_codes(os.path.join(GOLDEN, "synth_e2e", "d.fa"))
    return qn, q, dn, d


def path_str(p):
    return "".join(map(str, np.asarray(p).tolist()))
