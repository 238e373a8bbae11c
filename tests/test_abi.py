"""CPU: the C-ABI shared library loads and exports every symbol include/sift4g_b200.h declares; without a
CUDA device the product fails loudly instead of falling back."""
import os
import re

import pytest

from sift4g_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "sift4g_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(s4g_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_three_stages():
    names = _declared()
    for n in ("s4g_prefilter", "s4g_sw_score", "s4g_sw_align", "s4g_db_create", "s4g_db_open_fasta", "s4g_merge_candidates"):
        assert n in names


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    for n in _declared():
        assert hasattr(lib, n), "missing export %s" % n
    assert sorted(capi.SIGNATURES) == _declared()


def test_no_torch_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "sift4g_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)      # declarations only
    assert "torch" not in text.lower() and "at::" not in text and "std::" not in text


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.S4GError) as e:
        capi.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sift4g_b200")):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "s4g_oracle" not in src and "from oracle" not in src and "import oracle" not in src, os.path.join(dirpath, f)
    mk = open(os.path.join(ROOT, "sift4g_b200", "host", "Makefile")).read()
    assert "oracle" not in mk
