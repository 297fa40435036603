"""CPU-side checks of the drop-in boundary: libsbv.so loads, exports every symbol include/sbv.h
declares, and refuses to run without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "sbv.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sbv_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_reference_facing_entry_points():
    syms = _declared_symbols()
    for s in ["sbv_create", "sbv_destroy", "sbv_verify_batch", "sbv_verify_batch_der", "sbv_hash_verify_batch",
              "sbv_sha256_batch", "sbv_quorum", "sbv_verify_mixed", "sbv_set_keys", "sbv_compute_quorum"]:
        assert s in syms
    # every entry point cites the reference interface it replaces
    hdr = open(os.path.join(ROOT, "include", "sbv.h")).read()
    for ref in ["dependencies.go:54-71", "view.go:519-551", "util.go:183-187", "types.go:50-69"]:
        assert ref in hdr


def test_library_exports_every_declared_symbol():
    import consensus_b200 as sbv
    lib = sbv.load_library()
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(sbv.SYMBOLS) == sorted(_declared_symbols())


def test_compute_quorum_matches_reference_table():
    import consensus_b200 as sbv
    # /root/reference/internal/bft/util_test.go:144-154
    for n, f, q in [(4, 1, 3), (5, 1, 4), (6, 1, 4), (7, 2, 5), (8, 2, 6), (9, 2, 6), (10, 3, 7), (11, 3, 8), (12, 3, 8)]:
        assert sbv.compute_quorum(n) == (q, f)


def test_no_cpu_fallback():
    import torch
    import consensus_b200 as sbv
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the fault path is exercised on CPU-only boxes")
    with pytest.raises(sbv.EngineFault):
        sbv.Engine(n_devices=1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "consensus_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".inc")) and "host_tests" not in f:
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f
