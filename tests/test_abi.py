"""CPU tests of the drop-in boundary: the shared library loads without a GPU and exports every symbol that
include/pcf_b200.h declares; the ctypes table in _lib.py covers exactly the declared entry points."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pcf_b200.h")
LIB = os.path.join(ROOT, "ml-pointconvformer_b200", "libpcf_b200.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcfb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for must in ("pcfb_knn_packed", "pcfb_knn_inverse", "pcfb_gather", "pcfb_gather_backward", "pcfb_edge_geometry",
                 "pcfb_pconv_forward", "pcfb_pconv_backward", "pcfb_gridsub_emit"):
        assert must in syms


@pytest.mark.skipif(not os.path.exists(LIB), reason="libpcf_b200.so not built (run __graft_entry__.build())")
def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(LIB)                     # loads without a GPU (no CUDA call at load time)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    lib.pcfb_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.pcfb_version()


def test_ctypes_table_matches_header():
    import pcf_b200  # noqa: F401
    from pcf_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_header_cites_the_reference():
    src = open(HEADER).read()
    for cite in ("pcf.h:243-250", "pcf.h:213-224", "knn.cu:104-168", "knn_post_dataloader_utils.py:22-41",
                 "grid_subsampling.cpp:9-110", "layer_utils.py:176-231"):
        assert cite in src, cite


def test_sass_has_tcgen05_and_no_legacy_mma():
    """The tensor-core variant must really be tcgen05 (SASS UTC*MMA + LDTM), not mma.sync (HMMA)."""
    import shutil
    import subprocess
    if not os.path.exists(LIB) or shutil.which("cuobjdump") is None:
        pytest.skip("needs the built library and cuobjdump")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    assert re.search(r"\bUTC\w*MMA", sass), "no tcgen05.mma in SASS"
    assert "LDTM" in sass, "no tcgen05.ld in SASS"
    assert not re.search(r"\bHMMA\b", sass), "legacy mma.sync found"
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
