"""The C-ABI shared library loads and exports every symbol include/picopose_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "picopose_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"PP_API\s+[\w\s\*]+?\b(pp_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from picopose_b200 import build, _lib
    build.build()                      # cross-compiles for sm_100a without a GPU; no-op when up to date
    return _lib.load()


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("pp_corr_lookup", "pp_match_scores", "pp_match_prepare", "pp_topk", "pp_match_similarity",
                 "pp_init_correspondences", "pp_stage3_correspondences", "pp_bilinear_sample", "pp_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    from picopose_b200 import _lib
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes prototype in picopose_b200/_lib.py"
    assert set(_lib.SIGNATURES) <= set(declared_symbols())


def test_pure_host_entry_points(lib):
    assert lib.pp_version() == 100
    assert lib.pp_match_kp(1024, 0) == 1024 and lib.pp_match_kp(384, 0) == 384
    assert lib.pp_match_kp(40, 0) == 64                      # K padded to the 64-element swizzle row
    assert lib.pp_match_kp(384, 2) == 1152 and lib.pp_match_kp(384, 1) == 2304
    assert lib.pp_match_kp(0, 0) < 0 and lib.pp_match_kp(64, 7) < 0
    # two 64-bit keys per (b, n, t) + two per-detection int arrays; query bookkeeping = 3 words per patch + 2 per detection
    assert lib.pp_match_scores_workspace(1, 162, 1024) == 2 * 162 * 1024 * 8 + 2 * 256
    assert lib.pp_match_query_meta_bytes(2, 1024) == (3 * 2 * 1024 + 4) * 4
    assert lib.pp_match_similarity_workspace(2, 256) == 0          # the volume is written by the contraction's epilogue
    assert lib.pp_match_similarity_dense_workspace(2, 1024, 16, 16, 0) == 2 * 2 * 256 * 1024 * 2 + 2 * 2 * 256 * 4
    assert isinstance(lib.pp_launch_count(), int)


def test_sass_is_blackwell_native():
    """The stage-1 kernel must contain tcgen05 / TMA / TMEM-load instructions (SASS names), not legacy HMMA."""
    import shutil
    import subprocess
    from picopose_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    obj = os.path.join(build.LIB_DIR, "match_gemm.o")
    sass = subprocess.run([cuobjdump, "-sass", obj], capture_output=True, text=True, check=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass
