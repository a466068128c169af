"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/lgs_b200.h declares.
No compute calls here (no GPU in the build container)."""
import ctypes
import os
import re

from languagegroundedsemseg_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lgs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lgs_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lgs_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes SIGNATURES must mirror the header one to one"


def test_host_only_entry_points(lib):
    assert lib.lgs_version() >= 100
    assert lib.lgs_coord_limit() > 100000
    for n in (0, 1, 511, 512, 513, 150000, 600000):
        cap = lib.lgs_hash_capacity(n)
        assert cap >= 2 * n and cap & (cap - 1) == 0
        assert lib.lgs_coordmap_scratch_elems(n) >= 2 * n
    assert isinstance(lib.lgs_launch_count(), int)


def test_argument_validation_without_gpu(lib):
    # invalid arguments are rejected before any CUDA call
    rc = lib.lgs_conv_fwd(None, 10, 0, None, 0, 27, 8, None, 10, 0, None, None, 0, 0, None)
    assert rc == _lib.E_INVALID and b"lgs_conv_fwd" in lib.lgs_last_error()
    rc = lib.lgs_kmap_build(None, 10, None, None, 1024, 5, 1, 1, None, None, None)
    assert rc == _lib.E_INVALID
    rc = lib.lgs_clip_ce(None, 10, 96, None, 1000, None, -1, None, None, None, None, None)
    assert rc == _lib.E_UNSUPPORTED


def test_library_is_sm100a(lib):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
