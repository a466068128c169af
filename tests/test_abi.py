"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/lgs_b200.h declares.
No compute calls here (no GPU in the build container)."""
import ctypes
import os
import re

import pytest

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


_BINDING_PROBE = r'''
import sys, ctypes
sys.path.insert(0, ROOT)
import torch
from languagegroundedsemseg_b200 import _lib
lib = _lib.load()
x = torch.zeros(8)
out = [_lib.binding()]
calls = [
    ("lgs_conv_fwd", (None, 10, 0, None, 0, 27, 8, None, 10, 0, None, None, 0, 0, None)),
    ("lgs_conv_fwd", (_lib.ptr(x), 123456789012, 7, _lib.ptr(x), 0, 99, 8, None, -5, 0, None, None, 0, 0, None)),
    ("lgs_conv_wgrad", (_lib.ptr(x), 5, 0, _lib.ptr(x), 5, 8, None, 27, _lib.ptr(x), 0, 1, None)),
    ("lgs_bn_fwd", (_lib.ptr(x), None, 1000, 6, None, None, 1e-5, 0.1, 1, None, None, _lib.ptr(x), _lib.ptr(x), _lib.ptr(x),
                    _lib.ptr(x), None, None, None)),
    ("lgs_bn_bwd", (_lib.ptr(x), None, _lib.ptr(x), 10, 2048, None, _lib.ptr(x), _lib.ptr(x), 0, _lib.ptr(x), None, None, None,
                    _lib.ptr(x), None, None)),
    ("lgs_seg_ce", (_lib.ptr(x), 5, 7, _lib.ptr(x), -1, _lib.ptr(x), _lib.ptr(x), None, ctypes.c_void_p(0))),
    ("lgs_weight_prep", (_lib.ptr(x), 27, 0, 8, 2, _lib.ptr(x), None, 0, None)),
    ("lgs_weight_prep_batch", (None, 3, 10, 2, 0, None)),
    ("lgs_kmap_build", (None, 10, None, None, 1024, 5, 1, 1, None, None, None)),
    ("lgs_clip_ce", (None, 10, 96, None, 1000, None, -1, None, None, None, None, None)),
]
for name, args in calls:
    rc = getattr(lib, name)(*args)
    out.append((name, rc, lib.lgs_last_error().decode()))
print(repr(out))
'''


def test_native_binding_matches_ctypes(lib):
    """the generated CPython binding (csrc/_lgs_fast*.so, LGS_FAST_BIND=1) hands the C library the same argument values
    as ctypes: identical return codes and identical error strings (which echo the integer arguments) on calls that stop
    at argument validation — nothing here needs a GPU"""
    import subprocess
    import sys
    res = {}
    for mode in ("0", "1"):
        r = subprocess.run([sys.executable, "-c", f"ROOT={ROOT!r}\n" + _BINDING_PROBE], capture_output=True, text=True,
                           env=dict(os.environ, LGS_FAST_BIND=mode), timeout=300)
        assert r.returncode == 0, r.stderr
        res[mode] = eval(r.stdout.strip().splitlines()[-1])
    assert res["0"][0] == "ctypes" and res["1"][0] == "native"
    assert res["0"][1:] == res["1"][1:]
    assert all(rc != 0 for _, rc, _ in res["0"][1:])


_SITES_PROBE = r'''
import sys, re
sys.path.insert(0, ROOT)
import torch
from tests import stub_engine
from languagegroundedsemseg_b200 import _lib, losses, minkowski as E
stub_engine.install(setattr, real_library=True)
losses._stream = lambda: None
with _lib.trace() as t:
    # kernel-map call sites (CoordinateManager._table / ._transpose) on hand-made coordinate maps
    mgr = E.CoordinateManager(D=3)
    for ts, n in ((1, 40), (2, 12)):
        cm = E._CoordMap()
        cm.coords = torch.zeros((n, 4), dtype=torch.int32)
        cm.tkeys, cm.tvals = torch.zeros(128, dtype=torch.int64), torch.zeros(128, dtype=torch.int32)
        cm.n, cm.capacity = n, 128
        mgr._maps[E.CoordinateMapKey([ts] * 3)] = cm
    k1, k2 = E.CoordinateMapKey([1, 1, 1]), E.CoordinateMapKey([2, 2, 2])
    mgr.kernel_map(k1, k1, [3, 3, 3], [1, 1, 1])
    mgr.kernel_map(k1, k2, [2, 2, 2], [1, 1, 1])
    mgr.kernel_map(k2, k1, [2, 2, 2], [1, 1, 1], True)
    # loss call sites
    x = torch.randn(33, 200, requires_grad=True)
    y = torch.randint(-1, 200, (33,))
    losses._SegCEFn.apply(x, y, -1).backward()
    f = torch.randn(50, 96, requires_grad=True)
    a = torch.nn.functional.normalize(torch.randn(200, 96), dim=1)
    lab = torch.randint(-1, 200, (50,))
    for algo in ("tc", "simt"):
        losses.set_clip_algo(algo)
        losses._ClipCEFn.apply(f, a, lab, -1)
    neg = torch.randint(0, 200, (50, 3), dtype=torch.int32)
    losses._ClipHingeFn.apply(f, a, lab, neg, -1, 0.0, 0.6, 1.0)
lines = [re.sub(r"0x[0-9a-f]+", "A", l).replace("(nil)", "0") for l in t.lines]
print(_lib.binding())
print(*lines, sep=chr(10))
'''


def test_remaining_call_sites_through_both_bindings(lib):
    """kernel-map construction and the three loss entry points as the facade calls them, recorded by the library
    (lgs_trace_begin): same calls and argument values through ctypes and through the native binding"""
    import subprocess
    import sys
    res = {}
    for mode in ("0", "1"):
        r = subprocess.run([sys.executable, "-c", f"ROOT={ROOT!r}\n" + _SITES_PROBE], capture_output=True, text=True,
                           env=dict(os.environ, LGS_FAST_BIND=mode), timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out = r.stdout.strip().splitlines()
        res[mode] = (out[0], out[1:])
    assert (res["0"][0], res["1"][0]) == ("ctypes", "native")
    assert res["0"][1] == res["1"][1]
    names = [l.split()[0] for l in res["0"][1]]
    assert names == ["lgs_kmap_build", "lgs_kmap_build", "lgs_kmap_transpose", "lgs_seg_ce", "lgs_clip_ce_tc", "lgs_clip_ce",
                     "lgs_clip_hinge"], names


def test_native_binding_fuzz_against_ctypes(lib):
    """random argument tuples (full int32 / int64 ranges, pointers up to 2^63, floats) through every wrapper of the native
    binding and through ctypes, in one process: traced entries must record identical lines, untraced ones (the *_supported
    queries) must return the same value"""
    import ctypes as C
    import random
    if _lib.binding() != "native":
        pytest.skip("native binding not built")
    fast, cdll = _lib._fast, lib._cdll
    rng = random.Random(1234)
    gen = {C.c_void_p: lambda: rng.choice([None, 0, rng.randrange(1, 2 ** 63), rng.randrange(1, 2 ** 47)]),
           C.c_int64: lambda: rng.choice([0, -1, 2 ** 63 - 1, -2 ** 63, rng.randrange(-2 ** 40, 2 ** 40)]),
           C.c_int32: lambda: rng.choice([0, 1, -1, 2 ** 31 - 1, -2 ** 31, rng.randrange(-10 ** 6, 10 ** 6)]),
           C.c_float: lambda: rng.choice([0.0, 1e-5, -3.5, 0.1, 1e30, rng.random()])}
    # lgs_program_run dereferences its handle (it IS the caller of the recorded entry points): not fuzzable with random pointers
    wrapped = [n for n in _lib.SIGNATURES if hasattr(fast, n) and not n.startswith("lgs_program_")]
    assert len(wrapped) >= 15
    for name in wrapped:
        res, argtypes = _lib.SIGNATURES[name]
        cfn = getattr(cdll, name)
        for _ in range(40):
            args = [gen[t]() for t in argtypes]
            outs = []
            for fn in (cfn, getattr(fast, name)):
                lib.lgs_trace_begin()
                rc = fn(*args)
                need = lib.lgs_trace_end(None, 0)
                buf = C.create_string_buffer(int(need))
                lib.lgs_trace_end(buf, need)
                outs.append((rc, buf.value))
            assert outs[0] == outs[1], (name, args, outs)
