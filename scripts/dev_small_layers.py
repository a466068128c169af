"""GPU time of the conv kernels on the coarse levels of the 150 K-voxel scene, straight through the C ABI in a tight
loop (no Python-side per-layer work, so the numbers are kernel + memset time, not host issue time), for several
decompositions (LGS_TC_SPLIT_TARGET = CTAs aimed for when a map has fewer 128-row tiles than SMs)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from languagegroundedsemseg_b200 import _lib, minkowski as E

lib = _lib.load()
c, f, _ = bench.make_scene(0)
st = E.SparseTensor(torch.from_numpy(f).cuda(), torch.from_numpy(c).cuda())
mgr, key = st.coordinate_manager, st.coordinate_map_key
levels = []
for lvl in range(5):
    levels.append((key, mgr.kernel_map(key, key, [3, 3, 3], [1, 1, 1])))
    if lvl < 4:
        key = mgr.stride(key, 2)
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
CASES = [(1, 32, 32), (1, 96, 96), (2, 64, 64), (2, 128, 128), (3, 128, 128), (3, 256, 256), (4, 256, 256)]
TARGETS = [int(t) for t in os.environ.get("TARGETS", "37,74,148,296").split(",")]
REP = 50
print(f"{'level rows cin->cout':28s} " + " ".join(f"T={t:<4d}" for t in TARGETS) + "   (us per launch, fwd | wgrad)")
for lvl, cin, cout in CASES:
    key, km = levels[lvl]
    n = km.n_out
    x = torch.randn(n, cin, device="cuda")
    gy = torch.randn(n, cout, device="cuda")
    w = torch.randn(27, cin, cout, device="cuda") * 0.05
    wf = torch.empty((2, 27, cout, cin), device="cuda")
    _lib.check(lib.lgs_weight_prep(_lib.ptr(w), 27, cin, cout, 2, _lib.ptr(wf), None, _lib.F32, stream))
    out = torch.empty(n, cout, device="cuda")
    gw = torch.empty(27, cin, cout, device="cuda")
    res = []
    for t in TARGETS:
        os.environ["LGS_TC_SPLIT_TARGET"] = str(t)

        def fwd():
            _lib.check(lib.lgs_conv_fwd(_lib.ptr(x), n, cin, _lib.ptr(wf), _lib.W_KNC_SPLIT, 27, cout, _lib.ptr(km.fwd_table),
                                        n, 0, None, _lib.ptr(out), _lib.F32, _lib.ALGO_TC3, stream))

        def wgrad():
            _lib.check(lib.lgs_conv_wgrad(_lib.ptr(x), n, cin, _lib.ptr(gy), n, cout, _lib.ptr(km.fwd_table), 27, _lib.ptr(gw),
                                          _lib.F32, _lib.ALGO_TC3, stream))
        r = []
        for fn in (fwd, wgrad):
            for _ in range(5):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(REP):
                fn()
            e1.record()
            torch.cuda.synchronize()
            r.append(1e3 * e0.elapsed_time(e1) / REP)
        res.append(r)
    print(f"L{lvl} {n:7d} {cin:3d}->{cout:3d}".ljust(28) + " " + " ".join(f"{a:5.1f}|{b:5.1f}" for a, b in res), flush=True)
