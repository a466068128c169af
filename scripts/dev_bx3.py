"""Development harness (GPU): the bf16x3 convolution kernel vs the exact SIMT kernel — correctness over shapes and forced
decompositions, then timing against the 3xTF32 kernel on the config-2 scene.  python scripts/dev_bx3.py [check|time|all]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from languagegroundedsemseg_b200 import minkowski as E, _lib, scenes

lib = _lib.load()
what = sys.argv[1] if len(sys.argv) > 1 else "all"


def tune(**kw):
    for k in ("bx3_tm", "bx3_rt", "bx3_ks", "bx3_ns", "bx3_sa"):
        _lib.check(lib.lgs_tune(k.encode(), int(kw.get(k, 0))))


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp(min=1e-30)).item()


def layer(c, cin, cout, ks=3, bias=False, seed=0):
    torch.manual_seed(seed)
    x = E.SparseTensor(torch.zeros(c.shape[0], 1).cuda(), torch.from_numpy(c).cuda())
    m, k = x.coordinate_manager, x.coordinate_map_key
    km = m.kernel_map(k, k, [ks] * 3, [1, 1, 1]) if ks > 1 else None
    f = torch.randn(c.shape[0], cin, device="cuda")
    w = torch.randn(ks ** 3, cin, cout, device="cuda") / np.sqrt(ks ** 3 * cin)
    b = torch.randn(1, cout, device="cuda") if bias else None
    return f, w, b, km


def check():
    bad = 0
    for nvox in (700, 9000, 40000):
        c, _, _ = scenes.synthetic_voxel_scene(1, nvox)
        for cin, cout, ks, bias in [(96, 96, 3, False), (32, 32, 3, False), (128, 96, 3, False), (256, 256, 3, False), (4, 32, 3, False),
                                    (96, 200, 1, True), (200, 96, 1, False), (384, 256, 3, False), (136, 264, 3, True), (64, 128, 3, False)]:
            if nvox == 40000 and cin * cout > 128 * 128:
                continue
            f, w, b, km = layer(c, cin, cout, ks, bias)
            E.set_conv_algo("simt")
            ref = E.sparse_conv(f, w, b, km)
            E.set_conv_algo("bx3")
            for kw in ({}, dict(bx3_tm=1), dict(bx3_tm=2, bx3_rt=77), dict(bx3_ks=3), dict(bx3_tm=1, bx3_ks=27), dict(bx3_sa=2),
                       dict(bx3_ns=2) if cout >= 128 else dict(bx3_tm=3)):
                if ks == 1 and kw.get("bx3_ks", 0) > 1:
                    continue
                tune(**kw)
                try:
                    out = E.sparse_conv(f, w, b, km)
                    torch.cuda.synchronize()
                    e = rel(out, ref)
                except Exception as ex:  # noqa: BLE001
                    e = float("nan")
                    print("   EXC", ex)
                flag = "" if e < 1e-4 else "   <<<<<< BAD"
                bad += bool(flag)
                print(f"n={c.shape[0]:6d} {cin:3d}->{cout:3d} ks={ks} bias={int(bias)} {str(kw):34s} rel err {e:.2e}{flag}", flush=True)
            tune()
    print("BAD =", bad)


def direct(f, w, b, km, algo):
    """closure launching ONLY the convolution kernel through the C ABI (weights prepared once, output preallocated)"""
    K, cin, cout = w.shape
    A = {"tc": (_lib.ALGO_TC3, 2, _lib.W_KNC_SPLIT), "tf32": (_lib.ALGO_TC, 1, _lib.W_KNC), "bx3": (_lib.ALGO_BX3, 3, _lib.W_BX3)}[algo]
    wf = E._operand_buffer(lib, A[1], K, cout, cin, torch.float32, f.device)
    _lib.check(lib.lgs_weight_prep(_lib.ptr(w), K, cin, cout, A[1], _lib.ptr(wf), None, _lib.F32, E._stream()))
    n = f.shape[0]
    out = torch.empty(n, cout, device=f.device)
    tab = _lib.ptr(km.fwd_table) if km is not None else None
    st = E._stream()
    args = (_lib.ptr(f), n, cin, _lib.ptr(wf), A[2], K, cout, tab, n, 0, _lib.ptr(b), _lib.ptr(out), _lib.F32, A[0], st)
    keep = (wf, out)

    def run():
        _lib.check(lib.lgs_conv_fwd(*args))
    run.keep = keep
    return run


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def time_layers():
    for nvox, shapes in ((150000, [(96, 96), (128, 96), (32, 32), (4, 32)]), (40000, [(32, 32), (96, 96), (128, 96), (64, 32)]),
                         (10600, [(64, 64), (128, 128), (192, 128)]), (2300, [(128, 128), (256, 256), (384, 256)]), (500, [(256, 256)])):
        c, _, _ = scenes.synthetic_voxel_scene(0, nvox)
        for cin, cout in shapes:
            f, w, b, km = layer(c, cin, cout)
            row = f"n={c.shape[0]:6d} {cin:3d}->{cout:3d}:"
            for algo in ("tc", "tf32", "bx3"):
                E.set_conv_algo(algo)
                row += f"  {algo} {timeit(direct(f, w, b, km, algo)):7.1f} us"
            print(row, flush=True)
    # decomposition sweep on the dominant layer
    c, _, _ = scenes.synthetic_voxel_scene(0, 150000)
    f, w, b, km = layer(c, 96, 96)
    E.set_conv_algo("bx3")
    for kw in (dict(bx3_tm=4), dict(bx3_tm=3), dict(bx3_tm=2), dict(bx3_tm=4, bx3_sa=6)):
        tune(**kw)
        print(f"L0 96->96 {kw}: {timeit(direct(f, w, b, km, 'bx3')):7.1f} us", flush=True)
    tune()


if what in ("check", "all"):
    check()
if what in ("time", "all"):
    time_layers()
