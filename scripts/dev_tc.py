"""Development harness (GPU): tcgen05 conv kernel vs the exact SIMT kernel on the same inputs, plus timings."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from languagegroundedsemseg_b200 import minkowski as E, _lib, scenes
from tests.helpers import random_sparse_coords

torch.manual_seed(0)
lib = _lib.load()

def run(algo, feats, w, km, bias=None):
    return E._SparseConvFn.apply(feats, w, bias, km, E._ALGO[algo])

def check(name, c, cin, cout, ks=3, dtype=torch.float32, nscale=1.0):
    cc = torch.from_numpy(c).cuda()
    x = E.SparseTensor(torch.zeros(c.shape[0], 1).cuda(), cc)
    m, k = x.coordinate_manager, x.coordinate_map_key
    km = m.kernel_map(k, k, [ks] * 3, [1, 1, 1]) if ks > 1 else None
    f = torch.randn(c.shape[0], cin, device="cuda").to(dtype).requires_grad_(True)
    K = ks ** 3
    w = (torch.randn(K, cin, cout, device="cuda") / np.sqrt(K * cin)).requires_grad_(True)
    if ks == 1:
        w = w.detach().view(cin, cout).requires_grad_(True)
    gy = torch.randn(c.shape[0], cout, device="cuda").to(dtype)
    res = {}
    for algo in ("simt", "tc", "tf32"):
        f.grad = None
        y = run(algo, f, w, km)
        y.backward(gy)
        torch.cuda.synchronize()
        res[algo] = (y.detach().float(), f.grad.detach().float(), w.grad.detach().float().clone())
        w.grad = None
    msg = ""
    for a in ("tc", "tf32"):
        e_out = ((res[a][0] - res["simt"][0]).abs().max() / res["simt"][0].abs().max()).item()
        e_gin = ((res[a][1] - res["simt"][1]).abs().max() / res["simt"][1].abs().max()).item()
        e_gw = ((res[a][2] - res["simt"][2]).abs().max() / res["simt"][2].abs().max()).item()
        msg += f" | {a}: out={e_out:.1e} dgrad={e_gin:.1e} wgrad={e_gw:.1e}"
    print(f"{name:14s} n={c.shape[0]:7d} {cin:4d}->{cout:4d} ks={ks} {str(dtype)[6:]:8s}{msg}", flush=True)
    return km, f, w

rng = np.random.default_rng(0)
c_small = random_sparse_coords(rng, 5000, extent=24, batches=2)
for cin, cout in [(32, 32), (96, 96), (64, 128), (128, 96), (256, 256), (384, 256), (16, 48), (96, 200), (200, 96), (20, 36)]:
    check("random-sparse", c_small, cin, cout)
check("1x1", c_small, 128, 96, ks=1)
check("bf16", c_small, 64, 96, dtype=torch.bfloat16)
check("bf16", c_small, 96, 96, dtype=torch.bfloat16)
check("bf16-32", c_small, 32, 32, dtype=torch.bfloat16)
check("c512", c_small, 512, 512)

# timing on the config-2 scene
c, _, _ = scenes.synthetic_voxel_scene(0, 150000)
for cin, cout, dtype in [(96, 96, torch.float32), (128, 96, torch.float32), (32, 32, torch.float32), (96, 96, torch.bfloat16)]:
    km, f, w = check("scene150k", c, cin, cout, dtype=dtype)
    P = int(km.counts.sum().item())
    for algo in ("simt", "tc", "tf32"):
        with torch.no_grad():
            for _ in range(3):
                run(algo, f, w, km)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                run(algo, f, w, km)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        s = 2 if dtype == torch.bfloat16 else 4
        byts = P * (cin + cout) * s + 8 * P + 27 * cin * cout * s
        print(f"   {algo:5s} fwd {ms:8.3f} ms  {2*P*cin*cout/ms/1e9:8.2f} TFLOP/s  gather-model {byts/ms/1e6:8.1f} GB/s  (P={P})", flush=True)
        gw = torch.empty(27, cin, cout, device="cuda")
        a = E._ALGO[algo]
        fd, gyd = f.detach().contiguous(), torch.randn(c.shape[0], cout, device="cuda").to(dtype)
        st = E._stream()
        def wg():
            _lib.check(lib.lgs_conv_wgrad(_lib.ptr(fd), fd.shape[0], cin, _lib.ptr(gyd), gyd.shape[0], cout, _lib.ptr(km.fwd_table), 27, _lib.ptr(gw), E._dtype_code(fd), a, st))
        for _ in range(2): wg()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): wg()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"   {algo:5s} wgrad {ms:8.3f} ms  {2*P*cin*cout/ms/1e9:8.2f} TFLOP/s", flush=True)
