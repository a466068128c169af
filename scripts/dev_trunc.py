"""Does tcgen05 kind::tf32 truncate (ignore) the low 13 mantissa bits of fp32 operands?  Compare bitwise."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from languagegroundedsemseg_b200 import minkowski as E, _lib
from tests.helpers import random_sparse_coords
lib = _lib.load()
rng = np.random.default_rng(0)
c = random_sparse_coords(rng, 20000, extent=40, batches=1)
x = E.SparseTensor(torch.zeros(c.shape[0], 1).cuda(), torch.from_numpy(c).cuda())
m, k = x.coordinate_manager, x.coordinate_map_key
km = m.kernel_map(k, k, [3, 3, 3], [1, 1, 1])
n, cin, cout = c.shape[0], 96, 96
torch.manual_seed(0)
f = torch.randn(n, cin, device="cuda")
w = torch.randn(27, cin, cout, device="cuda") / 50
mask = lambda t: (t.view(torch.int32) & -8192).view(torch.float32)
def run(ff, ww):
    wf = torch.empty(1, 27, cout, cin, device="cuda")
    _lib.check(lib.lgs_weight_prep(_lib.ptr(ww.contiguous()), 27, cin, cout, 1, _lib.ptr(wf), None, 0, E._stream()))
    out = torch.empty(n, cout, device="cuda")
    _lib.check(lib.lgs_conv_fwd(_lib.ptr(ff.contiguous()), n, cin, _lib.ptr(wf), _lib.W_KNC, 27, cout, _lib.ptr(km.fwd_table), n, 0, None, _lib.ptr(out), 0, _lib.ALGO_TC, E._stream()))
    torch.cuda.synchronize()
    return out
a = run(f, w)
print("A truncated by HW (masking features changes nothing):", torch.equal(a, run(mask(f), w)))
print("B truncated by HW (masking weights changes nothing): ", torch.equal(a, run(f, mask(w))))
print("max |diff| A-masked:", (a - run(mask(f), w)).abs().max().item(), " B-masked:", (a - run(f, mask(w))).abs().max().item())
