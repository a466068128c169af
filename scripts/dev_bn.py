import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from languagegroundedsemseg_b200 import minkowski as E
torch.manual_seed(0)
for n, c in ((149106, 96), (149106, 32), (38506, 96), (2094, 256)):
    x = torch.randn(n, c, device="cuda", requires_grad=True)
    res = torch.randn(n, c, device="cuda")
    bn = torch.nn.BatchNorm1d(c).cuda().train()
    gz = torch.randn(n, c, device="cuda")
    def fused():
        p = E._PendingBN(bn, x); p.res = res
        z = E._bn_act(p, relu=True); z.backward(gz); x.grad = None
    def aten():
        z = torch.relu(bn(x) + res); z.backward(gz); x.grad = None
    for name, fn in (("fused", fused), ("aten", aten)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): fn()
        b.record(); torch.cuda.synchronize()
        print(f"n={n:7d} c={c:4d} {name:6s} fwd+bwd {a.elapsed_time(b)/20*1e3:8.1f} us  (unroll={os.environ.get('LGS_BN_UNROLL','4')})")
