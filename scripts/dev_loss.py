"""Timing of the fused CLIP loss kernels at BASELINE config-3 size vs the torch formulation on the same GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from languagegroundedsemseg_b200 import losses as L
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for c in (96, 512):
    torch.manual_seed(0)
    n = 150_000
    Fe = torch.randn(n, c, device="cuda", requires_grad=True)
    A = torch.randn(200, c, device="cuda")
    y = torch.randint(-1, 200, (n,), device="cuda")
    crit = L.ContrastiveLanguageCELoss(num_labels=200)
    def ours():
        Fe.grad = None
        crit(Fe, y, A)[0].backward()
    def ref():   # the reference's formulation (ContrastiveLanguageLoss.py:224-237), materialising [n,200]
        Fe.grad = None
        S = F.normalize(Fe, dim=1) @ F.normalize(A, dim=1).t()
        F.cross_entropy(S, y, ignore_index=-1).backward()
    t1, t2 = timeit(ours), timeit(ref)
    flops = 2 * 2.0 * n * c * 200   # S and dF GEMMs
    print(f"CE  c={c:3d}: fused kernel fwd+bwd {t1:.3f} ms ({flops/t1/1e9:.1f} TFLOP/s)   torch ops {t2:.3f} ms")
    h = L.ContrastiveLanguageLoss(num_labels=200)
    neg = h.sample_negatives(y)
    def ours_h():
        Fe.grad = None
        h(Fe, y, A, neg_ids=neg)[0].backward()
    print(f"hinge c={c:3d}: fused kernel fwd+bwd {timeit(ours_h):.3f} ms")
