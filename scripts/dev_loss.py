"""Timing of the fused CLIP loss kernels at BASELINE config-3 size (150 K points x 200 anchors, C = 96 and 512):
kernel-only (direct C-ABI calls, CUDA events) for the tcgen05 and the SIMT kernel, plus the facade's fwd+bwd next to the
reference's torch formulation on the same GPU.  Writes gpurun_out/loss_bench.json."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from languagegroundedsemseg_b200 import _lib, losses as L

lib = _lib.load()
HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]


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
    return a.elapsed_time(b) / n


def stream():
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


out = {}
for c in (96, 512):
    torch.manual_seed(0)
    n, a = 150_000, 200
    Fe = torch.randn(n, c, device="cuda")
    A = F.normalize(torch.randn(a, c, device="cuda"), dim=1)
    y = torch.randint(-1, a, (n,), device="cuda")
    loss = torch.empty(n, device="cuda")
    pred = torch.empty(n, dtype=torch.int32, device="cuda")
    gf = torch.empty_like(Fe)
    ws = torch.empty(lib.lgs_clip_ce_tc_ws_elems(c, a), device="cuda")
    P = _lib.ptr

    def tc(grad=True):
        _lib.check(lib.lgs_clip_ce_tc(P(Fe), n, c, P(A), a, P(y), -1, P(loss), P(gf) if grad else None, P(pred), None, P(ws), stream()))

    def simt(grad=True):
        _lib.check(lib.lgs_clip_ce(P(Fe), n, c, P(A), a, P(y), -1, P(loss), P(gf) if grad else None, P(pred), None, stream()))

    t_tc, t_tc_f = timeit(tc), timeit(lambda: tc(False))
    t_simt = timeit(simt, 5)
    byts = 3 * n * c * 4 + n * 16            # F read by TMA, F re-read + dF written by the epilogue, labels/loss/pred
    flops = 2 * 2.0 * n * c * a
    crit = L.ContrastiveLanguageCELoss(num_labels=a)
    Fg = Fe.clone().requires_grad_(True)

    def ours():
        Fg.grad = None
        crit(Fg, y, A)[0].backward()

    def ref():   # the reference's formulation (ContrastiveLanguageLoss.py:224-237), materialising [n,200]
        Fg.grad = None
        S = F.normalize(Fg, dim=1) @ F.normalize(A, dim=1).t()
        F.cross_entropy(S, y, ignore_index=-1).backward()

    t_ours, t_ref = timeit(ours), timeit(ref)
    out[f"c{c}"] = {"n": n, "a": a, "tc_kernel_ms": round(t_tc, 4), "tc_fwd_only_ms": round(t_tc_f, 4), "simt_kernel_ms": round(t_simt, 4),
                    "tc_gbs": round(byts / t_tc / 1e6, 1), "tc_frac_of_hbm": round(byts / t_tc / 1e6 / HBM, 3),
                    "tc_tflops_useful": round(flops / t_tc / 1e9, 1), "facade_fwd_bwd_ms": round(t_ours, 4),
                    "torch_fwd_bwd_ms": round(t_ref, 4)}
    print(f"c={c}", out[f"c{c}"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/loss_bench.json", "w"), indent=1)
