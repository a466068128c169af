"""-m gpu: fused CLIP text-anchor loss kernels vs the reference-generated golden vectors and the oracle.
Tolerance: fp32 arithmetic, 1e-4 relative (exp/log intrinsics and summation order)."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["tc", "simt"])
def L(lib, request):
    """both kernels behind the same facade: the tcgen05 one (default) and the exact-fp32 SIMT one"""
    from languagegroundedsemseg_b200 import losses
    losses.set_clip_algo(request.param)
    yield losses
    losses.set_clip_algo("tc")


@pytest.mark.parametrize("tag", ["c96", "c512"])
def test_clip_ce_vs_reference_golden(L, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "clip_ce.npz"))
    F_ = torch.from_numpy(g[f"{tag}_F"]).cuda().requires_grad_(True)
    A, y = torch.from_numpy(g[f"{tag}_A"]).cuda(), torch.from_numpy(g[f"{tag}_y"]).cuda()
    loss = L.ContrastiveLanguageCELoss(num_labels=200, reduction="none")(F_, y, A)[0]
    np.testing.assert_allclose(loss.detach().cpu().numpy(), g[f"{tag}_loss"], rtol=1e-4, atol=1e-5)
    assert torch.all(loss[y == -1] == 0)
    crit = L.ContrastiveLanguageCELoss(num_labels=200, reduction="mean")
    lm = crit(F_, y, A)[0]
    assert abs(lm.item() - float(g[f"{tag}_mean"])) < 1e-4
    lm.backward()
    assert rel_err(F_.grad.cpu(), torch.from_numpy(g[f"{tag}_grad"])) < 1e-4
    assert torch.all(F_.grad[y == -1] == 0)
    # argmax predictions == feature_sim argmax of the oracle
    from oracle import losses_cpu
    S = losses_cpu.feature_sim(torch.from_numpy(g[f"{tag}_F"]), torch.from_numpy(g[f"{tag}_A"]))
    assert torch.equal(crit.last_pred.cpu().long(), S.argmax(1))
    assert torch.equal(L.feature_sim_argmax(F_.detach(), A).cpu(), S.argmax(1))


def test_clip_ce_anchor_grad_and_scale_invariance(L):
    from oracle import losses_cpu
    torch.manual_seed(0)
    n, c = 777, 96
    F_, A = torch.randn(n, c), torch.randn(200, c)
    y = torch.randint(0, 200, (n,))
    y[::7] = -1
    Fo, Ao = F_.clone().requires_grad_(True), A.clone().requires_grad_(True)
    losses_cpu.clip_ce_loss(Fo, y, Ao).backward()
    Fg, Ag = F_.cuda().requires_grad_(True), A.cuda().requires_grad_(True)
    lg = L.ContrastiveLanguageCELoss(num_labels=200)(Fg, y.cuda(), Ag)[0]
    lg.backward()
    assert rel_err(Fg.grad.cpu(), Fo.grad) < 1e-4 and rel_err(Ag.grad.cpu(), Ao.grad) < 1e-4
    l2 = L.ContrastiveLanguageCELoss(num_labels=200)(Fg.detach() * 3.7, y.cuda(), Ag.detach())[0]
    assert abs(l2.item() - lg.item()) < 1e-5


def test_clip_ce_full_size(L):
    """BASELINE config-3 shape: 150K points x 512-d vs 200 anchors; mean loss vs torch on the same device data"""
    torch.manual_seed(1)
    n, c = 150_000, 512
    F_ = torch.randn(n, c, device="cuda")
    A = torch.randn(200, c, device="cuda")
    y = torch.randint(-1, 200, (n,), device="cuda")
    loss, pred = L.clip_ce(F_, y, A)
    S = torch.nn.functional.normalize(F_, dim=1) @ torch.nn.functional.normalize(A, dim=1).t()
    ref = torch.nn.functional.cross_entropy(S, y, ignore_index=-1, reduction="none")
    assert rel_err(loss, ref) < 1e-4
    assert (pred.long() == S.argmax(1)).float().mean() > 0.9999


def test_clip_hinge_vs_oracle(L):
    from oracle import losses_cpu
    torch.manual_seed(2)
    n, c = 1500, 96
    F_, A = torch.randn(n, c), torch.randn(200, c)
    y = torch.randint(0, 200, (n,))
    y[::5] = -1
    crit = L.ContrastiveLanguageLoss(num_labels=200)
    neg = crit.sample_negatives(y.cuda())
    assert torch.all(neg.cpu() != y.clamp(min=0)[:, None]) and neg.min() >= 0 and neg.max() < 200
    Fo, Ao = F_.clone().requires_grad_(True), A.clone().requires_grad_(True)
    lo, po, no = losses_cpu.clip_hinge_loss(Fo, y, Ao, neg.cpu())
    lo.backward()
    Fg, Ag = F_.cuda().requires_grad_(True), A.cuda().requires_grad_(True)
    lg, pg, ng = crit(Fg, y.cuda(), Ag, neg_ids=neg)
    lg.backward()
    assert abs(lg.item() - lo.item()) < 1e-5
    assert rel_err(pg.cpu(), po) < 1e-5 and rel_err(ng.cpu(), no) < 1e-5
    assert rel_err(Fg.grad.cpu(), Fo.grad) < 1e-4
    # the anchors get their gradient too (it trains projection_layer of Res16UNet34CR_Proj, clip_models.py:197-200)
    assert Ag.grad is not None and rel_err(Ag.grad.cpu(), Ao.grad) < 1e-4
    # 'use all labels' setting of the reference (ContrastiveLanguageLoss.py:33-36)
    assert L.ContrastiveLanguageLoss(num_labels=200, num_negative_samples=-1).num_negative_samples == 200


@pytest.mark.parametrize("n,c,a", [(1, 96, 200), (127, 96, 200), (129, 16, 20), (1000, 100, 200), (333, 512, 200),
                                   (4097, 96, 208), (640, 32, 4), (2000, 192, 56)])
def test_clip_ce_tc_shapes_vs_oracle(lib, n, c, a):
    """tcgen05 kernel on ragged shapes (partial tiles, 16-column tails, K padding, multi-chunk dF) vs the fp32 oracle:
    loss / dF / d anchors within 1e-4 relative, argmax identical where the top-2 margin exceeds 1e-5."""
    from languagegroundedsemseg_b200 import losses
    from oracle import losses_cpu
    assert lib.lgs_clip_ce_tc_supported(c, a) == 1
    losses.set_clip_algo("tc")
    torch.manual_seed(n + c + a)
    F_, A = torch.randn(n, c) * 3.0, torch.randn(a, c)
    y = torch.randint(0, a, (n,))
    y[1::9] = -1            # row 0 stays labelled (n = 1: an all-ignored batch has no defined mean)
    Fo, Ao = F_.clone().requires_grad_(True), A.clone().requires_grad_(True)
    lo = losses_cpu.clip_ce_loss(Fo, y, Ao)
    lo.backward()
    Fg, Ag = F_.cuda().requires_grad_(True), A.cuda().requires_grad_(True)
    l0 = lib.lgs_launch_count()
    crit = losses.ContrastiveLanguageCELoss(num_labels=a)
    lg = crit(Fg, y.cuda(), Ag)[0]
    assert lib.lgs_launch_count() - l0 == 2          # weight_prep (anchor split) + the fused kernel
    lg.backward()
    assert abs(lg.item() - lo.item()) < 1e-4 * max(1.0, abs(lo.item()))
    assert rel_err(Fg.grad.cpu(), Fo.grad) < 1e-4
    assert rel_err(Ag.grad.cpu(), Ao.grad) < 1e-4
    assert torch.all(Fg.grad[y.cuda() == -1] == 0)
    S = losses_cpu.feature_sim(F_, A)
    top2 = S.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-5
    assert torch.equal(crit.last_pred.cpu().long()[clear], S.argmax(1)[clear])
    # per-row losses too
    lr = losses.ContrastiveLanguageCELoss(num_labels=a, reduction="none")(Fg.detach(), y.cuda(), Ag.detach())[0]
    Sd = S.double()
    ref = torch.nn.functional.cross_entropy(Sd, y, ignore_index=-1, reduction="none").float()
    np.testing.assert_allclose(lr.cpu().numpy(), ref.numpy(), rtol=1e-4, atol=2e-5)


def test_clip_ce_tc_rejects_unsupported_shapes(lib):
    from languagegroundedsemseg_b200 import _lib
    assert lib.lgs_clip_ce_tc_supported(96, 200) == 1
    assert lib.lgs_clip_ce_tc_supported(97, 200) == 0 and lib.lgs_clip_ce_tc_supported(96, 256) == 0
    rc = lib.lgs_clip_ce_tc(None, 10, 97, None, 200, None, -1, None, None, None, None, None, None)
    assert rc == _lib.E_UNSUPPORTED and b"lgs_clip_ce_tc" in lib.lgs_last_error()
    # the facade serves such shapes with the SIMT kernel
    from languagegroundedsemseg_b200 import losses
    F_, A = torch.randn(50, 97, device="cuda"), torch.randn(30, 97, device="cuda")
    y = torch.randint(0, 30, (50,), device="cuda")
    loss, _ = losses.clip_ce(F_, y, A)
    S = torch.nn.functional.normalize(F_, dim=1) @ torch.nn.functional.normalize(A, dim=1).t()
    assert rel_err(loss, torch.nn.functional.cross_entropy(S, y, reduction="none")) < 1e-4


@pytest.mark.parametrize("n,c,ignored", [(1, 200, 0.0), (7, 20, 0.3), (5000, 200, 0.1), (4097, 1024, 0.5), (300, 4, 0.0),
                                          (64, 200, 1.0)])
def test_seg_cross_entropy_vs_oracle(lib, n, c, ignored):
    """fused softmax cross-entropy (lgs_seg_ce) = nn.CrossEntropyLoss(ignore_index=-1): loss and d loss / d logits
    against the float64 oracle, including rows with large logits, all-ignored input and a scaled upstream gradient"""
    from languagegroundedsemseg_b200 import losses
    from oracle import losses_cpu
    torch.manual_seed(n + c)
    x = (torch.randn(n, c) * 4).requires_grad_(True)
    x.data[0, : min(c, 3)] += 60.0                       # needs the max-subtraction
    y = torch.randint(0, c, (n,))
    y[torch.rand(n) < ignored] = -1
    ref = losses_cpu.seg_ce_loss(x, y, -1)
    if ignored < 1.0:
        (ref * 0.5).backward()
    xg = x.detach().cuda().requires_grad_(True)
    out = losses.cross_entropy(xg, y.cuda(), ignore_index=-1)
    if ignored == 1.0:
        assert torch.isnan(out).item() == torch.isnan(ref).item()
        return
    (out * 0.5).backward()
    assert abs(out.item() - ref.item()) < 5e-6 * max(1.0, abs(ref.item()))
    assert rel_err(xg.grad.cpu(), x.grad.float()) < 2e-5
    assert torch.equal(xg.grad[y.cuda() == -1], torch.zeros_like(xg.grad[y.cuda() == -1]))


def test_seg_cross_entropy_full_size_and_fallback(lib):
    """BASELINE config-2 size (150 K points x 200 classes) vs ATen on the same device; class counts outside the kernel's
    envelope take ATen's path"""
    from languagegroundedsemseg_b200 import losses
    torch.manual_seed(0)
    x = torch.randn(150000, 200, device="cuda")
    y = torch.randint(-1, 200, (150000,), device="cuda")
    a = x.clone().requires_grad_(True)
    b = x.clone().requires_grad_(True)
    la = losses.CrossEntropyLoss(ignore_index=-1)(a, y)
    lb = torch.nn.functional.cross_entropy(b, y, ignore_index=-1)
    la.backward(), lb.backward()
    assert abs(la.item() - lb.item()) < 1e-5 * lb.item() and rel_err(a.grad, b.grad) < 1e-4
    z = torch.randn(50, 7, device="cuda", requires_grad=True)            # 7 classes: not a multiple of 4
    t = torch.randint(0, 7, (50,), device="cuda")
    assert torch.allclose(losses.cross_entropy(z, t), torch.nn.functional.cross_entropy(z, t))
