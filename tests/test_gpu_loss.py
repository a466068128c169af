"""-m gpu: fused CLIP text-anchor loss kernels vs the reference-generated golden vectors and the oracle.
Tolerance: fp32 arithmetic, 1e-4 relative (exp/log intrinsics and summation order)."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L(lib):
    from languagegroundedsemseg_b200 import losses
    return losses


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
    Fo = F_.clone().requires_grad_(True)
    lo, po, no = losses_cpu.clip_hinge_loss(Fo, y, A, neg.cpu())
    lo.backward()
    Fg = F_.cuda().requires_grad_(True)
    lg, pg, ng = crit(Fg, y.cuda(), A.cuda(), neg_ids=neg)
    lg.backward()
    assert abs(lg.item() - lo.item()) < 1e-5
    assert rel_err(pg.cpu(), po) < 1e-5 and rel_err(ng.cpu(), no) < 1e-5
    assert rel_err(Fg.grad.cpu(), Fo.grad) < 1e-4
