"""CPU tests of the facade's host logic against a stub of the C-ABI library (tests/stub_engine.py): no arithmetic, but
the autograd plumbing, call counts and cache-invalidation rules are the real code."""
import pytest
import torch

from tests import stub_engine

SIZES = {1: 600, 2: 200, 4: 70, 8: 25, 16: 9}


@pytest.fixture
def stub(monkeypatch):
    from languagegroundedsemseg_b200 import minkowski as E
    s = stub_engine.install(monkeypatch.setattr)
    yield s
    E.set_conv_bn_fusion(True), E.set_wgrad_overlap(True), E.set_batched_weight_prep(True)


@pytest.mark.parametrize("fuse,overlap,batch", [(1, 1, 1), (0, 0, 0), (1, 0, 1), (0, 1, 0)])
def test_unet34c_host_plumbing(stub, fuse, overlap, batch):
    """two SGD steps of Res16UNet34C through the facade: every parameter receives a gradient of its own shape, and the
    number of C-ABI launches per step is what the path promises (63 convs fwd, 62 dgrad, 63 wgrad; one batched
    weight-operand launch per step instead of 62 per-layer ones; 62 fused BatchNorms when conv+BN fusion is on)"""
    from languagegroundedsemseg_b200 import minkowski as E, nets
    E.set_conv_bn_fusion(fuse), E.set_wgrad_overlap(overlap), E.set_batched_weight_prep(batch)
    torch.manual_seed(0)
    net = nets.build_model("Res16UNet34C", 3, 200, nets.DefaultConfig()).train()
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9)
    mgr = stub_engine.FakeManager(SIZES)
    for _ in range(2):
        out, feat = net(stub_engine.sparse_input(SIZES[1], 3, mgr))
        assert out.F.shape == (SIZES[1], 200) and feat.F.shape == (SIZES[1], 96)
        opt.zero_grad(set_to_none=True)
        out.F.float().mean().backward()
        assert not [k for k, p in net.named_parameters() if p.grad is None or p.grad.shape != p.shape]
        opt.step()
    c = stub.calls
    assert c["lgs_conv_fwd"] == 2 * (63 + 62) and c["lgs_conv_wgrad"] == 2 * 63
    if batch:
        assert c["lgs_weight_prep_batch"] == 2 and c["lgs_weight_prep"] == 2      # conv0p1s1 (c_in 3 -> 4) pads per call
    else:
        assert c["lgs_weight_prep"] == 2 * 63 and "lgs_weight_prep_batch" not in c
    if fuse:
        assert c["lgs_bn_fwd"] == c["lgs_bn_bwd"] == 2 * 62
    net.eval()                      # BatchNorm not fusable: lazy convs materialise through .F
    with torch.no_grad():
        assert net(stub_engine.sparse_input(SIZES[1], 3, stub_engine.FakeManager(SIZES)))[0].F.shape == (SIZES[1], 200)


def test_weight_operand_cache_invalidation(stub):
    """cached tensor-core weight operands are re-derived after an optimiser step (torch's fused optimisers do not bump the
    parameter's version counter), after any weight-gradient computation, after an in-place edit, and on request"""
    from languagegroundedsemseg_b200 import minkowski as E
    conv = E.MinkowskiConvolution(16, 16, kernel_size=1, dimension=3)
    mgr = stub_engine.FakeManager({1: 100})

    def fwd():
        return conv(stub_engine.sparse_input(100, 16, mgr)).F

    def n():
        return stub.calls.get("lgs_weight_prep_batch", 0)

    fwd()
    a = n()
    fwd()
    assert n() == a                                   # cached
    opt = torch.optim.SGD(conv.parameters(), lr=0.1, fused=True)
    v = conv.kernel._version
    conv.kernel.grad = torch.zeros_like(conv.kernel)
    opt.step()
    if conv.kernel._version == v:                     # the reason rule (b) exists
        pass
    fwd()
    assert n() == a + 1                               # optimiser post-step hook
    fwd()
    assert n() == a + 1
    fwd().sum().backward()
    fwd()
    assert n() == a + 2                               # a weight gradient was computed
    with torch.no_grad():
        conv.kernel.mul_(2.0)
    fwd()
    assert n() == a + 3                               # version counter
    E.invalidate_weight_cache()
    fwd()
    assert n() == a + 4
