"""CPU dry run of the facade's host logic (no GPU, no arithmetic): the C-ABI library is replaced by a stub whose entry
points return LGS_OK, coordinate/kernel maps by tables of the right shapes.  Exercises the autograd plumbing of the
lazy conv / fused conv+BN node / batched weight prep / side-stream wgrad paths (argument counts, saved tensors,
gradient arity and shapes) before a GPU box is spent on them.  `python scripts/dev_dryrun_cpu.py`"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from languagegroundedsemseg_b200 import _lib, minkowski as E, nets


from tests import stub_engine
from tests.stub_engine import FakeManager


def main():
    stub = stub_engine.install(setattr)
    sizes = {1: 1000, 2: 300, 4: 100, 8: 30, 16: 10}
    for flags in ((True, True, True), (False, False, False), (True, False, True), (False, True, False)):
        E.set_conv_bn_fusion(flags[0]), E.set_wgrad_overlap(flags[1]), E.set_batched_weight_prep(flags[2])
        stub.calls.clear()
        torch.manual_seed(0)
        net = nets.build_model("Res16UNet34C", 3, 200, nets.DefaultConfig()).train()
        opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9)
        mgr = FakeManager(sizes)
        for step in range(2):
            x = E.SparseTensor._make(torch.randn(sizes[1], 3), E.CoordinateMapKey([1, 1, 1]), mgr)
            out, feat = net(x)
            assert out.F.shape == (sizes[1], 200) and feat.F.shape == (sizes[1], 96)
            loss = out.F.float().sum() * 0.0 + 1.0 * out.F.float().mean()
            opt.zero_grad(set_to_none=True)
            loss.backward()
            missing = [k for k, p in net.named_parameters() if p.grad is None or p.grad.shape != p.shape]
            assert not missing, missing
            opt.step()
        print(f"fuse_conv_bn={flags[0]} overlap={flags[1]} batch_prep={flags[2]}: ok; C-ABI calls per 2 steps:",
              dict(sorted(stub.calls.items())))
    # eval mode: BatchNorm is not fusable, lazy convs materialise through .F
    E.set_conv_bn_fusion(True)
    net.eval()
    with torch.no_grad():
        out, _ = net(E.SparseTensor._make(torch.randn(sizes[1], 3), E.CoordinateMapKey([1, 1, 1]), FakeManager(sizes)))
    assert out.F.shape == (sizes[1], 200)
    print("eval ok")


if __name__ == "__main__":
    main()
