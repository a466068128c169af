"""CPU dry run of the facade's host logic (no GPU, no arithmetic): the C-ABI library is replaced by a stub whose entry
points return LGS_OK, coordinate/kernel maps by tables of the right shapes.  Exercises the autograd plumbing of the
lazy conv / fused conv+BN node / batched weight prep / side-stream wgrad paths (argument counts, saved tensors,
gradient arity and shapes) before a GPU box is spent on them.  `python scripts/dev_dryrun_cpu.py`"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from languagegroundedsemseg_b200 import _lib, minkowski as E, nets


class StubLib:
    calls = {}

    def __getattr__(self, name):
        def f(*a):
            StubLib.calls[name] = StubLib.calls.get(name, 0) + 1
            return 1 if name.endswith("_supported") else 0
        return f


class FakeEvent:
    def record(self, s=None):
        pass

    def wait(self, s=None):
        pass


class FakeStream:
    cuda_stream = 0

    def wait_event(self, e):
        pass


class FakeManager:
    D = 3

    def __init__(self, sizes):
        self.sizes = sizes          # rows per tensor stride 1,2,4,8,16
        self.cache = {}

    def conv_maps(self, in_key, ks, stride, dil, transpose):
        ts = in_key.tensor_stride[0]
        out_ts = ts // stride if transpose else ts * stride
        ck = (ts, out_ts, ks)
        if ck not in self.cache:
            km = E.KernelMap()
            km.K, km.n_in, km.n_out = ks ** 3, self.sizes[ts], self.sizes[out_ts]
            km.fwd_table = torch.zeros((km.K, km.n_out), dtype=torch.int32)
            km.bwd_table = torch.zeros((km.K, km.n_in), dtype=torch.int32)
            km.bwd_reverse, km.counts = ts == out_ts, torch.zeros(km.K, dtype=torch.int32)
            self.cache[ck] = (E.CoordinateMapKey([out_ts] * 3), km)
        return self.cache[ck]


def main():
    stub = StubLib()
    _lib.load = lambda: stub
    E._stream = lambda: None
    E._scratch64 = lambda idx: torch.empty(16 * 1024, dtype=torch.float64)
    E._side_stream = lambda idx: (FakeStream(), FakeEvent(), FakeEvent())
    E._cur_stream_obj = lambda idx: FakeStream()
    sizes = {1: 1000, 2: 300, 4: 100, 8: 30, 16: 10}
    for flags in ((True, True, True), (False, False, False), (True, False, True), (False, True, False)):
        E.set_conv_bn_fusion(flags[0]), E.set_wgrad_overlap(flags[1]), E.set_batched_weight_prep(flags[2])
        StubLib.calls.clear()
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
              dict(sorted(StubLib.calls.items())))
    # eval mode: BatchNorm is not fusable, lazy convs materialise through .F
    E.set_conv_bn_fusion(True)
    net.eval()
    with torch.no_grad():
        out, _ = net(E.SparseTensor._make(torch.randn(sizes[1], 3), E.CoordinateMapKey([1, 1, 1]), FakeManager(sizes)))
    assert out.F.shape == (sizes[1], 200)
    print("eval ok")


if __name__ == "__main__":
    main()
