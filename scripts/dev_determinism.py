"""Is the eval forward bit-reproducible?  Same SparseTensor twice, inline vs prefetched construction, with the host fast
paths on and off (tests/test_gpu_nets.py::test_prefetcher_matches_inline_construction asserts torch.equal)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from languagegroundedsemseg_b200 import minkowski as E, nets, scenes

E.set_conv_algo("tc")
torch.manual_seed(42)
net = nets.build_model("Res16UNet14A", 3, 200, nets.DefaultConfig()).cuda().eval()
for fast in (True, False):
    E.set_conv_bn_fusion(fast), E.set_batched_weight_prep(fast)
    for seed, target in ((0, 4000), (1, 5500), (2, 7000), (3, 8500), (4, 40000)):
        c, f, _ = scenes.synthetic_voxel_scene(seed=seed, target_voxels=target)
        with torch.no_grad():
            st = E.SparseTensor(torch.from_numpy(f).cuda(), torch.from_numpy(c).cuda())
            outs = [net(st)[0].F.clone() for _ in range(4)]
            st2 = E.SparseTensor(torch.from_numpy(f).cuda(), torch.from_numpy(c).cuda())
            outs.append(net(st2)[0].F.clone())
        d = [(o - outs[0]).abs().max().item() for o in outs[1:]]
        print(f"fast={fast} voxels={c.shape[0]:6d} max|diff| vs first run: {d} scale {outs[0].abs().max().item():.3f}", flush=True)
# one layer at a time: which convolution shapes are not reproducible?
c, f, _ = scenes.synthetic_voxel_scene(seed=0, target_voxels=4000)
st = E.SparseTensor(torch.from_numpy(f).cuda(), torch.from_numpy(c).cuda())
mgr, key = st.coordinate_manager, st.coordinate_map_key
lvl = 0
with torch.no_grad():
    while True:
        n = mgr.size(key)
        for ch in (32, 64, 128, 256):
            torch.manual_seed(ch)
            conv = E.MinkowskiConvolution(ch, ch, kernel_size=3, dimension=3).cuda()
            x = E.SparseTensor(torch.randn(n, ch).cuda(), coordinate_map_key=key, coordinate_manager=mgr)
            a = conv(x).F.clone()
            diffs = [(conv(x).F - a).abs().max().item() for _ in range(5)]
            print(f"level {lvl} rows {n:5d} {ch:3d}->{ch:3d}: max diff over 5 repeats {max(diffs):.3e}", flush=True)
        if lvl == 4:
            break
        key = mgr.stride(key, 2)
        lvl += 1
