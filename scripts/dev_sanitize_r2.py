"""Small end-to-end case for compute-sanitizer, round-2 kernels: one native-driver training step (Res16UNet14A, bf16x3) on a
3 K-voxel scene — plan builder, conv_nb (incl. the reduction split on the coarse levels), conv_bx3, tcgen05 wgrad, the stem
wgrad, zero-fill kernels, fused BatchNorm, persistent seg_ce, vectorised colsum."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from languagegroundedsemseg_b200 import minkowski as E, nets, scenes
from languagegroundedsemseg_b200.program import NativeStep
E.set_conv_algo("bx3")
torch.manual_seed(42)
net = nets.build_model("Res16UNet14A", 3, 200, nets.DefaultConfig()).cuda().train()
c, f, l = scenes.synthetic_voxel_scene(seed=1, target_voxels=3000)
step = NativeStep(net)
loss = step.run(E.SparseTensor(torch.from_numpy(f).cuda(), torch.from_numpy(c).cuda()), torch.from_numpy(l).cuda())
torch.cuda.synchronize()
print("bx3 native step loss", float(loss))
