"""Small end-to-end case for compute-sanitizer: every kernel family once (maps, tc single/multi tile, wgrad, BN, losses)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from languagegroundedsemseg_b200 import minkowski as E, nets, scenes, losses, voxelizer
import numpy as np
for algo in ("tc", "tf32", "simt"):
    E.set_conv_algo(algo)
    torch.manual_seed(42)
    net = nets.build_model("Res16UNet14A", 3, 200, nets.DefaultConfig()).cuda().train()
    c, f, l = scenes.synthetic_voxel_scene(seed=1, target_voxels=22000)   # >= 148 tiles at level 0 -> multi-tile kernel
    out, feat = net(E.SparseTensor(torch.from_numpy(f).cuda(), torch.from_numpy(c).cuda()))
    loss = torch.nn.functional.cross_entropy(out.F, torch.from_numpy(l).cuda(), ignore_index=-1)
    A = torch.randn(200, 96, device="cuda")
    loss = loss + losses.ContrastiveLanguageCELoss(num_labels=200)(feat.F, torch.from_numpy(l).cuda(), A)[0]
    loss.backward()
    torch.cuda.synchronize()
    print(algo, "loss", float(loss))
xyz, _, _ = scenes.synthetic_scene(0, scale=0.3)
M = np.eye(4); M[:3, :3] *= 50.0
coords, uidx, inv = voxelizer.voxelize(torch.from_numpy(xyz).cuda(), M)
torch.cuda.synchronize()
print("voxelize", coords.shape)
