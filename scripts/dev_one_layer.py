"""One conv layer on the config-2 scene (for ncu captures): python scripts/dev_one_layer.py [cin] [cout] [algo] [dtype] [mode] [voxels]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from languagegroundedsemseg_b200 import minkowski as E, _lib, scenes
cin = int(sys.argv[1]) if len(sys.argv) > 1 else 96
cout = int(sys.argv[2]) if len(sys.argv) > 2 else 96
algo = sys.argv[3] if len(sys.argv) > 3 else "tc"
dtype = torch.bfloat16 if len(sys.argv) > 4 and sys.argv[4] == "bf16" else torch.float32
mode = sys.argv[5] if len(sys.argv) > 5 else "fwd"
voxels = int(sys.argv[6]) if len(sys.argv) > 6 else 150000
E.set_conv_algo(algo)
c, _, _ = scenes.synthetic_voxel_scene(0, voxels)
x = E.SparseTensor(torch.zeros(c.shape[0], 1).cuda(), torch.from_numpy(c).cuda())
m, k = x.coordinate_manager, x.coordinate_map_key
km = m.kernel_map(k, k, [3, 3, 3], [1, 1, 1])
torch.manual_seed(0)
f = torch.randn(c.shape[0], cin, device="cuda").to(dtype).requires_grad_(mode != "fwd")
w = (torch.randn(27, cin, cout, device="cuda") / np.sqrt(27 * cin)).requires_grad_(mode != "fwd")
gy = torch.randn(c.shape[0], cout, device="cuda").to(dtype)
for _ in range(3):
    y = E.sparse_conv(f, w, None, km)
    if mode != "fwd":
        y.backward(gy)
torch.cuda.synchronize()
print("done", y.shape)
