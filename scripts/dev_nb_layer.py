"""One neighbourhood-cache convolution (lgs_conv_fwd3) on a synthetic scene, for ncu captures:
python scripts/dev_nb_layer.py [voxels] [c_in] [c_out] [reps]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from languagegroundedsemseg_b200 import minkowski as E, _lib, scenes
voxels = int(sys.argv[1]) if len(sys.argv) > 1 else 450
c_in = int(sys.argv[2]) if len(sys.argv) > 2 else 256
c_out = int(sys.argv[3]) if len(sys.argv) > 3 else 256
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
lib = _lib.load()
for kv in filter(None, os.environ.get("LGS_TUNE", "").split(",")):
    k_, v_ = kv.split("=")
    assert lib.lgs_tune(k_.encode(), int(v_)) == 0
c, _, _ = scenes.synthetic_voxel_scene(11, voxels)
x = E.SparseTensor(torch.zeros(c.shape[0], 1).cuda(), torch.from_numpy(c).cuda())
m, k = x.coordinate_manager, x.coordinate_map_key
km = m.kernel_map(k, k, [3, 3, 3], [1, 1, 1])
assert km.plan is not None, km.plan_stats
n, K = km.n_out, 27
torch.manual_seed(0)
f = torch.randn(n, c_in, device="cuda")
w = torch.randn(K, c_in, c_out, device="cuda") / np.sqrt(K * c_in)
fwd = torch.empty(lib.lgs_weight_bx3_elems(K, c_out, c_in), dtype=torch.bfloat16, device="cuda")
bwd = torch.empty(lib.lgs_weight_bx3_elems(K, c_in, c_out), dtype=torch.bfloat16, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
_lib.check(lib.lgs_weight_prep(_lib.ptr(w), K, c_in, c_out, 3, _lib.ptr(fwd), _lib.ptr(bwd), _lib.F32, st))
out = torch.empty(n, c_out, device="cuda")
for _ in range(reps):
    _lib.check(lib.lgs_conv_fwd3(_lib.ptr(f), c_in, None, 0, n, _lib.ptr(fwd), K, c_out, _lib.ptr(km.fwd_table), _lib.ptr(km.plan), n, 0,
                                 None, _lib.ptr(out), None, st))
torch.cuda.synchronize()
print("done", n, out.abs().mean().item())
