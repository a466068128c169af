"""Development harness (GPU): time one conv layer for several table variants (run with LGS_TC_STAGES=n to sweep)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from languagegroundedsemseg_b200 import minkowski as E, _lib, scenes
lib = _lib.load()
cin = int(sys.argv[1]) if len(sys.argv) > 1 else 96
cout = int(sys.argv[2]) if len(sys.argv) > 2 else 96
c, _, _ = scenes.synthetic_voxel_scene(0, 150000)
x = E.SparseTensor(torch.zeros(c.shape[0], 1).cuda(), torch.from_numpy(c).cuda())
m, k = x.coordinate_manager, x.coordinate_map_key
km = m.kernel_map(k, k, [3, 3, 3], [1, 1, 1])
n = c.shape[0]
torch.manual_seed(0)
f = torch.randn(n, cin, device="cuda")
w = torch.randn(27, cin, cout, device="cuda") / np.sqrt(27 * cin)
out = torch.empty(n, cout, device="cuda")
tables = {"real": km.fwd_table}
t = torch.full_like(km.fwd_table, -1); t[13] = km.fwd_table[13]; tables["centre-only(1 offset)"] = t
t = km.fwd_table.clone(); t[t >= 0] = 0; tables["all-same-row(no misses)"] = t
ident = torch.arange(n, device="cuda", dtype=torch.int32)
t = ident[None, :].repeat(27, 1).contiguous(); tables["identity x27 (sequential rows)"] = t
for algo, nsplit, layout in (("tf32", 1, _lib.W_KNC), ("tc3", 2, _lib.W_KNC_SPLIT)):
    wf = torch.empty(nsplit, 27, cout, cin, device="cuda")
    _lib.check(lib.lgs_weight_prep(_lib.ptr(w), 27, cin, cout, nsplit, _lib.ptr(wf), None, 0, E._stream()))
    a = _lib.ALGO_TC if algo == "tf32" else _lib.ALGO_TC3
    for name, tab in tables.items():
        def run():
            _lib.check(lib.lgs_conv_fwd(_lib.ptr(f), n, cin, _lib.ptr(wf), layout, 27, cout, _lib.ptr(tab), n, 0, None, _lib.ptr(out), 0, a, E._stream()))
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        print(f"stages={os.environ.get('LGS_TC_STAGES','max')} {algo:5s} {name:32s} {e0.elapsed_time(e1)/10:7.3f} ms", flush=True)
