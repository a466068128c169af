"""torch.profiler view of one training step (warm): kernel-level aggregate + step wall time + GPU idle share."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from languagegroundedsemseg_b200 import minkowski as E
from torch.profiler import profile, ProfilerActivity
c, f, l = bench.make_scene(0)
dev = "cuda"
dc, df, dl = torch.from_numpy(c).to(dev), torch.from_numpy(f).to(dev), torch.from_numpy(l).to(dev)
net, opt = bench.build_net(None, dev, torch.float32)
for _ in range(5):
    bench.train_step(E.SparseTensor, net, opt, dc, df, dl)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        bench.train_step(E.SparseTensor, net, opt, dc, df, dl)
    torch.cuda.synchronize()
ev = prof.key_averages()
rows = [(e.key, e.device_time_total / 3e3, e.count // 3) for e in ev if e.device_time_total > 0 and e.device_type.name != "CPU" or (e.device_time_total > 0 and "Memcpy" in e.key)]
seen = {}
for e in ev:
    if getattr(e, "device_type", None) is not None and e.device_type.name == "CUDA":
        seen[e.key] = (e.device_time_total / 3e3, e.count // 3)
tot = sum(v[0] for v in seen.values())
print(f"GPU kernel time per step: {tot:.2f} ms (sum of kernel durations; equals GPU-busy time only with LGS_OVERLAP_WGRAD=0)")
for k, (t, n) in sorted(seen.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"{t:8.3f} ms {100*t/tot:5.1f}% x{n:4d}  {k[:110]}")
