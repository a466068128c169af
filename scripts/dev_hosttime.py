"""Where does the host spend a training step?  bench.py's prefetched loop with perf_counter around each phase (no syncs
added), next to the CUDA-event step time: host-bound if the host phases add up to the step time."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from languagegroundedsemseg_b200 import minkowski as E
from languagegroundedsemseg_b200.prefetch import SparseBatchPrefetcher

c, f, l = bench.make_scene(0)
dev = torch.device("cuda", 0)
dc, df, dl = torch.from_numpy(c).to(dev), torch.from_numpy(f).to(dev), torch.from_numpy(l).to(dev)
net, opt = bench.build_net(None, dev, torch.float32)
pf = SparseBatchPrefetcher(dev, torch.float32, threaded=os.environ.get("LGS_STAGE_THREAD", "0") != "0",
                           high_priority=os.environ.get("LGS_STAGE_PRIORITY", "0") != "0")
ticket = pf.stage(dc, df, dl)
acc = {}


def step(record):
    global ticket
    t = [time.perf_counter()]
    st, lab = pf.get(ticket)
    ticket = pf.stage(dc, df, dl)
    t.append(time.perf_counter())
    out, _ = net(st)
    loss = torch.nn.functional.cross_entropy(out.F.float(), lab, ignore_index=-1)
    t.append(time.perf_counter())
    opt.zero_grad(set_to_none=True)
    loss.backward()
    t.append(time.perf_counter())
    opt.step()
    t.append(time.perf_counter())
    if record:
        for name, a, b in zip(("stage_next_batch", "forward+loss", "backward", "optimizer"), t, t[1:]):
            acc[name] = acc.get(name, 0.0) + (b - a)


for _ in range(8):
    step(False)
torch.cuda.synchronize()
N = 30
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
w0 = time.perf_counter()
e0.record()
for _ in range(N):
    step(True)
e1.record()
w1 = time.perf_counter()
torch.cuda.synchronize()
print(f"step (CUDA events) {e0.elapsed_time(e1) / N:.2f} ms; host loop wall {1e3 * (w1 - w0) / N:.2f} ms/step")
for k, v in acc.items():
    print(f"  host {k:18s} {1e3 * v / N:6.2f} ms/step")
print(f"  host total          {1e3 * sum(acc.values()) / N:6.2f} ms/step")
