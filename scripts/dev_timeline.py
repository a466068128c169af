"""GPU timeline of the native-driver training step from torch.profiler (CUPTI): kernel-busy time per stream, union busy time,
idle gaps between consecutive kernels of the training stream, host time of NativeStep.run.
usage: python scripts/dev_timeline.py [out.txt]"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from languagegroundedsemseg_b200 import minkowski as E, ddp
from languagegroundedsemseg_b200.program import NativeStep
from torch.profiler import profile, ProfilerActivity

out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
E.set_conv_algo("bx3")
c, f, l = bench.make_scene(0)
dev = "cuda"
dc, df, dl = torch.from_numpy(c).to(dev), torch.from_numpy(f).to(dev), torch.from_numpy(l).to(dev)
net, opt = bench.build_net(None, dev, torch.float32)
reducer = ddp.GradAllReducer(net.parameters(), overlap=False)
native = NativeStep(net, ignore_index=-1, reducer=reducer, head=None)
sts = [E.SparseTensor(df, dc) for _ in range(9)]            # coordinate / kernel maps + plans built ahead (prefetch.py does this on a side stream)
for i in range(5):
    bench.train_step(None, net, opt, None, None, dl, reducer, st=sts[i], native=native)
torch.cuda.synchronize()
host = []
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        t0 = time.perf_counter()
        bench.train_step(None, net, opt, None, None, dl, reducer, st=sts[5 + i], native=native)
        host.append(time.perf_counter() - t0)
        torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "t.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
# the last step: kernels after the last big idle gap (>200 us) that follows a synchronize
starts = [0] + [i for i in range(1, len(ev)) if ev[i]["ts"] - (ev[i - 1]["ts"] + ev[i - 1]["dur"]) > 300]
step = ev[starts[-1]:]
t0, t1 = step[0]["ts"], max(e["ts"] + e["dur"] for e in step)
streams = {}
for e in step:
    streams.setdefault(e["args"].get("stream"), []).append(e)
print(f"step span {1e-3 * (t1 - t0):.3f} ms, {len(step)} kernels, host issue time of the step {1e3 * host[-1]:.2f} ms (train_step call, no sync)", file=out)
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in step)
busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
for s, e_ in iv[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        cur_s, cur_e = s, e_
    else:
        cur_e = max(cur_e, e_)
busy += cur_e - cur_s
print(f"GPU busy (union over streams) {1e-3 * busy:.3f} ms = {100 * busy / (t1 - t0):.1f} % of the span; idle {1e-3 * (t1 - t0 - busy):.3f} ms", file=out)
for sid, es in sorted(streams.items(), key=lambda kv: -len(kv[1])):
    tot = sum(e["dur"] for e in es)
    gaps = [es[i + 1]["ts"] - (es[i]["ts"] + es[i]["dur"]) for i in range(len(es) - 1)]
    small = [g for g in gaps if 0 <= g < 50]
    print(f"stream {sid}: {len(es)} kernels, {1e-3 * tot:.3f} ms of kernels; gaps to the next kernel of the same stream: "
          f"median {sorted(gaps)[len(gaps) // 2] if gaps else 0:.1f} us, sum of gaps < 50 us {1e-3 * sum(small):.3f} ms ({len(small)} gaps)", file=out)
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as tr:        # compact trace of the step: start us, duration us, stream, grid, kernel
        for e in step:
            tr.write(f"{e['ts'] - t0:10.1f} {e['dur']:8.1f} {e['args'].get('stream')} {e['args'].get('grid')} {e['name'].split('(')[0].replace('void ', '')[:60]}\n")
agg = {}
for e in step:
    n = e["name"].split("(")[0].replace("void ", "")
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += e["dur"]
print("kernel durations in situ (warm L2, concurrent streams):", file=out)
for n, (k, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"  {1e-3 * d:7.3f} ms x{k:4d}  {n[:90]}", file=out)
