"""cProfile of the host side of the training step (GPU box)."""
import sys, os, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from languagegroundedsemseg_b200 import minkowski as E
c, f, l = bench.make_scene(0)
dev = "cuda"
dc, df, dl = torch.from_numpy(c).to(dev), torch.from_numpy(f).to(dev), torch.from_numpy(l).to(dev)
net, opt = bench.build_net(None, dev, torch.float32)
for _ in range(5):
    bench.train_step(E.SparseTensor, net, opt, dc, df, dl)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    bench.train_step(E.SparseTensor, net, opt, dc, df, dl)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host issue time/step {1e3*(t1-t0)/10:.2f} ms; incl. drain {1e3*(t2-t0)/10:.2f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    bench.train_step(E.SparseTensor, net, opt, dc, df, dl)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
st.sort_stats("cumtime").print_stats(30)
