"""BASELINE configs 3 and 5 through the engine: Res16UNet34CR_Proj + CLIP CE loss; Res16UNet34D on a ~600K-voxel 1 cm scene."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from languagegroundedsemseg_b200 import minkowski as E, nets, scenes, losses
cfg = nets.DefaultConfig()
def run(name, target, voxel, steps=3, clip=False):
    c, f, l = scenes.synthetic_voxel_scene(0, target, voxel_size=voxel)
    dc, df, dl = torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda(), torch.from_numpy(l).cuda()
    torch.manual_seed(42)
    net = nets.build_model(name, 3, 200, cfg).cuda().train()
    if clip:
        net.representation_only(True)
        torch.manual_seed(1)
        anchors = torch.nn.functional.normalize(torch.randn(200, 512, device="cuda"), dim=1)
        crit = losses.ContrastiveLanguageCELoss(num_labels=200)
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, fused=True)
    def step():
        st = E.SparseTensor(df, dc)
        if clip and name.endswith("Proj"):
            feat, anc = net(st, anchors)
            loss = crit(feat.F, dl, anc)[0]
        elif clip:
            feat = net(st)
            loss = crit(feat.F, dl, anchors)[0]
        else:
            out, _ = net(st)
            loss = torch.nn.functional.cross_entropy(out.F, dl, ignore_index=-1)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss
    for _ in range(2): step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps): loss = step()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    print(f"{name:20s} voxels={c.shape[0]:7d} clip={clip} loss={loss.item():.4f} {ms:8.2f} ms/step {c.shape[0]/ms/1e3:6.2f} M voxels/s  peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
run("Res16UNet34CR_Proj", 150000, 0.02, clip=True)
run("Res16UNet34D", 150000, 0.02, clip=True)
run("Res16UNet34D", 600000, 0.01, clip=True)
run("Res16UNet14A", 150000, 0.02)
