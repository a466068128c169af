"""Development harness (GPU): whole-network error of each conv algo vs golden / oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from languagegroundedsemseg_b200 import minkowski as E, nets, scenes
from oracle import me_cpu
from tests.helpers import dense_cube

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
g = np.load(os.path.join(G, "unet14a_cube.npz"))
for algo in sys.argv[1:] or ["simt", "tc"]:
    E.set_conv_algo(algo)
    torch.manual_seed(42)
    net = nets.build_model("Res16UNet14A", 3, 200, nets.DefaultConfig()).cuda().train()
    torch.manual_seed(0)
    f = torch.rand(8000, 3) - 0.5
    with torch.no_grad():
        out, feat = net(E.SparseTensor(f.cuda(), torch.from_numpy(dense_cube(20)).cuda()))
    scale = np.abs(g["logits_rows"]).max()
    e = np.abs(out.F[::125].cpu().numpy() - g["logits_rows"])
    print(f"[{algo}] 14A cube logits: max err/scale {e.max()/scale:.2e}  rms err/rms {np.sqrt((e**2).mean())/np.sqrt((g['logits_rows']**2).mean()):.2e}")

for nvox in (3000, 20000):
    coords, feats, labels = scenes.synthetic_voxel_scene(seed=5, target_voxels=nvox)
    ref = None
    for algo in ["oracle"] + (sys.argv[1:] or ["simt", "tc"]):
        eng, dev = (me_cpu, "cpu") if algo == "oracle" else (None, "cuda")
        if eng is None:
            E.set_conv_algo(algo)
        torch.manual_seed(42)
        net = nets.build_model("Res16UNet34C", 3, 200, nets.DefaultConfig(), engine=eng).to(dev).train()
        ST = (eng or E).SparseTensor
        out, _ = net(ST(torch.from_numpy(feats).to(dev), torch.from_numpy(coords).to(dev)))
        loss = torch.nn.functional.cross_entropy(out.F, torch.from_numpy(labels).to(dev), ignore_index=-1)
        loss.backward()
        res = (loss.item(), out.F.detach().cpu(), {k: p.grad.cpu() for k, p in net.named_parameters()})
        if ref is None:
            ref = res
            sizes = [v.shape[0] for v in ST.__dict__.get('x', [])] if False else None
            continue
        lo = ((res[1] - ref[1]).abs().max() / ref[1].abs().max()).item()
        errs = {k: ((res[2][k] - v).norm() / v.norm().clamp(min=1e-20)).item() for k, v in ref[2].items()}
        errm = {k: ((res[2][k] - v).abs().max() / v.abs().max().clamp(min=1e-20)).item() for k, v in ref[2].items()}
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
        print(f"[{algo}] 34C n={coords.shape[0]} loss {res[0]:.6f} vs {ref[0]:.6f}  logits max err/scale {lo:.2e}  "
              f"grad L2 rel err: median {np.median(list(errs.values())):.2e} worst {[(k, f'{v:.1e}', f'max-norm {errm[k]:.1e}', f'|g|={ref[2][k].norm():.1e}') for k, v in worst]}")
