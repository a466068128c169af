"""2+ ranks (torchrun): after one NativeStep.run every rank must hold the SAME flat gradient buffer (each bucket reduced once,
after all of its gradients were written), equal to the mean of the ranks' local gradients (computed with a world-1 reducer).
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dev_ddp_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from languagegroundedsemseg_b200 import ddp, minkowski as E, nets, scenes
from languagegroundedsemseg_b200.program import NativeStep

rank, world, local = ddp.init_process_group("nccl")
torch.cuda.set_device(local)
E.set_conv_algo("bx3")
coords, feats, labels = scenes.synthetic_voxel_scene(seed=20 + rank, target_voxels=40000)
c, f, lab = (torch.from_numpy(a).cuda() for a in (coords, feats, labels))


def one_step(distributed):
    torch.manual_seed(7)
    net = nets.build_model("Res16UNet34C", 3, 200, nets.DefaultConfig()).cuda().train()
    red = ddp.GradAllReducer(net.parameters(), overlap=False)
    if not distributed:
        red.world = 1
    step = NativeStep(net, reducer=red)
    loss = step.run(E.SparseTensor(f, c), lab)
    torch.cuda.synchronize()
    return red.flat.clone(), float(loss)


g_dist, _ = one_step(True)
g_local, loss = one_step(False)
mean_local = g_local.clone()
dist.all_reduce(mean_local, op=dist.ReduceOp.AVG)
other = g_dist.clone()
dist.broadcast(other, src=0)
same = torch.equal(other, g_dist)
err = ((g_dist - mean_local).norm() / mean_local.norm()).item()
print(f"[rank {rank}] loss {loss:.4f}  flat gradients identical to rank 0: {same};  |reduced - mean(local)| / |mean(local)| = {err:.2e}", flush=True)
ok = torch.tensor([1 if (same and err < 5e-2) else 0], device="cuda")
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(ok.item()) == 1 else 1)
