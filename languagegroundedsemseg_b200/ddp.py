"""Scene-parallel training plumbing: one process per GPU, whole scenes per rank, DDP's bucketed
gradient all-reduce as the only collective (main.py:192-195 of the reference uses PL's DDPPlugin the same way).
BatchNorm statistics are deliberately NOT synchronised."""
import os

import torch
import torch.distributed as dist


def shard_scenes(n_scenes: int, world_size: int, rank: int):
    """Scene ids owned by `rank`: contiguous, balanced to within one scene, covering range(n_scenes) exactly once."""
    base, extra = divmod(n_scenes, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def init_process_group(backend=None):
    """Rendezvous from torchrun's env (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).  -> (rank, world, local_rank)"""
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return rank, world, local_rank


def wrap_ddp(net, local_rank=None):
    """DistributedDataParallel over the whole network; every parameter is used every step (static graph)."""
    from torch.nn.parallel import DistributedDataParallel as DDP
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return net
    if next(net.parameters()).is_cuda:
        return DDP(net, device_ids=[local_rank], find_unused_parameters=False, gradient_as_bucket_view=True)
    return DDP(net, find_unused_parameters=False)


class GradAllReducer:
    """The path's only collective, done once per step on ONE flat buffer: concat the gradients, NCCL all-reduce (mean),
    scatter back.  Replaces DistributedDataParallel's per-parameter autograd hooks and bucketing, whose host cost
    (~4.5 ms per step for the 374 parameter tensors of Res16UNet34C) lands on the critical path of a host-bound step;
    the 151 MB all-reduce itself takes ~0.4 ms over NVLink/NVSwitch.  Parameters must start identical on every rank
    (same seed, or call `broadcast_parameters`)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.sizes = [p.numel() for p in self.params]
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1

    def broadcast_parameters(self, src=0):
        if self.world > 1:
            for p in self.params:
                dist.broadcast(p.data, src)
            from . import minkowski
            minkowski.invalidate_weight_cache()     # `.data` writes do not bump the parameters' version counters

    def __call__(self):
        if self.world == 1:
            return
        grads = [p.grad.view(-1) for p in self.params]
        flat = torch.cat(grads)
        dist.all_reduce(flat, op=dist.ReduceOp.AVG if flat.is_cuda else dist.ReduceOp.SUM)
        if not flat.is_cuda:
            flat.div_(self.world)          # gloo has no AVG
        torch._foreach_copy_(grads, list(flat.split(self.sizes)))
