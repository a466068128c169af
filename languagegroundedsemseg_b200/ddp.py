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
    """The path's only collective: the gradients of all parameters live in ONE flat buffer (every `p.grad` is a view of it,
    so there is no concat before and no scatter after the collective), all-reduced (mean) over NCCL in a few contiguous
    buckets that are launched WHILE backward is still running: a bucket goes out as soon as the last of its parameters has
    its gradient (post-accumulate hooks count them down; backward produces the decoder's gradients first, so the buckets
    are cut in parameter order and complete from the tail).  NCCL runs them on its own stream next to the remaining
    backward kernels; `wait()` (or calling the reducer) joins them before the optimiser step.
    Replaces DistributedDataParallel's per-parameter bucketing (~4.5 ms of host time per step for the 374 parameter tensors
    of Res16UNet34C) and round 1's cat -> all-reduce -> scatter (3.5 ms per step, not overlapped).  (main.py:192-195)

    Contract: parameters start identical on every rank (same seed, or `broadcast_parameters`); gradients are zeroed with
    `zero_grad()` of this object (one memset) or `optimizer.zero_grad(set_to_none=False)` — `set_to_none=True` would detach
    the views, which `zero_grad()`/`attach()` repair.  Parameters that received no gradient in a step contribute zeros.
    BatchNorm running statistics stay rank-local during training; `sync_buffers()` averages them over the
    ranks — call it before saving a checkpoint or evaluating."""

    def __init__(self, params, buckets=3, module=None, overlap=True):
        self.params = [p for p in params if p.requires_grad]
        self.sizes = [p.numel() for p in self.params]
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.module = module
        self.overlap = overlap
        total = sum(self.sizes)
        p0 = self.params[0]
        self.flat = torch.zeros(total, dtype=p0.dtype, device=p0.device)
        self.offsets = [0]
        for n in self.sizes:
            self.offsets.append(self.offsets[-1] + n)
        # buckets: contiguous parameter ranges of roughly equal bytes
        buckets = max(1, min(int(buckets), len(self.params)))
        self.bounds, acc, target, k = [0], 0, total / buckets, 1
        for i, n in enumerate(self.sizes):
            acc += n
            if acc >= k * target and len(self.bounds) < buckets and i + 1 < len(self.params):
                self.bounds.append(i + 1)
                k += 1
        self.bounds.append(len(self.params))
        self.bucket_of = []
        for b in range(len(self.bounds) - 1):
            self.bucket_of += [b] * (self.bounds[b + 1] - self.bounds[b])
        self._pending = [0] * (len(self.bounds) - 1)
        self._works = []
        self._fired = [False] * (len(self.bounds) - 1)
        self.attach()
        self._hooks = []
        if self.world > 1 and overlap:
            for i, p in enumerate(self.params):
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(self.bucket_of[i])))
        self._arm()

    def set_bounds(self, bounds):
        """re-cut the buckets at the given parameter indices (ascending, first 0, last len(params)).  The native step driver
        aligns them with the points of backward at which whole groups of layers have their gradients, so that a bucket can go
        out the moment its group is done.  Not with per-parameter hooks (overlap=True: they captured the old bucket ids)."""
        bounds = sorted(set(int(b) for b in bounds))
        if self._hooks:
            raise RuntimeError("set_bounds: this reducer launches its buckets from gradient hooks")
        if bounds[0] != 0 or bounds[-1] != len(self.params):
            raise ValueError("set_bounds: bounds must start at 0 and end at len(params)")
        self.bounds = bounds
        self.bucket_of = []
        for b in range(len(bounds) - 1):
            self.bucket_of += [b] * (bounds[b + 1] - bounds[b])
        self._pending = [0] * (len(bounds) - 1)
        self._fired = [False] * (len(bounds) - 1)
        self._arm()

    # -- gradient views ---------------------------------------------------------------------------------------------
    def attach(self):
        """(re)point every parameter's .grad at its slice of the flat buffer"""
        for p, o, n in zip(self.params, self.offsets, self.sizes):
            p.grad = self.flat[o:o + n].view_as(p)

    def zero_grad(self):
        self.flat.zero_()
        for p, o, n in zip(self.params, self.offsets, self.sizes):
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + o * self.flat.element_size():
                p.grad = self.flat[o:o + n].view_as(p)
        self._arm()

    def _arm(self):
        for b in range(len(self._pending)):
            self._pending[b] = self.bounds[b + 1] - self.bounds[b]
            self._fired[b] = False
        self._works = []

    # -- the collective ---------------------------------------------------------------------------------------------
    def bucket_slice(self, b):
        return self.flat[self.offsets[self.bounds[b]]:self.offsets[self.bounds[b + 1]]]

    def reduce_bucket(self, b):
        """launch the all-reduce of bucket b now (the explicit step program calls this at its bucket boundaries)"""
        if self.world == 1 or self._fired[b]:
            return
        self._fired[b] = True
        t = self.bucket_slice(b)
        if t.is_cuda:
            self._works.append(dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=True))
        else:                                   # gloo (CPU tests): no AVG, synchronous
            dist.all_reduce(t)
            t.div_(self.world)

    def _make_hook(self, b):
        def hook(_p):
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self.reduce_bucket(b)
        return hook

    def wait(self):
        """all buckets reduced (those whose parameters never got a gradient this step are sent now); the current stream
        waits for the collectives"""
        if self.world == 1:
            return
        for b in range(len(self._pending)):
            self.reduce_bucket(b)
        for w in self._works:
            w.wait()
        self._works = []

    __call__ = wait

    # -- parameters / buffers ---------------------------------------------------------------------------------------
    def broadcast_parameters(self, src=0):
        if self.world > 1:
            for p in self.params:
                dist.broadcast(p.data, src)
            from . import minkowski
            minkowski.invalidate_weight_cache()     # `.data` writes do not bump the parameters' version counters

    def sync_buffers(self, module=None):
        """average the floating-point buffers (BatchNorm running_mean / running_var) over the ranks and take rank 0's
        integer buffers (num_batches_tracked): what a rank-0 checkpoint should contain when BatchNorm statistics were
        not synchronised during training (the reference syncs them every step, main.py:122-123)"""
        module = module if module is not None else self.module
        if module is None or self.world == 1:
            return
        for b in module.buffers():
            if b.is_floating_point():
                if b.is_cuda:
                    dist.all_reduce(b, op=dist.ReduceOp.AVG)
                else:
                    dist.all_reduce(b)
                    b.div_(self.world)
            else:
                dist.broadcast(b, 0)
