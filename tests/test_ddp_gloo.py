"""N>1 host logic on CPU: world_size-2 gloo, one scene per rank, DDP gradient all-reduce == mean of per-rank gradients."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_scenes_partition():
    from languagegroundedsemseg_b200.ddp import shard_scenes
    for n, w in [(8, 8), (8, 2), (7, 4), (3, 8), (0, 2)]:
        parts = [shard_scenes(n, w, r) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def _scene_grads(rank, engine):
    from languagegroundedsemseg_b200 import nets, scenes
    torch.manual_seed(42)
    net = nets.build_model("Res16UNet14A", 3, 20, nets.DefaultConfig(), engine=engine).train()
    coords, feats, labels = scenes.synthetic_voxel_scene(seed=10 + rank, target_voxels=600)
    labels = labels % 20
    return net, coords, feats, labels


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    from languagegroundedsemseg_b200 import ddp
    from oracle import me_cpu
    r, w, _ = ddp.init_process_group("gloo")
    assert (r, w) == (rank, world)
    net, coords, feats, labels = _scene_grads(rank, me_cpu)
    if os.environ.get("LGS_TEST_USE_DDP"):
        model = ddp.wrap_ddp(net)                       # stock DistributedDataParallel
        reducer = None
    elif os.environ.get("LGS_TEST_STAGED"):
        # the native step driver's protocol: no hooks, buckets re-cut at layer-group boundaries, sent group by group
        model, reducer = net, ddp.GradAllReducer(net.parameters(), overlap=False)
        n = len(reducer.params)
        reducer.set_bounds([0, n // 5, n // 2, n - 7, n])
    else:
        model, reducer = net, ddp.GradAllReducer(net.parameters())   # bench.py's path: one flat all-reduce
    out, _ = model(me_cpu.SparseTensor(torch.from_numpy(feats), torch.from_numpy(coords)))
    loss = torch.nn.functional.cross_entropy(out.F, torch.from_numpy(labels), ignore_index=-1)
    loss.backward()
    if reducer is not None and os.environ.get("LGS_TEST_STAGED"):
        for first in (n - 7, n // 2):                   # "decoder done", "stage 4 done": buckets at or above `first` go out
            for b in range(len(reducer._pending)):
                if reducer.bounds[b] >= first:
                    reducer.reduce_bucket(b)
        reducer.wait()                                  # the rest
    elif reducer is not None:
        reducer()
    torch.save({k: p.grad.clone() for k, p in net.named_parameters()}, os.path.join(out_dir, f"g{rank}.pt"))
    n = torch.tensor([coords.shape[0]], dtype=torch.float64)
    dist.all_reduce(n)                                  # bench.py's voxel total over ranks
    torch.save(n, os.path.join(out_dir, f"n{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("use_ddp", [False, True, "staged"])
def test_ddp_gloo_two_ranks(tmp_path, use_ddp, monkeypatch):
    monkeypatch.delenv("LGS_TEST_USE_DDP", raising=False)
    monkeypatch.delenv("LGS_TEST_STAGED", raising=False)
    if use_ddp == "staged":
        monkeypatch.setenv("LGS_TEST_STAGED", "1")
    elif use_ddp:
        monkeypatch.setenv("LGS_TEST_USE_DDP", "1")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g0, g1 = torch.load(tmp_path / "g0.pt"), torch.load(tmp_path / "g1.pt")
    # after DDP's all-reduce both ranks hold the same (averaged) gradients
    for k in g0:
        assert torch.equal(g0[k], g1[k]), k
    # ... equal to the mean of the two scenes' gradients computed serially
    from oracle import me_cpu
    ref, total = None, 0
    for rank in range(2):
        net, coords, feats, labels = _scene_grads(rank, me_cpu)
        out, _ = net(me_cpu.SparseTensor(torch.from_numpy(feats), torch.from_numpy(coords)))
        torch.nn.functional.cross_entropy(out.F, torch.from_numpy(labels), ignore_index=-1).backward()
        g = {k: p.grad for k, p in net.named_parameters()}
        ref = g if ref is None else {k: ref[k] + g[k] for k in g}
        total += coords.shape[0]
    for k in g0:
        np.testing.assert_allclose(g0[k].numpy(), (ref[k] / 2).numpy(), rtol=2e-4, atol=1e-6, err_msg=k)
    assert torch.load(tmp_path / "n0.pt").item() == total == torch.load(tmp_path / "n1.pt").item()
