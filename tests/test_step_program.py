"""CPU: the explicit training-step program (languagegroundedsemseg_b200/step.py).
 (1) data flow: run on the CPU oracle engine (tests/oracle_backend.py) it gives the loss, logits and EVERY parameter
     gradient that autograd gives for the same oracle network, and the same BatchNorm running statistics;
 (2) call sequence: run on the facade backend with the C library recording its calls (lgs_trace_begin, no GPU) it issues
     exactly the calls the module-by-module facade + autograd issue for the same step."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from languagegroundedsemseg_b200 import nets, scenes, step
from oracle import me_cpu
from tests.oracle_backend import OracleBackend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name,voxels", [("Res16UNet14A", 1500), ("Res16UNet34C", 900)])
def test_program_gradients_equal_autograd_on_the_oracle_engine(name, voxels):
    coords, feats, labels = scenes.synthetic_voxel_scene(seed=4, target_voxels=voxels)
    c, f, lab = torch.from_numpy(coords), torch.from_numpy(feats), torch.from_numpy(labels)
    res = {}
    for mode in ("autograd", "program"):
        torch.manual_seed(42)
        net = nets.build_model(name, 3, 200, nets.DefaultConfig(), engine=me_cpu).train()
        st = me_cpu.SparseTensor(f, c)
        if mode == "autograd":
            out, _ = net(st)
            loss = torch.nn.functional.cross_entropy(out.F, lab.long(), ignore_index=-1)
            loss.backward()
            logits = out.F.detach()
        else:
            prog = step.StepProgram(net, OracleBackend())
            with torch.no_grad():
                loss = prog.run(st, lab, ignore_index=-1)
            logits = prog.logits
        res[mode] = (loss.item(), logits, {k: p.grad.clone() for k, p in net.named_parameters()},
                     {k: v.clone() for k, v in net.state_dict().items() if "running" in k or "tracked" in k})
    a, p = res["autograd"], res["program"]
    assert abs(a[0] - p[0]) < 1e-6 * abs(a[0])
    assert torch.allclose(a[1], p[1], rtol=1e-5, atol=1e-6)
    assert a[2].keys() == p[2].keys() and len(a[2]) > 90
    worst = max(((p[2][k] - g).norm() / g.norm().clamp(min=1e-20)).item() for k, g in a[2].items())
    assert worst < 1e-4, worst                       # same arithmetic, node-local autograd: fp32 round-off only
    for k, v in a[3].items():
        assert torch.allclose(p[3][k].float(), v.float(), rtol=1e-5, atol=1e-7), k


@pytest.mark.parametrize("name", ["Res16UNet34CR_Proj", "Res16UNet34D"])
def test_program_with_clip_head_equals_autograd_on_the_oracle_engine(name):
    """BASELINE configs 3 / 5 in small: the CLIP pre-training nets (representation_only) with the text-anchor
    cross-entropy as the head of `run_features` — incl. the learned anchor projection of 34CR_Proj"""
    from oracle import losses_cpu
    coords, feats, labels = scenes.synthetic_voxel_scene(seed=9, target_voxels=700)
    c, f, lab = torch.from_numpy(coords), torch.from_numpy(feats), torch.from_numpy(labels)
    torch.manual_seed(1)
    anchors = torch.nn.functional.normalize(torch.randn(200, 512), dim=1)
    res = {}
    for mode in ("autograd", "program"):
        torch.manual_seed(42)
        net = nets.build_model(name, 3, 200, nets.DefaultConfig(), engine=me_cpu).train()
        net.representation_only(True)
        st = me_cpu.SparseTensor(f, c)

        def head(x):
            anc = net.projection_layer(anchors.unsqueeze(-1)).squeeze() if name.endswith("Proj") else anchors
            return losses_cpu.clip_ce_loss(x, lab, anc, ignore_label=-1, reduction="mean")
        if mode == "autograd":
            out = net(st, anchors)[0] if name.endswith("Proj") else net(st)
            loss = head(out.F)
            loss.backward()
        else:
            loss = step.StepProgram(net, OracleBackend()).run_features(st, head)
        res[mode] = (loss.item(), {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None})
    a, p = res["autograd"], res["program"]
    assert abs(a[0] - p[0]) < 1e-6 * abs(a[0]) and a[1].keys() == p[1].keys()
    if name.endswith("Proj"):
        assert "projection_layer.weight" in p[1]
    worst = max(((p[1][k] - g).norm() / g.norm().clamp(min=1e-20)).item() for k, g in a[1].items())
    assert worst < 1e-4, worst


_TRACE = r'''
import sys, re
sys.path.insert(0, ROOT)
import torch
from tests import stub_engine
from languagegroundedsemseg_b200 import _lib, losses, minkowski as E, nets, step
stub_engine.install(setattr, real_library=True)
losses._stream = lambda: None
SIZES = {1: 600, 2: 200, 4: 70, 8: 25, 16: 9}
torch.manual_seed(0)
net = nets.build_model(NAME, 3, 200, nets.DefaultConfig()).train()
mgr = stub_engine.FakeManager(SIZES)
lab = torch.randint(-1, 200, (SIZES[1],))
anchors = torch.nn.functional.normalize(torch.randn(200, 512), dim=1)
if net.flavour != "seg":
    net.representation_only(True)
prog = step.StepProgram(net)
with _lib.trace() as t:
    for _ in range(2):
        for p in net.parameters():
            p.grad = None
        if net.flavour == "seg":
            if MODE == "facade":
                out, _ = net(stub_engine.sparse_input(SIZES[1], 3, mgr))
                losses._SegCEFn.apply(out.F, lab, -1).backward()
            else:
                prog.run(stub_engine.sparse_input(SIZES[1], 3, mgr), lab, -1)
        else:
            crit = losses.ContrastiveLanguageCELoss(num_labels=200, ignore_label=-1)
            if MODE == "facade":
                feat, anc = net(stub_engine.sparse_input(SIZES[1], 3, mgr), anchors)
                crit(feat.F, lab, anc)[0].backward()
            else:
                prog.run_features(stub_engine.sparse_input(SIZES[1], 3, mgr),
                                  lambda x: crit(x, lab, net.projection_layer(anchors.unsqueeze(-1)).squeeze())[0])
        assert not [k for k, p in net.named_parameters() if p.grad is None or p.grad.shape != p.shape]
        E.invalidate_weight_cache()              # what an optimiser step would do
names = {}
for k, v in list(net.named_parameters()) + list(net.named_buffers()):
    names[v.data_ptr()] = k
for mn, m in net.named_modules():
    for key, bufs in getattr(m, "_prep_bufs", {}).items():
        for tag, b in zip(("w_fwd", "w_bwd"), bufs[:2]):
            if b is not None:
                names[b.data_ptr()] = f"{mn}.{tag}"
for (ts, out_ts, ks), (_, km) in mgr.cache.items():
    names[km.fwd_table.data_ptr()] = f"table{ks}:{ts}->{out_ts}:fwd"
    names[km.bwd_table.data_ptr()] = f"table{ks}:{ts}->{out_ts}:bwd"
for i, h in enumerate(E._scratch64(None).halves):
    names[h.value] = f"bn_scratch{i}"
def canon(m):
    return "0" if m.group(0) == "(nil)" else names.get(int(m.group(0), 16), "A")
print(*[re.sub(r"0x[0-9a-f]+|\(nil\)", canon, l) for l in t.lines], sep=chr(10))
'''


@pytest.mark.parametrize("name", ["Res16UNet34C", "Res16UNet14A", "Res16UNet34CR_Proj"])
def test_program_issues_the_facades_calls(lib, name):
    out = {}
    for mode in ("facade", "program"):
        code = f"ROOT={ROOT!r}\nNAME={name!r}\nMODE={mode!r}\n" + _TRACE
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        out[mode] = r.stdout.strip().splitlines()
    a, b = out["facade"], out["program"]
    assert len(a) == len(b) and len(a) > 300 and len(a) % 2 == 0
    # compare the SECOND step: by then every persistent tensor (cached weight operands, neighbour tables) exists and stays
    # alive, so no temporary of the traced step can share an address — and hence a name — with one of them
    a, b = a[len(a) // 2:], b[len(b) // 2:]
    for i, (x, y) in enumerate(zip(a, b)):
        assert x == y, (i, x, y)
    assert sum(l.startswith("lgs_seg_ce" if name != "Res16UNet34CR_Proj" else "lgs_clip_ce_tc") for l in a) == 1
