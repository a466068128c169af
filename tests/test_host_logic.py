"""CPU tests of the facade's host logic against a stub of the C-ABI library (tests/stub_engine.py): no arithmetic, but
the autograd plumbing, call counts and cache-invalidation rules are the real code."""
import os

import pytest
import torch

from tests import stub_engine

SIZES = {1: 600, 2: 200, 4: 70, 8: 25, 16: 9}


@pytest.fixture
def stub(monkeypatch):
    from languagegroundedsemseg_b200 import minkowski as E
    s = stub_engine.install(monkeypatch.setattr)
    yield s
    E.set_conv_bn_fusion(True), E.set_wgrad_overlap(True), E.set_batched_weight_prep(True)


@pytest.mark.parametrize("fuse,overlap,batch", [(1, 1, 1), (0, 0, 0), (1, 0, 1), (0, 1, 0)])
def test_unet34c_host_plumbing(stub, fuse, overlap, batch):
    """two SGD steps of Res16UNet34C through the facade: every parameter receives a gradient of its own shape, and the
    number of C-ABI launches per step is what the path promises (63 convs fwd, 62 dgrad, 63 wgrad; one batched
    weight-operand launch per step instead of 62 per-layer ones; 62 fused BatchNorms when conv+BN fusion is on)"""
    from languagegroundedsemseg_b200 import minkowski as E, nets
    E.set_conv_bn_fusion(fuse), E.set_wgrad_overlap(overlap), E.set_batched_weight_prep(batch)
    torch.manual_seed(0)
    net = nets.build_model("Res16UNet34C", 3, 200, nets.DefaultConfig()).train()
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9)
    mgr = stub_engine.FakeManager(SIZES)
    for _ in range(2):
        out, feat = net(stub_engine.sparse_input(SIZES[1], 3, mgr))
        assert out.F.shape == (SIZES[1], 200) and feat.F.shape == (SIZES[1], 96)
        opt.zero_grad(set_to_none=True)
        out.F.float().mean().backward()
        assert not [k for k, p in net.named_parameters() if p.grad is None or p.grad.shape != p.shape]
        opt.step()
    c = stub.calls
    assert c["lgs_conv_fwd"] == 2 * (63 + 62) and c["lgs_conv_wgrad"] == 2 * 63
    if batch:
        assert c["lgs_weight_prep_batch"] == 2 and c["lgs_weight_prep"] == 2      # conv0p1s1 (c_in 3 -> 4) pads per call
    else:
        assert c["lgs_weight_prep"] == 2 * 63 and "lgs_weight_prep_batch" not in c
    if fuse:
        assert c["lgs_bn_fwd"] == c["lgs_bn_bwd"] == 2 * 62
    net.eval()                      # BatchNorm not fusable: lazy convs materialise through .F
    with torch.no_grad():
        assert net(stub_engine.sparse_input(SIZES[1], 3, stub_engine.FakeManager(SIZES)))[0].F.shape == (SIZES[1], 200)


def test_weight_operand_cache_invalidation(stub):
    """cached tensor-core weight operands are re-derived after an optimiser step (torch's fused optimisers do not bump the
    parameter's version counter), after any weight-gradient computation, after an in-place edit, and on request"""
    from languagegroundedsemseg_b200 import minkowski as E
    conv = E.MinkowskiConvolution(16, 16, kernel_size=1, dimension=3)
    mgr = stub_engine.FakeManager({1: 100})

    def fwd():
        return conv(stub_engine.sparse_input(100, 16, mgr)).F

    def n():
        return stub.calls.get("lgs_weight_prep_batch", 0)

    fwd()
    a = n()
    fwd()
    assert n() == a                                   # cached
    opt = torch.optim.SGD(conv.parameters(), lr=0.1, fused=True)
    v = conv.kernel._version
    conv.kernel.grad = torch.zeros_like(conv.kernel)
    opt.step()
    if conv.kernel._version == v:                     # the reason rule (b) exists
        pass
    fwd()
    assert n() == a + 1                               # optimiser post-step hook
    fwd()
    assert n() == a + 1
    fwd().sum().backward()
    fwd()
    assert n() == a + 2                               # a weight gradient was computed
    with torch.no_grad():
        conv.kernel.mul_(2.0)
    fwd()
    assert n() == a + 3                               # version counter
    E.invalidate_weight_cache()
    fwd()
    assert n() == a + 4


_TRACE_PROBE = r'''
import os, sys, re, types
sys.path.insert(0, ROOT)
import torch
from tests import stub_engine
from languagegroundedsemseg_b200 import _lib, minkowski as E, nets
stub_engine.install(setattr, real_library=True)
SIZES = {1: 600, 2: 200, 4: 70, 8: 25, 16: 9}
MODEL = os.environ.get("LGS_PROBE_MODEL", "Res16UNet34C")
torch.manual_seed(0)
if os.environ.get("LGS_PROBE_REFERENCE"):
    # the reference's own, unmodified models/ package over the facade installed as `MinkowskiEngine`
    import languagegroundedsemseg_b200 as lgs
    lgs.install_as_minkowski()
    sys.path.insert(0, os.environ["LGS_PROBE_REFERENCE"])
    import models
    cfg = types.SimpleNamespace(bn_momentum=0.02, conv1_kernel_size=3, dilations=[1, 1, 1, 1])
    net = models.load_model(MODEL)(3, 200, cfg).train()
else:
    net = nets.build_model(MODEL, 3, 200, nets.DefaultConfig()).train()
opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9)
mgr = stub_engine.FakeManager(SIZES)
anchors = torch.randn(200, 512)
with _lib.trace() as t:
    for _ in range(2):
        x = stub_engine.sparse_input(SIZES[1], 3, mgr)
        if MODEL == "Res16UNet34CR_Proj":
            (out, _), proj = net(x, anchors)
            extra = proj.float().mean()
        else:
            out, _ = net(x)
            extra = 0.0
        opt.zero_grad(set_to_none=True)
        (out.F.float().mean() + extra).backward()
        opt.step()
# canonical form (addresses differ from process to process and the allocator re-uses them): a pointer that is the address
# of a persistent tensor — parameter, buffer, cached weight operand, neighbour table, BatchNorm scratch half — becomes that
# tensor's NAME (an exact-address match, so this also proves the pointer value arrived intact); any other non-null pointer
# (activations, gradients, temporaries) becomes "A"; NULL becomes "0"
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
sc = E._scratch64(None)
for i, h in enumerate(sc.halves):
    names[h.value] = f"bn_scratch{i}"
def canon(m):
    if m.group(0) == "(nil)":
        return "0"
    return names.get(int(m.group(0), 16), "A")
lines = [re.sub(r"0x[0-9a-f]+|\(nil\)", canon, l) for l in t.lines]
print(_lib.binding())
print("\n".join(lines))
'''


def test_both_bindings_issue_identical_calls(lib):
    """two SGD steps of Res16UNet34C through the facade with the REAL library recording its calls (lgs_trace_begin: the
    compute entry points log their arguments and return; no GPU needed): the ctypes binding and the generated native
    binding (LGS_FAST_BIND=1) must produce the same call sequence with the same integer / float arguments and the same
    pointer structure — i.e. the native binding marshals every argument of every hot entry point like ctypes does"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for mode in ("0", "1"):
        r = subprocess.run([sys.executable, "-c", f"ROOT={root!r}\n" + _TRACE_PROBE], capture_output=True, text=True,
                           env=dict(os.environ, LGS_FAST_BIND=mode), timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        out = r.stdout.strip().splitlines()
        res[mode] = (out[0], out[1:])
    assert res["0"][0] == "ctypes" and res["1"][0] == "native"
    a, b = res["0"][1], res["1"][1]
    per_step = 125 + 63 + 62 + 62 + 1 + 1
    assert len(a) == len(b) == 2 * per_step
    # the entry names and every non-pointer argument must agree over both steps; the NAMED pointers are compared on the
    # second step only: in the first one a freed temporary can share its address with a persistent tensor allocated
    # later (cached weight operands, tables), and which temporaries do differs from process to process
    strip = lambda l: " ".join(t if (t[0].isdigit() or t[0] == "-" or t.startswith("lgs_")) else "P" for t in l.split())  # noqa: E731
    for x, y in zip(a, b):
        assert strip(x) == strip(y)
    a, b = a[per_step:], b[per_step:]
    for x, y in zip(a, b):
        assert x == y
    names = [l.split()[0] for l in a]
    assert names.count("lgs_conv_fwd") == 63 + 62 and names.count("lgs_conv_wgrad") == 63
    assert names.count("lgs_bn_fwd") == names.count("lgs_bn_bwd") == 62 and names.count("lgs_weight_prep_batch") == 1
    named = sum(tok not in ("A", "0") and not tok[0].isdigit() and not tok.startswith("lgs_") and not tok.startswith("-")
                for l in a for tok in l.split())
    assert named > 750                   # parameters, buffers, weight operands, tables and scratch halves were all recognised
    # BatchNorm scratch halves alternate: the half a call clears is the half the next call accumulates into
    bn = [l.split() for l in a if l.startswith("lgs_bn_")]
    for prev, nxt in zip(bn, bn[1:]):
        p_next = prev[-3] if prev[0] == "lgs_bn_fwd" else prev[-2]       # d_scratch_next of the earlier call
        n_acc = nxt[-4] if nxt[0] == "lgs_bn_fwd" else nxt[-3]           # d_scratch of the later call
        assert p_next == n_acc and p_next.startswith("bn_scratch")


def _one_step_trace(mode="0", model="Res16UNet34C", reference=None):
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, LGS_FAST_BIND=mode, LGS_PROBE_MODEL=model)
    env.pop("LGS_PROBE_REFERENCE", None)
    if reference:
        env["LGS_PROBE_REFERENCE"] = reference
    r = subprocess.run([sys.executable, "-c", f"ROOT={root!r}\n" + _TRACE_PROBE], capture_output=True, text=True,
                       env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()[1:]
    assert len(lines) % 2 == 0
    return lines[len(lines) // 2:]          # the second (steady-state) step


def test_facade_call_sequence_is_the_committed_one(lib, golden_dir):
    """the exact sequence of C-ABI calls (entry, sizes, flags, which table / weight operand / BatchNorm tensor) of one
    Res16UNet34C training step — the contract a native step driver has to reproduce — against the committed trace"""
    import os
    want = [l for l in open(os.path.join(golden_dir, "facade_trace_unet34c.txt")).read().splitlines() if not l.startswith("#")]
    got = _one_step_trace()
    assert len(got) == len(want) == 314
    for i, (g, w) in enumerate(zip(got, want)):
        assert g == w, (i, g, w)


REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference checkout not present")
@pytest.mark.parametrize("model", ["Res16UNet34C", "Res16UNet14A", "Res16UNet34CR_Proj", "Res16UNet34D"])
def test_reference_models_run_through_the_facade_like_nets(lib, golden_dir, model):
    """The drop-in claim, exercised: the reference's UNMODIFIED models/ package (models/res16unet.py:196-270,
    models/clip_models.py:95-215, models/modules/resnet_block.py:41-57) runs forward + backward + SGD through the facade
    installed as `MinkowskiEngine`, with the library recording every C-ABI call, and must issue exactly the call sequence
    nets.py issues for the same topology — same entry points, sizes, flags, tables, weight operands, BatchNorm tensors.
    This fails if the lazy conv / BatchNorm deferral reacts differently to the reference's own call pattern (separate
    module calls, `out += residual`, NoReluBlock, me.cat).  For Res16UNet34C that sequence is also the committed golden."""
    got = _one_step_trace(model=model, reference=REF)
    mine = _one_step_trace(model=model)
    assert len(got) == len(mine) > 150
    for i, (g, w) in enumerate(zip(got, mine)):
        assert g == w, (i, g, w)
    if model == "Res16UNet34C":
        want = [l for l in open(os.path.join(golden_dir, "facade_trace_unet34c.txt")).read().splitlines() if not l.startswith("#")]
        assert got == want


if __name__ == "__main__":
    import sys
    if "--regen" in sys.argv:
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "facade_trace_unet34c.txt")
        head = [l for l in open(path).read().splitlines() if l.startswith("#")]
        open(path, "w").write("\n".join(head + _one_step_trace()) + "\n")
        print("regenerated", path)


def test_bench_roofline_report_on_synthetic_launch_records():
    """bench.py's post-processing (the `roofline` object of the JSON line) on hand-made per-launch records: the dominant
    group is the one with the largest summed time, achieved = gather-model bytes of one launch / its mean duration"""
    import os
    import sys
    import types
    import torch
    sys_path_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if sys_path_root not in sys.path:
        sys.path.insert(0, sys_path_root)
    import bench

    class Ev:
        def __init__(self, t):
            self.t = t

        def elapsed_time(self, other):
            return other.t - self.t
    km_l0 = types.SimpleNamespace(K=27)
    km_l1 = types.SimpleNamespace(K=27)
    pairs = {id(km_l0): 2_218_378, id(km_l1): 600_000}
    prof = []
    for step in range(2):
        for _ in range(6):
            prof.append((("fwd", 27, 96, 96, 149106, 149106, km_l0, torch.float32), (Ev(0.0), Ev(0.37))))
        for _ in range(6):
            prof.append((("dgrad", 27, 96, 96, 38506, 38506, km_l1, torch.float32), (Ev(0.0), Ev(0.14))))
        for _ in range(3):
            prof.append((("wgrad", 27, 96, 96, 149106, 149106, km_l0, torch.float32), (Ev(0.0), Ev(0.345))))
        prof.append((("fwd", 1, 96, 200, 149106, 149106, None, torch.float32), (Ev(0.0), Ev(0.137))))
    r = bench.roofline_report(prof, 16.0, "tc", "f32", 2, pair_counts=pairs)
    by = 2_218_378 * (96 + 96) * 4 + 8 * 2_218_378 + 27 * 96 * 96 * 4
    assert r["algorithmic_bytes_per_launch"] == by and r["launches_per_step"] == 6
    assert abs(r["avg_launch_ms"] - 0.37) < 1e-9 and abs(r["achieved"] - by / 0.37e-3 / 1e9) < 0.1
    assert "96->96" in r["kernel"] and r["bound"] == "hbm" and 0 < r["frac"] < 1.5
    assert abs(r["share_of_step"] - 6 * 0.37 / 16.0) < 1e-3
    assert r["all_wgrad"]["share_of_step"] == round(3 * 0.345 / 16.0, 3)
    assert r["traffic"] is None or isinstance(r["traffic"], int)


def test_bank_colour_scheme_of_the_neighbourhood_plan():
    """csrc/nbplan.cu, "Colours": g = (x + 3y + 5z) / step mod 8 is additive, so a translation by any kernel offset shifts the
    colours of a group of voxels by one constant — eight pairwise different colours stay pairwise different for all 27 offsets
    (the bank-conflict-free cache reads of conv_nb.cu rest on this) — and every axis-aligned plane carries all eight colours."""
    import itertools
    import numpy as np
    rng = np.random.default_rng(0)
    colour = lambda c, step: ((c[..., 0] // step) + 3 * (c[..., 1] // step) + 5 * (c[..., 2] // step)) & 7  # noqa: E731
    for step in (1, 2, 8):
        pts = rng.integers(-50, 50, size=(4000, 3)) * step
        g = colour(pts, step)
        octet = np.stack([pts[np.flatnonzero(g == k)[0]] for k in range(8)])          # eight voxels of eight different colours
        for off in itertools.product((-1, 0, 1), repeat=3):
            moved = colour(octet + np.array(off) * step, step)
            assert len(set(moved.tolist())) == 8
            assert len(set(((moved - colour(octet, step)) & 7).tolist())) == 1        # one constant shift
        for axis in range(3):                                                          # planes x / y / z = const
            plane = pts.copy()
            plane[:, axis] = 3 * step
            assert set(colour(plane, step).tolist()) == set(range(8))
