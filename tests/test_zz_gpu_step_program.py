"""-m gpu (collected last on purpose): the explicit step program (languagegroundedsemseg_b200/step.py) on the device vs
the module-by-module facade + autograd — same kernels in the same order (tests/test_step_program.py proves that on the
CPU with the call recorder), so loss, logits, gradients and BatchNorm statistics must agree up to the arrival-order noise
of the red.add reductions.  """
import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


def _grad_errs(x, ref):
    return sorted(((x[k] - g).norm() / g.norm().clamp(min=1e-20)).item() for k, g in ref.items())


def _one_step(E, nets, c, f, lab, mode, algo, **kw):
    """loss, logits, gradients, BatchNorm statistics of one fwd+bwd of Res16UNet34C through `mode`"""
    from languagegroundedsemseg_b200.program import NativeStep
    from languagegroundedsemseg_b200.step import StepProgram
    E.set_conv_algo(algo)
    torch.manual_seed(42)
    net = nets.build_model("Res16UNet34C", 3, 200, nets.DefaultConfig()).cuda().train()
    st = E.SparseTensor(f, c)
    if mode == "facade":
        out, _ = net(st)
        loss = torch.nn.functional.cross_entropy(out.F, lab.long(), ignore_index=-1)
        loss.backward()
        logits = out.F.detach()
    elif mode == "program":
        prog = StepProgram(net)
        loss = prog.run(st, lab, ignore_index=-1)
        logits = prog.logits
    else:
        step = NativeStep(net, keep_logits=True, **kw)
        loss = step.run(st, lab)
        logits = step.logits
    torch.cuda.synchronize()
    return (loss.item(), logits.cpu(), {k: p.grad.detach().cpu().clone() for k, p in net.named_parameters()},
            {k: v.cpu().clone() for k, v in net.state_dict().items() if "running" in k or "tracked" in k})


def _check_against_exact(name, cand, facades, exact):
    """`cand` (another driver of the same kernels) must sit as close to the exact-fp32 result as the facade does: the
    per-parameter gradient errors of this ill-conditioned problem (random init, random labels, BatchNorm everywhere) are
    dominated by how rounding noise is amplified, not by who issued the kernels, so the yardstick is the facade's own
    distance from the exact SIMT kernels — measured in the same test, twice (run-to-run spread of the red.add sums)."""
    e_f = [_grad_errs(x[2], exact[2]) for x in facades]
    e_c = _grad_errs(cand[2], exact[2])
    n = len(e_c)
    med_f, worst_f = max(e[n // 2] for e in e_f), max(e[-1] for e in e_f)
    print(f"[{name}] vs exact fp32: facade median {med_f:.2e} worst {worst_f:.2e};  candidate median {e_c[n // 2]:.2e} worst {e_c[-1]:.2e}")
    # The noise is bimodal: identical runs of ONE driver land at a median of either ~4.4e-3 or ~9.8e-3 ('tc') depending on how a
    # handful of arrival-order-dependent sums round (3 x 6 repetitions in profiles/r2_noise_gate_runs.txt: the facade itself
    # measured 4.56e-3, 4.45e-3 and 9.79e-3).  Two facade samples can both fall into the low mode, so the gate also admits the
    # high mode's level; a driver that issues a wrong, missing or misordered kernel is off by O(1), not by 1e-2.
    assert e_c[n // 2] < max(2 * med_f, 3e-2) and e_c[-1] < max(2 * worst_f, 6e-2), (e_c[n // 2], e_c[-1], med_f, worst_f)
    a = facades[0]
    assert abs(a[0] - cand[0]) < 1e-4 * abs(a[0]), (a[0], cand[0])
    assert rel_err(cand[1], a[1]) < 1e-4
    assert a[2].keys() == cand[2].keys()
    for k, v in a[3].items():
        # running statistics of the coarsest levels (a few hundred rows) carry the noise of every layer above them
        assert torch.allclose(cand[3][k].float(), v.float(), rtol=2e-3, atol=2e-5), k


@pytest.mark.parametrize("algo", ["bx3", "tc"])
def test_step_program_matches_facade_on_device(lib, algo):
    """The Python step program issues the facade's kernels in the facade's order (tests/test_step_program.py proves it with
    the call recorder), so on the device the two can differ only by the engine's run-to-run noise (wgrad and the split
    small-map convolutions reduce through red.add in arrival order).  Round 1 gated this with a constant chosen on the CPU
    and failed on the driver's box (3.2e-3 median against 1e-3); the gate is now measured: see _check_against_exact."""
    from languagegroundedsemseg_b200 import minkowski as E, nets, scenes
    coords, feats, labels = scenes.synthetic_voxel_scene(seed=6, target_voxels=6000)
    c, f, lab = (torch.from_numpy(a).cuda() for a in (coords, feats, labels))
    exact = _one_step(E, nets, c, f, lab, "facade", "simt")
    facades = [_one_step(E, nets, c, f, lab, "facade", algo) for _ in range(2)]
    cand = _one_step(E, nets, c, f, lab, "program", algo)
    E.set_conv_algo("bx3")
    _check_against_exact(f"StepProgram {algo}", cand, facades, exact)


@pytest.mark.parametrize("fuse", [True, False])
@pytest.mark.parametrize("voxels", [6000, 40000])
def test_native_step_matches_facade_on_device(lib, voxels, fuse):
    """the native step driver (program.NativeStep: one lgs_program_run per step) against the module-by-module facade +
    autograd on the same inputs and weights: loss, per-point logits, EVERY parameter gradient and the BatchNorm running
    statistics.  40 K voxels puts level 0 above the row count from which the BatchNorm statistics come out of the
    convolution epilogue (fuse=True)."""
    from languagegroundedsemseg_b200 import minkowski as E, nets, scenes
    coords, feats, labels = scenes.synthetic_voxel_scene(seed=6, target_voxels=voxels)
    c, f, lab = (torch.from_numpy(a).cuda() for a in (coords, feats, labels))
    exact = _one_step(E, nets, c, f, lab, "facade", "simt")
    facades = [_one_step(E, nets, c, f, lab, "facade", "bx3") for _ in range(2)]
    cand = _one_step(E, nets, c, f, lab, "native", "bx3", fuse_bn_stats=fuse)
    _check_against_exact(f"NativeStep fuse={fuse} n={c.shape[0]}", cand, facades, exact)


def test_native_step_trains(lib):
    """ten SGD steps through NativeStep and through the facade from the same initialisation follow the same loss curve"""
    from languagegroundedsemseg_b200 import minkowski as E, nets, scenes
    from languagegroundedsemseg_b200.program import NativeStep
    E.set_conv_algo("bx3")
    coords, feats, labels = scenes.synthetic_voxel_scene(seed=8, target_voxels=8000)
    c, f, lab = (torch.from_numpy(a).cuda() for a in (coords, feats, labels))
    curves = {}
    for mode in ("facade", "native"):
        torch.manual_seed(42)
        net = nets.build_model("Res16UNet34C", 3, 200, nets.DefaultConfig()).cuda().train()
        step = NativeStep(net) if mode == "native" else None
        opt = torch.optim.SGD(net.parameters(), lr=0.05, momentum=0.9)
        ls = []
        for _ in range(10):
            st = E.SparseTensor(f, c)
            if step is not None:
                loss = step.run(st, lab)
            else:
                out, _ = net(st)
                loss = torch.nn.functional.cross_entropy(out.F, lab.long(), ignore_index=-1)
                opt.zero_grad(set_to_none=True)
                loss.backward()
            opt.step()
            ls.append(loss.item())
        curves[mode] = ls
    a, b = curves["facade"], curves["native"]
    print(f"[NativeStep trains] facade {a[0]:.4f} -> {a[-1]:.4f}, native {b[0]:.4f} -> {b[-1]:.4f}, "
          f"largest relative gap over the 10 steps {max(abs(x - y) / abs(x) for x, y in zip(a, b)):.2e}")
    assert a[-1] < a[0] - 0.3 and b[-1] < b[0] - 0.3, (a, b)         # it learns
    assert all(abs(x - y) < 0.02 * abs(x) for x, y in zip(a, b)), (a, b)
