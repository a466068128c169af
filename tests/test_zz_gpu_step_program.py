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


@pytest.mark.parametrize("algo", ["bx3", "tc"])
def test_step_program_matches_facade_on_device(lib, algo):
    """The program issues the facade's kernels in the facade's order, so the two can differ only by the engine's own
    run-to-run noise: wgrad and the split small-map convolutions reduce through red.add in arrival order, and this randomly
    initialised BatchNorm network with random labels amplifies a 1e-7 perturbation of a weight gradient to ~1e-3 on the
    parameter gradients (round-1 driver run: 3.2e-3 median between facade and program).  The gate is therefore set from a
    measured baseline — the facade against ITSELF on identical inputs — instead of a constant chosen on the CPU."""
    from languagegroundedsemseg_b200 import minkowski as E, nets, scenes
    from languagegroundedsemseg_b200.step import StepProgram
    E.set_conv_algo(algo)
    coords, feats, labels = scenes.synthetic_voxel_scene(seed=6, target_voxels=6000)
    c, f, lab = (torch.from_numpy(a).cuda() for a in (coords, feats, labels))
    res = {}
    for mode in ("facade", "facade2", "program"):
        torch.manual_seed(42)
        net = nets.build_model("Res16UNet34C", 3, 200, nets.DefaultConfig()).cuda().train()
        st = E.SparseTensor(f, c)
        if mode != "program":
            out, _ = net(st)
            loss = torch.nn.functional.cross_entropy(out.F, lab.long(), ignore_index=-1)
            loss.backward()
            logits = out.F.detach()
        else:
            prog = StepProgram(net)
            loss = prog.run(st, lab, ignore_index=-1)
            logits = prog.logits
        res[mode] = (loss.item(), logits.cpu(), {k: p.grad.detach().cpu() for k, p in net.named_parameters()},
                     {k: v.cpu().clone() for k, v in net.state_dict().items() if "running" in k or "tracked" in k})
    E.set_conv_algo("bx3")
    a, a2, p = res["facade"], res["facade2"], res["program"]
    assert abs(a[0] - p[0]) < 1e-4 * abs(a[0])
    assert rel_err(p[1], a[1]) < 1e-4
    assert a[2].keys() == p[2].keys()
    noise = _grad_errs(a2[2], a[2])                  # the engine against itself
    errs = _grad_errs(p[2], a[2])
    n = len(errs)
    print(f"[{algo}] facade vs facade: median {noise[n // 2]:.2e} worst {noise[-1]:.2e};  program vs facade: "
          f"median {errs[n // 2]:.2e} worst {errs[-1]:.2e}")
    assert errs[n // 2] < max(3 * noise[n // 2], 1e-5) and errs[-1] < max(3 * noise[-1], 1e-4), (errs[n // 2], errs[-1], noise[n // 2], noise[-1])
    assert errs[-1] < 5e-2
    for k, v in a[3].items():
        assert torch.allclose(p[3][k].float(), v.float(), rtol=1e-4, atol=1e-6), k
