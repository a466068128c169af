"""-m gpu (collected last on purpose): the explicit step program (languagegroundedsemseg_b200/step.py) on the device vs
the module-by-module facade + autograd — same kernels in the same order (tests/test_step_program.py proves that on the
CPU with the call recorder), so loss, logits, gradients and BatchNorm statistics must agree up to the arrival-order noise
of the red.add reductions.  Written after this round's GPU budget was spent: first executed by the round-end run."""
import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


def test_step_program_matches_facade_on_device(lib):
    from languagegroundedsemseg_b200 import minkowski as E, nets, scenes
    from languagegroundedsemseg_b200.step import StepProgram
    E.set_conv_algo("tc")
    coords, feats, labels = scenes.synthetic_voxel_scene(seed=6, target_voxels=6000)
    c, f, lab = (torch.from_numpy(a).cuda() for a in (coords, feats, labels))
    res = {}
    for mode in ("facade", "program"):
        torch.manual_seed(42)
        net = nets.build_model("Res16UNet34C", 3, 200, nets.DefaultConfig()).cuda().train()
        st = E.SparseTensor(f, c)
        if mode == "facade":
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
    a, p = res["facade"], res["program"]
    assert abs(a[0] - p[0]) < 1e-4 * abs(a[0])
    assert rel_err(p[1], a[1]) < 1e-4
    assert a[2].keys() == p[2].keys()
    errs = sorted(((p[2][k] - g).norm() / g.norm().clamp(min=1e-20)).item() for k, g in a[2].items())
    assert errs[len(errs) // 2] < 1e-3 and errs[-1] < 5e-2, (errs[len(errs) // 2], errs[-1])
    for k, v in a[3].items():
        assert torch.allclose(p[3][k].float(), v.float(), rtol=1e-4, atol=1e-6), k
