"""CPU: reference-format checkpoints load into the engine's networks (languagegroundedsemseg_b200/checkpoint.py,
restating lib/utils.py:17-45 + main.py:103-119).  The source state dicts come from the reference's OWN model classes when
the checkout is present (build container), else from nets.py (same keys, checked by test_reference_models_on_facade)."""
import os
import sys
import types

import pytest
import torch

from languagegroundedsemseg_b200 import checkpoint, nets

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


_GEN = r'''
import sys, types, torch
sys.path.insert(0, ROOT)
import languagegroundedsemseg_b200 as lgs
lgs.install_as_minkowski()
sys.path.insert(0, REF)
import models                                   # the reference's own package, unmodified
cfg = types.SimpleNamespace(bn_momentum=0.02, conv1_kernel_size=3, dilations=[1, 1, 1, 1])
torch.manual_seed(7)
torch.save(models.load_model(NAME)(3, N_OUT, cfg).state_dict(), OUT)
'''
_cache = {}


def _source_state(name, n_out):
    """state dict of the reference's own model class (built in a subprocess so that the reference's `models` package is
    not left bound to this process's modules); nets.py's when there is no reference checkout"""
    if (name, n_out) not in _cache:
        if os.path.isdir(os.path.join(REF, "models")):
            import subprocess
            import tempfile
            out = os.path.join(tempfile.mkdtemp(), "ref_state.pth")
            code = f"ROOT={ROOT!r}\nREF={REF!r}\nNAME={name!r}\nN_OUT={n_out}\nOUT={out!r}\n" + _GEN
            r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stderr
            _cache[(name, n_out)] = torch.load(out)
        else:
            cfg = types.SimpleNamespace(bn_momentum=0.02, conv1_kernel_size=3, dilations=[1, 1, 1, 1])
            torch.manual_seed(7)
            _cache[(name, n_out)] = nets.build_model(name, 3, n_out, cfg).state_dict()
    return _cache[(name, n_out)]


@pytest.mark.parametrize("prefix", ["", "model.", "module.model.", "module.", "encoder."])
def test_prefixed_checkpoints_round_trip(lib, tmp_path, prefix):
    src = _source_state("Res16UNet34C", 200)
    path = tmp_path / "weights.pth"
    torch.save({"state_dict": {prefix + k: v for k, v in src.items()}, "epoch": 3, "arch": "Res16UNet34C"}, path)
    torch.manual_seed(99)
    net = nets.build_model("Res16UNet34C", 3, 200, nets.DefaultConfig())
    loaded, skipped, untouched = checkpoint.load_reference_checkpoint(net, str(path))
    assert not skipped and not untouched and len(loaded) == len(src)
    got = net.state_dict()
    assert all(torch.equal(got[k], v) for k, v in src.items())
    torch.manual_seed(99)
    strict = nets.build_model("Res16UNet34C", 3, 200, nets.DefaultConfig())
    checkpoint.load_reference_checkpoint(strict, {prefix + k: v for k, v in src.items()}, lenient=False)
    assert all(torch.equal(strict.state_dict()[k], v) for k, v in src.items())


def test_lenient_loading_keeps_mismatched_heads(lib):
    """main.py:109-117: a 20-class checkpoint loaded into a 200-class model takes everything but the `final` head"""
    src = _source_state("Res16UNet34C", 20)
    torch.manual_seed(5)
    net = nets.build_model("Res16UNet34C", 3, 200, nets.DefaultConfig())
    before = {k: v.clone() for k, v in net.state_dict().items()}
    loaded, skipped, untouched = checkpoint.load_reference_checkpoint(net, {"state_dict": {"model." + k: v for k, v in src.items()}})
    assert skipped == ["final.bias", "final.kernel"] and untouched == ["final.bias", "final.kernel"]
    got = net.state_dict()
    for k in loaded:
        assert torch.equal(got[k], src[k])
    for k in untouched:
        assert torch.equal(got[k], before[k])
    with pytest.raises(RuntimeError):
        checkpoint.load_reference_checkpoint(net, {"state_dict": src}, lenient=False)


def test_loading_expires_cached_weight_operands(lib, monkeypatch):
    from languagegroundedsemseg_b200 import minkowski as E
    from tests import stub_engine
    stub = stub_engine.install(monkeypatch.setattr)
    conv = E.MinkowskiConvolution(16, 16, kernel_size=1, dimension=3)
    mgr = stub_engine.FakeManager({1: 50})
    conv(stub_engine.sparse_input(50, 16, mgr)).F
    n = stub.calls["lgs_weight_prep_batch"]
    conv(stub_engine.sparse_input(50, 16, mgr)).F
    assert stub.calls["lgs_weight_prep_batch"] == n
    checkpoint.load_reference_checkpoint(conv, {"kernel": torch.ones(16, 16)})
    conv(stub_engine.sparse_input(50, 16, mgr)).F
    assert stub.calls["lgs_weight_prep_batch"] == n + 1
