"""Drop-in check at the reference's call sites (build container only: /root/reference is not on the GPU box).
The reference's unmodified models/ package is imported against the CUDA facade installed as `MinkowskiEngine`
(module construction needs no GPU) and compared with nets.py: same state-dict keys, shapes and — for the same seed —
values, so the graphs bench.py runs are the reference's graphs."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

SCRIPT = r'''
import sys, types, torch
sys.path.insert(0, ROOT)
import languagegroundedsemseg_b200 as lgs
ME = lgs.install_as_minkowski()
sys.path.insert(0, REF)
import models                                   # the reference's own package, unmodified
from languagegroundedsemseg_b200 import nets
cfg = types.SimpleNamespace(bn_momentum=0.02, conv1_kernel_size=3, dilations=[1, 1, 1, 1])
for name in ("Res16UNet14A", "Res16UNet34C", "Res16UNet34CR", "Res16UNet34CR_Proj", "Res16UNet34D"):
    torch.manual_seed(42); ref = models.load_model(name)(3, 200, cfg)
    torch.manual_seed(42); mine = nets.build_model(name, 3, 200, cfg)
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a) == list(b), name
    assert all(a[k].shape == b[k].shape and torch.equal(a[k], b[k]) for k in a), name
    convs = [m for m in ref.modules() if isinstance(m, (ME.MinkowskiConvolution, ME.MinkowskiConvolutionTranspose))]
    assert len(convs) == {"Res16UNet14A": 33}.get(name, 63), (name, len(convs))   # SURVEY.md App. B: 33 / 59+4
    assert sum(isinstance(m, ME.MinkowskiBatchNorm) for m in ref.modules()) in (32, 62)
print("OK")
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference checkout not present")
def test_reference_models_construct_on_cuda_facade(lib):
    code = f"ROOT={ROOT!r}\nREF={REF!r}\n" + SCRIPT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
