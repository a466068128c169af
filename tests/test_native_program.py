"""CPU tests of the native step driver (languagegroundedsemseg_b200/program.py + csrc/program.cu) with the library's call
recorder: the program is built from the network's modules, executed by lgs_program_run with recording on (every entry
point logs its arguments and returns; no GPU), and its call sequence is held against the facade's committed trace
(tests/golden/facade_trace_unet34c.txt) — same convolutions, weight gradients and BatchNorm calls, same sizes, same order."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

_PROBE = r'''
import os, sys, types
sys.path.insert(0, ROOT)
import torch
from tests import stub_engine
from languagegroundedsemseg_b200 import _lib, minkowski as E, nets
from languagegroundedsemseg_b200.program import NativeStep
SIZES = {1: 600, 2: 200, 4: 70, 8: 25, 16: 9}
MODEL = os.environ.get("LGS_PROBE_MODEL", "Res16UNet34C")
torch.manual_seed(0)
if os.environ.get("LGS_PROBE_REFERENCE"):
    import languagegroundedsemseg_b200 as lgs
    lgs.install_as_minkowski()
    sys.path.insert(0, os.environ["LGS_PROBE_REFERENCE"])
    import models
    cfg = types.SimpleNamespace(bn_momentum=0.02, conv1_kernel_size=3, dilations=[1, 1, 1, 1])
    net = models.load_model(MODEL)(3, 200, cfg).train()
else:
    net = nets.build_model(MODEL, 3, 200, nets.DefaultConfig()).train()
step = NativeStep(net, _dry=True)
mgr = stub_engine.FakeManager(SIZES)
x = stub_engine.sparse_input(SIZES[1], 3, mgr)
labels = torch.zeros(SIZES[1], dtype=torch.long)
with _lib.trace() as t:
    step.run(x, labels)
    step.run(x, labels)
print(step.n_ops, len(t.lines))
print("\n".join(t.lines))
'''


def _program_trace(model="Res16UNet34C", reference=None):
    env = dict(os.environ, LGS_PROBE_MODEL=model)
    env.pop("LGS_PROBE_REFERENCE", None)
    if reference:
        env["LGS_PROBE_REFERENCE"] = reference
    r = subprocess.run([sys.executable, "-c", f"ROOT={ROOT!r}\n" + _PROBE], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = r.stdout.strip().splitlines()
    n_ops, n_lines = map(int, lines[0].split())
    calls = lines[1:]
    assert len(calls) == n_lines and n_lines % 2 == 0
    return n_ops, calls[:n_lines // 2], calls[n_lines // 2:]


def _conv_key_facade(t):      # lgs_conv_fwd in n_in c_in w layout K c_out table n_out reverse bias out dtype algo stream
    return (int(t[2]), int(t[3]), int(t[6]), int(t[7]), int(t[9]), int(t[10]), t[11] != "0")


def _conv_key_program(t):     # lgs_conv_fwd2 in c_in in2 c_in2 n_in w K c_out table n_out reverse bias out sums stream
    return (int(t[5]), int(t[2]), int(t[7]), int(t[8]), int(t[10]), int(t[11]), t[12] != "(nil)")


def test_program_issues_the_facades_convolutions_in_order(golden_dir):
    want = [l.split() for l in open(os.path.join(golden_dir, "facade_trace_unet34c.txt")).read().splitlines() if not l.startswith("#")]
    n_ops, first, second = _program_trace()
    strip = lambda ls: [" ".join(tok if not tok.startswith("0x") else "P" for tok in l.split()) for l in ls]  # noqa: E731
    assert strip(first) == strip(second)                      # every step is the same program
    got = [l.split() for l in second]
    names = [g[0] for g in got]
    assert names.count("lgs_program_run") == 0                # the driver itself is not a recorded entry
    assert names.count("lgs_conv_fwd2") == 63 + 62 and names.count("lgs_conv_wgrad") == 63
    assert names.count("lgs_bn_fwd") + names.count("lgs_bn_fwd2") == 62 and names.count("lgs_bn_bwd") == 62
    assert names.count("lgs_weight_prep_batch") == 1 and names.count("lgs_seg_ce") == 1 and names.count("lgs_colsum") == 1
    # convolutions: same (n_in, c_in, K, c_out, n_out, mirrored table, bias) sequence as the facade's forward + dgrad calls
    fac = [_conv_key_facade(t) for t in want if t[0] == "lgs_conv_fwd"]
    prog = [_conv_key_program(t) for t in got if t[0] == "lgs_conv_fwd2"]
    assert prog == fac
    # weight gradients: same entry point, same sizes; same order except that the driver issues the four level-0 3x3x3 decoder
    # wgrads (block8) after the decoder's backward instead of next to their dgrads (program.py, LGS_DEFER_WGRAD)
    wg = lambda ls: [(t[2], t[3], t[5], t[6], t[8]) for t in ls if t[0] == "lgs_conv_wgrad"]  # noqa: E731
    a, b = wg(got), wg(want)
    assert sorted(a) == sorted(b)
    n0 = max(int(t[0]) for t in b)                       # rows of level 0
    late = [t for t in b if int(t[0]) == n0 and t[4] == "27" and int(t[1]) >= 64][:4]
    assert len(late) == 4
    rest = list(b)
    for t in late:
        rest.remove(t)                                   # first occurrence = the decoder's (backward visits block8 first)
    i = next(k for k in range(len(a)) if a[k:k + 4] == late)
    assert a[:i] + a[i + 4:] == rest
    # BatchNorm: (rows, channels, relu, residual present) forward; (rows, channels, relu, d_residual wanted) backward
    bnf = lambda ls: [(t[3], t[4], t[9], t[2] not in ("0", "(nil)")) for t in ls if t[0] in ("lgs_bn_fwd", "lgs_bn_fwd2")]  # noqa: E731
    assert bnf(got) == bnf(want)
    bnb = lambda ls: [(t[4], t[5], t[9], t[11] not in ("0", "(nil)")) for t in ls if t[0] == "lgs_bn_bwd"]  # noqa: E731
    assert bnb(got) == bnb(want)
    assert 300 < n_ops < 450


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference checkout not present")
def test_program_from_the_reference_models_equals_program_from_nets():
    """NativeStep built from the reference's own models.Res16UNet34C instance issues the calls of the one built from nets.py"""
    strip = lambda ls: [" ".join(tok if not tok.startswith("0x") else "P" for tok in l.split()) for l in ls]  # noqa: E731
    a = _program_trace(reference=REF)
    b = _program_trace()
    assert a[0] == b[0] and strip(a[2]) == strip(b[2])
