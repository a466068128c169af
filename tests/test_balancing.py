"""CPU: on-device `sample_categories_for_balancing` (languagegroundedsemseg_b200/losses.py) vs the reference function
(lib/losses/utils.py:13-77; imported in a subprocess with a stub for its unused torchmetrics import) and by its properties."""
import os
import subprocess
import sys
import types

import pytest
import torch

from languagegroundedsemseg_b200 import losses

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup(n=5000, L=40, seed=0):
    torch.manual_seed(seed)
    cats = torch.zeros(L, 3, dtype=torch.bool)
    perm = torch.randperm(L)
    cats[perm[:10], 0] = True
    cats[perm[10:25], 1] = True
    cats[perm[25:], 2] = True
    ds = types.SimpleNamespace(NUM_LABELS=L, frequency_organized_cats=cats)
    targets = torch.randint(0, L, (n,))
    targets[torch.rand(n) < 0.1] = -1
    return ds, targets, torch.rand(n)


def _cfg(h, c):
    return types.SimpleNamespace(ignore_label=-1, balanced_sample_head_ratio=h, balanced_sample_common_ratio=c)


def test_keeps_the_reference_quota_of_every_class():
    ds, targets, _ = _setup()
    cats = ds.frequency_organized_cats
    n = targets.shape[0]
    g = torch.Generator().manual_seed(5)
    mean, (hl, cl, tl), items = losses.sample_categories_for_balancing(torch.ones(n), _cfg(0.5, 0.25), ds, targets, generator=g)
    want = sum(round(0.5 * k) if cats[c, 0] else round(0.25 * k) if cats[c, 1] else k
               for c, k in enumerate(torch.bincount(targets[targets >= 0], minlength=40).tolist()))
    assert round(float(mean) * n) == want
    assert items.shape == (int((targets != -1).sum()), 3) and bool((items.sum(1) == 1).all())
    assert hl.numel() + cl.numel() + tl.numel() == items.shape[0]
    # per class: exactly round(ratio * count) survivors
    lossv = torch.arange(n, dtype=torch.float32) + 1
    keep = None
    for _ in range(3):                      # different draws, same counts
        m = losses.sample_categories_for_balancing(lossv, _cfg(0.5, -1), ds, targets)[0]
        assert float(m) > 0
    one_hot_loss = torch.zeros(n)
    cls = int(torch.nonzero(cats[:, 0])[0])
    one_hot_loss[targets == cls] = 1.0
    kept = float(losses.sample_categories_for_balancing(one_hot_loss, _cfg(0.5, -1), ds, targets)[0]) * n
    assert round(kept) == round(0.5 * int((targets == cls).sum()))


def test_loss_given_for_valid_points_only():
    """the reference accepts the loss of the valid points alone (lib/losses/utils.py:17-19)"""
    ds, targets, loss = _setup()
    valid = targets != -1
    a = losses.sample_categories_for_balancing(loss[valid], _cfg(-1, -1), ds, targets)
    b = losses.sample_categories_for_balancing(loss, _cfg(-1, -1), ds, targets)
    assert torch.equal(a[2], b[2]) and abs(float(a[0]) - float(loss[valid].mean())) < 1e-6


_DIFF = r'''
import sys, types, warnings
import numpy as np, torch
warnings.filterwarnings("ignore")
tm = types.ModuleType("torchmetrics"); tm.Metric = type("Metric", (), {"__init__": lambda self, *a, **k: None})
sys.modules["torchmetrics"] = tm                 # imported at module level by lib/losses/utils.py, unused by this function
sys.path.insert(0, ROOT)
import importlib.util
spec = importlib.util.spec_from_file_location("ref_loss_utils", "/root/reference/lib/losses/utils.py")
RU = importlib.util.module_from_spec(spec); spec.loader.exec_module(RU)
from languagegroundedsemseg_b200 import losses
from tests.test_balancing import _setup, _cfg
for seed in range(3):
    ds, targets, loss = _setup(seed=seed)
    r = RU.sample_categories_for_balancing(loss.clone(), _cfg(-1, -1), ds, targets.clone())
    a = losses.sample_categories_for_balancing(loss.clone(), _cfg(-1, -1), ds, targets.clone())
    assert float(r[0]) == float(a[0]) and torch.equal(r[2], a[2]) and all(torch.equal(x, y) for x, y in zip(r[1], a[1]))
    np.random.seed(seed)
    r = RU.sample_categories_for_balancing(torch.ones_like(loss), _cfg(0.5, 0.25), ds, targets.clone())
    a = losses.sample_categories_for_balancing(torch.ones_like(loss), _cfg(0.5, 0.25), ds, targets.clone())
    assert abs(float(r[0]) - float(a[0])) < 1e-7 and torch.equal(r[2], a[2])      # same number of kept points
print("OK")
'''


@pytest.mark.skipif(not os.path.isdir("/root/reference/lib"), reason="reference checkout not present")
def test_matches_reference_function():
    r = subprocess.run([sys.executable, "-c", f"ROOT={ROOT!r}\n" + _DIFF], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
