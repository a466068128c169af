"""The oracle (and nets.py's topology) against the committed golden vectors produced by the REFERENCE's own
classes (tests/golden/make_golden.py) — CPU only."""
import os

import numpy as np
import torch

from languagegroundedsemseg_b200 import nets, scenes
from oracle import losses_cpu, me_cpu, voxelize_cpu
from tests.helpers import dense_cube


def test_unet14a_cube_matches_reference_models(golden_dir):
    g = np.load(os.path.join(golden_dir, "unet14a_cube.npz"))
    torch.manual_seed(42)
    net = nets.build_model("Res16UNet14A", 3, 200, nets.DefaultConfig(), engine=me_cpu).train()
    c = dense_cube(20)
    torch.manual_seed(0)
    f = torch.rand(8000, 3) - 0.5
    st = me_cpu.SparseTensor(f, torch.from_numpy(c))
    with torch.no_grad():
        out, feat = net(st)
    mgr = st.coordinate_manager
    sizes = [mgr._coords[k].shape[0] for k in sorted(mgr._coords, key=lambda k: k.tensor_stride)]
    assert sizes == g["level_sizes"].tolist() == [8000, 1000, 125, 27, 8]
    pairs = [sum(len(i) for i, _ in km) for ck, km in mgr._kmaps.items() if ck[2] == (3, 3, 3)]
    assert pairs == g["pairs3"].tolist() == [195112, 21952, 2197, 343, 64]
    np.testing.assert_allclose(out.F[:64].numpy(), g["logits_head"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out.F[::125].numpy(), g["logits_rows"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(feat.F[::125].numpy(), g["feat_rows"], rtol=1e-4, atol=1e-5)


def test_unet34c_small_fwd_bwd_matches_reference_models(golden_dir):
    g = np.load(os.path.join(golden_dir, "unet34c_small.npz"))
    torch.manual_seed(42)
    net = nets.build_model("Res16UNet34C", 3, 200, nets.DefaultConfig(), engine=me_cpu).train()
    coords, feats, labels = scenes.synthetic_voxel_scene(seed=3, target_voxels=3000)
    assert coords.shape[0] == int(g["n"])
    st = me_cpu.SparseTensor(torch.from_numpy(feats), torch.from_numpy(coords))
    out, _ = net(st)
    loss = torch.nn.functional.cross_entropy(out.F, torch.from_numpy(labels), ignore_index=-1)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-4
    np.testing.assert_allclose(out.F.detach()[::50].numpy(), g["logits_rows"], rtol=1e-3, atol=1e-4)
    gn = dict(net.named_parameters())
    for k, v in zip(g["grad_keys"], g["grad_norms"]):
        assert abs(gn[str(k)].grad.norm().item() - v) <= 1e-3 * max(v, 1e-3)
    np.testing.assert_allclose(net.conv0p1s1.kernel.grad.numpy(), g["conv0_grad"], rtol=1e-3, atol=1e-5)


def test_clip_ce_oracle_matches_reference_class(golden_dir):
    g = np.load(os.path.join(golden_dir, "clip_ce.npz"))
    for tag in ("c96", "c512"):
        F_ = torch.from_numpy(g[f"{tag}_F"]).requires_grad_(True)
        A, y = torch.from_numpy(g[f"{tag}_A"]), torch.from_numpy(g[f"{tag}_y"])
        loss = losses_cpu.clip_ce_loss(F_, y, A, reduction="none")
        np.testing.assert_allclose(loss.detach().numpy(), g[f"{tag}_loss"], rtol=1e-5, atol=1e-6)
        lm = losses_cpu.clip_ce_loss(F_, y, A, reduction="mean")
        assert abs(lm.item() - float(g[f"{tag}_mean"])) < 1e-5
        lm.backward()
        np.testing.assert_allclose(F_.grad.numpy(), g[f"{tag}_grad"], rtol=1e-4, atol=1e-7)
        assert torch.all(loss[y == -1] == 0)


def test_clip_pretraining_nets_match_reference_models(golden_dir):
    """nets.py's CLIP pre-training topologies + the oracle loss vs the reference's own clip_models.py classes and its own
    ContrastiveLanguageCELoss (BASELINE configs 3 / 5 in small)"""
    g = np.load(os.path.join(golden_dir, "clip_nets.npz"))
    coords, feats, labels = scenes.synthetic_voxel_scene(seed=9, target_voxels=1500)
    assert coords.shape[0] == int(g["n"])
    torch.manual_seed(1)
    anchors = torch.nn.functional.normalize(torch.randn(200, 512), dim=1)
    for name in ("Res16UNet34CR_Proj", "Res16UNet34D"):
        torch.manual_seed(42)
        net = nets.build_model(name, 3, 200, nets.DefaultConfig(), engine=me_cpu).train()
        net.representation_only(True)
        st = me_cpu.SparseTensor(torch.from_numpy(feats), torch.from_numpy(coords))
        feat, anc = net(st, anchors) if name.endswith("Proj") else (net(st), anchors)
        loss = losses_cpu.clip_ce_loss(feat.F, torch.from_numpy(labels), anc, ignore_label=-1, reduction="mean")
        loss.backward()
        assert abs(loss.item() - float(g[f"{name}_loss"])) < 1e-5
        np.testing.assert_allclose(feat.F.detach()[::25].numpy(), g[f"{name}_feat_rows"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(anc.detach()[::20].numpy(), g[f"{name}_anchor_rows"], rtol=1e-5, atol=1e-6)
        assert abs(net.conv0p1s1.kernel.grad.norm().item() - float(g[f"{name}_g_conv0"])) < 1e-3 * float(g[f"{name}_g_conv0"])
        assert abs(net.block8[0].conv1.kernel.grad.norm().item() - float(g[f"{name}_g_block8"])) < 1e-3 * float(g[f"{name}_g_block8"])


def test_voxelize_oracle_matches_reference_voxelizer(golden_dir):
    g = np.load(os.path.join(golden_dir, "voxelize.npz"))
    q, uidx, _ = voxelize_cpu.voxelize(g["pts"], g["M"])
    assert np.array_equal(q, g["coords"])
    assert np.all(np.diff(uidx) > 0)
