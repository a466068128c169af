"""Generate the committed golden fixtures by importing the REFERENCE's own Python from /root/reference
(build container only — the GPU box has no /root/reference, so the vectors travel as .npz files).

  unet14a_cube.npz   reference models/res16unet.py::Res16UNet14A driven through the oracle ME shim on the
                     BASELINE config-1 input (dense 20^3 cube): level sizes, 3^3 pair counts, logits sample.
                     Pins nets.py's topology + the oracle's composition (MinkowskiEngine itself is absent: the
                     conv arithmetic remains "parity unpinned", see oracle/me_cpu.py).
  unet34c_small.npz  reference Res16UNet34C fwd+bwd (CrossEntropy) on a small synthetic scene: loss, logits
                     sample, a few gradient norms.
  clip_ce.npz        reference lib/losses/ContrastiveLanguageLoss.py::ContrastiveLanguageCELoss outputs (a genuine
                     reference implementation: pins oracle/losses_cpu.py and the CUDA loss kernel).
  clip_nets.npz      reference models/clip_models.py::Res16UNet34CR_Proj and ::Res16UNet34D (representation_only) driven
                     through the oracle ME shim, with the reference's own ContrastiveLanguageCELoss on top (BASELINE
                     configs 3 / 5 in small): per-point features, projected anchors, loss, two gradient norms.
  augment.npz        reference lib/transforms.py classes (ElasticDistortion, RandomDropout, RandomHorizontalFlip,
                     ChromaticAutoContrast / Translation / Jitter composed as in lib/datasets/scannet.py) on seeded
                     inputs with seeded python/numpy RNGs: pins languagegroundedsemseg_b200/augment.py.
  voxelize.npz       reference lib/voxelizer.py::Voxelizer.voxelize (affine + floor by the reference's numpy code;
                     de-duplication by the oracle's sparse_quantize).

Run:  python tests/golden/make_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import me_cpu  # noqa: E402

ME = me_cpu.install()
sys.path.insert(0, "/root/reference")
import models  # noqa: E402  (the reference's own package, unmodified)

from languagegroundedsemseg_b200 import scenes  # noqa: E402

torch.set_num_threads(8)
cfg = types.SimpleNamespace(bn_momentum=0.02, conv1_kernel_size=3, dilations=[1, 1, 1, 1])


def cube_input(n=20):
    g = np.arange(n)
    zz, yy, xx = np.meshgrid(g, g, g, indexing="ij")
    coords = np.stack([np.zeros(n ** 3, np.int64), xx.ravel(), yy.ravel(), zz.ravel()], 1).astype(np.int32)
    torch.manual_seed(0)
    feats = torch.rand(n ** 3, 3) - 0.5
    return coords, feats


def unet14a():
    torch.manual_seed(42)
    net = models.load_model("Res16UNet14A")(3, 200, cfg)
    net.train()
    coords, feats = cube_input()
    st = ME.SparseTensor(feats, torch.from_numpy(coords))
    with torch.no_grad():
        out, feat = net(st)
    mgr = st.coordinate_manager
    sizes = [mgr._coords[k].shape[0] for k in sorted(mgr._coords, key=lambda k: k.tensor_stride)]
    pairs = [sum(len(i) for i, _ in km) for ck, km in mgr._kmaps.items() if ck[2] == (3, 3, 3)]
    np.savez_compressed(os.path.join(HERE, "unet14a_cube.npz"), level_sizes=np.array(sizes), pairs3=np.array(pairs),
                        logits_head=out.F[:64].numpy(), logits_rows=out.F[::125].numpy(),
                        feat_rows=feat.F[::125].numpy(), logits_abs_mean=np.float64(out.F.abs().mean().item()))
    print("unet14a", sizes, pairs, out.F.abs().mean().item())


def unet34c():
    torch.manual_seed(42)
    net = models.load_model("Res16UNet34C")(3, 200, cfg)
    net.train()
    coords, feats, labels = scenes.synthetic_voxel_scene(seed=3, target_voxels=3000)
    st = ME.SparseTensor(torch.from_numpy(feats), torch.from_numpy(coords))
    out, feat = net(st)
    loss = torch.nn.functional.cross_entropy(out.F, torch.from_numpy(labels), ignore_index=-1)
    loss.backward()
    g = {k: p.grad.norm().item() for k, p in net.named_parameters()}
    keys = ["conv0p1s1.kernel", "block1.0.conv1.kernel", "block4.5.conv2.kernel", "convtr4p16s2.kernel",
            "block8.1.conv2.kernel", "final.kernel", "final.bias", "bn0.bn.weight"]
    np.savez_compressed(os.path.join(HERE, "unet34c_small.npz"), n=np.int64(coords.shape[0]),
                        loss=np.float64(loss.item()), logits_rows=out.F.detach()[::50].numpy(),
                        grad_keys=np.array(keys), grad_norms=np.array([g[k] for k in keys]),
                        conv0_grad=net.conv0p1s1.kernel.grad.numpy())
    print("unet34c", coords.shape[0], loss.item(), [g[k] for k in keys])


def clip_ce():
    sys.modules.setdefault("joblib", __import__("joblib"))
    from lib.losses.ContrastiveLanguageLoss import ContrastiveLanguageCELoss
    lcfg = types.SimpleNamespace(ignore_label=-1, num_negative_samples=3, contrast_neg_thresh=0.6,
                                 contrast_pos_thresh=0.0, contrast_neg_weight=1.0,
                                 instance_augmentation_color_aug_prob=0.0, scannet_path="/nonexistent",
                                 projection_model_path="none", representation_distance_type="cos")
    out = {}
    for tag, (n, c) in {"c96": (300, 96), "c512": (130, 512)}.items():
        g = torch.Generator().manual_seed(7 + c)
        F_ = torch.randn(n, c, generator=g) * 1.7
        A = torch.randn(200, c, generator=g)
        y = torch.randint(0, 200, (n,), generator=g)
        y[torch.rand(n, generator=g) < 0.15] = -1
        F_.requires_grad_(True)
        crit = ContrastiveLanguageCELoss(lcfg, 200, reduction="none")
        loss = crit(F_, y, A)[0]
        crit_m = ContrastiveLanguageCELoss(lcfg, 200, reduction="mean")
        lm = crit_m(F_, y, A)[0]
        lm.backward()
        out.update({f"{tag}_F": F_.detach().numpy(), f"{tag}_A": A.numpy(), f"{tag}_y": y.numpy(),
                    f"{tag}_loss": loss.detach().numpy(), f"{tag}_mean": np.float64(lm.item()),
                    f"{tag}_grad": F_.grad.numpy()})
        print("clip_ce", tag, lm.item())
    np.savez_compressed(os.path.join(HERE, "clip_ce.npz"), **out)


def clip_nets():
    sys.modules.setdefault("joblib", __import__("joblib"))
    from lib.losses.ContrastiveLanguageLoss import ContrastiveLanguageCELoss
    lcfg = types.SimpleNamespace(ignore_label=-1, num_negative_samples=3, contrast_neg_thresh=0.6,
                                 contrast_pos_thresh=0.0, contrast_neg_weight=1.0,
                                 instance_augmentation_color_aug_prob=0.0, scannet_path="/nonexistent",
                                 projection_model_path="none", representation_distance_type="cos")
    coords, feats, labels = scenes.synthetic_voxel_scene(seed=9, target_voxels=1500)
    torch.manual_seed(1)
    anchors = torch.nn.functional.normalize(torch.randn(200, 512), dim=1)
    out = {"n": np.int64(coords.shape[0])}
    for name in ("Res16UNet34CR_Proj", "Res16UNet34D"):
        torch.manual_seed(42)
        net = models.load_model(name)(3, 200, cfg)
        net.train()
        net.representation_only(True)
        st = ME.SparseTensor(torch.from_numpy(feats), torch.from_numpy(coords))
        if name.endswith("Proj"):
            feat, anc = net(st, anchors)
        else:
            feat, anc = net(st), anchors
        loss = ContrastiveLanguageCELoss(lcfg, 200, reduction="mean")(feat.F, torch.from_numpy(labels).long(), anc)[0]
        loss.backward()
        out.update({f"{name}_feat_rows": feat.F.detach()[::25].numpy(), f"{name}_loss": np.float64(loss.item()),
                    f"{name}_anchor_rows": anc.detach()[::20].numpy(),
                    f"{name}_g_conv0": np.float64(net.conv0p1s1.kernel.grad.norm().item()),
                    f"{name}_g_block8": np.float64(net.block8[0].conv1.kernel.grad.norm().item())})
        print("clip_nets", name, coords.shape[0], tuple(feat.F.shape), loss.item())
    np.savez_compressed(os.path.join(HERE, "clip_nets.npz"), **out)


AUG_SEEDS = (0, 4, 7)      # seed 4 and 7 exercise RandomDropout (5000 -> 4000 points at full size)


def augment_pipeline(mod):
    return mod.Compose([mod.ElasticDistortion([(4.0, 1.6), (16.0, 6.4)]), mod.RandomDropout(0.2),
                        mod.RandomHorizontalFlip("z", False), mod.ChromaticAutoContrast(), mod.ChromaticTranslation(0.1),
                        mod.ChromaticJitter(0.05)])


def augment_inputs(n=1500):
    rng = np.random.default_rng(0)
    return ((rng.random((n, 3)) * np.array([300, 250, 120])).astype(np.float32),
            (rng.random((n, 3)) * 255).astype(np.float32), rng.integers(0, 20, n).astype(np.int32))


def augment():
    import random
    for m in ("matplotlib", "open3d"):          # imported at module level by lib/transforms.py, unused by these classes
        sys.modules.setdefault(m, types.ModuleType(m))
    import lib.transforms as RT
    c0, f0, l0 = augment_inputs()
    out = {}
    for seed in AUG_SEEDS:
        random.seed(seed)
        np.random.seed(seed)
        c, f, l = augment_pipeline(RT)(c0.copy(), f0.copy(), l0.copy())
        out.update({f"s{seed}_coords": c, f"s{seed}_feats": f, f"s{seed}_labels": l})
        print("augment", seed, c.shape)
        random.seed(seed)
        out[f"s{seed}_hue"] = RT.HueSaturationTranslation(0.5, 0.2)(None, np.floor(f0).copy(), None)[1]
    # batch assembly with a point budget (cfl_collate_fn_factory): 4 scenes, limits that keep 4 / 3 / 2 of them
    items = [(c0[:n].astype(np.int32), f0[:n], l0[:n], f"scene{n}") for n in (100, 250, 70, 400)]
    for limit in (0, 500, 360):
        bc, bf, bl, names = RT.cfl_collate_fn_factory(limit)(items)
        out.update({f"collate{limit}_coords": bc.numpy(), f"collate{limit}_feats": bf.numpy(),
                    f"collate{limit}_labels": bl.numpy(), f"collate{limit}_names": np.array(names)})
    np.savez_compressed(os.path.join(HERE, "augment.npz"), **out)


def voxelize():
    from lib.voxelizer import Voxelizer
    rng = np.random.default_rng(5)
    pts = (rng.random((6000, 3)) * np.array([3.0, 2.5, 2.0])).astype(np.float32)
    pts[:50] = pts[50:100]  # exact duplicates
    feats = rng.random((6000, 3)).astype(np.float32)
    labels = rng.integers(0, 20, 6000).astype(np.int32)
    np.random.seed(11)
    vox = Voxelizer(voxel_size=0.05, clip_bound=None, use_augmentation=True, scale_augmentation_bound=(0.9, 1.1),
                    rotation_augmentation_bound=((-np.pi / 64, np.pi / 64), (-np.pi / 64, np.pi / 64), (-np.pi, np.pi)),
                    translation_augmentation_ratio_bound=None, ignore_label=255)
    c, f, l, (M_v, M_r) = vox.voxelize(pts, feats, labels)
    Mfull = M_r @ M_v
    np.savez_compressed(os.path.join(HERE, "voxelize.npz"), pts=pts, M=Mfull, coords=c.astype(np.int32),
                        feats_kept=f, labels_kept=l)
    print("voxelize", pts.shape, "->", c.shape)


if __name__ == "__main__":
    only = sys.argv[1:]            # e.g. `make_golden.py clip_nets` regenerates one fixture
    for fn in (unet14a, unet34c, clip_ce, clip_nets, augment, voxelize):
        if not only or fn.__name__ in only:
            fn()
