"""-m gpu: coordinate maps and kernel maps of the CUDA engine vs the oracle — bit-exact (integer work)."""
import numpy as np
import pytest
import torch

from tests.helpers import dense_cube, pair_set, pairs_array, random_sparse_coords

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E(lib):
    from languagegroundedsemseg_b200 import minkowski
    return minkowski


def _both(E, c, feats=None):
    from oracle import me_cpu
    f = feats if feats is not None else torch.zeros(c.shape[0], 1)
    return E.SparseTensor(f.cuda(), torch.from_numpy(c).cuda()), me_cpu.SparseTensor(f, torch.from_numpy(c))


def _check_all_maps(E, c, levels=4):
    g, o = _both(E, c)
    gm, om = g.coordinate_manager, o.coordinate_manager
    gk, ok = g.coordinate_map_key, o.coordinate_map_key
    assert np.array_equal(g.C.cpu().numpy(), o.C.numpy())                      # row order preserved / first occurrence
    assert np.array_equal(g.inverse_mapping.cpu().numpy(), o.inverse_mapping)
    assert np.array_equal(g.unique_index.cpu().numpy(), o.unique_index)
    for _ in range(levels + 1):
        # 3^3 map on this level: identical tables row for row (orders agree), plus canonical-form check
        gkm = gm.kernel_map(gk, gk, [3, 3, 3], [1, 1, 1])
        okm = om.kernel_map(ok, ok, [3, 3, 3], [1, 1, 1])
        gp = gm.kernel_map_pairs(gkm)
        for a, b in zip(pairs_array(gp), pairs_array(okm)):
            assert np.array_equal(a, b)
        assert gkm.counts.cpu().tolist() == [len(i) for i, _ in okm]
        if _ == levels:
            break
        gk2, ok2 = gm.stride(gk, 2), om.stride(ok, 2)
        gc, oc = gm.get_coordinates(gk2).cpu().numpy(), om.get_coordinates(ok2)
        assert np.array_equal(gc, oc)                                           # strided map, first-occurrence order
        gd = gm.kernel_map(gk, gk2, [2, 2, 2], [1, 1, 1])
        od = om.kernel_map(ok, ok2, [2, 2, 2], [1, 1, 1])
        for a, b in zip(pairs_array(gm.kernel_map_pairs(gd)), pairs_array(od)):
            assert np.array_equal(a, b)
        gu = gm.kernel_map(gk2, gk, [2, 2, 2], [1, 1, 1], True)
        ou = om.kernel_map(ok2, ok, [2, 2, 2], [1, 1, 1], True)
        cin, cout = gm.get_coordinates(gk2).cpu().numpy(), gm.get_coordinates(gk).cpu().numpy()
        assert pair_set(gm.kernel_map_pairs(gu), cin, cout) == pair_set(ou, oc, om.get_coordinates(ok))
        # dgrad table of the strided conv is the transpose: (k,i)->o
        bt = gd.bwd_table.cpu().numpy()
        ft = gd.fwd_table.cpu().numpy()
        kk, oo = np.nonzero(ft >= 0)
        assert np.array_equal(bt[kk, ft[kk, oo]], oo) and (bt >= 0).sum() == (ft >= 0).sum()
        gk, ok = gk2, ok2


def test_dense_cube20_kat(E):
    c = dense_cube(20)
    g, _ = _both(E, c)
    m, k = g.coordinate_manager, g.coordinate_map_key
    sizes, pairs = [m.size(k)], []
    for lvl in range(5):
        pairs.append(int(m.kernel_map(k, k, [3, 3, 3], [1, 1, 1]).counts.sum().item()))
        if lvl < 4:
            k = m.stride(k, 2)
            sizes.append(m.size(k))
    assert sizes == [8000, 1000, 125, 27, 8]
    assert pairs == [195112, 21952, 2197, 343, 64]          # (3m-2)^3


@pytest.mark.parametrize("seed,n,extent,dups", [(0, 3000, 24, 0), (1, 20000, 64, 0), (2, 500, 6, 300), (3, 1, 4, 0),
                                               (4, 40, 200, 5)])
def test_random_sparse_maps(E, seed, n, extent, dups):
    rng = np.random.default_rng(seed)
    _check_all_maps(E, random_sparse_coords(rng, n, extent=extent, batches=3, duplicates=dups))


def test_scene_150k_maps_canonical(E):
    from languagegroundedsemseg_b200 import scenes
    c, _, _ = scenes.synthetic_voxel_scene(0, 150000)
    c[:, 1:] += np.array([37, 91, 5], np.int32)              # the trainer's random translation (pl_BaselineTrainer.py:294)
    _check_all_maps(E, c, levels=4)


def test_symmetry_property_full_size(E):
    """size-independent property: M_{K-1-k} is M_k with (in,out) swapped, on a 600K-voxel scene"""
    from languagegroundedsemseg_b200 import scenes
    c, _, _ = scenes.synthetic_voxel_scene(1, 600000, voxel_size=0.01)
    g = E.SparseTensor(torch.zeros(c.shape[0], 1).cuda(), torch.from_numpy(c).cuda())
    m, k = g.coordinate_manager, g.coordinate_map_key
    t = m.kernel_map(k, k, [3, 3, 3], [1, 1, 1]).fwd_table
    n = t.shape[1]
    assert torch.equal(t[13], torch.arange(n, device="cuda", dtype=torch.int32))
    for kk in (0, 4, 12):
        o = torch.nonzero(t[kk] >= 0).squeeze(1)
        i = t[kk][o].long()
        assert torch.equal(t[26 - kk][i].long(), o)
        assert int((t[kk] >= 0).sum()) == int((t[26 - kk] >= 0).sum())


def test_empty_and_errors(E):
    from languagegroundedsemseg_b200 import _lib
    g = E.SparseTensor(torch.zeros(0, 3).cuda(), torch.zeros((0, 4), dtype=torch.int32).cuda())
    assert g.C.shape == (0, 4)
    big = torch.tensor([[0, 1 << 20, 0, 0]], dtype=torch.int32).cuda()
    with pytest.raises(_lib.EngineError) as e:
        E.SparseTensor(torch.zeros(1, 3).cuda(), big)
    assert e.value.code == _lib.E_RANGE


def test_sparse_quantize_and_voxelize(E, golden_dir):
    import os
    from languagegroundedsemseg_b200 import voxelizer
    from oracle import me_cpu, voxelize_cpu
    rng = np.random.default_rng(3)
    c = rng.integers(-5, 5, (4000, 3)).astype(np.float64) + rng.random((4000, 3)) * 0.9
    c[:3] = 0.0
    uc, idx = E.utils.sparse_quantize(c, return_index=True)
    ouc, oidx = me_cpu.utils.sparse_quantize(c, return_index=True)
    assert np.array_equal(uc, ouc) and np.array_equal(idx, oidx) and idx[0] == 0
    g = np.load(os.path.join(golden_dir, "voxelize.npz"))
    coords, uidx, inv = voxelizer.voxelize(torch.from_numpy(g["pts"]).cuda(), g["M"], batch_index=2)
    assert np.array_equal(coords[:, 1:].cpu().numpy(), g["coords"]) and torch.all(coords[:, 0] == 2)
    q, ou, oinv = voxelize_cpu.voxelize(g["pts"], g["M"])
    assert np.array_equal(uidx.cpu().numpy(), ou) and np.array_equal(inv.cpu().numpy(), oinv)
    # full size: 0.9M raw points -> ~150K voxels, idempotence + sortedness
    from languagegroundedsemseg_b200 import scenes
    xyz, _, _ = scenes.synthetic_scene(0)
    M = np.eye(4)
    M[:3, :3] *= 50.0
    coords, uidx, inv = voxelizer.voxelize(torch.from_numpy(xyz).cuda(), M)
    assert torch.all(uidx[1:] > uidx[:-1])
    q2, u2, _ = voxelize_cpu.voxelize(xyz, M)
    assert np.array_equal(coords[:, 1:].cpu().numpy(), q2) and np.array_equal(uidx.cpu().numpy(), u2)
    c2, u3, _ = voxelizer.voxelize(coords[:, 1:].float().add(0.5).div(50.0), M)
    assert torch.equal(c2, coords) and torch.equal(u3, torch.arange(coords.shape[0], device="cuda"))
