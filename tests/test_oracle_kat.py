"""Known-answer tests that pin the oracle — CPU only."""
import numpy as np
import pytest
import torch

from oracle import me_cpu as ME
from oracle import voxelize_cpu
from tests.helpers import dense_cube, random_sparse_coords


def _st(coords, feats):
    return ME.SparseTensor(feats, torch.from_numpy(coords))


def _conv(cin, cout, ks, stride=1, transpose=False, bias=False):
    cls = ME.MinkowskiConvolutionTranspose if transpose else ME.MinkowskiConvolution
    return cls(cin, cout, kernel_size=ks, stride=stride, dilation=1, bias=bias, dimension=3)


@pytest.mark.parametrize("n", [4, 7, 20])
def test_dense_cube_pair_counts(n):
    c = dense_cube(n)
    x = _st(c, torch.ones(n ** 3, 1))
    conv = _conv(1, 1, 3)
    conv(x)
    mgr = x.coordinate_manager
    km = mgr.kernel_map(x.coordinate_map_key, x.coordinate_map_key, [3, 3, 3], [1, 1, 1])
    assert sum(len(i) for i, _ in km) == (3 * n - 2) ** 3
    assert len(km[13][0]) == n ** 3
    for k, (ii, _) in enumerate(km):
        d = [k % 3 - 1, (k // 3) % 3 - 1, k // 9 - 1]
        assert len(ii) == np.prod([n - abs(a) for a in d])


def test_level_sizes_cube20():
    c = dense_cube(20)
    x = _st(c, torch.ones(8000, 1))
    mgr, key = x.coordinate_manager, x.coordinate_map_key
    sizes = [8000]
    for _ in range(4):
        key = mgr.stride(key, 2)
        sizes.append(mgr.size(key))
    assert sizes == [8000, 1000, 125, 27, 8]


def test_neighbour_count_conv():
    n = 6
    c = dense_cube(n)
    x = _st(c, torch.ones(n ** 3, 1))
    conv = _conv(1, 1, 3)
    with torch.no_grad():
        conv.kernel.fill_(1.0)
    out = conv(x).F.detach().squeeze(1).numpy()
    border = ((c[:, 1:] == 0) | (c[:, 1:] == n - 1)).sum(1)
    expect = np.array([27, 18, 12, 8])[border]
    assert np.array_equal(out, expect)


def test_centre_identity_and_row_order():
    rng = np.random.default_rng(0)
    c = random_sparse_coords(rng, 500)
    f = torch.randn(c.shape[0], 5)
    x = _st(c, f)
    assert np.array_equal(x.C.numpy(), c)  # duplicate-free input keeps its row order
    conv = _conv(5, 5, 3)
    with torch.no_grad():
        conv.kernel.zero_()
        conv.kernel[13] = torch.eye(5)
    assert torch.equal(conv(x).F, f)


@pytest.mark.parametrize("k", [0, 5, 14, 26])
def test_one_hot_offset(k):
    rng = np.random.default_rng(1)
    c = random_sparse_coords(rng, 800, extent=10, batches=1)
    f = torch.randn(c.shape[0], 2)
    x = _st(c, f)
    conv = _conv(2, 2, 3)
    with torch.no_grad():
        conv.kernel.zero_()
        conv.kernel[k] = torch.eye(2)
    out = conv(x).F
    off = np.array([k % 3 - 1, (k // 3) % 3 - 1, k // 9 - 1])  # x fastest
    lut = {tuple(r): i for i, r in enumerate(c)}
    for o in range(c.shape[0]):
        q = c[o].copy()
        q[1:] += off
        i = lut.get(tuple(q))
        expect = f[i] if i is not None else torch.zeros(2)
        assert torch.equal(out[o], expect)


def test_down_up_identities():
    n = 8
    c = dense_cube(n)
    x = _st(c, torch.ones(n ** 3, 1))
    down, up = _conv(1, 1, 2, 2), _conv(1, 1, 2, 2, transpose=True)
    with torch.no_grad():
        down.kernel.fill_(1.0)
        up.kernel.fill_(1.0)
    y = down(x)
    assert y.tensor_stride == [2, 2, 2] and torch.all(y.F == 8)
    z = up(y)
    assert z.coordinate_map_key == x.coordinate_map_key and torch.all(z.F == 8)
    mgr = x.coordinate_manager
    kd = mgr.kernel_map(x.coordinate_map_key, y.coordinate_map_key, [2, 2, 2], [1, 1, 1])
    ku = mgr.kernel_map(y.coordinate_map_key, x.coordinate_map_key, [2, 2, 2], [1, 1, 1], True)
    assert sum(len(i) for i, _ in kd) == n ** 3 == sum(len(i) for i, _ in ku)


def test_transpose_lands_on_encoder_map_random():
    rng = np.random.default_rng(2)
    c = random_sparse_coords(rng, 1500, extent=20)
    x = _st(c, torch.randn(c.shape[0], 3))
    y = _conv(3, 4, 2, 2)(x)
    y2 = _conv(4, 4, 2, 2)(y)
    z = _conv(4, 3, 2, 2, transpose=True)(y2)
    assert z.coordinate_map_key == y.coordinate_map_key
    assert set(map(tuple, z.C.numpy())) == set(map(tuple, y.C.numpy()))
    # strided coords: distinct floor(c/2)*2 including negatives
    expect = {(r[0], *(np.floor_divide(r[1:], 2) * 2)) for r in c}
    assert set(map(tuple, y.C.numpy())) == expect


def test_sparse_quantize_first_occurrence():
    rng = np.random.default_rng(3)
    c = rng.integers(-5, 5, (400, 3)).astype(np.float64) + rng.random((400, 3)) * 0.9
    c[:3] = 0.0
    uc, idx = ME.utils.sparse_quantize(c, return_index=True)
    assert idx[0] == 0 and np.all(np.diff(idx) > 0)
    q = np.floor(c).astype(np.int32)
    assert np.array_equal(idx, np.sort(np.unique(q, axis=0, return_index=True)[1]))
    assert np.array_equal(uc, q[idx])


def test_duplicates_keep_first_row():
    c = np.array([[0, 1, 1, 1], [0, 2, 2, 2], [0, 1, 1, 1], [0, 3, 3, 3], [0, 2, 2, 2]], np.int32)
    f = torch.arange(5.0)[:, None]
    x = _st(c, f)
    assert np.array_equal(x.C.numpy(), c[[0, 1, 3]])
    assert torch.equal(x.F.squeeze(1), torch.tensor([0.0, 1.0, 3.0]))
    assert np.array_equal(x.inverse_mapping, [0, 1, 0, 2, 1])


def test_sparse_collate():
    a, b = np.zeros((3, 3), np.int32), np.ones((2, 3), np.int32)
    bc, f, l = ME.utils.sparse_collate([a, b], [torch.zeros(3, 2), torch.ones(2, 2)], [torch.zeros(3), torch.ones(2)])
    assert bc.dtype == torch.int32 and bc[:, 0].tolist() == [0, 0, 0, 1, 1] and f.shape == (5, 2) and l.shape == (5,)


def test_conv_gradcheck_fp64():
    rng = np.random.default_rng(4)
    occ = rng.random((4, 4, 4)) < 0.6
    xyz = np.argwhere(occ)
    c = np.concatenate([np.zeros((xyz.shape[0], 1), int), xyz], 1).astype(np.int32)
    x = _st(c, torch.zeros(c.shape[0], 3))
    mgr, key = x.coordinate_manager, x.coordinate_map_key
    km = mgr.kernel_map(key, key, [3, 3, 3], [1, 1, 1])
    f = torch.randn(c.shape[0], 3, dtype=torch.float64, requires_grad=True)
    w = torch.randn(27, 3, 5, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda a, b: ME.sparse_conv(a, b, km, c.shape[0]), (f, w))


def test_voxelize_oracle_matches_numpy_matmul():
    rng = np.random.default_rng(6)
    pts = (rng.random((3000, 3)) * 4 - 2).astype(np.float32)
    M = np.eye(4)
    M[:3, :3] = np.diag([50.0, 50.0, 50.0]) @ np.array([[0.8, -0.6, 0], [0.6, 0.8, 0], [0, 0, 1.0]])
    M[:3, 3] = [0.3, -1.2, 2.0]
    q = voxelize_cpu.affine_floor(pts, M)
    homo = np.hstack((pts, np.ones((3000, 1), dtype=pts.dtype)))
    ref = np.floor(homo @ M.T[:, :3])        # the reference's expression, lib/voxelizer.py:138-139
    assert (q != ref.astype(np.int32)).sum() <= 2   # only points within an ulp of a voxel face may differ
