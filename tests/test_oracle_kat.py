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


def _densify(coords, feats, lo, size):
    """[n, 4] (b, x, y, z) sparse rows -> dense [B, C, Z, Y, X] grid (zeros elsewhere), grid origin `lo`"""
    b = int(coords[:, 0].max()) + 1
    g = torch.zeros(b, feats.shape[1], size, size, size, dtype=feats.dtype)
    x, y, z = (torch.from_numpy(coords[:, i] - lo).long() for i in (1, 2, 3))
    g[torch.from_numpy(coords[:, 0]).long(), :, z, y, x] = feats
    return g


def test_conv3_equals_dense_conv3d_at_active_sites():
    """An anchor that does not come from this repository's reading of MinkowskiEngine: a stride-1 3x3x3 sparse convolution is
    torch's DENSE conv3d (cross-correlation, zero padding) sampled at the active voxels, with W[k] at kernel position
    (kz, ky, kx) = (k // 9, (k // 3) % 3, k % 3) — ME's x-fastest offset enumeration with centred odd kernels (App. A.5).
    Values, input gradient and weight gradient, fp64."""
    rng = np.random.default_rng(3)
    c = random_sparse_coords(rng, 500, extent=9, batches=2)
    cin, cout = 5, 4
    f = torch.from_numpy(rng.standard_normal((c.shape[0], cin))).double().requires_grad_(True)
    conv = _conv(cin, cout, 3, bias=True).double()
    y = conv(_st(c, f))
    assert torch.equal(y.C, torch.from_numpy(c))
    gy = torch.from_numpy(rng.standard_normal(tuple(y.F.shape)))
    y.F.backward(gy)
    lo, size = int(c[:, 1:].min()), int(c[:, 1:].max() - c[:, 1:].min()) + 1
    f2 = f.detach().clone().requires_grad_(True)
    w = conv.kernel.detach().clone().requires_grad_(True)                       # [27, cin, cout]
    wd = w.reshape(3, 3, 3, cin, cout).permute(4, 3, 0, 1, 2)                    # [cout, cin, kz, ky, kx]
    dense = torch.nn.functional.conv3d(_densify(c, f2, lo, size), wd, conv.bias.detach().reshape(-1), padding=1)
    x_, y_, z_ = (torch.from_numpy(c[:, i] - lo).long() for i in (1, 2, 3))
    ref = dense[torch.from_numpy(c[:, 0]).long(), :, z_, y_, x_]
    assert torch.allclose(y.F, ref, rtol=1e-12, atol=1e-12)
    ref.backward(gy)
    assert torch.allclose(f.grad, f2.grad, rtol=1e-11, atol=1e-12)
    assert torch.allclose(conv.kernel.grad, w.grad, rtol=1e-11, atol=1e-12)


def test_strided_conv2_equals_dense_conv3d_stride2():
    """2x2x2 stride-2 down-sampling: coarse voxel = floor(c / 2) * 2, fine voxel at offset (dx, dy, dz) in {0, 1}^3 from it
    contributes through W[dx + 2 dy + 4 dz] (even kernels are not centred, App. A.5) = dense conv3d with stride 2, no padding,
    on a grid whose origin is even; sampled at the coarse map's coordinates / 2"""
    rng = np.random.default_rng(5)
    c = random_sparse_coords(rng, 400, extent=10, batches=1)
    c[:, 1:] -= c[:, 1:].min()                                                   # origin 0 (even)
    cin, cout = 3, 6
    f = torch.from_numpy(rng.standard_normal((c.shape[0], cin))).double()
    conv = _conv(cin, cout, 2, stride=2).double()
    y = conv(_st(c, f))
    size = int(c[:, 1:].max()) + 2
    size += size % 2
    wd = conv.kernel.detach().reshape(2, 2, 2, cin, cout).permute(4, 3, 0, 1, 2)
    dense = torch.nn.functional.conv3d(_densify(c, f, 0, size), wd, stride=2)
    oc = y.C.numpy()
    assert (oc[:, 1:] % 2 == 0).all() and y.tensor_stride == [2, 2, 2]
    ref = dense[torch.from_numpy(oc[:, 0]).long(), :, torch.from_numpy(oc[:, 3] // 2).long(), torch.from_numpy(oc[:, 2] // 2).long(),
                torch.from_numpy(oc[:, 1] // 2).long()]
    assert torch.allclose(y.F.detach(), ref, rtol=1e-12, atol=1e-12)
    # every non-zero cell of the dense result is an active coarse voxel and vice versa (the coarse map is exactly the support)
    support = (_densify(c, torch.ones(c.shape[0], 1, dtype=torch.float64), 0, size).reshape(1, 1, size // 2, 2, size // 2, 2, size // 2, 2)
               .amax((3, 5, 7)) > 0).sum().item()
    assert support == oc.shape[0]


def test_transposed_conv2_equals_dense_conv_transpose3d():
    """2x2x2 stride-2 up-sampling onto the encoder's fine map: out[p] = in[parent(p)] . W[k(p)], k(p) = p's offset inside its
    parent, x fastest = dense conv_transpose3d with stride 2 (every fine cell receives exactly one product) sampled at the
    fine map's voxels"""
    rng = np.random.default_rng(8)
    c = random_sparse_coords(rng, 300, extent=8, batches=1)
    c[:, 1:] -= c[:, 1:].min()
    cmid, cout = 4, 3
    x = _st(c, torch.from_numpy(rng.standard_normal((c.shape[0], 2))).double())
    down = _conv(2, cmid, 2, stride=2).double()
    up = _conv(cmid, cout, 2, stride=2, transpose=True).double()
    mid = down(x)
    y = up(mid)
    assert torch.equal(y.C, x.C) and y.tensor_stride == [1, 1, 1]                 # lands on the encoder's map, same row order
    mc = mid.C.numpy()
    size = int(c[:, 1:].max()) + 2
    size += size % 2
    coarse = torch.zeros(1, cmid, size // 2, size // 2, size // 2, dtype=torch.float64)
    iz, iy, ix = (torch.from_numpy(mc[:, i] // 2).long() for i in (3, 2, 1))
    coarse[0, :, iz, iy, ix] = mid.F.detach().t()
    wd = up.kernel.detach().reshape(2, 2, 2, cmid, cout).permute(3, 4, 0, 1, 2)   # [c_in, c_out, kz, ky, kx]
    dense = torch.nn.functional.conv_transpose3d(coarse, wd, stride=2)
    ref = dense[0, :, torch.from_numpy(c[:, 3]).long(), torch.from_numpy(c[:, 2]).long(), torch.from_numpy(c[:, 1]).long()].t()
    assert torch.allclose(y.F.detach(), ref, rtol=1e-12, atol=1e-12)
