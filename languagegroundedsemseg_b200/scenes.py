"""Synthetic ScanNet-shaped scenes: a room shell (floor + 4 partial walls) and boxes,
sampled at ~4 points per voxel face with Gaussian jitter, to be quantised at 2 cm.  numpy only (host-side data
synthesis is outside the timed path)."""
import numpy as np


def _rect(rng, origin, u, v, density):
    """points on the parallelogram origin + a*u + b*v, a,b in [0,1]; density = points per m^2."""
    area = np.linalg.norm(np.cross(u, v))
    n = max(int(area * density), 1)
    a, b = rng.random(n), rng.random(n)
    return origin[None, :] + a[:, None] * u[None, :] + b[:, None] * v[None, :]


def synthetic_scene(seed=0, voxel_size=0.02, room=(2.6, 2.3, 2.2), n_boxes=6, points_per_voxel_face=4.0,
                    jitter_voxels=0.2, scale=1.0):
    """-> (xyz float32 [P,3] metres, rgb float32 [P,3] in [0,255], labels int32 [P] in [-1,200))."""
    rng = np.random.default_rng(seed)
    L, W, H = (room[0] * scale, room[1] * scale, room[2] * scale)
    density = points_per_voxel_face / (voxel_size * voxel_size)
    ex, ey, ez = np.eye(3)
    parts = [_rect(rng, np.zeros(3), ex * L, ey * W, density)]
    walls = [(np.zeros(3), ex * L), (np.array([0, W, 0.0]), ex * L), (np.zeros(3), ey * W),
             (np.array([L, 0, 0.0]), ey * W)]
    for o, u in walls:
        h = H * rng.uniform(0.55, 1.0)
        cov = rng.uniform(0.6, 0.95)
        start = rng.uniform(0, 1 - cov)
        parts.append(_rect(rng, o + u * start, u * cov, ez * h, density))
    for _ in range(n_boxes):
        sz = rng.uniform([0.3, 0.3, 0.3], [1.0, 0.9, 1.1]) * scale
        p = np.array([rng.uniform(0, L - sz[0]), rng.uniform(0, W - sz[1]), 0.0])
        sx, sy, sz_ = ex * sz[0], ey * sz[1], ez * sz[2]
        parts += [_rect(rng, p + sz_, sx, sy, density), _rect(rng, p, sx, sz_, density),
                  _rect(rng, p + sy, sx, sz_, density), _rect(rng, p, sy, sz_, density),
                  _rect(rng, p + sx, sy, sz_, density)]
    xyz = np.concatenate(parts, 0)
    xyz += rng.normal(0, jitter_voxels * voxel_size, xyz.shape)
    n = xyz.shape[0]
    rgb = rng.uniform(0, 255, (n, 3)).astype(np.float32)
    labels = rng.integers(0, 200, n).astype(np.int32)
    labels[rng.random(n) < 0.1] = -1
    return xyz.astype(np.float32), rgb, labels


def quantise_numpy(xyz, voxel_size=0.02):
    """Host-side first-occurrence quantisation (numpy) used only to *prepare* synthetic inputs of a known size."""
    q = np.floor(xyz.astype(np.float64) / voxel_size).astype(np.int64)
    key = (q[:, 0] + (1 << 20)) * (1 << 42) + (q[:, 1] + (1 << 20)) * (1 << 21) + (q[:, 2] + (1 << 20))
    _, first = np.unique(key, return_index=True)
    first.sort()
    return q[first].astype(np.int32), first


def synthetic_voxel_scene(seed=0, target_voxels=150_000, voxel_size=0.02, batch_index=0, tol=0.05):
    """Scene whose quantised size is target_voxels +-tol: the room footprint is scaled, a few secant steps.
    -> coords int32 [N,4], feats float32 [N,3] (rgb/255-0.5), labels int64 [N]"""
    scale = 1.0
    for _ in range(8):
        xyz, rgb, lab = synthetic_scene(seed, voxel_size, scale=scale)
        q, first = quantise_numpy(xyz, voxel_size)
        if abs(q.shape[0] - target_voxels) <= tol * target_voxels:
            break
        scale *= (target_voxels / q.shape[0]) ** 0.5   # every surface scales with scale^2
    coords = np.concatenate([np.full((q.shape[0], 1), batch_index, np.int32), q], 1)
    feats = (rgb[first] / 255.0 - 0.5).astype(np.float32)
    return coords, feats, lab[first].astype(np.int64)
