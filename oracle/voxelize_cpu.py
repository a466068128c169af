"""ORACLE (test infrastructure, NOT product code) — CPU restatement of single-view voxelisation.

Follows /root/reference/lib/voxelizer.py:129-149: ``coords_aug = floor([xyz,1] @ M^T[:, :3])`` in float64,
then ``ME.utils.sparse_quantize(coords_aug, return_index=True)`` = keep the FIRST point (file order) per voxel,
kept rows in ascending original index.

The affine is evaluated in a fixed order without FMA contraction — ((x*m0 + y*m1) + z*m2) + m3 — so that the
CUDA kernel can match bit for bit; numpy's BLAS matmul in the reference may associate differently, which can
only matter for points landing within 1 ulp of a voxel face.  Pinned against the reference's own
``Voxelizer.voxelize`` (imported from /root/reference in the build container) in tests/golden/make_golden.py.
"""
import numpy as np

from .me_cpu import _encode, first_occurrence_unique


def affine_floor(xyz, M):
    """xyz [N,3] float; M [4,4] float64 homogeneous (voxelizer.py:44-74).  -> int32 [N,3]."""
    p = np.asarray(xyz, dtype=np.float64)
    M = np.asarray(M, dtype=np.float64)
    out = np.empty((p.shape[0], 3), np.float64)
    for j in range(3):
        out[:, j] = ((p[:, 0] * M[j, 0] + p[:, 1] * M[j, 1]) + p[:, 2] * M[j, 2]) + M[j, 3]
    return np.floor(out).astype(np.int32)


def voxelize(xyz, M):
    """-> (voxel coords int32 [N',3], unique_index int64 [N'] ascending = first point of each voxel)."""
    q = affine_floor(xyz, M)
    q4 = np.concatenate([np.zeros((q.shape[0], 1), np.int32), q], 1)
    uidx, inv = first_occurrence_unique(_encode(q4))
    return q[uidx], uidx, inv
