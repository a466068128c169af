"""On-device voxelisation: float64 affine + floor + first-occurrence de-duplication.
Replaces the arithmetic of lib/voxelizer.py:129-149 (single-view branch); the transformation matrix itself
(get_transformation_matrix, :44-74) stays host-side numpy in the caller."""
import ctypes

import numpy as np
import torch

from . import _lib
from .minkowski import _build_coordmap, _stream


def voxelize(xyz: torch.Tensor, M, batch_index: int = 0):
    """xyz [N,3] float32 CUDA; M 4x4 (or 3x4) float64 host matrix (voxelizer.py:138: floor([xyz,1] @ M^T[:, :3])).
    -> (coords int32 [N',4] (batch,x,y,z), unique_index int64 [N'] ascending = first point per voxel, inverse [N])"""
    lib = _lib.load()
    if not xyz.is_cuda:
        raise RuntimeError("lgs_b200 has no CPU path: voxelize needs a CUDA tensor")
    xyz = xyz.float().contiguous()
    n = xyz.shape[0]
    m = np.ascontiguousarray(np.asarray(M, dtype=np.float64)[:3, :4])
    coords = torch.empty((n, 4), dtype=torch.int32, device=xyz.device)
    _lib.check(lib.lgs_voxelize_affine(_lib.ptr(xyz), n, m.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                       int(batch_index), _lib.ptr(coords), _stream()))
    cm, uidx, inv = _build_coordmap(coords, 1, True)
    return cm.coords, uidx.long(), inv.long()
