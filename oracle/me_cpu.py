"""ORACLE (test infrastructure, NOT product code) — CPU restatement of MinkowskiEngine 0.5.4 semantics.

PARITY UNPINNED: the arithmetic of this path lives in the un-vendored third-party dependency
``MinkowskiEngine==0.5.4`` (pin: /root/reference/config/lg_semseg.yml:204, install note README.md:46-53).
It is absent from /root/reference, not installed, and the reference holds no tests / golden vectors for it
(SURVEY.md §4).  This file restates ME's *published* CPU algorithm and is pinned
only by (i) the closed-form known-answer tests of SURVEY.md Appendix C (tests/test_oracle_kat.py) and
(ii) driving the reference's own unmodified ``models/`` package through it (tests/golden/make_golden.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may
import this module.  The product package ``languagegroundedsemseg_b200`` never does.

What is restated, and the reference call site that fixes which semantics matter:
  * SparseTensor(features, coordinates)         lib/train_test/pl_BaselineTrainer.py:300
  * KernelGenerator / MinkowskiConvolution      models/modules/common.py:179-203
  * MinkowskiConvolutionTranspose               models/modules/common.py:206-236
  * MinkowskiBatchNorm(.bn) / ReLU / cat / +=   models/modules/common.py:17-19, models/res16unet.py:194,237,
                                                 models/modules/resnet_block.py:54, models/resnet.py:78-82
  * utils.sparse_quantize / sparse_collate      lib/voxelizer.py:142, lib/transforms.py:421

Algorithm = ME's CPU path: first-occurrence-ordered coordinate maps, per-kernel-offset
``index_select -> mm -> index_add_`` convolution (autograd gives dgrad / wgrad of exactly that formula).
"""
from __future__ import annotations

import collections
import collections.abc
import sys
import types
from enum import Enum

import numpy as np
import torch
import torch.nn as nn

# Python-3.12 compatibility for the 2021-era reference (models/modules/common.py:81 uses collections.Sequence).
for _n in ("Sequence", "Iterable"):
    if not hasattr(collections, _n):
        setattr(collections, _n, getattr(collections.abc, _n))


# ----------------------------------------------------------------------------------------------------------
# enums / kernel generator  (Appendix A.5)
# ----------------------------------------------------------------------------------------------------------
class RegionType(Enum):
    HYPER_CUBE = 0
    HYPER_CROSS = 1
    CUSTOM = 2


def _as_list(v, D):
    if isinstance(v, torch.Tensor):
        v = v.tolist()
    if isinstance(v, (list, tuple)):
        assert len(v) == D
        return [int(x) for x in v]
    return [int(v)] * D


class KernelGenerator:
    def __init__(self, kernel_size=-1, stride=1, dilation=1, is_transpose=False, region_type=RegionType.HYPER_CUBE,
                 region_offsets=None, expand_coordinates=False, axis_types=None, dimension=-1):
        assert dimension > 0
        self.dimension = dimension
        self.kernel_size = _as_list(kernel_size, dimension)
        self.kernel_stride = _as_list(stride, dimension)
        self.kernel_dilation = _as_list(dilation, dimension)
        self.region_type = region_type
        self.region_offsets = region_offsets
        self.axis_types = axis_types
        if region_type != RegionType.HYPER_CUBE:
            raise NotImplementedError("oracle restates HYPER_CUBE only (all in-scope call sites, SURVEY App. B)")
        self.kernel_volume = int(np.prod(self.kernel_size))


def kernel_offsets(kernel_size, tensor_stride, dilation):
    """Offsets [K,3], x fastest (k = ix + ks*iy + ks^2*iz); odd ks centred, even ks starts at 0 (App. A.5)."""
    axes = []
    for ks, ts, d in zip(kernel_size, tensor_stride, dilation):
        r = np.arange(ks) - (ks // 2 if ks % 2 == 1 else 0)
        axes.append(r * d * ts)
    kz, ky, kx = np.meshgrid(axes[2], axes[1], axes[0], indexing="ij")
    return np.stack([kx.ravel(), ky.ravel(), kz.ravel()], 1).astype(np.int64)


# ----------------------------------------------------------------------------------------------------------
# coordinate manager  (Appendix A.2, A.6, A.7, A.11)
# ----------------------------------------------------------------------------------------------------------
_R = 1 << 20  # per-axis key radix; coordinates must lie in [-2^19, 2^19)


def _encode(c: np.ndarray) -> np.ndarray:
    c = c.astype(np.int64)
    assert c[:, 1:].min(initial=0) >= -(_R // 2) and c[:, 1:].max(initial=0) < _R // 2
    return ((c[:, 0] * _R + (c[:, 1] + _R // 2)) * _R + (c[:, 2] + _R // 2)) * _R + (c[:, 3] + _R // 2)


def first_occurrence_unique(keys: np.ndarray):
    """-> (unique_index ascending = first occurrence of each distinct key, inverse map row->unique row)."""
    _, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")  # unique rows in order of first appearance (ME CPU order)
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    return first[order], rank[inv.reshape(-1)]


class CoordinateMapKey:
    def __init__(self, tensor_stride, tag=""):
        self.tensor_stride = tuple(int(t) for t in tensor_stride)
        self.tag = tag

    def get_tensor_stride(self):
        return list(self.tensor_stride)

    def __eq__(self, o):
        return isinstance(o, CoordinateMapKey) and (self.tensor_stride, self.tag) == (o.tensor_stride, o.tag)

    def __hash__(self):
        return hash((self.tensor_stride, self.tag))

    def __repr__(self):
        return f"CoordinateMapKey(stride={list(self.tensor_stride)}, tag={self.tag!r})"


class CoordinateManager:
    def __init__(self, D=3):
        self.D = D
        self._coords = {}   # key -> np.int32 [N,4]
        self._sorted = {}   # key -> (sorted keys, argsort)
        self._kmaps = {}

    # --- maps ------------------------------------------------------------------------------------------
    def insert_and_map(self, coords: np.ndarray, tensor_stride=(1, 1, 1)):
        uidx, inv = first_occurrence_unique(_encode(coords))
        key = CoordinateMapKey(tensor_stride)
        self._coords[key] = np.ascontiguousarray(coords[uidx].astype(np.int32))
        return key, uidx, inv

    def get_coordinates(self, key):
        return self._coords[key]

    def size(self, key):
        return self._coords[key].shape[0]

    def stride(self, key, stride):
        s = _as_list(stride, self.D)
        new_ts = tuple(t * q for t, q in zip(key.tensor_stride, s))
        nkey = CoordinateMapKey(new_ts)
        if nkey not in self._coords:
            c = self._coords[key].astype(np.int64).copy()
            for a in range(self.D):
                c[:, 1 + a] = np.floor_divide(c[:, 1 + a], new_ts[a]) * new_ts[a]
            uidx, _ = first_occurrence_unique(_encode(c))
            self._coords[nkey] = np.ascontiguousarray(c[uidx].astype(np.int32))
        return nkey

    def key_with_stride(self, tensor_stride):
        k = CoordinateMapKey(tensor_stride)
        if k not in self._coords:
            raise RuntimeError(f"no coordinate map with tensor stride {list(tensor_stride)} in this manager")
        return k

    def _lookup(self, key):
        if key not in self._sorted:
            e = _encode(self._coords[key])
            o = np.argsort(e, kind="stable")
            self._sorted[key] = (e[o], o)
        return self._sorted[key]

    def kernel_map(self, in_key, out_key, kernel_size, dilation, is_transpose=False):
        """list over k of (in_idx, out_idx) int64 arrays; M_k = {(i,o): C_in[i] = C_out[o] + off_k} (App. A.6).
        Transposed: the map of the forward conv fine->coarse with (in,out) swapped (App. A.7)."""
        ck = (in_key, out_key, tuple(kernel_size), tuple(dilation), bool(is_transpose))
        if ck in self._kmaps:
            return self._kmaps[ck]
        if is_transpose:
            fwd = self.kernel_map(out_key, in_key, kernel_size, dilation, False)
            res = [(o, i) for (i, o) in fwd]
        else:
            offs = kernel_offsets(kernel_size, in_key.tensor_stride, dilation)
            sk, so = self._lookup(in_key)
            cout = self._coords[out_key].astype(np.int64)
            res = []
            for off in offs:
                q = cout.copy()
                q[:, 1:] += off[None, :]
                e = _encode(q)
                pos = np.searchsorted(sk, e)
                pos[pos >= sk.size] = 0
                hit = sk[pos] == e if sk.size else np.zeros(e.shape, bool)
                out_idx = np.nonzero(hit)[0]
                res.append((so[pos[hit]].astype(np.int64), out_idx.astype(np.int64)))
        self._kmaps[ck] = res
        return res


# ----------------------------------------------------------------------------------------------------------
# SparseTensor  (Appendix A.2, A.10)
# ----------------------------------------------------------------------------------------------------------
class SparseTensor:
    def __init__(self, features, coordinates=None, coordinate_map_key=None, coordinate_manager=None,
                 tensor_stride=1, device=None, **_ignored):
        if coordinate_map_key is None:
            assert coordinates is not None
            c = coordinates.detach().cpu().numpy() if isinstance(coordinates, torch.Tensor) else np.asarray(coordinates)
            c = np.floor(c).astype(np.int32) if c.dtype.kind == "f" else c.astype(np.int32)
            mgr = coordinate_manager or CoordinateManager(D=c.shape[1] - 1)
            key, uidx, inv = mgr.insert_and_map(c, _as_list(tensor_stride, c.shape[1] - 1))
            if uidx.size != c.shape[0]:  # duplicates: keep one row per coordinate (first; ME: RANDOM_SUBSAMPLE)
                features = features[torch.from_numpy(uidx)]
            self.unique_index, self.inverse_mapping = uidx, inv
            coordinate_map_key, coordinate_manager = key, mgr
        self._F = features
        self.coordinate_map_key = coordinate_map_key
        self.coordinate_manager = coordinate_manager

    @property
    def F(self):
        return self._F

    @property
    def feats(self):
        return self._F

    @property
    def C(self):
        return torch.from_numpy(self.coordinate_manager.get_coordinates(self.coordinate_map_key))

    coordinates = C

    @property
    def tensor_stride(self):
        return list(self.coordinate_map_key.tensor_stride)

    @property
    def device(self):
        return self._F.device

    @property
    def D(self):
        return self.coordinate_manager.D

    @property
    def shape(self):
        return self._F.shape

    def _same(self, o):
        if o.coordinate_map_key != self.coordinate_map_key or o.coordinate_manager is not self.coordinate_manager:
            raise RuntimeError("SparseTensor arithmetic needs identical coordinate_map_key (App. A.10)")

    def __add__(self, o):
        self._same(o)
        return SparseTensor(self._F + o._F, coordinate_map_key=self.coordinate_map_key,
                            coordinate_manager=self.coordinate_manager)

    def __iadd__(self, o):  # resnet_block.py:54  `out += residual`
        self._same(o)
        self._F = self._F + o._F
        return self

    def __len__(self):
        return self._F.shape[0]


def cat(*sts):
    """Column concat of same-key tensors in argument order (res16unet.py:237; App. A.10)."""
    for s in sts[1:]:
        sts[0]._same(s)
    return SparseTensor(torch.cat([s.F for s in sts], 1), coordinate_map_key=sts[0].coordinate_map_key,
                        coordinate_manager=sts[0].coordinate_manager)


# ----------------------------------------------------------------------------------------------------------
# modules  (Appendix A.6-A.10)
# ----------------------------------------------------------------------------------------------------------
class MinkowskiNetwork(nn.Module):
    def __init__(self, D):
        super().__init__()
        self.D = D


def sparse_conv(feats, kernel, kmap, n_out, bias=None):
    """out[o] = sum_k sum_{(i,o) in M_k} in[i] @ W[k]; ME CPU algorithm: gather -> GEMM -> scatter-add per offset."""
    out = feats.new_zeros((n_out, kernel.shape[-1]))
    for k, (ii, oo) in enumerate(kmap):
        if ii.size == 0:
            continue
        ii_t, oo_t = torch.from_numpy(ii), torch.from_numpy(oo)
        out = out.index_add(0, oo_t, feats.index_select(0, ii_t) @ kernel[k])
    if bias is not None:
        out = out + bias
    return out


class _ConvBase(nn.Module):
    TRANSPOSE = False

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, dimension=None):
        super().__init__()
        assert dimension is not None and dimension > 0
        if kernel_generator is None:
            kernel_generator = KernelGenerator(kernel_size, stride, dilation, dimension=dimension)
        self.kernel_generator = kernel_generator
        self.in_channels, self.out_channels, self.dimension = in_channels, out_channels, dimension
        self.use_mm = (not self.TRANSPOSE) and kernel_generator.kernel_volume == 1 and \
            all(s == 1 for s in kernel_generator.kernel_stride)
        K = kernel_generator.kernel_volume
        shape = (in_channels, out_channels) if self.use_mm else (K, in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):  # App. A.8
        with torch.no_grad():
            K = self.kernel_generator.kernel_volume
            n = (self.out_channels if self.TRANSPOSE else self.in_channels) * K
            stdv = 1.0 / np.sqrt(n)
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def forward(self, x: SparseTensor):
        mgr, kg = x.coordinate_manager, self.kernel_generator
        if self.use_mm:
            out = x.F @ self.kernel
            if self.bias is not None:
                out = out + self.bias
            return SparseTensor(out, coordinate_map_key=x.coordinate_map_key, coordinate_manager=mgr)
        in_key = x.coordinate_map_key
        if self.TRANSPOSE:
            ts = [t // s for t, s in zip(in_key.tensor_stride, kg.kernel_stride)]
            out_key = mgr.key_with_stride(ts)
        else:
            out_key = mgr.stride(in_key, kg.kernel_stride) if any(s > 1 for s in kg.kernel_stride) else in_key
        kmap = mgr.kernel_map(in_key, out_key, kg.kernel_size, kg.kernel_dilation, self.TRANSPOSE)
        out = sparse_conv(x.F, self.kernel, kmap, mgr.size(out_key), self.bias)
        return SparseTensor(out, coordinate_map_key=out_key, coordinate_manager=mgr)


class MinkowskiConvolution(_ConvBase):
    TRANSPOSE = False


class MinkowskiConvolutionTranspose(_ConvBase):
    TRANSPOSE = True


class MinkowskiBatchNorm(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)

    def forward(self, x: SparseTensor):
        return SparseTensor(self.bn(x.F), coordinate_map_key=x.coordinate_map_key,
                            coordinate_manager=x.coordinate_manager)


class MinkowskiSyncBatchNorm(MinkowskiBatchNorm):
    @classmethod
    def convert_sync_batchnorm(cls, module, process_group=None):  # main.py:122-123
        for name, child in list(module.named_children()):
            if isinstance(child, MinkowskiBatchNorm):
                child.bn = nn.SyncBatchNorm.convert_sync_batchnorm(child.bn, process_group)
            else:
                cls.convert_sync_batchnorm(child, process_group)
        return module


class MinkowskiReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()
        self.inplace = inplace

    def forward(self, x: SparseTensor):
        return SparseTensor(torch.relu(x.F), coordinate_map_key=x.coordinate_map_key,
                            coordinate_manager=x.coordinate_manager)


def _stub(name):
    class _S(nn.Module):
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name}: outside the hot path (SURVEY.md §8b, import-surface row)")
    _S.__name__ = name
    return _S


for _n in ("MinkowskiInstanceNorm", "MinkowskiSumPooling", "MinkowskiAvgPooling", "MinkowskiAvgUnpooling",
           "MinkowskiPoolingTranspose", "MinkowskiGlobalPooling", "MinkowskiBroadcastAddition",
           "MinkowskiBroadcastMultiplication", "MinkowskiLinear", "MinkowskiSigmoid", "MinkowskiMaxPooling",
           "MinkowskiGlobalMaxPooling", "MinkowskiDropout", "MinkowskiBroadcast", "MinkowskiConvolutionFunction"):
    globals()[_n] = _stub(_n)


def convert_to_int_tensor(arg, dimension):
    return torch.IntTensor(_as_list(arg, dimension))


def convert_region_type(*a, **k):
    raise NotImplementedError("ME 0.4-era API, outside the hot path")


# ----------------------------------------------------------------------------------------------------------
# utils  (Appendix A.3, A.4)
# ----------------------------------------------------------------------------------------------------------
def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False,
                    return_inverse=False, quantization_size=None, **_):
    """No-labels mode only: floor, keep first occurrence, ascending original order (lib/voxelizer.py:142)."""
    assert labels is None, "oracle restates the no-labels mode used by lib/voxelizer.py:142"
    is_t = isinstance(coordinates, torch.Tensor)
    c = coordinates.numpy() if is_t else np.asarray(coordinates)
    if quantization_size is not None:
        c = c / quantization_size
    q = np.floor(c).astype(np.int32)
    q4 = np.concatenate([np.zeros((q.shape[0], 1), np.int32), q], 1)
    uidx, inv = first_occurrence_unique(_encode(q4))
    uc = q[uidx]
    wrap = (lambda a: torch.from_numpy(a)) if is_t else (lambda a: a)
    ret = [wrap(uc)]
    if features is not None:
        ret.append(features[uidx])
    if return_index:
        ret.append(wrap(uidx))
    if return_inverse:
        ret.append(wrap(inv))
    return ret[0] if len(ret) == 1 else tuple(ret)


def batched_coordinates(coords_list, dtype=torch.int32):
    out = []
    for b, c in enumerate(coords_list):
        c = torch.as_tensor(c)
        out.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=dtype), torch.floor(c).to(dtype)], 1))
    return torch.cat(out, 0)


def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
    """(int32 [sum N, 1+D] with batch column prepended, cat(feats), cat(labels))  (lib/transforms.py:421)."""
    bc = batched_coordinates(coords, dtype)
    f = torch.cat([torch.as_tensor(x) for x in feats], 0)
    if labels is None:
        return bc, f
    return bc, f, torch.cat([torch.as_tensor(x) for x in labels], 0)


utils = types.ModuleType("MinkowskiEngine.utils")
utils.sparse_quantize = sparse_quantize
utils.sparse_collate = sparse_collate
utils.batched_coordinates = batched_coordinates

MinkowskiOps = types.ModuleType("MinkowskiEngine.MinkowskiOps")
MinkowskiOps.cat = cat

__version__ = "0.5.4-oracle"


def install(name="MinkowskiEngine"):
    """Register this module under the name the reference imports (models/modules/common.py:9)."""
    me = sys.modules[__name__]
    sys.modules[name] = me
    sys.modules[name + ".MinkowskiOps"] = MinkowskiOps
    sys.modules[name + ".utils"] = utils
    return me
