"""Host-side mirror of the reference-facing interface: the Python names of package ``MinkowskiEngine`` that
RozDavid/LanguageGroundedSemseg uses, backed by the lgs_b200 C-ABI CUDA engine.

    import languagegroundedsemseg_b200 as lgs
    lgs.install_as_minkowski()          # `import MinkowskiEngine as ME` in the reference now resolves here
    import models                        # /root/reference/models, unchanged

Reference call sites served (relative to /root/reference):
  SparseTensor                      lib/train_test/pl_BaselineTrainer.py:300, pl_RepresentationTrainer.py:183
  KernelGenerator, RegionType       models/modules/common.py:55-64, 192-193
  MinkowskiConvolution(/Transpose)  models/modules/common.py:195-203, 228-236
  MinkowskiBatchNorm (.bn)          models/modules/common.py:17-19, models/resnet.py:78-82
  MinkowskiReLU, cat, +=            models/res16unet.py:5-6,194,237; models/modules/resnet_block.py:54
  MinkowskiSyncBatchNorm            main.py:122-123
  utils.sparse_quantize / collate   lib/voxelizer.py:142, lib/transforms.py:421

All arithmetic runs in the CUDA library; there is no CPU path behind these names.
"""
from __future__ import annotations

import collections
import collections.abc
import ctypes
import os
import sys
import types
import weakref
from enum import Enum

import numpy as np
import torch
import torch.nn as nn

from .. import _lib

# Python-3.12 compatibility for the 2021-era reference callers (models/modules/common.py:81, lib/voxelizer.py:53)
for _n in ("Sequence", "Iterable"):
    if not hasattr(collections, _n):
        setattr(collections, _n, getattr(collections.abc, _n))

# ----------------------------------------------------------------------------------------------------------
# engine-wide settings
# ----------------------------------------------------------------------------------------------------------
# 'simt'  exact fp32 FMA kernels (parity anchor, slow)
# 'bx3'   tcgen05 tensor cores, parity-grade: bf16x3 error-compensated products for fp32 features (2^-16 per product,
#         fp32 accumulate; whole-net logits ~1e-4 of the fp32 reference); bf16 MMA for bf16 features             [default]
# 'tc'    tcgen05 3xTF32 products (2^-21 per product) at twice the tensor-core time and weight traffic of 'bx3'
# 'tf32'  tcgen05 single-pass TF32 (fast, ~7e-4 relative error per layer; not parity grade)
_ALGO = {"simt": _lib.ALGO_SIMT, "tc": _lib.ALGO_TC3, "tf32": _lib.ALGO_TC, "bx3": _lib.ALGO_BX3}
_state = {"algo": _ALGO[os.environ.get("LGS_CONV_ALGO", "bx3")], "profile": None, "fuse_bn": True,
          # conv -> BatchNorm (+ residual) (+ ReLU) as ONE autograd node (the conv is evaluated lazily); kernels unchanged
          "fuse_conv_bn": os.environ.get("LGS_FUSE_CONV_BN", "1") != "0",
          # backward: wgrad on a side stream while dgrad runs on the training stream (joined before the node returns)
          "overlap_wgrad": os.environ.get("LGS_OVERLAP_WGRAD", "1") != "0",
          # tensor-core operand forms of ALL layers' weights in one launch per parameter update
          "batch_prep": os.environ.get("LGS_BATCH_PREP", "1") != "0"}


def profile_begin():
    """Start recording one (kind, shape, kernel map, start/end CUDA events) entry per engine convolution launch —
    bench.py's live per-kernel timing.  Events are recorded on the launching stream."""
    _state["profile"] = []


def profile_end():
    rec, _state["profile"] = _state["profile"], None
    return rec


class _NoTimer:
    __slots__ = ()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NO_TIMER = _NoTimer()


class _Timed:
    """context manager: CUDA events around one C-ABI launch when profiling is on"""
    __slots__ = ("meta", "ev")

    def __new__(cls, *a):
        if _state["profile"] is None:
            return _NO_TIMER
        return object.__new__(cls)

    def __init__(self, kind, K, c_in, c_out, n_in, n_out, km, dtype):
        self.meta = (kind, K, c_in, c_out, n_in, n_out, km, dtype)

    def __enter__(self):
        if _state["profile"] is not None:
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.ev[0].record()
        return self

    def __exit__(self, *a):
        if _state["profile"] is not None:
            self.ev[1].record()
            _state["profile"].append((self.meta, self.ev))
        return False


def set_conv_algo(name: str):
    """'simt' exact fp32 FMA | 'bx3' tcgen05 bf16x3 (default, parity grade) | 'tc' tcgen05 3xTF32 | 'tf32' single-pass TF32."""
    _state["algo"] = _ALGO[name]


def get_conv_algo() -> str:
    return {v: k for k, v in _ALGO.items()}[_state["algo"]]


def _stream():
    # torch.cuda.current_stream() costs ~16 us per call (device-index plumbing); the raw getter is ~0.3 us
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


def set_conv_bn_fusion(flag: bool):
    """conv -> BatchNorm (+ residual) (+ ReLU) recorded as one autograd node (default on); same kernels either way."""
    _state["fuse_conv_bn"] = bool(flag)


def set_wgrad_overlap(flag: bool):
    """Backward: launch wgrad on a side stream next to dgrad (default on); joined before the backward node returns."""
    _state["overlap_wgrad"] = bool(flag)


def set_batched_weight_prep(flag: bool):
    """One lgs_weight_prep_batch launch per parameter update instead of one lgs_weight_prep per layer call."""
    _state["batch_prep"] = bool(flag)


_bn_scratch = {}
_BN_SCRATCH_DOUBLES = 16 * 1024


class _Scratch:
    """fp64 accumulator scratch of the BatchNorm kernels, one per (device, stream): two zero-initialised halves used
    alternately — every lgs_bn_fwd / lgs_bn_bwd call accumulates into the current half and clears the other one for the
    next call on that stream (launches on one stream are ordered), so no memset node is issued per BatchNorm."""
    __slots__ = ("buf", "cur", "halves")

    def __init__(self, device):
        self.buf = torch.zeros((2, _BN_SCRATCH_DOUBLES), dtype=torch.float64, device=device)
        base = self.buf.data_ptr()
        self.halves = (ctypes.c_void_p(base), ctypes.c_void_p(base + 8 * _BN_SCRATCH_DOUBLES))
        self.cur = 0

    def pair(self):
        """(accumulators, half to clear) as c_void_p"""
        return self.halves[self.cur], self.halves[1 - self.cur]

    def done(self, ok):
        if ok:
            self.cur = 1 - self.cur
        else:                    # a failed call may have left partial sums behind
            self.buf.zero_()


def _scratch64(dev_index):
    key = (dev_index, torch._C._cuda_getCurrentRawStream(dev_index))
    t = _bn_scratch.get(key)
    if t is None:
        t = _bn_scratch[key] = _Scratch(torch.device("cuda", dev_index))
    return t


_stream_objs = {}


def _cur_stream_obj(dev_index):
    """torch.cuda.Stream object of the current stream (cached by raw handle: torch.cuda.current_stream() costs ~16 us)"""
    raw = torch._C._cuda_getCurrentRawStream(dev_index)
    so = _stream_objs.get((dev_index, raw))
    if so is None:
        so = _stream_objs[(dev_index, raw)] = torch.cuda.current_stream(dev_index)
    return so


_side = {}


def _side_stream(dev_index):
    r = _side.get(dev_index)
    if r is None:
        with torch.cuda.device(dev_index):
            r = _side[dev_index] = (torch.cuda.Stream(), torch.cuda.Event(), torch.cuda.Event())
    return r


_tc_ok_cache = {}


def _tc_supported(lib, c_in, c_out, dt):
    key = (c_in, c_out, dt)
    r = _tc_ok_cache.get(key)
    if r is None:
        r = _tc_ok_cache[key] = bool(lib.lgs_conv_tc_supported(c_in, c_out, dt))
    return r


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.bfloat16:
        return _lib.BF16
    raise TypeError(f"engine features must be float32 or bfloat16, got {t.dtype}")


# ----------------------------------------------------------------------------------------------------------
# enums / kernel generator
# ----------------------------------------------------------------------------------------------------------
class RegionType(Enum):
    HYPER_CUBE = 0
    HYPER_CROSS = 1
    CUSTOM = 2


def _as_list(v, D):
    if isinstance(v, torch.Tensor):
        v = v.tolist()
    if isinstance(v, (list, tuple)):
        if len(v) != D:
            raise ValueError(f"expected {D} values, got {v}")
        return [int(x) for x in v]
    return [int(v)] * D


def _uniform(v, what):
    if any(x != v[0] for x in v):
        raise NotImplementedError(f"anisotropic {what} {v} is outside the hot path (SURVEY.md App. B: all isotropic)")
    return v[0]


class KernelGenerator:
    def __init__(self, kernel_size=-1, stride=1, dilation=1, is_transpose=False,
                 region_type=RegionType.HYPER_CUBE, region_offsets=None, expand_coordinates=False,
                 axis_types=None, dimension=-1):
        if dimension != 3:
            raise NotImplementedError("lgs_b200 implements D = 3 (all in-scope call sites)")
        if region_type != RegionType.HYPER_CUBE:
            raise NotImplementedError("lgs_b200 implements RegionType.HYPER_CUBE (SURVEY.md App. B)")
        self.dimension = dimension
        self.kernel_size = _as_list(kernel_size, dimension)
        self.kernel_stride = _as_list(stride, dimension)
        self.kernel_dilation = _as_list(dilation, dimension)
        self.region_type, self.region_offsets, self.axis_types = region_type, region_offsets, axis_types
        self.kernel_volume = int(np.prod(self.kernel_size))


# ----------------------------------------------------------------------------------------------------------
# coordinate manager
# ----------------------------------------------------------------------------------------------------------
class CoordinateMapKey:
    def __init__(self, tensor_stride, tag=""):
        self.tensor_stride = tuple(int(t) for t in tensor_stride)
        self.tag = tag
        self._hash = hash((self.tensor_stride, tag))

    def get_tensor_stride(self):
        return list(self.tensor_stride)

    def __eq__(self, o):
        return self is o or (isinstance(o, CoordinateMapKey) and self.tensor_stride == o.tensor_stride and self.tag == o.tag)

    def __hash__(self):
        return self._hash

    def __repr__(self):
        return f"CoordinateMapKey(stride={list(self.tensor_stride)}, tag={self.tag!r})"


class _CoordMap:
    """One coordinate map: rows (device int32 [n,4]) + its cuckoo table.  All buffers are torch tensors."""
    __slots__ = ("coords", "tkeys", "tvals", "n", "capacity")


_PLAN_STATUS = {"buf": None, "next": 0}          # pinned ring of plan-builder status words (cudaHostAlloc once, not per plan)


def _plan_status_slot():
    if _PLAN_STATUS["buf"] is None:
        _PLAN_STATUS["buf"] = torch.empty(256, 2, dtype=torch.int32).pin_memory()
    i = _PLAN_STATUS["next"]
    _PLAN_STATUS["next"] = (i + 1) % 256
    return _PLAN_STATUS["buf"][i]


class KernelMap:
    """Output-stationary neighbour tables of one (in map, out map, kernel) triple.
    fwd_table [K, n_out] rows of the input map; bwd_table [K, n_in] rows of the output map (dgrad)."""
    __slots__ = ("K", "n_in", "n_out", "fwd_table", "bwd_table", "bwd_reverse", "counts", "_plan", "_plan_status", "_plan_event",
                 "_plan_builder", "plan_stats")
    # plan: neighbourhood plan of a same-map 3^3 table (lgs_nbplan_build; None when not built / not usable).  Built with the
    # kernel map (on the staging stream, a step ahead) once some convolution has asked for one, lazily at the first request
    # before that (bf16 / 'tc' / 'simt' runs never ask and never pay for plans).  The builder's status (did a supertile overflow
    # the cache?) comes back through pinned memory behind an event and is only waited for at the first use of the plan —
    # normally a step later, so building plans never blocks the host.

    def __init__(self):
        self._plan = self._plan_status = self._plan_event = self._plan_builder = self.plan_stats = None

    @property
    def plan(self):
        if self._plan_builder is not None:
            build, self._plan_builder = self._plan_builder, None
            _state["plans_wanted"] = True
            build()
        if self._plan_event is not None:
            self._plan_event.synchronize()
            self._plan_event = None
            self.plan_stats = (int(self._plan_status[0]), int(self._plan_status[1]))
            if self.plan_stats[0] != 0:
                self._plan = None                 # a supertile touches more unique rows than the cache holds: table-driven kernel
        return self._plan


def _build_coordmap(coords: torch.Tensor, quant: int, want_maps: bool):
    lib = _lib.load()
    n = coords.shape[0]
    dev = coords.device
    cap = lib.lgs_hash_capacity(n)
    cm = _CoordMap()
    for attempt in range(3):
        cm.tkeys = torch.empty(cap, dtype=torch.int64, device=dev)
        cm.tvals = torch.empty(cap, dtype=torch.int32, device=dev)
        out_coords = torch.empty((n, 4), dtype=torch.int32, device=dev)
        uidx = torch.empty(n, dtype=torch.int32, device=dev) if want_maps else None
        inv = torch.empty(n, dtype=torch.int32, device=dev) if want_maps else None
        scratch = torch.empty(lib.lgs_coordmap_scratch_elems(n), dtype=torch.int32, device=dev)
        d_n = torch.empty(1, dtype=torch.int32, device=dev)
        h_n = ctypes.c_int64(0)
        rc = lib.lgs_coordmap_build(_lib.ptr(coords), n, quant, _lib.ptr(cm.tkeys), _lib.ptr(cm.tvals), cap,
                                    _lib.ptr(out_coords), _lib.ptr(uidx), _lib.ptr(inv), _lib.ptr(scratch),
                                    _lib.ptr(d_n), ctypes.byref(h_n), _stream())
        if rc == _lib.E_HASH_FULL and attempt < 2:
            cap *= 2
            continue
        _lib.check(rc)
        break
    cm.n, cm.capacity = int(h_n.value), cap
    cm.coords = out_coords[: cm.n]
    if want_maps:
        return cm, uidx[: cm.n], inv
    return cm


class CoordinateManager:
    # Map requests seen on earlier managers (tensor strides / kernel maps a network asks for, in order).  A strided map
    # build needs one host sync (the row count sizes the next buffers); done lazily inside the forward pass each sync
    # drains the launch queue.  New managers therefore pre-build everything previous ones were asked for right at
    # SparseTensor creation — ME's manager also caches per (key, stride, kernel) but builds on first use.
    _learned_strides = []      # [(from tensor stride, factor)]
    _learned_kmaps = []        # [(in stride, out stride, ks, dilation, transpose)]
    prebuild = True

    def __init__(self, D=3):
        if D != 3:
            raise NotImplementedError("lgs_b200 implements D = 3")
        self.D = D
        self._maps = {}
        self._kmaps = {}
        self._conv_cache = {}

    def _prebuild(self):
        if not CoordinateManager.prebuild:
            return
        for ts, factor in list(CoordinateManager._learned_strides):
            k = CoordinateMapKey(ts)
            if k in self._maps:
                self.stride(k, factor)
        for in_ts, out_ts, ks, dil, tr in list(CoordinateManager._learned_kmaps):
            ik, ok = CoordinateMapKey(in_ts), CoordinateMapKey(out_ts)
            if ik in self._maps and ok in self._maps:
                self.kernel_map(ik, ok, [ks] * 3, [dil] * 3, tr)

    # -- coordinate maps --------------------------------------------------------------------------------
    def insert_and_map(self, coords: torch.Tensor, tensor_stride=(1, 1, 1)):
        key = CoordinateMapKey(tensor_stride)
        cm, uidx, inv = _build_coordmap(coords, 1, True)
        self._maps[key] = cm
        self._prebuild()
        return key, uidx, inv

    def stride(self, key, stride):
        s = _as_list(stride, self.D)
        nkey = CoordinateMapKey([t * q for t, q in zip(key.tensor_stride, s)])
        req = (key.tensor_stride, tuple(s))
        if req not in CoordinateManager._learned_strides:
            CoordinateManager._learned_strides.append(req)
        if nkey not in self._maps:
            self._maps[nkey] = _build_coordmap(self._maps[key].coords, _uniform(list(nkey.tensor_stride), "stride"),
                                               False)
        return nkey

    def key_with_stride(self, tensor_stride):
        k = CoordinateMapKey(tensor_stride)
        if k not in self._maps:
            raise RuntimeError(f"no coordinate map with tensor stride {list(tensor_stride)} in this manager")
        return k

    def get_coordinates(self, key):
        return self._maps[key].coords

    def size(self, key):
        return self._maps[key].n

    # -- kernel maps ------------------------------------------------------------------------------------
    def _table(self, in_key, out_key, ks, dil):
        lib = _lib.load()
        cin, cout = self._maps[in_key], self._maps[out_key]
        K = ks ** 3
        table = torch.empty((K, cout.n), dtype=torch.int32, device=cout.coords.device)
        counts = torch.empty(K, dtype=torch.int32, device=cout.coords.device)
        _lib.check(lib.lgs_kmap_build(_lib.ptr(cout.coords), cout.n, _lib.ptr(cin.tkeys), _lib.ptr(cin.tvals),
                                      cin.capacity, ks, _uniform(list(in_key.tensor_stride), "tensor stride"), dil,
                                      _lib.ptr(table), _lib.ptr(counts), _stream()))
        return table, counts

    def _plan(self, km, cmap, step):
        """neighbourhood plan of a same-map 3^3 kernel map: supertiles of spatially close output rows, their unique input
        rows and the table in local indices (csrc/nbplan.cu), consumed by lgs_conv_fwd3.  No host sync: see KernelMap.plan."""
        lib = _lib.load()
        if not _state.get("nbplan", True) or not lib.lgs_nbplan_supported(km.n_out, km.K):
            return
        dev = km.fwd_table.device
        plan = torch.empty(lib.lgs_nbplan_bytes(km.n_out, km.K) // 4, dtype=torch.int32, device=dev)
        scratch = torch.empty(lib.lgs_nbplan_scratch_bytes(km.n_out) // 4, dtype=torch.int32, device=dev)
        _lib.check(lib.lgs_nbplan_build(_lib.ptr(cmap.coords), km.n_out, _lib.ptr(km.fwd_table), km.K, step, _lib.ptr(plan),
                                        _lib.ptr(scratch), None, _stream()))
        km._plan_status = _plan_status_slot()
        km._plan_status.copy_(plan[8:10], non_blocking=True)          # header words 8, 9: overflow flag, largest unique-row count
        km._plan_event = torch.cuda.Event()
        km._plan_event.record(torch.cuda.current_stream())
        km._plan = plan

    def _transpose(self, table, n_in):
        lib = _lib.load()
        K, n_out = table.shape
        tt = torch.empty((K, n_in), dtype=torch.int32, device=table.device)
        _lib.check(lib.lgs_kmap_transpose(_lib.ptr(table), K, n_out, n_in, _lib.ptr(tt), _stream()))
        return tt

    def kernel_map(self, in_key, out_key, kernel_size, dilation, is_transpose=False) -> KernelMap:
        ks, dil = _uniform(list(kernel_size), "kernel size"), _uniform(list(dilation), "dilation")
        ck = (in_key, out_key, ks, dil, bool(is_transpose))
        km = self._kmaps.get(ck)
        if km is not None:
            return km
        req = (in_key.tensor_stride, out_key.tensor_stride, ks, dil, bool(is_transpose))
        if req not in CoordinateManager._learned_kmaps:
            CoordinateManager._learned_kmaps.append(req)
        km = KernelMap()
        km.K = ks ** 3
        if is_transpose:
            # map of the forward conv fine(out_key) -> coarse(in_key) with (in, out) swapped (App. A.7)
            down = self.kernel_map(out_key, in_key, kernel_size, dilation, False)
            km.n_in, km.n_out = down.n_out, down.n_in
            km.fwd_table, km.bwd_table, km.bwd_reverse, km.counts = down.bwd_table, down.fwd_table, False, down.counts
        else:
            km.n_in, km.n_out = self._maps[in_key].n, self._maps[out_key].n
            km.fwd_table, km.counts = self._table(in_key, out_key, ks, dil)
            if in_key == out_key and ks % 2 == 1:
                # C_in[i] = C[o] + off_k  <=>  C[o] = C_in[i] + off_{K-1-k}: dgrad reads the same table mirrored
                km.bwd_table, km.bwd_reverse = km.fwd_table, True
                cmap, step = self._maps[out_key], _uniform(list(in_key.tensor_stride), "tensor stride") * dil
                if _state.get("plans_wanted"):
                    self._plan(km, cmap, step)
                else:
                    km._plan_builder = lambda km=km, cmap=cmap, step=step: self._plan(km, cmap, step)
            else:
                km.bwd_table, km.bwd_reverse = self._transpose(km.fwd_table, km.n_in), False
        self._kmaps[ck] = km
        return km

    def tensors(self):
        """every device buffer this manager owns (coordinate rows, cuckoo tables, neighbour tables)"""
        seen = set()
        for cm in self._maps.values():
            for t in (cm.coords, cm.tkeys, cm.tvals):
                if id(t) not in seen:
                    seen.add(id(t))
                    yield t
        for km in self._kmaps.values():
            for t in (km.fwd_table, km.bwd_table, km.counts, km._plan):
                if t is not None and id(t) not in seen:
                    seen.add(id(t))
                    yield t

    def conv_maps(self, in_key, ks, stride, dil, transpose):
        """(output key, kernel map) of one convolution call — one dict lookup on the hot path."""
        ck = (in_key, ks, stride, dil, transpose)
        r = self._conv_cache.get(ck)
        if r is None:
            if transpose:
                out_key = self.key_with_stride([t // stride for t in in_key.tensor_stride])
            else:
                out_key = self.stride(in_key, [stride] * self.D) if stride > 1 else in_key
            r = self._conv_cache[ck] = (out_key, self.kernel_map(in_key, out_key, [ks] * self.D, [dil] * self.D, transpose))
        return r

    def kernel_map_pairs(self, km: KernelMap):
        """ME-style pair lists [(in_idx, out_idx)] per offset, derived from the table (for inspection / tests)."""
        res = []
        for k in range(km.K):
            o = torch.nonzero(km.fwd_table[k] >= 0).squeeze(1)
            res.append((km.fwd_table[k][o].long(), o))
        return res


# ----------------------------------------------------------------------------------------------------------
# SparseTensor
# ----------------------------------------------------------------------------------------------------------
class _PendingConv:
    """A convolution whose launch is deferred until its output is needed: if the consumer is a fusable BatchNorm, conv,
    BatchNorm (+ residual) (+ ReLU) run inside ONE autograd node (_ConvBNActFn) — same kernels, half the autograd /
    Python dispatch work.  `.F` on the output materialises the plain convolution."""
    __slots__ = ("conv", "x", "km", "plain")

    def __init__(self, conv, x, km):
        self.conv, self.x, self.km, self.plain = conv, x, km, None

    def out_rows(self):
        return self.km.n_out if self.km is not None else self.x.shape[0]


class _PendingBN:
    """A BatchNorm whose application is deferred so that a following `+= residual` and ReLU fuse into one kernel pair
    (lgs_bn_fwd / lgs_bn_bwd).  The reference calls bn -> (+= residual) -> relu as separate modules
    (models/modules/resnet_block.py:41-57, models/res16unet.py:196-270); nothing in the model code changes.
    `conv` is the still-pending convolution that feeds it (then `x` is that convolution's INPUT features)."""
    __slots__ = ("bn", "x", "res", "plain", "consumed", "conv")

    def __init__(self, bn, x, conv=None):
        self.bn, self.x, self.res, self.plain, self.consumed, self.conv = bn, x, None, None, False, conv


class SparseTensor:
    def __init__(self, features, coordinates=None, coordinate_map_key=None, coordinate_manager=None,
                 tensor_stride=1, device=None, _pending=None, **_ignored):
        self._pending = _pending
        for k, v in _ignored.items():
            # ME arguments that change semantics must not vanish silently: duplicates are always resolved by keeping the
            # first row (RANDOM_SUBSAMPLE, ME's default and what the reference relies on)
            if k == "quantization_mode" and "RANDOM_SUBSAMPLE" not in str(v) and v is not None:
                raise NotImplementedError(f"SparseTensor(quantization_mode={v}): lgs_b200 implements RANDOM_SUBSAMPLE only")
            if k not in ("quantization_mode", "minkowski_algorithm", "allocator_type", "requires_grad"):
                import warnings
                warnings.warn(f"lgs_b200 SparseTensor ignores the argument {k!r}", stacklevel=2)
        if coordinate_map_key is None:
            if coordinates is None:
                raise ValueError("SparseTensor needs coordinates or a coordinate_map_key")
            dev = features.device if features.is_cuda else torch.device("cuda", torch.cuda.current_device())
            if device is not None:
                dev = torch.device(device)
            if dev.type != "cuda":
                raise RuntimeError("lgs_b200 has no CPU path: SparseTensor needs a CUDA device")
            c = torch.as_tensor(coordinates)
            if c.is_floating_point():
                c = torch.floor(c)
            c = c.to(device=dev, dtype=torch.int32).contiguous()
            features = features.to(dev)
            mgr = coordinate_manager or CoordinateManager(D=c.shape[1] - 1)
            key, uidx, inv = mgr.insert_and_map(c, _as_list(tensor_stride, c.shape[1] - 1))
            if uidx.shape[0] != c.shape[0]:  # duplicates: one row per coordinate, first occurrence kept
                features = features.index_select(0, uidx.long())
            self.unique_index, self.inverse_mapping = uidx, inv
            coordinate_map_key, coordinate_manager = key, mgr
        self._F = features
        self.coordinate_map_key = coordinate_map_key
        self.coordinate_manager = coordinate_manager

    @property
    def F(self):
        p = self._pending
        if p is not None:
            if p.plain is None:
                if type(p) is _PendingConv:      # materialise the deferred convolution
                    p.plain = sparse_conv(p.x, p.conv._parameters["kernel"], p.conv.bias, p.km, module=p.conv)
                else:                            # materialise the deferred BatchNorm (+ residual), no ReLU
                    p.plain = _bn_act(p, relu=False)
            self._F, self._pending = p.plain, None
        return self._F

    feats = F

    @property
    def C(self):
        return self.coordinate_manager.get_coordinates(self.coordinate_map_key)

    coordinates = C

    @property
    def tensor_stride(self):
        return list(self.coordinate_map_key.tensor_stride)

    @property
    def device(self):
        return (self._pending.x if self._pending is not None else self._F).device

    @property
    def D(self):
        return self.coordinate_manager.D

    @property
    def shape(self):
        return self.F.shape

    def __len__(self):
        return self.F.shape[0]

    def _same(self, o):
        if o.coordinate_map_key != self.coordinate_map_key or o.coordinate_manager is not self.coordinate_manager:
            raise RuntimeError("SparseTensor arithmetic needs identical coordinate_map_key")

    @classmethod
    def _make(cls, F, key, mgr, pending=None):
        """op-output constructor: no coordinate work, no argument parsing (called ~250 times per step)"""
        o = object.__new__(cls)
        o._F, o.coordinate_map_key, o.coordinate_manager, o._pending = F, key, mgr, pending
        return o

    def _like(self, F):
        return SparseTensor._make(F, self.coordinate_map_key, self.coordinate_manager)

    def __add__(self, o):
        self._same(o)
        return self._like(self.F + o.F)

    def __iadd__(self, o):  # models/modules/resnet_block.py:54  `out += residual`
        self._same(o)
        p = self._pending
        if type(p) is _PendingBN and p.res is None and p.plain is None:
            p.res = o.F              # fused into the deferred BatchNorm's epilogue
            return self
        self._F = self.F + o.F
        return self


def cat(*sts):
    for s in sts[1:]:
        sts[0]._same(s)
    return sts[0]._like(torch.cat([s.F for s in sts], 1))


# ----------------------------------------------------------------------------------------------------------
# fused BatchNorm (+ residual) (+ ReLU)
# ----------------------------------------------------------------------------------------------------------
def _bn_fwd_impl(x, res, gamma, beta, bn, relu, update_running):
    lib = _lib.load()
    if not x.is_contiguous():
        x = x.contiguous()
    n, c = x.shape
    if res is not None and not res.is_contiguous():
        res = res.contiguous()
    z = torch.empty_like(x)
    stats = torch.empty((2, c), dtype=torch.float32, device=x.device)       # save_mean, save_invstd
    bufs = bn._buffers                    # plain dict reads: nn.Module.__getattr__ costs ~1 us per parameter / buffer access
    rm = bufs["running_mean"] if update_running else None
    rv = bufs["running_var"] if update_running else None
    sc = _scratch64(x.device.index)
    acc, nxt = sc.pair()
    rc = lib.lgs_bn_fwd(_lib.ptr(x), _lib.ptr(res), n, c, _lib.ptr(gamma), _lib.ptr(beta), float(bn.eps),
                        float(bn.momentum), 1 if relu else 0, _lib.ptr(rm), _lib.ptr(rv), _lib.ptr(z),
                        ctypes.c_void_p(stats.data_ptr()), ctypes.c_void_p(stats.data_ptr() + 4 * c), acc, nxt,
                        _lib.ptr(bufs["num_batches_tracked"]) if update_running else None, _stream())
    sc.done(rc == _lib.OK)
    _lib.check(rc)
    return x, z, stats


def _bn_bwd_impl(x, z, gamma, stats, dz, relu, need_dres):
    lib = _lib.load()
    if not dz.is_contiguous():
        dz = dz.contiguous()
    n, c = x.shape
    dx = torch.empty_like(x)
    dres = torch.empty_like(x) if need_dres else None
    dgb = torch.empty((2, c), dtype=torch.float32, device=x.device)
    sc = _scratch64(x.device.index)
    acc, nxt = sc.pair()
    sp, gp = stats.data_ptr(), dgb.data_ptr()
    rc = lib.lgs_bn_bwd(_lib.ptr(x), _lib.ptr(z), _lib.ptr(dz), n, c, _lib.ptr(gamma), ctypes.c_void_p(sp),
                        ctypes.c_void_p(sp + 4 * c), 1 if relu else 0, _lib.ptr(dx), _lib.ptr(dres), ctypes.c_void_p(gp),
                        ctypes.c_void_p(gp + 4 * c), acc, nxt, _stream())
    sc.done(rc == _lib.OK)
    _lib.check(rc)
    return dx, dres, dgb[0], dgb[1]


class _BNActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, res, gamma, beta, bn, relu, update_running):
        x, z, stats = _bn_fwd_impl(x, res, gamma, beta, bn, relu, update_running)
        ctx.save_for_backward(x, z if relu else None, gamma, stats)
        ctx.relu, ctx.has_res = relu, res is not None
        return z

    @staticmethod
    def backward(ctx, dz):
        x, z, gamma, stats = ctx.saved_tensors
        dx, dres, dg, db = _bn_bwd_impl(x, z, gamma, stats, dz, ctx.relu, ctx.has_res and ctx.needs_input_grad[1])
        return dx, dres, dg, db, None, None, None


def _bn_act(p: _PendingBN, relu: bool):
    if p.conv is not None:
        c = p.conv
        bp = p.bn._parameters
        out = _ConvBNActFn.apply(c.x, c.conv._parameters["kernel"], bp["weight"], bp["bias"], p.res, c.km, _state["algo"],
                                 p.bn, relu, not p.consumed, c.conv)
    else:
        bp = p.bn._parameters
        out = _BNActFn.apply(p.x, p.res, bp["weight"], bp["bias"], p.bn, relu, not p.consumed)
    p.consumed = True          # running statistics are updated once per BatchNorm call
    return out


def _bn_fusable_meta(bn, dtype, n, c):
    return (bn.training and dtype is torch.float32 and _state["fuse_bn"] and type(bn) is nn.BatchNorm1d
            and bn.track_running_stats and bn.affine and bn.momentum is not None and n >= 1 and c % 4 == 0 and c <= 1024)


def _bn_fusable(bn, F):
    return F.is_cuda and F.dim() == 2 and _bn_fusable_meta(bn, F.dtype, F.shape[0], F.shape[1])


def set_bn_fusion(flag: bool):
    """Deferred BatchNorm -> (+ residual) -> ReLU fusion (default on); off = plain ATen ops per module."""
    _state["fuse_bn"] = bool(flag)


# ----------------------------------------------------------------------------------------------------------
# convolution
# ----------------------------------------------------------------------------------------------------------
class _WeightPrep:
    """Tensor-core operand forms (lgs_weight_prep) of every registered convolution's weights, refreshed by ONE
    lgs_weight_prep_batch launch instead of one launch per layer call.  Buffers are persistent per module.
    A cached operand is trusted only while (a) no convolution weight gradient has been computed since it was derived
    (every training step therefore re-derives all of them, once, at its first convolution), (b) no torch optimiser has
    stepped (global post-step hook) and (c) the parameter's version counter and storage are unchanged.  Version
    counters alone are not enough: torch's fused optimisers update parameters without bumping them.
    Parameters changed behind all three (e.g. `p.data.add_()` in inference code) need `invalidate_weight_cache()`."""

    def __init__(self):
        self.mods = weakref.WeakSet()
        self.desc = {}          # (device, dt, nsplit) -> descriptor cache, see refresh()
        self.epoch = 0          # bumped whenever the parameters may have changed
        self.dirty = False      # a weight gradient was computed / an optimiser stepped since the last refresh
        self.hooked = False

    def register(self, mod):
        self.hook_optimizers()
        self.mods.add(mod)

    def invalidate(self):
        self.dirty = True

    def hook_optimizers(self):
        if not self.hooked:
            self.hooked = True
            try:
                from torch.optim.optimizer import register_optimizer_step_post_hook
                register_optimizer_step_post_hook(lambda *a, **k: self.invalidate())
            except ImportError:      # older torch: the gradient rule (a) still covers training loops
                pass

    def lookup(self, mod, w3, dt, nsplit, tdtype):
        """(w_fwd, w_bwd) for `mod` (either may be None if the tensor-core kernels do not take that direction), or None
        when this module is not handled here."""
        if self.dirty:
            self.dirty = False
            self.epoch += 1
        st = mod._prep
        sig = (self.epoch, w3._version, w3.data_ptr(), dt, nsplit)
        if st is None or st[0] != sig:
            self.refresh(w3.device, dt, nsplit, tdtype)
            st = mod._prep
            if st is None or st[0] != sig:
                return None
        return st[1], st[2]

    def refresh(self, device, dt, nsplit, tdtype):
        lib = _lib.load()
        ck = (device, dt, nsplit)
        cached = self.desc.get(ck)
        if cached is not None and cached[4] == len(self.mods):
            # fast path (every training step): same set of layers, weights still at the same addresses -> reuse the
            # device descriptor table, relaunch, stamp every layer's operands with the new epoch / version
            live = []
            for ref, wptr, fb, bb in cached[5]:
                m = ref()
                if m is None:
                    break
                w = m._parameters["kernel"]
                if w.data_ptr() != wptr or w.dtype is not torch.float32:
                    break
                live.append((m, w, wptr, fb, bb))
            else:
                _lib.check(lib.lgs_weight_prep_batch(_lib.ptr(cached[1]), cached[2], cached[3], nsplit, dt, _stream()))
                ep = self.epoch
                for m, w, wptr, fb, bb in live:
                    m._prep = ((ep, w._version, wptr, dt, nsplit), fb, bb)
                return
        es = 2 if dt == _lib.BF16 else 4
        entries = []
        for m in list(self.mods):
            w = m._parameters["kernel"]
            if w.device != device or w.dtype is not torch.float32 or not w.is_contiguous():
                continue
            K, c_in, c_out = (1,) + tuple(w.shape) if w.dim() == 2 else tuple(w.shape)
            if (c_in * es) % 16 != 0:
                continue                                   # padded per call (conv0p1s1)
            f_ok, b_ok = _tc_supported(lib, c_in, c_out, dt), _tc_supported(lib, c_out, c_in, dt)
            if not (f_ok or b_ok):
                continue
            bufs = m._prep_bufs.get((dt, nsplit))
            if bufs is None or bufs[2] != w.data_ptr():
                fb = _operand_buffer(lib, nsplit, K, c_out, c_in, tdtype, device) if f_ok else None
                bb = _operand_buffer(lib, nsplit, K, c_in, c_out, tdtype, device) if b_ok else None
                bufs = m._prep_bufs[(dt, nsplit)] = (fb, bb, w.data_ptr())
            entries.append((m, w, K, c_in, c_out, bufs[0], bufs[1]))
        if not entries:
            self.desc.pop(ck, None)
            return
        # shapes belong in the key: a new layer's tensors can land on a dead layer's addresses
        key = tuple((e[1].data_ptr(), e[5].data_ptr() if e[5] is not None else 0,
                     e[6].data_ptr() if e[6] is not None else 0, e[2], e[3], e[4]) for e in entries)
        if cached is None or cached[0] != key:
            rows, tile0 = [], 0
            for (_, w, K, c_in, c_out, fb, bb), k3 in zip(entries, key):
                rows.append([k3[0], k3[1], k3[2], K, c_in, c_out, tile0, 0])
                tile0 += K * ((c_in + 31) // 32) * ((c_out + 31) // 32)
            table = torch.tensor(rows, dtype=torch.int64).to(device)
            cached = (key, table, len(rows), tile0)
        # (key, descriptor table, layers, tiles, registered modules at build time, [(weakref(layer), weight address, operands)])
        cached = self.desc[ck] = cached[:4] + (len(self.mods), [(weakref.ref(e[0]), e[1].data_ptr(), e[5], e[6])
                                                                for e in entries])
        _lib.check(lib.lgs_weight_prep_batch(_lib.ptr(cached[1]), cached[2], cached[3], nsplit, dt, _stream()))
        for m, w, _, _, _, fb, bb in entries:
            m._prep = ((self.epoch, w._version, w.data_ptr(), dt, nsplit), fb, bb)


_weight_prep = _WeightPrep()


def invalidate_weight_cache():
    """Force the next convolution call to re-derive the tensor-core weight operands (needed only after in-place
    parameter updates that bypass the version counter, e.g. through `.data`)."""
    _weight_prep.invalidate()


class _ConvMeta:
    __slots__ = ("km", "algo", "bwd_tc", "tc_layout", "dims", "w_shape", "w_dtype", "has_bias", "c_in_true")


def _operand_buffer(lib, nsplit, K, rows, red, dtype, device):
    """tensor-core weight operand of one direction: [nsplit, K, rows, red] in the feature dtype, or the LGS_W_BX3 form"""
    if nsplit == 3:
        return torch.empty(lib.lgs_weight_bx3_elems(K, rows, red), dtype=torch.bfloat16, device=device)
    return torch.empty((nsplit, K, rows, red), dtype=dtype, device=device)


def _conv_fwd_impl(feats, weight, bias, km, algo, need_dgrad, module=None):
    """one convolution launch (+ weight operand prep); returns (out, feats as saved for wgrad, dgrad weights, meta)"""
    lib = _lib.load()
    if not feats.is_contiguous():
        feats = feats.contiguous()
    w3 = weight.view(1, *weight.shape) if weight.dim() == 2 else weight
    K, c_in, c_out = w3.shape
    dt = _lib.F32 if feats.dtype is torch.float32 else _dtype_code(feats)
    if (algo == _lib.ALGO_TC3 or algo == _lib.ALGO_BX3) and dt == _lib.BF16:
        algo = _lib.ALGO_TC                                   # bf16 features: plain bf16 tensor-core products
    n_in = feats.shape[0]
    n_out = km.n_out if km is not None else n_in
    out = torch.empty((n_out, c_out), dtype=feats.dtype, device=feats.device)
    b32 = bias.detach().float().contiguous().view(-1) if bias is not None else None
    w32 = None

    def weights32():
        w = w3.detach()
        if w.dtype is not torch.float32 or not w.is_contiguous():
            w = w.float().contiguous()
        if padded:
            w = torch.nn.functional.pad(w, (0, 0, 0, c_in - c_in_true))
        return w
    # A tiny channel count (the 3 colour channels of conv0p1s1) is zero-padded to a 16-byte row so that the layer
    # takes the tensor-core kernels; the padded weight rows are zero and the padded gradients are dropped.
    c_in_true = c_in
    row_bytes = c_in * feats.element_size()
    padded = algo != _lib.ALGO_SIMT and c_in < 16 and row_bytes % 16 != 0
    if padded:
        c_in = -(-row_bytes // 16) * 16 // feats.element_size()
        feats = torch.nn.functional.pad(feats, (0, c_in - c_in_true))
    # tensor-core operand forms of the weights; a direction the TC kernels do not take (e.g. c_in = 3) runs on the
    # exact SIMT kernel with the parameter itself
    fwd_tc = algo != _lib.ALGO_SIMT and _tc_supported(lib, c_in, c_out, dt)
    bwd_tc = algo != _lib.ALGO_SIMT and need_dgrad and _tc_supported(lib, c_out, c_in, dt)
    nsplit = 3 if algo == _lib.ALGO_BX3 else (2 if algo == _lib.ALGO_TC3 else 1)
    w_fwd = w_bwd = None
    if fwd_tc or bwd_tc:
        pre = None
        if (module is not None and not padded and _state["batch_prep"] and weight.dtype is torch.float32
                and weight.is_contiguous()):
            pre = _weight_prep.lookup(module, w3, dt, nsplit, feats.dtype)     # one launch for all layers, per update
        if pre is not None and (pre[0] is not None or not fwd_tc) and (pre[1] is not None or not bwd_tc):
            w_fwd, w_bwd = pre[0], (pre[1] if bwd_tc else None)
        else:
            if fwd_tc:
                w_fwd = _operand_buffer(lib, nsplit, K, c_out, c_in, feats.dtype, feats.device)
            if bwd_tc:
                w_bwd = _operand_buffer(lib, nsplit, K, c_in, c_out, feats.dtype, feats.device)
            w32 = weights32()
            _lib.check(lib.lgs_weight_prep(_lib.ptr(w32), K, c_in, c_out, nsplit, _lib.ptr(w_fwd), _lib.ptr(w_bwd), dt,
                                           _stream()))
    tc_layout = _lib.W_BX3 if nsplit == 3 else (_lib.W_KNC_SPLIT if nsplit == 2 else _lib.W_KNC)
    if fwd_tc:
        wf, layout, a = w_fwd, tc_layout, algo
    else:
        w32 = weights32() if w32 is None else w32
        wf, layout, a = w32.to(feats.dtype), _lib.W_KCN, _lib.ALGO_SIMT
    with _Timed("fwd", K, c_in, c_out, n_in, n_out, km, feats.dtype):
        if a == _lib.ALGO_BX3 and dt == _lib.F32 and km is not None and km.plan is not None:
            # large same-map 3^3 map: neighbourhood-cache kernel (same products and per-row accumulation order)
            _lib.check(lib.lgs_conv_fwd3(_lib.ptr(feats), c_in, None, 0, n_in, _lib.ptr(wf), K, c_out, _lib.ptr(km.fwd_table),
                                         _lib.ptr(km.plan), n_out, 0, _lib.ptr(b32), _lib.ptr(out), None, _stream()))
        else:
            _lib.check(lib.lgs_conv_fwd(_lib.ptr(feats), n_in, c_in, _lib.ptr(wf), layout, K, c_out,
                                        _lib.ptr(km.fwd_table) if km is not None else None, n_out, 0, _lib.ptr(b32),
                                        _lib.ptr(out), dt, a, _stream()))
    if need_dgrad and not bwd_tc:
        w_bwd = (weights32() if w32 is None else w32).to(feats.dtype)                           # [K, c_in, c_out] read as LGS_W_KNC by the SIMT dgrad
    m = _ConvMeta()
    m.km, m.algo, m.bwd_tc, m.tc_layout = km, algo, bwd_tc, tc_layout
    m.dims, m.w_shape, m.w_dtype, m.has_bias = (K, c_in, c_out), weight.shape, weight.dtype, bias is not None
    m.c_in_true = c_in_true
    return out, feats, w_bwd, m


def _conv_bwd_impl(m, feats, w_bwd, gout, need_gin, need_gw, need_gb):
    """dgrad (the forward kernel on the transposed problem) + wgrad (+ bias gradient) of one convolution"""
    lib = _lib.load()
    km, algo = m.km, m.algo
    if not gout.is_contiguous():
        gout = gout.contiguous()
    K, c_in, c_out = m.dims
    n_in, n_out = feats.shape[0], gout.shape[0]
    dt = _lib.F32 if feats.dtype is torch.float32 else _dtype_code(feats)
    gin = gw = gb = None
    table = _lib.ptr(km.fwd_table) if km is not None else None
    joined = None
    if need_gw:
        _weight_prep.dirty = True       # a weight gradient exists: cached tensor-core weight operands expire
        gw = torch.empty((K, c_in, c_out), dtype=torch.float32, device=feats.device)
        if need_gin and _state["overlap_wgrad"] and _state["profile"] is None:
            # fork: wgrad on the side stream (it reads feats / gout, ready on this stream now), dgrad on this stream;
            # joined below, before anything can consume gw — every buffer stays owned by the training stream
            side, ev_fork, ev_join = _side_stream(feats.device.index)
            main = _cur_stream_obj(feats.device.index)
            ev_fork.record(main)
            side.wait_event(ev_fork)
            _lib.check(lib.lgs_conv_wgrad(_lib.ptr(feats), n_in, c_in, _lib.ptr(gout), n_out, c_out, table, K,
                                          _lib.ptr(gw), dt, algo, ctypes.c_void_p(side.cuda_stream)))
            ev_join.record(side)
            joined = main
    if need_gin:
        # dgrad = the same kernel on the transposed problem; W[k] ([c_in,c_out]) is its K-major B operand as is
        gin = torch.empty((n_in, c_in), dtype=feats.dtype, device=feats.device)
        layout, a = (m.tc_layout, algo) if m.bwd_tc else (_lib.W_KNC, _lib.ALGO_SIMT)
        with _Timed("dgrad", K, c_out, c_in, n_out, n_in, km, feats.dtype):
            if a == _lib.ALGO_BX3 and dt == _lib.F32 and km is not None and km.bwd_reverse and km.plan is not None:
                _lib.check(lib.lgs_conv_fwd3(_lib.ptr(gout), c_out, None, 0, n_out, _lib.ptr(w_bwd), K, c_in, _lib.ptr(km.bwd_table),
                                             _lib.ptr(km.plan), n_in, 1, None, _lib.ptr(gin), None, _stream()))
            else:
                _lib.check(lib.lgs_conv_fwd(_lib.ptr(gout), n_out, c_out, _lib.ptr(w_bwd), layout, K, c_in,
                                            _lib.ptr(km.bwd_table) if km is not None else None, n_in,
                                            1 if (km is not None and km.bwd_reverse) else 0, None, _lib.ptr(gin), dt,
                                            a, _stream()))
    if need_gw:
        if joined is not None:
            joined.wait_event(ev_join)                        # training stream waits for the side-stream wgrad
        else:
            with _Timed("wgrad", K, c_in, c_out, n_in, n_out, km, feats.dtype):
                _lib.check(lib.lgs_conv_wgrad(_lib.ptr(feats), n_in, c_in, _lib.ptr(gout), n_out, c_out, table, K,
                                              _lib.ptr(gw), dt, algo, _stream()))
        if m.c_in_true != c_in:
            gw = gw[:, :m.c_in_true, :].contiguous()
        gw = gw.view(m.w_shape).to(m.w_dtype)
    if need_gb:
        gb = gout.float().sum(0, keepdim=True)
    if gin is not None and m.c_in_true != c_in:
        gin = gin[:, :m.c_in_true].contiguous()
    return gin, gw, gb


class _SparseConvFn(torch.autograd.Function):
    """out = conv(feats, W) through the C ABI; backward = dgrad (same kernel, W^T, mirrored/transposed table) + wgrad."""

    @staticmethod
    def forward(ctx, feats, weight, bias, km, algo, module):
        out, feats, w_bwd, ctx.meta = _conv_fwd_impl(feats, weight, bias, km, algo, ctx.needs_input_grad[0], module)
        ctx.save_for_backward(feats, w_bwd)
        return out

    @staticmethod
    def backward(ctx, gout):
        feats, w_bwd = ctx.saved_tensors
        m = ctx.meta
        gin, gw, gb = _conv_bwd_impl(m, feats, w_bwd, gout, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                     m.has_bias and ctx.needs_input_grad[2])
        return gin, gw, gb, None, None, None


class _ConvBNActFn(torch.autograd.Function):
    """z = relu?( BatchNorm(conv(feats, W)) [+ residual] ) as ONE autograd node: lgs_conv_fwd + lgs_bn_fwd forward,
    lgs_bn_bwd + dgrad + wgrad backward.  (models/modules/resnet_block.py:41-57: conv -> norm -> (+=) -> relu)"""

    @staticmethod
    def forward(ctx, feats, weight, gamma, beta, res, km, algo, bn, relu, update_running, module):
        y, feats, w_bwd, ctx.meta = _conv_fwd_impl(feats, weight, None, km, algo, ctx.needs_input_grad[0], module)
        y, z, stats = _bn_fwd_impl(y, res, gamma, beta, bn, relu, update_running)
        ctx.save_for_backward(feats, w_bwd, y, z if relu else None, gamma, stats)
        ctx.relu, ctx.has_res = relu, res is not None
        return z

    @staticmethod
    def backward(ctx, dz):
        feats, w_bwd, y, z, gamma, stats = ctx.saved_tensors
        dy, dres, dg, db = _bn_bwd_impl(y, z, gamma, stats, dz, ctx.relu, ctx.has_res and ctx.needs_input_grad[4])
        gin, gw, _ = _conv_bwd_impl(ctx.meta, feats, w_bwd, dy, ctx.needs_input_grad[0], ctx.needs_input_grad[1], False)
        return gin, gw, dg, db, dres, None, None, None, None, None, None


def sparse_conv(feats, weight, bias, km, algo=None, module=None):
    return _SparseConvFn.apply(feats, weight, bias, km, _state["algo"] if algo is None else algo, module)


class MinkowskiNetwork(nn.Module):
    def __init__(self, D):
        super().__init__()
        self.D = D


class _ConvBase(nn.Module):
    TRANSPOSE = False

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, dimension=None):
        super().__init__()
        if dimension is None or dimension <= 0:
            raise ValueError("dimension must be given")
        if expand_coordinates:
            raise NotImplementedError("expand_coordinates is outside the hot path")
        if kernel_generator is None:
            kernel_generator = KernelGenerator(kernel_size, stride, dilation, dimension=dimension)
        self.kernel_generator = kernel_generator
        self.in_channels, self.out_channels, self.dimension = in_channels, out_channels, dimension
        self.use_mm = (not self.TRANSPOSE) and kernel_generator.kernel_volume == 1 and \
            all(s == 1 for s in kernel_generator.kernel_stride)
        K = kernel_generator.kernel_volume
        # state-dict ABI of ME 0.5.4: kernel [K,Cin,Cout] ([Cin,Cout] for 1x1), bias [1,Cout]
        self.kernel = nn.Parameter(torch.empty((in_channels, out_channels) if self.use_mm
                                               else (K, in_channels, out_channels)))
        self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
        # scalars of the (isotropic) kernel, computed once: forward() runs ~125 times per step
        self._ks = _uniform(list(kernel_generator.kernel_size), "kernel size")
        self._dil = _uniform(list(kernel_generator.kernel_dilation), "dilation")
        self._stride = _uniform(list(kernel_generator.kernel_stride), "stride")
        self._prep, self._prep_bufs = None, {}      # cached tensor-core weight operands (_WeightPrep)
        _weight_prep.register(self)
        self.reset_parameters()

    def __getstate__(self):
        # derived operand caches (up to 4x the weight bytes) are not state: keep them out of pickles / deep copies
        d = self.__dict__.copy()
        d["_prep"], d["_prep_bufs"] = None, {}
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        _weight_prep.register(self)

    def reset_parameters(self):
        with torch.no_grad():
            K = self.kernel_generator.kernel_volume
            n = (self.out_channels if self.TRANSPOSE else self.in_channels) * K
            stdv = 1.0 / np.sqrt(n)
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def forward(self, x: SparseTensor):
        mgr = x.coordinate_manager
        in_key = x.coordinate_map_key
        if self.use_mm:
            out_key, km = in_key, None
        else:
            out_key, km = mgr.conv_maps(in_key, self._ks, self._stride, self._dil, self.TRANSPOSE)
        F = x.F
        if self.bias is None and _state["fuse_conv_bn"] and F.dtype is torch.float32:
            # deferred: a following fusable BatchNorm turns conv + BN (+ residual) (+ ReLU) into one autograd node
            return SparseTensor._make(None, out_key, mgr, _PendingConv(self, F, km))
        return SparseTensor._make(sparse_conv(F, self._parameters["kernel"], self.bias, km, module=self), out_key, mgr)

    def extra_repr(self):
        kg = self.kernel_generator
        return (f"in={self.in_channels}, out={self.out_channels}, kernel_size={kg.kernel_size}, "
                f"stride={kg.kernel_stride}, dilation={kg.kernel_dilation}")


class MinkowskiConvolution(_ConvBase):
    TRANSPOSE = False


class MinkowskiConvolutionTranspose(_ConvBase):
    TRANSPOSE = True


# ----------------------------------------------------------------------------------------------------------
# row-wise layers (ATen on .F; fusion into the conv epilogue is SURVEY.md §8f rank 1)
# ----------------------------------------------------------------------------------------------------------
class MinkowskiBatchNorm(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)

    def forward(self, x: SparseTensor):
        p = x._pending
        bn = self._modules["bn"]
        if (type(p) is _PendingConv and p.plain is None
                and _bn_fusable_meta(bn, p.x.dtype, p.out_rows(), p.conv.out_channels)):
            return SparseTensor._make(None, x.coordinate_map_key, x.coordinate_manager, _PendingBN(bn, p.x, p))
        F = x.F
        if _bn_fusable(bn, F):
            return SparseTensor._make(None, x.coordinate_map_key, x.coordinate_manager, _PendingBN(bn, F))
        return x._like(bn(F))


class MinkowskiSyncBatchNorm(MinkowskiBatchNorm):
    @classmethod
    def convert_sync_batchnorm(cls, module, process_group=None):  # main.py:122-123
        for _, child in list(module.named_children()):
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
        p = x._pending
        if type(p) is _PendingBN and p.plain is None:
            return x._like(_bn_act(p, relu=True))      # BatchNorm (+ residual) + ReLU in one kernel pair
        return x._like(torch.relu(x.F))


def _stub(name):
    class _S(nn.Module):
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name}: outside the hot path (SURVEY.md §8b, import-surface row)")
    _S.__name__ = _S.__qualname__ = name
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
# utils
# ----------------------------------------------------------------------------------------------------------
def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False,
                    return_inverse=False, quantization_size=None, device=None, **_):
    """First-occurrence voxel de-duplication on the GPU (lib/voxelizer.py:142).  numpy in -> numpy out."""
    if labels is not None:
        raise NotImplementedError("label-voting mode of sparse_quantize is not used by the reference path")
    is_np = isinstance(coordinates, np.ndarray)
    c = torch.as_tensor(coordinates)
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    c = c.to(dev)
    if quantization_size is not None:
        c = c / quantization_size
    if c.is_floating_point():
        c = torch.floor(c)
    c4 = torch.cat([torch.zeros((c.shape[0], 1), dtype=torch.int32, device=dev), c.to(torch.int32)], 1).contiguous()
    cm, uidx, inv = _build_coordmap(c4, 1, True)
    conv = (lambda t: t.cpu().numpy()) if is_np else (lambda t: t.to(torch.as_tensor(coordinates).device))
    ret = [conv(cm.coords[:, 1:])]
    if features is not None:
        ret.append(features[conv(uidx.long())])
    if return_index:
        ret.append(conv(uidx.long()))
    if return_inverse:
        ret.append(conv(inv.long()))
    return ret[0] if len(ret) == 1 else tuple(ret)


def batched_coordinates(coords_list, dtype=torch.int32, device=None):
    out = []
    for b, c in enumerate(coords_list):
        c = torch.as_tensor(c)
        if c.is_floating_point():
            c = torch.floor(c)
        out.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=dtype), c.to(dtype)], 1))
    r = torch.cat(out, 0)
    return r.to(device) if device is not None else r


def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
    """(int32 [sum N, 1+D] with the batch column prepended, cat(feats), cat(labels))  (lib/transforms.py:421)."""
    bc = batched_coordinates(coords, dtype, device)
    f = torch.cat([torch.as_tensor(x) for x in feats], 0)
    if labels is None:
        return bc, f
    return bc, f, torch.cat([torch.as_tensor(x) for x in labels], 0)


utils = types.ModuleType(__name__ + ".utils")
utils.sparse_quantize = sparse_quantize
utils.sparse_collate = sparse_collate
utils.batched_coordinates = batched_coordinates

MinkowskiOps = types.ModuleType(__name__ + ".MinkowskiOps")
MinkowskiOps.cat = cat

__version__ = "0.5.4+lgs_b200"


def install(name="MinkowskiEngine"):
    """Register this facade under the package name the reference imports."""
    me = sys.modules[__name__]
    sys.modules[name] = me
    sys.modules[name + ".MinkowskiOps"] = MinkowskiOps
    sys.modules[name + ".utils"] = utils
    return me
