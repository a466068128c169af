"""Native step driver: forward + loss + backward of a Res16UNet as ONE C-ABI call (include/lgs_b200.h, lgs_program_*).

    step = NativeStep(net)                       # net: nets.Res16UNet OR the reference's own models.Res16UNet34C instance
    loss = step.run(sparse_input, labels)        # == criterion(net(x)[0].F, labels); loss.backward()   (gradients in p.grad)
    optimizer.step()

The network object keeps owning its parameters, BatchNorm buffers and gradients (state-dict, optimiser, checkpoints are
untouched); what changes is who walks the layers.  The facade walks them in Python, one module call and one autograd node
at a time (~14.8 ms of host time per Res16UNet34C step) — here `NativeStep.__init__` walks them ONCE and
writes down a straight-line program of engine calls (the order is the one `step.StepProgram` executes, which
tests/test_step_program.py proves equal to autograd's), and every step is then a single `lgs_program_run` that issues
those calls from C++ with buffer addresses bump-allocated in one arena.

Reference call sites replaced (relative to /root/reference): the body of `training_step` / `model_step`
lib/train_test/pl_BaselineTrainer.py:157-160, 288-309 (forward through models/res16unet.py:196-270, criterion :343-350,
`loss.backward()` under Lightning).

Differences from the facade's call sequence, all inside this file: BatchNorm column statistics come out of the
convolution's epilogue (lgs_conv_fwd2 d_bn_sums + lgs_bn_fwd2 stats_ready) instead of a separate pass; `cat`, the column
split of its gradient, gradient sums and the classifier's bias gradient are lgs_copy2d / lgs_add / lgs_colsum instead of
ATen kernels; parameter gradients are WRITTEN (not accumulated) into `p.grad`, which are views of one flat buffer
(`ddp.GradAllReducer`), so the NCCL all-reduce needs no packing and its buckets can be launched between two ranges of the
program while backward is still running.
"""
import ctypes
import struct

import numpy as np
import os

import torch

from . import _lib
from . import minkowski as E
from .ddp import GradAllReducer

OP_WORDS = 18
OP_WEIGHT_PREP, OP_CONV, OP_WGRAD, OP_BN_FWD, OP_BN_BWD, OP_COPY2D, OP_ADD, OP_SEG_CE, OP_COLSUM, OP_JOIN = range(1, 11)


def _f32_bits(x):
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


class _Buf:
    """symbolic buffer: external slot or arena intermediate of rows(level) x channels"""
    __slots__ = ("id", "level", "c")

    def __init__(self, id, level, c):
        self.id, self.level, self.c = id, level, c


class NativeStep:
    LEVELS = 5

    def __init__(self, net, ignore_index=-1, reducer=None, fuse_bn_stats=True, keep_logits=False, overlap_wgrad=True, head=None,
                 _dry=False):
        """head=None: segmentation step (classifier `final` + mean cross-entropy with ignore_index, BASELINE configs 1/2/4).
        head=callable(features [n0, C] requiring grad, labels) -> scalar loss: the U-Net body runs natively, the head under
        autograd (the CLIP pre-training nets of configs 3 / 5: anchor projection + fused text-anchor loss; it is tiny), and its
        gradient w.r.t. the features feeds the native backward."""
        convs = [m for m in net.modules() if isinstance(m, (E.MinkowskiConvolution, E.MinkowskiConvolutionTranspose))]
        if not convs or (head is None and getattr(net, "final", None) is None):
            raise NotImplementedError("NativeStep drives Res16UNet topologies built on the lgs_b200 facade (segmentation flavour, or any "
                                      "flavour with a `head`)")
        self.head = head
        p0 = next(net.parameters())
        if p0.dtype is not torch.float32 or (not p0.is_cuda and not _dry):
            # _dry: host-logic tests run the program with the library's call recorder on (lgs_trace_begin), nothing executes
            raise RuntimeError("NativeStep needs the network's fp32 parameters on a CUDA device (there is no CPU path)")
        self._dry = _dry
        self.net, self.device, self.ignore_index = net, p0.device, int(ignore_index)
        self.lib = _lib.load()
        self.fuse_stats, self.keep_logits, self.overlap_wgrad = bool(fuse_bn_stats), bool(keep_logits), bool(overlap_wgrad)
        self.reducer = reducer if reducer is not None else GradAllReducer(net.parameters(), overlap=False)
        self.reducer.attach()
        self._ops, self._bufs = [], []
        self._ext_static, self._ext_dynamic, self._keep, self._ext_cache = [], {}, [], {}    # [(slot, getter)], name -> slot, ...
        self._n_ext = 0
        # The four big level-0 decoder wgrads (block8: 1.4 ms of full-GPU work) are issued after the decoder's backward instead
        # of next to their dgrads: there they compete with the level-0 dgrads for every SM, later they run under the
        # latency-bound coarse encoder stages.  10.74 -> 10.63 ms on one box (profiles/r2_wgrad_deferral.txt; deferring level 1
        # as well, or spreading them over the encoder stages, is slower).  LGS_DEFER_WGRAD=0 switches it off.
        self._defer_wgrad = os.environ.get("LGS_DEFER_WGRAD", "1") != "0"
        self._deferred = []
        self._tables = {}             # (level_in, ks, stride, transpose, 'fwd'|'bwd') -> buffer
        self._plans = {}              # level -> buffer (neighbourhood plan of the level's 3^3 same-map kernel map)
        self._layers = []             # (conv module, weight source tensor, K, c_in_padded, c_out, fwd operand, bwd operand)
        self.marks = {}
        self._build()
        self._finish()

    # ---- buffers ---------------------------------------------------------------------------------------------------
    def _ext(self, getter=None, name=None, key=None):
        """external pointer slot.  getter: () -> tensor, re-evaluated whenever the module's tensors may have moved (static
        slot: parameters, gradients, buffers, operands);  name: filled per batch (dynamic slot: features, labels, tables)"""
        if key is not None and key in self._ext_cache:
            return self._ext_cache[key]
        slot = self._n_ext
        self._n_ext += 1
        self._bufs.append((0, slot, 0, 0))
        b = _Buf(len(self._bufs) - 1, None, None)
        if getter is not None:
            self._ext_static.append((slot, getter))
        else:
            self._ext_dynamic[name] = slot
        if key is not None:
            self._ext_cache[key] = b
        return b

    def _param(self, p):
        return self._ext(lambda: p, key=("p", id(p)))

    def _grad(self, p):
        return self._ext(lambda: p.grad, key=("g", id(p)))

    def _buffer(self, mod, name):
        return self._ext(lambda: mod._buffers[name], key=("b", id(mod), name))

    def _tensor(self, t):
        self._keep.append(t)
        return self._ext(lambda: t, key=("t", id(t)))

    def _new(self, level, c, elem=4):
        self._bufs.append((1, level, c, elem))
        return _Buf(len(self._bufs) - 1, level, c)

    def _op(self, code, *args):
        row = [code] + [a.id if isinstance(a, _Buf) else (-1 if a is None else int(a)) for a in args]
        assert len(row) <= OP_WORDS
        row = row + [0] * (OP_WORDS - len(row))
        if code == OP_CONV:
            if len(args) < 15:
                row[15] = -1            # no neighbourhood plan
            if len(args) < 16:
                row[16] = -1            # no addend
        self._ops.append(row)

    def _table(self, level_in, ks, stride, transpose, which):
        key = (level_in, ks, stride, transpose, which)
        if key not in self._tables:
            self._tables[key] = self._ext(name=("table",) + key)
        return self._tables[key]

    def _plan(self, conv, level):
        """external slot of the neighbourhood plan of a same-map 3^3 convolution's kernel map (NULL at run time when the
        manager built none: small map, or a supertile overflowed the cache) — word 15 of OP_CONV"""
        if conv.use_mm or conv._ks != 3 or conv._stride != 1 or conv.TRANSPOSE:
            return None
        if level not in self._plans:
            self._plans[level] = self._ext(name=("plan", level))
        return self._plans[level]

    # ---- layers ----------------------------------------------------------------------------------------------------
    def _conv_info(self, conv):
        """register a convolution: tensor-core operand buffers, padded weight if the input width is not a multiple of 4"""
        for rec in self._layers:
            if rec["mod"] is conv:
                return rec
        w = conv._parameters["kernel"]
        K = 1 if w.dim() == 2 else w.shape[0]
        c_in, c_out = w.shape[-2], w.shape[-1]
        c_pad = -(-c_in // 4) * 4
        if c_out % 4 != 0:
            raise NotImplementedError(f"NativeStep: {c_in}->{c_out} convolution (output channels must be a multiple of 4)")
        rec = {"mod": conv, "K": K, "c_in": c_in, "c_pad": c_pad, "c_out": c_out, "w": self._param(w), "gw": self._grad(w)}
        rec["src"] = lambda: conv._parameters["kernel"]
        if c_pad != c_in:
            wp = torch.zeros((K, c_pad, c_out), dtype=torch.float32, device=self.device)
            rec["w_pad"] = self._tensor(wp)
            rec["src"] = lambda: wp
        fwd = torch.empty(self.lib.lgs_weight_bx3_elems(K, c_out, c_pad), dtype=torch.bfloat16, device=self.device)
        rec["w_fwd_t"], rec["w_fwd"] = fwd, self._tensor(fwd)
        rec["w_bwd_t"] = rec["w_bwd"] = None
        self._layers.append(rec)
        return rec

    def _need_bwd_operand(self, rec):
        if rec["w_bwd"] is None:
            t = torch.empty(self.lib.lgs_weight_bx3_elems(rec["K"], rec["c_pad"], rec["c_out"]), dtype=torch.bfloat16, device=self.device)
            rec["w_bwd_t"], rec["w_bwd"] = t, self._tensor(t)

    def _geometry(self, conv, level):
        if conv.use_mm:
            return level, None, None, 0
        ks, stride, tr = conv._ks, conv._stride, conv.TRANSPOSE
        out_level = level - 1 if tr else (level + 1 if stride > 1 else level)
        fwd = self._table(level, ks, stride, tr, "fwd")
        if stride == 1 and ks % 2 == 1:
            return out_level, fwd, fwd, 1                 # same map, odd kernel: dgrad reads the table mirrored
        return out_level, fwd, self._table(level, ks, stride, tr, "bwd"), 0

    def _conv_fwd(self, conv, x, level, bias=None, stats=False):
        rec = self._conv_info(conv)
        out_level, t_fwd, _, _ = self._geometry(conv, level)
        assert x.c == rec["c_pad"], (x.c, rec["c_pad"])
        y = self._new(out_level, rec["c_out"])
        self._op(OP_CONV, x, x.c, None, 0, level, rec["w_fwd"], rec["K"], rec["c_out"], t_fwd, out_level, 0, bias, y, 1 if stats else 0,
                 self._plan(conv, level))
        return y, out_level

    def _conv_bwd(self, conv, x, level, dy, need_gin=True, addend=None):
        """wgrad (side stream next to dgrad, like the facade's _conv_bwd_impl) + dgrad; returns d x"""
        rec = self._conv_info(conv)
        out_level, t_fwd, t_bwd, rev = self._geometry(conv, level)
        side = 1 if (need_gin and self.overlap_wgrad) else 0
        if rec["c_pad"] != rec["c_in"]:
            gw = self._new(-1, rec["K"] * rec["c_pad"] * rec["c_out"])
            self._op(OP_WGRAD, x, rec["c_pad"], level, dy, rec["c_out"], out_level, t_fwd, rec["K"], gw, _lib.ALGO_BX3, 0)
            self._op(OP_COPY2D, gw, rec["c_pad"] * rec["c_out"], 0, rec["gw"], rec["c_in"] * rec["c_out"], 0, -1, rec["K"],
                     rec["c_in"] * rec["c_out"])
        else:
            args = (OP_WGRAD, x, rec["c_pad"], level, dy, rec["c_out"], out_level, t_fwd, rec["K"], rec["gw"], _lib.ALGO_BX3, side)
            if self._defer_wgrad and side and level == 0 and rec["K"] == 27 and rec["c_pad"] >= 64 and "decoder_done" not in self.marks:
                self._deferred.append(args)
            else:
                self._op(*args)
        if not need_gin:
            return None
        self._need_bwd_operand(rec)
        gin = self._new(level, rec["c_pad"])
        # addend: d x = dgrad + addend in the convolution's epilogue (lgs_conv_fwd4) instead of a separate add pass
        self._op(OP_CONV, dy, rec["c_out"], None, 0, out_level, rec["w_bwd"], rec["K"], rec["c_pad"], t_bwd, level, rev, None, gin, 0,
                 self._plan(conv, level), addend)
        return gin

    def _cbr_fwd(self, conv, bn, x, level, relu, res=None):
        """conv -> BatchNorm (+ residual) (+ ReLU): the facade's fused node (models/modules/resnet_block.py:41-57)"""
        fuse = self.fuse_stats
        y, out_level = self._conv_fwd(conv, x, level, stats=fuse)
        c = y.c
        bp, bb = bn._parameters, bn._buffers
        z, stats = self._new(out_level, c), self._new(-1, 2 * c)
        self._op(OP_BN_FWD, y, res, out_level, c, self._param(bp["weight"]), self._param(bp["bias"]), _f32_bits(bn.eps),
                 _f32_bits(bn.momentum), 1 if relu else 0, self._buffer(bn, "running_mean"), self._buffer(bn, "running_var"), z, stats,
                 self._buffer(bn, "num_batches_tracked"), 1 if fuse else 0)
        del bb
        return z, out_level, (conv, bn, x, level, y, z, stats, relu, res is not None)

    def _cbr_bwd(self, node, dz, need_gin=True, addend=None):
        conv, bn, x, level, y, z, stats, relu, has_res = node
        out_level = y.level
        c = y.c
        bp = bn._parameters
        dy = self._new(out_level, c)
        dres = self._new(out_level, c) if has_res else None
        self._op(OP_BN_BWD, y, z if relu else None, dz, out_level, c, self._param(bp["weight"]), stats, 1 if relu else 0, dy, dres,
                 self._grad(bp["weight"]), self._grad(bp["bias"]))
        return self._conv_bwd(conv, x, level, dy, need_gin, addend), dres

    def _add(self, a, b):
        out = self._new(a.level, a.c)
        self._op(OP_ADD, a, b, out, a.level, a.c)
        return out

    # ---- blocks (order of step.StepProgram) ---------------------------------------------------------------------
    @staticmethod
    def _block_parts(blk):
        ds = getattr(blk, "downsample", None)
        final_relu = getattr(blk, "final_relu", None)
        if final_relu is None:            # the reference's classes: BasicBlock ends with ReLU, NoReluBlock does not
            final_relu = type(blk).__name__ != "NoReluBlock"
        return blk.conv1, blk.norm1.bn, blk.conv2, blk.norm2.bn, ((ds[0], ds[1].bn) if ds is not None else None), final_relu

    def _block_fwd(self, blk, x, level):
        conv1, bn1, conv2, bn2, ds, final_relu = self._block_parts(blk)
        h, _, n1 = self._cbr_fwd(conv1, bn1, x, level, True)
        nd, r = None, x
        if ds is not None:
            r, _, nd = self._cbr_fwd(ds[0], ds[1], x, level, False)
        y, _, n2 = self._cbr_fwd(conv2, bn2, h, level, final_relu, res=r)
        return y, (n1, nd, n2)

    def _block_bwd(self, nodes, dy):
        n1, nd, n2 = nodes
        dh, dres = self._cbr_bwd(n2, dy)
        dxr = self._cbr_bwd(nd, dres)[0] if nd is not None else dres
        # d x = dgrad of the block's first convolution + the gradient of the identity / downsample path.  LGS_FUSE_ADDEND=1 adds
        # it in the convolution's epilogue (lgs_conv_fwd4): 23 fewer launches per step but 0.13 ms SLOWER — the epilogue's extra
        # row loads lengthen the dgrad kernels, while the separate add pass hides behind the wgrad side stream
        # (profiles/r2_addend_fusion.txt); off by default
        if os.environ.get("LGS_FUSE_ADDEND", "0") == "0":
            return self._add(self._cbr_bwd(n1, dh)[0], dxr)
        return self._cbr_bwd(n1, dh, addend=dxr)[0]

    def _stage_fwd(self, stage, x, level):
        nodes = []
        for blk in stage:
            x, n = self._block_fwd(blk, x, level)
            nodes.append(n)
        return x, nodes

    def _stage_bwd(self, nodes, d):
        for n in reversed(nodes):
            d = self._block_bwd(n, d)
        return d

    # ---- the whole step ------------------------------------------------------------------------------------------
    ENC = [("conv1p1s2", "bn1", "block1"), ("conv2p2s2", "bn2", "block2"), ("conv3p4s2", "bn3", "block3"), ("conv4p8s2", "bn4", "block4")]
    DEC = [("convtr4p16s2", "bntr4", "block5"), ("convtr5p8s2", "bntr5", "block6"), ("convtr6p4s2", "bntr6", "block7"),
           ("convtr7p2s2", "bntr7", "block8")]

    def _build(self):
        net = self.net
        c0 = self._conv_info(net.conv0p1s1)
        self.in_channels, self.in_pad = c0["c_in"], c0["c_pad"]
        # -- weights: padded copies, then every layer's tensor-core operands in one launch (descriptor table filled in _finish)
        for rec in [c0]:
            if "w_pad" in rec:
                self._op(OP_COPY2D, rec["w"], rec["c_in"] * rec["c_out"], 0, rec["w_pad"], rec["c_pad"] * rec["c_out"], 0, -1, rec["K"],
                         rec["c_in"] * rec["c_out"])
        self._desc = self._ext(name="desc")
        self._prep_op = len(self._ops)
        self._op(OP_WEIGHT_PREP, self._desc, 0, 0, 3)
        # -- forward
        x = self._ext(name="feats")
        x.level, x.c = 0, self.in_pad
        a, level, n0 = self._cbr_fwd(net.conv0p1s1, net.bn0.bn, x, 0, True)
        skips, enc, dec = [a], [], []
        for i, (cn, bn, bl) in enumerate(self.ENC):
            a, level, nd = self._cbr_fwd(getattr(net, cn), getattr(net, bn).bn, a, level, True)
            a, nb = self._stage_fwd(getattr(net, bl), a, level)
            enc.append((nd, nb))
            if i < 3:
                skips.append(a)
        for i, (cn, bn, bl) in enumerate(self.DEC):
            a, level, nt = self._cbr_fwd(getattr(net, cn), getattr(net, bn).bn, a, level, True)
            skip = skips[3 - i]
            cat = self._new(level, a.c + skip.c)
            self._op(OP_COPY2D, a, a.c, 0, cat, cat.c, 0, level, 0, a.c)
            self._op(OP_COPY2D, skip, skip.c, 0, cat, cat.c, a.c, level, 0, skip.c)
            width = a.c
            a, nb = self._stage_fwd(getattr(net, bl), cat, level)
            dec.append((nt, width, nb, cat.c))
        self._logits = None
        if self.head is None:
            # -- classifier + loss
            fin = net.final
            rf = self._conv_info(fin)
            bias = self._param(fin._parameters["bias"]) if fin._parameters.get("bias") is not None else None
            if self.keep_logits:
                self._logits = self._ext(name="logits")
                self._logits.level, self._logits.c = 0, rf["c_out"]
                self._op(OP_CONV, a, a.c, None, 0, 0, rf["w_fwd"], rf["K"], rf["c_out"], None, 0, 0, bias, self._logits, 0)
                logits = self._logits
            else:
                logits, _ = self._conv_fwd(fin, a, 0, bias=bias)
            self.n_classes = rf["c_out"]
            self._loss_t = torch.zeros((), dtype=torch.float32, device=self.device)
            ws = self._new(-1, 4, 8)
            dlog = self._new(0, rf["c_out"])
            self._op(OP_SEG_CE, logits, 0, rf["c_out"], self._ext(name="labels"), self.ignore_index, ws, self._tensor(self._loss_t), dlog)
            self.marks["forward_end"] = len(self._ops)
            # -- backward (reverse creation order, as autograd runs this graph)
            d = self._conv_bwd(fin, a, 0, dlog)
            if bias is not None:
                self._op(OP_COLSUM, dlog, 0, rf["c_out"], self._grad(fin._parameters["bias"]))
        else:
            # -- per-point features out to the head, their gradient back in
            self.feat_channels = a.c
            fo = self._ext(name="features")
            self._op(OP_COPY2D, a, a.c, 0, fo, a.c, 0, 0, 0, a.c)
            self.marks["forward_end"] = len(self._ops)
            d = self._ext(name="dfeatures")
            d.level, d.c = 0, a.c
        dskip = [None] * 4
        for i in (3, 2, 1, 0):
            nt, width, nb, cat_c = dec[i]
            d = self._stage_bwd(nb, d)
            lvl = d.level
            ds_, du = self._new(lvl, cat_c - width), self._new(lvl, width)
            self._op(OP_COPY2D, d, cat_c, width, ds_, cat_c - width, 0, lvl, 0, cat_c - width)
            self._op(OP_COPY2D, d, cat_c, 0, du, width, 0, lvl, 0, width)
            dskip[3 - i] = ds_
            d = self._cbr_bwd(nt, du)[0]
        self._op(OP_JOIN)
        self.marks["decoder_done"] = len(self._ops)          # every decoder / classifier gradient is complete here
        for args in self._deferred:                           # ... except the deferred level-0 wgrads, issued here
            self._op(*args)
        for i in (3, 2, 1, 0):
            nd, nb = enc[i]
            if i < 3:
                d = self._add(d, dskip[i + 1])
            d = self._stage_bwd(nb, d)
            d = self._cbr_bwd(nd, d)[0]
            self.marks[f"encoder_done_{i}"] = len(self._ops)  # gradients of encoder stage i (and of everything after it) are complete
        d = self._add(d, dskip[0])
        self._cbr_bwd(n0, d, need_gin=False)
        self._op(OP_JOIN)

    def _finish(self):
        lib, dev = self.lib, self.device
        tile0 = 0
        for r in self._layers:
            tile0 += r["K"] * ((r["c_pad"] + 31) // 32) * ((r["c_out"] + 31) // 32)
        self._ops[self._prep_op][2:4] = [len(self._layers), tile0]
        self._desc_slot = self._ext_dynamic.pop("desc")
        self._desc_t = None
        self._scratch = torch.zeros(2 * 16384, dtype=torch.float64, device=dev)
        # LGS_SIDE_PRIORITY=1: wgrad side stream at high priority (experiment knob, profiles/r2_stream_priority.txt)
        self._side = torch.cuda.Stream(dev, priority=-1 if os.environ.get("LGS_SIDE_PRIORITY", "0") != "0" else 0) if not self._dry else None
        ops = np.asarray(self._ops, dtype=np.int64)
        bufs = np.asarray(self._bufs, dtype=np.int64)
        self.n_ops = ops.shape[0]
        h = ctypes.c_void_p()
        _lib.check(lib.lgs_program_create(ops.ctypes.data_as(ctypes.c_void_p), ops.shape[0], bufs.ctypes.data_as(ctypes.c_void_p),
                                          bufs.shape[0], self.LEVELS, self._n_ext, ctypes.byref(h)))
        self._handle = h
        self._ext_arr = (ctypes.c_void_p * self._n_ext)()
        self._rows_arr = (ctypes.c_int64 * self.LEVELS)()
        self._arena = None
        self._keys = [E.CoordinateMapKey([1 << l] * 3) for l in range(self.LEVELS)]
        self._sig = None
        ids = {id(p): i for i, p in enumerate(self.reducer.params)}
        self._enc_last = max(ids[id(p)] for p in self.net.block4.parameters())
        # all-reduce buckets cut where backward completes whole groups of layers: [decoder + classifier] after the decoder,
        # [stage 4: conv4p8s2 .. block4 = 56 % of the parameters] and [stage 3] right after their backward, the small rest
        # (stem, stages 1-2) at the end — instead of three equal-byte buckets of which two could only go out after backward
        self._stage_first = {}
        for i, names in enumerate(self.ENC):
            ps = [p for nm in names for p in getattr(self.net, nm).parameters()]
            self._stage_first[i] = min(ids[id(p)] for p in ps)
        if self.reducer.world > 1 and not self.reducer._hooks:
            self.reducer.set_bounds([0, self._stage_first[2], self._stage_first[3], self._enc_last + 1, len(self.reducer.params)])
        written = set()
        for r in self._layers:
            written.add(id(r["mod"]._parameters["kernel"]))
        for key in self._ext_cache:
            if key[0] == "g":
                written.add(key[1])
        self._unwritten = [p for p in self.reducer.params if id(p) not in written]
        self._bind_static()

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                self.lib.lgs_program_destroy(h)
            except Exception:  # noqa: BLE001  (interpreter shutdown)
                pass

    # ---- per step ----------------------------------------------------------------------------------------------------
    def _signature(self):
        ps = self.reducer.params
        return (ps[0].data_ptr(), ps[-1].data_ptr(), ps[0].grad.data_ptr() if ps[0].grad is not None else 0,
                ps[-1].grad.data_ptr() if ps[-1].grad is not None else 0)

    def _bind_static(self):
        """(re)read the addresses of parameters, gradients, buffers — they move when the module is cast, reloaded or its
        gradients are re-attached — and rebuild the weight-operand descriptor table of lgs_weight_prep_batch from them"""
        if any(p.grad is None or p.grad.dtype is not torch.float32 for p in self.reducer.params):
            self.reducer.attach()
        for slot, getter in self._ext_static:
            self._ext_arr[slot] = getter().data_ptr()
        rows, tile0 = [], 0
        for r in self._layers:          # {weight, forward operand, dgrad operand (0 = none), K, c_in, c_out, first tile, 0}
            rows.append([r["src"]().data_ptr(), r["w_fwd_t"].data_ptr(), r["w_bwd_t"].data_ptr() if r["w_bwd_t"] is not None else 0,
                         r["K"], r["c_pad"], r["c_out"], tile0, 0])
            tile0 += r["K"] * ((r["c_pad"] + 31) // 32) * ((r["c_out"] + 31) // 32)
        self._desc_t = torch.tensor(rows, dtype=torch.int64).to(self.device)
        self._ext_arr[self._desc_slot] = self._desc_t.data_ptr()
        self._sig = self._signature()

    def _bind_batch(self, st, labels):
        mgr = st.coordinate_manager
        ext, dyn = self._ext_arr, self._ext_dynamic
        # kernel maps first, in network order: asking for a strided map builds the coarser coordinate map (once per
        # manager; managers created later pre-build what earlier ones were asked for, on the staging stream)
        for key, buf in self._tables.items():
            level, ks, stride, tr, which = key
            _, km = mgr.conv_maps(self._keys[level], ks, stride, 1, tr)
            t = km.fwd_table if which == "fwd" else km.bwd_table
            ext[self._bufs[buf.id][1]] = t.data_ptr()
        for level, buf in self._plans.items():
            _, km = mgr.conv_maps(self._keys[level], 3, 1, 1, False)
            ext[self._bufs[buf.id][1]] = km.plan.data_ptr() if km.plan is not None else 0
        for l, key in enumerate(self._keys):
            self._rows_arr[l] = mgr.size(key)
        f = st.F
        if f.shape[1] != self.in_pad:
            f = torch.nn.functional.pad(f, (0, self.in_pad - f.shape[1]))
        f = f.contiguous()
        lab = labels.long().contiguous()
        keep = [f, lab]
        ext[dyn["feats"]] = f.data_ptr()
        if "labels" in dyn:
            ext[dyn["labels"]] = lab.data_ptr()
        if self._logits is not None:
            self.logits = torch.empty((self._rows_arr[0], self.n_classes), dtype=torch.float32, device=self.device)
            ext[dyn["logits"]] = self.logits.data_ptr()
        if self.head is not None:
            self.features = torch.empty((self._rows_arr[0], self.feat_channels), dtype=torch.float32, device=self.device)
            ext[dyn["features"]] = self.features.data_ptr()
            ext[dyn["dfeatures"]] = self.features.data_ptr()       # placeholder until the head has run
        need = int(self.lib.lgs_program_arena_bytes(self._handle, ctypes.addressof(self._rows_arr)))
        if self._arena is None or self._arena.numel() < need:
            self._arena = torch.empty(int(need * 1.1) + 4096, dtype=torch.uint8, device=self.device)
        return keep

    def run_range(self, begin, end, join=True):
        """ops [begin, end); join=False leaves the wgrad side stream unjoined at the end of the range (the caller's collective
        waits on the side stream instead, so the training stream never stalls at a bucket boundary)"""
        rc = self.lib.lgs_program_run2(self._handle, begin, end, ctypes.addressof(self._rows_arr), ctypes.addressof(self._ext_arr),
                                       self._arena.data_ptr(), self._arena.numel(), self._scratch.data_ptr(),
                                       0 if self._dry else torch._C._cuda_getCurrentRawStream(self.device.index),
                                       0 if self._dry else self._side.cuda_stream, 0 if join else 1)
        if rc != _lib.OK:
            self._scratch.zero_()
            self.lib.lgs_program_reset(self._handle)
            _lib.check(rc)

    def run(self, st, labels):
        """forward + loss + backward; gradients are WRITTEN into every parameter's .grad (views of the reducer's flat
        buffer).  With more than one rank the all-reduce buckets are launched between the two halves of backward and joined
        before returning.  Returns the loss (0-dim device tensor, valid on the current stream)."""
        if self._signature() != self._sig:
            self._bind_static()
        self._keepalive = self._bind_batch(st, labels)
        E._weight_prep.dirty = True           # the facade's cached operands expire with every gradient step, as in autograd
        red = self.reducer
        if red.world > 1:
            red._arm()
        begin = 0
        loss = self._loss_t if self.head is None else None
        if self.head is not None:
            fwd_end = self.marks["forward_end"]
            self.run_range(0, fwd_end)
            for p in self._unwritten:                      # parameters only the head (or nothing) touches: autograd accumulates
                p.grad.zero_()
            leaf = self.features.requires_grad_(True)
            with torch.enable_grad():
                loss = self.head(leaf, labels)
                loss.backward()
            self._dfeat = leaf.grad.contiguous()
            self._ext_arr[self._ext_dynamic["dfeatures"]] = self._dfeat.data_ptr()
            loss = loss.detach()
            begin = fwd_end
        if red.world > 1:
            # a bucket goes out as soon as every parameter in it has its gradient: after the decoder, after encoder stage 4,
            # after stage 3 (run_range joins the wgrad side stream at the end of each range); the rest with red.wait()
            pos = begin
            cps = [("decoder_done", self._enc_last + 1), ("encoder_done_3", self._stage_first[3]), ("encoder_done_2", self._stage_first[2])]
            if self._deferred:          # the decoder's gradients are complete only after the deferred wgrads: first send at stage 4
                cps = cps[1:]
            main = torch.cuda.current_stream(self.device)
            for mark, first in cps:
                self.run_range(pos, self.marks[mark], join=False)
                pos = self.marks[mark]
                # the collective must see the wgrads (side stream) AND the BatchNorm / bias gradients (training stream) issued so
                # far: the side stream waits for the training stream's position, NCCL's stream for the side stream — the training
                # stream itself waits for nothing
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    for b in range(len(red._pending)):
                        if red.bounds[b] >= first:
                            red.reduce_bucket(b)
            self.run_range(pos, self.n_ops)
            red.wait()
        else:
            self.run_range(begin, self.n_ops)
        return loss
