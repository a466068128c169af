"""Explicit training step of the Res16UNet family: forward, loss and backward as ONE straight-line program over the
engine's entry points — no autograd graph, no nn.Module dispatch, no SparseTensor objects per layer.

Why: the step is host-bound as much as GPU-bound: ~12 ms of Python per 16 ms step, of which the autograd
engine, `Function.apply` and `Module.__call__` are the largest shares.  This program issues exactly the C-ABI calls the
facade issues for the same network (tests/test_step_program.py compares the recorded call traces, and checks the
gradients of the whole program against autograd on the CPU oracle engine), so it is also the executable specification of
the native (C++) step driver planned next.

Scope: nets.Res16UNet topologies in training mode.  `run` = segmentation nets (Res16UNet14A/18A/34A/34C: BASELINE
configs 1, 2, 4) with the fused mean cross-entropy; `run_features` = any flavour with a caller-supplied head (the CLIP
pre-training nets of configs 3 / 5).  One call = forward + loss + backward, parameter gradients accumulated into `.grad`
like `loss.backward()` would.  Opt-in (`StepProgram(net).run(st, labels)`); the module-by-module facade path
remains the default and the reference-facing one.

Backward order.  Nodes run in reverse creation order (what autograd does for this graph: among the ready nodes the one
created last).  Inside a residual block the creation order is conv1-bn1-relu, [downsample conv-bn], conv2-bn2(+res)-relu
(the facade evaluates lazily: the projection of the skip is materialised when `out += residual` reads it)."""
import torch

from . import _lib


class FacadeBackend:
    """the product backend: the same helper functions the MinkowskiEngine facade calls (minkowski/__init__.py)"""

    def __init__(self):
        from . import minkowski as E
        self.E = E

    def begin(self, st):
        self.mgr = st.coordinate_manager
        self.algo = self.E._state["algo"]
        return st.F, st.coordinate_map_key

    def maps(self, conv, key):
        if conv.use_mm:
            return key, None
        return self.mgr.conv_maps(key, conv._ks, conv._stride, conv._dil, conv.TRANSPOSE)

    def conv_fwd(self, x, conv, km, need_dgrad):
        y, feats, w_bwd, meta = self.E._conv_fwd_impl(x, conv._parameters["kernel"], conv.bias, km, self.algo, need_dgrad, conv)
        return y, (feats, w_bwd, meta)

    def conv_bwd(self, saved, gout, need_gin, need_gb=False):
        feats, w_bwd, meta = saved
        return self.E._conv_bwd_impl(meta, feats, w_bwd, gout, need_gin, True, need_gb)

    def bn_fwd(self, y, res, bn, relu):
        p = bn._parameters
        x, z, stats = self.E._bn_fwd_impl(y, res, p["weight"], p["bias"], bn, relu, True)
        return z, (x, z if relu else None, p["weight"], stats, relu)

    def bn_bwd(self, saved, dz, need_dres):
        x, z, gamma, stats, relu = saved
        return self.E._bn_bwd_impl(x, z, gamma, stats, dz, relu, need_dres)

    def ce(self, logits, labels, ignore_index):
        lib = _lib.load()
        n, c = logits.shape
        if not (logits.dtype is torch.float32 and lib.lgs_seg_ce_supported(c)):
            raise NotImplementedError("StepProgram needs fp32 logits with a class count that is a multiple of 4 (<= 1024)")
        logits = logits.contiguous()
        labels = labels.long().contiguous()
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        ws = torch.empty(4, dtype=torch.float64, device=logits.device)
        g = torch.empty_like(logits)
        _lib.check(lib.lgs_seg_ce(_lib.ptr(logits), n, c, _lib.ptr(labels), int(ignore_index), _lib.ptr(ws), _lib.ptr(loss),
                                  _lib.ptr(g), self.E._stream()))
        return loss, g

    cat = staticmethod(lambda a, b: torch.cat([a, b], 1))
    add = staticmethod(lambda a, b: a + b)


def _acc(p, g):
    """what AccumulateGrad does: the first gradient is kept as is, later ones are added"""
    if g is None or not p.requires_grad:
        return
    if g.shape != p.shape:
        g = g.view(p.shape)
    p.grad = g if p.grad is None else p.grad + g


class StepProgram:
    def __init__(self, net, backend=None):
        if not hasattr(net, "_enc") or not hasattr(net, "_dec"):
            raise NotImplementedError("StepProgram drives nets.Res16UNet topologies")
        self.net = net
        self.B = backend if backend is not None else FacadeBackend()

    # ---- conv + BatchNorm (+ residual) (+ ReLU): one fused node of the facade ---------------------------------
    def _cbr_fwd(self, conv, bn, x, key, relu, res=None, need_dgrad=True):
        B = self.B
        out_key, km = B.maps(conv, key)
        y, csave = B.conv_fwd(x, conv, km, need_dgrad)
        z, bsave = B.bn_fwd(y, res, bn, relu)
        return z, out_key, (conv, bn, csave, bsave, res is not None)

    def _cbr_bwd(self, node, dz, need_gin=True):
        conv, bn, csave, bsave, has_res = node
        dy, dres, dg, db = self.B.bn_bwd(bsave, dz, has_res)
        gin, gw, _ = self.B.conv_bwd(csave, dy, need_gin)
        _acc(conv._parameters["kernel"], gw)
        _acc(bn._parameters["weight"], dg)
        _acc(bn._parameters["bias"], db)
        return gin, dres

    # ---- residual block / stage ------------------------------------------------------------------------------
    def _block_fwd(self, blk, x, key):
        h, _, n1 = self._cbr_fwd(blk.conv1, blk.norm1.bn, x, key, True)
        nd, r = None, x
        if blk.downsample is not None:
            r, _, nd = self._cbr_fwd(blk.downsample[0], blk.downsample[1].bn, x, key, False)
        y, _, n2 = self._cbr_fwd(blk.conv2, blk.norm2.bn, h, key, blk.final_relu, res=r)
        return y, (n1, nd, n2)

    def _block_bwd(self, nodes, dy):
        n1, nd, n2 = nodes
        dh, dres = self._cbr_bwd(n2, dy)
        dxr = self._cbr_bwd(nd, dres)[0] if nd is not None else dres
        dx1 = self._cbr_bwd(n1, dh)[0]
        return self.B.add(dx1, dxr)

    def _stage_fwd(self, stage, x, key):
        nodes = []
        for blk in stage:
            x, n = self._block_fwd(blk, x, key)
            nodes.append(n)
        return x, nodes

    def _stage_bwd(self, nodes, d):
        for n in reversed(nodes):
            d = self._block_bwd(n, d)
        return d

    # ---- the U-Net body -------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _features_fwd(self, st):
        net, B = self.net, self.B
        x, key = B.begin(st)
        a, key, n0 = self._cbr_fwd(net.conv0p1s1, net.bn0.bn, x, key, True, need_dgrad=False)
        skips, enc, dec = [a], [], []
        for lvl, (cn, bn, bl) in enumerate(net._enc):
            a, key, nd = self._cbr_fwd(getattr(net, cn), getattr(net, bn).bn, a, key, True)
            a, nb = self._stage_fwd(getattr(net, bl), a, key)
            enc.append((nd, nb))
            if lvl < 3:
                skips.append(a)
        for lvl, (cn, bn, bl) in enumerate(net._dec):
            a, key, nt = self._cbr_fwd(getattr(net, cn), getattr(net, bn).bn, a, key, True)
            width = a.shape[1]
            a = B.cat(a, skips[3 - lvl])
            a, nb = self._stage_fwd(getattr(net, bl), a, key)
            dec.append((nt, width, nb))
        return a, key, (n0, enc, dec)

    @torch.no_grad()
    def _features_bwd(self, nodes, d):
        """backward of the body in reverse creation order; `d` = gradient w.r.t. the per-point features"""
        B = self.B
        n0, enc, dec = nodes
        dskip = [None] * 4
        for lvl in (3, 2, 1, 0):
            nt, width, nb = dec[lvl]
            d = self._stage_bwd(nb, d)
            dskip[3 - lvl] = d[:, width:]
            d = self._cbr_bwd(nt, d[:, :width])[0]
        for lvl in (3, 2, 1, 0):
            nd, nb = enc[lvl]
            if lvl < 3:
                d = B.add(d, dskip[lvl + 1])           # the stage's output also fed the decoder through `cat`
            d = self._stage_bwd(nb, d)
            d = self._cbr_bwd(nd, d)[0]
        d = B.add(d, dskip[0])
        self._cbr_bwd(n0, d, need_gin=False)

    # ---- the steps --------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def run(self, st, labels, ignore_index=-1):
        """segmentation flavour: forward + `final` classifier + mean cross-entropy (ignore_index) + backward; returns the
        loss (0-dim tensor on the device); the logits are left in `self.logits`"""
        net, B = self.net, self.B
        if getattr(net, "flavour", None) != "seg" or net.final is None:
            raise NotImplementedError("run() needs the segmentation flavour (final classifier); use run_features()")
        a, key, nodes = self._features_fwd(st)
        _, km = B.maps(net.final, key)
        logits, fsave = B.conv_fwd(a, net.final, km, True)
        loss, d = B.ce(logits, labels, ignore_index)
        d, gw, gb = B.conv_bwd(fsave, d, True, net.final.bias is not None)
        _acc(net.final._parameters["kernel"], gw)
        if net.final.bias is not None:
            _acc(net.final._parameters["bias"], gb)
        self._features_bwd(nodes, d)
        self.logits = logits
        return loss

    def run_features(self, st, head):
        """any flavour (the CLIP pre-training nets of BASELINE configs 3 / 5: Res16UNet34CR / 34CR_Proj / 34D with
        representation_only): the U-Net body runs as the explicit program, `head(features) -> scalar loss` runs under
        autograd (it is tiny: anchor projection + the fused text-anchor loss, or any criterion) and its gradient w.r.t.
        the features is fed to the body's backward.  Returns the loss; the features are left in `self.features`."""
        a, _, nodes = self._features_fwd(st)
        leaf = a.detach().requires_grad_(True)
        with torch.enable_grad():
            loss = head(leaf)
            loss.backward()
        self._features_bwd(nodes, leaf.grad)
        self.features = a
        return loss.detach()
