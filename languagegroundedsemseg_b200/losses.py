"""CLIP text-anchor losses through the fused CUDA kernels (include/lgs_b200.h: lgs_clip_ce / lgs_clip_hinge).

Mirrors the reference classes' call signature — ``criterion(features, labels, anchor_feats)``:
  ContrastiveLanguageCELoss   lib/losses/ContrastiveLanguageLoss.py:197-237
  ContrastiveLanguageLoss     lib/losses/ContrastiveLanguageLoss.py:13-194 (hinge; 'cos' distance)
  feature_sim (argmax preds)  lib/losses/utils.py:80-103
"""
import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


# 'tc'  : tcgen05 kernel (3xTF32 GEMM + fused softmax-CE + second GEMM for dF) whenever the shape allows   [default]
# 'simt': the exact-fp32 FMA kernel (parity anchor; also serves shapes outside the tensor-core envelope)
_clip = {"algo": "tc"}


def set_clip_algo(name: str):
    if name not in ("tc", "simt"):
        raise ValueError(name)
    _clip["algo"] = name


def _stream():
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


class _ClipCEFn(torch.autograd.Function):
    """per-row loss; gradient w.r.t. features computed in the same launch; anchors' gradient (learned projection,
    models/clip_models.py:197-200) through the optional [n,a] grad-logits output and one plain GEMM."""

    @staticmethod
    def forward(ctx, feats, anchors_n, labels, ignore_label):
        lib = _lib.load()
        feats = feats.float().contiguous()
        anchors_n = anchors_n.float().contiguous()
        labels = labels.long().contiguous()
        n, c = feats.shape
        a = anchors_n.shape[0]
        loss = torch.empty(n, dtype=torch.float32, device=feats.device)
        pred = torch.empty(n, dtype=torch.int32, device=feats.device)
        need_f = ctx.needs_input_grad[0]
        need_a = ctx.needs_input_grad[1]
        gf = torch.empty_like(feats) if need_f else None
        gl = torch.empty((n, a), dtype=torch.float32, device=feats.device) if need_a else None
        if _clip["algo"] == "tc" and lib.lgs_clip_ce_tc_supported(c, a):
            ws = torch.empty(lib.lgs_clip_ce_tc_ws_elems(c, a), dtype=torch.float32, device=feats.device)
            _lib.check(lib.lgs_clip_ce_tc(_lib.ptr(feats), n, c, _lib.ptr(anchors_n), a, _lib.ptr(labels),
                                          int(ignore_label), _lib.ptr(loss), _lib.ptr(gf), _lib.ptr(pred), _lib.ptr(gl),
                                          _lib.ptr(ws), _stream()))
        else:
            _lib.check(lib.lgs_clip_ce(_lib.ptr(feats), n, c, _lib.ptr(anchors_n), a, _lib.ptr(labels),
                                       int(ignore_label), _lib.ptr(loss), _lib.ptr(gf), _lib.ptr(pred), _lib.ptr(gl),
                                       _stream()))
        ctx.save_for_backward(gf, gl, feats if need_a else None)
        ctx.mark_non_differentiable(pred)
        return loss, pred

    @staticmethod
    def backward(ctx, gloss, _gpred):
        gf, gl, feats = ctx.saved_tensors
        g_feats = gf * gloss[:, None] if gf is not None else None
        g_anch = None
        if gl is not None:
            g_anch = (gl * gloss[:, None]).t() @ F.normalize(feats, p=2, dim=1)
        return g_feats, g_anch, None, None


def clip_ce(feats, labels, anchor_feats, ignore_label=-1):
    """-> (per-point loss [n], argmax prediction [n]); anchors are L2-normalised here (tiny [a,c] op)."""
    return _ClipCEFn.apply(feats, F.normalize(anchor_feats.float(), p=2, dim=1), labels, ignore_label)


class ContrastiveLanguageCELoss(nn.Module):
    def __init__(self, config=None, num_labels=200, reduction="mean", ignore_label=None):
        super().__init__()
        self.ignore_label = ignore_label if ignore_label is not None else getattr(config, "ignore_label", -1)
        self.num_labels, self.reduction = num_labels, reduction
        self.last_pred = None

    def forward(self, features, labels, anchor_feats, preds=None):
        if features.dim() != 2:
            raise ValueError("`features` needs to be [n_points, feat_dim]")
        loss, self.last_pred = clip_ce(features, labels, anchor_feats, self.ignore_label)
        if self.reduction == "mean":   # nn.CrossEntropyLoss(ignore_index): mean over non-ignored points
            loss = loss.sum() / (labels != self.ignore_label).sum().clamp(min=1)
        elif self.reduction == "sum":
            loss = loss.sum()
        return loss, torch.zeros(1), loss   # same 3-tuple as the reference (:239)


class _SegCEFn(torch.autograd.Function):
    """mean softmax cross-entropy over the non-ignored points + its gradient in one pass (lgs_seg_ce)"""

    @staticmethod
    def forward(ctx, logits, labels, ignore_index):
        lib = _lib.load()
        n, c = logits.shape
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        ws = torch.empty(4, dtype=torch.float64, device=logits.device)
        g = torch.empty_like(logits) if ctx.needs_input_grad[0] else None
        _lib.check(lib.lgs_seg_ce(_lib.ptr(logits), n, c, _lib.ptr(labels), int(ignore_index), _lib.ptr(ws),
                                  _lib.ptr(loss), _lib.ptr(g), _stream()))
        ctx.save_for_backward(g)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        (g,) = ctx.saved_tensors
        # out of place: a second backward through this node (retain_graph, two losses sharing it) must see the same `g`
        return (g * gloss if g is not None else None), None, None


def cross_entropy(input, target, ignore_index=-100):
    """`nn.CrossEntropyLoss(ignore_index=...)(soutput.F, target)` of lib/train_test/pl_BaselineTrainer.py:343,350
    (criterion from loss_by_name, :96; reduction 'mean') as one fused pass over the logits.  Shapes the kernel does not
    take (class count not a multiple of 4 or above 1024, non-fp32 logits) use ATen's cross_entropy on the same device."""
    if (input.is_cuda and input.dim() == 2 and input.dtype is torch.float32 and input.shape[0] > 0
            and _lib.load().lgs_seg_ce_supported(input.shape[1])):
        return _SegCEFn.apply(input.contiguous(), target.long().contiguous(), ignore_index)
    return F.cross_entropy(input.float(), target.long(), ignore_index=ignore_index)


class CrossEntropyLoss(nn.Module):
    """drop-in for the `nn.CrossEntropyLoss(ignore_index=config.ignore_label)` the reference trains with"""

    def __init__(self, ignore_index=-100):
        super().__init__()
        self.ignore_index = ignore_index

    def forward(self, input, target):
        return cross_entropy(input, target, self.ignore_index)


class _ClipHingeFn(torch.autograd.Function):
    """per-point hinge loss  total_i = pos_i + neg_weight * neg_i  (differentiable) and its two parts (for logging).
    Gradients: features from the kernel; anchors as a scatter-add of coef * normalize(F_i) rows —
    d pos_i / dA[y_i] = -[pos_i > 0] F^_i,  d neg_i / dA[n_ij] = [neg_i > 0] F^_i / n_neg  (ContrastiveLanguageLoss.py:
    184-192 backpropagates through `anchor_feats`, which trains `projection_layer` of Res16UNet34CR_Proj)."""

    @staticmethod
    def forward(ctx, feats, anchors_n, labels, neg_ids, ignore_label, pos_thresh, neg_thresh, neg_weight):
        lib = _lib.load()
        feats = feats.float().contiguous()
        anchors_n = anchors_n.float().contiguous()
        labels = labels.long().contiguous()
        neg_ids = neg_ids.to(torch.int32).contiguous()
        n, c = feats.shape
        pos = torch.empty(n, dtype=torch.float32, device=feats.device)
        neg = torch.empty_like(pos)
        gf = torch.empty_like(feats) if ctx.needs_input_grad[0] else None
        _lib.check(lib.lgs_clip_hinge(_lib.ptr(feats), n, c, _lib.ptr(anchors_n), anchors_n.shape[0], _lib.ptr(labels),
                                      _lib.ptr(neg_ids), neg_ids.shape[1], int(ignore_label), float(pos_thresh),
                                      float(neg_thresh), float(neg_weight), _lib.ptr(pos), _lib.ptr(neg),
                                      _lib.ptr(gf), _stream()))
        ctx.save_for_backward(gf, pos, neg, feats, labels, neg_ids)
        ctx.neg_weight, ctx.n_anchors, ctx.ignore = float(neg_weight), anchors_n.shape[0], int(ignore_label)
        ctx.mark_non_differentiable(pos, neg)
        return pos + float(neg_weight) * neg, pos, neg

    @staticmethod
    def backward(ctx, gtot, _gpos, _gneg):
        gf, pos, neg, feats, labels, neg_ids = ctx.saved_tensors
        g_feats = gf * gtot[:, None] if (gf is not None and ctx.needs_input_grad[0]) else None
        g_anch = None
        if ctx.needs_input_grad[1]:
            fh = F.normalize(feats, p=2, dim=1)
            valid = labels != ctx.ignore
            g_anch = torch.zeros((ctx.n_anchors, feats.shape[1]), dtype=torch.float32, device=feats.device)
            cp = torch.where(valid & (pos > 0), -gtot, torch.zeros_like(gtot))
            g_anch.index_add_(0, labels.clamp(min=0), fh * cp[:, None])
            n_neg = neg_ids.shape[1]
            cn = torch.where(valid & (neg > 0), gtot * (ctx.neg_weight / n_neg), torch.zeros_like(gtot))
            for j in range(n_neg):
                g_anch.index_add_(0, neg_ids[:, j].long(), fh * cn[:, None])
        return g_feats, g_anch, None, None, None, None, None, None


class ContrastiveLanguageLoss(nn.Module):
    """Hinge variant.  Negatives are drawn on the device (uniform over anchors != label, the reference's
    clip_uniform_sampling branch :131-134) unless `neg_ids` is given."""

    def __init__(self, config=None, num_labels=200, reduction="mean", ignore_label=None, num_negative_samples=3,
                 pos_thresh=0.0, neg_thresh=0.6, neg_weight=1.0):
        super().__init__()
        g = lambda k, d: getattr(config, k, d) if config is not None else d
        self.ignore_label = ignore_label if ignore_label is not None else g("ignore_label", -1)
        self.num_labels, self.reduction = num_labels, reduction
        self.num_negative_samples = g("num_negative_samples", num_negative_samples)
        if self.num_negative_samples < 0:      # 'use all labels' (ContrastiveLanguageLoss.py:33-36)
            self.num_negative_samples = num_labels
        self.pos_thresh = g("contrast_pos_thresh", pos_thresh)
        self.neg_thresh = g("contrast_neg_thresh", neg_thresh)
        self.neg_weight = g("contrast_neg_weight", neg_weight)

    def sample_negatives(self, labels, generator=None):
        n, a = labels.shape[0], self.num_labels
        r = torch.randint(0, a - 1, (n, self.num_negative_samples), device=labels.device, generator=generator)
        y = labels.long().clamp(min=0)[:, None]
        return (r + (r >= y).long()).to(torch.int32)   # uniform over {0..a-1} \ {y}

    def forward(self, features, labels, anchor_feats, preds=None, neg_ids=None):
        if features.dim() != 2:
            raise ValueError("`features` needs to be [n_points, feat_dim]")
        if anchor_feats.dim() == 3:
            anchor_feats = anchor_feats[:, 0, :]
        if neg_ids is None:
            neg_ids = self.sample_negatives(labels)
        an = F.normalize(anchor_feats.float(), p=2, dim=1)
        total, pos, neg = _ClipHingeFn.apply(features, an, labels, neg_ids, self.ignore_label, self.pos_thresh,
                                             self.neg_thresh, self.neg_weight)
        loss = total.mean() if self.reduction == "mean" else total      # = pos.mean() + w * neg.mean()
        return loss, pos, neg


def feature_sim_argmax(feats, anchor_feats):
    """argmax_c cos(F_i, A_c) (lib/losses/utils.py:99-103 + the argmax at pl_RepresentationTrainer.py:238-239)."""
    n = feats.shape[0]
    labels = torch.full((n,), -1, dtype=torch.long, device=feats.device)
    with torch.no_grad():
        _, pred = clip_ce(feats.detach(), labels, anchor_feats, -1)
    return pred.long()


def sample_categories_for_balancing(loss, config, dataset, targets, outputs=None, generator=None):
    """lib/losses/utils.py:13-77 without the per-class host loop (`.item()` syncs, np.random.choice on the CPU):
    per-point loss [n] (or only the valid points'), targets [n]; `dataset.frequency_organized_cats` bool [NUM_LABELS,3]
    marks head / common / tail categories.  Head (common) points are kept with probability mass
    `config.balanced_sample_head_ratio` (`..._common_ratio`): exactly round(ratio * count) points of every head (common)
    class, chosen uniformly on the device; ratio <= 0 keeps them all (the reference's default, -1); tail points are always
    kept.  Returns (mean of the masked loss, (head, common, tail) losses detached, [n_valid,3] membership) like the
    reference.  With ratios <= 0 the result equals the reference's exactly; with sampling the kept COUNT per class is the
    reference's and the choice is uniform, but the random stream is torch's, not numpy's."""
    ignore = config.ignore_label
    if loss.shape[0] != targets.shape[0]:
        targets = targets[targets != ignore]
    dev = loss.device
    valid = targets != ignore
    cats = dataset.frequency_organized_cats.to(dev)
    t = targets.long().clamp(min=0)
    member = cats[t] & valid[:, None]                              # [n,3] head / common / tail
    member[:, 2] = valid & ~member[:, 0] & ~member[:, 1]           # the reference's `else` branch: everything else is tail
    keep = member[:, 2].clone()
    n_labels = cats.shape[0]
    for col, ratio in ((0, config.balanced_sample_head_ratio), (1, config.balanced_sample_common_ratio)):
        sel = member[:, col]
        if ratio > 0.0:
            # rank every selected point inside its class by a random key; keep the first round(ratio * count) of each class
            counts = torch.bincount(t[sel], minlength=n_labels)
            quota = torch.round(ratio * counts.double()).long()   # python's round() of the reference is half-to-even too
            key = torch.rand(t.shape[0], device=dev, generator=generator)
            order = torch.argsort(torch.where(sel, t.double() + key.double(), torch.full_like(key, float("inf")).double()))
            sorted_cls = t[order]
            first = torch.searchsorted(sorted_cls[: int(sel.sum())].contiguous(), torch.arange(n_labels, device=dev))
            rank = torch.empty_like(t)
            n_sel = int(sel.sum())
            rank[order[:n_sel]] = torch.arange(n_sel, device=dev) - first[sorted_cls[:n_sel]]
            keep |= sel & (rank < quota[t])
        else:
            keep |= sel
    head_loss, common_loss, tail_loss = (loss[member[:, c]].detach() for c in range(3))
    masked = loss * keep.to(loss.dtype)
    return masked.mean(), (head_loss, common_loss, tail_loss), member[valid]
