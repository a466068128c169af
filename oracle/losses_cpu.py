"""ORACLE (test infrastructure, NOT product code) — CPU restatement of the CLIP text-anchor losses.

Follows /root/reference/lib/losses/ContrastiveLanguageLoss.py and lib/losses/utils.py:
  * clip_ce_loss      ContrastiveLanguageCELoss.forward :224-237 with feat_dist 'cos' :215-218
                      (S = normalize(F) @ normalize(A)^T, no temperature, CrossEntropyLoss(ignore_index))
  * clip_hinge_loss   ContrastiveLanguageLoss.forward :184-192 with feat_dist 'cos' :87-93
                      (the negative anchor ids, drawn on the host by np.random.choice at :131-138 of the
                       reference, are an explicit INPUT here so the restatement is deterministic)
  * feature_sim       lib/losses/utils.py:99-103 (cosine branch) -> argmax prediction

PINNED: clip_ce_loss is checked against the reference class itself, imported from /root/reference in the
build container; the vectors are committed under tests/golden/ (tests/golden/make_golden.py).
The hinge class needs CUDA in the reference (torch.cuda.FloatTensor, :113-114) and is therefore pinned
only through its formula (parity unpinned for that variant).
"""
import torch
import torch.nn.functional as F


def cosine_logits(feats, anchors):
    return F.normalize(feats, p=2, dim=1) @ F.normalize(anchors, p=2, dim=1).t()


def clip_ce_loss(feats, labels, anchors, ignore_label=-1, reduction="mean"):
    return F.cross_entropy(cosine_logits(feats, anchors), labels.long(), ignore_index=ignore_label,
                           reduction=reduction)


def seg_ce_loss(logits, labels, ignore_label=-1):
    """nn.CrossEntropyLoss(ignore_index=...) on the class logits (lib/train_test/pl_BaselineTrainer.py:343,350), evaluated
    in float64 so that it can referee fp32 implementations; the formula IS torch's, so this one is pinned by construction."""
    return F.cross_entropy(logits.double(), labels.long(), ignore_index=ignore_label)


def clip_hinge_loss(feats, labels, anchors, neg_ids, pos_thresh=0.0, neg_thresh=0.6, neg_weight=1.0,
                    ignore_label=-1, reduction="mean"):
    """neg_ids: [N, k] anchor ids of the sampled negatives (ignored rows may hold anything valid)."""
    S = cosine_logits(feats, anchors)
    valid = labels != ignore_label
    y = labels.long().clamp(min=0)
    pos_d = 1.0 - S.gather(1, y[:, None]).squeeze(1)
    neg_d = 1.0 - S.gather(1, neg_ids.long()).mean(1)
    pos_d = torch.where(valid, pos_d, torch.zeros_like(pos_d))     # feat_dist :93  loss[target==ignore] = 0
    neg_d = torch.where(valid, neg_d, torch.zeros_like(neg_d))
    pos_loss = torch.relu(pos_d - pos_thresh)
    neg_loss = torch.relu(neg_thresh - neg_d)
    if reduction == "mean":
        return pos_loss.mean() + neg_loss.mean() * neg_weight, pos_loss, neg_loss
    return pos_loss + neg_loss * neg_weight, pos_loss, neg_loss


def feature_sim(feats, anchors):
    return cosine_logits(feats, anchors)
