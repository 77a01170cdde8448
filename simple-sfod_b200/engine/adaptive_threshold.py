"""Adaptive per-class confidence threshold: the FlexMatch-style alternative pseudo-label filter of
the reference (``ADAPTIVE_THRESHOLD.ENABLED``), same names and arithmetic:

* ``AdaptiveConfidenceBasedSelfTrainingLoss`` -- reference daod/modeling/adaptive_thresh/adaptive_confidence.py:6-33
  ("convex" map ``threshold * acc / (2 - acc)``, ``confidence >= ...``);
* ``adaptive_threshold_bbox`` / ``prediction_threshold_bbox`` -- reference
  daod/engine/trainers/source_free_adaptive_teacher.py:185-254;
* ``count_label_prediction`` / ``update_adaptive_threshold`` -- reference :282-310 (including the hard-coded classes 0 and 2,
  a known reference defect that is preserved as the default and exposed as ``frozen_classes``).

The reference moves every mask through numpy on the host (``valid_map.cpu().numpy()`` per image) and hard-codes
``.cuda()``; here the selection is one ``sfod_class_threshold_select`` launch and the class histogram of a whole batch one
``sfod_class_histogram`` launch, on whatever CUDA device the detections live on.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
from torch import Tensor, nn

from .. import ops
from ..structures import Boxes, Instances


class AdaptiveConfidenceBasedSelfTrainingLoss(nn.Module):
    def __init__(self, threshold: float, num_classes: int, device=None):
        super().__init__()
        self.threshold = threshold
        self.num_classes = num_classes
        self.classwise_acc = torch.ones((self.num_classes,), device=device if device is not None else "cuda")

    def update(self, selected_labels: Tensor) -> None:
        """Update dynamic per-class accuracy."""
        if selected_labels.nelement() > 0:
            sigma = selected_labels.bincount(minlength=self.num_classes)
            self.classwise_acc = sigma / sigma.max()

    def class_thresholds(self) -> Tensor:
        acc = self.classwise_acc.to(torch.float32)
        return self.threshold * (acc / (2.0 - acc))                                   # convex

    def forward(self, confidence: Tensor, pseudo_labels: Tensor) -> Tensor:
        """Float mask (1 = keep) like the reference; the index form used by the filters is ``select``."""
        keep = self.select(confidence, pseudo_labels)
        mask = torch.zeros_like(confidence, dtype=torch.float32)
        mask[keep] = 1.0
        return mask

    def select(self, confidence: Tensor, pseudo_labels: Tensor) -> Tensor:
        n = confidence.shape[0]
        if n == 0:
            return torch.empty(0, dtype=torch.int64, device=confidence.device)
        counts = torch.tensor([n], dtype=torch.int32, device=confidence.device)
        idx, cnt = ops.class_threshold_select(confidence.reshape(1, n), pseudo_labels.reshape(1, n), counts,
                                              self.class_thresholds().to(confidence.device))
        return idx[0, : int(cnt.item())]


def _filtered(inst: Instances, keep: Tensor, as_gt: bool) -> Instances:
    out = Instances(inst.image_size)
    boxes = Boxes(inst.pred_boxes.tensor[keep, :])
    if as_gt:
        out.gt_boxes, out.gt_classes = boxes, inst.pred_classes[keep]
    else:
        out.pred_boxes, out.pred_classes = boxes, inst.pred_classes[keep]
    out.scores = inst.scores[keep]
    return out


def adaptive_threshold_bbox(criterion: AdaptiveConfidenceBasedSelfTrainingLoss, proposal_bbox_inst: Instances, thres: float = 0.7,
                            proposal_type: str = "roih") -> Instances:
    if proposal_type == "rpn":
        from .pseudo_label import threshold_bbox
        return threshold_bbox(proposal_bbox_inst, thres, "rpn")
    return _filtered(proposal_bbox_inst, criterion.select(proposal_bbox_inst.scores, proposal_bbox_inst.pred_classes), as_gt=True)


def prediction_threshold_bbox(criterion: AdaptiveConfidenceBasedSelfTrainingLoss, proposal_bbox_inst: Instances, thres: float = 0.7,
                              proposal_type: str = "roih") -> Instances:
    return _filtered(proposal_bbox_inst, criterion.select(proposal_bbox_inst.scores, proposal_bbox_inst.pred_classes), as_gt=False)


def count_label_prediction(predictions_roih_unsup_q: List[Instances], num_classes: int, bbox_threshold: float) -> Tensor:
    """Per-class count of predictions with ``score > bbox_threshold`` over the batch -> (K) float tensor (the row the
    reference writes into ``reserve_matrix``)."""
    if len(predictions_roih_unsup_q) == 0:
        raise ValueError("empty prediction list")
    dev = predictions_roih_unsup_q[0].scores.device
    stride = max(1, max(len(p) for p in predictions_roih_unsup_q))
    S = len(predictions_roih_unsup_q)
    scores = torch.zeros((S, stride), dtype=torch.float32, device=dev)
    classes = torch.zeros((S, stride), dtype=torch.int64, device=dev)
    for i, p in enumerate(predictions_roih_unsup_q):
        scores[i, : len(p)] = p.scores
        classes[i, : len(p)] = p.pred_classes
    counts = torch.tensor([len(p) for p in predictions_roih_unsup_q], dtype=torch.int32, device=dev)
    return ops.class_histogram(scores, classes, counts, num_classes, bbox_threshold).to(torch.float32)


def update_adaptive_threshold(criterion: AdaptiveConfidenceBasedSelfTrainingLoss, reserve_matrix: Tensor,
                              frozen_classes: Sequence[int] = (0, 2)) -> None:
    """reference :298-310: class-wise counter over the reserve window -> accuracy in [0, 1]; the frozen classes keep acc 1."""
    classwise_counter = reserve_matrix.sum(dim=0)
    for c in frozen_classes:
        classwise_counter[c] = 0
    criterion.classwise_acc = classwise_counter / max(classwise_counter.max(), 1)
    for c in frozen_classes:
        criterion.classwise_acc[c] = 1
