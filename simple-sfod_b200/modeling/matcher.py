"""detectron2.modeling.matcher.Matcher, detectron2.modeling.sampling.subsample_labels and
proposal_utils.add_ground_truth_to_proposals: the student-side labelling glue that sits between NMS and
ROIAlign in every training step.  Plain torch index arithmetic, device-agnostic; the random
permutations go through ``randperm`` so that tests (and a caller that wants RNG parity) can inject the permutation."""
from __future__ import annotations

import math
from typing import Callable, List, Tuple, Union

import torch
from torch import Tensor

from ..structures import Boxes, Instances


def nonzero_tuple(x: Tensor):
    return x.nonzero(as_tuple=True) if x.dim() > 0 else x.unsqueeze(0).nonzero(as_tuple=True)


class Matcher:
    def __init__(self, thresholds: List[float], labels: List[int], allow_low_quality_matches: bool = False):
        thresholds = list(thresholds)
        assert thresholds[0] > 0
        thresholds.insert(0, -float("inf"))
        thresholds.append(float("inf"))
        assert all(low <= high for low, high in zip(thresholds[:-1], thresholds[1:]))
        assert all(l in [-1, 0, 1] for l in labels)
        assert len(labels) == len(thresholds) - 1
        self.thresholds, self.labels, self.allow_low_quality_matches = thresholds, list(labels), allow_low_quality_matches

    def __call__(self, match_quality_matrix: Tensor) -> Tuple[Tensor, Tensor]:
        """match_quality_matrix (M gt x N predictions) -> (matches (N) int64, match_labels (N) int8)."""
        assert match_quality_matrix.dim() == 2
        if match_quality_matrix.numel() == 0:
            default_matches = match_quality_matrix.new_full((match_quality_matrix.size(1),), 0, dtype=torch.int64)
            default_match_labels = match_quality_matrix.new_full((match_quality_matrix.size(1),), self.labels[0], dtype=torch.int8)
            return default_matches, default_match_labels
        assert torch.all(match_quality_matrix >= 0)
        matched_vals, matches = match_quality_matrix.max(dim=0)
        match_labels = matches.new_full(matches.size(), 1, dtype=torch.int8)
        for (l, low, high) in zip(self.labels, self.thresholds[:-1], self.thresholds[1:]):
            low_high = (matched_vals >= low) & (matched_vals < high)
            match_labels[low_high] = l
        if self.allow_low_quality_matches:
            self.set_low_quality_matches_(match_labels, match_quality_matrix)
        return matches, match_labels

    def match_boxes(self, gt_boxes, boxes) -> Tuple[Tensor, Tensor]:
        """``self(pairwise_iou(gt_boxes, boxes))`` for two ``Boxes``.  On CUDA tensors this is one fused native call
        (``sfod_iou_match``: the M x N matrix is never stored and nothing synchronises); elsewhere the two torch steps."""
        g, b = gt_boxes.tensor, boxes.tensor
        if g.is_cuda and b.is_cuda:
            from .. import ops
            matches, match_labels, _ = ops.iou_match(g, b, self.thresholds[1:-1], self.labels, self.allow_low_quality_matches)
            return matches, match_labels
        from ..structures import pairwise_iou
        return self(pairwise_iou(gt_boxes, boxes))

    def set_low_quality_matches_(self, match_labels: Tensor, match_quality_matrix: Tensor) -> None:
        highest_quality_foreach_gt, _ = match_quality_matrix.max(dim=1)
        _, pred_inds_with_highest_quality = nonzero_tuple(match_quality_matrix == highest_quality_foreach_gt[:, None])
        match_labels[pred_inds_with_highest_quality] = 1


def subsample_labels(labels: Tensor, num_samples: int, positive_fraction: float, bg_label: int,
                     randperm: Callable = torch.randperm) -> Tuple[Tensor, Tensor]:
    positive = nonzero_tuple((labels != -1) & (labels != bg_label))[0]
    negative = nonzero_tuple(labels == bg_label)[0]
    num_pos = int(num_samples * positive_fraction)
    num_pos = min(positive.numel(), num_pos)
    num_neg = num_samples - num_pos
    num_neg = min(negative.numel(), num_neg)
    perm1 = randperm(positive.numel(), device=positive.device)[:num_pos]
    perm2 = randperm(negative.numel(), device=negative.device)[:num_neg]
    return positive[perm1], negative[perm2]


def add_ground_truth_to_proposals(gt: Union[List[Instances], List[Boxes]], proposals: List[Instances]) -> List[Instances]:
    assert gt is not None
    if len(proposals) != len(gt):
        raise ValueError("proposals and gt should have the same length as the number of images!")
    if len(proposals) == 0:
        return proposals
    return [add_ground_truth_to_proposals_single_image(g, p) for g, p in zip(gt, proposals)]


def add_ground_truth_to_proposals_single_image(gt: Union[Instances, Boxes], proposals: Instances) -> Instances:
    gt_boxes = gt if isinstance(gt, Boxes) else gt.gt_boxes
    device = proposals.objectness_logits.device
    gt_logit_value = math.log((1.0 - 1e-10) / (1 - (1.0 - 1e-10)))   # a logit whose sigmoid is ~1
    gt_logits = gt_logit_value * torch.ones(len(gt_boxes), device=device)
    gt_proposal = Instances(proposals.image_size)
    gt_proposal.proposal_boxes = gt_boxes
    gt_proposal.objectness_logits = gt_logits
    for key in proposals.get_fields().keys():
        assert gt_proposal.has(key), "The attribute '{}' in `proposals` does not exist in `gt`".format(key)
    return Instances.cat([proposals, gt_proposal])


def smooth_l1_loss(input: Tensor, target: Tensor, beta: float, reduction: str = "none") -> Tensor:
    """fvcore.nn.smooth_l1_loss."""
    if beta < 1e-5:
        loss = torch.abs(input - target)
    else:
        n = torch.abs(input - target)
        cond = n < beta
        loss = torch.where(cond, 0.5 * n ** 2 / beta, n - 0.5 * beta)
    if reduction == "mean":
        loss = loss.mean() if loss.numel() > 0 else 0.0 * loss.sum()
    elif reduction == "sum":
        loss = loss.sum()
    return loss


def dense_box_regression_loss(anchors: List[Union[Boxes, Tensor]], box2box_transform, pred_anchor_deltas: List[Tensor],
                              gt_boxes: List[Tensor], fg_mask: Tensor, box_reg_loss_type: str = "smooth_l1",
                              smooth_l1_beta: float = 0.0) -> Tensor:
    """detectron2.modeling.box_regression._dense_box_regression_loss (smooth_l1 only: the default of every shipped config)."""
    if isinstance(anchors[0], Boxes):
        anchors = type(anchors[0]).cat(anchors).tensor
    else:
        anchors = torch.cat(anchors)
    if box_reg_loss_type != "smooth_l1":
        raise ValueError(f"Invalid dense box regression loss type '{box_reg_loss_type}'")
    gt_anchor_deltas = torch.stack([box2box_transform.get_deltas(anchors, k) for k in gt_boxes])
    return smooth_l1_loss(torch.cat(pred_anchor_deltas, dim=1)[fg_mask], gt_anchor_deltas[fg_mask], beta=smooth_l1_beta, reduction="sum")


def cross_entropy(input: Tensor, target: Tensor, *, reduction: str = "mean", **kwargs) -> Tensor:
    """detectron2.layers.cross_entropy (safe for an empty batch)."""
    if target.numel() == 0 and reduction == "mean":
        return input.sum() * 0.0
    return torch.nn.functional.cross_entropy(input, target, reduction=reduction, **kwargs)
