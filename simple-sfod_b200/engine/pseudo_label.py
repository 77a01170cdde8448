"""Confidence-threshold pseudo-label filter: ``threshold_bbox`` / ``process_pseudo_label`` of the reference trainer
(reference daod/engine/trainers/source_free_adaptive_teacher.py:150-183 and :256-280), same names and arguments.

When the detections come from ``FastRCNNOutputLayers.inference`` of this package, the fused post-processing call
has already counted the detections above the threshold on the device (they are score-descending, so the selected set
is a prefix); the filter is then a slice and costs no kernel and no host sync.  For any other ``Instances`` the
selection runs on ``ops.threshold_select`` (one launch for the whole list).
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from .. import ops
from ..structures import Boxes, Instances


def _prefix_from_fused(inst: Instances, thres: float):
    """Length of the ``score > thres`` prefix if ``inst`` is an untouched result of the fused post-processing call."""
    batch, i = getattr(inst, "_sfod_batch", None), getattr(inst, "_sfod_index", None)
    if batch is None or i is None or float(batch.pseudo_thresh) != float(thres) or not inst.has("scores"):
        return None
    det, pseudo = batch.host_counts()
    same_storage = inst.scores.untyped_storage().data_ptr() == batch.scores.untyped_storage().data_ptr()
    return pseudo[i] if same_storage and len(inst) == det[i] else None


def _select(values: torch.Tensor, thres: float) -> torch.Tensor:
    """Indices (ascending) of ``values > thres`` via the native kernel."""
    n = values.shape[0]
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device=values.device)
    counts = torch.tensor([n], dtype=torch.int32, device=values.device)
    idx, cnt = ops.threshold_select(values.reshape(1, n), counts, thres)
    return idx[0, : int(cnt.item())]


def threshold_bbox(proposal_bbox_inst: Instances, thres: float = 0.7, proposal_type: str = "roih") -> Instances:
    image_shape = proposal_bbox_inst.image_size
    new_proposal_inst = Instances(image_shape)
    if proposal_type == "rpn":
        keep = _select(proposal_bbox_inst.objectness_logits, thres)
        new_proposal_inst.gt_boxes = Boxes(proposal_bbox_inst.proposal_boxes.tensor[keep, :])
        new_proposal_inst.objectness_logits = proposal_bbox_inst.objectness_logits[keep]
    elif proposal_type == "roih":
        k = _prefix_from_fused(proposal_bbox_inst, thres)
        keep = slice(0, k) if k is not None else _select(proposal_bbox_inst.scores, thres)
        new_proposal_inst.gt_boxes = Boxes(proposal_bbox_inst.pred_boxes.tensor[keep, :])
        new_proposal_inst.gt_classes = proposal_bbox_inst.pred_classes[keep]
        new_proposal_inst.scores = proposal_bbox_inst.scores[keep]
    return new_proposal_inst


def process_pseudo_label(proposals_rpn_unsup_k: List[Instances], cur_threshold: float, proposal_type: str,
                         pseudo_label_method: str = "", criterion=None) -> Tuple[List[Instances], float]:
    list_instances = []
    num_proposal_output = 0.0
    for proposal_bbox_inst in proposals_rpn_unsup_k:
        if pseudo_label_method == "thresholding":
            proposal_bbox_inst = threshold_bbox(proposal_bbox_inst, thres=cur_threshold, proposal_type=proposal_type)
        elif pseudo_label_method in ("adaptive_thresholding", "prediction_thresholding"):
            if criterion is None:
                raise ValueError(f"{pseudo_label_method} needs the trainer's self_training_criterion (pass criterion=...)")
            from .adaptive_threshold import adaptive_threshold_bbox, prediction_threshold_bbox
            fn = adaptive_threshold_bbox if pseudo_label_method == "adaptive_thresholding" else prediction_threshold_bbox
            proposal_bbox_inst = fn(criterion, proposal_bbox_inst, thres=cur_threshold, proposal_type=proposal_type)
        else:
            raise ValueError("Unkown pseudo label boxes methods")
        num_proposal_output += len(proposal_bbox_inst)
        list_instances.append(proposal_bbox_inst)
    num_proposal_output = num_proposal_output / len(proposals_rpn_unsup_k)
    return list_instances, num_proposal_output
