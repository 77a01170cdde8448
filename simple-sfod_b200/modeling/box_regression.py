"""detectron2.modeling.box_regression.Box2BoxTransform with the decode on the sm_100a kernel.

``apply_deltas`` is what the reference reaches through ``RPN._decode_proposals`` (called from reference
daod/modeling/proposal_generator/rpn.py:54) and ``FastRCNNOutputLayers.predict_boxes`` (reference
daod/modeling/roi_heads/source_free_fast_rcnn.py:16); SURVEY.md A-2 restates its arithmetic."""
from __future__ import annotations

import math
from typing import Tuple

import torch
from torch import Tensor

from .. import ops

_DEFAULT_SCALE_CLAMP = math.log(1000.0 / 16)


class Box2BoxTransform:
    def __init__(self, weights: Tuple[float, float, float, float], scale_clamp: float = _DEFAULT_SCALE_CLAMP):
        self.weights = tuple(float(w) for w in weights)
        self.scale_clamp = scale_clamp

    def get_deltas(self, src_boxes: Tensor, target_boxes: Tensor) -> Tensor:
        """Training-side regression targets (plain torch: not on the pseudo-labelling path)."""
        assert isinstance(src_boxes, Tensor), type(src_boxes)
        assert isinstance(target_boxes, Tensor), type(target_boxes)
        src_w = src_boxes[:, 2] - src_boxes[:, 0]
        src_h = src_boxes[:, 3] - src_boxes[:, 1]
        src_cx = src_boxes[:, 0] + 0.5 * src_w
        src_cy = src_boxes[:, 1] + 0.5 * src_h
        tw = target_boxes[:, 2] - target_boxes[:, 0]
        th = target_boxes[:, 3] - target_boxes[:, 1]
        tcx = target_boxes[:, 0] + 0.5 * tw
        tcy = target_boxes[:, 1] + 0.5 * th
        wx, wy, ww, wh = self.weights
        deltas = torch.stack((wx * (tcx - src_cx) / src_w, wy * (tcy - src_cy) / src_h,
                              ww * torch.log(tw / src_w), wh * torch.log(th / src_h)), dim=1)
        assert (src_w > 0).all().item(), "Input boxes to Box2BoxTransform are not valid!"
        return deltas

    def apply_deltas(self, deltas: Tensor, boxes: Tensor) -> Tensor:
        """deltas (N, k*4), boxes (N, 4) -> (N, k*4); fp32, dw/dh clamped to ``scale_clamp``."""
        return ops.apply_deltas(deltas.float(), boxes.to(torch.float32), self.weights, self.scale_clamp)
