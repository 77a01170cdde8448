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
        """deltas (N, k*4), boxes (N, 4) -> (N, k*4); fp32, dw/dh clamped to ``scale_clamp``.
        Pseudo-labelling (no autograd) runs the native decode.  When a gradient is asked for (detectron2's decode is
        differentiable: giou / diou box losses, the reference's ``bpc_loss`` on ``convert_bbox_scores`` outputs) the same formula
        runs as torch ops -- the kernel has no backward, and silently returning a non-differentiable result would be wrong."""
        if torch.is_grad_enabled() and (deltas.requires_grad or boxes.requires_grad):
            return self._apply_deltas_torch(deltas, boxes)
        return ops.apply_deltas(deltas.float(), boxes.to(torch.float32), self.weights, self.scale_clamp)

    def _apply_deltas_torch(self, deltas: Tensor, boxes: Tensor) -> Tensor:
        """detectron2 Box2BoxTransform.apply_deltas, operation by operation."""
        deltas = deltas.float()
        boxes = boxes.to(deltas.dtype)
        widths = boxes[:, 2] - boxes[:, 0]
        heights = boxes[:, 3] - boxes[:, 1]
        ctr_x = boxes[:, 0] + 0.5 * widths
        ctr_y = boxes[:, 1] + 0.5 * heights
        wx, wy, ww, wh = self.weights
        dx = deltas[:, 0::4] / wx
        dy = deltas[:, 1::4] / wy
        dw = torch.clamp(deltas[:, 2::4] / ww, max=self.scale_clamp)
        dh = torch.clamp(deltas[:, 3::4] / wh, max=self.scale_clamp)
        pred_ctr_x = dx * widths[:, None] + ctr_x[:, None]
        pred_ctr_y = dy * heights[:, None] + ctr_y[:, None]
        pred_w = torch.exp(dw) * widths[:, None]
        pred_h = torch.exp(dh) * heights[:, None]
        x1, y1 = pred_ctr_x - 0.5 * pred_w, pred_ctr_y - 0.5 * pred_h
        x2, y2 = pred_ctr_x + 0.5 * pred_w, pred_ctr_y + 0.5 * pred_h
        return torch.stack((x1, y1, x2, y2), dim=-1).reshape(deltas.shape)
