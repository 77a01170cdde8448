"""detectron2.modeling.poolers.ROIPooler on the sm_100a ROIAlign / ROIPool kernels.

Constructed by the reference at daod/modeling/roi_heads/source_free_adaptive_teacher_roi_heads.py:42-47 and called
at :117 with ``[x.proposal_boxes for x in proposals]``.  Every shipped config is single-level; the
multi-level FPN assignment is kept for API completeness (plain torch index arithmetic around the same kernels)."""
from __future__ import annotations

import math
from typing import List, Tuple, Union

import torch
from torch import Tensor, nn

from .. import ops
from ..structures import Boxes


def assign_boxes_to_levels(box_lists: List[Boxes], min_level: int, max_level: int, canonical_box_size: int, canonical_level: int):
    box_sizes = torch.sqrt(torch.cat([b.area() for b in box_lists]))
    level_assignments = torch.floor(canonical_level + torch.log2(box_sizes / canonical_box_size + 1e-8))
    level_assignments = torch.clamp(level_assignments, min=min_level, max=max_level)
    return level_assignments.to(torch.int64) - min_level


def convert_boxes_to_pooler_format(box_lists: List[Boxes]) -> Tensor:
    boxes = torch.cat([x.tensor for x in box_lists], dim=0)
    sizes = [len(b) for b in box_lists]
    idx = torch.repeat_interleave(torch.arange(len(box_lists), dtype=boxes.dtype, device=boxes.device),
                                  torch.tensor(sizes, device=boxes.device), output_size=sum(sizes))
    return torch.cat([idx[:, None], boxes], dim=1)


class ROIPooler(nn.Module):
    def __init__(self, output_size: Union[int, Tuple[int, int]], scales, sampling_ratio: int, pooler_type: str,
                 canonical_box_size: int = 224, canonical_level: int = 4):
        super().__init__()
        if isinstance(output_size, int):
            output_size = (output_size, output_size)
        assert len(output_size) == 2 and isinstance(output_size[0], int) and isinstance(output_size[1], int)
        self.output_size = output_size
        if pooler_type not in ("ROIAlign", "ROIAlignV2", "ROIPool"):
            raise ValueError("Unknown pooler type: {}".format(pooler_type))
        self.pooler_type = pooler_type
        self.sampling_ratio = sampling_ratio
        self.scales = tuple(scales)
        min_level = -(math.log2(self.scales[0]))
        max_level = -(math.log2(self.scales[-1]))
        assert math.isclose(min_level, int(min_level)) and math.isclose(max_level, int(max_level)), "Featuremap stride is not power of 2!"
        self.min_level, self.max_level = int(min_level), int(max_level)
        assert len(self.scales) == self.max_level - self.min_level + 1, "[ROIPooler] Sizes of input featuremaps do not form a pyramid!"
        assert 0 <= self.min_level <= self.max_level
        self.canonical_level = canonical_level
        assert canonical_box_size > 0
        self.canonical_box_size = canonical_box_size

    def _pool_level(self, x: Tensor, rois: Tensor, scale: float) -> Tensor:
        if self.pooler_type == "ROIPool":
            return ops.roi_pool(x, rois, self.output_size, scale)
        return ops.roi_align(x, rois, self.output_size, scale, self.sampling_ratio, aligned=(self.pooler_type == "ROIAlignV2"))

    def forward(self, x: List[Tensor], box_lists: List[Boxes]) -> Tensor:
        num_level_assignments = len(self.scales)
        assert isinstance(x, list) and isinstance(box_lists, list), "Arguments to pooler must be lists"
        assert len(x) == num_level_assignments, "unequal value, num_level_assignments={}, but x is list of {} Tensors".format(
            num_level_assignments, len(x))
        assert len(box_lists) == x[0].size(0), "unequal value, x[0] batch dim 0 is {}, but box_list has length {}".format(
            x[0].size(0), len(box_lists))
        if len(box_lists) == 0:
            return torch.zeros((0, x[0].shape[1]) + self.output_size, device=x[0].device, dtype=x[0].dtype)
        pooler_fmt_boxes = convert_boxes_to_pooler_format(box_lists)
        if num_level_assignments == 1:
            return self._pool_level(x[0], pooler_fmt_boxes, self.scales[0])
        level_assignments = assign_boxes_to_levels(box_lists, self.min_level, self.max_level, self.canonical_box_size,
                                                   self.canonical_level)
        num_boxes = pooler_fmt_boxes.size(0)
        output = torch.zeros((num_boxes, x[0].shape[1], self.output_size[0], self.output_size[1]), dtype=x[0].dtype, device=x[0].device)
        for level, (x_level, scale) in enumerate(zip(x, self.scales)):
            inds = torch.nonzero(level_assignments == level, as_tuple=True)[0]
            output.index_put_((inds,), self._pool_level(x_level, pooler_fmt_boxes[inds], scale))
        return output
