"""detectron2.modeling.anchor_generator.DefaultAnchorGenerator.

The cell anchors live on the host as well as in a buffer: the RPN kernel regenerates the (H, W, A) grid in
closed form from them (0 bytes of anchor traffic), while ``forward`` still returns the ``List[Boxes]`` the
reference passes around (reference daod/modeling/proposal_generator/rpn.py:25)."""
from __future__ import annotations

import math
from typing import List, Sequence

import torch
from torch import nn

from ..registry import ANCHOR_GENERATOR_REGISTRY
from ..structures import Boxes, ShapeSpec


def _broadcast_params(params, num_features: int, name: str):
    assert isinstance(params, (list, tuple)), f"{name} in anchor generator has to be a list! Got {params}."
    assert len(params), f"{name} in anchor generator cannot be empty!"
    if not isinstance(params[0], (list, tuple)):
        return [params] * num_features
    if len(params) == 1:
        return list(params) * num_features
    assert len(params) == num_features, f"Got {name} of length {len(params)} in anchor generator, but the number of input features is {num_features}!"
    return params


@ANCHOR_GENERATOR_REGISTRY.register()
class DefaultAnchorGenerator(nn.Module):
    box_dim: int = 4

    def __init__(self, cfg_or_sizes=None, input_shape: List[ShapeSpec] = None, *, sizes=None, aspect_ratios=None, strides=None,
                 offset: float = 0.5):
        super().__init__()
        if sizes is None and cfg_or_sizes is not None and hasattr(cfg_or_sizes, "MODEL"):
            cfg = cfg_or_sizes
            sizes = cfg.MODEL.ANCHOR_GENERATOR.SIZES
            aspect_ratios = cfg.MODEL.ANCHOR_GENERATOR.ASPECT_RATIOS
            strides = [x.stride for x in input_shape]
            offset = cfg.MODEL.ANCHOR_GENERATOR.OFFSET
        elif sizes is None:
            sizes = cfg_or_sizes
        self.strides = list(strides)
        self.num_features = len(self.strides)
        sizes = _broadcast_params(sizes, self.num_features, "sizes")
        aspect_ratios = _broadcast_params(aspect_ratios, self.num_features, "aspect_ratios")
        cells = [self.generate_cell_anchors(s, a) for s, a in zip(sizes, aspect_ratios)]
        self.host_cell_anchors = [c.clone() for c in cells]
        for i, c in enumerate(cells):
            self.register_buffer(f"cell_anchors_{i}", c, persistent=False)
        self.offset = offset
        assert 0.0 <= self.offset < 1.0, self.offset

    @property
    def cell_anchors(self) -> List[torch.Tensor]:
        return [getattr(self, f"cell_anchors_{i}") for i in range(self.num_features)]

    @property
    def num_anchors(self) -> List[int]:
        return [len(c) for c in self.host_cell_anchors]

    @property
    def num_cell_anchors(self) -> List[int]:
        return self.num_anchors

    @staticmethod
    def generate_cell_anchors(sizes: Sequence[float] = (32, 64, 128, 256, 512), aspect_ratios: Sequence[float] = (0.5, 1, 2)):
        anchors = []
        for size in sizes:
            area = size ** 2.0
            for aspect_ratio in aspect_ratios:
                w = math.sqrt(area / aspect_ratio)
                h = aspect_ratio * w
                anchors.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
        return torch.tensor(anchors)

    def _grid_anchors(self, grid_sizes) -> List[torch.Tensor]:
        anchors = []
        for size, stride, base in zip(grid_sizes, self.strides, self.cell_anchors):
            gh, gw = size
            sx = torch.arange(self.offset * stride, gw * stride, step=stride, dtype=torch.float32, device=base.device)
            sy = torch.arange(self.offset * stride, gh * stride, step=stride, dtype=torch.float32, device=base.device)
            shift_y, shift_x = torch.meshgrid(sy, sx, indexing="ij")
            shift_x, shift_y = shift_x.reshape(-1), shift_y.reshape(-1)
            shifts = torch.stack((shift_x, shift_y, shift_x, shift_y), dim=1)
            anchors.append((shifts.view(-1, 1, 4) + base.view(1, -1, 4)).reshape(-1, 4))
        return anchors

    def forward(self, features: List[torch.Tensor]) -> List[Boxes]:
        grid_sizes = [f.shape[-2:] for f in features]
        return [Boxes(x) for x in self._grid_anchors(grid_sizes)]
