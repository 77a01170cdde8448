"""VGG16(-BN) backbone with the reference's stage layout and state_dict keys (reference daod/modeling/meta_arch/vgg.py:10-118).

The convolutions and max-pools stay on cuDNN (BASELINE.json north_star); what this module adds is the BatchNorm path:
in train()/no_grad forwards -- every teacher forward and every AdaBN iteration of the reference
-- each ``conv -> BN -> ReLU`` triple runs ``conv`` on cuDNN and ``BN(+ReLU)`` on the two sm_100a kernels (statistics at
4 B/element, fused normalise+ReLU in place at 8 B/element).  Stages: vgg0..vgg4 with strides 2..32 (vgg.py:60-74).
"""
from __future__ import annotations

from typing import Dict, List, Union, cast

from torch import Tensor, nn

from ..registry import BACKBONE_REGISTRY
from ..structures import ShapeSpec
from .batch_norm import SfodBatchNorm2d

cfgs: Dict[str, List[Union[str, int]]] = {
    "vgg16": [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"],
}


def make_layers(cfg: List[Union[str, int]], batch_norm: bool = False) -> List[nn.Module]:
    layers: List[nn.Module] = []
    in_channels = 3
    for v in cfg:
        if v == "M":
            layers += [nn.MaxPool2d(kernel_size=2, stride=2)]
        else:
            v = cast(int, v)
            conv2d = nn.Conv2d(in_channels, v, kernel_size=3, padding=1)
            layers += [conv2d, SfodBatchNorm2d(v), nn.ReLU(inplace=True)] if batch_norm else [conv2d, nn.ReLU(inplace=True)]
            in_channels = v
    return layers


def _is_pool2(m: nn.Module) -> bool:
    def two(v):
        return v == 2 or v == (2, 2)
    return (isinstance(m, nn.MaxPool2d) and two(m.kernel_size) and two(m.stride) and m.padding in (0, (0, 0))
            and m.dilation in (1, (1, 1)) and not m.ceil_mode and not m.return_indices)


class _Stage(nn.Sequential):
    """nn.Sequential (same child names, hence the reference's state_dict keys) whose forward hands the elementwise
    neighbours of a natively executed BatchNorm to the BN kernels: the bias of the preceding convolution (which then
    runs bias-free on cuDNN), the following ReLU and, at the end of a stage, the 2x2 max-pool.  Per element the
    arithmetic is unchanged: fl(conv + bias) -> normalise -> ReLU -> max."""

    def forward(self, x: Tensor) -> Tensor:
        mods = list(self)
        i, n = 0, len(mods)
        while i < n:
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < n else None
            if isinstance(m, nn.Conv2d) and isinstance(nxt, SfodBatchNorm2d) and nxt._native_ok(x):
                z = nn.functional.conv2d(x, m.weight, None, m.stride, m.padding, m.dilation, m.groups)  # bias goes to the BN kernels
                relu = i + 2 < n and isinstance(mods[i + 2], nn.ReLU)
                pool = relu and i + 3 < n and _is_pool2(mods[i + 3])
                x = nxt(z, fuse_relu=relu, inplace=True, pre_bias=m.bias, fuse_maxpool=pool)  # z is private to this forward
                i += 2 + int(relu) + int(pool)
                continue
            if isinstance(m, SfodBatchNorm2d) and isinstance(nxt, nn.ReLU) and m._native_ok(x):
                x = m(x, fuse_relu=True, inplace=i > 0)  # in place only on an intermediate produced inside this stage
                i += 2
                continue
            x = m(x)
            i += 1
        return x


class vgg_backbone(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        layers = make_layers(cfgs["vgg16"], batch_norm=cfg.VGG.BN)
        self._initialize_weights(layers)
        if cfg.VGG.BN:
            cuts = [(0, 7), (7, 14), (14, 24), (24, 34), (34, len(layers))]
        else:  # the reference's slicing assumes BN (vgg.py:70-74); without BN the same stages end after each pool
            cuts = [(0, 5), (5, 10), (10, 17), (17, 24), (24, len(layers))]
        out_channels, out_strides = [64, 128, 256, 512, 512], [2, 4, 8, 16, 32]
        self._out_feature_channels, self._out_feature_strides, self._stage_names = {}, {}, []
        self.stages = []
        for i, (a, b) in enumerate(cuts):
            name = "vgg{}".format(i)
            stage = _Stage(*layers[a:b])
            self.add_module(name, stage)
            self.stages.append(stage)
            self._stage_names.append(name)
            self._out_feature_channels[name] = out_channels[i]
            self._out_feature_strides[name] = out_strides[i]
        self._out_features = self._stage_names

    @property
    def size_divisibility(self) -> int:
        return 0

    def output_shape(self) -> Dict[str, ShapeSpec]:
        return {n: ShapeSpec(channels=self._out_feature_channels[n], stride=self._out_feature_strides[n]) for n in self._out_features}

    def forward(self, x: Tensor) -> Dict[str, Tensor]:
        features = {}
        for name, stage in zip(self._stage_names, self.stages):
            x = stage(x)
            features[name] = x
        return features

    @staticmethod
    def _initialize_weights(layers) -> None:
        for m in layers:
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)


@BACKBONE_REGISTRY.register()
def build_vgg_backbone(cfg, _=None):
    return vgg_backbone(cfg)
