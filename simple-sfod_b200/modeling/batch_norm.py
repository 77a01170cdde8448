"""BatchNorm2d whose train-mode / no_grad forward (the AdaBN and mean-teacher case) runs on the sm_100a kernels.

In the reference every teacher forward happens in ``train()`` mode under ``torch.no_grad()`` (reference
daod/engine/trainers/source_free_adaptive_teacher.py:385-390; the model is never ``.eval()``-ed, :61-65) and
AdaBN is 1400 such forwards after ``reset_bn_stats`` (reference daod/engine/trainers/base.py:270-337).  Those
forwards need batch statistics, the momentum update of the running statistics and the normalised output, but no
autograd graph: exactly what ``ops.bn_train_forward`` computes (statistics pass at 4 B/element, fused
normalise+ReLU pass at 8 B/element).  Every other mode defers to ``nn.BatchNorm2d`` (cuDNN).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor, nn

from .. import ops


class SfodBatchNorm2d(nn.BatchNorm2d):
    """Drop-in ``nn.BatchNorm2d`` (same parameters, buffers and state_dict keys).

    ``process_group``: when set (``True`` = default group), train-mode/no_grad forwards all-reduce the per-channel
    (sum, sum of squares, count) so that every rank normalises with the statistics of the concatenated batch
    (the reference keeps per-rank statistics, which is the ``None`` default).
    """

    def __init__(self, *args, process_group=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.process_group = process_group

    def _native_ok(self, x: Tensor) -> bool:
        return (self.training and not torch.is_grad_enabled() and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4
                and self.track_running_stats and self.momentum is not None)

    def forward(self, x: Tensor, fuse_relu: bool = False, inplace: bool = False, pre_bias: Optional[Tensor] = None,
                fuse_maxpool: bool = False, residual: Optional[Tensor] = None) -> Tensor:
        """``pre_bias`` / ``fuse_relu`` / ``fuse_maxpool`` / ``residual`` let a caller that owns the surrounding ``conv -> BN
        [+ shortcut] -> ReLU [-> MaxPool2d(2, 2)]`` chain (``modeling/vgg.py``, ``modeling/resnet.py``) hand the neighbours'
        elementwise work to the BN kernels.

        The native kernels serve exactly one mode -- train() under no_grad on CUDA fp32 NCHW/NHWC, the mode of every teacher
        forward and AdaBN iteration of the reference.  Any other mode (autograd needed: the student; eval: frozen statistics;
        CPU tensors) is NOT this library's hot path and runs the stock ``nn.BatchNorm2d`` of PyTorch (cuDNN / ATen), i.e. the
        reference's own operator, with the fused neighbours applied in plain torch.  That is a deliberate operator boundary, not
        a fallback of the kernels: ``ops.bn_train_forward`` itself raises on anything it does not implement."""
        if self._native_ok(x):
            return ops.bn_train_forward(x, self.weight, self.bias, self.running_mean, self.running_var,
                                        self.num_batches_tracked, self.momentum, self.eps, fuse_relu=fuse_relu,
                                        inplace=inplace, group=self.process_group, pre_bias=pre_bias, fuse_maxpool=fuse_maxpool,
                                        residual=residual)
        if pre_bias is not None:
            x = x + pre_bias.view(1, -1, 1, 1)
        y = super().forward(x)
        if residual is not None:
            y = y + residual
        if fuse_relu:
            y = torch.relu_(y) if not torch.is_grad_enabled() else torch.relu(y)
        return torch.nn.functional.max_pool2d(y, 2, 2) if fuse_maxpool else y


def convert_batchnorm(module: nn.Module, process_group=None) -> nn.Module:
    """Swap every ``nn.BatchNorm2d`` of ``module`` for a ``SfodBatchNorm2d`` sharing the same tensors
    (recursive, in place; the traversal of reference base.py:325-328)."""
    for name, child in list(module.named_children()):
        if isinstance(child, nn.BatchNorm2d) and not isinstance(child, SfodBatchNorm2d):
            new = SfodBatchNorm2d(child.num_features, eps=child.eps, momentum=child.momentum, affine=child.affine,
                                  track_running_stats=child.track_running_stats, process_group=process_group)
            new.weight, new.bias = child.weight, child.bias
            if child.track_running_stats:
                # running stats may already be Parameters (after reset_bn_stats); keep whatever they are
                for key in ("running_mean", "running_var", "num_batches_tracked"):
                    if key in child._parameters:
                        new._buffers.pop(key, None)
                        new._parameters[key] = child._parameters[key]
                    else:
                        new._buffers[key] = child._buffers[key]
            new.train(child.training)
            setattr(module, name, new)
        else:
            convert_batchnorm(child, process_group)
    return module
