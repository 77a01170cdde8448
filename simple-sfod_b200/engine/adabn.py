"""AdaBN running-statistic recomputation: ``reset_bn_stats`` / ``recursive_traversal`` / ``adabn_refinement`` /
``test_refinement`` of the reference (reference daod/engine/trainers/base.py:270-337), same names.

Semantics preserved: the statistics are zeroed/oned (twice), then the model runs in ``train()``
mode under ``no_grad`` for at most 1400(+1) batches; each BN layer's running statistics follow the momentum-0.1
recursion.  What changes is where the BN arithmetic runs: BN layers are swapped for ``SfodBatchNorm2d`` whose
train/no_grad forward is the sm_100a statistics + fused normalise kernels.  The evaluation / checkpoint steps of
the reference's ``test_refinement`` (:300-303) are orchestration and are left to the caller (``after`` callback).
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional

import torch
from torch import nn

from ..modeling.batch_norm import convert_batchnorm


def reset_bn_stats(module: nn.Module) -> None:
    """Reset running statistics in the BatchNorm layers (they become non-trainable Parameters, reference :318-323)."""
    if isinstance(module, nn.BatchNorm2d):
        module.running_mean = nn.Parameter(torch.zeros_like(module.running_mean), requires_grad=False)
        module.running_var = nn.Parameter(torch.ones_like(module.running_var), requires_grad=False)


def recursive_traversal(module: nn.Module) -> None:
    for child in module.children():
        reset_bn_stats(child)
        recursive_traversal(child)


def test_refinement(model: nn.Module, data_loader_iter: Iterable, max_iters: int = 1400,
                    forward: Optional[Callable] = None, after: Optional[Callable] = None, process_group=None) -> int:
    """reference base.py:270-315.  ``forward(model, data)`` defaults to ``model(data)``; returns the number of
    batches consumed (the reference breaks once ``i > 1400`` *after* running batch i)."""
    convert_batchnorm(model, process_group)
    model.train()
    i = 0
    it = iter(data_loader_iter)
    while True:
        i += 1
        try:
            data = next(it)
        except StopIteration:
            i -= 1
            break
        if data is None or (hasattr(data, "__len__") and len(data) == 0):
            i -= 1
            break
        with torch.no_grad():
            result = forward(model, data) if forward is not None else model(data)
        del data, result
        if i > max_iters:
            break
    if after is not None:
        after(model)
    return i


def adabn_refinement(model: nn.Module, data_loader_iter: Iterable, max_iters: int = 1400, **kw) -> int:
    """reference base.py:330-337."""
    recursive_traversal(model)
    recursive_traversal(model)
    return test_refinement(model, data_loader_iter, max_iters, **kw)
