"""A minimal ``detectron2`` namespace backed by this package, for environments where detectron2 itself is absent.

The reference's plugin modules import their base classes and helpers from detectron2
(reference daod/modeling/proposal_generator/rpn.py:5-7, daod/modeling/roi_heads/source_free_adaptive_teacher_roi_heads.py:4-22,
daod/modeling/roi_heads/source_free_fast_rcnn.py:7-12, daod/modeling/meta_arch/vgg.py:6-7).  ``install()`` registers
stand-in modules under those exact dotted names whose symbols are the classes of this package (``RPN``, ``StandardROIHeads``,
``FastRCNNOutputLayers``, ``ROIPooler``, ``Box2BoxTransform``, ``Boxes`` / ``Instances`` / ``ImageList``, the registries, ...), so
that the reference's OWN ``PseudoLabRPN`` / ``SourceFreeAdaptiveTeacherStandardROIHeads`` / ``SourceFreeFastRCNNOutputLayers`` /
``vgg_backbone`` sources import unchanged and run on the B200 kernels (tests/test_reference_plugins_cpu.py does exactly that with
the files under /root/reference).  Only the symbols on the hot path are provided; it is NOT a detectron2 replacement.
When the real detectron2 is importable ``install()`` does nothing unless ``force=True``.
"""
from __future__ import annotations

import importlib
import sys
import types
from typing import Dict

import torch
from torch import nn

from . import config as _config
from . import ops, structures
from .modeling import box_regression, fast_rcnn, matcher, poolers, proposal_generator, roi_heads
from .registry import Registry
from .structures import ShapeSpec
from .utils import events


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__dict__["__sfod_shim__"] = True
    return m


def configurable(init_func=None, *, from_config=None):
    """detectron2.config.configurable, reduced to what the hot-path classes need: this package's classes already accept
    ``(cfg, input_shape)`` directly, so the decorator is the identity (on ``__init__`` and on functions)."""
    if init_func is not None:
        return init_func
    return lambda f: f


class Backbone(nn.Module):
    """detectron2.modeling.backbone.Backbone (reference daod/modeling/meta_arch/vgg.py:34 subclasses it)."""

    @property
    def size_divisibility(self) -> int:
        return 0

    @property
    def padding_constraints(self) -> Dict[str, int]:
        return {}

    def output_shape(self):
        return {name: ShapeSpec(channels=self._out_feature_channels[name], stride=self._out_feature_strides[name])
                for name in self._out_features}


class _NotOnHotPath:
    def __init__(self, *a, **k):
        raise NotImplementedError(f"{type(self).__name__} is outside the hot path (SURVEY.md section 8) and not provided by the shim")


class FPN(_NotOnHotPath):
    pass


class LastLevelMaxPool(_NotOnHotPath):
    pass


class LastLevelP6P7(_NotOnHotPath):
    pass


def _get_fed_loss_cls_weights(*a, **k):
    raise NotImplementedError("federated loss is not used by any shipped config")


def install(force: bool = False) -> bool:
    """Registers the stand-in modules in ``sys.modules``.  Returns True if the shim is (now) active."""
    if "detectron2" in sys.modules and getattr(sys.modules["detectron2"], "__sfod_shim__", False):
        return True
    if not force:
        try:
            importlib.import_module("detectron2")
            return False
        except Exception:
            pass
    regs = {n: Registry(n) for n in ("PROPOSAL_GENERATOR", "ROI_HEADS", "ROI_BOX_HEAD", "BACKBONE", "META_ARCH", "RPN_HEAD", "ANCHOR_GENERATOR")}
    regs["ROI_BOX_HEAD"]._do_register("FastRCNNConvFCHead", roi_heads.FastRCNNConvFCHead)
    regs["RPN_HEAD"]._do_register("StandardRPNHead", proposal_generator.StandardRPNHead)

    def build_box_head(cfg, input_shape):
        return regs["ROI_BOX_HEAD"].get(cfg.MODEL.ROI_BOX_HEAD.NAME)(cfg, input_shape)

    def build_backbone(cfg, input_shape=None):
        return regs["BACKBONE"].get(cfg.MODEL.BACKBONE.NAME)(cfg, input_shape or ShapeSpec(channels=len(cfg.MODEL.PIXEL_MEAN)))

    def build_proposal_generator(cfg, input_shape):
        return regs["PROPOSAL_GENERATOR"].get(cfg.MODEL.PROPOSAL_GENERATOR.NAME)(cfg, input_shape)

    def build_roi_heads(cfg, input_shape):
        return regs["ROI_HEADS"].get(cfg.MODEL.ROI_HEADS.NAME)(cfg, input_shape)

    def cat(tensors, dim: int = 0):
        assert isinstance(tensors, (list, tuple))
        return tensors[0] if len(tensors) == 1 else torch.cat(tensors, dim)

    def get_world_size() -> int:
        d = torch.distributed
        return d.get_world_size() if d.is_available() and d.is_initialized() else 1

    mods = {
        "detectron2": _mod("detectron2", __path__=[]),
        "detectron2.config": _mod("detectron2.config", configurable=configurable, get_cfg=_config.get_cfg, CfgNode=_config.CfgNode),
        "detectron2.structures": _mod("detectron2.structures", Boxes=structures.Boxes, Instances=structures.Instances,
                                      ImageList=structures.ImageList, pairwise_iou=structures.pairwise_iou),
        "detectron2.layers": _mod("detectron2.layers", ShapeSpec=ShapeSpec, batched_nms=ops.batched_nms, cat=cat,
                                  cross_entropy=matcher.cross_entropy, nonzero_tuple=matcher.nonzero_tuple),
        "detectron2.utils": _mod("detectron2.utils", __path__=[]),
        "detectron2.utils.events": _mod("detectron2.utils.events", EventStorage=events.EventStorage, get_event_storage=events.get_event_storage),
        "detectron2.utils.comm": _mod("detectron2.utils.comm", get_world_size=get_world_size, is_main_process=lambda: True),
        "detectron2.data": _mod("detectron2.data", __path__=[]),
        "detectron2.data.detection_utils": _mod("detectron2.data.detection_utils", get_fed_loss_cls_weights=_get_fed_loss_cls_weights),
        "detectron2.modeling": _mod("detectron2.modeling", __path__=[], build_backbone=build_backbone, build_proposal_generator=build_proposal_generator,
                                    build_roi_heads=build_roi_heads, META_ARCH_REGISTRY=regs["META_ARCH"], BACKBONE_REGISTRY=regs["BACKBONE"],
                                    ROI_HEADS_REGISTRY=regs["ROI_HEADS"], ROI_BOX_HEAD_REGISTRY=regs["ROI_BOX_HEAD"], Backbone=Backbone,
                                    StandardROIHeads=roi_heads._StandardROIHeadsBase),
        "detectron2.modeling.backbone": _mod("detectron2.modeling.backbone", __path__=[], Backbone=Backbone, BACKBONE_REGISTRY=regs["BACKBONE"]),
        "detectron2.modeling.backbone.fpn": _mod("detectron2.modeling.backbone.fpn", FPN=FPN, LastLevelMaxPool=LastLevelMaxPool, LastLevelP6P7=LastLevelP6P7),
        "detectron2.modeling.box_regression": _mod("detectron2.modeling.box_regression", Box2BoxTransform=box_regression.Box2BoxTransform,
                                                   _dense_box_regression_loss=matcher.dense_box_regression_loss),
        "detectron2.modeling.poolers": _mod("detectron2.modeling.poolers", ROIPooler=poolers.ROIPooler, assign_boxes_to_levels=poolers.assign_boxes_to_levels),
        "detectron2.modeling.matcher": _mod("detectron2.modeling.matcher", Matcher=matcher.Matcher),
        "detectron2.modeling.sampling": _mod("detectron2.modeling.sampling", subsample_labels=matcher.subsample_labels),
        "detectron2.modeling.proposal_generator": _mod("detectron2.modeling.proposal_generator", __path__=[], RPN=proposal_generator.RPN,
                                                       PROPOSAL_GENERATOR_REGISTRY=regs["PROPOSAL_GENERATOR"]),
        "detectron2.modeling.proposal_generator.build": _mod("detectron2.modeling.proposal_generator.build", PROPOSAL_GENERATOR_REGISTRY=regs["PROPOSAL_GENERATOR"],
                                                             build_proposal_generator=build_proposal_generator),
        "detectron2.modeling.proposal_generator.proposal_utils": _mod("detectron2.modeling.proposal_generator.proposal_utils",
                                                                      add_ground_truth_to_proposals=matcher.add_ground_truth_to_proposals),
        "detectron2.modeling.roi_heads": _mod("detectron2.modeling.roi_heads", __path__=[], ROI_HEADS_REGISTRY=regs["ROI_HEADS"],
                                              StandardROIHeads=roi_heads._StandardROIHeadsBase, build_roi_heads=build_roi_heads),
        "detectron2.modeling.roi_heads.fast_rcnn": _mod("detectron2.modeling.roi_heads.fast_rcnn", FastRCNNOutputLayers=fast_rcnn.FastRCNNOutputLayers),
        "detectron2.modeling.roi_heads.box_head": _mod("detectron2.modeling.roi_heads.box_head", build_box_head=build_box_head,
                                                       ROI_BOX_HEAD_REGISTRY=regs["ROI_BOX_HEAD"]),
    }
    for name, m in mods.items():
        sys.modules[name] = m
    for name, m in mods.items():   # parent.child attributes
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(mods[parent], child, m)
    mods["detectron2"].registries = regs
    return True


def uninstall() -> None:
    for name in [n for n, m in sys.modules.items() if n.split(".")[0] == "detectron2" and getattr(m, "__sfod_shim__", False)]:
        del sys.modules[name]
