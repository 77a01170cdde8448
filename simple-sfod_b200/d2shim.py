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

import functools
import importlib
import sys
import types
from typing import Dict

import torch
from torch import nn

from . import config as _config
from . import ops, structures
from .engine import export as _export
from .engine import hooks as _hooks
from .modeling import box_regression, fast_rcnn, matcher, poolers, proposal_generator, roi_heads
from .modeling import resnet as _resnet
from .registry import Registry
from .structures import ShapeSpec
from .utils import events


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__dict__["__sfod_shim__"] = True
    return m


def _called_with_cfg(*args, **kwargs) -> bool:
    if len(args) and isinstance(args[0], _config.CfgNode):
        return True
    return isinstance(kwargs.get("cfg", None), _config.CfgNode)


def configurable(init_func=None, *, from_config=None):
    """detectron2.config.configurable: an ``__init__`` decorated with it accepts either explicit arguments or a ``cfg`` first
    argument, in which case ``cls.from_config(cfg, ...)`` supplies the explicit arguments (the reference's meta-architectures,
    daod/modeling/meta_arch/source_free_adaptive_teacher_rcnn.py:27-92, are written that way).  The classes of this package
    take ``(cfg, input_shape)`` directly and do not use the decorator."""
    if init_func is not None:
        assert from_config is None, "from_config is only for the function form of @configurable"

        @functools.wraps(init_func)
        def wrapped(self, *args, **kwargs):
            if _called_with_cfg(*args, **kwargs):
                try:
                    from_config_func = type(self).from_config
                except AttributeError as e:
                    raise AttributeError("Class with @configurable must have a 'from_config' classmethod.") from e
                explicit_args = from_config_func(*args, **kwargs)
                init_func(self, **explicit_args)
            else:
                init_func(self, *args, **kwargs)
        return wrapped

    def wrapper(orig_func):
        @functools.wraps(orig_func)
        def wrapped(*args, **kwargs):
            if _called_with_cfg(*args, **kwargs):
                return orig_func(**from_config(*args, **kwargs))
            return orig_func(*args, **kwargs)
        wrapped.from_config = from_config
        return wrapped
    return wrapper


class GeneralizedRCNN(nn.Module):
    """detectron2.modeling.meta_arch.rcnn.GeneralizedRCNN reduced to what the reference's subclasses inherit
    (daod/modeling/meta_arch/source_free_adaptive_teacher_rcnn.py:24 calls ``super(GeneralizedRCNN, self).__init__()`` and sets
    every attribute itself; it uses ``device``, ``preprocess_image`` and ``inference`` of the base)."""

    @property
    def device(self) -> torch.device:
        return self.pixel_mean.device

    def preprocess_image(self, batched_inputs):
        images = [x["image"].to(self.device) for x in batched_inputs]
        if images[0].is_cuda and all(im.dtype in (torch.uint8, torch.float32) for im in images):
            ms = self.__dict__.get("_sfod_mean_std")
            if ms is None:   # one device read for the lifetime of the module
                ms = (tuple(self.pixel_mean.flatten().tolist()), tuple(self.pixel_std.flatten().tolist()))
                self.__dict__["_sfod_mean_std"] = ms
            batch, sizes = ops.normalize_pad(images, ms[0], ms[1], self.backbone.size_divisibility)
            return structures.ImageList(batch, sizes)
        images = [(x - self.pixel_mean) / self.pixel_std for x in images]
        return structures.ImageList.from_tensors(images, self.backbone.size_divisibility)

    def inference(self, batched_inputs, detected_instances=None, do_postprocess: bool = True):
        assert not self.training
        images = self.preprocess_image(batched_inputs)
        features = self.backbone(images.tensor)
        if detected_instances is not None:
            raise NotImplementedError("inference with given boxes is not used on the hot path")
        if self.proposal_generator is not None:
            proposals, _ = self.proposal_generator(images, features, None)
        else:
            proposals = [x["proposals"].to(self.device) for x in batched_inputs]
        results, _ = self.roi_heads(images, features, proposals, None)
        if do_postprocess:
            return GeneralizedRCNN._postprocess(results, batched_inputs, images.image_sizes)
        return results

    @staticmethod
    def _postprocess(instances, batched_inputs, image_sizes):
        from .engine.export import detector_postprocess
        out = []
        for results_per_image, input_per_image, image_size in zip(instances, batched_inputs, image_sizes):
            height = input_per_image.get("height", image_size[0])
            width = input_per_image.get("width", image_size[1])
            out.append({"instances": detector_postprocess(results_per_image, height, width)})
        return out

    def visualize_training(self, *a, **k):
        raise NotImplementedError("visualisation (VIS_PERIOD > 0) is outside the hot path")


class DatasetEvaluator:
    """detectron2.evaluation.DatasetEvaluator interface (imported by daod/loss/bpc_loss.py:3)."""

    def reset(self):
        pass

    def process(self, inputs, outputs):
        pass

    def evaluate(self):
        pass


def _convert_image_to_rgb(*a, **k):
    raise NotImplementedError("convert_image_to_rgb is only used by visualisation, outside the hot path")


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

    def _dist_on() -> bool:
        d = torch.distributed
        return d.is_available() and d.is_initialized()

    def get_world_size() -> int:
        return torch.distributed.get_world_size() if _dist_on() else 1

    def get_rank() -> int:
        return torch.distributed.get_rank() if _dist_on() else 0

    def is_main_process() -> bool:
        return get_rank() == 0

    def synchronize() -> None:
        if _dist_on() and get_world_size() > 1:
            torch.distributed.barrier()

    def gather(data, dst: int = 0, group=None):
        """detectron2.utils.comm.gather: list of every rank's picklable ``data`` on ``dst``, [] elsewhere (reference base.py:198)."""
        if get_world_size() == 1:
            return [data]
        out = [None] * get_world_size() if get_rank() == dst else None
        torch.distributed.gather_object(data, out, dst=dst, group=group)
        return out if get_rank() == dst else []

    def all_gather(data, group=None):
        if get_world_size() == 1:
            return [data]
        out = [None] * get_world_size()
        torch.distributed.all_gather_object(out, data, group=group)
        return out

    mods = {
        "detectron2": _mod("detectron2", __path__=[]),
        "detectron2.config": _mod("detectron2.config", configurable=configurable, get_cfg=_config.get_cfg, CfgNode=_config.CfgNode),
        "detectron2.structures": _mod("detectron2.structures", Boxes=structures.Boxes, Instances=structures.Instances,
                                      ImageList=structures.ImageList, pairwise_iou=structures.pairwise_iou),
        "detectron2.layers": _mod("detectron2.layers", ShapeSpec=ShapeSpec, batched_nms=ops.batched_nms, cat=cat,
                                  cross_entropy=matcher.cross_entropy, nonzero_tuple=matcher.nonzero_tuple,
                                  # reference daod/modeling/roi_heads/box_head.py:9, daod/engine/trainers/base.py:22
                                  Conv2d=_resnet.Conv2d, get_norm=_resnet.get_norm, FrozenBatchNorm2d=_resnet.FrozenBatchNorm2d),
        "detectron2.utils": _mod("detectron2.utils", __path__=[]),
        "detectron2.utils.events": _mod("detectron2.utils.events", EventStorage=events.EventStorage, get_event_storage=events.get_event_storage),
        "detectron2.utils.comm": _mod("detectron2.utils.comm", get_world_size=get_world_size, get_rank=get_rank, is_main_process=is_main_process,
                                    synchronize=synchronize, gather=gather, all_gather=all_gather),
        "detectron2.data": _mod("detectron2.data", __path__=[]),
        "detectron2.data.detection_utils": _mod("detectron2.data.detection_utils", get_fed_loss_cls_weights=_get_fed_loss_cls_weights,
                                                convert_image_to_rgb=_convert_image_to_rgb),
        "detectron2.evaluation": _mod("detectron2.evaluation", DatasetEvaluator=DatasetEvaluator),
        "detectron2.engine": _mod("detectron2.engine", __path__=[], HookBase=_hooks.HookBase, TrainerBase=_hooks.TrainerBase),
        "detectron2.engine.hooks": _mod("detectron2.engine.hooks", HookBase=_hooks.HookBase),
        "detectron2.engine.train_loop": _mod("detectron2.engine.train_loop", HookBase=_hooks.HookBase, TrainerBase=_hooks.TrainerBase),
        "detectron2.modeling.postprocessing": _mod("detectron2.modeling.postprocessing", detector_postprocess=_export.detector_postprocess),
        "detectron2.modeling.meta_arch": _mod("detectron2.modeling.meta_arch", __path__=[], GeneralizedRCNN=GeneralizedRCNN,
                                              META_ARCH_REGISTRY=regs["META_ARCH"]),
        "detectron2.modeling.meta_arch.build": _mod("detectron2.modeling.meta_arch.build", META_ARCH_REGISTRY=regs["META_ARCH"]),
        "detectron2.modeling.meta_arch.rcnn": _mod("detectron2.modeling.meta_arch.rcnn", GeneralizedRCNN=GeneralizedRCNN),
        "detectron2.modeling": _mod("detectron2.modeling", __path__=[], build_backbone=build_backbone, build_proposal_generator=build_proposal_generator,
                                    build_roi_heads=build_roi_heads, META_ARCH_REGISTRY=regs["META_ARCH"], BACKBONE_REGISTRY=regs["BACKBONE"],
                                    ROI_HEADS_REGISTRY=regs["ROI_HEADS"], ROI_BOX_HEAD_REGISTRY=regs["ROI_BOX_HEAD"], Backbone=Backbone,
                                    StandardROIHeads=roi_heads._StandardROIHeadsBase, GeneralizedRCNN=GeneralizedRCNN),
        "detectron2.modeling.backbone": _mod("detectron2.modeling.backbone", __path__=[], Backbone=Backbone, BACKBONE_REGISTRY=regs["BACKBONE"],
                                             build_backbone=build_backbone),
        "detectron2.modeling.backbone.fpn": _mod("detectron2.modeling.backbone.fpn", FPN=FPN, LastLevelMaxPool=LastLevelMaxPool, LastLevelP6P7=LastLevelP6P7),
        "detectron2.modeling.box_regression": _mod("detectron2.modeling.box_regression", Box2BoxTransform=box_regression.Box2BoxTransform,
                                                   _dense_box_regression_loss=matcher.dense_box_regression_loss),
        "detectron2.modeling.poolers": _mod("detectron2.modeling.poolers", ROIPooler=poolers.ROIPooler, assign_boxes_to_levels=poolers.assign_boxes_to_levels),
        "detectron2.modeling.matcher": _mod("detectron2.modeling.matcher", Matcher=matcher.Matcher),
        "detectron2.modeling.sampling": _mod("detectron2.modeling.sampling", subsample_labels=matcher.subsample_labels),
        "detectron2.modeling.proposal_generator": _mod("detectron2.modeling.proposal_generator", __path__=[], RPN=proposal_generator.RPN,
                                                       PROPOSAL_GENERATOR_REGISTRY=regs["PROPOSAL_GENERATOR"],
                                                       build_proposal_generator=build_proposal_generator),
        "detectron2.modeling.proposal_generator.build": _mod("detectron2.modeling.proposal_generator.build", PROPOSAL_GENERATOR_REGISTRY=regs["PROPOSAL_GENERATOR"],
                                                             build_proposal_generator=build_proposal_generator),
        "detectron2.modeling.proposal_generator.proposal_utils": _mod("detectron2.modeling.proposal_generator.proposal_utils",
                                                                      add_ground_truth_to_proposals=matcher.add_ground_truth_to_proposals),
        "detectron2.modeling.roi_heads": _mod("detectron2.modeling.roi_heads", __path__=[], ROI_HEADS_REGISTRY=regs["ROI_HEADS"],
                                              StandardROIHeads=roi_heads._StandardROIHeadsBase, build_roi_heads=build_roi_heads,
                                              FastRCNNOutputLayers=fast_rcnn.FastRCNNOutputLayers),
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
