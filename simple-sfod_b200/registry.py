"""Plugin registries with detectron2's names and behaviour (fvcore.common.registry.Registry).

The reference selects its hot-path plugins by name through these registries:
``PROPOSAL_GENERATOR_REGISTRY`` (reference daod/modeling/proposal_generator/rpn.py:10, chosen by
``MODEL.PROPOSAL_GENERATOR.NAME``) and ``ROI_HEADS_REGISTRY`` (reference
daod/modeling/roi_heads/source_free_adaptive_teacher_roi_heads.py:25, chosen by ``MODEL.ROI_HEADS.NAME``).
When the real detectron2 is importable its registries are reused, so the B200 plugins register into the
very objects ``build_proposal_generator`` / ``build_roi_heads`` look names up in.
"""
from __future__ import annotations

from typing import Any, Dict, Iterator, Optional, Tuple


class Registry:
    def __init__(self, name: str) -> None:
        self._name = name
        self._obj_map: Dict[str, Any] = {}

    def _do_register(self, name: str, obj: Any) -> None:
        assert name not in self._obj_map, "An object named '{}' was already registered in '{}' registry!".format(name, self._name)
        self._obj_map[name] = obj

    def register(self, obj: Any = None) -> Any:
        if obj is None:
            def deco(func_or_class: Any) -> Any:
                self._do_register(func_or_class.__name__, func_or_class)
                return func_or_class
            return deco
        self._do_register(obj.__name__, obj)
        return obj

    def get(self, name: str) -> Any:
        ret = self._obj_map.get(name)
        if ret is None:
            raise KeyError("No object named '{}' found in '{}' registry!".format(name, self._name))
        return ret

    def __contains__(self, name: str) -> bool:
        return name in self._obj_map

    def __iter__(self) -> Iterator[Tuple[str, Any]]:
        return iter(self._obj_map.items())

    def __repr__(self) -> str:
        return "Registry of {}: {}".format(self._name, sorted(self._obj_map))


def _d2_registry(path: str, attr: str) -> Optional[Any]:
    try:  # pragma: no cover - detectron2 is absent in the build environment
        import importlib
        return getattr(importlib.import_module(path), attr)
    except Exception:
        return None


PROPOSAL_GENERATOR_REGISTRY = _d2_registry("detectron2.modeling.proposal_generator.build", "PROPOSAL_GENERATOR_REGISTRY") or Registry("PROPOSAL_GENERATOR")
ROI_HEADS_REGISTRY = _d2_registry("detectron2.modeling.roi_heads", "ROI_HEADS_REGISTRY") or Registry("ROI_HEADS")
ROI_BOX_HEAD_REGISTRY = _d2_registry("detectron2.modeling.roi_heads.box_head", "ROI_BOX_HEAD_REGISTRY") or Registry("ROI_BOX_HEAD")
BACKBONE_REGISTRY = _d2_registry("detectron2.modeling.backbone", "BACKBONE_REGISTRY") or Registry("BACKBONE")
META_ARCH_REGISTRY = _d2_registry("detectron2.modeling.meta_arch", "META_ARCH_REGISTRY") or Registry("META_ARCH")
RPN_HEAD_REGISTRY = _d2_registry("detectron2.modeling.proposal_generator.rpn", "RPN_HEAD_REGISTRY") or Registry("RPN_HEAD")
ANCHOR_GENERATOR_REGISTRY = _d2_registry("detectron2.modeling.anchor_generator", "ANCHOR_GENERATOR_REGISTRY") or Registry("ANCHOR_GENERATOR")


def build_proposal_generator(cfg, input_shape):
    """detectron2.modeling.proposal_generator.build.build_proposal_generator."""
    name = cfg.MODEL.PROPOSAL_GENERATOR.NAME
    if name == "PrecomputedProposals":
        return None
    return PROPOSAL_GENERATOR_REGISTRY.get(name)(cfg, input_shape)


def build_roi_heads(cfg, input_shape):
    """detectron2.modeling.roi_heads.build_roi_heads."""
    return ROI_HEADS_REGISTRY.get(cfg.MODEL.ROI_HEADS.NAME)(cfg, input_shape)


def build_backbone(cfg, input_shape=None):
    """detectron2.modeling.build_backbone."""
    from .structures import ShapeSpec
    if input_shape is None:
        input_shape = ShapeSpec(channels=len(cfg.MODEL.PIXEL_MEAN))
    return BACKBONE_REGISTRY.get(cfg.MODEL.BACKBONE.NAME)(cfg, input_shape)


def build_model(cfg):
    """detectron2.modeling.build_model (device placement included)."""
    import torch
    model = META_ARCH_REGISTRY.get(cfg.MODEL.META_ARCHITECTURE)(cfg)
    model.to(torch.device(cfg.MODEL.DEVICE))
    return model
