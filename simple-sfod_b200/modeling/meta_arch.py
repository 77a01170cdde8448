"""Teacher-side meta-architecture: the ``unsup_data_weak`` branch of the reference's
``SourceFreeAdaptiveTeacherGeneralizedRCNN`` (reference daod/modeling/meta_arch/source_free_adaptive_teacher_rcnn.py:310-339)
plus detectron2's ``GeneralizedRCNN.inference``.  This is the caller of the hot path: preprocess ->
backbone (cuDNN convs + native BN) -> PseudoLabRPN -> ROI heads -> detections; the pseudo-label filter is fused in.
The student-side branches ``supervised`` / ``supervised_target`` (reference ...rcnn.py:233-300) are provided in their
detector-loss form so that a student can train through the same plugins (RPN + ROI-head losses; ROIAlign forward and
backward on the sm_100a kernels).  ``bpc_loss`` enters the reference's total with weight 0 (reference
daod/engine/trainers/source_free_adaptive_teacher.py:549-550) and is returned as an exact zero; the ``domain_classifier``
branch (DA baselines) is out of scope and raises NotImplementedError.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
from torch import Tensor, nn

from ..registry import META_ARCH_REGISTRY, build_backbone, build_proposal_generator, build_roi_heads
from .. import ops
from ..structures import ImageList


class _GradientScalar(torch.autograd.Function):
    """Gradient-reversal layer (reference daod/modeling/dann/dann.py:33-51): identity forward, grad * alpha backward."""

    @staticmethod
    def forward(ctx, x, alpha):
        ctx.alpha = alpha
        return x.view_as(x)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.clone() * ctx.alpha, None


class _ImageDomainClassifier(nn.Module):
    """Parameter layout of the reference's image-level discriminator (reference daod/modeling/dann/dann.py:10-29).  It is not
    evaluated on the teacher's pseudo-labelling branch; it exists so that the teacher/student state_dicts -- and therefore
    the EMA update (47 636 547 elements) -- have the reference's keys and sizes."""

    def __init__(self, in_channels: int, ndf1: int = 256, ndf2: int = 128):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, ndf1, 3, padding=1)
        self.conv2 = nn.Conv2d(ndf1, ndf2, 3, padding=1)
        self.conv3 = nn.Conv2d(ndf2, ndf2, 3, padding=1)
        self.classifier = nn.Conv2d(ndf2, 1, 3, padding=1)

    def forward(self, x: Tensor) -> Tensor:
        act = nn.functional.leaky_relu
        return self.classifier(act(self.conv3(act(self.conv2(act(self.conv1(x), 0.2)), 0.2)), 0.2))


class _InstanceDomainClassifier(nn.Module):
    """Parameter layout of the reference's instance-level discriminator (reference daod/modeling/dann/dann.py:97-155)."""

    def __init__(self, in_channels: int, levels: List[str]):
        super().__init__()
        for level in levels:
            for i, (a, b) in enumerate(((in_channels, 1024), (1024, 1024), (1024, 1)), start=1):
                fc = nn.Linear(a, b)
                nn.init.normal_(fc.weight, std=0.01)
                nn.init.constant_(fc.bias, 0)
                self.add_module("da_ins_fc{}_level_{}".format(i, level), fc)


@META_ARCH_REGISTRY.register()
class SourceFreeAdaptiveTeacherGeneralizedRCNN(nn.Module):
    def __init__(self, cfg=None, *, backbone: nn.Module = None, proposal_generator: nn.Module = None, roi_heads: nn.Module = None,
                 pixel_mean=(103.530, 116.280, 123.675), pixel_std=(1.0, 1.0, 1.0), dis_type: Optional[str] = None,
                 ins_dc: bool = False):
        super().__init__()
        if cfg is not None:
            dis_type, ins_dc = cfg.SEMISUPNET.DIS_TYPE, cfg.SEMISUPNET.INS_DC
            backbone = build_backbone(cfg)
            proposal_generator = build_proposal_generator(cfg, backbone.output_shape())
            roi_heads = build_roi_heads(cfg, backbone.output_shape())
            pixel_mean, pixel_std = cfg.MODEL.PIXEL_MEAN, cfg.MODEL.PIXEL_STD
        self.backbone, self.proposal_generator, self.roi_heads = backbone, proposal_generator, roi_heads
        self.register_buffer("pixel_mean", torch.tensor(pixel_mean).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.tensor(pixel_std).view(-1, 1, 1), False)
        assert self.pixel_mean.shape == self.pixel_std.shape
        self._mean_std = (tuple(float(v) for v in pixel_mean), tuple(float(v) for v in pixel_std))
        self.dis_type, self.ins_dc = dis_type, ins_dc
        if dis_type is not None and dis_type in getattr(backbone, "_out_feature_channels", {}):
            self.DC_img = _ImageDomainClassifier(backbone._out_feature_channels[dis_type])  # reference ...rcnn.py:68
            if ins_dc:
                self.DC_ins = _InstanceDomainClassifier(roi_heads.box_predictor.cls_score.in_features, [dis_type])  # :71

    @property
    def device(self) -> torch.device:
        return self.pixel_mean.device

    def preprocess_image(self, batched_inputs: List[Dict[str, Tensor]]) -> ImageList:
        """d2 GeneralizedRCNN.preprocess_image (reference ...rcnn.py:92-104): normalise, pad, batch."""
        images = [x["image"].to(self.device, non_blocking=True) for x in batched_inputs]
        if images[0].is_cuda and all(im.dtype in (torch.uint8, torch.float32) for im in images):
            # one native pass per image: normalise + zero-pad straight into the batch slot (sfod_normalize_pad)
            batch, sizes = ops.normalize_pad(images, self._mean_std[0], self._mean_std[1], self.backbone.size_divisibility)
            return ImageList(batch, sizes)
        images = [(x - self.pixel_mean) / self.pixel_std for x in images]
        return ImageList.from_tensors(images, self.backbone.size_divisibility)

    def preprocess_batch(self, images: Tensor) -> ImageList:
        """Same arithmetic for an already batched (N, 3, H, W) tensor of equally sized images (one launch)."""
        images = images.to(self.device, non_blocking=True)
        if images.is_cuda and images.dtype in (torch.uint8, torch.float32):
            nhwc = images.is_contiguous(memory_format=torch.channels_last) and not images.is_contiguous()
            x, sizes = ops.normalize_pad(images, self._mean_std[0], self._mean_std[1], 0, channels_last=nhwc)
            return ImageList(x, sizes)
        x = (images - self.pixel_mean) / self.pixel_std
        return ImageList(x, [tuple(images.shape[-2:])] * images.shape[0])

    def forward(self, batched_inputs, branch: str = "supervised", given_proposals=None, val_mode: bool = False):
        if not self.training and not val_mode:
            return self.inference(batched_inputs)
        if branch not in ("unsup_data_weak", "supervised", "supervised_target"):
            raise NotImplementedError(f"branch {branch!r} (domain-adversarial baselines, SURVEY.md section 2 rows 8-9) is out of scope")
        images = self.preprocess_batch(batched_inputs) if isinstance(batched_inputs, Tensor) else self.preprocess_image(batched_inputs)
        features = self.backbone(images.tensor)
        if branch in ("supervised", "supervised_target"):
            gt_instances = [x["instances"].to(self.device) for x in batched_inputs] if "instances" in batched_inputs[0] else None
            proposals_rpn, proposal_losses = self.proposal_generator(images, features, gt_instances)
            out = self.roi_heads(images, features, proposals_rpn, compute_loss=True, targets=gt_instances, branch=branch)
            detector_losses = out[1]
            losses = {}
            losses.update(detector_losses)
            losses.update(proposal_losses)
            if branch == "supervised":
                if hasattr(self, "DC_img"):      # reference ...rcnn.py:234-236, :259
                    d_out = self.DC_img(_GradientScalar.apply(features[self.dis_type], -1.0))
                    losses["loss_DC_img_s"] = nn.functional.binary_cross_entropy_with_logits(d_out, torch.zeros_like(d_out)) * 0.001
                return losses, [], []
            proposals_roih, _ = self.roi_heads(images, features, proposals_rpn, targets=None, compute_loss=False, branch=branch)
            losses["loss_bpc"] = next(iter(detector_losses.values())).new_zeros(())   # weight 0 in the reference's total
            return losses, proposals_roih, [], []
        proposals_rpn, _ = self.proposal_generator(images, features, None, compute_loss=False)
        proposals_roih, _ = self.roi_heads(images, features, proposals_rpn, targets=None, compute_loss=False, branch=branch)
        return {}, proposals_rpn, proposals_roih

    @torch.no_grad()
    def inference(self, batched_inputs, detected_instances=None, do_postprocess: bool = False):
        assert not self.training
        images = self.preprocess_batch(batched_inputs) if isinstance(batched_inputs, Tensor) else self.preprocess_image(batched_inputs)
        features = self.backbone(images.tensor)
        proposals, _ = self.proposal_generator(images, features, None)
        results, _ = self.roi_heads(images, features, proposals, None, compute_loss=False)
        return [{"instances": r} for r in results]
