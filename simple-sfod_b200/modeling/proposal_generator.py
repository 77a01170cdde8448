"""RPN plugins of the reference (``PseudoLabRPN`` / ``DARPN``, reference daod/modeling/proposal_generator/rpn.py:10-113)
on top of a detectron2-shaped ``RPN`` base whose ``predict_proposals`` is ONE fused call into libsfod_b200
(decode + per-image top-k + clip + non-empty filter + NMS + post top-k for all images).

The loss half of the reference's forward (``label_and_sample_anchors`` + ``losses``, rpn.py:43-50) is the student's
training step: it is restated from detectron2 0.6 in plain torch (``modeling/matcher.py``) so that
a student can train through the same plugin; it launches no kernel of this library.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor, nn

from .. import ops
from ..registry import PROPOSAL_GENERATOR_REGISTRY, RPN_HEAD_REGISTRY
from ..structures import Boxes, ImageList, Instances, LazyInstances, ShapeSpec, pairwise_iou
from ..utils.events import get_event_storage
from .anchor_generator import DefaultAnchorGenerator
from .box_regression import Box2BoxTransform
from .matcher import Matcher, dense_box_regression_loss, subsample_labels


@RPN_HEAD_REGISTRY.register()
class StandardRPNHead(nn.Module):
    """detectron2 StandardRPNHead: 3x3 conv + ReLU, then 1x1 objectness (A) and 1x1 anchor deltas (4A).
    The convolutions stay on cuDNN (BASELINE.json north_star)."""

    def __init__(self, cfg_or_channels=None, input_shape: List[ShapeSpec] = None, *, in_channels: int = None,
                 num_anchors: int = None, box_dim: int = 4):
        super().__init__()
        if in_channels is None and hasattr(cfg_or_channels, "MODEL"):
            cfg = cfg_or_channels
            chans = [s.channels for s in input_shape]
            assert len(set(chans)) == 1, "Each level must have the same channel!"
            in_channels = chans[0]
            ag = DefaultAnchorGenerator(cfg, input_shape)
            assert len(set(ag.num_anchors)) == 1, "Each level must have the same number of anchors per spatial position"
            num_anchors, box_dim = ag.num_anchors[0], ag.box_dim
        elif in_channels is None:
            in_channels = cfg_or_channels
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)
        self.objectness_logits = nn.Conv2d(in_channels, num_anchors, kernel_size=1, stride=1)
        self.anchor_deltas = nn.Conv2d(in_channels, num_anchors * box_dim, kernel_size=1, stride=1)
        for layer in (self.conv, self.objectness_logits, self.anchor_deltas):
            nn.init.normal_(layer.weight, std=0.01)
            nn.init.constant_(layer.bias, 0)

    def forward(self, features: List[Tensor]):
        pred_objectness_logits, pred_anchor_deltas = [], []
        for x in features:
            t = torch.relu(self.conv(x))
            pred_objectness_logits.append(self.objectness_logits(t))
            pred_anchor_deltas.append(self.anchor_deltas(t))
        return pred_objectness_logits, pred_anchor_deltas


class ProposalBatch:
    """Padded device-side result of one ``rpn_select`` call: boxes (N, P, 4), logits (N, P), count / invalid (N) int32.
    The host copy of the counts is fetched at most once, and only if somebody needs per-image lengths."""

    def __init__(self, boxes: Tensor, logits: Tensor, count: Tensor, invalid: Tensor, image_sizes, training: bool):
        self.boxes, self.logits, self.count, self.invalid = boxes, logits, count, invalid
        self.image_sizes = list(image_sizes)
        self.training = training
        self._host = None

    def set_host_counts(self, counts: List[int], invalid: List[int]) -> None:
        """Lets a later read of the same batch (the detections' count read) deliver these values for free."""
        if self._host is None:
            self._host = (list(counts), list(invalid))
            self._check()

    def _check(self) -> None:
        if self.training and any(v > 0 for v in self._host[1]):   # detectron2 raises here, in training, before NMS
            raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")

    def host_counts(self) -> List[int]:
        if self._host is None:
            h = torch.stack([self.count, self.invalid]).cpu().tolist()   # the one device->host read
            self._host = (h[0], h[1])
            self._check()
        return self._host[0]

    def instances(self) -> List[Instances]:
        def make(i):
            def materialize(inst):
                k = self.host_counts()[i]
                inst.proposal_boxes = Boxes(self.boxes[i, :k])
                inst.objectness_logits = self.logits[i, :k]
            return LazyInstances(self.image_sizes[i], materialize, source=self, index=i)
        return [make(i) for i in range(len(self.image_sizes))]


class RPN(nn.Module):
    """detectron2.modeling.proposal_generator.RPN (inference half)."""

    def __init__(self, cfg=None, input_shape: Dict[str, ShapeSpec] = None, *, in_features: List[str] = None, head: nn.Module = None,
                 anchor_generator: nn.Module = None, box2box_transform: Box2BoxTransform = None,
                 pre_nms_topk: Tuple[int, int] = (12000, 6000), post_nms_topk: Tuple[int, int] = (2000, 1000),
                 nms_thresh: float = 0.7, min_box_size: float = 0.0, loss_weight=1.0, anchor_matcher: Matcher = None,
                 batch_size_per_image: int = 256, positive_fraction: float = 0.5, anchor_boundary_thresh: float = -1.0,
                 box_reg_loss_type: str = "smooth_l1", smooth_l1_beta: float = 0.0):
        super().__init__()
        if cfg is not None:
            anchor_matcher = Matcher(cfg.MODEL.RPN.IOU_THRESHOLDS, cfg.MODEL.RPN.IOU_LABELS, allow_low_quality_matches=True)
            batch_size_per_image = cfg.MODEL.RPN.BATCH_SIZE_PER_IMAGE
            positive_fraction = cfg.MODEL.RPN.POSITIVE_FRACTION
            anchor_boundary_thresh = cfg.MODEL.RPN.BOUNDARY_THRESH
            box_reg_loss_type = cfg.MODEL.RPN.BBOX_REG_LOSS_TYPE
            smooth_l1_beta = cfg.MODEL.RPN.SMOOTH_L1_BETA
            in_features = cfg.MODEL.RPN.IN_FEATURES
            shapes = [input_shape[f] for f in in_features]
            anchor_generator = DefaultAnchorGenerator(cfg, shapes)
            head = RPN_HEAD_REGISTRY.get(cfg.MODEL.RPN.HEAD_NAME)(cfg, shapes)
            box2box_transform = Box2BoxTransform(weights=cfg.MODEL.RPN.BBOX_REG_WEIGHTS)
            pre_nms_topk = (cfg.MODEL.RPN.PRE_NMS_TOPK_TRAIN, cfg.MODEL.RPN.PRE_NMS_TOPK_TEST)
            post_nms_topk = (cfg.MODEL.RPN.POST_NMS_TOPK_TRAIN, cfg.MODEL.RPN.POST_NMS_TOPK_TEST)
            nms_thresh = cfg.MODEL.RPN.NMS_THRESH
            min_box_size = cfg.MODEL.PROPOSAL_GENERATOR.MIN_SIZE
            loss_weight = {"loss_rpn_cls": cfg.MODEL.RPN.LOSS_WEIGHT,
                           "loss_rpn_loc": cfg.MODEL.RPN.BBOX_REG_LOSS_WEIGHT * cfg.MODEL.RPN.LOSS_WEIGHT}
        self.in_features = in_features
        self.rpn_head = head
        self.anchor_generator = anchor_generator
        self.box2box_transform = box2box_transform
        self.pre_nms_topk = {True: pre_nms_topk[0], False: pre_nms_topk[1]}
        self.post_nms_topk = {True: post_nms_topk[0], False: post_nms_topk[1]}
        self.nms_thresh = nms_thresh
        self.min_box_size = float(min_box_size)
        if isinstance(loss_weight, float):
            loss_weight = {"loss_rpn_cls": loss_weight, "loss_rpn_loc": loss_weight}
        self.loss_weight = loss_weight
        self.anchor_matcher = anchor_matcher or Matcher([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True)
        self.batch_size_per_image, self.positive_fraction = batch_size_per_image, positive_fraction
        self.anchor_boundary_thresh = anchor_boundary_thresh
        self.box_reg_loss_type, self.smooth_l1_beta = box_reg_loss_type, smooth_l1_beta

    # ------------------------------------------------------------------ training half (student): detectron2 0.6 in plain torch
    def _subsample_labels(self, label: Tensor) -> Tensor:
        pos_idx, neg_idx = subsample_labels(label, self.batch_size_per_image, self.positive_fraction, 0)
        label.fill_(-1)
        label.scatter_(0, pos_idx, 1)
        label.scatter_(0, neg_idx, 0)
        return label

    @torch.no_grad()
    def label_and_sample_anchors(self, anchors: List[Boxes], gt_instances: List[Instances]):
        """d2 RPN.label_and_sample_anchors -> (gt_labels: List[(R,) in {-1, 0, 1}], matched_gt_boxes: List[(R, 4)]).
        On CUDA the per-image ``_subsample_labels`` calls (two ``nonzero`` syncs + two randperm each) become one sampler launch
        for the whole batch followed by a scatter of the sampled positions -- no host synchronisation at all."""
        anchors = Boxes.cat(anchors)
        gt_boxes = [x.gt_boxes for x in gt_instances]
        gt_labels, matched_gt_boxes = [], []
        batched = anchors.tensor.is_cuda and hasattr(self.anchor_matcher, "match_boxes")
        for gt_boxes_i in gt_boxes:
            if hasattr(self.anchor_matcher, "match_boxes"):   # fused pairwise_iou + Matcher on the device
                matched_idxs, gt_labels_i = self.anchor_matcher.match_boxes(gt_boxes_i, anchors)
            else:
                matched_idxs, gt_labels_i = self.anchor_matcher(pairwise_iou(gt_boxes_i, anchors))
            gt_labels_i = gt_labels_i.to(device=gt_boxes_i.device)
            if self.anchor_boundary_thresh >= 0:
                raise NotImplementedError("RPN.BOUNDARY_THRESH >= 0 is not used by any shipped config")
            if not batched:
                gt_labels_i = self._subsample_labels(gt_labels_i)
            if len(gt_boxes_i) == 0:
                matched_gt_boxes_i = torch.zeros_like(anchors.tensor)
            else:
                matched_gt_boxes_i = gt_boxes_i[matched_idxs].tensor
            gt_labels.append(gt_labels_i)
            matched_gt_boxes.append(matched_gt_boxes_i)
        if batched and len(gt_labels):
            gt_labels = list(self._subsample_labels_batched(torch.stack(gt_labels)))
        return gt_labels, matched_gt_boxes

    _sample_calls = 0

    def _sampling_seed(self) -> int:
        """Key of the counter-based sampler: a function of torch's seed and of how many batches were sampled so far."""
        type(self)._sample_calls += 1
        return (torch.initial_seed() * 0x9E3779B97F4A7C15 + 0x5DEECE66D * type(self)._sample_calls) & ((1 << 64) - 1)

    def _subsample_labels_batched(self, labels: Tensor) -> Tensor:
        """labels (N, R) int8 in {-1, 0, 1} -> the same with everything but the sampled positives (1) / negatives (0) set to -1."""
        N, R = labels.shape
        sampled, counts = ops.subsample_labels_batched(labels.reshape(-1).to(torch.int64), [R] * N, self.batch_size_per_image,
                                                       self.positive_fraction, 0, self._sampling_seed())
        S = sampled.shape[1]
        ar = torch.arange(S, device=labels.device).unsqueeze(0)
        nfg, ntot = counts[:, 0:1].to(torch.int64), counts.sum(dim=1, keepdim=True).to(torch.int64)
        vals = (ar < nfg).to(labels.dtype)                                   # 1 for the sampled positives, 0 for the negatives
        idx = torch.where(ar < ntot, sampled, torch.full_like(sampled, R))   # padding goes to a scratch column
        out = labels.new_full((N, R + 1), -1)
        out.scatter_(1, idx, vals)
        return out[:, :R]

    def losses(self, anchors: List[Boxes], pred_objectness_logits: List[Tensor], gt_labels: List[Tensor],
               pred_anchor_deltas: List[Tensor], gt_boxes: List[Tensor]) -> Dict[str, Tensor]:
        """d2 RPN.losses: BCE-with-logits on the sampled anchors + box regression on the positives, both divided by
        batch_size_per_image * num_images."""
        num_images = len(gt_labels)
        gt_labels = torch.stack(gt_labels)
        pos_mask = gt_labels == 1
        storage = get_event_storage()
        storage.put_scalar("rpn/num_pos_anchors", pos_mask.sum().item() / num_images)
        storage.put_scalar("rpn/num_neg_anchors", (gt_labels == 0).sum().item() / num_images)
        localization_loss = dense_box_regression_loss(anchors, self.box2box_transform, pred_anchor_deltas, gt_boxes, pos_mask,
                                                      box_reg_loss_type=self.box_reg_loss_type, smooth_l1_beta=self.smooth_l1_beta)
        valid_mask = gt_labels >= 0
        objectness_loss = torch.nn.functional.binary_cross_entropy_with_logits(
            torch.cat(pred_objectness_logits, dim=1)[valid_mask], gt_labels[valid_mask].to(torch.float32), reduction="sum")
        normalizer = self.batch_size_per_image * num_images
        losses = {"loss_rpn_cls": objectness_loss / normalizer, "loss_rpn_loc": localization_loss / normalizer}
        return {k: v * self.loss_weight.get(k, 1.0) for k, v in losses.items()}

    # ------------------------------------------------------------------ inference half: the hot path
    def _flatten_head_outputs(self, pred_objectness_logits: List[Tensor], pred_anchor_deltas: List[Tensor]):
        """reference rpn.py:28-41: (N, A, Hi, Wi) -> (N, Hi*Wi*A); (N, A*B, Hi, Wi) -> (N, Hi*Wi*A, B)."""
        B = self.anchor_generator.box_dim
        logits = [s.permute(0, 2, 3, 1).flatten(1) for s in pred_objectness_logits]
        deltas = [x.view(x.shape[0], -1, B, x.shape[-2], x.shape[-1]).permute(0, 3, 4, 1, 2).flatten(1, -2) for x in pred_anchor_deltas]
        return logits, deltas

    def select_proposals(self, pred_objectness_logits: List[Tensor], pred_anchor_deltas: List[Tensor], image_sizes,
                         feat_hw: Optional[List[Tuple[int, int]]] = None, anchors: Optional[List[Boxes]] = None):
        """Device-side result of ``predict_proposals`` without the host read of the counts:
        (boxes (N, P, 4), logits (N, P), src_index (N, P), count (N) int32, invalid (N) int32).
        The head outputs may be passed flattened (reference rpn.py:28-41) or as they lie ((N, A, H, W) / (N, 4A, H, W)): the
        single-level kernel chain reads the latter directly, so the teacher's forward makes no permute copy."""
        if len(pred_objectness_logits) != 1:
            if pred_objectness_logits[0].dim() == 4:
                pred_objectness_logits, pred_anchor_deltas = self._flatten_head_outputs(pred_objectness_logits, pred_anchor_deltas)
            return self._select_proposals_multilevel(pred_objectness_logits, pred_anchor_deltas, image_sizes, feat_hw, anchors)
        ag = self.anchor_generator
        kw = {}
        if feat_hw is not None and isinstance(ag, DefaultAnchorGenerator) and ag.num_anchors[0] <= 64:
            # our own forward: the grid is known, the kernel regenerates the anchors in closed form (0 B of traffic)
            kw = dict(cell_anchors=ag.host_cell_anchors[0], feat_hw=feat_hw[0], stride=ag.strides[0], anchor_offset=ag.offset)
        else:
            kw = dict(anchors=anchors[0].tensor)
        return ops.rpn_select(pred_objectness_logits[0], pred_anchor_deltas[0], image_sizes,
                              weights=self.box2box_transform.weights, scale_clamp=self.box2box_transform.scale_clamp,
                              pre_nms_topk=self.pre_nms_topk[self.training], post_nms_topk=self.post_nms_topk[self.training],
                              nms_thresh=self.nms_thresh, min_box_size=self.min_box_size, **kw)

    def _select_proposals_multilevel(self, pred_objectness_logits, pred_anchor_deltas, image_sizes, feat_hw, anchors):
        """detectron2 ``find_top_rpn_proposals`` for several feature levels (FPN-style ``RPN.IN_FEATURES``; the reference registers
        ``build_vgg_fpn_backbone``, daod/modeling/meta_arch/vgg.py:121-143, although no shipped YAML selects it).  Functional, not
        fused: per level the single-level kernel chain decodes, filters and takes the level's top ``pre_nms_topk`` with suppression
        switched off; the levels are then concatenated per image and go through ``batched_nms`` with the level as the category
        (one host read of the per-level counts + one per image for the keep count, like torchvision's own GPU path)."""
        ag = self.anchor_generator
        N = pred_objectness_logits[0].shape[0]
        pre, post = self.pre_nms_topk[self.training], self.post_nms_topk[self.training]
        per_level = []
        for l, (lg, dl) in enumerate(zip(pred_objectness_logits, pred_anchor_deltas)):
            if feat_hw is not None and isinstance(ag, DefaultAnchorGenerator) and ag.num_anchors[l] <= 64:
                kw = dict(cell_anchors=ag.host_cell_anchors[l], feat_hw=feat_hw[l], stride=ag.strides[l], anchor_offset=ag.offset)
            else:
                kw = dict(anchors=anchors[l].tensor)
            k = min(int(lg.shape[1]), pre)
            per_level.append(ops.rpn_select(lg, dl, image_sizes, weights=self.box2box_transform.weights,
                                            scale_clamp=self.box2box_transform.scale_clamp, pre_nms_topk=pre, post_nms_topk=k,
                                            nms_thresh=2.0, min_box_size=self.min_box_size, **kw))      # IoU never exceeds 2: keep all
        counts = torch.stack([torch.stack([r[3], r[4]]) for r in per_level]).cpu().tolist()           # [level][count|invalid][image]
        dev = pred_objectness_logits[0].device
        out_b = torch.zeros((N, post, 4), dtype=torch.float32, device=dev)
        out_l = torch.zeros((N, post), dtype=torch.float32, device=dev)
        out_s = torch.full((N, post), -1, dtype=torch.int64, device=dev)
        kept, invalid, base = [], [], [0]
        for lg in pred_objectness_logits:
            base.append(base[-1] + int(lg.shape[1]))
        for n in range(N):
            bx = torch.cat([r[0][n, :counts[l][0][n]] for l, r in enumerate(per_level)])
            sc = torch.cat([r[1][n, :counts[l][0][n]] for l, r in enumerate(per_level)])
            src = torch.cat([r[2][n, :counts[l][0][n]] + base[l] for l, r in enumerate(per_level)])
            lvl = torch.cat([torch.full((counts[l][0][n],), l, dtype=torch.int64, device=dev) for l in range(len(per_level))])
            keep = ops.batched_nms(bx, sc, lvl, self.nms_thresh)[:post]
            k = int(keep.numel())
            out_b[n, :k], out_l[n, :k], out_s[n, :k] = bx[keep], sc[keep], src[keep]
            kept.append(k)
            invalid.append(sum(int(counts[l][1][n]) for l in range(len(per_level))))
        return out_b, out_l, out_s, torch.tensor(kept, dtype=torch.int32, device=dev), torch.tensor(invalid, dtype=torch.int32, device=dev)

    def _closed_form_anchors(self) -> bool:
        ag = self.anchor_generator
        return isinstance(ag, DefaultAnchorGenerator) and ag.num_features == 1 and ag.num_anchors[0] <= 64

    @torch.no_grad()
    def predict_proposals(self, anchors: List[Boxes], pred_objectness_logits: List[Tensor], pred_anchor_deltas: List[Tensor],
                          image_sizes: List[Tuple[int, int]], feat_hw: Optional[List[Tuple[int, int]]] = None) -> List[Instances]:
        """d2 RPN.predict_proposals: List[Instances{proposal_boxes, objectness_logits}], score-descending."""
        boxes, logits, _, count, invalid = self.select_proposals(pred_objectness_logits, pred_anchor_deltas, image_sizes, feat_hw, anchors)
        # No host read here: the Instances are cut out of the padded batch lazily (ROI heads of this package consume the
        # padded batch directly; the counts arrive with the detections' single device->host read).
        return ProposalBatch(boxes, logits, count, invalid, image_sizes, self.training).instances()

    def forward(self, images: ImageList, features: Dict[str, Tensor], gt_instances: Optional[List[Instances]] = None):
        feats = [features[f] for f in self.in_features]
        anchors = self.anchor_generator(feats)
        pred_objectness_logits, pred_anchor_deltas = self.rpn_head(feats)
        feat_hw = [tuple(f.shape[-2:]) for f in feats]
        if self.training:
            pred_objectness_logits, pred_anchor_deltas = self._flatten_head_outputs(pred_objectness_logits, pred_anchor_deltas)
            gt_labels, gt_boxes = self.label_and_sample_anchors(anchors, gt_instances)
            losses = self.losses(anchors, pred_objectness_logits, gt_labels, pred_anchor_deltas, gt_boxes)
        else:
            losses = {}
        proposals = self.predict_proposals(anchors, pred_objectness_logits, pred_anchor_deltas, images.image_sizes, feat_hw)
        return proposals, losses


@PROPOSAL_GENERATOR_REGISTRY.register()
class PseudoLabRPN(RPN):
    """reference daod/modeling/proposal_generator/rpn.py:10-58: an RPN that can skip its loss so that a
    label-free teacher can emit proposals.  Same forward signature and return value."""

    def forward(self, images: ImageList, features: Dict[str, Tensor], gt_instances: Optional[List[Instances]] = None,
                compute_loss: bool = True, compute_val_loss: bool = False):
        feats = [features[f] for f in self.in_features]
        need_loss = (self.training and compute_loss) or compute_val_loss
        # the selection kernel regenerates the anchor grid in closed form; the anchor tensor is only built when a loss needs it
        anchors = self.anchor_generator(feats) if (need_loss or not self._closed_form_anchors()) else None
        pred_objectness_logits, pred_anchor_deltas = self.rpn_head(feats)
        feat_hw = [tuple(f.shape[-2:]) for f in feats]
        if need_loss:   # the flattened copies of rpn.py:28-41 are made for the losses only; the selection reads the head outputs as they lie
            pred_objectness_logits, pred_anchor_deltas = self._flatten_head_outputs(pred_objectness_logits, pred_anchor_deltas)
            gt_labels, gt_boxes = self.label_and_sample_anchors(anchors, gt_instances)
            losses = self.losses(anchors, pred_objectness_logits, gt_labels, pred_anchor_deltas, gt_boxes)
            losses = {k: v * self.loss_weight.get(k, 1.0) for k, v in losses.items()}
        else:  # inference
            losses = {}
        proposals = self.predict_proposals(anchors, pred_objectness_logits, pred_anchor_deltas, images.image_sizes, feat_hw)
        return proposals, losses


@PROPOSAL_GENERATOR_REGISTRY.register()
class DARPN(RPN):
    """reference daod/modeling/proposal_generator/rpn.py:61-113: an RPN that accepts unlabeled inputs -- losses only
    when training AND ``gt_instances`` is given, proposals always."""

    def forward(self, images: ImageList, features: Dict[str, Tensor], gt_instances: Optional[List[Instances]] = None):
        feats = [features[f] for f in self.in_features]
        anchors = self.anchor_generator(feats)
        pred_objectness_logits, pred_anchor_deltas = self.rpn_head(feats)
        feat_hw = [tuple(f.shape[-2:]) for f in feats]
        if self.training and gt_instances is not None:
            pred_objectness_logits, pred_anchor_deltas = self._flatten_head_outputs(pred_objectness_logits, pred_anchor_deltas)
            gt_labels, gt_boxes = self.label_and_sample_anchors(anchors, gt_instances)
            losses = self.losses(anchors, pred_objectness_logits, gt_labels, pred_anchor_deltas, gt_boxes)
        else:
            losses = {}
        proposals = self.predict_proposals(anchors, pred_objectness_logits, pred_anchor_deltas, images.image_sizes, feat_hw)
        return proposals, losses
