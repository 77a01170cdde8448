"""Fast R-CNN output layers with the inference half on the fused sm_100a post-processing call.

``FastRCNNOutputLayers`` mirrors detectron2's class; ``inference`` -- reached from reference
daod/modeling/roi_heads/source_free_adaptive_teacher_roi_heads.py:161 -- is ONE call into libsfod_b200 for all images:
softmax + class-specific decode + clip + ``score > thresh`` + per-class NMS + top-k (the per-image body is on disk at
reference daod/modeling/roi_heads/fast_rcnn.py:108-142), and the same call already counts the detections above the
pseudo-label threshold (``threshold_bbox``, reference daod/engine/trainers/source_free_adaptive_teacher.py:167-181).
``SourceFreeFastRCNNOutputLayers.convert_bbox_scores`` is the reference's no-NMS sibling
(reference daod/modeling/roi_heads/source_free_fast_rcnn.py:15-147).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor, nn

from .. import ops
from ..structures import Boxes, Instances, ShapeSpec
from ..utils.events import get_event_storage
from .box_regression import Box2BoxTransform
from .matcher import cross_entropy, dense_box_regression_loss, nonzero_tuple


def packed_proposal_batch(proposals: List[Instances]):
    """The padded RPN batch behind ``proposals`` if they are the untouched, complete, in-order lazy views of ONE batch."""
    if not proposals or not all(hasattr(p, "packed_source") for p in proposals):
        return None
    srcs = [p.packed_source() for p in proposals]
    if any(s is None for s in srcs) or any(s[0] is not srcs[0][0] for s in srcs):
        return None
    pb = srcs[0][0]
    if [s[1] for s in srcs] != list(range(len(pb.image_sizes))):
        return None
    return pb


class DetectionBatch:
    """Device-side result of the fused post-processing of one batch (padded to ``topk`` per image)."""

    def __init__(self, out: Dict[str, Tensor], image_sizes, pseudo_thresh: float, proposal_batch=None):
        self.boxes, self.scores, self.classes, self.rows = out["boxes"], out["scores"], out["classes"], out["rows"]
        self.count, self.pseudo_count = out["count"], out["pseudo_count"]
        self.image_sizes = list(image_sizes)
        self.pseudo_thresh = pseudo_thresh
        self.proposal_batch = proposal_batch          # padded RPN result these detections came from (packed path)
        self._host: Optional[List[List[int]]] = None

    def host_counts(self) -> List[List[int]]:
        """[[detections per image], [pseudo-labels per image]] -- the ONE device->host read of the batch; on the packed
        path the same read also brings the RPN's proposal counts / non-finite counts."""
        if self._host is None:
            pb = self.proposal_batch
            if pb is not None and pb._host is None:
                h = torch.stack([self.count, self.pseudo_count, pb.count, pb.invalid]).cpu().tolist()
                self._host = h[:2]
                pb.set_host_counts(h[2], h[3])
            else:
                self._host = torch.stack([self.count, self.pseudo_count]).cpu().tolist()
        return self._host

    def instances(self) -> Tuple[List[Instances], List[Tensor]]:
        det, _ = self.host_counts()
        res, kept = [], []
        for i, size in enumerate(self.image_sizes):
            k = det[i]
            r = Instances(size)
            r.pred_boxes = Boxes(self.boxes[i, :k])
            r.scores = self.scores[i, :k]
            r.pred_classes = self.classes[i, :k]
            res.append(r)
            kept.append(self.rows[i, :k])
        return res, kept

    def lazy_instances(self) -> List[Instances]:
        """The same detections as ``instances()[0]`` without touching the host now."""
        from ..structures import LazyInstances

        def make(i):
            def materialize(inst):
                k = self.host_counts()[0][i]
                inst.pred_boxes = Boxes(self.boxes[i, :k])
                inst.scores = self.scores[i, :k]
                inst.pred_classes = self.classes[i, :k]
            r = LazyInstances(self.image_sizes[i], materialize, source=self, index=i)
            object.__setattr__(r, "_sfod_batch", self)
            object.__setattr__(r, "_sfod_index", i)
            return r
        return [make(i) for i in range(len(self.image_sizes))]

    def pseudo_labels(self) -> List[Instances]:
        """threshold_bbox(proposal_type='roih') of every image: detections are score-descending, so the set with
        ``score > thres`` is the prefix of length pseudo_count."""
        _, ps = self.host_counts()
        res = []
        for i, size in enumerate(self.image_sizes):
            k = ps[i]
            r = Instances(size)
            r.gt_boxes = Boxes(self.boxes[i, :k])
            r.gt_classes = self.classes[i, :k]
            r.scores = self.scores[i, :k]
            res.append(r)
        return res


def _log_classification_stats(pred_logits: Tensor, gt_classes: Tensor, prefix: str = "fast_rcnn") -> None:
    """d2 fast_rcnn._log_classification_stats."""
    num_instances = gt_classes.numel()
    if num_instances == 0:
        return
    pred_classes = pred_logits.argmax(dim=1)
    bg_class_ind = pred_logits.shape[1] - 1
    fg_inds = (gt_classes >= 0) & (gt_classes < bg_class_ind)
    num_fg = fg_inds.nonzero().numel()
    fg_gt_classes = gt_classes[fg_inds]
    fg_pred_classes = pred_classes[fg_inds]
    num_false_negative = (fg_pred_classes == bg_class_ind).nonzero().numel()
    num_accurate = (pred_classes == gt_classes).nonzero().numel()
    fg_num_accurate = (fg_pred_classes == fg_gt_classes).nonzero().numel()
    storage = get_event_storage()
    storage.put_scalar(f"{prefix}/cls_accuracy", num_accurate / num_instances)
    if num_fg > 0:
        storage.put_scalar(f"{prefix}/fg_cls_accuracy", fg_num_accurate / num_fg)
        storage.put_scalar(f"{prefix}/false_negative", num_false_negative / num_fg)


class FastRCNNOutputLayers(nn.Module):
    """detectron2.modeling.roi_heads.fast_rcnn.FastRCNNOutputLayers: two linear layers (K+1 scores, 4K deltas)."""

    def __init__(self, cfg_or_shape, input_shape: Optional[ShapeSpec] = None, *, box2box_transform: Box2BoxTransform = None,
                 num_classes: int = None, test_score_thresh: float = 0.0, test_nms_thresh: float = 0.5,
                 test_topk_per_image: int = 100, cls_agnostic_bbox_reg: bool = False, pseudo_label_thresh: float = 0.8,
                 smooth_l1_beta: float = 0.0, box_reg_loss_type: str = "smooth_l1", loss_weight=1.0):
        super().__init__()
        if hasattr(cfg_or_shape, "MODEL"):
            cfg = cfg_or_shape
            box2box_transform = Box2BoxTransform(weights=cfg.MODEL.ROI_BOX_HEAD.BBOX_REG_WEIGHTS)
            num_classes = cfg.MODEL.ROI_HEADS.NUM_CLASSES
            cls_agnostic_bbox_reg = cfg.MODEL.ROI_BOX_HEAD.CLS_AGNOSTIC_BBOX_REG
            test_score_thresh = cfg.MODEL.ROI_HEADS.SCORE_THRESH_TEST
            test_nms_thresh = cfg.MODEL.ROI_HEADS.NMS_THRESH_TEST
            test_topk_per_image = cfg.TEST.DETECTIONS_PER_IMAGE
            pseudo_label_thresh = cfg.SEMISUPNET.BBOX_THRESHOLD
            smooth_l1_beta = cfg.MODEL.ROI_BOX_HEAD.SMOOTH_L1_BETA
            box_reg_loss_type = cfg.MODEL.ROI_BOX_HEAD.BBOX_REG_LOSS_TYPE
            loss_weight = {"loss_box_reg": cfg.MODEL.ROI_BOX_HEAD.BBOX_REG_LOSS_WEIGHT}
        else:
            input_shape = cfg_or_shape
        if isinstance(input_shape, int):
            input_shape = ShapeSpec(channels=input_shape)
        self.num_classes = num_classes
        input_size = input_shape.channels * (input_shape.width or 1) * (input_shape.height or 1)
        self.cls_score = nn.Linear(input_size, num_classes + 1)
        num_bbox_reg_classes = 1 if cls_agnostic_bbox_reg else num_classes
        box_dim = len(box2box_transform.weights)
        self.bbox_pred = nn.Linear(input_size, num_bbox_reg_classes * box_dim)
        nn.init.normal_(self.cls_score.weight, std=0.01)
        nn.init.normal_(self.bbox_pred.weight, std=0.001)
        for l in (self.cls_score, self.bbox_pred):
            nn.init.constant_(l.bias, 0)
        self.box2box_transform = box2box_transform
        self.test_score_thresh = test_score_thresh
        self.test_nms_thresh = test_nms_thresh
        self.test_topk_per_image = test_topk_per_image
        self.pseudo_label_thresh = pseudo_label_thresh
        self.smooth_l1_beta, self.box_reg_loss_type = smooth_l1_beta, box_reg_loss_type
        if isinstance(loss_weight, float):
            loss_weight = {"loss_cls": loss_weight, "loss_box_reg": loss_weight}
        self.loss_weight = loss_weight

    def forward(self, x: Tensor) -> Tuple[Tensor, Tensor]:
        if x.dim() > 2:
            x = torch.flatten(x, start_dim=1)
        return self.cls_score(x), self.bbox_pred(x)

    # ------------------------------------------------------------------ training half (student): detectron2 0.6 in plain torch
    def losses(self, predictions: Tuple[Tensor, Tensor], proposals: List[Instances]) -> Dict[str, Tensor]:
        """d2 FastRCNNOutputLayers.losses: mean cross-entropy + class-specific smooth-L1 box regression on the foreground."""
        scores, proposal_deltas = predictions
        gt_classes = (torch.cat([p.gt_classes for p in proposals], dim=0) if len(proposals)
                      else torch.empty(0, dtype=torch.int64, device=scores.device))
        _log_classification_stats(scores, gt_classes)
        if len(proposals):
            proposal_boxes = torch.cat([p.proposal_boxes.tensor for p in proposals], dim=0)
            assert not proposal_boxes.requires_grad, "Proposals should not require gradients!"
            gt_boxes = torch.cat([(p.gt_boxes if p.has("gt_boxes") else p.proposal_boxes).tensor for p in proposals], dim=0)
        else:
            proposal_boxes = gt_boxes = torch.empty((0, 4), device=proposal_deltas.device)
        losses = {"loss_cls": cross_entropy(scores, gt_classes, reduction="mean"),
                  "loss_box_reg": self.box_reg_loss(proposal_boxes, gt_boxes, proposal_deltas, gt_classes)}
        return {k: v * self.loss_weight.get(k, 1.0) for k, v in losses.items()}

    def box_reg_loss(self, proposal_boxes: Tensor, gt_boxes: Tensor, pred_deltas: Tensor, gt_classes: Tensor) -> Tensor:
        box_dim = proposal_boxes.shape[1]
        fg_inds = nonzero_tuple((gt_classes >= 0) & (gt_classes < self.num_classes))[0]
        if pred_deltas.shape[1] == box_dim:
            fg_pred_deltas = pred_deltas[fg_inds]
        else:
            fg_pred_deltas = pred_deltas.view(-1, self.num_classes, box_dim)[fg_inds, gt_classes[fg_inds]]
        loss_box_reg = dense_box_regression_loss([proposal_boxes[fg_inds]], self.box2box_transform, [fg_pred_deltas.unsqueeze(0)],
                                                 [gt_boxes[fg_inds]], ..., self.box_reg_loss_type, self.smooth_l1_beta)
        return loss_box_reg / max(gt_classes.numel(), 1.0)

    @torch.no_grad()
    def predict_boxes_for_gt_classes(self, predictions: Tuple[Tensor, Tensor], proposals: List[Instances]):
        if not len(proposals):
            return []
        scores, proposal_deltas = predictions
        proposal_boxes = torch.cat([p.proposal_boxes.tensor for p in proposals], dim=0)
        N, B = proposal_boxes.shape
        predict_boxes = self.box2box_transform.apply_deltas(proposal_deltas.detach(), proposal_boxes)
        K = predict_boxes.shape[1] // B
        if K > 1:
            gt_classes = torch.cat([p.gt_classes for p in proposals], dim=0)
            gt_classes = gt_classes.clamp_(0, K - 1)
            predict_boxes = predict_boxes.view(N, K, B)[torch.arange(N, dtype=torch.long, device=predict_boxes.device), gt_classes]
        return predict_boxes.split([len(p) for p in proposals])

    def predict_boxes(self, predictions: Tuple[Tensor, Tensor], proposals: List[Instances]) -> Tuple[Tensor, ...]:
        if not len(proposals):
            return ()
        _, proposal_deltas = predictions
        num_prop_per_image = [len(p) for p in proposals]
        proposal_boxes = torch.cat([p.proposal_boxes.tensor for p in proposals], dim=0)
        return self.box2box_transform.apply_deltas(proposal_deltas, proposal_boxes).split(num_prop_per_image)

    def predict_probs(self, predictions: Tuple[Tensor, Tensor], proposals: List[Instances]) -> Tuple[Tensor, ...]:
        scores, _ = predictions
        num_inst_per_image = [len(p) for p in proposals]
        if torch.is_grad_enabled() and scores.requires_grad:   # differentiable in detectron2; the kernel has no backward
            return torch.softmax(scores, dim=-1).split(num_inst_per_image, dim=0)
        return ops.softmax_lastdim(scores).split(num_inst_per_image, dim=0)

    @torch.no_grad()
    def inference_batch(self, predictions: Tuple[Tensor, Tensor], proposals: List[Instances],
                        pseudo_thresh: Optional[float] = None) -> DetectionBatch:
        """The fused call; nothing is read back to the host until the caller asks for Instances."""
        scores, proposal_deltas = predictions
        image_sizes = [p.image_size for p in proposals]
        thr = self.pseudo_label_thresh if pseudo_thresh is None else pseudo_thresh
        kw = dict(weights=self.box2box_transform.weights, scale_clamp=self.box2box_transform.scale_clamp,
                  score_thresh=self.test_score_thresh, nms_thresh=self.test_nms_thresh, topk=self.test_topk_per_image, pseudo_thresh=thr)
        pb = packed_proposal_batch(proposals)
        if pb is not None and scores.shape[0] == pb.boxes.shape[0] * pb.boxes.shape[1]:
            # packed path: predictions were computed on the padded (N * P) rows; valid rows come from the device-side counts
            out = ops.frcnn_postprocess(scores, proposal_deltas, pb.boxes.reshape(-1, 4), pb.count, image_sizes,
                                        rows_stride=pb.boxes.shape[1], **kw)
            return DetectionBatch(out, image_sizes, thr, proposal_batch=pb)
        rows = [len(p) for p in proposals]
        proposal_boxes = torch.cat([p.proposal_boxes.tensor for p in proposals], dim=0)
        out = ops.frcnn_postprocess(scores, proposal_deltas, proposal_boxes, rows, image_sizes, **kw)
        return DetectionBatch(out, image_sizes, thr)

    defer_host_read: bool = False   # True: detections are handed out as LazyInstances (no D2H read inside the forward)

    def inference(self, predictions: Tuple[Tensor, Tensor], proposals: List[Instances]):
        """d2 FastRCNNOutputLayers.inference -> (List[Instances{pred_boxes, scores, pred_classes}], List[kept row indices]).
        The per-image lengths need the device-side counts on the host: ONE read for the whole batch, here -- or, with
        ``defer_host_read`` (CUDA-graph capture, engine/graph.py), when somebody first looks at a result."""
        batch = self.inference_batch(predictions, proposals)
        if self.defer_host_read:
            return batch.lazy_instances(), [batch.rows[i] for i in range(len(batch.image_sizes))]
        instances, kept = batch.instances()
        for i, inst in enumerate(instances):  # keeps the fused pseudo-label counts reachable from the d2-shaped result
            inst._sfod_batch, inst._sfod_index = batch, i
        return instances, kept


class SourceFreeFastRCNNOutputLayers(FastRCNNOutputLayers):
    """reference daod/modeling/roi_heads/source_free_fast_rcnn.py:14-147."""

    @torch.no_grad()
    def convert_bbox_scores(self, predictions, proposals):
        """Decode + softmax + clip, ``scores > 0`` and NO NMS: dense per-class instances for ``bpc_loss``
        (reference source_free_fast_rcnn.py:82-147).  Decode and softmax run on the sm_100a kernels; the boolean-mask
        gathers of this student-side helper are plain torch indexing."""
        boxes = self.predict_boxes(predictions, proposals)
        scores = self.predict_probs(predictions, proposals)
        image_shapes = [x.image_size for x in proposals]
        results, kept = [], []
        for b, s, shape in zip(boxes, scores, image_shapes):
            valid_mask = torch.isfinite(b).all(dim=1) & torch.isfinite(s).all(dim=1)
            if not valid_mask.all():
                b, s = b[valid_mask], s[valid_mask]
            s = s[:, :-1]
            k = b.shape[1] // 4
            bx = Boxes(b.reshape(-1, 4))
            bx.clip(shape)
            b3 = bx.tensor.view(-1, k, 4)
            filter_mask = s > 0
            filter_inds = filter_mask.nonzero()
            bsel = b3[filter_inds[:, 0], 0] if k == 1 else b3[filter_mask]
            result = Instances(shape)
            result.pred_boxes = Boxes(bsel)
            result.scores = s[filter_mask]
            result.pred_classes = filter_inds[:, 1]
            results.append(result)
            kept.append(filter_inds[:, 0])
        return results, kept
