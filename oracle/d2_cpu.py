"""CPU oracle for the simple-SFOD pseudo-labelling hot path.  TEST INFRASTRUCTURE ONLY.

This module is *not* part of the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product package (``simple-sfod_b200``) never imports anything from ``oracle/``.

What it is
----------
A line-by-line CPU restatement of what detectron2 0.6 executes for the reference's
hot path, running on the *real* native CPU ops that the reference
binds to: ``torchvision.ops.{nms,batched_nms,roi_align,roi_pool}`` and ATen
(``sort``, ``softmax``, ``exp``, ``batch_norm``).  detectron2 itself cannot be
installed in this environment (no network, not in the wheelhouse), therefore the
detectron2 *glue* is restated from its published 0.6 source,
anchored on the reference's own call sites:

* ``daod/modeling/proposal_generator/rpn.py:25-58``  (tensor layouts, predict_proposals)
* ``daod/modeling/roi_heads/fast_rcnn.py:88-142``    (verbatim copy of d2's
  ``fast_rcnn_inference_single_image`` -- the only on-disk statement)
* ``daod/modeling/roi_heads/source_free_fast_rcnn.py:82-147`` (no-NMS variant)
* ``daod/modeling/roi_heads/source_free_adaptive_teacher_roi_heads.py:108-163``
* ``daod/engine/trainers/source_free_adaptive_teacher.py:150-183`` (threshold_bbox)
* ``daod/engine/trainers/source_free_adaptive_teacher.py:583-603`` (EMA)
* ``daod/engine/trainers/base.py:270-337``           (AdaBN)

Pinning status
--------------
The reference ships no tests and no golden vectors.  Pins:

1. **The reference's own on-disk functions, executed on the CPU** (tests/ref_exec.py imports / compiles them from
   /root/reference on the detectron2 shim, copying nothing): ``fast_rcnn_inference_single_image_with_mcd`` (fast_rcnn.py:88-142),
   ``convert_bbox_scores`` / ``fast_rcnn_inference_single_image_new`` (source_free_fast_rcnn.py:15-36,82-147), ``threshold_bbox``,
   ``adaptive_threshold_bbox``, ``prediction_threshold_bbox``, ``process_pseudo_label``, ``count_label_prediction``,
   ``update_adaptive_threshold`` (source_free_adaptive_teacher.py:150-310), ``_update_teacher_model`` (:583-603 and
   adaptive_teacher.py:339-358), ``reset_bn_stats`` / ``recursive_traversal`` (base.py:318-328),
   ``AdaptiveConfidenceBasedSelfTrainingLoss`` (adaptive_confidence.py:6-33) and the flatten of ``PseudoLabRPN.forward``
   (rpn.py:25-58).  tests/test_oracle_vs_reference_cpu.py demands bit-equality with this module (live, in the build
   container) and tests/golden/ref_exec.npz -- written FROM those functions by tests/golden/make_golden_ref.py -- carries the
   same pin to the GPU box.
2. The native CPU kernels the reference binds to (torchvision 0.26 ``nms`` / ``batched_nms`` / ``roi_align`` / ``roi_pool``,
   ATen): called live by this module and frozen in tests/golden/{nms,roi,dense}.npz; known answers of SURVEY.md App. B in
   tests/golden/kat.json.
3. Still **unpinned against a running detectron2** (none is installable): the pieces of detectron2 0.6 that are NOT on disk in
   the reference -- ``Box2BoxTransform.apply_deltas`` (pinned instead on torchvision's ``BoxCoder.decode_single``, its
   importable twin), ``find_top_rpn_proposals``, ``DefaultAnchorGenerator``, ``ROIPooler`` glue, ``Matcher`` /
   ``subsample_labels`` (pinned on torchvision's detection ``Matcher``).  They follow SURVEY.md Appendix A.

Tie rule: ``torch.sort``/``topk`` on CPU are unspecified on exact ties
(SURVEY.md B-4); the oracle canonicalises to value-descending / index-ascending
(``stable=True``), which is what the CUDA path implements.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
import torchvision

SCALE_CLAMP = math.log(1000.0 / 16)  # d2 box_regression._DEFAULT_SCALE_CLAMP


# --------------------------------------------------------------------------- anchors
def generate_cell_anchors(sizes=(32, 64, 128, 256, 512), aspect_ratios=(0.5, 1.0, 2.0)) -> torch.Tensor:
    """d2 DefaultAnchorGenerator.generate_cell_anchors (SURVEY A-1): size-major, ratio-minor."""
    anchors = []
    for size in sizes:
        area = size ** 2.0
        for ar in aspect_ratios:
            w = math.sqrt(area / ar)
            h = ar * w
            anchors.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
    return torch.tensor(anchors, dtype=torch.float32)


def grid_anchors(grid_size: Tuple[int, int], stride: int, cell_anchors: torch.Tensor, offset: float = 0.0) -> torch.Tensor:
    """d2 DefaultAnchorGenerator._grid_anchors for one level -> (H*W*A, 4), (H, W, A) order."""
    gh, gw = grid_size
    sx = torch.arange(offset * stride, gw * stride, step=stride, dtype=torch.float32)
    sy = torch.arange(offset * stride, gh * stride, step=stride, dtype=torch.float32)
    shift_y, shift_x = torch.meshgrid(sy, sx, indexing="ij")
    shift_x = shift_x.reshape(-1)
    shift_y = shift_y.reshape(-1)
    shifts = torch.stack((shift_x, shift_y, shift_x, shift_y), dim=1)
    return (shifts.view(-1, 1, 4) + cell_anchors.view(1, -1, 4)).reshape(-1, 4)


# --------------------------------------------------------------------------- box coder
def exp_correctly_rounded(x: torch.Tensor) -> torch.Tensor:
    """fp32 exp rounded from the fp64 result.  This is the arithmetic the CUDA path (and
    the C oracle) define; ATen-CPU's ``torch.exp`` (MKL VML / Sleef u10) differs from it
    by one ulp on ~1.1 % of inputs (measured; see DESIGN.md "exp")."""
    return torch.exp(x.double()).float()


def apply_deltas(deltas: torch.Tensor, boxes: torch.Tensor, weights=(1.0, 1.0, 1.0, 1.0),
                 scale_clamp: float = SCALE_CLAMP, exp: Callable = torch.exp) -> torch.Tensor:
    """d2 Box2BoxTransform.apply_deltas (SURVEY A-2).  deltas (R, 4k), boxes (R, 4)."""
    deltas = deltas.float()
    boxes = boxes.to(deltas.dtype)
    widths = boxes[:, 2] - boxes[:, 0]
    heights = boxes[:, 3] - boxes[:, 1]
    ctr_x = boxes[:, 0] + 0.5 * widths
    ctr_y = boxes[:, 1] + 0.5 * heights
    wx, wy, ww, wh = weights
    dx = deltas[:, 0::4] / wx
    dy = deltas[:, 1::4] / wy
    dw = deltas[:, 2::4] / ww
    dh = deltas[:, 3::4] / wh
    dw = torch.clamp(dw, max=scale_clamp)
    dh = torch.clamp(dh, max=scale_clamp)
    pred_ctr_x = dx * widths[:, None] + ctr_x[:, None]
    pred_ctr_y = dy * heights[:, None] + ctr_y[:, None]
    pred_w = exp(dw) * widths[:, None]
    pred_h = exp(dh) * heights[:, None]
    x1 = pred_ctr_x - 0.5 * pred_w
    y1 = pred_ctr_y - 0.5 * pred_h
    x2 = pred_ctr_x + 0.5 * pred_w
    y2 = pred_ctr_y + 0.5 * pred_h
    pred_boxes = torch.stack((x1, y1, x2, y2), dim=-1)
    return pred_boxes.reshape(deltas.shape)


def get_deltas(src_boxes: torch.Tensor, target_boxes: torch.Tensor, weights=(1.0, 1.0, 1.0, 1.0)) -> torch.Tensor:
    """d2 Box2BoxTransform.get_deltas."""
    src_w = src_boxes[:, 2] - src_boxes[:, 0]
    src_h = src_boxes[:, 3] - src_boxes[:, 1]
    src_cx = src_boxes[:, 0] + 0.5 * src_w
    src_cy = src_boxes[:, 1] + 0.5 * src_h
    tw = target_boxes[:, 2] - target_boxes[:, 0]
    th = target_boxes[:, 3] - target_boxes[:, 1]
    tcx = target_boxes[:, 0] + 0.5 * tw
    tcy = target_boxes[:, 1] + 0.5 * th
    wx, wy, ww, wh = weights
    dx = wx * (tcx - src_cx) / src_w
    dy = wy * (tcy - src_cy) / src_h
    dw = ww * torch.log(tw / src_w)
    dh = wh * torch.log(th / src_h)
    return torch.stack((dx, dy, dw, dh), dim=1)


def clip_boxes_(boxes: torch.Tensor, image_size: Tuple[int, int]) -> torch.Tensor:
    """d2 Boxes.clip: box_size is (h, w); x in [0, w], y in [0, h].  In place on (M, 4)."""
    h, w = image_size
    boxes[:, 0].clamp_(min=0, max=w)
    boxes[:, 1].clamp_(min=0, max=h)
    boxes[:, 2].clamp_(min=0, max=w)
    boxes[:, 3].clamp_(min=0, max=h)
    return boxes


def nonempty(boxes: torch.Tensor, threshold: float = 0.0) -> torch.Tensor:
    widths = boxes[:, 2] - boxes[:, 0]
    heights = boxes[:, 3] - boxes[:, 1]
    return (widths > threshold) & (heights > threshold)


# --------------------------------------------------------------------------- NMS
def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """d2 layers.batched_nms (0.6) == torchvision.ops.batched_nms(boxes.float(), ...) (SURVEY A-4).

    The vanilla strategy's final ``sort(descending=True)`` is *unstable* on CPU (probed), so the
    oracle re-sorts the kept set stably (score desc, index asc) -- identical on tie-free scores.
    """
    keep = torchvision.ops.batched_nms(boxes.float(), scores, idxs, iou_threshold)
    if keep.numel() > 1:
        ks = scores[keep]
        # canonical tie order
        order = torch.argsort(keep, stable=True)
        keep = keep[order]
        ks = scores[keep]
        keep = keep[torch.sort(ks, descending=True, stable=True)[1]]
    return keep


# --------------------------------------------------------------------------- RPN
def rpn_flatten_head_outputs(objectness: List[torch.Tensor], deltas: List[torch.Tensor], box_dim: int = 4):
    """reference rpn.py:27-41: (N,A,H,W)->(N,HWA); (N,A*B,H,W)->(N,HWA,B)."""
    logits = [s.permute(0, 2, 3, 1).flatten(1) for s in objectness]
    dl = [x.view(x.shape[0], -1, box_dim, x.shape[-2], x.shape[-1]).permute(0, 3, 4, 1, 2).flatten(1, -2)
          for x in deltas]
    return logits, dl


def decode_proposals(anchors: List[torch.Tensor], pred_anchor_deltas: List[torch.Tensor],
                     weights=(1.0, 1.0, 1.0, 1.0), exp: Callable = torch.exp) -> List[torch.Tensor]:
    """d2 RPN._decode_proposals."""
    N = pred_anchor_deltas[0].shape[0]
    out = []
    for anchors_i, d_i in zip(anchors, pred_anchor_deltas):
        B = anchors_i.size(1)
        d_i = d_i.reshape(-1, B)
        a_i = anchors_i.unsqueeze(0).expand(N, -1, -1).reshape(-1, B)
        out.append(apply_deltas(d_i, a_i, weights, exp=exp).view(N, -1, B))
    return out


def find_top_rpn_proposals(proposals: List[torch.Tensor], pred_objectness_logits: List[torch.Tensor],
                           image_sizes: Sequence[Tuple[int, int]], nms_thresh: float, pre_nms_topk: int,
                           post_nms_topk: int, min_box_size: float, training: bool):
    """d2 proposal_utils.find_top_rpn_proposals (0.6) (SURVEY A-3).

    Returns per image dict(proposal_boxes (k,4), objectness_logits (k,), src_index (k,) int64 --
    the flat anchor index of each kept proposal, level-concatenated, used by the parity tests).
    """
    num_images = len(image_sizes)
    topk_scores, topk_proposals, level_ids, topk_src = [], [], [], []
    batch_idx = torch.arange(num_images)
    base = 0
    for level_id, (proposals_i, logits_i) in enumerate(zip(proposals, pred_objectness_logits)):
        hwa = logits_i.shape[1]
        num_proposals_i = min(hwa, pre_nms_topk)
        logits_s, idx = logits_i.sort(descending=True, dim=1, stable=True)   # canonical tie rule
        topk_scores_i = logits_s.narrow(1, 0, num_proposals_i)
        topk_idx = idx.narrow(1, 0, num_proposals_i)
        topk_proposals.append(proposals_i[batch_idx[:, None], topk_idx])
        topk_scores.append(topk_scores_i)
        topk_src.append(topk_idx + base)
        level_ids.append(torch.full((num_proposals_i,), level_id, dtype=torch.int64))
        base += hwa
    topk_scores = torch.cat(topk_scores, dim=1)
    topk_proposals = torch.cat(topk_proposals, dim=1)
    topk_src = torch.cat(topk_src, dim=1)
    level_ids = torch.cat(level_ids, dim=0)

    results = []
    for n, image_size in enumerate(image_sizes):
        boxes = topk_proposals[n].clone()
        scores_per_img = topk_scores[n]
        src = topk_src[n]
        lvl = level_ids
        valid_mask = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores_per_img)
        if not valid_mask.all():
            if training:
                raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")
            boxes, scores_per_img, lvl, src = boxes[valid_mask], scores_per_img[valid_mask], lvl[valid_mask], src[valid_mask]
        clip_boxes_(boxes, image_size)
        keep = nonempty(boxes, threshold=min_box_size)
        if keep.sum().item() != len(boxes):
            boxes, scores_per_img, lvl, src = boxes[keep], scores_per_img[keep], lvl[keep], src[keep]
        keep = batched_nms(boxes, scores_per_img, lvl, nms_thresh)
        keep = keep[:post_nms_topk]
        results.append(dict(image_size=tuple(image_size), proposal_boxes=boxes[keep],
                            objectness_logits=scores_per_img[keep], src_index=src[keep]))
    return results


def rpn_predict_proposals(anchors: List[torch.Tensor], pred_objectness_logits: List[torch.Tensor],
                          pred_anchor_deltas: List[torch.Tensor], image_sizes, nms_thresh=0.7,
                          pre_nms_topk=12000, post_nms_topk=2000, min_box_size=0.0, training=True,
                          weights=(1.0, 1.0, 1.0, 1.0), exp: Callable = torch.exp):
    """d2 RPN.predict_proposals, called from reference rpn.py:54-56."""
    with torch.no_grad():
        pred = decode_proposals(anchors, pred_anchor_deltas, weights, exp=exp)
        return find_top_rpn_proposals(pred, pred_objectness_logits, image_sizes, nms_thresh,
                                      pre_nms_topk, post_nms_topk, min_box_size, training)


# --------------------------------------------------------------------------- ROI pooling
def convert_boxes_to_pooler_format(box_lists: List[torch.Tensor]) -> torch.Tensor:
    """d2 poolers.convert_boxes_to_pooler_format: (R,5) [batch_idx, x1, y1, x2, y2]."""
    parts = []
    for i, b in enumerate(box_lists):
        parts.append(torch.cat([torch.full((len(b), 1), i, dtype=b.dtype), b], dim=1))
    return torch.cat(parts, dim=0) if parts else torch.zeros((0, 5))


def roi_pooler(features: torch.Tensor, box_lists: List[torch.Tensor], output_size=7, scale=1.0 / 32,
               sampling_ratio=0, pooler_type="ROIAlignV2") -> torch.Tensor:
    """d2 ROIPooler.forward for the single-level case used by every shipped config
    (built at reference ...roi_heads.py:42-47, called :117) (SURVEY A-5)."""
    rois = convert_boxes_to_pooler_format(box_lists).to(features.dtype)
    osz = (output_size, output_size) if isinstance(output_size, int) else tuple(output_size)
    if pooler_type == "ROIAlignV2":
        return torchvision.ops.roi_align(features, rois, osz, scale, sampling_ratio, aligned=True)
    if pooler_type == "ROIAlign":
        return torchvision.ops.roi_align(features, rois, osz, scale, sampling_ratio, aligned=False)
    if pooler_type == "ROIPool":
        return torchvision.ops.roi_pool(features, rois, osz, scale)
    raise ValueError(pooler_type)


# --------------------------------------------------------------------------- Fast R-CNN outputs
def predict_boxes(proposal_deltas: torch.Tensor, proposal_boxes: List[torch.Tensor],
                  weights=(10.0, 10.0, 5.0, 5.0), exp: Callable = torch.exp) -> List[torch.Tensor]:
    """d2 FastRCNNOutputLayers.predict_boxes (SURVEY A-6)."""
    n_per = [len(p) for p in proposal_boxes]
    pb = torch.cat(proposal_boxes, dim=0)
    return list(apply_deltas(proposal_deltas, pb, weights, exp=exp).split(n_per))


def predict_probs(scores: torch.Tensor, n_per: List[int]) -> List[torch.Tensor]:
    """d2 FastRCNNOutputLayers.predict_probs."""
    return list(F.softmax(scores, dim=-1).split(n_per, dim=0))


def fast_rcnn_inference_single_image(boxes, scores, image_shape, score_thresh, nms_thresh, topk_per_image):
    """Restates reference daod/modeling/roi_heads/fast_rcnn.py:108-142 (identical to d2's)."""
    valid_mask = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores).all(dim=1)
    row_ids = torch.arange(boxes.shape[0])
    if not valid_mask.all():
        boxes = boxes[valid_mask]
        scores = scores[valid_mask]
        row_ids = row_ids[valid_mask]
    scores = scores[:, :-1]
    num_bbox_reg_classes = boxes.shape[1] // 4
    boxes = clip_boxes_(boxes.reshape(-1, 4).clone(), image_shape).view(-1, num_bbox_reg_classes, 4)
    filter_mask = scores > score_thresh
    filter_inds = filter_mask.nonzero()
    if num_bbox_reg_classes == 1:
        boxes = boxes[filter_inds[:, 0], 0]
    else:
        boxes = boxes[filter_mask]
    scores = scores[filter_mask]
    keep = batched_nms(boxes, scores, filter_inds[:, 1], nms_thresh)
    if topk_per_image >= 0:
        keep = keep[:topk_per_image]
    boxes, scores, filter_inds = boxes[keep], scores[keep], filter_inds[keep]
    return dict(image_size=tuple(image_shape), pred_boxes=boxes, scores=scores,
                pred_classes=filter_inds[:, 1], kept_rows=row_ids[filter_inds[:, 0]])


def fast_rcnn_inference(boxes: List[torch.Tensor], scores: List[torch.Tensor], image_shapes,
                        score_thresh=0.05, nms_thresh=0.5, topk_per_image=100):
    return [fast_rcnn_inference_single_image(b, s, sh, score_thresh, nms_thresh, topk_per_image)
            for s, b, sh in zip(scores, boxes, image_shapes)]


def convert_bbox_scores_single_image(boxes, scores, image_shape):
    """Restates reference source_free_fast_rcnn.py:100-147: decode+clip, ``scores > 0``, no NMS."""
    valid_mask = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores).all(dim=1)
    row_ids = torch.arange(boxes.shape[0])
    if not valid_mask.all():
        boxes, scores, row_ids = boxes[valid_mask], scores[valid_mask], row_ids[valid_mask]
    scores = scores[:, :-1]
    k = boxes.shape[1] // 4
    boxes = clip_boxes_(boxes.reshape(-1, 4).clone(), image_shape).view(-1, k, 4)
    filter_mask = scores > 0
    filter_inds = filter_mask.nonzero()
    boxes = boxes[filter_inds[:, 0], 0] if k == 1 else boxes[filter_mask]
    scores = scores[filter_mask]
    return dict(image_size=tuple(image_shape), pred_boxes=boxes, scores=scores,
                pred_classes=filter_inds[:, 1], kept_rows=row_ids[filter_inds[:, 0]])


def box_predictor_inference(cls_logits: torch.Tensor, proposal_deltas: torch.Tensor,
                            proposal_boxes: List[torch.Tensor], image_shapes, score_thresh=0.05,
                            nms_thresh=0.5, topk_per_image=100, weights=(10.0, 10.0, 5.0, 5.0),
                            exp: Callable = torch.exp, softmax: Optional[Callable] = None):
    """d2 FastRCNNOutputLayers.inference (called at reference ...roi_heads.py:161)."""
    n_per = [len(p) for p in proposal_boxes]
    bxs = predict_boxes(proposal_deltas, proposal_boxes, weights, exp=exp)
    if softmax is None:
        prs = predict_probs(cls_logits, n_per)
    else:
        prs = list(softmax(cls_logits).split(n_per, dim=0))
    return fast_rcnn_inference(bxs, prs, image_shapes, score_thresh, nms_thresh, topk_per_image)


def softmax_defined(x: torch.Tensor) -> torch.Tensor:
    """The softmax arithmetic the CUDA path / C oracle define (DESIGN.md "softmax"):
    m = max; e_k = fp32(exp_fp64(x_k - m)); s = sequential fp32 sum k=0..K; p_k = e_k / s."""
    x = x.float()
    m = x.max(dim=-1, keepdim=True).values
    e = exp_correctly_rounded(x - m)
    s = torch.zeros(x.shape[:-1] + (1,), dtype=torch.float32)
    for k in range(x.shape[-1]):
        s = s + e[..., k:k + 1]
    return e / s


# --------------------------------------------------------------------------- pseudo-label filter
def threshold_bbox(inst: dict, thres: float = 0.7, proposal_type: str = "roih") -> dict:
    """Restates reference source_free_adaptive_teacher.py:150-183."""
    if proposal_type == "rpn":
        valid_map = inst["objectness_logits"] > thres
        return dict(image_size=inst["image_size"], gt_boxes=inst["proposal_boxes"][valid_map, :],
                    objectness_logits=inst["objectness_logits"][valid_map])
    elif proposal_type == "roih":
        valid_map = inst["scores"] > thres
        return dict(image_size=inst["image_size"], gt_boxes=inst["pred_boxes"][valid_map, :],
                    gt_classes=inst["pred_classes"][valid_map], scores=inst["scores"][valid_map])
    raise ValueError(proposal_type)


def process_pseudo_label(insts: List[dict], cur_threshold: float, proposal_type: str, method="thresholding"):
    """Restates reference source_free_adaptive_teacher.py:256-280 (thresholding branch)."""
    if method != "thresholding":
        raise ValueError("Unkown pseudo label boxes methods")
    out, n = [], 0.0
    for inst in insts:
        r = threshold_bbox(inst, thres=cur_threshold, proposal_type=proposal_type)
        n += len(r["gt_boxes"])
        out.append(r)
    return out, n / len(insts)


# --------------------------------------------------------------------------- EMA
@torch.no_grad()
def update_teacher_model(student_sd: Dict[str, torch.Tensor], teacher_sd: Dict[str, torch.Tensor],
                         keep_rate: float = 0.9996, ddp_prefix: bool = False) -> "OrderedDict[str, torch.Tensor]":
    """Restates reference source_free_adaptive_teacher.py:583-603; returns the dict that the
    reference then feeds to ``load_state_dict`` (which copy_()s into the teacher's dtypes)."""
    if ddp_prefix:
        student_sd = {k[7:]: v for k, v in student_sd.items()}
    new = OrderedDict()
    for key, value in teacher_sd.items():
        if key in student_sd.keys():
            new[key] = student_sd[key] * (1 - keep_rate) + value * keep_rate
        else:
            raise Exception("{} is not found in student model".format(key))
    return new


@torch.no_grad()
def load_state_dict_like(teacher_sd: Dict[str, torch.Tensor], new_sd: Dict[str, torch.Tensor]) -> None:
    """What nn.Module.load_state_dict does with the EMA result: ``param.copy_(input_param)``."""
    for k, v in new_sd.items():
        teacher_sd[k].copy_(v)


# --------------------------------------------------------------------------- AdaBN
def reset_bn_stats(module: torch.nn.Module) -> None:
    """Restates reference base.py:318-323 (stats become non-trainable Parameters)."""
    if isinstance(module, torch.nn.BatchNorm2d):
        module.running_mean = torch.nn.Parameter(torch.zeros_like(module.running_mean), requires_grad=False)
        module.running_var = torch.nn.Parameter(torch.ones_like(module.running_var), requires_grad=False)


def recursive_traversal(module: torch.nn.Module) -> None:
    """Restates reference base.py:325-328."""
    for child in module.children():
        reset_bn_stats(child)
        recursive_traversal(child)


@torch.no_grad()
def bn_train_forward(x, running_mean, running_var, weight, bias, momentum=0.1, eps=1e-5):
    """What nn.BatchNorm2d does in train() mode under no_grad (SURVEY A-7); updates running stats in place."""
    return F.batch_norm(x, running_mean, running_var, weight, bias, True, momentum, eps)


@torch.no_grad()
def adabn_recompute(model: torch.nn.Module, batches, max_iters: int = 1400) -> int:
    """Restates the statistic-relevant part of reference base.py:270-337: reset (twice), then
    train-mode forwards under no_grad for at most ``max_iters``+1 batches (the reference breaks when
    ``i > 1400`` *after* running the batch)."""
    recursive_traversal(model)
    recursive_traversal(model)
    model.train()
    i = 0
    for data in batches:
        i += 1
        model(data)
        if i > max_iters:
            break
    return i


# --------------------------------------------------------------------------- student-side labelling / losses
def pairwise_iou(boxes1: torch.Tensor, boxes2: torch.Tensor) -> torch.Tensor:
    """d2 structures.pairwise_iou on raw (M,4)/(N,4) tensors."""
    a1 = (boxes1[:, 2] - boxes1[:, 0]) * (boxes1[:, 3] - boxes1[:, 1])
    a2 = (boxes2[:, 2] - boxes2[:, 0]) * (boxes2[:, 3] - boxes2[:, 1])
    wh = (torch.min(boxes1[:, None, 2:], boxes2[:, 2:]) - torch.max(boxes1[:, None, :2], boxes2[:, :2])).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    return torch.where(inter > 0, inter / (a1[:, None] + a2 - inter), torch.zeros(1))


def matcher(mqm: torch.Tensor, thresholds: Sequence[float], labels: Sequence[int], allow_low_quality: bool):
    """d2 Matcher.__call__ (A-9): RPN uses ([0.3, 0.7], [0, -1, 1], True), ROI heads ([0.5], [0, 1], False)."""
    n = mqm.shape[1]
    if mqm.numel() == 0:
        return torch.zeros(n, dtype=torch.int64), torch.full((n,), labels[0], dtype=torch.int8)
    vals, matches = mqm.max(dim=0)
    out = torch.ones(n, dtype=torch.int8)
    th = [-float("inf")] + list(thresholds) + [float("inf")]
    for l, lo, hi in zip(labels, th[:-1], th[1:]):
        out[(vals >= lo) & (vals < hi)] = l
    if allow_low_quality:
        best, _ = mqm.max(dim=1)
        out[(mqm == best[:, None]).nonzero()[:, 1]] = 1
    return matches, out


def subsample_labels(labels, num_samples, positive_fraction, bg_label, randperm=torch.randperm):
    positive = ((labels != -1) & (labels != bg_label)).nonzero().flatten()
    negative = (labels == bg_label).nonzero().flatten()
    num_pos = min(positive.numel(), int(num_samples * positive_fraction))
    num_neg = min(negative.numel(), num_samples - num_pos)
    return positive[randperm(positive.numel())[:num_pos]], negative[randperm(negative.numel())[:num_neg]]


def rpn_label_and_sample_anchors(anchors: torch.Tensor, gt_boxes: List[torch.Tensor], batch_size_per_image=256,
                                 positive_fraction=0.5, randperm=torch.randperm):
    """d2 RPN.label_and_sample_anchors (called at reference rpn.py:45)."""
    labels_out, boxes_out = [], []
    for gt in gt_boxes:
        idx, lab = matcher(pairwise_iou(gt, anchors), [0.3, 0.7], [0, -1, 1], True)
        pos, neg = subsample_labels(lab, batch_size_per_image, positive_fraction, 0, randperm)
        lab = torch.full_like(lab, -1)
        lab[pos] = 1
        lab[neg] = 0
        labels_out.append(lab)
        boxes_out.append(gt[idx] if len(gt) else torch.zeros_like(anchors))
    return labels_out, boxes_out


def rpn_losses(anchors, logits, gt_labels, deltas, gt_boxes, batch_size_per_image=256, weights=(1.0, 1.0, 1.0, 1.0)):
    """d2 RPN.losses (smooth_l1 with beta 0 == L1), logits (N, R), deltas (N, R, 4)."""
    lab = torch.stack(gt_labels)
    pos = lab == 1
    tgt = torch.stack([get_deltas(anchors, g, weights) for g in gt_boxes])
    loc = F.l1_loss(deltas[pos], tgt[pos], reduction="sum")
    valid = lab >= 0
    cls = F.binary_cross_entropy_with_logits(logits[valid], lab[valid].float(), reduction="sum")
    norm = batch_size_per_image * len(gt_labels)
    return {"loss_rpn_cls": cls / norm, "loss_rpn_loc": loc / norm}


def label_and_sample_proposals(proposals: List[dict], targets: List[dict], num_classes=8, batch_size_per_image=512,
                               positive_fraction=0.25, append_gt=True, randperm=torch.randperm):
    """reference source_free_adaptive_teacher_roi_heads.py:165-215 on dicts(proposal_boxes, objectness_logits) / dicts(gt_boxes, gt_classes)."""
    out = []
    gt_logit = math.log((1.0 - 1e-10) / (1 - (1.0 - 1e-10)))
    for p, t in zip(proposals, targets):
        boxes, logits = p["proposal_boxes"], p["objectness_logits"]
        if append_gt:
            boxes = torch.cat([boxes, t["gt_boxes"]])
            logits = torch.cat([logits, gt_logit * torch.ones(len(t["gt_boxes"]))])
        idx, lab = matcher(pairwise_iou(t["gt_boxes"], boxes), [0.5], [0, 1], False)
        if t["gt_classes"].numel() > 0:
            cls = t["gt_classes"][idx].clone()
            cls[lab == 0] = num_classes
            cls[lab == -1] = -1
        else:
            cls = torch.zeros_like(idx) + num_classes
        fg, bg = subsample_labels(cls, batch_size_per_image, positive_fraction, num_classes, randperm)
        sel = torch.cat([fg, bg])
        gtb = t["gt_boxes"][idx[sel]] if len(t["gt_boxes"]) else torch.zeros((len(sel), 4))
        out.append(dict(image_size=p.get("image_size"), proposal_boxes=boxes[sel], objectness_logits=logits[sel], gt_classes=cls[sel], gt_boxes=gtb))
    return out


def fast_rcnn_losses(scores, proposal_deltas, proposals: List[dict], num_classes=8, weights=(10.0, 10.0, 5.0, 5.0)):
    """d2 FastRCNNOutputLayers.losses: mean CE + class-specific L1 (smooth_l1 beta 0) over the foreground / #proposals."""
    gt_classes = torch.cat([p["gt_classes"] for p in proposals])
    pb = torch.cat([p["proposal_boxes"] for p in proposals])
    gb = torch.cat([p["gt_boxes"] for p in proposals])
    loss_cls = F.cross_entropy(scores, gt_classes, reduction="mean")
    fg = ((gt_classes >= 0) & (gt_classes < num_classes)).nonzero().flatten()
    pred = proposal_deltas.view(-1, num_classes, 4)[fg, gt_classes[fg]]
    loss_box = F.l1_loss(pred, get_deltas(pb[fg], gb[fg], weights), reduction="sum") / max(gt_classes.numel(), 1.0)
    return {"loss_cls": loss_cls, "loss_box_reg": loss_box}


# --------------------------------------------------------------------------- adaptive per-class threshold
def adaptive_confidence_mask(confidence: torch.Tensor, pseudo_labels: torch.Tensor, threshold: float, classwise_acc: torch.Tensor):
    """reference daod/modeling/adaptive_thresh/adaptive_confidence.py:21-33 ("convex" map)."""
    return (confidence >= threshold * (classwise_acc[pseudo_labels] / (2. - classwise_acc[pseudo_labels]))).float()


def adaptive_threshold_bbox(inst: dict, threshold: float, classwise_acc: torch.Tensor, as_gt: bool = True) -> dict:
    """reference source_free_adaptive_teacher.py:185-228 (roih branch; as_gt=False: prediction_threshold_bbox :230-254)."""
    mask = adaptive_confidence_mask(inst["scores"], inst["pred_classes"], threshold, classwise_acc)
    valid = (mask == 1).nonzero().flatten()
    kb, kc = ("gt_boxes", "gt_classes") if as_gt else ("pred_boxes", "pred_classes")
    return {"image_size": inst.get("image_size"), kb: inst["pred_boxes"][valid, :], kc: inst["pred_classes"][valid], "scores": inst["scores"][valid]}


def count_label_prediction(preds: List[dict], num_classes: int, bbox_threshold: float) -> torch.Tensor:
    """reference source_free_adaptive_teacher.py:282-296."""
    reserve = torch.zeros(num_classes)
    for p in preds:
        reserve += p["pred_classes"][p["scores"] > bbox_threshold].bincount(minlength=num_classes)
    return reserve


def update_adaptive_threshold(reserve_matrix: torch.Tensor) -> torch.Tensor:
    """reference source_free_adaptive_teacher.py:298-310 -> new classwise_acc (classes 0 and 2 hard-coded, as in the reference)."""
    counter = reserve_matrix.sum(dim=0)
    counter[0] = 0
    counter[2] = 0
    acc = counter / max(counter.max(), 1)
    acc[0] = 1
    acc[2] = 1
    return acc
