"""Pseudo-label export: the on-disk format that links AdaBN inference to the fixed-pseudo-label
training stage of the reference.

* ``detector_postprocess`` -- detectron2.modeling.postprocessing.detector_postprocess (boxes rescaled to the output
  resolution, clipped, empty boxes dropped) as applied by ``GeneralizedRCNN.inference``;
* ``instances_to_coco_json`` -- reference daod/evaluation/sim_cocoevaluator.py:65-123 (box branch: XYXY -> XYWH, score,
  category id), plus ``batch_to_coco_json`` that serialises a whole fused ``DetectionBatch`` with ONE device->host copy;
* ``prediction_to_gt`` -- reference cityscapes-to-coco-conversion/prediction_to_gt.py:21-45: detections with
  ``score >= 0.7`` become the ``annotations`` of the target-domain training json (ids from 1).
Host-side formatting only; the arithmetic that matters (which detections exist) happened in the fused kernels.
"""
from __future__ import annotations

import copy
from typing import Dict, List, Optional, Sequence

import torch

from ..structures import Instances


def detector_postprocess(results: Instances, output_height: int, output_width: int) -> Instances:
    if isinstance(output_width, torch.Tensor):
        output_width, output_height = float(output_width), float(output_height)
    new_size = (int(output_height), int(output_width))
    scale_x, scale_y = output_width / results.image_size[1], output_height / results.image_size[0]
    out = Instances(new_size, **results.get_fields())
    if out.has("pred_boxes"):
        output_boxes = out.pred_boxes.clone()
        out.pred_boxes = output_boxes
    elif out.has("proposal_boxes"):
        output_boxes = out.proposal_boxes.clone()
        out.proposal_boxes = output_boxes
    else:
        return out
    output_boxes.scale(scale_x, scale_y)
    output_boxes.clip(out.image_size)
    return out[output_boxes.nonempty()]


def _xyxy_to_xywh(boxes: torch.Tensor) -> torch.Tensor:
    b = boxes.clone()
    b[:, 2] -= b[:, 0]
    b[:, 3] -= b[:, 1]
    return b


def instances_to_coco_json(instances: Instances, img_id: int, id_map: Optional[Dict[int, int]] = None) -> List[dict]:
    num_instance = len(instances)
    if num_instance == 0:
        return []
    boxes = _xyxy_to_xywh(instances.pred_boxes.tensor.detach().cpu()).tolist()
    scores = instances.scores.tolist()
    classes = instances.pred_classes.tolist()
    return [{"image_id": img_id, "category_id": id_map[classes[k]] if id_map else classes[k], "bbox": boxes[k], "score": scores[k]}
            for k in range(num_instance)]


def batch_to_coco_json(batch, image_ids: Sequence[int], id_map: Optional[Dict[int, int]] = None) -> List[dict]:
    """All detections of a fused post-processing batch (``FastRCNNOutputLayers.inference_batch``) in COCO result format,
    image by image in score order, with one device->host transfer for the whole batch."""
    det, _ = batch.host_counts()
    packed = torch.cat([batch.boxes, batch.scores.unsqueeze(-1), batch.classes.to(batch.boxes.dtype).unsqueeze(-1)], dim=-1).cpu()
    out = []
    for i, img_id in enumerate(image_ids):
        k = det[i]
        if k == 0:
            continue
        rows = packed[i, :k]
        xywh = _xyxy_to_xywh(rows[:, :4]).tolist()
        for j in range(k):
            c = int(rows[j, 5])
            out.append({"image_id": img_id, "category_id": id_map[c] if id_map else c, "bbox": xywh[j], "score": float(rows[j, 4])})
    return out


def prediction_to_gt(predictions: List[dict], dataset: dict, score_thresh: float = 0.7) -> dict:
    """Returns a copy of ``dataset`` (a COCO-format dict) whose ``annotations`` are the predictions with
    ``score >= score_thresh`` ({image_id, bbox, category_id, id}, ids from 1, input order)."""
    ground_truths = []
    idd = 1
    for prediction in predictions:
        if prediction["score"] < score_thresh:
            continue
        ground_truths.append({"image_id": prediction["image_id"], "bbox": prediction["bbox"],
                              "category_id": prediction["category_id"], "id": idd})
        idd += 1
    out = copy.copy(dataset)
    out["annotations"] = ground_truths
    return out


def load_pseudo_label_annotations(dataset: dict, image_sizes: Optional[Dict[int, tuple]] = None, device=None,
                                  filter_empty: bool = True) -> Dict[int, Instances]:
    """What the fixed-pseudo-label stage sees when it trains on the json written above: the reference registers that file with
    ``register_coco_instances`` (reference daod/data/datasets.py:46-63), i.e. detectron2's ``load_coco_json`` +
    ``annotations_to_instances`` -- XYWH_ABS boxes become XYXY ``gt_boxes``, COCO category ids are mapped to contiguous
    ``gt_classes`` in sorted-id order.  Returns {image_id: Instances(gt_boxes, gt_classes)} (images without annotations get an
    empty Instances when ``dataset["images"]`` lists them).  ``image_sizes`` overrides / supplies (height, width) per image id.
    ``filter_empty``: drop boxes without area, as detectron2's DatasetMapper does (``utils.filter_empty_instances``) -- a
    zero-width pseudo-label (a detection clipped at the image border) would otherwise put log(0) into the regression targets."""
    from ..structures import Boxes
    cats = sorted(c["id"] for c in dataset.get("categories", []))
    if not cats:
        cats = sorted({a["category_id"] for a in dataset["annotations"]})
    cat_to_contiguous = {c: i for i, c in enumerate(cats)}
    sizes = {im["id"]: (int(im.get("height", 0)), int(im.get("width", 0))) for im in dataset.get("images", [])}
    if image_sizes:
        sizes.update(image_sizes)
    per_image: Dict[int, list] = {i: [] for i in sizes}
    for a in dataset["annotations"]:
        per_image.setdefault(a["image_id"], []).append(a)
    out = {}
    for img_id, anns in per_image.items():
        inst = Instances(sizes.get(img_id, (0, 0)))
        b = torch.tensor([a["bbox"] for a in anns], dtype=torch.float32).reshape(-1, 4)
        b[:, 2] += b[:, 0]
        b[:, 3] += b[:, 1]
        c = torch.tensor([cat_to_contiguous[a["category_id"]] for a in anns], dtype=torch.int64)
        if device is not None:
            b, c = b.to(device), c.to(device)
        if filter_empty and len(anns):
            keep = ((b[:, 2] - b[:, 0]) > 1e-5) & ((b[:, 3] - b[:, 1]) > 1e-5)
            b, c = b[keep], c[keep]
        inst.gt_boxes = Boxes(b)
        inst.gt_classes = c
        out[img_id] = inst
    return out
