"""CPU restatement of the reference's teacher pseudo-labelling step, end to end.  TEST INFRASTRUCTURE ONLY
(imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by the product).

One call = what reference daod/engine/trainers/source_free_adaptive_teacher.py:385-390 + :256-280 execute for a batch of
weakly augmented target images on detectron2's CPU path:
  preprocess (reference daod/modeling/meta_arch/source_free_adaptive_teacher_rcnn.py:92-104)
  -> VGG16-BN backbone in train() mode under no_grad (reference daod/modeling/meta_arch/vgg.py:10-98; BN uses batch
     statistics and updates its running statistics)
  -> PseudoLabRPN (reference daod/modeling/proposal_generator/rpn.py:16-58) with the TRAIN top-k pair
  -> ROIPooler/ROIAlignV2 -> FastRCNNConvFCHead -> FastRCNNOutputLayers.inference
     (reference daod/modeling/roi_heads/source_free_adaptive_teacher_roi_heads.py:108-163)
  -> threshold_bbox (reference source_free_adaptive_teacher.py:150-183)
and, optionally, the EMA teacher update (:583-603).  Dense layers run on ATen CPU (conv2d / batch_norm / linear), the
detection glue on ``oracle.d2_cpu`` (torchvision CPU nms / roi_align).  It consumes a plain state_dict with the
reference's key names, so it shares no code with the product package.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

from . import d2_cpu as o

VGG16 = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"]
_STAGE_CUTS = [(0, 7), (7, 14), (14, 24), (24, 34), (34, 44)]  # reference vgg.py:70-74 (module indices with BN)


def _vgg_layer_table() -> List[Tuple[str, str]]:
    """[(kind, state_dict prefix)] for the 44 modules of make_layers(vgg16, batch_norm=True), keyed as the reference's stages."""
    mods = []
    for v in VGG16:
        mods += ["pool"] if v == "M" else ["conv", "bn", "relu"]
    table = []
    for s, (a, b) in enumerate(_STAGE_CUTS):
        for local, kind in enumerate(mods[a:b]):
            table.append((kind, f"backbone.vgg{s}.{local}"))
    return table


@torch.no_grad()
def backbone_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, training: bool = True, momentum: float = 0.1, eps: float = 1e-5):
    for kind, pre in _vgg_layer_table():
        if kind == "conv":
            x = F.conv2d(x, sd[pre + ".weight"], sd[pre + ".bias"], padding=1)
        elif kind == "bn":
            if training:
                sd[pre + ".num_batches_tracked"] += 1
            x = F.batch_norm(x, sd[pre + ".running_mean"], sd[pre + ".running_var"], sd[pre + ".weight"], sd[pre + ".bias"],
                             training, momentum, eps)
        elif kind == "relu":
            x = F.relu_(x)
        else:
            x = F.max_pool2d(x, 2, 2)
    return x  # vgg4


@torch.no_grad()
def resnet_c4_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, training: bool = True, momentum: float = 0.1, eps: float = 1e-5,
                      blocks=(3, 4, 23)):
    """detectron2 ResNet-C4 (BasicStem + res2..res4 BottleneckBlocks, STRIDE_IN_1X1) from a state_dict with detectron2's keys
    (``backbone.stem.conv1.*``, ``backbone.res{s}.{i}.{shortcut,conv1,conv2,conv3}.*``) -- the backbone of reference
    configs/r101_c4_cs_foggy_adaptive_teacher_source_free.yaml.  A norm layer WITHOUT ``num_batches_tracked`` in the state_dict
    is a FrozenBatchNorm2d (detectron2 ``FREEZE_AT``) and always uses its stored statistics; the others are nn.BatchNorm2d in
    train() mode when ``training`` (batch statistics, running-stat update)."""
    def norm(t, pre):
        frozen = (pre + ".num_batches_tracked") not in sd
        if training and not frozen:
            sd[pre + ".num_batches_tracked"] += 1
        return F.batch_norm(t, sd[pre + ".running_mean"], sd[pre + ".running_var"], sd[pre + ".weight"], sd[pre + ".bias"],
                            training and not frozen, momentum, eps)

    def conv(t, pre, stride=1, padding=0):
        return norm(F.conv2d(t, sd[pre + ".weight"], None, stride, padding), pre + ".norm")

    x = F.relu_(conv(x, "backbone.stem.conv1", 2, 3))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for s, n in zip((2, 3, 4), blocks):
        for i in range(n):
            pre = f"backbone.res{s}.{i}"
            stride = 2 if (i == 0 and s > 2) else 1
            out = F.relu_(conv(x, pre + ".conv1", stride))          # STRIDE_IN_1X1: the stride sits on the first 1x1
            out = F.relu_(conv(out, pre + ".conv2", 1, 1))
            out = conv(out, pre + ".conv3")
            shortcut = conv(x, pre + ".shortcut", stride) if (pre + ".shortcut.weight") in sd else x
            out += shortcut
            x = F.relu_(out)
    return x  # res4


@torch.no_grad()
def teacher_pseudo_label(sd: Dict[str, torch.Tensor], images_u8: torch.Tensor, *, training: bool = True, num_classes: int = 8,
                         bbox_threshold: float = 0.8, pixel_mean=(103.530, 116.280, 123.675), pixel_std=(1.0, 1.0, 1.0),
                         sizes=(32, 64, 128, 256, 512), ratios=(0.5, 1.0, 2.0), stride: int = 32,
                         pre_nms_topk=(12000, 6000), post_nms_topk=(2000, 1000)):
    """images_u8: (N, 3, H, W) uint8/float CPU tensor.  Returns (proposals_rpn, proposals_roih, pseudo_labels) as lists of dicts."""
    N, _, H, W = images_u8.shape
    image_sizes = [(H, W)] * N
    x = (images_u8.float() - torch.tensor(pixel_mean).view(1, 3, 1, 1)) / torch.tensor(pixel_std).view(1, 3, 1, 1)
    if "backbone.stem.conv1.weight" in sd:      # ResNet-C4 (R101-C4 configs): stride 16, anchors 64..512
        feat = resnet_c4_forward(sd, x, training)
    else:
        feat = backbone_forward(sd, x, training)
    # RPN head + flatten (reference rpn.py:27-41)
    t = F.relu(F.conv2d(feat, sd["proposal_generator.rpn_head.conv.weight"], sd["proposal_generator.rpn_head.conv.bias"], padding=1))
    obj = F.conv2d(t, sd["proposal_generator.rpn_head.objectness_logits.weight"], sd["proposal_generator.rpn_head.objectness_logits.bias"])
    dl = F.conv2d(t, sd["proposal_generator.rpn_head.anchor_deltas.weight"], sd["proposal_generator.rpn_head.anchor_deltas.bias"])
    logits, deltas = o.rpn_flatten_head_outputs([obj], [dl])
    anchors = o.grid_anchors(tuple(feat.shape[-2:]), stride, o.generate_cell_anchors(sizes, ratios))
    k = 0 if training else 1
    props = o.rpn_predict_proposals([anchors], logits, deltas, image_sizes, 0.7, pre_nms_topk[k], post_nms_topk[k], 0.0, training)
    # ROI heads (inference path)
    boxes = [p["proposal_boxes"] for p in props]
    pooled = o.roi_pooler(feat, boxes, 7, 1.0 / stride, 0, "ROIAlignV2")
    h = F.relu(F.linear(pooled.flatten(1), sd["roi_heads.box_head.fc1.weight"], sd["roi_heads.box_head.fc1.bias"]))
    h = F.relu(F.linear(h, sd["roi_heads.box_head.fc2.weight"], sd["roi_heads.box_head.fc2.bias"]))
    cls = F.linear(h, sd["roi_heads.box_predictor.cls_score.weight"], sd["roi_heads.box_predictor.cls_score.bias"])
    reg = F.linear(h, sd["roi_heads.box_predictor.bbox_pred.weight"], sd["roi_heads.box_predictor.bbox_pred.bias"])
    dets = o.box_predictor_inference(cls, reg, boxes, image_sizes, 0.05, 0.5, 100)
    pseudo, _ = o.process_pseudo_label(dets, bbox_threshold, "roih", "thresholding")
    return props, dets, pseudo


@torch.no_grad()
def ema_update(student_sd: Dict[str, torch.Tensor], teacher_sd: Dict[str, torch.Tensor], keep_rate: float = 0.9996) -> None:
    o.load_state_dict_like(teacher_sd, o.update_teacher_model(student_sd, teacher_sd, keep_rate))
