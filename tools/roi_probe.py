"""Runs the ROIAlign forward / backward of one configuration a few times (for ncu captures)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sfod_b200  # noqa
from sfod_b200 import ops, synth

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="R101"); ap.add_argument("--n", type=int, default=1); ap.add_argument("--r", type=int, default=2000)
ap.add_argument("--iters", type=int, default=3); ap.add_argument("--layout", default="nhwc"); ap.add_argument("--bwd", type=int, default=0)
a = ap.parse_args()
cfg = getattr(synth, a.cfg)
dev = "cuda"
x = synth.features(cfg, a.n, 1).to(dev)
if a.layout == "nhwc":
    x = x.contiguous(memory_format=torch.channels_last)
logits, deltas, cell, anchors = synth.rpn_head_outputs(cfg, a.n, 2)
boxes, lg, src, cnt, _ = ops.rpn_select(logits.to(dev), deltas.to(dev), [cfg["image"]] * a.n, cell_anchors=cell,
                                        feat_hw=(cfg["H"], cfg["W"]), stride=cfg["stride"], post_nms_topk=a.r)
rois = ops.convert_boxes_to_roi_format([boxes[i] for i in range(a.n)])
if a.bwd:
    rois = rois[torch.randperm(rois.shape[0], device=dev)[:a.bwd * a.n]].contiguous()
    xg = x.clone().requires_grad_(True)
    y = ops.roi_align(xg, rois, (7, 7), 1.0 / cfg["stride"], 0, True)
    g = torch.randn_like(y)
    for _ in range(a.iters):
        torch.autograd.grad(y, xg, g, retain_graph=True)
else:
    for _ in range(a.iters):
        y = ops.roi_align(x, rois, (7, 7), 1.0 / cfg["stride"], 0, True)
torch.cuda.synchronize()
w = (rois[:, 3] - rois[:, 1]) / cfg["stride"]; h = (rois[:, 4] - rois[:, 2]) / cfg["stride"]
print("rois", rois.shape[0], "mean w/h cells", w.mean().item(), h.mean().item(), "max", w.max().item(), h.max().item())
