"""Executes the ctypes snippets of INTEGRATION.md section 1 (torchvision-compatible nms; sfod_rpn_select with the head outputs
as they lie) on cuda:0 and compares them with ``sfod_b200.ops`` -- the documentation is run, not trusted."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir(ROOT)
import torch  # noqa: E402
import sfod_b200  # noqa: E402,F401
from sfod_b200 import ops, synth  # noqa: E402

text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
ns = {}
exec(blocks[0], ns)   # nms stub
exec(blocks[1], ns)   # RpnParams / predict_proposals (re-uses `lib`, `C`, `torch` of the first block)
dev = "cuda:0"

boxes, scores = synth.boxes_high_suppression(4000, 3)
keep = ns["nms"](boxes.to(dev), scores.to(dev), 0.5)
ref = ops.nms(boxes.to(dev), scores.to(dev), 0.5)
assert torch.equal(keep, ref), "nms snippet differs from ops.nms"

cfg = synth.V
logits, deltas, cell, anchors = synth.rpn_head_outputs(cfg, 2, 11)
N, H, W = 2, cfg["H"], cfg["W"]
A = logits.shape[1] // (H * W)
lg4 = logits.view(N, H, W, A).permute(0, 3, 1, 2).contiguous().to(dev)
dl4 = deltas.view(N, H, W, A, 4).permute(0, 3, 4, 1, 2).reshape(N, 4 * A, H, W).contiguous().to(dev)
hw = torch.tensor([[600, 1200], [576, 1100]], dtype=torch.int32, device=dev)
got = ns["predict_proposals"](lg4, dl4, cell, cfg["stride"], hw)
want = ops.rpn_select(logits.to(dev), deltas.to(dev), [(600, 1200), (576, 1100)], cell_anchors=cell, feat_hw=(H, W), stride=cfg["stride"])
for g, w_ in zip(got, want):
    assert torch.equal(g, w_), "predict_proposals snippet differs from ops.rpn_select"
print("INTEGRATION.md snippets ok: nms", int(keep.numel()), "kept; proposals", got[3].tolist())
