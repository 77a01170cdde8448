"""Generates the golden fixtures of tests/golden/ from the INSTALLED native CPU kernels the reference binds to
(torchvision 0.26.0 CPU ``nms`` / ``batched_nms`` / ``roi_align`` / ``roi_pool``, ATen ``batch_norm`` / elementwise, and
torchvision's ``BoxCoder.decode_single`` as the importable twin of detectron2's ``Box2BoxTransform.apply_deltas``,
SURVEY.md A-2).  Run once in the build container:  ``python tests/golden/make_golden.py``.

The reference ships no tests and no golden vectors of its own, so these files -- plus the known-answer
vectors of SURVEY.md Appendix B re-derived below into kat.json -- are what pins the oracle.  Nothing here imports
``oracle/`` or the product: the fixtures are independent ground truth.
"""
import json
import math
import os

import numpy as np
import torch
import torchvision
from torchvision.models.detection._utils import BoxCoder

HERE = os.path.dirname(os.path.abspath(__file__))


def boxes_clustered(n, seed, w=1200, h=600, clusters=40):
    g = torch.Generator().manual_seed(seed)
    ctr = torch.rand(clusters, 2, generator=g) * torch.tensor([w, h], dtype=torch.float32)
    wh = torch.rand(clusters, 2, generator=g) * 200 + 20
    base = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)
    idx = torch.randint(0, clusters, (n,), generator=g)
    b = base[idx] + torch.randn(n, 4, generator=g) * 4
    b = torch.stack([torch.minimum(b[:, 0], b[:, 2]), torch.minimum(b[:, 1], b[:, 3]),
                     torch.maximum(b[:, 0], b[:, 2]) + 1e-3, torch.maximum(b[:, 1], b[:, 3]) + 1e-3], 1)
    s = torch.randn(n, generator=g) + torch.arange(n, dtype=torch.float32) * 2.0 ** -20
    return b.contiguous(), s.contiguous()


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    kat = {"versions": {"torch": torch.__version__, "torchvision": torchvision.__version__}}

    # ---- B-1 NMS semantics
    b = torch.tensor([[0, 0, 10, 10], [0, 0, 10, 10], [20, 20, 30, 30], [20, 20, 30, 30]], dtype=torch.float32)
    kat["nms_ties_lowest_index"] = torchvision.ops.nms(b, torch.ones(4), 0.5).tolist()
    b = torch.tensor([[5, 5, 5, 5], [5, 5, 5, 5], [0, 0, 10, 10]], dtype=torch.float32)
    kat["nms_degenerate_never_suppressed"] = torchvision.ops.nms(b, torch.tensor([0.9, 0.8, 0.7]), 0.5).tolist()
    b = torch.tensor([[0, 0, 10, 10], [0, 0, 10, 5]], dtype=torch.float32)
    kat["nms_iou_equal_threshold_kept"] = torchvision.ops.nms(b, torch.tensor([0.9, 0.8]), 0.5).tolist()
    # ---- B-3 ROIAlign / ROIPool / apply_deltas / anchors
    x = torch.arange(25, dtype=torch.float32).reshape(1, 1, 5, 5)
    rois = torch.tensor([[0, 1, 1, 3, 3]], dtype=torch.float32)
    kat["roi_align_aligned_false"] = torchvision.ops.roi_align(x, rois, (4, 4), 1.0, 0, False).reshape(4, 4).tolist()
    kat["roi_align_aligned_true"] = torchvision.ops.roi_align(x, rois, (4, 4), 1.0, 0, True).reshape(4, 4).tolist()
    kat["roi_pool_2x2"] = torchvision.ops.roi_pool(x, rois, (2, 2), 1.0).reshape(2, 2).tolist()
    coder = BoxCoder((1.0, 1.0, 1.0, 1.0), bbox_xform_clip=math.log(1000.0 / 16))
    d = torch.tensor([[0.1, -0.2, 0.3, 10.0]]); bx = torch.tensor([[10.0, 20.0, 50.0, 100.0]])
    kat["apply_deltas_clamped"] = [float(v) for v in coder.decode_single(d, bx)[0]]
    kat["scale_clamp"] = math.log(1000.0 / 16)
    kat["cell_anchor_32_ar05"] = [-math.sqrt(32 ** 2 / 0.5) / 2, -0.5 * math.sqrt(32 ** 2 / 0.5) / 2]
    # ---- B-5 EMA constants
    kat["ema_one_minus_k_f32"] = float(np.float32(1 - 0.9996)); kat["ema_k_f32"] = float(np.float32(0.9996))
    kat["ema_int64"] = int((torch.tensor(1000) * (1 - 0.9996) + torch.tensor(1000) * 0.9996).to(torch.int64))
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1, sort_keys=True)

    # ---- NMS fixtures (keep indices are the bit-exact contract)
    out = {}
    for n, seed in ((257, 1), (1500, 2)):
        b, s = boxes_clustered(n, seed)
        out[f"boxes_{n}"] = b.numpy(); out[f"scores_{n}"] = s.numpy()
        for thr in (0.5, 0.7):
            out[f"keep_{n}_{int(thr * 10)}"] = torchvision.ops.nms(b, s, thr).numpy()
    # batched: coordinate trick (n <= 1000 on CPU) and vanilla (n > 1000)
    for n, seed in ((900, 3), (1200, 4)):
        b, s = boxes_clustered(n, seed)
        idx = torch.randint(0, 8, (n,), generator=torch.Generator().manual_seed(seed))
        out[f"bboxes_{n}"] = b.numpy(); out[f"bscores_{n}"] = s.numpy(); out[f"bidx_{n}"] = idx.numpy()
        out[f"bkeep_{n}"] = torchvision.ops.batched_nms(b, s, idx, 0.5).numpy()
    np.savez_compressed(os.path.join(HERE, "nms.npz"), **out)

    # ---- ROIAlign / ROIPool fixtures (forward bit-exact for the C oracle, backward 1e-5)
    g = torch.Generator().manual_seed(10)
    x = torch.randn(2, 6, 18, 37, generator=g)
    ctr = torch.rand(24, 2, generator=g) * torch.tensor([1200.0, 600.0]); wh = torch.rand(24, 2, generator=g) ** 2 * torch.tensor([900.0, 450.0]) + 4
    rois = torch.cat([torch.randint(0, 2, (24, 1), generator=g).float(), ctr - wh / 2, ctr + wh / 2], 1)
    rois = torch.cat([rois, torch.tensor([[0, -500, -500, -100, -100], [1, 100, 100, 100, 100], [0, 300, 300, 200, 250], [1, 0, 0, 1200, 600]], dtype=torch.float32)])
    out = {"x": x.numpy(), "rois": rois.numpy()}
    gout = torch.randn(rois.shape[0], 6, 7, 7, generator=g)
    out["grad_out"] = gout.numpy()
    for aligned in (True, False):
        xr = x.clone().requires_grad_(True)
        y = torchvision.ops.roi_align(xr, rois, (7, 7), 1 / 32, 0, aligned)
        y.backward(gout)
        out[f"align_{int(aligned)}"] = y.detach().numpy(); out[f"align_grad_{int(aligned)}"] = xr.grad.numpy()
    xr = x.clone().requires_grad_(True)
    y = torchvision.ops.roi_pool(xr, rois, (7, 7), 1 / 32)
    y.backward(gout)
    out["pool"] = y.detach().numpy(); out["pool_grad"] = xr.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "roi.npz"), **out)

    # ---- box decode (BoxCoder twin), softmax, EMA, BatchNorm
    g = torch.Generator().manual_seed(20)
    ctr = torch.rand(200, 2, generator=g) * 1000; wh = torch.rand(200, 2, generator=g) * 300 + 1
    boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)
    deltas = torch.randn(200, 32, generator=g) * 2
    deltas[0, 3] = 100.0
    dec = BoxCoder((10.0, 10.0, 5.0, 5.0), bbox_xform_clip=math.log(1000.0 / 16)).decode_single(deltas, boxes).reshape(200, 32)
    logits = torch.randn(200, 9, generator=g) * 4
    s = torch.randn(4099, generator=g); t = torch.randn(4099, generator=g)
    ema = s * (1 - 0.9996) + t * 0.9996
    xb = torch.randn(2, 5, 33, 47, generator=g) * 3 + torch.linspace(-40, 40, 5).view(1, -1, 1, 1)
    bn = torch.nn.BatchNorm2d(5)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5, generator=g); bn.bias.uniform_(-1, 1, generator=g)
        bn.running_mean.normal_(generator=g); bn.running_var.uniform_(0.5, 2.0, generator=g)
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    bn.train()
    with torch.no_grad():
        yb = bn(xb)
    np.savez_compressed(os.path.join(HERE, "dense.npz"), boxes=boxes.numpy(), deltas=deltas.numpy(), decoded=dec.numpy(), logits=logits.numpy(),
                        softmax=torch.softmax(logits, -1).numpy(), ema_s=s.numpy(), ema_t=t.numpy(), ema_out=ema.numpy(),
                        bn_x=xb.numpy(), bn_w=bn.weight.detach().numpy(), bn_b=bn.bias.detach().numpy(), bn_rm0=rm0.numpy(), bn_rv0=rv0.numpy(),
                        bn_rm1=bn.running_mean.numpy(), bn_rv1=bn.running_var.numpy(), bn_y=yb.numpy())
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
