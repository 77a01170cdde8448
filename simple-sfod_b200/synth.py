"""Synthetic inputs of SURVEY.md section 8(d) (seeded, tie-free), shared by tests, smoke() and bench.py.
Pure torch-CPU generation (no oracle dependency); callers move tensors to the GPU."""
from __future__ import annotations

import math
from typing import Tuple

import torch

V = dict(name="V", C=512, H=18, W=37, stride=32, sizes=(32, 64, 128, 256, 512), ratios=(0.5, 1.0, 2.0), image=(600, 1200))
R101 = dict(name="R", C=1024, H=38, W=75, stride=16, sizes=(64, 128, 256, 512), ratios=(0.5, 1.0, 2.0), image=(600, 1200))


def cell_anchors(sizes, ratios) -> torch.Tensor:
    out = []
    for s in sizes:
        area = s ** 2.0
        for ar in ratios:
            w = math.sqrt(area / ar)
            h = ar * w
            out.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
    return torch.tensor(out, dtype=torch.float32)


def grid_anchors(H, W, stride, cell) -> torch.Tensor:
    sx = torch.arange(0, W * stride, step=stride, dtype=torch.float32)
    sy = torch.arange(0, H * stride, step=stride, dtype=torch.float32)
    yy, xx = torch.meshgrid(sy, sx, indexing="ij")
    shifts = torch.stack((xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)), dim=1)
    return (shifts.view(-1, 1, 4) + cell.view(1, -1, 4)).reshape(-1, 4)


def tie_free(x: torch.Tensor) -> torch.Tensor:
    """Adds arange * 2^-20 jitter along the last dim.  This makes exact ties rare, not impossible;
    both the oracle and the CUDA path break remaining ties canonically (value desc, index asc)."""
    n = x.shape[-1]
    return x + torch.arange(n, dtype=torch.float32) * 2.0 ** -20


def rpn_head_outputs(cfg, N: int, seed: int, delta_std: float = 0.5):
    """(logits (N,HWA), deltas (N,HWA,4), cell (A,4), anchors (HWA,4))"""
    g = torch.Generator().manual_seed(seed)
    cell = cell_anchors(cfg["sizes"], cfg["ratios"])
    anchors = grid_anchors(cfg["H"], cfg["W"], cfg["stride"], cell)
    hwa = anchors.shape[0]
    logits = tie_free(torch.randn(N, hwa, generator=g))
    deltas = torch.randn(N, hwa, 4, generator=g) * delta_std
    return logits, deltas, cell, anchors


def boxes_low_suppression(cfg, n: int, seed: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """anchors (+) N(0,0.5) deltas, clipped, top-n by logit: ~73 % survive NMS 0.7 at n=9990."""
    logits, deltas, cell, anchors = rpn_head_outputs(cfg, 1, seed)
    d = deltas[0]
    w, h = anchors[:, 2] - anchors[:, 0], anchors[:, 3] - anchors[:, 1]
    cx, cy = anchors[:, 0] + 0.5 * w, anchors[:, 1] + 0.5 * h
    pcx, pcy = d[:, 0] * w + cx, d[:, 1] * h + cy
    pw, ph = torch.exp(d[:, 2].clamp(max=4.135)) * w, torch.exp(d[:, 3].clamp(max=4.135)) * h
    boxes = torch.stack((pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph), dim=1)
    ih, iw = cfg["image"]
    boxes[:, 0::2] = boxes[:, 0::2].clamp(0, iw)
    boxes[:, 1::2] = boxes[:, 1::2].clamp(0, ih)
    keep = ((boxes[:, 2] - boxes[:, 0]) > 0) & ((boxes[:, 3] - boxes[:, 1]) > 0)
    boxes, sc = boxes[keep], logits[0][keep]
    order = sc.sort(descending=True, stable=True)[1][:n]
    return boxes[order].contiguous(), sc[order].contiguous()


def boxes_high_suppression(n: int, seed: int, image=(600, 1200), clusters: int = 200) -> Tuple[torch.Tensor, torch.Tensor]:
    """`clusters` random centres x jittered copies: the distribution a trained RPN produces."""
    g = torch.Generator().manual_seed(seed)
    h, w = image
    ctr = torch.rand(clusters, 2, generator=g) * torch.tensor([w, h], dtype=torch.float32)
    wh = torch.rand(clusters, 2, generator=g) * 200 + 20
    base = torch.cat([ctr - wh / 2, ctr + wh / 2], dim=1)
    idx = torch.randint(0, clusters, (n,), generator=g)
    b = base[idx] + torch.randn(n, 4, generator=g) * 4.0
    b[:, 0::2] = b[:, 0::2].clamp(0, w)
    b[:, 1::2] = b[:, 1::2].clamp(0, h)
    b = torch.stack([torch.minimum(b[:, 0], b[:, 2]), torch.minimum(b[:, 1], b[:, 3]),
                     torch.maximum(b[:, 0], b[:, 2]) + 1e-3, torch.maximum(b[:, 1], b[:, 3]) + 1e-3], dim=1)
    s = tie_free(torch.randn(n, generator=g))
    return b.contiguous(), s.contiguous()


def features(cfg, N: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(N, cfg["C"], cfg["H"], cfg["W"], generator=g)


def random_rois(N: int, R: int, seed: int, image=(600, 1200)) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    h, w = image
    ctr = torch.rand(R, 2, generator=g) * torch.tensor([w, h], dtype=torch.float32)
    wh = torch.rand(R, 2, generator=g) ** 2 * torch.tensor([w * 0.8, h * 0.8]) + 4
    b = torch.cat([ctr - wh / 2, ctr + wh / 2], dim=1)
    bi = torch.randint(0, N, (R, 1), generator=g).float()
    return torch.cat([bi, b], dim=1).contiguous()


def box_head_outputs(R: int, K: int, seed: int, logit_std: float = 4.0, delta_std: float = 1.0):
    g = torch.Generator().manual_seed(seed)
    cls = torch.randn(R, K + 1, generator=g) * logit_std
    dl = torch.randn(R, 4 * K, generator=g) * delta_std
    return cls, dl
