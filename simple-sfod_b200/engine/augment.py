"""Strong augmentation of the student's batch on the GPU.

The reference builds, per image and on the CPU (PIL), reference daod/data/detection_utils.py:7-37:

    RandomApply([ColorJitter(0.4, 0.4, 0.4, 0.1)], p=0.8) -> RandomGrayscale(p=0.2) -> RandomApply([GaussianBlur((0.1, 2.0))], p=0.5)
    -> ToTensor -> RandomErasing(0.7, (0.05, 0.2), (0.3, 3.3), "random") -> RandomErasing(0.5, (0.02, 0.2), (0.1, 6), "random")
    -> RandomErasing(0.3, (0.02, 0.2), (0.05, 8), "random") -> ToPILImage

and applies it in the data mapper (daod/data/mappers/two_crop_augmentation_mapper.py:141-157).  Here the random DECISIONS are
drawn on the host with the same distributions and in the same order as torchvision draws them (``draw_params``), and the
pixel work runs on the batch in HBM (``ops.color_jitter`` / ``ops.gaussian_blur_pil`` / ``ops.random_erase_``: 4 launches per batch
instead of ~10 PIL passes per image on one CPU core).  Images are RGB uint8 (N, 3, H, W), as the mapper's PIL images are.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
from torch import Tensor

from .. import ops

JITTER = dict(brightness=(0.6, 1.4), contrast=(0.6, 1.4), saturation=(0.6, 1.4), hue=(-0.1, 0.1), p=0.8)   # ColorJitter(0.4, 0.4, 0.4, 0.1)
GRAYSCALE_P = 0.2
BLUR = dict(sigma=(0.1, 2.0), p=0.5)
ERASERS = [dict(p=0.7, scale=(0.05, 0.2), ratio=(0.3, 3.3)), dict(p=0.5, scale=(0.02, 0.2), ratio=(0.1, 6.0)),
           dict(p=0.3, scale=(0.02, 0.2), ratio=(0.05, 8.0))]


def _uniform(g: torch.Generator, lo: float, hi: float) -> float:
    return float(torch.empty(1).uniform_(lo, hi, generator=g))


def draw_erase_rect(g: torch.Generator, H: int, W: int, scale, ratio):
    """torchvision RandomErasing.get_params (10 attempts; None if no rectangle fits)."""
    area = H * W
    log_ratio = (math.log(ratio[0]), math.log(ratio[1]))
    for _ in range(10):
        erase_area = area * _uniform(g, scale[0], scale[1])
        aspect_ratio = math.exp(_uniform(g, log_ratio[0], log_ratio[1]))
        h = int(round(math.sqrt(erase_area * aspect_ratio)))
        w = int(round(math.sqrt(erase_area / aspect_ratio)))
        if not (h < H and w < W):
            continue
        i = int(torch.randint(0, H - h + 1, size=(1,), generator=g))
        j = int(torch.randint(0, W - w + 1, size=(1,), generator=g))
        return (i, j, h, w)
    return None


def draw_params(N: int, H: int, W: int, generator: Optional[torch.Generator] = None) -> List[Dict]:
    """The per-image decisions of the reference's strong augmentation: same distributions, same order of draws."""
    g = generator if generator is not None else torch.default_generator
    out = []
    for _ in range(N):
        p: Dict = {"order": [], "factors": [], "grayscale": False, "sigma": None, "rects": []}
        if not (JITTER["p"] < float(torch.rand(1, generator=g))):           # RandomApply
            fn_idx = torch.randperm(4, generator=g).tolist()                # ColorJitter.get_params
            vals = [_uniform(g, *JITTER["brightness"]), _uniform(g, *JITTER["contrast"]), _uniform(g, *JITTER["saturation"]),
                    _uniform(g, *JITTER["hue"])]
            p["order"], p["factors"] = fn_idx, [vals[k] for k in fn_idx]
        p["grayscale"] = bool(float(torch.rand(1, generator=g)) < GRAYSCALE_P)
        if not (BLUR["p"] < float(torch.rand(1, generator=g))):
            p["sigma"] = _uniform(g, *BLUR["sigma"])
        for e in ERASERS:
            if float(torch.rand(1, generator=g)) < e["p"]:
                r = draw_erase_rect(g, H, W, e["scale"], e["ratio"])
                if r is not None:
                    p["rects"].append(r)
        out.append(p)
    return out


def strong_augment(images: Tensor, params: Optional[List[Dict]] = None, generator: Optional[torch.Generator] = None,
                   noise: Optional[Tensor] = None, seed: Optional[int] = None, blur: str = "pil", arithmetic: str = "pil") -> Tensor:
    """(N, 3, H, W) uint8 RGB CUDA batch -> strongly augmented uint8 batch (a new tensor).

    ``arithmetic="pil"`` (default): ColorJitter / RandomGrayscale in Pillow's arithmetic (torchvision's PIL path, which the
    reference's mapper takes: daod/data/mappers/two_crop_augmentation_mapper.py:141-157), bit for bit; ``"tensor"``: torchvision's
    uint8-tensor arithmetic.

    ``blur="pil"`` (default) is the reference's filter, ``PIL.ImageFilter.GaussianBlur(radius=sigma)`` (reference
    daod/data/transforms/augmentations.py:18-21) reproduced bit for bit (``ops.gaussian_blur_pil``: Pillow's three extended
    box-blur passes per axis); ``blur="gaussian"`` is a true Gaussian convolution with torchvision's taps (``ops.gaussian_blur``)."""
    if blur not in ("pil", "gaussian"):
        raise ValueError("blur must be 'pil' or 'gaussian'")
    N, _, H, W = images.shape
    if params is None:
        params = draw_params(N, H, W, generator)
    x = ops.color_jitter(images, params, arithmetic=arithmetic)
    if any(p["sigma"] is not None for p in params):
        x = (ops.gaussian_blur_pil if blur == "pil" else ops.gaussian_blur)(x, [p["sigma"] for p in params])
    if any(p["rects"] for p in params):
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,), generator=generator if generator is not None else torch.default_generator))
        ops.random_erase_(x, [p["rects"] for p in params], noise=noise, seed=seed)
    return x
