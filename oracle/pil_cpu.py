"""CPU restatement of Pillow's ``ImageFilter.GaussianBlur`` -- TEST INFRASTRUCTURE ONLY (never imported by the product).

The reference's strong augmentation blurs with Pillow (reference daod/data/transforms/augmentations.py:18-21:
``x.filter(ImageFilter.GaussianBlur(radius=sigma))``, sigma ~ U(0.1, 2.0), built at daod/data/detection_utils.py:17).  Pillow is a
third-party dependency of the reference (not vendored in /root/reference; the version installed here, against which this file is
pinned live by tests/test_oracle_cpu.py, is 12.2.0).  Its published algorithm, src/libImaging/BoxBlur.c:

* ``_gaussian_blur_radius(radius, passes=3)``: sigma^2 / passes -> ideal box length L = sqrt(12 s2 + 1) -> integer part
  l = floor((L - 1) / 2) and fractional part a = (2l + 1)(l(l + 1) - 3 s2) / (6 (s2 - (l + 1)^2)); float32 arithmetic.
* ``ImagingBoxBlur``: ``passes`` horizontal extended-box passes, transpose, ``passes`` more, transpose back.
* ``ImagingHorizontalBoxBlur`` / ``ImagingLineBoxBlur*``: r = int(R), ww = uint32(2^24 / (2R + 1)), fw = (2^24 - (2r + 1) ww) / 2,
  out[x] = (ww * sum_{|d|<=r} in[c(x+d)] + fw * (in[c(x-r-1)] + in[c(x+r+1)]) + 2^23) >> 24 with c = clamp to the line (the C code
  keeps a running sum; the closed form is the same integer), each pass rounded to uint8.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def gaussian_blur_radius(radius: float, passes: int = 3) -> np.float32:
    radius = f32(radius)
    sigma2 = f32(radius * radius / f32(passes))
    L = f32(np.sqrt(12.0 * float(sigma2) + 1.0))
    l = f32(np.floor((float(L) - 1.0) / 2.0))
    a = f32(f32(f32(2) * l + f32(1)) * f32(f32(l * f32(l + f32(1))) - f32(f32(3) * sigma2)))
    a = f32(a / f32(f32(6) * f32(sigma2 - f32(f32(l + f32(1)) * f32(l + f32(1))))))
    return f32(l + a)


def box_weights(float_radius) -> tuple:
    fr = f32(float_radius)
    r = int(fr)
    ww = int(f32(16777216.0) / f32(fr * f32(2) + f32(1)))
    fw = ((1 << 24) - (r * 2 + 1) * ww) // 2
    return r, ww, fw


def horizontal_box_blur(img: np.ndarray, float_radius) -> np.ndarray:
    """One extended-box pass along axis 1 of an (H, W, C) uint8 array."""
    r, ww, fw = box_weights(float_radius)
    H, W, _ = img.shape
    x = np.arange(W)
    src = img.astype(np.int64)
    acc = np.zeros_like(src)
    for d in range(-r, r + 1):
        acc += src[:, np.clip(x + d, 0, W - 1), :]
    far = src[:, np.clip(x - r - 1, 0, W - 1), :] + src[:, np.clip(x + r + 1, 0, W - 1), :]
    return ((acc * ww + far * fw + (1 << 23)) >> 24).astype(np.uint8)


def gaussian_blur(img: np.ndarray, sigma: float, passes: int = 3) -> np.ndarray:
    """(H, W, C) uint8 -> what ``Image.fromarray(img).filter(ImageFilter.GaussianBlur(radius=sigma))`` returns."""
    R = gaussian_blur_radius(sigma, passes)
    out = img
    if R != 0:
        for _ in range(passes):
            out = horizontal_box_blur(out, R)
        t = np.ascontiguousarray(out.transpose(1, 0, 2))
        for _ in range(passes):
            t = horizontal_box_blur(t, R)
        out = np.ascontiguousarray(t.transpose(1, 0, 2))
    return out
