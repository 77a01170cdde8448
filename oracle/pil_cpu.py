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


# --------------------------------------------------------------------------------------------------------------------------
# ColorJitter / RandomGrayscale on PIL images: torchvision/transforms/_functional_pil.py (adjust_brightness / contrast / saturation
# -> ImageEnhance.Brightness / Contrast / Color -> Image.blend(degenerate, image, factor); adjust_hue -> convert("HSV"), shift,
# convert back; to_grayscale -> convert("L")) over Pillow's libImaging Blend.c / Convert.c.  The reference reaches this path because
# its mapper converts every image to PIL before the strong augmentation (reference
# daod/data/mappers/two_crop_augmentation_mapper.py:141-157).  Pinned live against torchvision + Pillow by tests/test_oracle_cpu.py.
def to_L(rgb: np.ndarray) -> np.ndarray:
    r, g, b = (rgb[..., k].astype(np.int64) for k in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def blend(degenerate: np.ndarray, image: np.ndarray, alpha: float) -> np.ndarray:
    """Image.blend(im1=degenerate, im2=image, alpha): float32 `im1 + alpha * (im2 - im1)`, truncated; clipped outside [0, 1]."""
    a = f32(alpha)
    i1, i2 = degenerate.astype(np.int32), image.astype(np.int32)
    t = (i1.astype(f32) + (a * (i2 - i1).astype(f32)).astype(f32)).astype(f32)
    if f32(0) <= a <= f32(1):
        return t.astype(np.int32).astype(np.uint8)
    return np.where(t <= 0, 0, np.where(t >= 255, 255, t.astype(np.int32))).astype(np.uint8)


def adjust_brightness(img: np.ndarray, factor: float) -> np.ndarray:
    return blend(np.zeros_like(img), img, factor)


def adjust_contrast(img: np.ndarray, factor: float) -> np.ndarray:
    mean = int(to_L(img).astype(np.float64).sum() / (img.shape[0] * img.shape[1]) + 0.5)
    return blend(np.full_like(img, mean), img, factor)


def adjust_saturation(img: np.ndarray, factor: float) -> np.ndarray:
    return blend(np.repeat(to_L(img)[..., None], 3, -1), img, factor)


def rgb_to_hsv(img: np.ndarray) -> np.ndarray:
    """Convert.c rgb2hsv_row: float variables, double literals (so `2.0 + rc - bc` and `h / 6.0 + 1.0` are evaluated in double)."""
    r, g, b = (img[..., k].astype(np.int32) for k in range(3))
    maxc, minc = np.maximum(r, np.maximum(g, b)), np.minimum(r, np.minimum(g, b))
    eq = maxc == minc
    cr = (maxc - minc).astype(f32)
    crs = np.where(eq, f32(1), cr)
    s = (cr / np.where(eq, f32(1), maxc.astype(f32))).astype(f32)
    rc, gc, bc = (((maxc - c).astype(f32) / crs).astype(f32) for c in (r, g, b))
    d = np.float64
    h = np.where(r == maxc, (bc - gc).astype(d), np.where(g == maxc, 2.0 + rc.astype(d) - bc.astype(d), 4.0 + gc.astype(d) - rc.astype(d))).astype(f32)
    h = np.fmod(h.astype(d) / 6.0 + 1.0, 1.0).astype(f32)
    uh = np.clip((h.astype(d) * 255.0).astype(np.int64), 0, 255)
    us = np.clip((s.astype(d) * 255.0).astype(np.int64), 0, 255)
    return np.stack([np.where(eq, 0, uh), np.where(eq, 0, us), maxc], -1).astype(np.uint8)


def hsv_to_rgb(hsv: np.ndarray) -> np.ndarray:
    """Convert.c hsv2rgb: i = floor(h * 6 / 255), p / q / t = round(v * (1 - ...)) with C's round (half away from zero)."""
    d = np.float64
    h, s, v = hsv[..., 0].astype(f32), hsv[..., 1], hsv[..., 2]
    h6 = h.astype(d) * 6.0 / 255.0
    i = np.floor(h6).astype(np.int64)
    f = (h6 - i.astype(f32).astype(d)).astype(f32).astype(d)
    fs = (s.astype(f32).astype(d) / 255.0).astype(f32).astype(d)
    vf = v.astype(f32).astype(d)

    def c_round(x):
        return np.where(x >= 0, np.floor(x + 0.5), np.ceil(x - 0.5)).astype(np.int64)
    p = np.clip(c_round(vf * (1.0 - fs)), 0, 255)
    q = np.clip(c_round(vf * (1.0 - fs * f)), 0, 255)
    t = np.clip(c_round(vf * (1.0 - fs * (1.0 - f))), 0, 255)
    vv = v.astype(np.int64)
    k = i % 6
    out = np.stack([np.choose(k, [vv, q, p, p, t, vv]), np.choose(k, [t, vv, vv, q, p, p]), np.choose(k, [p, p, t, vv, vv, q])], -1)
    return np.where((s == 0)[..., None], vv[..., None], out).astype(np.uint8)


def adjust_hue(img: np.ndarray, hue_factor: float) -> np.ndarray:
    hsv = rgb_to_hsv(img).copy()
    shift = int(hue_factor * 255) & 0xFF            # np.int32(hue_factor * 255).astype(np.uint8)
    hsv[..., 0] = ((hsv[..., 0].astype(np.int32) + shift) & 0xFF).astype(np.uint8)
    return hsv_to_rgb(hsv)


def color_jitter(img: np.ndarray, order, factors, grayscale: bool = False) -> np.ndarray:
    """ColorJitter.forward's loop over the drawn permutation (0 brightness, 1 contrast, 2 saturation, 3 hue), then RandomGrayscale."""
    fns = (adjust_brightness, adjust_contrast, adjust_saturation, adjust_hue)
    for o, f in zip(order, factors):
        img = fns[o](img, f)
    if grayscale:
        img = np.repeat(to_L(img)[..., None], 3, -1)
    return img
