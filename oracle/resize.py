"""CPU oracle for the image geometry of `load_image` (reference img2latex/data/utils.py:18-90):
`ResizeWithAspectRatio` (img2latex/data/transforms.py:9-56) = aspect-preserving LANCZOS resize to the
target height, then white right-padding or a centre crop to the target width.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The resampling arithmetic lives in a third-party dependency that is not vendored under
/root/reference: **Pillow** (`pillow` is an unpinned requirement; 12.2.0 is installed in the build
container and is the de-facto pin).  `Image.resize(size, LANCZOS)` for 8-bit modes is restated here
from Pillow's published algorithm (src/libImaging/Resample.c: `precompute_coeffs`,
`normalize_coeffs_8bpc`, `ImagingResampleHorizontal_8bpc`, `ImagingResampleVertical_8bpc`,
`ImagingResample`): per output pixel a window of `support * max(scale, 1)` source pixels around the
centre `(xx + 0.5) * scale`, truncated-sinc weights normalised to sum 1, rounded to 22-bit fixed
point; a horizontal pass then a vertical pass, each accumulating in int32 from a rounding bias of
`1 << 21`, shifting right by 22 and clamping to [0, 255] (the intermediate image is uint8).
`convert("L")` is restated from src/libImaging/Convert.c (`rgb2l`: ITU-R 601-2 luma in 16-bit fixed
point).  Pinned against live Pillow by tests/golden/make_golden.py::resize_case (reference
`ResizeWithAspectRatio` executed on seeded images) and tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

__all__ = ["lanczos_coeffs", "resize_lanczos_u8", "resize_with_aspect_ratio", "rgb_to_l", "aspect_width",
           "PRECISION_BITS", "LANCZOS_SUPPORT"]

PRECISION_BITS = 32 - 8 - 2          # Resample.c: #define PRECISION_BITS (32 - 8 - 2)
LANCZOS_SUPPORT = 3.0                # Resample.c: static struct filter LANCZOS = {lanczos_filter, 3.0}


def _sinc(x: float) -> float:
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos(x: float) -> float:
    """Resample.c lanczos_filter: truncated sinc, a = 3."""
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3)
    return 0.0


def _bicubic(x: float) -> float:
    """Resample.c bicubic_filter (Keys, a = -0.5, support 2): Pillow's default for `Image.resize(size)`."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


_FILTERS = {"lanczos": (_lanczos, LANCZOS_SUPPORT), "bicubic": (_bicubic, 2.0)}


def lanczos_coeffs(in_size: int, out_size: int, resample: str = "lanczos") -> Tuple[int, np.ndarray, np.ndarray]:
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full box (in0 = 0, in1 = in_size).
    Returns ksize, bounds (out_size, 2) int32 = (xmin, count), kk (out_size, ksize) int32."""
    _filter, fsupport = _FILTERS[resample]
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = fsupport * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)        # C (int) cast: truncation toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)       # arithmetic shift, clip8_lookups


def _pass(img: np.ndarray, out_size: int, axis: int, resample: str = "lanczos") -> np.ndarray:
    """One resampling pass along `axis` of an (H, W, C) uint8 array."""
    in_size = img.shape[axis]
    _, bounds, kk = lanczos_coeffs(in_size, out_size, resample)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        xmin, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.tensordot(kk[xx, :n].astype(np.int64), src[xmin:xmin + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        acc = ((acc + (1 << 31)) % (1 << 32)) - (1 << 31)                # int32 wrap-around of the C accumulator
        out[xx] = _clip8(acc)
    return np.moveaxis(out, 0, axis)


def resize_lanczos_u8(img: np.ndarray, out_w: int, out_h: int, resample: str = "lanczos") -> np.ndarray:
    """`Image.resize((out_w, out_h), LANCZOS)` for modes L / RGB (resample="bicubic": `Image.resize(size)`,
    the PIL branch of Predictor._prepare_image, training/predictor.py:436).  img: (H, W) or (H, W, C) uint8.
    Resample.c ImagingResample: horizontal pass first (skipped when the width is unchanged), then the
    vertical pass (skipped when the height is unchanged)."""
    squeeze = img.ndim == 2
    a = img[:, :, None] if squeeze else img
    if out_w <= 0 or out_h <= 0:
        raise ValueError("height and width must be > 0")                 # Pillow raises ValueError
    if a.shape[1] != out_w:
        a = _pass(a, out_w, 1, resample)
    if a.shape[0] != out_h:
        a = _pass(a, out_h, 0, resample)
    a = np.ascontiguousarray(a)
    return a[:, :, 0] if squeeze else a


def aspect_width(width: int, height: int, target_height: int) -> int:
    """transforms.py:31-35: int(round(target_height * (width / height))) -- Python's round-half-even."""
    return int(round(target_height * (width / height)))


def _blank(img: np.ndarray, target_height: int, target_width: int) -> np.ndarray:
    """`Image.new(img.mode, size, 255)` (transforms.py:29,45-47): for mode L that is white, but for mode RGB
    Pillow reads the integer colour 255 as 0x0000FF = (R=255, G=0, B=0) -- the reference pads RGB images
    with RED (pinned by tests/golden/resize.npz)."""
    out = np.zeros((target_height, target_width) + img.shape[2:], np.uint8)
    if img.ndim == 2:
        out[:] = 255
    else:
        out[:, :, 0] = 255
    return out


def resize_with_aspect_ratio(img: np.ndarray, target_height: int, target_width: int) -> np.ndarray:
    """transforms.py:26-56 on an (H, W) / (H, W, C) uint8 array."""
    height, width = img.shape[0], img.shape[1]
    if height == 0:                                                      # transforms.py:28-29
        return _blank(img, target_height, target_width)
    new_width = aspect_width(width, height, target_height)
    r = resize_lanczos_u8(img, new_width, target_height)
    if new_width == target_width:
        return r
    if new_width < target_width:                                         # right padding, 44-50
        out = _blank(img, target_height, target_width)
        out[:, :new_width] = r
        return out
    left = (new_width - target_width) // 2                               # centre crop, 51-56
    return np.ascontiguousarray(r[:, left:left + target_width])


def rgb_to_l(img: np.ndarray) -> np.ndarray:
    """`Image.convert("L")` from RGB (Convert.c rgb2l / L24): (R*19595 + G*38470 + B*7471 + 0x8000) >> 16."""
    a = img.astype(np.int64)
    return ((a[..., 0] * 19595 + a[..., 1] * 38470 + a[..., 2] * 7471 + 0x8000) >> 16).astype(np.uint8)
