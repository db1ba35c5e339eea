"""Host-side coefficient tables for the Pillow-exact uint8 bilinear resize kernel (csrc/vit_kernels.cu).

`ResizeLongestSide.apply_image` (upstream segment_anything/utils/transforms.py; reference call path
sam_pt/modeling/sam_pt.py:849 -> SamPredictor.set_image) resizes with PIL: an antialiased, separable, 22-bit fixed-point
filter (Pillow src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal/Vertical_8bpc).
The tables below restate precompute_coeffs + normalize_coeffs_8bpc in float64; the kernel does the integer passes."""
from __future__ import annotations

import math
from functools import lru_cache
from typing import Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2


@lru_cache(maxsize=64)
def bilinear_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """-> (bounds int32 [out,2] = (first input index, tap count), coeffs int32 [out, ksize], ksize)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale  # bilinear filter support is 1.0
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            a = -a if a < 0 else a
            w = 1.0 - a if a < 1.0 else 0.0
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            kk[xx, :xmax] /= ww
        bounds[xx, 0], bounds[xx, 1] = xmin, xmax
    scaled = kk * (1 << PRECISION_BITS)
    coeffs = np.where(kk < 0, (-0.5 + scaled).astype(np.int64), (0.5 + scaled).astype(np.int64)).astype(np.int32)
    return bounds, coeffs, ksize


def resize_reference_numpy(img_hwc_u8: np.ndarray, out_hw: Tuple[int, int]) -> np.ndarray:
    """Integer emulation of the two CUDA passes (used by the CPU test that pins the tables against PIL itself)."""
    H, W, C = img_hwc_u8.shape
    Ho, Wo = out_hw
    hb, hk, _ = bilinear_coeffs(W, Wo)
    vb, vk, _ = bilinear_coeffs(H, Ho)
    src = img_hwc_u8.astype(np.int64)
    tmp = np.zeros((H, Wo, C), dtype=np.int64)
    for xo in range(Wo):
        x0, n = hb[xo]
        acc = (1 << (PRECISION_BITS - 1)) + (src[:, x0:x0 + n, :] * hk[xo, :n][None, :, None].astype(np.int64)).sum(axis=1)
        tmp[:, xo, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
    out = np.zeros((Ho, Wo, C), dtype=np.int64)
    for yo in range(Ho):
        y0, n = vb[yo]
        acc = (1 << (PRECISION_BITS - 1)) + (tmp[y0:y0 + n, :, :] * vk[yo, :n][:, None, None].astype(np.int64)).sum(axis=0)
        out[yo] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return out.astype(np.uint8)
