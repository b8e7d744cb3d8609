"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the pre-processing in front of the LPSR forward.

Reference: inference/run.py:80-96 `preprocess_for_sr`: BGR -> RGB (cv2.cvtColor), PIL `Image.resize((192, 32), Image.BICUBIC)`,
torchvision `ToTensor()` (uint8 HWC -> float32 CHW / 255), `unsqueeze(0)`.

The resize is Pillow's `ImagingResample` (Pillow 12.2.0, src/libImaging/Resample.c; a third-party dependency that is not vendored in
the reference): an antialiased separable convolution in 8-bit fixed point --
  * bicubic kernel with a = -0.5, support 2, stretched by max(scale, 1) when downscaling (Resample.c: bicubic_filter,
    precompute_coeffs), weights normalised per output pixel in double precision,
  * weights rounded to 22 fractional bits (PRECISION_BITS = 32 - 8 - 2, normalize_coeffs_8bpc: (int)(+-0.5 + w * 2^22)),
  * horizontal pass first (only the rows the vertical pass needs), each pass accumulating in int32 from 2^21 and clipping
    (acc >> 22) to [0, 255] (ImagingResampleHorizontal_8bpc / ImagingResampleVertical_8bpc), uint8 between the passes.
Pinned by tests/golden/preprocess_*.npz, generated with Pillow itself (tests/golden/make_golden_preprocess.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bicubic_filter(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c: precompute_coeffs + normalize_coeffs_8bpc for the full-image box.  Returns (bounds [out,2], kk int32 [out,ksize])."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def pil_bicubic_resize_u8(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """img: uint8 [H, W, C].  Same result as PIL.Image.fromarray(img).resize((out_w, out_h), Image.BICUBIC)."""
    H, W, C = img.shape
    cur = img
    bounds_v, kk_v = precompute_coeffs(H, out_h) if out_h != H else (None, None)
    if out_w != W:
        bounds_h, kk_h = precompute_coeffs(W, out_w)
        # rows the vertical pass reads (Resample.c: ybox_first / ybox_last)
        if bounds_v is not None:
            y0 = int(bounds_v[0, 0])
            y1 = int(bounds_v[-1, 0] + bounds_v[-1, 1])
        else:
            y0, y1 = 0, H
        tmp = np.zeros((y1 - y0, out_w, C), dtype=np.uint8)
        src = img[y0:y1].astype(np.int64)
        for xx in range(out_w):
            xmin, xmax = bounds_h[xx]
            acc = np.full((y1 - y0, C), 1 << (PRECISION_BITS - 1), dtype=np.int64)
            for x in range(xmax):
                acc += src[:, xmin + x, :] * int(kk_h[xx, x])
            tmp[:, xx, :] = _clip8(acc)
        cur = tmp
        if bounds_v is not None:
            bounds_v = bounds_v.copy()
            bounds_v[:, 0] -= y0
    if out_h != H:
        out = np.zeros((out_h, cur.shape[1], C), dtype=np.uint8)
        src = cur.astype(np.int64)
        for yy in range(out_h):
            ymin, ymax = bounds_v[yy]
            acc = np.full((cur.shape[1], C), 1 << (PRECISION_BITS - 1), dtype=np.int64)
            for y in range(ymax):
                acc += src[ymin + y] * int(kk_v[yy, y])
            out[yy] = _clip8(acc)
        cur = out
    return cur


def preprocess_for_sr(plate_bgr: np.ndarray, target_size=(192, 32)) -> np.ndarray:
    """inference/run.py:80-96.  plate_bgr uint8 [H, W, 3] (BGR) -> float32 [1, 3, 32, 192] in [0, 1]."""
    rgb = plate_bgr[:, :, ::-1]
    small = pil_bicubic_resize_u8(np.ascontiguousarray(rgb), target_size[0], target_size[1])
    return (small.astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1)[None]
