"""Operator-level entry points of the C ABI (unit parity tests): device fp32 NCHW in / out, same kernels the forward uses."""
from __future__ import annotations

import torch

from . import capi


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def pixel_unshuffle2(x: torch.Tensor) -> torch.Tensor:
    """nn.PixelUnshuffle(2) (reference lpsr.py:72,79) through the NHWC address map of the fused DConv kernels."""
    lib = capi.load_library()
    x = x.float().contiguous()
    B, Cc, H, W = x.shape
    y = torch.empty((B, Cc * 4, H // 2, W // 2), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        capi.check(lib.lpsr_op_pixel_unshuffle2(x.data_ptr(), y.data_ptr(), B, Cc, H, W, _stream(x)), None, "pixel_unshuffle2")
    return y


def pixel_shuffle2(x: torch.Tensor) -> torch.Tensor:
    """nn.PixelShuffle(2) (reference lpsr.py:88,95)."""
    lib = capi.load_library()
    x = x.float().contiguous()
    B, Cc, H, W = x.shape
    y = torch.empty((B, Cc // 4, H * 2, W * 2), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        capi.check(lib.lpsr_op_pixel_shuffle2(x.data_ptr(), y.data_ptr(), B, Cc, H, W, _stream(x)), None, "pixel_shuffle2")
    return y


def conv2d(model, x: torch.Tensor, weight: torch.Tensor, bias=None, relu: bool = False) -> torch.Tensor:
    """nn.Conv2d(k in {1,3}, padding='same') (+ReLU) with the kernel ``model``'s precision mode selects."""
    lib = capi.load_library()
    x = x.float().contiguous()
    weight = weight.float().contiguous()
    B, Cin, H, W = x.shape
    Cout, _, k, _ = weight.shape
    y = torch.empty((B, Cout, H, W), dtype=torch.float32, device=x.device)
    h = model._handle(x.device)
    bptr = bias.float().contiguous().data_ptr() if bias is not None else None
    with torch.cuda.device(x.device):
        torch.cuda.synchronize(x.device)
        capi.check(lib.lpsr_op_conv2d(h, x.data_ptr(), weight.data_ptr(), bptr, y.data_ptr(), B, Cin, Cout, k, H, W,
                                      1 if relu else 0, _stream(x)), h, "lpsr_op_conv2d")
    return y


def preprocess_for_sr_batch(model, plates, target_size=(192, 32), device=None) -> torch.Tensor:
    """Batched, on-device version of the reference's per-plate ``preprocess_for_sr`` (inference/run.py:80-96):
    ``plates`` is a list of BGR uint8 HxWx3 numpy arrays (any sizes, e.g. all plates of a frame); the result is the float32
    ``[B, 3, H_t, W_t]`` CUDA tensor ``torch.cat([preprocess_for_sr(p) for p in plates])`` would give (BGR->RGB, Pillow bicubic
    resize, / 255), bit for bit, ready for ``model(x)``.  One packed H2D copy and one kernel launch for the whole batch."""
    import numpy as np
    lib = capi.load_library()
    dev = torch.device(device) if device is not None else next(model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("preprocess_for_sr_batch needs a CUDA device (there is no CPU fallback)")
    W_t, H_t = int(target_size[0]), int(target_size[1])
    B = len(plates)
    hs, ws, offs, total = [], [], [], 0
    for p in plates:
        if not isinstance(p, np.ndarray) or p.dtype != np.uint8 or p.ndim != 3 or p.shape[2] != 3 or p.shape[0] < 1 or p.shape[1] < 1:
            raise ValueError("plates must be non-empty uint8 HxWx3 (BGR) arrays")
        hs.append(p.shape[0]); ws.append(p.shape[1]); offs.append(total)
        total += p.shape[0] * p.shape[1] * 3
    out = torch.empty((B, 3, H_t, W_t), dtype=torch.float32, device=dev)
    if B == 0:
        return out
    # pinned staging from torch's caching host allocator (reuse is stream-ordered by the allocator itself)
    packed = torch.empty(total, dtype=torch.uint8, pin_memory=True)
    np.concatenate([np.ascontiguousarray(p).reshape(-1) for p in plates], out=packed.numpy())
    h = model._handle(dev)
    with torch.cuda.device(dev):
        crops = packed.to(dev, non_blocking=True)
        a_off = np.asarray(offs, dtype=np.int64); a_h = np.asarray(hs, dtype=np.int32); a_w = np.asarray(ws, dtype=np.int32)
        capi.check(lib.lpsr_preprocess_resize(h, crops.data_ptr(), a_off.ctypes.data, a_h.ctypes.data, a_w.ctypes.data, B, H_t, W_t,
                                              out.data_ptr(), _stream(out)), h, "lpsr_preprocess_resize")
        crops.record_stream(torch.cuda.current_stream(dev))
    return out


def enhance_plates(model, plates, target_size=(192, 32)):
    """The LPSR stage of the reference's per-plate loop (inference/run.py:200-203) for ALL plates of a frame / clip at once:
    ``preprocess_for_sr`` -> ``sr_model(...)`` -> ``.squeeze(0).cpu().permute(1, 2, 0).numpy() * 255`` -> ``astype(np.uint8)``.
    Returns a list of uint8 ``[H_t, W_t, 1]`` arrays (what the reference hands to its OCR step before the colour conversion), using one
    packed H2D copy, one pre-processing launch, one forward and one uint8 D2H copy instead of a sync per plate."""
    x = preprocess_for_sr_batch(model, plates, target_size)
    if x.shape[0] == 0:
        return []
    with torch.no_grad():
        y = model(x)
    # numpy's float32 * 255 followed by astype(uint8) truncates toward zero; the sigmoid output is in (0, 1) so no wrap-around
    u8 = (y * 255.0).to(torch.uint8).permute(0, 2, 3, 1).contiguous().cpu().numpy()
    return [u8[i] for i in range(u8.shape[0])]
