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
