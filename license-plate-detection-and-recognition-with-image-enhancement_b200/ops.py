"""Operator-level entry points of the C ABI (unit parity tests): device fp32 NCHW in / out, same kernels the forward uses."""
from __future__ import annotations

import numpy as np
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


def non_max_suppression(prediction: torch.Tensor, conf_thres: float = 0.25, iou_thres: float = 0.45, classes=None, agnostic: bool = False,
                        multi_label: bool = False, max_det: int = 300, max_candidates: int = 2048):
    """The reference's ``non_max_suppression`` (yolov5/utils/general.py:677-760; call site my_models/detection.py:44-54) for a whole batch
    of images in one CUDA launch: ``prediction`` is the YOLOv5 head's output, a CUDA float32 tensor [B, N, 5 + nc]; returns a list of B CUDA
    tensors [n, 6] (x1, y1, x2, y2, conf, cls), bit-identical to the reference on CPU (same order of the kept boxes).  Same argument names
    and checks as the reference; the input tensor is NOT modified (the reference zeroes the objectness of out-of-range boxes in place).
    ``max_candidates`` bounds the detections per image BEFORE suppression the workspace holds (doubled automatically up to 8192)."""
    if not (0 <= conf_thres <= 1):
        raise ValueError(f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0")
    if not (0 <= iou_thres <= 1):
        raise ValueError(f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0")
    if prediction.dim() != 3 or prediction.shape[2] < 6:
        raise ValueError("prediction must be [B, N, 5 + nc]")
    if not prediction.is_cuda or prediction.dtype != torch.float32:
        raise ValueError("prediction must be a CUDA float32 tensor (there is no CPU path)")
    B, N, row = prediction.shape
    if B == 0:
        return []
    if N == 0:
        return [prediction.new_zeros((0, 6)) for _ in range(B)]
    pred = prediction.contiguous()
    lib = capi.load_library()
    dev = pred.device
    cls = np.ascontiguousarray(np.asarray(classes if classes is not None else [], dtype=np.int32))
    cap = max(64, (int(max_candidates) + 63) // 64 * 64)
    with torch.cuda.device(dev):
        while True:
            out = torch.empty(B, max_det, 6, dtype=torch.float32, device=dev)
            cnt = torch.empty(B, dtype=torch.int32, device=dev)
            nbytes = lib.lpsr_op_yolo_nms_workspace_bytes(B, cap, int(max_det), int(cls.size))
            if nbytes == 0:
                raise ValueError("bad max_candidates / max_det")
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            capi.check(lib.lpsr_op_yolo_nms(None, pred.data_ptr(), B, N, row - 5, float(conf_thres), float(iou_thres),
                                            cls.ctypes.data if cls.size else None, int(cls.size), 1 if agnostic else 0, 1 if multi_label else 0,
                                            int(max_det), cap, out.data_ptr(), cnt.data_ptr(), ws.data_ptr(), nbytes, _stream(pred)),
                       None, "lpsr_op_yolo_nms")
            counts = cnt.cpu().tolist()                        # the one synchronisation: the reference returns ragged tensors
            if min(counts) >= 0:
                break
            if cap >= 8192:
                raise RuntimeError(f"more than {cap} detections before suppression in one image: raise conf_thres (the reference cuts at 30000 "
                                   "by an unstable argsort; this path reports instead of truncating)")
            if cap >= (int(max_candidates) + 63) // 64 * 64 and int(max_candidates) < 2048:
                raise RuntimeError(f"more than max_candidates = {max_candidates} detections before suppression in one image")
            cap = min(8192, cap * 2)
    return [out[b, :counts[b]] for b in range(B)]
