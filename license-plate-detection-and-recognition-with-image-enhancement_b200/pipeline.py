"""Batched, headless LPSR stage of the reference's video pipeline (BASELINE.json configs[4]).

The reference (inference/run.py:153-214) walks the plates of a frame one by one: crop -> straighten -> ``format_long_plate``
(run.py:21-52) -> OCR -> ``preprocess_for_sr`` (80-96) -> ``.to(device)`` -> ``sr_model`` -> ``.cpu()`` (200-202) -> ``*255`` ->
``astype(uint8)`` -> ``cv2.cvtColor(RGB2BGR)`` (203-204) -> ``restack_to_square`` (55-78, 206) -> OCR, with a host<->device round
trip and a stream synchronisation per plate.  Here the plates of a whole frame (or clip) go through ONE pre-processing launch, ONE
LPSR forward and ONE uint8 device->host copy; the steps either side are the same array operations in numpy.

Out of scope: the YOLOv5 plate detector and OCR model themselves -- boxes are INJECTED (the detector weights are not
shipped with the reference, ``.MISSING_LARGE_BLOBS``), and ``straighten_license_plate`` (Canny + Hough, CPU glue) is a caller-supplied
hook that defaults to the identity.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .ops import preprocess_for_sr_batch

Box = Tuple[int, int, int, int]     # x1, y1, x2, y2 in frame pixels (what Detection.detect(bb_scale=True) returns, run.py:175-181)


def format_long_plate(plate: np.ndarray, aspect_ratio_threshold: float = 1.5) -> Tuple[np.ndarray, bool]:
    """Two-row plate -> one row (reference run.py:21-52): a plate whose w/h is at most the threshold is cut at ``h // 2`` and the top
    ``h // 2`` rows are placed left of the bottom ``h // 2`` rows (the middle row of an odd height is dropped).  Returns
    ``(image, changed)``; long plates and empty images come back unchanged."""
    h, w = plate.shape[:2]
    if h == 0 or w == 0 or (w / h) > aspect_ratio_threshold:
        return plate, False
    half = h // 2
    return np.concatenate([plate[:half], plate[h - half:]], axis=1), True


def restack_to_square(long_plate: np.ndarray, aspect_ratio_threshold: float = 1.5) -> np.ndarray:
    """One-row plate -> two rows (reference run.py:55-78): the left ``w // 2`` columns go on top of the right ``w // 2`` columns; an
    image that is already square-ish (w/h below the threshold) or empty is returned as is."""
    h, w = long_plate.shape[:2]
    if h == 0 or w == 0 or (w / h) < aspect_ratio_threshold:
        return long_plate
    half = w // 2
    return np.concatenate([long_plate[:, :half], long_plate[:, w - half:]], axis=0)


def gray_to_bgr(gray_hw1: np.ndarray) -> np.ndarray:
    """``cv2.cvtColor(x, cv2.COLOR_RGB2BGR)`` on the reference's (H, W, 1) uint8 SR output (run.py:204): OpenCV replicates the single
    channel into three, so the result is (H, W, 3) with equal channels."""
    g = gray_hw1.reshape(gray_hw1.shape[0], gray_hw1.shape[1])
    return np.repeat(g[:, :, None], 3, axis=2)


@dataclass
class PlateResult:
    frame: int                       # index of the frame in the clip
    rank: int                        # 1-based rank by box area inside its frame (the reference keeps the 3 largest, run.py:180)
    box: Box
    was_formatted: bool              # format_long_plate changed the layout
    ocr_input: np.ndarray            # what the reference feeds its "before" OCR call (run.py:196-197)
    sr_bgr: np.ndarray               # hr_plate_img of run.py:204: uint8 (H_t, W_t, 3)
    sr_for_ocr: np.ndarray           # what the reference feeds its "after" OCR call (run.py:206)


@dataclass
class StageTimes:
    plates: int = 0
    lpsr_ms: float = 0.0             # CUDA-event time of pre-processing + forward + uint8 conversion (the LPSR stage on the device)
    forward_ms: float = 0.0          # the forward alone
    d2h_ms: float = 0.0
    host_ms: float = 0.0             # crop / format / restack on the host
    calls: int = 0
    per_call: List[Tuple[int, float]] = field(default_factory=list)   # (plates, lpsr_ms)


def select_plates(boxes: Sequence[Box], top_k: int = 3) -> List[Box]:
    """Largest ``top_k`` boxes by area, as the reference does after filtering the class name (run.py:176-180)."""
    return sorted(boxes, key=lambda b: (b[2] - b[0]) * (b[3] - b[1]), reverse=True)[:top_k]


def enhance_frames(model, frames: Sequence[np.ndarray], boxes_per_frame: Sequence[Sequence[Box]], top_k: int = 3,
                   target_size: Tuple[int, int] = (192, 32), straighten: Optional[Callable[[np.ndarray], np.ndarray]] = None,
                   times: Optional[StageTimes] = None) -> List[PlateResult]:
    """The reference's per-frame inner loop (run.py:180-214) for a clip of frames with injected plate boxes, batched over every plate
    of the clip.  Frames are BGR uint8 (as ``cv2.VideoCapture.read`` returns them).  ``times`` (optional) accumulates CUDA-event timings
    of the LPSR stage in isolation (BASELINE configs[4])."""
    import time
    t_host = time.perf_counter()
    meta, long_plates = [], []
    for fi, (frame, boxes) in enumerate(zip(frames, boxes_per_frame)):
        for rank, (x1, y1, x2, y2) in enumerate(select_plates(boxes, top_k), start=1):
            raw = frame[int(y1):int(y2), int(x1):int(x2)]
            if raw.size == 0:                                        # run.py:189
                continue
            straight = straighten(raw) if straighten is not None else raw
            if straight.size == 0:                                   # run.py:192
                continue
            long_img, changed = format_long_plate(straight)
            ocr_in = restack_to_square(long_img) if changed else long_img.copy()
            meta.append((fi, rank, (int(x1), int(y1), int(x2), int(y2)), changed, ocr_in))
            long_plates.append(np.ascontiguousarray(long_img))
    host_s = time.perf_counter() - t_host
    if not long_plates:
        return []
    dev = next(model.parameters()).device
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.no_grad():
        ev[0].record()
        x = preprocess_for_sr_batch(model, long_plates, target_size)          # run.py:201 for all plates: one H2D copy, one launch
        ev[1].record()
        y = model(x)                                                          # run.py:202
        ev[2].record()
        u8_dev = (y * 255.0).to(torch.uint8).permute(0, 2, 3, 1).contiguous()   # run.py:203: float32 * 255 -> astype(uint8) truncates
        ev[3].record()
        t0 = time.perf_counter()
        u8 = u8_dev.cpu().numpy()
        d2h_s = time.perf_counter() - t0
    t_host = time.perf_counter()
    out = []
    for (fi, rank, box, changed, ocr_in), g in zip(meta, u8):
        bgr = gray_to_bgr(g)                                                  # run.py:204
        out.append(PlateResult(fi, rank, box, changed, ocr_in, bgr, restack_to_square(bgr) if changed else bgr.copy()))   # run.py:206
    host_s += time.perf_counter() - t_host
    if times is not None:
        torch.cuda.synchronize(dev)
        times.plates += len(out)
        times.calls += 1
        lp = ev[0].elapsed_time(ev[3])
        times.lpsr_ms += lp
        times.forward_ms += ev[1].elapsed_time(ev[2])
        times.d2h_ms += 1e3 * d2h_s
        times.host_ms += 1e3 * host_s
        times.per_call.append((len(out), lp))
    return out


# ---- synthetic clip for the harness (no plate images ship with the reference) ------------------------------------------------------
def synthetic_clip(n_frames: int, plates_per_frame: int = 3, size: Tuple[int, int] = (1080, 1920), seed: int = 0):
    """1080p BGR frames with plate-like rectangles (bright background, dark glyph-like bars, one- and two-row aspect ratios) pasted at
    random positions, and the boxes a detector would return for them (run.py:175 with ``bb_scale=True``)."""
    rng = np.random.default_rng(seed)
    H, W = size
    frames, boxes = [], []
    for _ in range(n_frames):
        frame = rng.integers(0, 96, (H, W, 3), dtype=np.uint8)
        bxs = []
        for k in range(plates_per_frame):
            two_row = bool(rng.integers(0, 2))
            ph = int(rng.integers(36, 80)) if two_row else int(rng.integers(18, 48))
            pw = int(ph * rng.uniform(1.15, 1.45)) if two_row else int(ph * rng.uniform(3.2, 5.0))
            x1 = int(rng.integers(0, W - pw)); y1 = int(rng.integers(0, H - ph))
            plate = np.full((ph, pw, 3), 235, dtype=np.uint8)
            rows = 2 if two_row else 1
            for r in range(rows):
                y0 = int((r + 0.15) * ph / rows); y2 = int((r + 0.85) * ph / rows)
                n_glyph = int(rng.integers(4, 8))
                for gi in range(n_glyph):
                    gx = int((gi + 0.2) * pw / n_glyph); gw = max(1, int(0.5 * pw / n_glyph))
                    plate[y0:y2, gx:gx + gw] = int(rng.integers(10, 60))
            noise = rng.normal(0, 6, plate.shape)
            frame[y1:y1 + ph, x1:x1 + pw] = np.clip(plate + noise, 0, 255).astype(np.uint8)
            bxs.append((x1, y1, x1 + pw, y1 + ph))
        frames.append(frame)
        boxes.append(bxs)
    return frames, boxes
