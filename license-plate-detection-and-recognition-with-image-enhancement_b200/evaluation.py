"""OCR-quality harness around the LPSR stage: the metric logic of the reference's ``evaluation/eval.py`` with the
super-resolution path batched through the B200 forward.

The reference (evaluation/eval.py:96-160) walks a folder whose file names are the ground-truth plate strings and, per image, runs its
YOLOv5 OCR twice -- on the raw crop (eval.py:112) and on the super-resolved crop (eval.py:115-122: ``preprocess_for_sr`` -> ``sr_model`` ->
``clip(0, 1)`` -> ``* 255`` -> ``astype(uint8)`` -> BGR) -- and reports exact-match accuracy and the character error rate (CER, Levenshtein
distance / len(ground truth), eval.py:28-33) with and without SR plus their differences (eval.py:135-160).

Here:
  * ``levenshtein`` is a pure-Python restatement of the edit distance the reference takes from the ``Levenshtein`` package (absent from this
    image; unit insert / delete / substitute costs), ``calculate_cer`` / ``ground_truth_from_filename`` mirror eval.py:22-33;
  * ``sort_license_plate_detections`` mirrors my_utils/utils.py:7-72 (reading order of character boxes: rows by centre-y gaps, then left to right),
    ``detections_to_text`` the string assembly of eval.py:49-53;
  * ``super_resolve_batch`` is the SR path of eval.py:115-120 for ALL images at once (one pre-processing launch, one forward, one uint8 copy);
  * ``evaluate`` produces the reference's report numbers.  The OCR network itself (YOLOv5 ``weights/char.pt``) is NOT part
    of this repository: the caller passes ``ocr(image_bgr) -> str`` (or ``-> detections``), e.g. a wrapper of the reference's ``Detection``;
  * ``detections_from_predictions`` / ``texts_from_predictions`` are the half of row n4 that is: everything ``Detection.char_detection_yolo``
    does AFTER the network (my_models/detection.py:44-71) -- ``non_max_suppression`` on the GPU for all plates at once
    (``lpsr_b200.non_max_suppression``, bit-identical to yolov5/utils/general.py:677-760), the detection rows, the reading-order text.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, Dict, List, Sequence, Tuple, Union

import numpy as np
import torch

from .ops import non_max_suppression, preprocess_for_sr_batch

Detection = Tuple[str, float, Tuple[float, float, float, float]]     # (class name, confidence, (x1, y1, x2, y2)): Detection.detect's rows


def ground_truth_from_filename(filename: str) -> str:
    """eval.py:22-24: the plate text is the upper-cased file name without its extension."""
    return os.path.splitext(os.path.basename(filename))[0].upper()


def levenshtein(a: str, b: str) -> int:
    """Edit distance with unit insert / delete / substitute costs (what ``Levenshtein.distance`` returns, eval.py:32)."""
    if a == b:
        return 0
    if len(a) < len(b):
        a, b = b, a
    if not b:
        return len(a)
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def calculate_cer(ground_truth: str, ocr_result: str) -> float:
    """Character error rate, eval.py:28-33 (an empty ground truth counts 1.0 when the OCR returned anything, else 0.0)."""
    if not ground_truth:
        return 1.0 if ocr_result else 0.0
    return levenshtein(ground_truth, ocr_result) / len(ground_truth)


def sort_license_plate_detections(detections: Sequence[Detection]) -> List[Detection]:
    """Reading order of character detections (my_utils/utils.py:7-72): sort by centre y, start a new row wherever consecutive centres are
    more than ``max(10, (y_last - y_first) / 5)`` apart, sort each row by centre x, concatenate the rows."""
    if not detections:
        return []
    items = [((d[2][0] + d[2][2]) / 2, (d[2][1] + d[2][3]) / 2, d) for d in detections]
    items.sort(key=lambda t: t[1])
    thr = max(10, (items[-1][1] - items[0][1]) / 5)
    rows, cur = [], [items[0]]
    for prev, it in zip(items, items[1:]):
        if abs(it[1] - prev[1]) > thr:
            rows.append(cur)
            cur = [it]
        else:
            cur.append(it)
    rows.append(cur)
    out: List[Detection] = []
    for row in rows:
        row.sort(key=lambda t: t[0])
        out.extend(t[2] for t in row)
    return out


def detections_to_text(detections: Sequence[Detection]) -> str:
    """eval.py:49-53: class names of the sorted detections, upper-cased and concatenated."""
    return "".join(str(d[0]).upper() for d in sort_license_plate_detections(detections))


def detections_from_predictions(prediction: torch.Tensor, names: Sequence[str], conf_thres: float, iou_thres: float, classes=None,
                                agnostic_nms: bool = True, max_det: int = 1000) -> List[List[Detection]]:
    """``Detection.char_detection_yolo`` from the network output on (my_models/detection.py:44-71, ``bb_scale=False``) for a whole batch:
    ``prediction`` = the YOLOv5 head's output ``model(img)[0]``, CUDA float32 [B, N, 5 + nc].  One NMS launch, one device-to-host copy;
    rows are ``(names[int(cls)], conf, (x1, y1, x2, y2))`` in the reference's order (the reference keeps ``str(conf)``; the float is kept here)."""
    dets = non_max_suppression(prediction, conf_thres=conf_thres, iou_thres=iou_thres, classes=classes, agnostic=agnostic_nms,
                               multi_label=True, max_det=max_det)
    out: List[List[Detection]] = []
    for d in dets:
        rows = d.cpu().tolist()
        out.append([(names[int(r[5])], r[4], (r[0], r[1], r[2], r[3])) for r in rows])
    return out


def texts_from_predictions(prediction: torch.Tensor, names: Sequence[str], conf_thres: float, iou_thres: float, **kw) -> List[str]:
    """Plate strings of a batch of OCR predictions: ``detections_from_predictions`` + the reading-order assembly of eval.py:49-53."""
    return [detections_to_text(d) for d in detections_from_predictions(prediction, names, conf_thres, iou_thres, **kw)]


def super_resolve_batch(model, images_bgr: Sequence[np.ndarray], target_size: Tuple[int, int] = (192, 32)) -> List[np.ndarray]:
    """The SR path of eval.py:115-120 for a whole list of BGR crops: ``preprocess_for_sr`` (bit-exact batch version) -> ``sr_model`` ->
    ``np.clip(0, 1)`` -> ``* 255`` -> ``astype(np.uint8)`` -> 3-channel BGR uint8 ``[H_t, W_t, 3]`` per image.  (eval.py:120 asks cv2 for
    COLOR_RGB2BGR on the model's one-channel output, which cv2 rejects; inference/run.py:204 uses COLOR_GRAY2BGR for the same step -- the grey
    value replicated to three channels -- and that is what is produced here.)"""
    x = preprocess_for_sr_batch(model, list(images_bgr), target_size)
    if x.shape[0] == 0:
        return []
    with torch.no_grad():
        y = model(x)
    u8 = (y.clamp(0.0, 1.0) * 255.0).to(torch.uint8).permute(0, 2, 3, 1).contiguous().cpu().numpy()   # truncation, as numpy's astype
    return [np.repeat(u8[i], 3, axis=2) for i in range(u8.shape[0])]


@dataclass
class EvalReport:
    """The numbers of the reference's final report (eval.py:135-160); percentages as printed there."""
    num_images: int
    exact_without_sr: int
    exact_with_sr: int
    accuracy_without_sr: float
    accuracy_with_sr: float
    cer_without_sr: float
    cer_with_sr: float
    accuracy_improvement: float
    cer_reduction: float
    rows: List[Dict[str, str]]

    def as_dict(self) -> Dict[str, object]:
        return {k: getattr(self, k) for k in ("num_images", "exact_without_sr", "exact_with_sr", "accuracy_without_sr", "accuracy_with_sr",
                                              "cer_without_sr", "cer_with_sr", "accuracy_improvement", "cer_reduction")}


def _to_text(r: Union[str, Sequence[Detection]]) -> str:
    return r if isinstance(r, str) else detections_to_text(r)


def evaluate(model, samples: Sequence[Tuple[str, np.ndarray]], ocr: Callable[[np.ndarray], Union[str, Sequence[Detection]]],
             target_size: Tuple[int, int] = (192, 32)) -> EvalReport:
    """eval.py:96-160 over ``samples`` = (file name or ground-truth string, BGR uint8 crop).  ``ocr`` is called on a COPY of the raw crop and
    of the super-resolved crop (eval.py:112,122) and returns the plate text or the detection rows; the SR crops come from one batched pass."""
    samples = [(n, im) for n, im in samples if im is not None and getattr(im, "size", 0) > 0]     # eval.py:107-109: unreadable images are skipped
    n = len(samples)
    sr_images = super_resolve_batch(model, [im for _, im in samples], target_size) if n else []
    tot = {"without": [0.0, 0], "with": [0.0, 0]}
    rows = []
    for (name, raw), sr in zip(samples, sr_images):
        gt = ground_truth_from_filename(name)
        t_raw, t_sr = _to_text(ocr(raw.copy())), _to_text(ocr(sr.copy()))
        for key, txt in (("without", t_raw), ("with", t_sr)):
            tot[key][0] += calculate_cer(gt, txt)
            tot[key][1] += int(gt == txt)
        rows.append({"file": name, "gt": gt, "ocr_original": t_raw, "ocr_sr": t_sr})
    d = max(n, 1)
    acc0, acc1 = tot["without"][1] / d * 100, tot["with"][1] / d * 100
    cer0, cer1 = tot["without"][0] / d * 100, tot["with"][0] / d * 100
    return EvalReport(n, tot["without"][1], tot["with"][1], acc0, acc1, cer0, cer1, acc1 - acc0, cer0 - cer1, rows)
