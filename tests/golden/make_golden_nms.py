"""Generates tests/golden/nms_cases.npz: inputs and outputs of the UNMODIFIED reference post-processing
(/root/reference/yolov5/utils/general.py:677-760 `non_max_suppression`, which calls torchvision.ops.nms) for
 * predictions of the reference's own OCR model (/root/reference/weights/char.pt, 128 x 128 input, my_models/detection.py:40-54) on
   synthetic plate images, with the call site's arguments (agnostic, multi-label, max_det = 1000, IoU 0.3), and
 * synthetic stress predictions (clustered boxes, tied scores, out-of-range sizes, class offsets, class filter, max_det cut).
Run in the build container (needs /root/reference; matplotlib / seaborn are absent here and only used by the reference's plotting
module, so they are stubbed before the import -- the function under test is untouched):  python tests/golden/make_golden_nms.py
"""
import importlib, os, sys, types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nms_cases.npz")


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Stub(self.__name__ + "." + k)

    def __call__(self, *a, **k):
        return None


def reference():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.font_manager", "seaborn"):
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _Stub(name)
    sys.path.insert(0, os.path.join(REF, "yolov5"))
    import utils  # noqa: F401  (yolov5/utils)
    sys.modules["utils.plots"] = _Stub("utils.plots")     # its class body downloads a font at import time
    from utils.general import non_max_suppression
    import models.experimental  # noqa: F401  (import order the reference's attempt_load uses)
    ck = torch.load(os.path.join(REF, "weights", "char.pt"), map_location="cpu", weights_only=False)
    model = ck["model"].float().eval()
    for m in model.modules():
        if type(m).__name__ == "Upsample" and not hasattr(m, "recompute_scale_factor"):
            m.recompute_scale_factor = None
    return non_max_suppression, model


def plate_images(rng, n):
    """n synthetic 128 x 128 RGB plates: light background, dark glyph-like strokes in two rows."""
    imgs = np.empty((n, 3, 128, 128), np.float32)
    for k in range(n):
        img = np.full((128, 128), 0.85 + 0.1 * rng.random(), np.float32)
        for row in range(2):
            y0 = 18 + 56 * row
            for c in range(4 + int(rng.integers(0, 2))):
                x0 = 8 + 24 * c + int(rng.integers(-2, 3))
                img[y0:y0 + 40, x0:x0 + 4] = 0.1
                img[y0:y0 + 4, x0:x0 + 16] = 0.1
                if rng.random() < 0.6:
                    img[y0 + 18:y0 + 22, x0:x0 + 16] = 0.1
                if rng.random() < 0.6:
                    img[y0:y0 + 40, x0 + 12:x0 + 16] = 0.1
                if rng.random() < 0.5:
                    img[y0 + 36:y0 + 40, x0:x0 + 16] = 0.1
        img += 0.03 * rng.standard_normal(img.shape).astype(np.float32)
        imgs[k] = np.clip(img, 0, 1)[None]
    return imgs


def synthetic_pred(rng, B, N, nc, n_clusters, ties=False, tiny=False):
    """YOLO-shaped predictions: N anchors, a few clusters of strongly overlapping boxes, the rest low objectness."""
    p = np.zeros((B, N, 5 + nc), np.float32)
    for b in range(B):
        p[b, :, 0:2] = rng.uniform(0, 128, (N, 2))
        p[b, :, 2:4] = rng.uniform(4, 40, (N, 2))
        p[b, :, 4] = rng.uniform(0, 0.2, N)
        p[b, :, 5:] = rng.uniform(0, 0.3, (N, nc))
        idx = rng.permutation(N)
        per = 12
        for c in range(n_clusters):
            rows = idx[c * per:(c + 1) * per]
            centre = rng.uniform(10, 118, 2)
            size = rng.uniform(10, 30, 2)
            p[b, rows, 0:2] = centre + rng.normal(0, 1.5, (per, 2))
            p[b, rows, 2:4] = size + rng.normal(0, 1.0, (per, 2))
            p[b, rows, 4] = rng.uniform(0.5, 1.0, per)
            cls = int(rng.integers(0, nc))
            p[b, rows, 5 + cls] = rng.uniform(0.5, 1.0, per)
            if nc > 1:
                p[b, rows, 5 + (cls + 1) % nc] = rng.uniform(0.2, 0.9, per)
        if ties:      # quantised scores: many exact ties, the stable sort's order matters
            p[b, :, 4] = np.round(p[b, :, 4] * 8) / 8
            p[b, :, 5:] = np.round(p[b, :, 5:] * 4) / 4
        if tiny:      # out-of-range sizes (min_wh / max_wh) and empty boxes
            p[b, idx[:40], 2] = rng.choice([0.0, 1.0, 1.99, 8000.0], 40)
            p[b, idx[:40], 4] = 0.9
    return p.astype(np.float32)


def main():
    nms_ref, model = reference()
    rng = np.random.default_rng(20260101)
    cases = {}

    def add(name, pred, **kw):
        args = dict(conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False, max_det=300)
        args.update(kw)
        out = nms_ref(torch.from_numpy(pred.copy()), **args)          # the reference modifies its input: give it a copy
        cases[name + ".pred"] = pred
        cases[name + ".args"] = np.array([args["conf_thres"], args["iou_thres"], float(args["agnostic"]), float(args["multi_label"]),
                                          float(args["max_det"])], np.float64)
        cases[name + ".classes"] = np.asarray(args["classes"] if args["classes"] is not None else [], np.int64)
        cases[name + ".counts"] = np.array([o.shape[0] for o in out], np.int64)
        cases[name + ".out"] = (torch.cat(list(out), 0).numpy().astype(np.float32) if sum(o.shape[0] for o in out)
                                else np.zeros((0, 6), np.float32))
        print(f"{name:28s} pred {pred.shape}  detections per image {[o.shape[0] for o in out]}")

    with torch.no_grad():
        pred = model(torch.from_numpy(plate_images(rng, 4)))[0].numpy().astype(np.float32)
    # the OCR call site (detection.py:44-54; run.py:112 --iou 0.3); two confidence levels (random glyphs score low with the trained model)
    add("char_pt_call_site_conf0.25", pred, conf_thres=0.25, iou_thres=0.3, agnostic=True, multi_label=True, max_det=1000)
    add("char_pt_call_site_conf0.02", pred, conf_thres=0.02, iou_thres=0.3, agnostic=True, multi_label=True, max_det=1000)
    add("char_pt_best_class_conf0.01", pred, conf_thres=0.01, iou_thres=0.3, agnostic=False, multi_label=False, max_det=1000)
    add("synthetic_multilabel", synthetic_pred(rng, 3, 512, 36, 10), conf_thres=0.25, iou_thres=0.3, agnostic=True, multi_label=True, max_det=1000)
    add("synthetic_class_offsets", synthetic_pred(rng, 2, 512, 36, 12), conf_thres=0.25, iou_thres=0.45, agnostic=False, multi_label=True)
    add("synthetic_ties", synthetic_pred(rng, 2, 768, 8, 20, ties=True), conf_thres=0.25, iou_thres=0.5, agnostic=True, multi_label=True, max_det=1000)
    add("synthetic_out_of_range", synthetic_pred(rng, 2, 512, 4, 8, tiny=True), conf_thres=0.1, iou_thres=0.45)
    add("synthetic_class_filter", synthetic_pred(rng, 2, 512, 36, 14), conf_thres=0.25, iou_thres=0.45, classes=[0, 3, 7, 12, 35], multi_label=True)
    add("synthetic_max_det_cut", synthetic_pred(rng, 1, 1024, 2, 60), conf_thres=0.2, iou_thres=0.9, agnostic=True, multi_label=True, max_det=25)
    add("synthetic_single_class", synthetic_pred(rng, 2, 256, 1, 6), conf_thres=0.25, iou_thres=0.45, multi_label=True)
    add("nothing_passes", synthetic_pred(rng, 2, 128, 5, 0), conf_thres=0.9, iou_thres=0.45)
    np.savez_compressed(OUT, **cases)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
