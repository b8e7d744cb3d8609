"""SURVEY 8f row n4 (post-processing half): YOLOv5 `non_max_suppression` around the OCR model.
CPU: the numpy oracle against vectors generated from the unmodified reference (tests/golden/make_golden_nms.py).
GPU: `lpsr_b200.non_max_suppression` (one CUDA kernel behind lpsr_op_yolo_nms) against the same vectors and against the oracle on
seeded random predictions -- index / compare work in the reference's float32 operation order, so the bar is bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import nms_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nms_cases.npz")


def _cases():
    d = np.load(GOLD)
    names = sorted({k.rsplit(".", 1)[0] for k in d.files})
    return d, names


D, NAMES = _cases()


def _generator():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_nms", os.path.join(os.path.dirname(GOLD), "make_golden_nms.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)          # importing it does not touch /root/reference (only main() does)
    return mod


def _args(name):
    conf, iou, agn, ml, md = D[name + ".args"]
    cls = D[name + ".classes"]
    return dict(conf_thres=float(conf), iou_thres=float(iou), agnostic=bool(agn), multi_label=bool(ml), max_det=int(md),
                classes=[int(c) for c in cls] if cls.size else None)


def _split(name):
    out, counts = D[name + ".out"], D[name + ".counts"]
    o = np.cumsum(np.concatenate(([0], counts)))
    return [out[o[i]:o[i + 1]] for i in range(len(counts))]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_vectors(name):
    got = orc.non_max_suppression(D[name + ".pred"], **_args(name))
    want = _split(name)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.shape == w.shape and g.dtype == np.float32
        assert np.array_equal(g, w)                      # bit-exact: same float32 operations in the same order


def test_oracle_nms_tie_order_and_empty():
    boxes = np.array([[0, 0, 10, 10], [0, 0, 10, 10], [20, 20, 30, 30], [0, 0, 0, 0], [0, 0, 0, 0]], np.float32)
    scores = np.array([0.5, 0.5, 0.5, 0.9, 0.9], np.float32)
    # the two empty boxes have IoU 0/0 = NaN (never suppressed); among equal scores the lower index comes first
    assert orc.nms(boxes, scores, 0.5).tolist() == [3, 4, 0, 2]
    assert orc.nms(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 0.5).shape == (0,)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_nms_matches_reference_vectors(name):
    import lpsr_b200
    pred = torch.from_numpy(D[name + ".pred"]).cuda()
    keep = pred.clone()
    got = lpsr_b200.non_max_suppression(pred, **_args(name))
    assert torch.equal(pred, keep)                       # the input is not modified
    want = _split(name)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert tuple(g.shape) == w.shape and g.dtype == torch.float32 and g.is_cuda
        assert np.array_equal(g.cpu().numpy(), w)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,B,N,nc,kw", [
    (1, 5, 512, 36, dict(conf_thres=0.25, iou_thres=0.3, agnostic=True, multi_label=True, max_det=1000)),
    (2, 3, 3000, 3, dict(conf_thres=0.3, iou_thres=0.6, agnostic=False, multi_label=True, max_det=300)),
    (3, 2, 1000, 80, dict(conf_thres=0.4, iou_thres=0.45, agnostic=False, multi_label=False, max_det=50)),
    (4, 96, 512, 36, dict(conf_thres=0.25, iou_thres=0.3, agnostic=True, multi_label=True, max_det=1000)),   # a clip's worth of plates
    (5, 1, 1, 1, dict(conf_thres=0.0, iou_thres=0.0, agnostic=True, multi_label=True, max_det=1)),
])
def test_gpu_nms_matches_oracle_random(seed, B, N, nc, kw):
    import lpsr_b200
    synthetic_pred = _generator().synthetic_pred
    rng = np.random.default_rng(seed)
    pred = synthetic_pred(rng, B, N, nc, min(N // 12, 24), ties=(seed % 2 == 0), tiny=(seed == 3))
    want = orc.non_max_suppression(pred, **kw)
    got = lpsr_b200.non_max_suppression(torch.from_numpy(pred).cuda(), **kw)
    for g, w in zip(got, want):
        assert np.array_equal(g.cpu().numpy(), w)


@pytest.mark.gpu
def test_gpu_nms_rejects_bad_arguments_and_overflow():
    import lpsr_b200
    p = torch.rand(1, 64, 9, device="cuda")
    with pytest.raises(ValueError):
        lpsr_b200.non_max_suppression(p, conf_thres=1.5)
    with pytest.raises(ValueError):
        lpsr_b200.non_max_suppression(p[0], conf_thres=0.5)
    with pytest.raises(ValueError):
        lpsr_b200.non_max_suppression(p.cpu(), conf_thres=0.5)
    many = torch.ones(1, 600, 9, device="cuda")
    many[..., 2:4] = 10.0
    with pytest.raises(RuntimeError):                     # 600 x 4 candidates > capacity 1024: reported, never truncated silently
        lpsr_b200.non_max_suppression(many, conf_thres=0.1, multi_label=True, max_candidates=1024)
    assert lpsr_b200.non_max_suppression(torch.zeros(0, 10, 9, device="cuda")) == []
    assert [tuple(t.shape) for t in lpsr_b200.non_max_suppression(torch.zeros(2, 0, 9, device="cuda"))] == [(0, 6), (0, 6)]


@pytest.mark.gpu
def test_detection_rows_and_texts_from_ocr_predictions():
    """my_models/detection.py:44-71 after the network, on predictions of the reference's own OCR model (weights/char.pt) for four synthetic
    plates: rows (name, conf, box) equal the reference's NMS output mapped through the class names, and the plate strings follow."""
    from lpsr_b200 import evaluation as ev
    name = "char_pt_call_site_conf0.25"
    names = list("0123456789abcdefghijklmnopqrstuvwxyz")
    a = _args(name)
    pred = torch.from_numpy(D[name + ".pred"]).cuda()
    rows = ev.detections_from_predictions(pred, names, a["conf_thres"], a["iou_thres"])
    want = _split(name)
    assert [len(r) for r in rows] == [w.shape[0] for w in want] and sum(len(r) for r in rows) > 0
    for r, w in zip(rows, want):
        for (nm, conf, box), ref in zip(r, w):
            assert nm == names[int(ref[5])] and np.float32(conf) == ref[4] and np.array_equal(np.asarray(box, np.float32), ref[:4])
    texts = ev.texts_from_predictions(pred, names, a["conf_thres"], a["iou_thres"])
    assert texts == [ev.detections_to_text([(names[int(x[5])], float(x[4]), tuple(float(v) for v in x[:4])) for x in w]) for w in want]
    assert all(t == t.upper() and len(t) == w.shape[0] for t, w in zip(texts, want))
