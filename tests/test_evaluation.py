"""OCR-quality harness (lpsr_b200/evaluation.py) against vectors generated from the unmodified reference
(tests/golden/make_golden_eval.py: evaluation/eval.py:22-33, my_utils/utils.py:7-72), plus the batched SR path of eval.py:115-120 on the GPU."""
import json
import os

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from conftest import GOLDEN
from lpsr_b200 import evaluation as ev

CASES = json.load(open(os.path.join(GOLDEN, "eval_cases.json")))


def test_levenshtein_and_cer_match_reference_vectors():
    for c in CASES["cer"]:
        assert ev.levenshtein(c["gt"], c["ocr"]) == c["distance"], c
        assert ev.calculate_cer(c["gt"], c["ocr"]) == pytest.approx(c["cer"], abs=0, rel=0), c
    assert ev.levenshtein("KITTEN", "SITTING") == 3 and ev.levenshtein("SATURDAY", "SUNDAY") == 3 and ev.levenshtein("", "ABC") == 3
    assert ev.calculate_cer("", "") == 0.0 and ev.calculate_cer("", "X") == 1.0      # eval.py:30-31


@settings(max_examples=200, deadline=None)
@given(st.text("AB1", max_size=8), st.text("AB1", max_size=8), st.text("AB1", max_size=8))
def test_levenshtein_is_a_metric(a, b, c):
    d = ev.levenshtein
    assert d(a, a) == 0 and d(a, b) == d(b, a) and (d(a, b) == 0) == (a == b)
    assert d(a, c) <= d(a, b) + d(b, c)
    assert abs(len(a) - len(b)) <= d(a, b) <= max(len(a), len(b))


def test_ground_truth_from_filename_matches_reference():
    for c in CASES["names"]:
        assert ev.ground_truth_from_filename(c["name"]) == c["gt"]


def test_sort_license_plate_detections_matches_reference():
    for c in CASES["sort"]:
        dets = [(cls, conf, tuple(box)) for cls, conf, box in c["detections"]]
        out = ev.sort_license_plate_detections(dets)
        assert [dets.index(o) if dets.count(o) == 1 else None for o in out] == [i if dets.count(dets[i]) == 1 else None for i in c["order"]]
        assert ev.detections_to_text(dets) == c["text"]
    assert ev.sort_license_plate_detections([]) == []


def test_report_arithmetic_with_stub_sr_and_ocr(monkeypatch):
    """eval.py:96-160 with the SR pass and the OCR replaced by stubs: the report's counts, percentages and differences."""
    raws = {"51F12345.jpg": np.full((20, 60, 3), 1, np.uint8), "30A99999.png": np.full((18, 70, 3), 2, np.uint8),
            "29H12345.jpeg": np.full((25, 80, 3), 3, np.uint8), "unreadable.png": None}
    monkeypatch.setattr(ev, "super_resolve_batch", lambda model, imgs, target_size=(192, 32): [np.full((32, 192, 3), int(im[0, 0, 0]) + 100, np.uint8) for im in imgs])
    table = {1: "51F1234", 101: "51F12345", 2: "30A99999", 102: "30A99999", 3: "Z9H1234", 103: "29H12845"}

    def ocr(img):
        assert img.flags.writeable
        return table[int(img[0, 0, 0])]
    r = ev.evaluate(None, list(raws.items()), ocr)
    assert r.num_images == 3 and r.exact_without_sr == 1 and r.exact_with_sr == 2
    assert r.accuracy_without_sr == pytest.approx(100 / 3) and r.accuracy_with_sr == pytest.approx(200 / 3)
    assert r.cer_without_sr == pytest.approx((1 / 8 + 0 + 2 / 8) / 3 * 100) and r.cer_with_sr == pytest.approx((0 + 0 + 1 / 8) / 3 * 100)
    assert r.accuracy_improvement == pytest.approx(100 / 3) and r.cer_reduction == pytest.approx(r.cer_without_sr - r.cer_with_sr)
    assert [row["gt"] for row in r.rows] == ["51F12345", "30A99999", "29H12345"]
    # detection rows instead of strings go through the reference's reading-order sort
    det = ev.evaluate(None, [("AB.png", raws["51F12345.jpg"])], lambda img: [("b", 0.9, (30, 0, 40, 10)), ("a", 0.8, (0, 1, 10, 11))])
    assert det.rows[0]["ocr_original"] == "AB" and det.exact_without_sr == 1
    assert ev.evaluate(None, [], ocr).num_images == 0


@pytest.mark.gpu
def test_super_resolve_batch_matches_reference_recipe(shipped_weights):
    """The SR path of eval.py:115-120 (per image, reference arithmetic on CPU) against the batched GPU pass, fp32 mode: at most one level on
    at most 1 % of the pixels; fp16 (the benchmarked mode): mean difference below one level."""
    import lpsr_b200
    from oracle import lpsr_torch_port as port, preprocess_oracle as pre
    rng = np.random.default_rng(7)
    imgs = []
    for h, w in [(20, 64), (33, 120), (16, 50), (48, 200), (24, 96)]:
        lo = rng.integers(0, 256, (max(h // 4, 2), max(w // 4, 2), 3), dtype=np.uint8)
        imgs.append(np.ascontiguousarray(np.kron(lo, np.ones((4, 4, 1), np.uint8))[:h, :w]))
    Wt = port.to_torch_weights(shipped_weights)
    refs = []
    for im in imgs:
        y = port.lpsr_forward(torch.from_numpy(pre.preprocess_for_sr(im)), Wt).squeeze(0).permute(1, 2, 0).numpy()
        refs.append(np.repeat((np.clip(y, 0, 1) * 255).astype(np.uint8), 3, axis=2))
    for prec in ("fp32", "fp16"):
        m = lpsr_b200.LPSR(3, 32, 16, 4, 4, None, precision=prec).load_live_weights(shipped_weights).to("cuda:0").eval()
        out = ev.super_resolve_batch(m, imgs)
        assert len(out) == len(imgs)
        for a, b in zip(out, refs):
            assert a.shape == b.shape == (32, 192, 3) and a.dtype == np.uint8
            d = np.abs(a.astype(np.int32) - b.astype(np.int32))
            if prec == "fp32":
                assert d.max() <= 1 and float((d != 0).mean()) <= 0.01
            else:
                assert float(d.mean()) <= 0.75
        # the report runs end to end on the GPU path with a deterministic stand-in OCR (mean brightness of the left / right half as two "characters")
        ocr = lambda img: f"{int(img[:, : img.shape[1] // 2].mean()) // 32}{int(img[:, img.shape[1] // 2:].mean()) // 32}"
        rep = ev.evaluate(m, [(f"{i}.png", im) for i, im in enumerate(imgs)], ocr)
        assert rep.num_images == len(imgs) and 0 <= rep.cer_with_sr and len(rep.rows) == len(imgs)
    assert ev.super_resolve_batch(m, []) == []
