"""CPU: host-side logic of the drop-in module and the C-ABI library (load + exports, no compute calls)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

import lpsr_b200
from conftest import ROOT


def _layout(golden_dir):
    return json.load(open(os.path.join(golden_dir, "state_dict_layout.json")))


def test_library_exports_every_symbol_declared_in_header():
    hdr = open(os.path.join(ROOT, "include", "lpsr_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(lpsr_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 19
    lib = ctypes.CDLL(lpsr_b200.capi.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in include/lpsr_b200.h but not exported"
    assert declared == set(lpsr_b200.capi.EXPORTS)
    assert lib.lpsr_abi_version() == lpsr_b200.capi.ABI_VERSION


def test_state_dict_has_reference_158_key_layout(golden_dir):
    lay = _layout(golden_dir)
    m = lpsr_b200.LPSR(3, 32, 16, 4, 4, None)
    sd = m.state_dict()
    assert len(sd) == lay["n_keys"] == 158
    assert list(sd.keys()) == [k["name"] for k in lay["keys"]]           # same names, same order
    for k in lay["keys"]:
        assert list(sd[k["name"]].shape) == k["shape"], k["name"]


def test_alias_and_dead_keys(golden_dir, shipped_weights):
    m = lpsr_b200.LPSR(num_channels=3, num_features=32, growth_rate=16, num_blocks=4, num_layers=4, scale_factor=None)
    sd = m.state_dict()
    # aliases share storage
    assert sd["auto_encoder.GA.0.0.dConv.0.weight"].data_ptr() == sd["auto_encoder.encoder.0.dConv.0.weight"].data_ptr()
    assert sd["auto_encoder.GA.1.3.dConv.1.bias"].data_ptr() == sd["auto_encoder.decoder.3.dConv.1.bias"].data_ptr()
    for i in (1, 3, 5, 7):
        assert sd[f"rdn.rdbs.{i}.conv_out.weight"].data_ptr() == sd["rdn.csar.conv_out.weight"].data_ptr()
    # dead RDBs exist and are not live
    assert "rdn.rdbs.4.lff.weight" in sd and "rdn.rdbs.6.alpha" in sd
    live = set(shipped_weights)
    assert len(live) == 64 and not any(k.startswith(("rdn.rdbs.4", "rdn.rdbs.6")) for k in live)
    uniq = {m.canonical_key(k) for k in sd}
    assert live <= uniq
    # 131,307 live parameters
    assert sum(int(np.prod(v.shape)) if v.shape else 1 for v in shipped_weights.values()) == 131307


def test_strict_load_and_round_trip(shipped_weights):
    m = lpsr_b200.LPSR(3, 32, 16, 4, 4, None)
    m.load_live_weights(shipped_weights)
    sd = m.state_dict()
    for k, v in shipped_weights.items():
        assert np.array_equal(sd[k].numpy(), np.asarray(v).reshape(sd[k].shape)), k
    m2 = lpsr_b200.LPSR(3, 32, 16, 4, 4, 1)
    res = m2.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    bad = dict(sd)
    bad.pop("rdn.rdbs.5.conv_out.bias")
    with pytest.raises(RuntimeError):
        m2.load_state_dict(bad, strict=True)            # strict, like every reference call site (run.py:127)
    # checkpoints wrapped as {'model_state_dict': ...} are unwrapped by the caller (run.py:126), not by the module
    wrapped = {"model_state_dict": sd}
    m2.load_state_dict(wrapped.get("model_state_dict", wrapped))


def test_random_init_is_deterministic_and_standard():
    torch.manual_seed(0)
    a = lpsr_b200.LPSR(3, 32, 16, 4, 4, None).state_dict()
    torch.manual_seed(0)
    b = lpsr_b200.LPSR(3, 32, 16, 4, 4, None).state_dict()
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert float(a["rdn.rdbs.0.alpha"]) == 1.0 and float(a["rdn.rdbs.6.alpha"]) == 1.0   # lpsr.py:58


def test_unsupported_ctor_dims_raise():
    with pytest.raises(ValueError):
        lpsr_b200.LPSR(3, 64, 16, 4, 4, None)
    with pytest.raises(ValueError):
        lpsr_b200.LPSR(3, 32, 16, 4, 4, None, out_channels=3)
    with pytest.raises(ValueError):
        lpsr_b200.LPSR(3, 32, 16, 4, 4, None, precision="int8")


def test_cpu_input_fails_loudly_no_fallback():
    m = lpsr_b200.LPSR(3, 32, 16, 4, 4, None).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.rand(1, 3, 32, 192))
    with pytest.raises(RuntimeError):
        m(torch.rand(3, 32, 192))


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_preprocess_batch_fails_loudly_without_gpu_and_validates_inputs():
    """the batched pre-processing has no CPU path either (the oracle is test-only), and rejects malformed crops."""
    m = lpsr_b200.LPSR(3, 32, 16, 4, 4, None)
    plate = np.zeros((20, 60, 3), dtype=np.uint8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lpsr_b200.preprocess_for_sr_batch(m, [plate])
    for bad in (np.zeros((20, 60), np.uint8), np.zeros((20, 60, 3), np.float32), np.zeros((0, 60, 3), np.uint8)):
        with pytest.raises(ValueError):
            lpsr_b200.preprocess_for_sr_batch(m, [bad], device="cuda:0")   # validated before any device work


def test_create_without_gpu_reports_error_not_fallback():
    lib = lpsr_b200.capi.load_library()
    cfg = lpsr_b200.capi.LpsrConfig(1, 0, 3, 32, 16, 4, 4, 1, 0)
    out = ctypes.c_void_p()
    rc = lib.lpsr_create(ctypes.byref(out), ctypes.byref(cfg))
    assert rc == -2 and out.value is None
    assert b"no CPU fallback" in lib.lpsr_last_error(None)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "license-plate-detection-and-recognition-with-image-enhancement_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle-free", ""), f"{f} mentions the oracle"


def test_shard_bounds_cover_batch_exactly():
    for B in (0, 1, 7, 8, 1024, 1025):
        for G in (1, 2, 3, 4, 8):
            spans = [lpsr_b200.shard_bounds(B, G, r) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(G - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_resample_weight_tables_match_oracle(tmp_path):
    """Host logic of the pre-processing (csrc/preprocess.cuh: resample_coeffs = Pillow's precompute_coeffs + normalize_coeffs_8bpc) against
    the oracle's restatement, without a GPU: a tiny nvcc-built host program prints the tables for several (in, out) sizes."""
    import shutil
    import subprocess
    from oracle import preprocess_oracle as pre
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    cases = [(47, 192), (300, 192), (192, 192), (11, 32), (100, 32), (1, 32), (5, 192), (960, 192)]
    src = tmp_path / "coeffs.cu"
    csrc = os.path.join(ROOT, "license-plate-detection-and-recognition-with-image-enhancement_b200", "csrc")
    body = "".join(f"  dump({a}, {b});\n" for a, b in cases)
    src.write_text('#include <cstdio>\n#include "preprocess.cuh"\n'
                   'static void dump(int in, int out) { std::vector<int> t; int b, k, ks; lpsr::resample_coeffs(in, out, t, b, k, ks);\n'
                   '  printf("%d %d %d", in, out, ks); for (int v : t) printf(" %d", v); printf("\\n"); }\n'
                   f'int main() {{\n{body}  return 0; }}\n')
    exe = tmp_path / "coeffs"
    subprocess.run([nvcc, "-I", csrc, "-o", str(exe), str(src)], check=True, capture_output=True)
    lines = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    assert len(lines) == len(cases)
    for line, (a, b) in zip(lines, cases):
        vals = [int(v) for v in line.split()]
        assert vals[:2] == [a, b]
        bounds, kk = pre.precompute_coeffs(a, b)
        assert vals[2] == kk.shape[1]
        assert vals[3:3 + 2 * b] == bounds.reshape(-1).tolist()
        assert vals[3 + 2 * b:] == kk.reshape(-1).tolist()


def test_pipeline_array_steps_match_reference_vectors(golden_dir):
    """format_long_plate / restack_to_square (inference/run.py:21-78) and the gray -> BGR step (run.py:204) against
    vectors produced by the reference's own functions and OpenCV (tests/golden/make_golden_pipeline.py), bit for bit."""
    from lpsr_b200 import pipeline as pl
    d = np.load(os.path.join(golden_dir, "pipeline_cases.npz"))
    n = int(d["n"])
    assert n >= 10
    n_changed = 0
    for i in range(n):
        f, changed = pl.format_long_plate(d[f"in_{i}"])
        assert changed == bool(d[f"changed_{i}"])
        n_changed += changed
        assert np.array_equal(f, d[f"fmt_{i}"])
        assert np.array_equal(pl.restack_to_square(f), d[f"restack_{i}"])
        assert np.array_equal(pl.restack_to_square(d[f"in_{i}"]), d[f"restack_raw_{i}"])
        assert np.array_equal(pl.gray_to_bgr(d[f"gray_{i}"]), d[f"bgr_{i}"])
    assert 0 < n_changed < n
    e = np.zeros((0, 5, 3), np.uint8)
    assert pl.format_long_plate(e)[1] is False and pl.restack_to_square(e) is e
    assert pl.select_plates([(0, 0, 2, 2), (0, 0, 10, 10), (0, 0, 5, 5), (0, 0, 1, 1)]) == [(0, 0, 10, 10), (0, 0, 5, 5), (0, 0, 2, 2)]


def test_synthetic_clip_is_deterministic_and_boxed():
    from lpsr_b200 import pipeline as pl
    f1, b1 = pl.synthetic_clip(2, 3, seed=4)
    f2, b2 = pl.synthetic_clip(2, 3, seed=4)
    assert b1 == b2 and all(np.array_equal(a, b) for a, b in zip(f1, f2))
    assert f1[0].shape == (1080, 1920, 3) and f1[0].dtype == np.uint8 and len(b1[0]) == 3
    for (x1, y1, x2, y2) in b1[0]:
        assert 0 <= x1 < x2 <= 1920 and 0 <= y1 < y2 <= 1080
