"""CPU: the oracle (numpy restatement) and the torch.nn.functional port against the golden vectors generated from the
unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import golden_cases, load_case
from oracle import lpsr_oracle as orc
from oracle import lpsr_torch_port as port

SHIPPED = [c for c in golden_cases() if not c.startswith("rand_")]
RANDOM = [c for c in golden_cases() if c.startswith("rand_")]


def _weights_for(case, shipped):
    if not case.startswith("rand_"):
        return shipped
    d = load_case(case)
    W = orc.random_weights(int(d["seed"]))
    if d["alphas"].size:
        W["rdn.rdbs.0.alpha"] = np.float32(d["alphas"][0]).reshape(())
        W["rdn.rdbs.2.alpha"] = np.float32(d["alphas"][1]).reshape(())
    return W


@pytest.mark.parametrize("case", SHIPPED + RANDOM)
def test_numpy_oracle_matches_reference_golden(case, shipped_weights):
    if "128x384" in case:
        pytest.skip("covered by the torch port (numpy im2col of this size is slow)")
    d = load_case(case)
    y = orc.lpsr_forward(d["x"], _weights_for(case, shipped_weights))
    assert y.shape == d["y"].shape
    assert np.abs(y - d["y"]).max() <= 2e-5     # fp32 summation-order noise only


@pytest.mark.parametrize("case", SHIPPED + RANDOM)
def test_torch_port_matches_reference_golden(case, shipped_weights):
    d = load_case(case)
    W = port.to_torch_weights(_weights_for(case, shipped_weights))
    y = port.lpsr_forward(torch.from_numpy(d["x"]), W).numpy()
    assert y.shape == d["y"].shape
    assert np.abs(y - d["y"]).max() <= 2e-5     # bit-identical on the generating host; ISA-dependent elsewhere


def test_numpy_oracle_float64_agrees(shipped_weights):
    d = load_case("u_b2_32x192")
    y64 = orc.lpsr_forward(d["x"], shipped_weights, dtype=np.float64)
    assert np.abs(y64 - d["y"]).max() <= 2e-5


def test_oracle_intermediates_match_reference_hooks(shipped_weights, golden_dir):
    t = np.load(golden_dir + "/taps_u_b1_16x32.npz")
    taps = {}
    y = orc.lpsr_forward(t["x"], shipped_weights, taps=taps)
    assert np.abs(y - t["y"]).max() <= 2e-5
    for name in ("ae.c0", "ae.enc0", "ae.enc1", "ae.dec0", "ae.dec1", "ae.out", "rdn.sfe1", "rdn.sfe2", "rdn.block0",
                 "rdn.block1", "rdn.block2", "rdn.block3", "csar1.x_in", "csar3.x_in", "rdn.out"):
        ref = t[name]
        scale = max(1.0, float(np.abs(ref).max()))
        assert taps[name].shape == ref.shape, name
        assert np.abs(taps[name] - ref).max() <= 2e-5 * scale, name


def test_pixel_shuffle_index_rules():
    """unshuffle: out[n, c*4+i*2+j, h, w] = in[n, c, 2h+i, 2w+j]; shuffle is its inverse; bit exact."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 3, 8, 12)).astype(np.float32)
    u = orc.pixel_unshuffle(x)
    for c in range(3):
        for i in range(2):
            for j in range(2):
                assert np.array_equal(u[:, c * 4 + i * 2 + j], x[:, c, i::2, j::2])
    assert np.array_equal(orc.pixel_shuffle(u), x)
    assert np.array_equal(u, torch.nn.functional.pixel_unshuffle(torch.from_numpy(x), 2).numpy())
    assert np.array_equal(orc.pixel_shuffle(u), torch.nn.functional.pixel_shuffle(torch.from_numpy(u), 2).numpy())


def test_pad_to_multiple_of_4_never_cropped(shipped_weights):
    d = load_case("u_b1_33x193_pad")
    assert d["y"].shape == (1, 1, 36, 196)
    d = load_case("u_b1_30x190_pad")
    assert d["y"].shape == (1, 1, 32, 192)


def test_survey_sanity_vector(shipped_weights):
    """SURVEY.md 8c golden sanity vector (seed 1234, 2x3x64x192), first crop only to keep the numpy run short."""
    x = torch.rand(2, 3, 64, 192, generator=torch.Generator().manual_seed(1234))
    y = port.lpsr_forward(x, port.to_torch_weights(shipped_weights)).numpy()
    assert abs(float(y.sum()) - 15583.47) < 0.05
    assert np.allclose(y[0, 0, 0, :4], [0.54689, 0.60430, 0.63589, 0.61160], atol=2e-5)
    assert np.allclose(y[1, 0, 63, 188:], [0.64432, 0.54506, 0.48248, 0.49594], atol=2e-5)


def _s2d_weights(w, b):
    """The weight transformation the tensor-core AutoEncoder uses (engine.cu: s2d_weights), restated in numpy: a ks x ks conv on the
    fine grid equals a 3x3 (5x5 for ks = 7) conv on the 2x coarser grid whose pixels carry their 2x2 fine pixels as channels in
    PixelUnshuffle order c*4 + i*2 + j:  w'[(co,I,J)][(ci,i,j)][th][tw] = w[co][ci][2th + i - I + R][2tw + j - J + R]."""
    cout, cin, ks, _ = w.shape
    R, RC = ks // 2, (1 if ks <= 5 else 2)
    KC = 2 * RC + 1
    ws = np.zeros((cout * 4, cin * 4, KC, KC), dtype=w.dtype)
    for I in range(2):
        for J in range(2):
            for i in range(2):
                for j in range(2):
                    for th in range(-RC, RC + 1):
                        dy = 2 * th + i - I + R
                        if not 0 <= dy < ks:
                            continue
                        for tw in range(-RC, RC + 1):
                            dx = 2 * tw + j - J + R
                            if 0 <= dx < ks:
                                ws[I * 2 + J::4, i * 2 + j::4, th + RC, tw + RC] = w[:, :, dy, dx]
    return ws, (None if b is None else np.repeat(b, 4))


@pytest.mark.parametrize("ks", [3, 5, 7])
def test_space_to_depth_conv_identity(ks):
    """unshuffle(conv(x, w)) == conv(unshuffle(x), s2d(w)) with 'same' zero padding: the identity that lets the 16-bit path run the
    AutoEncoder (lpsr.py:64-117) and shallowF1 (lpsr.py:195-197) on the half / quarter grid with the pixel (un)shuffles folded away."""
    rng = np.random.default_rng(ks)
    x = rng.standard_normal((2, 3, 12, 20)).astype(np.float32)
    w = rng.standard_normal((5, 3, ks, ks)).astype(np.float32)
    b = rng.standard_normal(5).astype(np.float32)
    ws, bs = _s2d_weights(w, b)
    fine = orc.pixel_unshuffle(orc.conv2d(x, w, b))
    coarse = orc.conv2d(orc.pixel_unshuffle(x), ws, bs)
    assert fine.shape == coarse.shape
    assert np.abs(fine - coarse).max() <= 2e-5


def test_dconv_is_one_dense_conv(shipped_weights):
    """DConv (depthwise k x k + bias, pointwise 1x1 + bias, no activation in between, lpsr.py:8-28) == one dense k x k conv with
    w[co][ci] = pw[co][ci] * dw[ci] and bias pw_b + pw @ dw_b: what the tensor-core path packs (engine.cu: compose_dconv)."""
    W = shipped_weights
    p = "auto_encoder.encoder.3.dConv."
    dw, dwb, pw, pwb = W[p + "0.weight"], W[p + "0.bias"], W[p + "1.weight"], W[p + "1.bias"]
    x = np.random.default_rng(0).standard_normal((1, dw.shape[0], 9, 11)).astype(np.float32)
    ref = orc.conv2d(orc.conv2d(x, dw, dwb, groups=dw.shape[0]), pw, pwb)
    dense_w = pw[:, :, 0, 0][:, :, None, None] * dw[:, 0][None]
    dense_b = pwb + pw[:, :, 0, 0] @ dwb
    assert np.abs(orc.conv2d(x, dense_w, dense_b) - ref).max() <= 1e-5


def test_preprocess_oracle_matches_reference_golden(golden_dir):
    """the oracle's restatement of Pillow's bicubic resample + BGR->RGB + ToTensor is bit-exact with the
    reference's own library calls (tests/golden/make_golden_preprocess.py) on every fixture."""
    import os
    from oracle import preprocess_oracle as pre
    d = np.load(os.path.join(golden_dir, "preprocess_cases.npz"))
    n = sum(1 for k in d.files if k.startswith("in_"))
    assert n >= 10
    for i in range(n):
        got = pre.preprocess_for_sr(d[f"in_{i}"])
        assert got.dtype == np.float32 and got.shape == (1, 3, 32, 192)
        assert np.array_equal(got, d[f"out_{i}"]), i
