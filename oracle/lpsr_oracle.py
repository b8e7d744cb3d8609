"""CPU oracle for the LPSR forward pass -- TEST INFRASTRUCTURE, NOT PRODUCT.

This file is a plain-numpy restatement of the algorithm executed by the reference's
``my_models/lpsr.py`` (class ``LPSR``, lpsr.py:252-274).  It exists only so that tests,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` leg can check / time the
CUDA path against an independent CPU statement of the same arithmetic.  Nothing in the
product package may import it; the product path fails loudly when the CUDA extension is
missing.

Parity pin: the reference has no tests of its own, so the pin
is "outputs of the reference itself run here": ``tests/golden/make_golden.py`` imports the
unmodified reference module from /root/reference with the shipped ``weights/best_model.pth``
and stores seeded input/output pairs plus per-stage intermediates under ``tests/golden/``;
``tests/test_oracle.py`` checks this restatement against every one of them.

All arrays are NCHW float32 (or float64 when ``dtype=np.float64`` is requested) exactly as
in the reference.  Weight names are the reference's state_dict keys (first alias of each
live tensor).
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# canonical live weight names (64 tensors) -- reference state_dict keys, lpsr.py ctor order
# --------------------------------------------------------------------------------------


def live_weight_names(num_layers: int = 4, executed_rdbs=(0, 2)) -> list[str]:
    names = ["auto_encoder.conv_in.weight"]
    for blk in ("encoder.0", "encoder.3", "decoder.0", "decoder.3"):
        for j in (0, 1):
            names += [f"auto_encoder.{blk}.dConv.{j}.weight", f"auto_encoder.{blk}.dConv.{j}.bias"]
    names += ["auto_encoder.conv_out.weight"]
    names += ["rdn.shallowF1.weight", "rdn.shallowF1.bias", "rdn.shallowF2.weight", "rdn.shallowF2.bias"]
    for p in ("conv_in.0", "conv_in.2", "ca.block.2", "ca.block.4", "sa.block.0", "sa.block.2", "conv_out"):
        names += [f"rdn.csar.{p}.weight", f"rdn.csar.{p}.bias"]
    for r in executed_rdbs:
        names.append(f"rdn.rdbs.{r}.alpha")
        for i in range(num_layers):
            names += [f"rdn.rdbs.{r}.layers.{i}.conv.weight", f"rdn.rdbs.{r}.layers.{i}.conv.bias"]
        names += [f"rdn.rdbs.{r}.lff.weight", f"rdn.rdbs.{r}.lff.bias"]
    names += ["rdn.gff.0.weight", "rdn.gff.0.bias", "rdn.gff.1.weight", "rdn.gff.1.bias"]
    names += ["final_conv.weight", "final_conv.bias"]
    return names


# --------------------------------------------------------------------------------------
# primitives (semantics of nn.Conv2d / PixelShuffle / PixelUnshuffle as used by lpsr.py)
# --------------------------------------------------------------------------------------


def conv2d(x, w, b=None, groups: int = 1):
    """Cross-correlation, stride 1, zero 'same' padding (k//2), NCHW.

    Matches every nn.Conv2d on the path (lpsr.py:15-24, 34-36, 67-69, 102-104, 159-172, 195-200,
    207-212, 262-267): all are stride 1, dilation 1, odd kernel, padding='same' or k//2.
    """
    n, cin, h, wd = x.shape
    cout, cin_g, kh, kw = w.shape
    assert cin_g * groups == cin and cout % groups == 0
    ph, pw = kh // 2, kw // 2
    xp = np.zeros((n, cin, h + 2 * ph, wd + 2 * pw), dtype=x.dtype)
    xp[:, :, ph:ph + h, pw:pw + wd] = x
    if groups == 1:
        # im2col: [N, Cin*kh*kw, H*W] then one GEMM per sample
        cols = np.empty((n, cin, kh, kw, h, wd), dtype=x.dtype)
        for i in range(kh):
            for j in range(kw):
                cols[:, :, i, j] = xp[:, :, i:i + h, j:j + wd]
        cols = cols.reshape(n, cin * kh * kw, h * wd)
        wm = w.reshape(cout, cin * kh * kw).astype(x.dtype)
        out = np.matmul(wm[None], cols).reshape(n, cout, h, wd)
    else:
        assert groups == cin and cout == cin and cin_g == 1, "only depthwise is on the path"
        out = np.zeros((n, cout, h, wd), dtype=x.dtype)
        for i in range(kh):
            for j in range(kw):
                out += xp[:, :, i:i + h, j:j + wd] * w[None, :, 0, i, j, None, None].astype(x.dtype)
    if b is not None:
        out = out + b.astype(x.dtype)[None, :, None, None]
    return out


def pixel_unshuffle(x, r: int = 2):
    """nn.PixelUnshuffle(r): out[n, c*r*r + i*r + j, h, w] = in[n, c, h*r + i, w*r + j] (lpsr.py:72,79)."""
    n, c, h, w = x.shape
    assert h % r == 0 and w % r == 0
    x = x.reshape(n, c, h // r, r, w // r, r)
    return np.ascontiguousarray(x.transpose(0, 1, 3, 5, 2, 4)).reshape(n, c * r * r, h // r, w // r)


def pixel_shuffle(x, r: int = 2):
    """nn.PixelShuffle(r): out[n, c, h*r + i, w*r + j] = in[n, c*r*r + i*r + j, h, w] (lpsr.py:88,95)."""
    n, c, h, w = x.shape
    assert c % (r * r) == 0
    co = c // (r * r)
    x = x.reshape(n, co, r, r, h, w)
    return np.ascontiguousarray(x.transpose(0, 1, 4, 2, 5, 3)).reshape(n, co, h * r, w * r)


def relu(x):
    return np.maximum(x, 0)


def sigmoid(x):
    # numerically stable logistic, same values as torch.sigmoid to ~1 ulp
    out = np.empty_like(x)
    pos = x >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-x[pos]))
    e = np.exp(x[~pos])
    out[~pos] = e / (1.0 + e)
    return out


# --------------------------------------------------------------------------------------
# modules
# --------------------------------------------------------------------------------------


def _dconv(x, W, prefix):
    """DConv (lpsr.py:8-28): depthwise kxk (+bias) then pointwise 1x1 (+bias)."""
    x = conv2d(x, W[prefix + ".dConv.0.weight"], W[prefix + ".dConv.0.bias"], groups=x.shape[1])
    return conv2d(x, W[prefix + ".dConv.1.weight"], W[prefix + ".dConv.1.bias"])


def pad_to_multiple_of_4(x):
    """AutoEncoder.forward lpsr.py:107-111: zero-pad bottom of H, then right of W, to multiples of 4."""
    n, c, h, w = x.shape
    hp = h + ((4 - h % 4) % 4)
    wp = w + ((4 - w % 4) % 4)
    if hp == h and wp == w:
        return x
    out = np.zeros((n, c, hp, wp), dtype=x.dtype)
    out[:, :, :h, :w] = x
    return out


def auto_encoder(x, W, taps=None):
    """AutoEncoder.forward (lpsr.py:106-117)."""
    p = "auto_encoder."
    x = pad_to_multiple_of_4(x)
    c0 = conv2d(x, W[p + "conv_in.weight"])                       # 3->12, no bias (lpsr.py:67-69)
    e = relu(pixel_unshuffle(_dconv(c0, W, p + "encoder.0")))      # 12->12 -> [48,H/2,W/2] (71-73)
    if taps is not None:
        taps["ae.enc0"] = e
    e = relu(pixel_unshuffle(_dconv(e, W, p + "encoder.3")))       # 48->12 -> [48,H/4,W/4] (74-80)
    if taps is not None:
        taps["ae.enc1"] = e
    d = relu(pixel_shuffle(_dconv(e, W, p + "decoder.0")))         # 48->48 -> [12,H/2,W/2] (83-89)
    if taps is not None:
        taps["ae.dec0"] = d
    d = relu(pixel_shuffle(_dconv(d, W, p + "decoder.3")))         # 12->48 -> [12,H,W]     (90-96)
    if taps is not None:
        taps["ae.dec1"] = d
    out = conv2d(c0 + d, W[p + "conv_out.weight"])                 # 12->3, no bias (102-104,115-116)
    if taps is not None:
        taps["ae.c0"] = c0
        taps["ae.out"] = out
    return out


def rdb(x, W, prefix, num_layers=4):
    """RDB.forward (lpsr.py:43-61): x + alpha * lff(dense-chain(x))."""
    cat = x
    for i in range(num_layers):
        y = relu(conv2d(cat, W[f"{prefix}.layers.{i}.conv.weight"], W[f"{prefix}.layers.{i}.conv.bias"]))
        cat = np.concatenate([cat, y], axis=1)                     # DenseLayer lpsr.py:39-40
    alpha = np.asarray(W[prefix + ".alpha"], dtype=x.dtype).reshape(())
    return x + alpha * conv2d(cat, W[prefix + ".lff.weight"], W[prefix + ".lff.bias"])


def csar(x, W, prefix="rdn.csar", taps=None, tag=""):
    """CSAR.forward (lpsr.py:180-186) incl. the x_in**2 channel branch."""
    x_in = conv2d(relu(conv2d(x, W[prefix + ".conv_in.0.weight"], W[prefix + ".conv_in.0.bias"])),
                  W[prefix + ".conv_in.2.weight"], W[prefix + ".conv_in.2.bias"])
    # ChannelAttention (lpsr.py:120-135): returns x_in * s_c
    mean = x_in.mean(axis=(2, 3), dtype=np.float64).astype(x.dtype)              # AdaptiveAvgPool2d(1)
    hid = relu(mean @ W[prefix + ".ca.block.2.weight"].T.astype(x.dtype) + W[prefix + ".ca.block.2.bias"].astype(x.dtype))
    s_c = sigmoid(hid @ W[prefix + ".ca.block.4.weight"].T.astype(x.dtype) + W[prefix + ".ca.block.4.bias"].astype(x.dtype))
    x_ca = x_in * s_c[:, :, None, None]
    # SpatialAttention (lpsr.py:138-153): per-pixel MLP 32->64->32 + sigmoid
    s_s = sigmoid(conv2d(relu(conv2d(x_in, W[prefix + ".sa.block.0.weight"], W[prefix + ".sa.block.0.bias"])),
                         W[prefix + ".sa.block.2.weight"], W[prefix + ".sa.block.2.bias"]))
    cat = np.concatenate([x_in * x_ca, x_in * s_s], axis=1)                      # lpsr.py:184
    if taps is not None:
        taps[f"csar{tag}.x_in"] = x_in
        taps[f"csar{tag}.s_c"] = s_c
    return x + conv2d(cat, W[prefix + ".conv_out.weight"], W[prefix + ".conv_out.bias"])


def rdn(x, W, num_blocks=4, num_layers=4, taps=None):
    """RDN.forward (lpsr.py:214-225).  Executes rdbs[0..num_blocks-1] only:
    even module indices are RDBs, odd ones the single shared CSAR (Q2)."""
    sfe1 = conv2d(x, W["rdn.shallowF1.weight"], W["rdn.shallowF1.bias"])
    sfe2 = conv2d(sfe1, W["rdn.shallowF2.weight"], W["rdn.shallowF2.bias"])
    if taps is not None:
        taps["rdn.sfe1"] = sfe1
        taps["rdn.sfe2"] = sfe2
    h = sfe2
    feats = []
    for i in range(num_blocks):
        if i % 2 == 0:
            h = rdb(h, W, f"rdn.rdbs.{i}", num_layers)
        else:
            h = csar(h, W, "rdn.csar", taps, tag=str(i))
        if taps is not None:
            taps[f"rdn.block{i}"] = h
        feats.append(h)
    g = conv2d(np.concatenate(feats, axis=1), W["rdn.gff.0.weight"], W["rdn.gff.0.bias"])
    g = conv2d(g, W["rdn.gff.1.weight"], W["rdn.gff.1.bias"])
    out = g + sfe1
    if taps is not None:
        taps["rdn.out"] = out
    return out


def lpsr_forward(x, W, num_blocks=4, num_layers=4, taps=None, dtype=np.float32):
    """LPSR.forward (lpsr.py:269-274): sigmoid(final_conv(rdn(auto_encoder(x)))).

    x: [B,3,H,W] in [0,1]; returns [B,1,ceil4(H),ceil4(W)].
    """
    x = np.asarray(x, dtype=dtype)
    W = {k: np.asarray(v, dtype=dtype) for k, v in W.items()}
    a = auto_encoder(x, W, taps)
    r = rdn(a, W, num_blocks, num_layers, taps)
    y = conv2d(r, W["final_conv.weight"], W["final_conv.bias"])
    return sigmoid(y)


def random_weights(seed: int = 0, scale: float = 1.0, num_layers: int = 4) -> dict:
    """Seeded random live weights with nn.Conv2d-like fan-in scaling (for tests without the checkpoint)."""
    rng = np.random.default_rng(seed)
    shapes = weight_shapes(num_layers)
    W = {}
    for k, shp in shapes.items():
        if k.endswith("alpha"):
            W[k] = np.asarray(1.0 + 0.1 * rng.standard_normal(), dtype=np.float32).reshape(())
            continue
        fan_in = int(np.prod(shp[1:])) if len(shp) > 1 else None
        if fan_in is None:  # bias: find its weight's fan-in
            wk = k[:-4] + "weight"
            fan_in = int(np.prod(shapes[wk][1:]))
        bound = scale / np.sqrt(fan_in)
        W[k] = rng.uniform(-bound, bound, size=shp).astype(np.float32)
    return W


def weight_shapes(num_layers: int = 4, F: int = 32, G: int = 16, C: int = 3, out_ch: int = 1) -> dict:
    E = 4 * C
    s = {"auto_encoder.conv_in.weight": (E, C, 3, 3), "auto_encoder.conv_out.weight": (C, E, 3, 3)}
    for blk, cin, cout in (("encoder.0", E, E), ("encoder.3", 4 * E, E), ("decoder.0", 4 * E, 4 * E), ("decoder.3", E, 4 * E)):
        s[f"auto_encoder.{blk}.dConv.0.weight"] = (cin, 1, 5, 5)
        s[f"auto_encoder.{blk}.dConv.0.bias"] = (cin,)
        s[f"auto_encoder.{blk}.dConv.1.weight"] = (cout, cin, 1, 1)
        s[f"auto_encoder.{blk}.dConv.1.bias"] = (cout,)
    s["rdn.shallowF1.weight"] = (F, C, 7, 7); s["rdn.shallowF1.bias"] = (F,)
    s["rdn.shallowF2.weight"] = (F, F, 3, 3); s["rdn.shallowF2.bias"] = (F,)
    s["rdn.csar.conv_in.0.weight"] = (F, F, 3, 3); s["rdn.csar.conv_in.0.bias"] = (F,)
    s["rdn.csar.conv_in.2.weight"] = (F, F, 3, 3); s["rdn.csar.conv_in.2.bias"] = (F,)
    s["rdn.csar.ca.block.2.weight"] = (F // 4, F); s["rdn.csar.ca.block.2.bias"] = (F // 4,)
    s["rdn.csar.ca.block.4.weight"] = (F, F // 4); s["rdn.csar.ca.block.4.bias"] = (F,)
    s["rdn.csar.sa.block.0.weight"] = (2 * F, F, 1, 1); s["rdn.csar.sa.block.0.bias"] = (2 * F,)
    s["rdn.csar.sa.block.2.weight"] = (F, 2 * F, 1, 1); s["rdn.csar.sa.block.2.bias"] = (F,)
    s["rdn.csar.conv_out.weight"] = (F, 2 * F, 1, 1); s["rdn.csar.conv_out.bias"] = (F,)
    for r in (0, 2):
        s[f"rdn.rdbs.{r}.alpha"] = ()
        for i in range(num_layers):
            s[f"rdn.rdbs.{r}.layers.{i}.conv.weight"] = (G, F + G * i, 3, 3)
            s[f"rdn.rdbs.{r}.layers.{i}.conv.bias"] = (G,)
        s[f"rdn.rdbs.{r}.lff.weight"] = (F, F + G * num_layers, 1, 1)
        s[f"rdn.rdbs.{r}.lff.bias"] = (F,)
    s["rdn.gff.0.weight"] = (F, 4 * F, 1, 1); s["rdn.gff.0.bias"] = (F,)
    s["rdn.gff.1.weight"] = (F, F, 3, 3); s["rdn.gff.1.bias"] = (F,)
    s["final_conv.weight"] = (out_ch, F, 3, 3); s["final_conv.bias"] = (out_ch,)
    return s
