"""torch.nn.functional port of the LPSR forward -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT.

Same executed graph as ``oracle/lpsr_oracle.py`` (and therefore as the reference's
``my_models/lpsr.py:269-274``), but expressed with ``torch.nn.functional`` calls so that it runs
through the very ATen/oneDNN CPU kernels the reference's ``nn.Module`` dispatches to.  It is used

* by ``bench.py`` as the ``cpu_baseline`` / ``--impl reference`` arm on the GPU box's host cores
  (``kind: "port"`` -- the reference itself is Python under /root/reference and cannot travel to
  the GPU box), and
* by tests as a fast checker at batch sizes the numpy oracle would take too long on.

It is validated against the real reference module in ``tests/golden/make_golden.py`` (bit-identical
on CPU, because op order and kernels are the same) and against the golden fixtures in
``tests/test_oracle.py``.  Only tests/, smoke() and bench.py may import this file.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _t(W, k):
    return W[k]


def _conv(x, W, name, bias=True, groups=1):
    w = W[name + ".weight"]
    b = W[name + ".bias"] if bias else None
    return F.conv2d(x, w, b, stride=1, padding=(w.shape[2] // 2, w.shape[3] // 2), groups=groups)


def _dconv(x, W, p):
    x = _conv(x, W, p + ".dConv.0", groups=x.shape[1])
    return _conv(x, W, p + ".dConv.1")


def auto_encoder(x, W):
    h, w = x.shape[-2:]
    if h % 4:
        x = F.pad(x, (0, 0, 0, 4 - h % 4))
    if w % 4:
        x = F.pad(x, (0, 4 - w % 4, 0, 0))
    p = "auto_encoder."
    c0 = _conv(x, W, p + "conv_in", bias=False)
    e = F.relu(F.pixel_unshuffle(_dconv(c0, W, p + "encoder.0"), 2))
    e = F.relu(F.pixel_unshuffle(_dconv(e, W, p + "encoder.3"), 2))
    d = F.relu(F.pixel_shuffle(_dconv(e, W, p + "decoder.0"), 2))
    d = F.relu(F.pixel_shuffle(_dconv(d, W, p + "decoder.3"), 2))
    return _conv(c0 + d, W, p + "conv_out", bias=False)


def rdb(x, W, p, num_layers=4):
    cat = x
    for i in range(num_layers):
        cat = torch.cat([cat, F.relu(_conv(cat, W, f"{p}.layers.{i}.conv"))], 1)
    return x + W[p + ".alpha"] * _conv(cat, W, p + ".lff")


def csar(x, W, p="rdn.csar"):
    x_in = _conv(F.relu(_conv(x, W, p + ".conv_in.0")), W, p + ".conv_in.2")
    m = F.adaptive_avg_pool2d(x_in, 1).flatten(1)
    s_c = torch.sigmoid(F.linear(F.relu(F.linear(m, W[p + ".ca.block.2.weight"], W[p + ".ca.block.2.bias"])),
                                 W[p + ".ca.block.4.weight"], W[p + ".ca.block.4.bias"]))
    x_ca = x_in * s_c[:, :, None, None]
    s_s = torch.sigmoid(_conv(F.relu(_conv(x_in, W, p + ".sa.block.0")), W, p + ".sa.block.2"))
    return x + _conv(torch.cat([x_in * x_ca, x_in * s_s], 1), W, p + ".conv_out")


def rdn(x, W, num_blocks=4, num_layers=4):
    sfe1 = _conv(x, W, "rdn.shallowF1")
    h = _conv(sfe1, W, "rdn.shallowF2")
    feats = []
    for i in range(num_blocks):
        h = rdb(h, W, f"rdn.rdbs.{i}", num_layers) if i % 2 == 0 else csar(h, W)
        feats.append(h)
    return _conv(_conv(torch.cat(feats, 1), W, "rdn.gff.0"), W, "rdn.gff.1") + sfe1


@torch.no_grad()
def lpsr_forward(x: torch.Tensor, W: dict, num_blocks: int = 4, num_layers: int = 4) -> torch.Tensor:
    """x: [B,3,H,W] fp32 CPU tensor; W: dict name -> torch tensor (live names)."""
    return torch.sigmoid(_conv(rdn(auto_encoder(x, W), W, num_blocks, num_layers), W, "final_conv"))


def to_torch_weights(W: dict, dtype=torch.float32) -> dict:
    return {k: torch.as_tensor(v).to(dtype) for k, v in W.items()}
