"""Scratch study (CPU): which tensors may be stored/multiplied in bf16 while keeping max|err| <= 1e-2
against the fp32 oracle with the shipped checkpoint.  Emulates bf16 operand rounding with fp32 accumulate."""
import sys, os, numpy as np, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import lpsr_torch_port as port
torch.set_num_threads(8)

def q(x):  # bf16 round-trip
    return x.to(torch.bfloat16).to(torch.float32)

def conv(x, W, name, qa, qw, bias=True, split_w=False, split_a=False):
    w = W[name + ".weight"]; b = W[name + ".bias"] if bias else None
    pad = (w.shape[2] // 2, w.shape[3] // 2)
    xa = q(x) if qa else x
    if qw:
        wh = q(w)
        y = F.conv2d(xa, wh, b, padding=pad)
        if split_w:
            y = y + F.conv2d(xa, q(w - wh), None, padding=pad)
        if split_a and qa:
            y = y + F.conv2d(q(x - xa), wh, None, padding=pad)
        return y
    return F.conv2d(xa, w, b, padding=pad)

def forward(x, W, cfg):
    a = port.auto_encoder(x, W)   # AE fp32 CUDA cores
    c = lambda t, n, g, **kw: conv(t, W, n, cfg[g][0], cfg[g][1], split_w=cfg[g][2] if len(cfg[g]) > 2 else False,
                                   split_a=cfg[g][3] if len(cfg[g]) > 3 else False, **kw)
    sfe1 = c(a, "rdn.shallowF1", "sfe1")
    h = c(sfe1, "rdn.shallowF2", "sfe2")
    feats = []
    for i in range(4):
        if i % 2 == 0:
            p = f"rdn.rdbs.{i}"
            cat = h
            for l in range(4):
                cat = torch.cat([cat, F.relu(c(cat, f"{p}.layers.{l}.conv", "rdb"))], 1)
            h = h + W[p + ".alpha"] * c(cat, p + ".lff", "lff")
        else:
            p = "rdn.csar"
            t = F.relu(c(h, p + ".conv_in.0", "csar_c1"))
            x_in = c(t, p + ".conv_in.2", "csar_c2")
            if cfg.get("x_in_bf16"): x_in = q(x_in)
            m = x_in.mean((2, 3))
            s_c = torch.sigmoid(F.linear(F.relu(F.linear(m, W[p + ".ca.block.2.weight"], W[p + ".ca.block.2.bias"])), W[p + ".ca.block.4.weight"], W[p + ".ca.block.4.bias"]))
            s_s = torch.sigmoid(c(F.relu(c(x_in, p + ".sa.block.0", "tail")), p + ".sa.block.2", "tail"))
            g = torch.cat([x_in * x_in * s_c[:, :, None, None], x_in * s_s], 1)
            h = h + c(g, p + ".conv_out", "tail_out")
        if cfg.get("res_bf16"): h = q(h)
        feats.append(h)
    g = c(c(torch.cat(feats, 1), "rdn.gff.0", "gff0"), "rdn.gff.1", "gff1") + sfe1
    return torch.sigmoid(c(g, "final_conv", "final"))

def smooth(b, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    lo = torch.rand(b, 3, h // 8, w // 8, generator=g)
    return F.interpolate(lo, size=(h, w), mode="bicubic", align_corners=False).clamp(0, 1)

if __name__ == "__main__":
    Wn = dict(np.load(os.path.join(ROOT, "tests/golden/weights_best_model.npz")))
    W = port.to_torch_weights(Wn)
    xs = {"uniform": torch.rand(8, 3, 64, 192, generator=torch.Generator().manual_seed(0)), "smooth": smooth(8, 64, 192, 1)}
    groups = ["sfe1", "sfe2", "rdb", "lff", "csar_c1", "csar_c2", "tail", "tail_out", "gff0", "gff1", "final"]
    FP = (False, False)
    def mk(default, **over):
        c = {g: default for g in groups}; c.update(over); return c
    cfgs = {
        "all fp32": mk(FP),
        "all bf16 x1": mk((True, True)),
        "all bf16 + x_in/res bf16": dict(mk((True, True)), x_in_bf16=True, res_bf16=True),
        "bf16 but tail fp32": mk((True, True), tail=FP, tail_out=FP),
        "bf16 but tail+sfe1 fp32": mk((True, True), tail=FP, tail_out=FP, sfe1=FP),
        "bf16 but tail+sfe1+final fp32": mk((True, True), tail=FP, tail_out=FP, sfe1=FP, final=FP),
        "bf16 w-split everywhere": mk((True, True, True)),
        "bf16 w-split, tail+sfe1 fp32": mk((True, True, True), tail=FP, tail_out=FP, sfe1=FP),
        "bf16 w+a split everywhere": mk((True, True, True, True)),
        "only rdb bf16": mk(FP, rdb=(True, True)),
        "only csar convs bf16": mk(FP, csar_c1=(True, True), csar_c2=(True, True)),
        "only csar c2 bf16": mk(FP, csar_c2=(True, True)),
        "only tail bf16": mk(FP, tail=(True, True)),
        "only tail_out bf16": mk(FP, tail_out=(True, True)),
        "only sfe1 bf16": mk(FP, sfe1=(True, True)),
        "only sfe2 bf16": mk(FP, sfe2=(True, True)),
        "only gff bf16": mk(FP, gff0=(True, True), gff1=(True, True)),
        "only final bf16": mk(FP, final=(True, True)),
        "only lff bf16": mk(FP, lff=(True, True)),
    }
    for xn, x in xs.items():
        with torch.no_grad():
            ref = port.lpsr_forward(x, W)
            for name, cfg in cfgs.items():
                y = forward(x, W, cfg)
                d = (y - ref).abs()
                mse = ((y - ref) ** 2).mean().item()
                print(f"{xn:8s} {name:36s} max {d.max().item():.2e}  mean {d.mean().item():.2e}  psnr-vs-ref {10*np.log10(1/max(mse,1e-30)):.1f}")

def more():
    Wn = dict(np.load(os.path.join(ROOT, "tests/golden/weights_best_model.npz")))
    W = port.to_torch_weights(Wn)
    xs = {"uniform": torch.rand(8, 3, 64, 192, generator=torch.Generator().manual_seed(0)), "smooth": smooth(8, 64, 192, 1),
          "smooth2": smooth(8, 64, 192, 2)}
    groups = ["sfe1", "sfe2", "rdb", "lff", "csar_c1", "csar_c2", "tail", "tail_out", "gff0", "gff1", "final"]
    FP = (False, False); B1 = (True, True); B3 = (True, True, True, True); BA = (True, True, False, True); BW = (True, True, True, False)
    def mk(default, **over):
        c = {g: default for g in groups}; c.update(over); return c
    cfgs = {
        "bf16x1": mk(B1),
        "1x1 convs x3, rest bf16x1": mk(B1, lff=B3, gff0=B3, tail=B3, tail_out=B3),
        "1x1 x3 + sfe1,final fp32": mk(B1, lff=B3, gff0=B3, tail=B3, tail_out=B3, sfe1=FP, final=FP),
        "1x1 x3 + sfe1,final fp32 + c2 x3": mk(B1, lff=B3, gff0=B3, tail=B3, tail_out=B3, sfe1=FP, final=FP, csar_c2=B3),
        "1x1 x3 + sfe1,final fp32 + c1,c2 x3": mk(B1, lff=B3, gff0=B3, tail=B3, tail_out=B3, sfe1=FP, final=FP, csar_c2=B3, csar_c1=B3),
        "rdb only x1, rest x3": mk(B3, rdb=B1),
        "rdb+sfe2+gff1 x1, rest x3": mk(B3, rdb=B1, sfe2=B1, gff1=B1),
        "all a-split(2 MMAs)": mk(BA),
        "all w-split(2 MMAs)": mk(BW),
    }
    for xn, x in xs.items():
        with torch.no_grad():
            ref = port.lpsr_forward(x, W)
            for name, cfg in cfgs.items():
                y = forward(x, W, cfg)
                d = (y - ref).abs()
                mse = ((y - ref) ** 2).mean().item()
                print(f"{xn:8s} {name:40s} max {d.max().item():.2e}  mean {d.mean().item():.2e}  psnr-vs-ref {10*np.log10(1/max(mse,1e-30)):.1f}")
    global q
    q = lambda x: x.to(torch.float16).to(torch.float32)
    for xn, x in xs.items():
        with torch.no_grad():
            ref = port.lpsr_forward(x, W)
            y = forward(x, W, mk(B1)); d = (y - ref).abs()
            print(f"{xn:8s} fp16x1 max {d.max().item():.2e} mean {d.mean().item():.2e}")
if __name__ == "__main__":
    more()
