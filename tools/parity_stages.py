"""Stage-by-stage parity on the GPU box: relative rms / max error of every tap of the CUDA forward against the reference arithmetic
(torch CPU port with the intermediates exposed), and the error of the remaining network when the GPU's tensor at that tap is fed to
the fp32 reference from there on ("error injected up to this stage").  Usage: python tools/parity_stages.py [precision] [case.npz]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.nn.functional as F
import lpsr_b200
from oracle import lpsr_torch_port as port

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
case = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "tests/golden/case_s_b8_64x192.npz")
Wn = dict(np.load(os.path.join(ROOT, "tests/golden/weights_best_model.npz")))
W = port.to_torch_weights(Wn)
torch.set_num_threads(os.cpu_count() or 8)
x = torch.from_numpy(np.load(case)["x"])


def trunk(ae, start=None, inject=None):
    """reference trunk from the AutoEncoder output; `start`/`inject`: replace the named tensor with `inject` and continue from there"""
    T = {}
    def put(name, v):
        if name == start:
            v = inject
        T[name] = v
        return v
    ae = put("ae.out", ae)
    sfe1 = put("rdn.sfe1", port._conv(ae, W, "rdn.shallowF1"))
    h = put("rdn.sfe2", port._conv(sfe1, W, "rdn.shallowF2"))
    feats = []
    for i in range(4):
        h = port.rdb(h, W, f"rdn.rdbs.{i}") if i % 2 == 0 else port.csar(h, W)
        h = put(f"rdn.block{i}", h)
        feats.append(h)
    g0 = put("rdn.gff0", port._conv(torch.cat(feats, 1), W, "rdn.gff.0"))
    g = put("rdn.out", port._conv(g0, W, "rdn.gff.1") + sfe1)
    T["y"] = torch.sigmoid(port._conv(g, W, "final_conv"))
    return T


with torch.no_grad():
    ref = trunk(port.auto_encoder(x, W))
    m = lpsr_b200.LPSR(3, 32, 16, 4, 4, None, precision=prec).load_live_weights(Wn).to("cuda:0").eval()
    y = m(x.to("cuda:0")).cpu()
    e = (y - ref["y"]).abs()
    print(f"{prec} final: max|err| {float(e.max()):.3e} rms {float(e.pow(2).mean().sqrt()):.3e}   env: "
          + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("LPSR_")))
    for name, ch in (("ae.out", 3), ("rdn.sfe1", 32), ("rdn.sfe2", 32), ("rdn.block0", 32), ("rdn.block1", 32), ("rdn.block2", 32),
                     ("rdn.block3", 32), ("rdn.gff0", 32), ("rdn.out", 32)):
        got = m.read_tap(name, x.shape, ch, 1).cpu()
        r = ref[name]
        d = got - r
        inj = trunk(ref["ae.out"], start=name, inject=got)["y"] if name != "rdn.sfe1" else None   # sfe1 also feeds the global residual
        if name == "rdn.sfe1":
            inj = trunk(ref["ae.out"], start=name, inject=got)["y"]
        ei = (inj - ref["y"]).abs()
        print(f"  {name:11s} |ref| max {float(r.abs().max()):9.2f} rms {float(r.pow(2).mean().sqrt()):8.3f}   err max {float(d.abs().max()):.3e} "
              f"rel-rms {float(d.pow(2).mean().sqrt() / r.pow(2).mean().sqrt()):.2e}   -> output err if the rest were exact: max {float(ei.max()):.3e} "
              f"rms {float(ei.pow(2).mean().sqrt()):.3e}")
