"""Parity report on the GPU box: max|err| of every 16-bit mode against the reference arithmetic (torch CPU port) on the golden cases
and on seeded smooth / uniform batches.  Usage: python tools/parity_report.py [precisions...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.nn.functional as F
import lpsr_b200
from oracle import lpsr_torch_port as port

W = dict(np.load(os.path.join(ROOT, "tests/golden/weights_best_model.npz")))
Wt = port.to_torch_weights(W)
precs = sys.argv[1:] or ["fp16", "bf16"]
torch.set_num_threads(os.cpu_count() or 8)

def smooth(b, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    lo = torch.rand(b, 3, max(h // 8, 1), max(w // 8, 1), generator=g)
    return F.interpolate(lo, size=(h, w), mode="bicubic", align_corners=False).clamp(0, 1).contiguous()

sets = {}
for s in range(31, 39):
    sets[f"smooth64x192_s{s}"] = smooth(8, 64, 192, s)
for s in range(41, 45):
    sets[f"smooth32x192_s{s}"] = smooth(8, 32, 192, s)
sets["smooth30x190"] = smooth(8, 30, 190, 51)
sets["uniform64x192"] = torch.rand(8, 3, 64, 192, generator=torch.Generator().manual_seed(0))
g = torch.Generator().manual_seed(3)
clean = smooth(8, 64, 192, 61)
sets["smooth+noise"] = (clean + 0.05 * torch.randn(clean.shape, generator=g)).clamp(0, 1)
sets["stripes"] = (torch.arange(192)[None, :] // 8).remainder(2).float().expand(2, 3, 64, 192).contiguous()
refs = {k: port.lpsr_forward(v, Wt) for k, v in sets.items()}
for prec in precs:
    m = lpsr_b200.LPSR(3, 32, 16, 4, 4, None, precision=prec).load_live_weights(W).to("cuda:0").eval()
    worst = 0.0
    errs = []
    for k, x in sets.items():
        y = m(x.to("cuda:0")).cpu()
        d = (y - refs[k]).abs()
        e = float(d.max())
        worst = max(worst, e)
        errs.append(d.flatten())
        if os.environ.get("PARITY_VERBOSE", "1") == "1":
            print(f"{prec} {k:22s} max|err| {e:.3e}  mean {float(d.mean()):.2e}", flush=True)
    allerr = torch.cat(errs)
    n = allerr.numel()
    q = lambda f: float(allerr.kthvalue(max(1, int(f * n))).values)
    print(f"{prec} SUMMARY worst {worst:.3e}  p99.99 {q(0.9999):.3e}  p99.9 {q(0.999):.3e}  rms {float((allerr.double() ** 2).mean().sqrt()):.3e}  "
          f"n>5e-3 {int((allerr > 5e-3).sum())}  n>1e-2 {int((allerr > 1e-2).sum())}  of {n}")
