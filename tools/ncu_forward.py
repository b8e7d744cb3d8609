"""Profiling driver: N forwards of the 16-bit path at batch B on resident inputs (run under ncu).
Usage: python tools/ncu_forward.py [B] [precision] [forwards]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lpsr_b200
from lpsr_b200 import LPSR

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "weights_best_model.npz"))
m = LPSR(3, 32, 16, 4, 4, None, precision=prec).load_live_weights({k: g[k] for k in g.files}).to("cuda").eval()
x = torch.rand(B, 3, 64, 192, device="cuda")
for _ in range(n):
    y = m(x)
torch.cuda.synchronize()
print("ok", float(y.mean()))
