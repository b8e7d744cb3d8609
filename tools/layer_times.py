"""Per-launch CUDA-event times of one forward (lpsr_forward_profiled), in launch order.
Usage: [LPSR_UMMA_DEBUG=<mask>] [LPSR_ROWCONV=0] python tools/layer_times.py [--batch 1024] [--precision fp16] [--filter rowconv]"""
import argparse, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, lpsr_b200
from lpsr_b200 import capi

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--height", type=int, default=64)
ap.add_argument("--width", type=int, default=192)
ap.add_argument("--precision", default="fp16")
ap.add_argument("--filter", default="")
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
W = dict(np.load(os.path.join(ROOT, "tests/golden/weights_best_model.npz")))
dev = torch.device("cuda:0")
m = lpsr_b200.LPSR(3, 32, 16, 4, 4, None, precision=a.precision).load_live_weights(W).to(dev).eval()
x = torch.rand(a.batch, 3, a.height, a.width, device=dev)
with torch.no_grad():
    y = m(x)
torch.cuda.synchronize()
lib = capi.load_library()
h = m._handle(dev)
ws = m._workspace(h, 0, a.batch, a.height, a.width)
cap = 64
acc, names = None, []
for _ in range(a.reps):
    ms = (C.c_float * cap)()
    nm = C.create_string_buffer(cap * 64)
    n = lib.lpsr_forward_profiled(h, x.data_ptr(), y.data_ptr(), a.batch, a.height, a.width, m._aligned_ptr(ws),
                                  ws.numel() - (m._aligned_ptr(ws) - ws.data_ptr()), torch.cuda.current_stream(dev).cuda_stream, ms, nm, 64, cap)
    capi.check(min(n, 0), h, "lpsr_forward_profiled")
    names = [nm.raw[i * 64:(i + 1) * 64].split(b"\0")[0].decode() for i in range(n)]
    acc = [ms[i] if acc is None else acc[i] + ms[i] for i in range(n)]
tot = sum(acc) / a.reps
print(f"mask {os.environ.get('LPSR_UMMA_DEBUG', '0')} rowconv {os.environ.get('LPSR_ROWCONV', '1')} total {tot:.3f} ms | " +
      " ".join(f"{nme.split(':')[0].split('.')[-1]}:{t / a.reps * 1e3:.0f}" for nme, t in zip(names, acc) if a.filter in nme))
