"""Profiling experiment: per-tile clock64 stamps of the tensor-core conv kernel (CTA 0).
Build with LPSR_NVCC_EXTRA=-DLPSR_UMMA_TRACE_BUILD, then: LPSR_UMMA_TRACE=1 python tools/umma_trace.py"""
import ctypes as C, math, os, sys
os.environ.setdefault("LPSR_UMMA_TRACE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, lpsr_b200
W = dict(np.load(os.path.join(ROOT, "tests/golden/weights_best_model.npz")))
m = lpsr_b200.LPSR(3, 32, 16, 4, 4, None, precision="bf16").load_live_weights(W).to("cuda:0").eval()
lib = lpsr_b200.capi.load_library()
for (ks, cin, cout) in [(3, 32, 16), (3, 64, 16), (3, 32, 32), (1, 96, 32), (1, 128, 32)]:
    B, H, Wd = 256, 64, 192
    x = torch.randn(B, cin, H, Wd, device="cuda:0")
    w = torch.randn(cout, cin, ks, ks, device="cuda:0") / math.sqrt(cin * ks * ks)
    b = torch.randn(cout, device="cuda:0")
    for _ in range(2):
        lpsr_b200.conv2d(m, x, w, b, relu=True)
    torch.cuda.synchronize()
    buf = (C.c_longlong * (512 * 8))()
    rc = lib.lpsr_debug_umma_trace(buf)
    t = np.array(buf[:], dtype=np.int64).reshape(512, 8)
    ok = (t[:, 2] > 0) & (t[:, 6] > 0)
    n = int(ok.sum())
    if n < 16:
        print(ks, cin, cout, "no trace", rc); continue
    t = t[ok][6:n - 3]
    per = (t[-1, 6] - t[0, 6]) / (len(t) - 1)
    print(f"conv {ks}x{ks} {cin}->{cout}: {len(t)} tiles traced (CTA 0), {per:.0f} clk per tile")
    print("  MMA warp : wait tempty %6.0f | issue+commit %6.0f | wait TMA data per tile %6.0f" % (
        np.mean(t[:, 1] - t[:, 0]), np.mean(t[:, 2] - t[:, 1]), np.mean(t[:, 7])))
    print("  epilogue : wait tfull %6.0f | TMEM loads %6.0f | rest of tile %6.0f | commit->tfull seen %6.0f | commit->tile done %6.0f" % (
        np.mean(t[:, 4] - t[:, 3]), np.mean(t[:, 5] - t[:, 4]), np.mean(t[:, 6] - t[:, 5]), np.mean(t[:, 4] - t[:, 2]), np.mean(t[:, 6] - t[:, 2])))
