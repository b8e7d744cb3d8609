"""Profiling experiment: clock64 stamps of the row-streaming conv kernel (CTA 0; MMA warp 0 and epilogue warp 0 of tile 0).
Usage: LPSR_ROW_TRACE=1 [LPSR_UMMA_DEBUG=<mask>] python tools/rowconv_trace.py"""
import ctypes as C, math, os, sys
os.environ.setdefault("LPSR_ROW_TRACE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, lpsr_b200
W = dict(np.load(os.path.join(ROOT, "tests/golden/weights_best_model.npz")))
m = lpsr_b200.LPSR(3, 32, 16, 4, 4, None, precision="fp16").load_live_weights(W).to("cuda:0").eval()
lib = lpsr_b200.capi.load_library()
for (cin, cout) in [(32, 16), (64, 16), (32, 32)]:
    B, H, Wd = 1024, 64, 192
    x = torch.randn(B, cin, H, Wd, device="cuda:0")
    w = torch.randn(cout, cin, 3, 3, device="cuda:0") / math.sqrt(cin * 9)
    b = torch.randn(cout, device="cuda:0")
    for _ in range(2):
        lpsr_b200.conv2d(m, x, w, b, relu=True)
    torch.cuda.synchronize()
    buf = (C.c_longlong * (512 * 8))()
    rc = lib.lpsr_debug_umma_trace(buf)
    t = np.array(buf[:], dtype=np.int64).reshape(512, 8)
    n = int((t[:, 0] > 0).sum())
    if n < 40:
        print(cin, cout, "no trace", rc, n); continue
    a = t[20:n - 20]
    per = np.diff(a[:, 0]).mean()
    print(f"conv 3x3 {cin}->{cout}: {n} steps traced, {per:.0f} clk per input row (MMA warp 0)")
    print("  MMA warp 0: wait staged row %5.0f | wait tile-0 drain %5.0f | issue tile 0 + commit %5.0f | its other tiles %5.0f | loop tail %5.0f" % (
        (a[:, 1] - a[:, 0]).mean(), (a[:, 2] - a[:, 1]).mean(), (a[:, 3] - a[:, 2]).mean(), (a[:, 4] - a[:, 3]).mean(),
        (a[1:, 0] - a[:-1, 4]).mean()))
    e = t[20:n - 24]
    ok = (e[:, 5] > 0) & (e[:, 7] > 0)
    e = e[ok]
    print("  epilogue (tile 0): period %5.0f | ld + reset + arrive %5.0f | convert + store %5.0f | next wait %5.0f" % (
        np.diff(e[:, 5]).mean(), (e[:, 6] - e[:, 5]).mean(), (e[:, 7] - e[:, 6]).mean(), (e[1:, 5] - e[:-1, 7]).mean()))
    # cross-warp latencies (interior of the first crop group only: step s stages input row s, output row g completes at step g + 1)
    g = np.arange(4, 56)
    print("  commit of row g (MMA stamp 3 at step g+1) -> epilogue sees it %5.0f | epilogue arrive -> MMA warp passes the drain wait of step g+2 %5.0f" % (
        (t[g, 5] - t[g + 1, 3]).mean(), (t[g + 2, 2] - t[g, 6]).mean()))
