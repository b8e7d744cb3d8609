"""Profiling experiment: per-tile clock64 stamps of the tensor-core conv kernel (CTA 0).  LPSR_UMMA_TRACE=1 python tools/umma_trace.py"""
import ctypes as C, math, os, sys
os.environ.setdefault("LPSR_UMMA_TRACE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
ge.build()
import numpy as np, torch, lpsr_b200
W = dict(np.load(os.path.join(ROOT, "tests/golden/weights_best_model.npz")))
m = lpsr_b200.LPSR(3, 32, 16, 4, 4, None, precision="bf16").load_live_weights(W).to("cuda:0").eval()
lib = lpsr_b200.capi.load_library()
for (ks, cin, cout) in [(3, 32, 16), (3, 80, 16), (3, 32, 32), (1, 96, 32)]:
    B, H, Wd = 256, 64, 192
    x = torch.randn(B, cin, H, Wd, device="cuda:0")
    w = torch.randn(cout, cin, ks, ks, device="cuda:0") / math.sqrt(cin * ks * ks)
    b = torch.randn(cout, device="cuda:0")
    for _ in range(2):
        lpsr_b200.conv2d(m, x, w, b, relu=True)
    torch.cuda.synchronize()
    buf = (C.c_longlong * (256 * 8))()
    rc = lib.lpsr_debug_umma_trace(buf)
    t = np.array(buf[:], dtype=np.int64).reshape(256, 8)
    t = t[(t[:, 2] > 0) & (t[:, 7] > 0)]
    n = len(t)
    if n < 8:
        print(ks, cin, cout, "no trace", rc); continue
    t = t[4:n - 2]
    print(f"conv {ks}x{ks} {cin}->{cout}: {len(t)} tiles traced (CTA 0)")
    print("  MMA warp : wait tempty %6.0f | issue+commit %6.0f | tile period %6.0f clk" % (
        np.mean(t[:, 1] - t[:, 0]), np.mean(t[:, 2] - t[:, 1]), np.mean(np.diff(t[:, 2]))))
    fold = t[:, 6] > 0
    if fold.all():
        print("  epilogue : wait tfull %6.0f | ld+xchg->bar %6.0f | bar wait %6.0f | shuffle+math+store %6.0f | commit->epilogue-done %6.0f clk" % (
            np.mean(t[:, 4] - t[:, 3]), np.mean(t[:, 5] - t[:, 4]), np.mean(t[:, 6] - t[:, 5]), np.mean(t[:, 7] - t[:, 6]), np.mean(t[:, 7] - t[:, 2])))
    else:
        print("  epilogue : wait tfull %6.0f | ld+math+store %6.0f | commit->epilogue-done %6.0f clk" % (
            np.mean(t[:, 4] - t[:, 3]), np.mean(t[:, 7] - t[:, 4]), np.mean(t[:, 7] - t[:, 2])))
    print("  latency commit->tfull seen by epilogue %6.0f clk" % np.mean(t[:, 4] - t[:, 2]))
