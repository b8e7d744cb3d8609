"""Profiling experiment: kernel-level time of one tensor-core conv launch under the LPSR_UMMA_DEBUG ablation switches
(1 skip MMAs, 2 skip stores, 4 skip TMA loads; trace builds add 8 skip shuffles, 16 skip the named barrier, 32 skip the store
function, 64 skip two of the three TMEM loads).  Usage: LPSR_UMMA_DEBUG=<mask> python tools/umma_ablate.py"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, lpsr_b200
W = dict(np.load(os.path.join(ROOT, "tests/golden/weights_best_model.npz")))
m = lpsr_b200.LPSR(3, 32, 16, 4, 4, None, precision="bf16").load_live_weights(W).to("cuda:0").eval()
B, H, Wd = 256, 64, 192
out = []
for (ks, cin, cout) in [(3, 32, 16), (3, 64, 16), (3, 32, 32), (1, 128, 32)]:
    x = torch.randn(B, cin, H, Wd, device="cuda:0")
    w = torch.randn(cout, cin, ks, ks, device="cuda:0") / math.sqrt(cin * ks * ks)
    b = torch.randn(cout, device="cuda:0")
    # op_conv2d = layout conversion + conv + layout conversion: time the whole call and subtract the same call with a 1x1 identity-size conv? simpler:
    # report the call time; the conversions are identical across ablation masks, so differences are the conv kernel's
    for _ in range(2):
        lpsr_b200.conv2d(m, x, w, b, relu=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lpsr_b200.conv2d(m, x, w, b, relu=True)
    e1.record()
    torch.cuda.synchronize()
    out.append(f"{ks}x{ks} {cin}->{cout}: {e0.elapsed_time(e1) / 5 * 1e3:7.1f} us/call")
print("mask", os.environ.get("LPSR_UMMA_DEBUG", "0"), " | ".join(out))
