"""GPU check of the fused RDB launch (csrc/rdb_chain.cuh): the forward with LPSR_RDB_CHAIN=1 must equal the layer-by-layer forward
(LPSR_RDB_CHAIN=0) BIT FOR BIT -- same MMAs in the same order per pixel, only the tiling and the path of g0..g2 (L2 scratch) differ.
    LPSR_RDB_CHAIN=0 python tools/chain_check.py save /tmp/ref.pt ; LPSR_RDB_CHAIN=1 python tools/chain_check.py compare /tmp/ref.pt"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import lpsr_b200

W = dict(np.load(os.path.join(ROOT, "tests/golden/weights_best_model.npz")))
shapes = [(2, 64, 192), (3, 32, 192), (1, 128, 384), (2, 30, 190), (1, 4, 4), (5, 20, 36), (2, 8, 1024), (1, 256, 8), (300, 64, 192)]
mode, path = sys.argv[1], sys.argv[2]
outs = {}
for prec in ("fp16", "bf16"):
    m = lpsr_b200.LPSR(3, 32, 16, 4, 4, None, precision=prec).load_live_weights(W).to("cuda:0").eval()
    for (b, h, w) in shapes:
        x = torch.rand(b, 3, h, w, generator=torch.Generator().manual_seed(b * 1000 + h + w)).to("cuda:0")
        y = m(x)
        y2 = m(x)
        torch.cuda.synchronize()
        assert torch.equal(y, y2), f"run-to-run mismatch {prec} {b,h,w}"
        outs[f"{prec}_{b}x{h}x{w}"] = y.cpu()
        print(prec, (b, h, w), "launches", m.launch_count(b, h, w), "finite", bool(torch.isfinite(y).all()), flush=True)
if mode == "save":
    torch.save(outs, path)
else:
    ref = torch.load(path)
    bad = 0
    for k, v in outs.items():
        d = (v - ref[k]).abs().max().item()
        print(f"{k:22s} max|chain - layerwise| = {d:.3e}")
        bad += d != 0.0
    print("CHAIN_CHECK", "OK" if not bad else f"MISMATCH in {bad} cases")
