"""Non-asserting GPU diagnostic: per-case / per-mode / per-tap errors of the CUDA path vs golden vectors.
Writes gpurun_out/diag.json.  Usage on the GPU box: python tools/gpu_diag.py"""
import json, math, os, sys, time
import numpy as np, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
ge.build()
import lpsr_b200
G = os.path.join(ROOT, "tests", "golden")
W = dict(np.load(os.path.join(G, "weights_best_model.npz")))
out = {"umma_env": os.environ.get("LPSR_UMMA", "1")}
dev = "cuda:0"
print(torch.cuda.get_device_name(0))
cases = sorted(f for f in os.listdir(G) if f.startswith("case_") and "rand" not in f)
for prec in ("fp32", "bf16", "fp16"):
    m = lpsr_b200.LPSR(3, 32, 16, 4, 4, None, precision=prec).load_live_weights(W).to(dev).eval()
    for c in cases:
        d = np.load(os.path.join(G, c))
        try:
            y = m(torch.from_numpy(d["x"]).to(dev)).cpu().numpy()
            err = float(np.abs(y - d["y"]).max()); nan = int(np.isnan(y).sum())
        except Exception as e:
            err, nan = str(e), -1
        out[f"{prec}/{c}"] = {"max_err": err, "nan": nan}
        print(prec, c, err, nan, flush=True)
    t = np.load(os.path.join(G, "taps_u_b1_16x32.npz"))
    x = torch.from_numpy(t["x"]).to(dev)
    y = m(x)
    for name, ch, div in (("ae.c0", 12, 1), ("ae.enc0", 48, 2), ("ae.enc1", 48, 4), ("ae.dec0", 12, 2), ("ae.out", 3, 1),
                          ("rdn.sfe1", 32, 1), ("rdn.sfe2", 32, 1), ("rdn.block0", 32, 1), ("rdn.block1", 32, 1),
                          ("rdn.block2", 32, 1), ("rdn.block3", 32, 1), ("csar3.x_in", 32, 1), ("rdn.out", 32, 1)):
        got = m.read_tap(name, x.shape, ch, div).cpu().numpy()
        e = float(np.abs(got - t[name]).max()); s = float(np.abs(t[name]).max())
        out[f"{prec}/tap/{name}"] = {"max_err": e, "scale": s}
        print(prec, "tap", name, f"{e:.3e} (scale {s:.2f})", flush=True)
# conv op matrix
for prec in ("bf16", "fp16"):
    m = lpsr_b200.LPSR(3, 32, 16, 4, 4, None, precision=prec).load_live_weights(W).to(dev).eval()
    for (ks, cin, cout, B, H, Wd) in [(1, 16, 16, 1, 8, 16), (1, 32, 32, 1, 16, 32), (3, 16, 16, 1, 8, 16), (3, 32, 16, 1, 32, 192), (3, 80, 16, 2, 16, 200),
                                      (3, 32, 32, 1, 64, 192), (1, 96, 32, 2, 32, 192), (1, 128, 32, 1, 64, 192), (1, 32, 64, 1, 24, 40), (3, 32, 32, 64, 64, 192)]:
        g = torch.Generator().manual_seed(1)
        dt = torch.bfloat16 if prec == "bf16" else torch.float16
        x = torch.randn(B, cin, H, Wd, generator=g).to(dt).float()
        w = (torch.randn(cout, cin, ks, ks, generator=g) / math.sqrt(cin * ks * ks)).to(dt).float()
        b = torch.randn(cout, generator=g) * 0.1
        ref = F.conv2d(x, w, b, padding=ks // 2)
        try:
            t0 = time.time()
            got = lpsr_b200.conv2d(m, x.to(dev), w.to(dev), b.to(dev)).cpu()
            e = float((got - ref).abs().max())
            # where is the error? interior vs border
            ei = float((got - ref)[:, :, 1:-1, 1:-1].abs().max()) if H > 2 else -1
        except Exception as ex:
            e, ei = str(ex), -1
        out[f"{prec}/conv/{ks}x{ks}_{cin}_{cout}_{B}x{H}x{Wd}"] = {"max_err": e, "interior": ei, "scale": float(ref.abs().max())}
        print(prec, "conv", ks, cin, cout, B, H, Wd, e, ei, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"diag_umma{out['umma_env']}.json"), "w"), indent=1)
