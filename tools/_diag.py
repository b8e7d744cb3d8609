import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import lpsr_b200
from lpsr_b200 import pipeline as pl
from oracle import lpsr_torch_port as port, preprocess_oracle as pre
W = dict(np.load("tests/golden/weights_best_model.npz")); Wt = port.to_torch_weights(W)
frames, boxes = pl.synthetic_clip(3, 3, seed=2)
m16 = lpsr_b200.LPSR(3, 32, 16, 4, 4, None, precision="fp16").load_live_weights(W).to("cuda:0").eval()
m32 = lpsr_b200.LPSR(3, 32, 16, 4, 4, None, precision="fp32").load_live_weights(W).to("cuda:0").eval()
r32 = pl.enhance_frames(m32, frames, boxes); r16 = pl.enhance_frames(m16, frames, boxes)
for a, b in zip(r32, r16):
    x1, y1, x2, y2 = a.box
    raw = frames[a.frame][y1:y2, x1:x2]
    long_img, _ = pl.format_long_plate(raw)
    x = torch.from_numpy(pre.preprocess_for_sr(np.ascontiguousarray(long_img)))
    yref = port.lpsr_forward(x, Wt)
    y16 = m16(x.cuda()).cpu(); y32 = m32(x.cuda()).cpu()
    d8 = np.abs(a.sr_bgr.astype(np.int32) - b.sr_bgr.astype(np.int32))
    e16 = (y16 - yref).abs()
    i = int(e16.argmax()); 
    print(f"plate f{a.frame} r{a.rank}: uint8 diff max {d8.max()} n>2 {(d8>2).sum()//3}; float err fp16 max {float(e16.max()):.3e} p99.9 {float(e16.flatten().kthvalue(int(0.999*e16.numel())).values):.2e} fp32 max {float((y32-yref).abs().max()):.2e}; yref at argmax {float(yref.flatten()[i]):.4f}")
