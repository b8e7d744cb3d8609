"""Profiling experiment: one forward as stream launches vs. as a replayed CUDA graph (same kernels, same buffers)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, lpsr_b200
W = dict(np.load(os.path.join(ROOT, "tests/golden/weights_best_model.npz")))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
m = lpsr_b200.LPSR(3, 32, 16, 4, 4, None, precision="bf16").load_live_weights(W).to("cuda:0").eval()
x = torch.rand(B, 3, 64, 192, device="cuda:0")
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
with torch.no_grad():
    y_ref = m(x).clone()
    t_stream = timeit(lambda: m(x))
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2): m(x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        y_g = m(x)
    t_graph = timeit(g.replay)
    print(f"B={B}: stream launches {t_stream:.3f} ms, graph replay {t_graph:.3f} ms, equal={torch.equal(y_g, y_ref)}")
