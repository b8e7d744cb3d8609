"""Side-by-side per-layer times of two bench.py JSON lines (A/B of kernel variants on the same box).
Usage: python tools/ab_compare.py a.jsonl b.jsonl"""
import json, sys
def load(p):
    for l in open(p):
        l = l.strip()
        if l.startswith('{'):
            return json.loads(l)
    raise SystemExit(f"{p}: no JSON line")
a, b = load(sys.argv[1]), load(sys.argv[2])
print(f"{'':34s} {'A':>9s} {'B':>9s} {'B/A':>6s}")
print(f"{'ms_per_step':34s} {a['ms_per_step']:9.3f} {b['ms_per_step']:9.3f} {b['ms_per_step']/a['ms_per_step']:6.3f}   clocks {a['clocks']['sm_mhz']} / {b['clocks']['sm_mhz']}")
print(f"{'e2e crops/s':34s} {a['e2e']['value']:9.0f} {b['e2e']['value']:9.0f}")
la, lb = a['layer_ms_per_forward'], b['layer_ms_per_forward']
for k in la:
    if k in lb:
        print(f"{k:34s} {la[k]:9.4f} {lb[k]:9.4f} {lb[k]/la[k]:6.3f}")
print(f"{'sum of layers':34s} {sum(la.values()):9.3f} {sum(lb.values()):9.3f}")
