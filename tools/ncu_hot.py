"""Hot SASS instructions of one launch from `ncu -i rep --page source --csv --kernel-id :::N` (stall samples per instruction).
Usage: python tools/ncu_hot.py source.csv [min_share]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[ix['# Samples']].isdigit()]
tot = sum(int(r[ix['# Samples']]) for r in data)
print(rows[0][1][:160]); print('total samples', tot, 'instructions', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print(' '.join(f"{k[6:]}={v*100//max(tot,1)}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for i, r in enumerate(data):
    s = int(r[ix['# Samples']]); ex = int(r[ix['Instructions Executed']])
    if s > tot * thr:
        top = sorted(((int(r[ix[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
        print(f"{i:5d} {r[ix['Source']].strip()[:72]:72s} {s*100/tot:5.1f}% ex={ex:8d} {top}")
