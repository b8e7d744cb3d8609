"""Warp-stall breakdown of one kernel from `ncu --set full --import-source on ... --page source --csv` (+ the details page):
samples per stall reason over the whole kernel and over its hottest instructions.  Usage: python tools/ncu_stalls.py source.csv details.txt"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
name = rows[0][1]
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
n = 0
for r in data:
    n += int(r[idx["# Samples"]])
    for h in reasons:
        tot[h] += int(r[idx[h]])
print(name[:110])
if len(sys.argv) > 2:
    for line in open(sys.argv[2]):
        if re.search(r"Duration|Issue Slots Busy|Eligible Warps Per Scheduler|Warp Cycles Per Issued Instruction|Registers Per Thread|Dynamic Shared Memory Per Block|DRAM Throughput", line):
            print("  " + " ".join(line.split()))
print(f"  warp-stall samples: {n}")
for h, v in tot.most_common(10):
    print(f"    {h:26s} {v:6d}  {100.0 * v / max(n, 1):5.1f} %")
base = int(data[0][0], 16)
print("  hottest instructions (offset, samples, SASS):")
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:12]:
    print(f"    {int(r[0], 16) - base:#07x} {int(r[idx['# Samples']]):5d}  {r[1].strip()[:90]}")
