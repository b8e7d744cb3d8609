"""Summarise an ncu report: per-kernel headline metrics and the top stall sites (source page).
Usage: python tools/ncu_top.py gpurun_out/prof.ncu-rep [kernel-index] [n-lines]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else None
nlines = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__cycles_elapsed.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
for k, r in enumerate(rows[2:]):
    print(f"[{k}] {r[idx['Kernel Name']][:80]}  grid {r[idx['Grid Size']]}")
    for w in want:
        if w in idx:
            print(f"      {w[:75]:75s} {r[idx[w]][:16]:>16s} {units[idx[w]]}")
if kidx is None:
    sys.exit(0)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(kidx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > idx["# Samples"]]
def num(x):
    try: return float(x)
    except: return 0.0
seen, uniq = set(), []
for r in data:
    if r[idx["Address"]] in seen: continue
    seen.add(r[idx["Address"]]); uniq.append(r)
tot = sum(num(r[idx["# Samples"]]) for r in uniq)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(f"--- kernel {kidx}: {rows[0][1][:80]}  total samples {tot:.0f}")
for r in sorted(uniq, key=lambda r: -num(r[idx["# Samples"]]))[:nlines]:
    s = num(r[idx["# Samples"]])
    st = sorted(((num(r[idx[h]]), h[6:]) for h in stalls if idx[h] < len(r)), reverse=True)[:2]
    print(f"{s:7.0f} {100*s/tot:5.1f}%  {r[idx['Address']][-5:]}  {r[idx['Source']][:60]:60s} {st}")
