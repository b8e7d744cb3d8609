"""Per-launch summary of an `ncu --set full ... --page raw --csv` export: duration, DRAM bytes (as channel planes of the
B x P pixel batch), achieved DRAM GB/s, DRAM %, L2 hit rate, issue-slot utilisation and TENSOR-PIPE ACTIVE % (the cycles the tensor pipe
is busy, sm__pipe_tensor_cycles_active -- what ncu's details page reports; an instruction-count ratio says nothing about tcgen05.mma, which
one thread issues for the whole CTA).
Usage: python tools/ncu_summary.py raw.csv B P"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
B, P = int(sys.argv[2]), int(sys.argv[3])
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
def val(r, name):
    s = r[idx[name]].replace(',', '') if name in idx else ''
    if s == '' or not s.replace('.', '').replace('-', '').replace('e', '').replace('+', '').isdigit(): return float('nan')
    m = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'us': 1, 'ms': 1e3, 'ns': 1e-3}.get(units[idx[name]], 1)
    return float(s) * m
DT = 'dram__throughput.avg.pct_of_peak_sustained_elapsed' if 'dram__throughput.avg.pct_of_peak_sustained_elapsed' in idx else 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'
L2 = 'lts__t_sector_hit_rate.pct'
TP = next((n for n in ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
                       'sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active') if n in idx), None) or \
     next((n for n in hdr if 'pipe_tensor' in n and 'cycles_active' in n and 'pct' in n), 'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active')
SMEM = 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'
tot = 0
print(f"{'kernel':34s} {'us':>8s} {'readMB':>8s} {'writeMB':>8s} {'planes':>7s} {'GB/s':>6s} {'dram%':>6s} {'l2hit%':>6s} {'issue%':>6s} {'TCpipe%':>7s} {'tcSmem%':>7s}")
print("# tensor-pipe metric:", TP)
for r in rows[2:]:
    name = r[idx['Kernel Name']]
    m = re.search(r'umma_conv_kernel<(.*?)>', name) or re.search(r'(rowconv_kernel<.*?>)', name) or re.search(r'(csar_tail_umma_kernel)', name)
    short = m.group(1).replace('__nv_bfloat16', 'bf16').replace('(int)', '').replace('__half', 'f16') if m else name[:32]
    d, w, t = val(r, 'dram__bytes_read.sum'), val(r, 'dram__bytes_write.sum'), val(r, 'gpu__time_duration.sum')
    tot += t
    print(f"{short:34s} {t:8.1f} {d/1e6:8.1f} {w/1e6:8.1f} {(d+w)/(B*P*2):7.1f} {(d+w)/t/1e3:6.0f} {val(r,DT):6.1f} {val(r,L2):6.1f} "
          f"{val(r,'sm__inst_issued.avg.pct_of_peak_sustained_active'):6.1f} {val(r,TP):7.1f} {val(r,SMEM):7.1f}")
print("total us", round(tot, 1))
