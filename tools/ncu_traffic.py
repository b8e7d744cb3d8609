"""Write profiles/r2_ncu_traffic.json from the raw-page CSV of an `ncu --set full` capture covering every tensor-core launch of
ONE 16-bit forward in launch order (tools/ncu_forward.py, `-k regex:"umma|csar_tail|rowconv" --launch-skip 25 --launch-count 25`).
Usage: python tools/ncu_traffic.py raw.csv B H W"""
import csv, json, os, sys
raw, B, H, W = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
def gb(r, name):
    v = float(r[idx[name]].replace(",", "")); u = units[idx[name]]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
def us(r):
    v = float(r[idx["gpu__time_duration.sum"]].replace(",", "")); u = units[idx["gpu__time_duration.sum"]]
    return v * {"us": 1.0, "ms": 1e3, "ns": 1e-3}[u]
launches = [{"kernel": r[idx["Kernel Name"]][:80], "us": us(r), "dram": gb(r, "dram__bytes_read.sum") + gb(r, "dram__bytes_write.sum")}
            for r in rows[2:]]
# launch order of one 16-bit forward (forward_impl.cuh): AutoEncoder (conv_in, enc0, enc1, dec0, dec1, conv_out), shallowF1, shallowF2,
# RDB (3 dense layers + fused last layer/lff), CSAR (conv_in.0, conv_in.2 + pool, fused tail), RDB, CSAR, gff.0, gff.1, final conv
names = (["ae.conv_in", "ae.enc0", "ae.enc1", "ae.dec0", "ae.dec1", "ae.conv_out", "sfe1", "sfe2"] + ["rdb0"] * 4 +
         ["csar1.conv_in"] * 2 + ["csar1.tail"] + ["rdb2"] * 4 + ["csar3.conv_in"] * 2 + ["csar3.tail"] + ["gff0", "gff1", "final"])
assert len(launches) == len(names), (len(launches), len(names))
out = {"batch": B, "pixels_per_crop": H * W, "source": os.path.basename(raw),
       "umma_dram_bytes_per_forward": sum(l["dram"] for l in launches),
       "tail_dram_bytes_per_forward": sum(l["dram"] for l, n in zip(launches, names) if n.endswith(".tail")),
       "launches": [dict(l, layer=n) for l, n in zip(launches, names)]}
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_ncu_traffic.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "launches"}))
