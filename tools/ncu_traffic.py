"""Write profiles/r1_ncu_traffic.json from an `ncu --set full` report of `bench.py --batch B --steps 1` covering every
tensor-core conv launch of ONE forward in launch order.  Usage: python tools/ncu_traffic.py rep.ncu-rep B H W"""
import csv, io, json, os, subprocess, sys
rep, B, H, W = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
def gb(r, name):
    v = float(r[idx[name]]); u = units[idx[name]]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
launches = []
for r in rows[2:]:
    launches.append({"kernel": r[idx["Kernel Name"]][:60], "us": float(r[idx["gpu__time_duration.sum"]]),
                     "dram": gb(r, "dram__bytes_read.sum") + gb(r, "dram__bytes_write.sum")})
# launch order of one bf16 forward (forward_impl.cuh): ae.conv_out, 7x7, sfe2, rdb0 x5, csar1 (c1, c2, sa1, gate, co), rdb2 x5,
# csar3 x5, gff0, gff1, final  -> tail launches are the 3 after each pair of csar conv_in launches
names = (["ae.conv_out", "sfe1", "sfe2"] + ["rdb0"] * 5 + ["csar1.conv_in"] * 2 + ["csar1.tail"] * 3 + ["rdb2"] * 5 +
         ["csar3.conv_in"] * 2 + ["csar3.tail"] * 3 + ["gff0", "gff1", "final"])
assert len(launches) == len(names), (len(launches), len(names))
out = {"batch": B, "pixels_per_crop": H * W, "source": os.path.basename(rep),
       "umma_dram_bytes_per_forward": sum(l["dram"] for l in launches),
       "tail_dram_bytes_per_forward": sum(l["dram"] for l, n in zip(launches, names) if n.endswith(".tail")),
       "launches": [dict(l, layer=n) for l, n in zip(launches, names)]}
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r1_ncu_traffic.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "launches"}))
