"""Opcode census of the built library: how many tcgen05 / TMA / TMEM instructions each kernel carries (cuobjdump -sass, no GPU needed).
UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, STG.E.ENL2.256 = 256-bit stores.
Usage: python tools/sass_census.py > profiles/r2_sass_opcode_census.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "license-plate-detection-and-recognition-with-image-enhancement_b200", "liblpsr_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ops = ["UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS", "STG.E.ENL2.256", "FFMA", "SHFL", "BAR.SYNC"]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts.setdefault(cur, collections.Counter())
        continue
    if cur:
        for o in ops:
            if re.search(r"\b" + re.escape(o), line):
                counts[cur][o] += 1
names = subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
agg = collections.OrderedDict()
for n, (k, c) in zip(names, counts.items()):
    short = re.sub(r"\(.*", "", n).replace("lpsr::", "").replace("__nv_bfloat16", "bf16").replace("__half", "f16").replace("(int)", "")
    a = agg.setdefault(short, collections.Counter())
    for o in ops:
        a[o] = max(a[o], c[o])          # the same instantiation appears once per translation unit
print(f"{'kernel (template arguments)':78s} " + " ".join(f"{o[:8]:>8s}" for o in ops))
tot = collections.Counter()
for k, c in agg.items():
    if not any(c.values()):
        continue
    print(f"{k[:78]:78s} " + " ".join(f"{c[o]:8d}" for o in ops))
    tot.update(c)
print(f"{'TOTAL (distinct instantiations)':78s} " + " ".join(f"{tot[o]:8d}" for o in ops))
