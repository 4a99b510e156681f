"""Hot instructions of one kernel from an .ncu-rep (source page): python scripts/ncu_src.py REP KERNEL_REGEX [TOP]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
si, so, ie = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
ci = [hdr.index(c) for c in cols]
tot = sum(int(r[si] or 0) for r in data)
print("total samples", tot, "instructions", sum(int(r[ie] or 0) for r in data))
for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][si] or 0))[:top]):
    r = data[i]
    st = {c[6:]: int(r[j] or 0) for c, j in zip(cols, ci) if int(r[j] or 0) > 0.1 * int(r[si] or 1)}
    print(i, r[si], r[ie], r[so][:80], st)
if len(sys.argv) > 5:
    a, b = int(sys.argv[4]), int(sys.argv[5])
    agg = {}
    for r in data[a:b]:
        for c, j in zip(cols, ci):
            agg[c[6:]] = agg.get(c[6:], 0) + int(r[j] or 0)
    print("region", a, b, "samples", sum(int(r[si] or 0) for r in data[a:b]), {k: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
