"""Summarise an `ncu --page source --csv --print-source sass` dump: the instructions with the most stall samples.

    ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv ; python scripts/ncu_src_top.py src.csv [kernel-index] [top]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 40
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
print("kernels:", [(k, rows[s][1][:60]) for k, s in enumerate(starts)])
s = starts[which]
e = starts[which + 1] if which + 1 < len(starts) else len(rows)
hdr = rows[s + 1]
data = [r for r in rows[s + 2:e] if len(r) == len(hdr)]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for r in data:
    for j in stalls:
        agg[hdr[j][6:]] = agg.get(hdr[j][6:], 0) + int(r[j])
print("stall totals:", dict(sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:ntop]
for i in sorted(top):
    r = data[i]
    st = {hdr[j][6:]: int(r[j]) for j in stalls if int(r[j]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(i, r[isamp], r[iex], r[ia].strip()[:80], st)
