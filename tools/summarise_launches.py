"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
cols = rows[hdr]
ik, im, iv = cols.index("Kernel Name"), cols.index("Metric Name"), cols.index("Metric Value")
agg = defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > iv and r[im] == "gpu__time_duration.sum":
        agg[r[ik].split("(")[0].replace("void ", "").replace("bgp::", "")].append(float(r[iv].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':40s} {'launches':>8s} {'mean us':>10s} {'total us':>10s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:40]:40s} {len(v):8d} {sum(v)/len(v)/1e3:10.1f} {sum(v)/1e3:10.1f} {100*sum(v)/tot:6.1f}%")
print(f"{'total':40s} {sum(len(v) for v in agg.values()):8d} {'':10s} {tot/1e3:10.1f}")
