"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]
ik, iv = H.index('Kernel Name'), H.index('Metric Value')
agg = defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > iv:
        agg[r[ik].split('(')[0][-48:]].append(float(r[iv].replace(',', '')))
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':48s} {'n':>4s} {'mean ms':>9s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:48s} {len(v):4d} {sum(v)/len(v)/1e6:9.3f} {100*sum(v)/tot:5.1f}%")
