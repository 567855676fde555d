"""Aggregate an `ncu --page source --print-source sass,cuda --csv` dump per CUDA source line.
usage: python tools/ncu_lines.py dump.csv [top]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, cur_line, cur_src = None, None, None, None
agg = defaultdict(lambda: [0, 0, 0])   # (file,line,src) -> [inst, thread_inst, samples]
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        i_inst = hdr.index("Instructions Executed"); i_thr = hdr.index("Thread Instructions Executed"); i_smp = hdr.index("# Samples")
        continue
    if r[0] != "":
        cur_line, cur_src = r[0], r[1].strip()
        continue
    try:
        key = (cur_file, int(cur_line), cur_src)
        agg[key][0] += int(r[i_inst]); agg[key][1] += int(r[i_thr]); agg[key][2] += int(r[i_smp])
    except (ValueError, TypeError):
        pass
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[2] for v in agg.values())
print(f"total warp-inst {tot_i:,}  samples {tot_s:,}")
for (f, l, s), v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"{100*v[2]/max(tot_s,1):5.1f}% smp {100*v[0]/max(tot_i,1):5.1f}% inst  thr/inst {v[1]/max(v[0],1):4.1f}  {f}:{l}  {s[:110]}")
