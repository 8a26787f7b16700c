"""Aggregate an ncu source page (sass,cuda correlation) by CUDA source line.
    ncu -i X.ncu-rep --page source --print-source sass,cuda --csv -k regex:KERNEL | python tools/ncu_hot_lines.py [N]"""
import collections
import csv
import sys

rows = list(csv.reader(sys.stdin))
agg = collections.OrderedDict()
cur_file, hdr = None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed")
        ws = hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr is None or len(r) <= ie or not r[0].isdigit():
        continue
    key = (cur_file, int(r[0]), r[1].strip())
    a = agg.setdefault(key, [0.0, 0.0])
    try:
        a[0] += float(r[ie] or 0)
        a[1] += float(r[ws] or 0)
    except ValueError:
        pass
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[1] for a in agg.values()) or 1
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
print(f"total warp-instructions {tot_i:.0f}, stall samples {tot_s:.0f}")
for (f, ln, src), (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:n]:
    print(f"{i / tot_i * 100:5.1f}% inst {s / tot_s * 100:5.1f}% stall  {f}:{ln}  {src[:100]}")
