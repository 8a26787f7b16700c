"""Reduce an ncu launch list (--metrics gpu__time_duration.sum --csv) of tools/step_launches.py to the
per-kernel share of ONE step (the second half of the list = the measured step).
    python tools/launch_share.py gpurun_out/prof/step_launches.csv > profiles/r01_step_launch_share.md"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hdr_i]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = []
for r in rows[hdr_i + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    unit = r[ui]
    us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    launches.append((r[ki], us))
half = launches     # (the step list holds exactly one iteration: tools/step_launches.py brackets it with cudaProfilerStart/Stop)
agg = collections.OrderedDict()
for name, us in half:
    short = re.sub(r"\(.*", "", name).replace("void ", "")
    short = re.sub(r"<.*", "", short) if not short.startswith("bdm::") else short
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += us
total = sum(a[1] for a in agg.values())
OURS = ("bdm::", "cv3::", "tc05::", "gnc::")      # namespaces of libbdm_b200.so's kernels as ncu prints them
ours = sum(a[1] for k, a in agg.items() if k.startswith(OURS))
print(f"# {'window of the bench command (graph replays)' if (len(sys.argv) > 2 and sys.argv[2] == 'all') else 'One PC^2 iteration, eager'} (B={__import__('os').environ.get('BDM_BATCH', '32')}, N=4096), serialised under ncu: {len(half)} launches, {total / 1e3:.2f} ms of kernel time")
print(f"# libbdm_b200 kernels: {ours / 1e3:.3f} ms = {ours / total * 100:.1f} % of the step's kernel time")
print("| kernel | launches | total us | share % |")
print("|---|---|---|---|")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"| {k[:90]} | {n} | {us:.1f} | {us / total * 100:.2f} |")
