"""Per-call breakdown of one PC^2 sampler iteration (eager, single stream, per-op CUDA events): every
libbdm_b200 op grouped by (op, argument shapes), sorted by total time.
    BDM_BATCH=32 python tools/op_profile.py > gpurun_out/op_profile.md"""
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import bdm_b200.denoiser as D  # noqa: E402
from bdm_b200 import backend  # noqa: E402

D.PLAN_AHEAD = False
b = int(os.environ.get("BDM_BATCH", "32"))
x, feats, cams = bench.make_inputs(b, 1234, "cuda:0")
sampler = bench.build_sampler(feats, cams, "cuda:0", mode="vanilla")
tt = torch.full((b,), 500, device="cuda:0", dtype=torch.long)
steps = 3
with torch.no_grad():
    sampler._pc2_eps(x, tt)
    torch.cuda.synchronize()
    backend.profile_start()
    for _ in range(steps):
        sampler._pc2_eps(x, tt)
    prof = backend.profile_stop()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sampler._pc2_eps(x, tt)
    e1.record()
    torch.cuda.synchronize()
ms_iter = e0.elapsed_time(e1) / steps
groups = collections.OrderedDict()
for op, calls in prof.items():
    for ms, shp in calls:
        key = op + " " + json.dumps(shp, default=lambda o: type(o).__name__)
        g = groups.setdefault(key, [op, shp, 0, 0.0])
        g[2] += 1
        g[3] += ms
tot = sum(g[3] for g in groups.values()) / steps
print(f"# PC^2 iteration, B={b}: {ms_iter:.3f} ms eager; libbdm_b200 ops {tot:.3f} ms")
print("| op + shapes | calls/iter | us/call | us/iter | GB/s or TFLOP/s |")
print("|---|---|---|---|---|")
for key, (op, shp, n, ms) in sorted(groups.items(), key=lambda kv: -kv[1][3]):
    work = bench.algorithmic_work(op, shp)
    rate = ""
    if work is not None:
        bound, units, unit = work
        per_call_s = ms / n * 1e-3
        rate = f"{units / per_call_s / (1e9 if bound == 'hbm' else 1e12):.0f} {'GB/s' if bound == 'hbm' else 'TFLOP/s'}"
    print(f"| {key[:150]} | {n // steps} | {ms / n * 1e3:.1f} | {ms / steps * 1e3:.1f} | {rate} |")
