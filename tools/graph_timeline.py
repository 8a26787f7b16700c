"""Timeline of one CUDA-graph replay of the PC^2 step (torch profiler / CUPTI): per-stream busy time, and
the idle gaps on the main stream with the kernels either side of them."""
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

x, feats, cams = bench.make_inputs(int(os.environ.get("BDM_BATCH", "32")), 1234, "cuda:0")
sampler = bench.build_sampler(feats, cams, "cuda:0", mode="vanilla")
with torch.no_grad():
    for _ in range(3):
        sampler.pc2_step(x, 500)
    sampler.enable_cuda_graphs(x)
    for _ in range(3):
        sampler.pc2_step(x, 500)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    with torch.no_grad():
        sampler.pc2_step(x, 500)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
print(f"{len(ev)} device activities, span {(t1 - t0) / 1e3:.3f} ms")
streams = {}
for e in ev:
    streams.setdefault(e["args"].get("stream"), []).append(e)
for sid, es in sorted(streams.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
    print(f"stream {sid}: {len(es)} activities, busy {sum(e['dur'] for e in es) / 1e3:.3f} ms")
main = max(streams.values(), key=lambda es: sum(e["dur"] for e in es))
gaps = []
small = sum(max(0, b["ts"] - (a["ts"] + a["dur"])) for a, b in zip(main, main[1:]))
print(f"main stream: total idle between consecutive activities {small / 1e3:.3f} ms")
for a, b in zip(main, main[1:]):
    g = b["ts"] - (a["ts"] + a["dur"])
    if g > 8:
        gaps.append((g, a["name"][:70], b["name"][:70], (a["ts"] - t0) / 1e3))
print(f"main stream: {len(gaps)} gaps > 8 us, total {sum(g[0] for g in gaps) / 1e3:.3f} ms")
for g in sorted(gaps, reverse=True)[:25]:
    print(f"  {g[0]:7.1f} us at {g[3]:6.3f} ms   after {g[1]}   before {g[2]}")
