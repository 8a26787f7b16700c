"""Graph-replay time of one PC^2 sampler iteration (B shapes) under a few global switches -- a quick A/B harness.
    BDM_BATCH=32 python tools/iter_time.py [cudnn_benchmark]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

if "cudnn_benchmark" in sys.argv:
    torch.backends.cudnn.benchmark = True
b = int(os.environ.get("BDM_BATCH", "32"))
x, feats, cams = bench.make_inputs(b, 1234, "cuda:0")
sampler = bench.build_sampler(feats, cams, "cuda:0", mode="vanilla")
sampler.enable_cuda_graphs(x)
g = sampler._graphs[("pc2", tuple(x.shape))]
g.run(x, 700, 5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g.x.copy_(x)
g.t.fill_(700)
e0.record()
for _ in range(40):
    g.graph.replay()
e1.record()
torch.cuda.synchronize()
print(f"{' '.join(sys.argv[1:]) or 'default':30s} B={b}: {e0.elapsed_time(e1) / 40:.3f} ms per PC^2 iteration", flush=True)
