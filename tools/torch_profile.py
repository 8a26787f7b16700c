"""Where the time of one PC^2 denoiser step goes (torch profiler, CUDA kernel table)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

x, feats, cams = bench.make_inputs(int(os.environ.get("BDM_BATCH", "32")), 1234, "cuda:0")
sampler = bench.build_sampler(feats, cams, "cuda:0", mode="vanilla")
for _ in range(3):
    with torch.no_grad():
        sampler.pc2_step(x, 500)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        with torch.no_grad():
            sampler.pc2_step(x, 500)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
