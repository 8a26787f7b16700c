"""Two warm-up steps (allocations, weight preparation) + one measured PC^2 step, eager, single stream: the command
whose launch list is committed under profiles/.  Only the measured step is inside cudaProfilerStart/Stop:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv python tools/step_launches.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import bdm_b200.denoiser as D  # noqa: E402

D.PLAN_AHEAD = False
x, feats, cams = bench.make_inputs(int(os.environ.get("BDM_BATCH", "32")), 1234, "cuda:0")
sampler = bench.build_sampler(feats, cams, "cuda:0", mode="vanilla")
for i in range(3):
    if i == 2:
        torch.cuda.cudart().cudaProfilerStart()
    with torch.no_grad():
        sampler._pc2_eps(x, torch.full((x.shape[0],), 500, device=x.device, dtype=torch.long))
    torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
