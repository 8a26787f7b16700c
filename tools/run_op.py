"""Run one hot-path op a few times at its largest PC^2 shape (for ncu captures).
    python tools/run_op.py voxelize|devoxelize|fps|ball_query|grouping|three_nn [--reps 3] [--batch 16]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bdm_b200 import backend as B  # noqa: E402
from tests import cases  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("op")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--regime", default="shape")
ap.add_argument("--channels", type=int, default=0)
a = ap.parse_args()
b = a.batch
rng = np.random.default_rng(1234)
co = torch.as_tensor(cases.cloud(rng, b, 4096, a.regime)).cuda()
nc = co - co.mean(2, keepdim=True)
nc = nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0) + 0.5
nc = torch.clamp(nc * 32, 0, 31).contiguous()
vox = torch.round(nc).to(torch.int32).contiguous()
idx = B.furthest_point_sampling(co, 1024)
cen = B.gather_features_forward(co, idx)
for _ in range(a.reps):
    if a.op == "voxelize":
        c = a.channels or 390
        B.avg_voxelize_forward(torch.randn(b, c, 4096, device="cuda"), vox, 32)
    elif a.op == "devoxelize":
        c = a.channels or 64
        B.trilinear_devoxelize_forward(32, False, nc, torch.randn(b, c, 32768, device="cuda"))
    elif a.op == "fps":
        B.furthest_point_sampling(co, 1024)
    elif a.op == "ball_query":
        B.ball_query(cen, co, 0.1, 32)
    elif a.op == "grouping":
        c = a.channels or 64
        nb = B.ball_query(cen, co, 0.1, 32)
        B.grouping_forward(torch.randn(b, c, 4096, device="cuda"), nb)
    elif a.op == "three_nn":
        c = a.channels or 192
        B.three_nearest_neighbors_interpolate_forward(co, cen, torch.randn(b, c, 1024, device="cuda"))
torch.cuda.synchronize()
