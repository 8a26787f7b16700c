"""Run one hot-path op a few times at its largest PC^2 shape (for ncu captures).
    python tools/run_op.py voxelize|devoxelize|fps|ball_query|grouping|three_nn|sparse_conv|attention|groupnorm_cl|
                           groupnorm_small|devox_cl [--reps 3] [--batch 16]"""
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
    elif a.op == "sparse_conv":      # compact averages -> (GEMM) -> gather, channels-last output
        c = a.channels or 64
        plan = B.voxel_plan(vox, 32)
        comp = B.avg_voxelize_compact(torch.randn(b, c, 4096, device="cuda"), plan)
        taps = torch.randn(b, 4096, 27 * 64, device="cuda")
        B.sparse_conv3_gather(taps, plan, channels_last=True)
    elif a.op == "attention":
        q, k, v = (torch.randn(b, 64, 4096, device="cuda") * 0.6 for _ in range(3))
        B.attention(q, k, v)
    elif a.op == "groupnorm_cl":
        c = a.channels or 64
        x = torch.randn(b, 32, 32, 32, c, device="cuda")
        w, bb = torch.randn(c, device="cuda"), torch.randn(c, device="cuda")
        part = torch.zeros(b, 128, c, 2, dtype=torch.float64, device="cuda")
        B.groupnorm_act_cl(x, 8, w, bb, 1e-5, True, conv_bias=bb, channel_sums=True, partials=part)
    elif a.op == "groupnorm_mid":    # cluster kernel: 256 KB groups, max over the 32 neighbours
        x = torch.randn(b, 64, 256, 32, device="cuda")
        w, bb = torch.randn(64, device="cuda"), torch.randn(64, device="cuda")
        B.groupnorm_act(x, 8, w, bb, 1e-5, True, conv_bias=bb, max_over_last=True)
    elif a.op == "groupnorm_small":  # one-pass kernel
        x = torch.randn(b, 256, 8, 8, 8, device="cuda")
        w, bb = torch.randn(256, device="cuda"), torch.randn(256, device="cuda")
        B.groupnorm_act(x, 8, w, bb, 1e-5, True, conv_bias=bb)
    elif a.op == "devox_cl":
        c = a.channels or 64
        B.trilinear_devoxelize_cl(torch.randn(b, 32, 32, 32, c, device="cuda"), nc, 32)
    elif a.op == "three_nn":
        c = a.channels or 192
        B.three_nearest_neighbors_interpolate_forward(co, cen, torch.randn(b, c, 1024, device="cuda"))
torch.cuda.synchronize()
