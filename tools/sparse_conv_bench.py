"""Time the pieces of the sparse first convolution (compact -> GEMM -> gather) against the dense route
(avg_voxelize fill -> cuDNN Conv3d) on the shapes of the PC^2 step.   python tools/sparse_conv_bench.py [--ncu]"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bdm_b200 import backend as B  # noqa: E402
from tests import cases  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ncu", action="store_true", help="run each piece twice, no timing (for ncu captures)")
ap.add_argument("--batch", type=int, default=16)
a = ap.parse_args()
b = a.batch


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


rng = np.random.default_rng(1234)
rows = []
for (cin, cout, n, r) in [(390, 32, 4096, 32), (32, 32, 4096, 32), (160, 64, 4096, 32), (64, 64, 4096, 32),
                          (64, 64, 1024, 16), (128, 128, 1024, 16)]:
    co = torch.as_tensor(cases.cloud(rng, b, n, "shape")).cuda()
    nc = co - co.mean(2, keepdim=True)
    nc = nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0) + 0.5
    vox = torch.round(torch.clamp(nc * r, 0, r - 1)).to(torch.int32).contiguous()
    plan = B.voxel_plan(vox, r)
    feats = torch.randn(b, cin, n, device="cuda")
    conv = nn.Conv3d(cin, cout, 3, padding=1).cuda().eval()
    wt = conv.weight.detach().permute(1, 2, 3, 4, 0).reshape(cin, -1).contiguous()
    torch.backends.cuda.matmul.allow_tf32 = True
    with torch.no_grad():
        compact = B.avg_voxelize_compact(feats, plan)
        taps = torch.matmul(compact.transpose(1, 2), wt)
        if a.ncu:
            for _ in range(2):
                B.avg_voxelize_compact(feats, plan)
                B.sparse_conv3_gather(taps, plan)
            continue
        t_compact = timeit(lambda: B.avg_voxelize_compact(feats, plan))
        t_gemm = timeit(lambda: torch.matmul(compact.transpose(1, 2), wt))
        t_gather = timeit(lambda: B.sparse_conv3_gather(taps, plan))
        t_fill = timeit(lambda: B.avg_voxelize_fill(feats, plan))
        grid = B.avg_voxelize_fill(feats, plan).view(b, cin, r, r, r)
        t_conv = timeit(lambda: conv._conv_forward(grid, conv.weight, None))
        nocc = (plan.cnt > 0).sum().item() / b
    row = {"cin": cin, "cout": cout, "n": n, "r": r, "occupied_per_shape": nocc, "compact_us": t_compact, "gemm_us": t_gemm,
           "gather_us": t_gather, "sparse_total_us": t_compact + t_gemm + t_gather, "fill_us": t_fill, "conv3d_us": t_conv,
           "dense_total_us": t_fill + t_conv,
           "gather_GBps": (b * cout * r ** 3 * 4 + b * nocc * 27 * cout * 4) / t_gather / 1e3}
    rows.append(row)
    print(json.dumps(row), flush=True)
if rows:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "sparse_conv_bench.json"), "w"), indent=1)
