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
# the other kernels of the channels-last voxel branch and the attention block, at the step's shapes
if not a.ncu:
    extra = []
    with torch.no_grad():
        for c in (32, 64):
            x = torch.randn(b, 32, 32, 32, c, device="cuda")
            w, bb = torch.randn(c, device="cuda"), torch.randn(c, device="cuda")
            t = timeit(lambda: B.groupnorm_act_cl(x, 8, w, bb, 1e-5, True, conv_bias=bb))
            nbytes = 3 * x.numel() * 4
            extra.append({"op": f"groupnorm_act_cl C={c} R=32 (stats + apply)", "us": t, "GBps": nbytes / t / 1e3,
                          "bytes": "2 reads + 1 write"})
            xn = x.permute(0, 4, 1, 2, 3).contiguous()
            t = timeit(lambda: B.groupnorm_act(xn, 8, w, bb, 1e-5, True, conv_bias=bb))
            extra.append({"op": f"groupnorm_act C={c} R=32 channel-first (stats + apply)", "us": t, "GBps": nbytes / t / 1e3})
            co = torch.as_tensor(cases.cloud(rng, b, 4096, "shape")).cuda()
            nc = co - co.mean(2, keepdim=True)
            nc = torch.clamp((nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0) + 0.5) * 32, 0, 31).contiguous()
            t = timeit(lambda: B.trilinear_devoxelize_cl(x, nc, 32))
            extra.append({"op": f"trilinear_devoxelize_cl C={c} R=32 N=4096", "us": t,
                          "GBps": (x.numel() * 4 + b * c * 4096 * 4 + b * 12 * 4096) / t / 1e3})
            t = timeit(lambda: B.trilinear_devoxelize_forward(32, False, nc, xn.view(b, c, -1)))
            extra.append({"op": f"trilinear_devoxelize C={c} R=32 N=4096 channel-first (incl. binning)", "us": t})
        x = torch.randn(b, 256, 8, 8, 8, device="cuda")
        w, bb = torch.randn(256, device="cuda"), torch.randn(256, device="cuda")
        t = timeit(lambda: B.groupnorm_act(x, 8, w, bb, 1e-5, True, conv_bias=bb))
        extra.append({"op": "groupnorm_act [16,256,8,8,8] one-pass", "us": t, "GBps": 2 * x.numel() * 4 / t / 1e3})
        q, k, v = (torch.randn(b, 64, 4096, device="cuda") * 0.6 for _ in range(3))
        t = timeit(lambda: B.attention(q, k, v))
        t_torch = timeit(lambda: torch.matmul(v, torch.softmax(torch.matmul(q.transpose(1, 2), k), -1).transpose(1, 2)), iters=5)
        extra.append({"op": "attention C=64 T=4096", "us": t, "torch_us": t_torch, "TFLOPs_fp32_equiv": 4 * b * 64 * 4096 ** 2 / t / 1e6})
    for r_ in extra:
        print(json.dumps(r_), flush=True)
    rows = rows + extra
if rows:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "sparse_conv_bench.json"), "w"), indent=1)
