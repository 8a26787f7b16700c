"""Per-op timing of the hot-path kernels at the PC^2 denoiser's shapes, ours and
-- when oracle/_ref is present -- the reference kernels recompiled for sm_100a.

    python tools/op_bench.py [--batch 16] [--regime shape] [--iters 20] [--no-ref] [--out gpurun_out/op_bench.json]

Timing: CUDA events on the current stream around `iters` back-to-back calls after 3 warm-ups; an
L2 flush (write of a 256 MB buffer) precedes every timed batch and the inputs of the big calls exceed
L2 anyway.  Algorithmic bytes per call follow SURVEY.md section 8d.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import cases  # noqa: E402

PEAK = 6556.5
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

VOX = [(390, 4096, 32), (32, 4096, 32), (128, 1024, 16), (192, 256, 8), (256, 64, 8), (256, 256, 8),
       (128, 1024, 16), (64, 4096, 32)]
VOX_MULT = [1, 1, 1, 1, 3, 3, 2, 2]
DEV = [(32, 4096, 32), (64, 1024, 16), (128, 256, 8), (256, 64, 8), (256, 256, 8), (128, 1024, 16), (64, 4096, 32)]
DEV_MULT = [2, 1, 1, 3, 3, 2, 2]
SA = [(4096, 1024, 0.1), (1024, 256, 0.2), (256, 64, 0.4), (64, 16, 0.8)]
GRP_C = [(3, 32, 64), (3, 64, 64), (3, 128, 64), (3, 320, 64)]
NN = [(576, 16, 64), (64, 16, 64), (320, 64, 256), (64, 64, 256), (320, 256, 1024), (64, 256, 1024),
      (192, 1024, 4096), (64, 1024, 4096)]

_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    _flush.fill_(1)


USE_GRAPH = True


def timeit(fn, iters, warm=3, graph=None):
    """Device time per call.  With USE_GRAPH the `iters` calls are captured into one CUDA graph and
    replayed, so Python / ctypes / allocator time between launches is not measured."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if USE_GRAPH if graph is None else graph:
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(iters):
                    fn()
            g.replay()
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                flush_l2()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) / iters)
            del g
            return min(ts)
        except Exception as e:  # e.g. an op that cannot be captured: fall back to eager timing
            print("graph capture failed:", repr(e)[:200], flush=True)
            torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / iters)
    return min(ts)  # ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--regime", default="shape")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--eager", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "op_bench.json"))
    args = ap.parse_args()
    B = args.batch
    global USE_GRAPH
    USE_GRAPH = not args.eager
    from bdm_b200 import backend as ours
    ref = None
    if not args.no_ref:
        from oracle import build_ref
        ref = build_ref.load_ref()
    impls = [("ours", ours)] + ([("ref", ref)] if ref is not None else [])
    rng = np.random.default_rng(1234)
    rows = []

    def add(op, shape, nbytes, fns, mult=1, work=None):
        row = dict(op=op, shape=shape, MB=nbytes / 1e6, calls_per_forward=mult)
        for name, fn in fns.items():
            # the reference launches half of its kernels on the legacy default stream (vox.cu:114,
            # trilinear_devox.cu:167, sampling.cu:171), which a stream capture does not see: eager only
            ms = timeit(fn, args.iters, graph=False if name == "ref" else None)
            row[f"{name}_ms"] = ms
            row[f"{name}_GBs"] = nbytes / ms / 1e6
            row[f"{name}_frac"] = nbytes / ms / 1e6 / PEAK
            if work is not None:
                row[f"{name}_Gtests_s"] = work / ms / 1e6
        rows.append(row)
        print(json.dumps(row), flush=True)

    # point pyramid (real FPS chain so ball query / 3-NN see realistic inputs)
    co = torch.as_tensor(cases.cloud(rng, B, 4096, args.regime)).cuda()
    levels = [co]
    for (n, m, rad) in SA:
        idx = ours.furthest_point_sampling(levels[-1], m)
        levels.append(ours.gather_features_forward(levels[-1], idx))

    for (c, n, r), mult in zip(VOX, VOX_MULT):
        lvl = {4096: 0, 1024: 1, 256: 2, 64: 3}[n]
        pts = levels[lvl]
        nc = pts - pts.mean(2, keepdim=True)
        nc = nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0) + 0.5
        nc = torch.clamp(nc * r, 0, r - 1)
        vox = torch.round(nc).to(torch.int32).contiguous()
        feat = torch.randn(B, c, n, device="cuda")
        nbytes = B * (4 * c * n + 12 * n + 4 * c * r ** 3 + 4 * n + 4 * r ** 3)
        add("avg_voxelize", (c, n, r), nbytes,
            {k: (lambda be=be: be.avg_voxelize_forward(feat, vox, r)) for k, be in impls}, mult)
        plan = ours.voxel_plan(vox, r)   # what the step does: one plan per (coords, R), shared by 2-3 calls
        add("avg_voxelize_fill (plan reused)", (c, n, r), nbytes, {"ours": lambda: ours.avg_voxelize_fill(feat, plan)}, 0)
        del feat
    for (c, n, r), mult in zip(DEV, DEV_MULT):
        lvl = {4096: 0, 1024: 1, 256: 2, 64: 3}[n]
        pts = levels[lvl]
        nc = pts - pts.mean(2, keepdim=True)
        nc = nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0) + 0.5
        nc = torch.clamp(nc * r, 0, r - 1).contiguous()
        grid = torch.randn(B, c, r ** 3, device="cuda")
        nbytes = B * (12 * n + 4 * c * min(r ** 3, 8 * n) + 4 * c * n)
        add("trilinear_devoxelize", (c, n, r), nbytes,
            {k: (lambda be=be: be.trilinear_devoxelize_forward(r, False, nc, grid)) for k, be in impls}, mult)
        dplan = ours.devoxelize_plan(nc, r)
        add("trilinear_devoxelize (plan reused)", (c, n, r), nbytes,
            {"ours": lambda: ours.trilinear_devoxelize_forward(r, False, nc, grid, dplan)}, 0)
        del grid
    for li, (n, m, rad) in enumerate(SA):
        pts, cen = levels[li], levels[li + 1]
        add("furthest_point_sampling", (n, m), B * (12 * n + 4 * m),
            {k: (lambda be=be: be.furthest_point_sampling(pts, m)) for k, be in impls}, 1, work=B * (m - 1) * n)
        add("ball_query", (m, n, rad), B * (12 * (n + m) + 4 * m * 32),
            {k: (lambda be=be: be.ball_query(cen, pts, rad, 32)) for k, be in impls}, 1, work=B * m * n)
        nb = ours.ball_query(cen, pts, rad, 32)
        for c in GRP_C[li]:
            feat = torch.randn(B, c, n, device="cuda")
            nbytes = B * (4 * m * 32 + 4 * c * min(n, m * 32) + 4 * c * m * 32)
            add("grouping", (c, n, m, 32), nbytes,
                {k: (lambda be=be: be.grouping_forward(feat, nb)) for k, be in impls}, 1)
    for (c, m, n) in NN:
        lvl = {4096: 0, 1024: 1, 256: 2, 64: 3}[n]
        pts, cen = levels[lvl], levels[lvl + 1]
        feat = torch.randn(B, c, m, device="cuda")
        nbytes = B * (12 * (n + m) + 4 * c * m + 4 * c * n + 24 * n)
        add("three_nn_interpolate", (c, m, n), nbytes,
            {k: (lambda be=be: be.three_nearest_neighbors_interpolate_forward(pts, cen, feat)) for k, be in impls},
            1, work=B * n * m)

    tot = {}
    for name, _ in impls:
        tot[name] = sum(r[f"{name}_ms"] * r["calls_per_forward"] for r in rows if f"{name}_ms" in r)
    summary = dict(batch=B, regime=args.regime, peak_GBs=PEAK, total_ms_per_forward=tot,
                   total_MB=sum(r["MB"] * r["calls_per_forward"] for r in rows))
    print(json.dumps(summary))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(dict(summary=summary, rows=rows), open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
