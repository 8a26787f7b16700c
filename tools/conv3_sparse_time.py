"""Dense vs window-skipping tcgen05 convolution on freshly voxelized clouds (first convolution of a PVConv block):
    python tools/conv3_sparse_time.py [--empty]
Times are pessimistic by ~10 % (the L2 flush before each call is followed by a synchronize); compare within a row."""
import sys, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bdm_b200 import backend as B
from tests import cases
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
def timed(fn, reps=5):
    for _ in range(2): fn()
    ms = []
    for _ in range(reps):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return sorted(ms)[len(ms)//2]
CASES = ((32, 64, 16, 32, "shape"),) if "--empty" in sys.argv else ((32, 64, 4096, 32, "shape"), (32, 64, 4096, 32, "noise"), (32, 64, 16, 32, "shape"), (32, 32, 4096, 32, "shape"), (32, 128, 1024, 16, "shape"))
for (b, c, n, r, regime) in CASES:
    rng = np.random.default_rng(1)
    co = cases.cloud(rng, b, n, regime)
    vox, _ = cases.vox_coords(co, r)
    plan = B.voxel_plan(torch.from_numpy(vox).cuda(), r)
    feats = torch.randn(b, c, n, device="cuda")
    w = torch.randn(c, c, 3, 3, 3, device="cuda") / (27 * c) ** 0.5
    bias = torch.randn(c, device="cuda")
    prepared = B.conv3_tc05_prepare(w, None, None, 1)
    planes = B.HalfPlanes(b, c, r, "cuda")
    B.conv3_tc05_fill_planes(B.avg_voxelize_compact(feats, plan, amax_into=prepared), plan, prepared, planes, amax_ready=True)
    td = timed(lambda: B.conv3_tc05(planes, prepared, c, bias=bias, stats=True, sparse=False))
    ts = timed(lambda: B.conv3_tc05(planes, prepared, c, bias=bias, stats=True, sparse=True))
    tn = timed(lambda: B.conv3_tc05(planes, prepared, c, bias=bias, stats=False, sparse=True))
    print(f"B={b} C={c} N={n} R={r} {regime}: dense {td*1e3:.1f} us, skipping {ts*1e3:.1f} us, skipping without stats {tn*1e3:.1f} us", flush=True)
