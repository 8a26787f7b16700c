"""Stand-alone timing of sparse_conv3_gather at the step's shapes (32 shapes, random taps, real voxel plans)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bdm_b200 import backend as B
from tests import cases
b = int(os.environ.get("BDM_BATCH", "32"))
rng = np.random.default_rng(1234)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for n, r, cout in ((4096, 32, 64), (4096, 32, 32), (1024, 16, 128), (256, 8, 256)):
    co = cases.cloud(rng, b, n, "shape")
    vox, _ = cases.vox_coords(co, r)
    plan = B.voxel_plan(torch.from_numpy(vox).cuda(), r)
    taps = torch.randn(b, n, 27 * cout, device="cuda")
    nocc = int((plan.cnt > 0).sum())
    fn = lambda: B.sparse_conv3_gather(taps, plan, channels_last=True, stats=True)
    for _ in range(2): fn()
    ms = []
    for _ in range(5):
        flush.max()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = sorted(ms)[2]
    nbytes = 4 * cout * (b * r ** 3 + 27 * nocc)
    print(f"gather N={n} R={r} Cout={cout}: {t * 1e3:7.1f} us  {nbytes / t / 1e6:6.0f} GB/s on {nbytes / 1e6:.0f} MB algorithmic ({nocc} occupied voxels)", flush=True)
