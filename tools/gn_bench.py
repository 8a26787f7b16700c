"""Stand-alone timing of the GroupNorm(+Swish) routes at the step's large shapes (CUDA-graph replays of 10 calls,
L2 flushed before each replay).   BDM_GN_CLUSTER=0|1 python tools/gn_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bdm_b200 import backend as B  # noqa: E402

b = int(os.environ.get("BDM_BATCH", "32"))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def bench(name, fn, nbytes, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.max() if os.environ.get('BDM_FLUSH', 'read') == 'read' else flush.zero_()   # read-flush leaves no dirty lines behind
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = sorted(ms)[len(ms) // 2]
    print(f"{name:60s} {t * 1e3:8.1f} us  {nbytes / t / 1e6:7.0f} GB/s (1R+1W)", flush=True)


for c, spatial, u in ((64, (1024, 32), 32), (32, (1024, 32), 0), (128, (256, 32), 32), (64, (256, 32), 0), (128, (4096,), 0),
                      (512, (16, 32), 32), (256, (64, 32), 32)):
    x = torch.randn((b, c) + spatial, device="cuda")
    w, bi, cb = (torch.randn(c, device="cuda") for _ in range(3))
    n = x.numel() * 4
    bench(f"channel-first {tuple(x.shape)} max_over_last={bool(u)}", lambda: B.groupnorm_act(x, 8, w, bi, 1e-5, True, conv_bias=cb, max_over_last=bool(u)),
          n * (1 + (1 / spatial[-1] if u else 1)))
for c, r in ((64, 32), (32, 32), (128, 16)):
    x = torch.randn(b, r, r, r, c, device="cuda")
    w, bi, cb = (torch.randn(c, device="cuda") for _ in range(3))
    n = x.numel() * 4
    bench(f"channels-last {tuple(x.shape)} stats+apply, sums", lambda: B.groupnorm_act_cl(x, 8, w, bi, 1e-5, True, conv_bias=cb, channel_sums="tiles"), 2 * n)
    part = torch.zeros(b, 128, c, 2, dtype=torch.float64, device="cuda")
    bench(f"channels-last {tuple(x.shape)} apply only (producer stats)", lambda: B.groupnorm_act_cl(x, 8, w, bi, 1e-5, True, conv_bias=cb, partials=part), 2 * n)

x = torch.randn(b, 32, 32, 32, 64, device="cuda")
y = torch.empty_like(x)
bench("torch copy_ of the same tensor (streaming ceiling, same method)", lambda: y.copy_(x), 2 * x.numel() * 4)
