"""Timing + float64 check of the fused attention at the step's shapes (B x 64 x 4096).
    BDM_ATTENTION=tc05x2|tc05x1|mma python tools/attn_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bdm_b200 import backend as B  # noqa: E402

for b in (2, 16, 32):
    g = torch.Generator(device="cuda").manual_seed(b)
    q, k, v = (torch.randn(b, 64, 4096, device="cuda", generator=g) * 0.6 for _ in range(3))
    got = B.attention(q, k, v)
    nb = min(b, 2)
    ref = torch.matmul(v[:nb].double(), torch.softmax(torch.matmul(q[:nb].double().transpose(1, 2), k[:nb].double()), -1).transpose(1, 2))
    err = (got[:nb].double() - ref).abs().max().item() / ref.abs().max().item()
    for _ in range(3):
        B.attention(q, k, v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        B.attention(q, k, v)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    print(f"B={b} T=4096: {us:.1f} us  ({4 * b * 4096 * 4096 * 64 / us / 1e6:.0f} TFLOP/s fp32-equivalent)  err vs float64 {err:.2e}", flush=True)
