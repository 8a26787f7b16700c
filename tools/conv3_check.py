"""Accuracy and timing of the tcgen05 3x3x3 convolution route (csrc/conv3_tc05.cu) against torch / cuDNN.

    python tools/conv3_check.py [--quick]

For each (batch, resolution, channels): GroupNorm+Swish -> fp16 chunk planes -> convolution (+ bias, + statistics),
compared with float64 torch and with cuDNN's TF32 and fp32 results on the same input; then timings (L2 flushed).
"""
import os
import sys

import torch
import torch.nn.functional as TF

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bdm_b200 import backend as B  # noqa: E402

quick = "--quick" in sys.argv
dev = "cuda"
torch.manual_seed(0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    ms = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return sorted(ms)[len(ms) // 2]


def one(b, r, cin, cout, accuracy=True, timing=True):
    x = torch.randn(b, r, r, r, cin, device=dev) * 1.7 + 0.3           # channels-last, bias-less "first conv" output
    cb = torch.randn(cin, device=dev) * 0.2
    gamma = torch.rand(cin, device=dev) + 0.5
    beta = torch.randn(cin, device=dev) * 0.1
    w = torch.randn(cout, cin, 3, 3, 3, device=dev) / (27 * cin) ** 0.5
    bias = torch.randn(cout, device=dev) * 0.1
    groups, eps = 8, 1e-5
    xs = x.reshape(b, -1, cin).double()
    partials = torch.stack([xs.sum(1), (xs * xs).sum(1)], dim=-1).reshape(b, 1, cin, 2).contiguous()
    prepared = B.conv3_tc05_prepare(w, gamma, beta, (cin // groups) * r ** 3)
    planes = B.HalfPlanes(b, cin, r, dev)
    torch.cuda.synchronize()
    hdr = prepared[:12].view(torch.float32).tolist()

    def ours():
        B.groupnorm_swish_half_planar(x, groups, gamma, beta, eps, True, cb, partials, prepared, planes)
        return B.conv3_tc05(planes, prepared, cout, bias=bias, stats=True)

    out, stats = ours()
    torch.cuda.synchronize()
    tag = f"B={b} R={r} Cin={cin} Cout={cout}"
    if accuracy:
        # float64 reference of the same computation
        xd = (x.double() + cb.double()).permute(0, 4, 1, 2, 3)
        yd = TF.group_norm(xd, groups, gamma.double(), beta.double(), eps)
        yd = yd * torch.sigmoid(yd)
        ref = TF.conv3d(yd, w.double(), bias.double(), padding=1).permute(0, 2, 3, 4, 1)
        y32 = yd.float().contiguous(memory_format=torch.channels_last_3d)
        w32 = w.contiguous(memory_format=torch.channels_last_3d)
        torch.backends.cudnn.allow_tf32 = True
        tf32 = TF.conv3d(y32, w32, bias, padding=1).permute(0, 2, 3, 4, 1)
        torch.backends.cudnn.allow_tf32 = False
        fp32 = TF.conv3d(y32, w32, bias, padding=1).permute(0, 2, 3, 4, 1)
        torch.backends.cudnn.allow_tf32 = True
        scale = ref.abs().max().item()
        e_ours = (out.double() - ref).abs().max().item() / scale
        e_tf32 = (tf32.double() - ref).abs().max().item() / scale
        e_fp32 = (fp32.double() - ref).abs().max().item() / scale
        rms = lambda a: ((a.double() - ref) ** 2).mean().sqrt().item() / scale
        od = out.double().reshape(b, -1, cout)
        cg = cout // groups
        want1 = od.sum(1).reshape(b, groups, cg).sum(-1)
        want2 = (od * od).sum(1).reshape(b, groups, cg).sum(-1)
        got = stats.reshape(b, groups, cg, 2)
        es = max(((got[:, :, 0, 0] - want1).abs() / (want2.sqrt() + 1)).max().item(),
                 ((got[:, :, 0, 1] - want2).abs() / (want2 + 1)).max().item())
        rest = got[:, :, 1:].abs().max().item() if cg > 1 else 0.0
        ok = e_ours <= 1.5 * e_tf32 + 1e-6 and es < 1e-5 and rest == 0.0
        print(f"{tag}: max err / peak  ours {e_ours:.2e}  cudnn-tf32 {e_tf32:.2e}  cudnn-fp32 {e_fp32:.2e} | rms ours {rms(out):.2e} "
              f"tf32 {rms(tf32):.2e} | stats rel err {es:.1e} | header {hdr} | {'OK' if ok else 'WRONG'}", flush=True)
    if timing:
        y32 = torch.randn(b, cin, r, r, r, device=dev).contiguous(memory_format=torch.channels_last_3d)
        w32 = w.contiguous(memory_format=torch.channels_last_3d)
        t_cudnn = timed(lambda: TF.conv3d(y32, w32, None, padding=1))
        t_apply = timed(lambda: B.groupnorm_swish_half_planar(x, groups, gamma, beta, eps, True, cb, partials, prepared, planes))
        t_conv = timed(lambda: B.conv3_tc05(planes, prepared, cout, bias=bias, stats=True))
        t_conv_ns = timed(lambda: B.conv3_tc05(planes, prepared, cout, bias=bias, stats=False))
        xx = x.reshape(b, -1, cin)
        t_gn = timed(lambda: B.groupnorm_act_cl(xx, groups, gamma, beta, eps, True, conv_bias=cb, partials=partials))
        fl = 2.0 * b * r ** 3 * 27 * cin * cout
        print(f"{tag}: conv {t_conv * 1e3:7.1f} us ({fl / t_conv / 1e9:6.0f} TFLOP/s; without stats {t_conv_ns * 1e3:7.1f}) | cuDNN tf32 "
              f"{t_cudnn * 1e3:7.1f} us ({fl / t_cudnn / 1e9:6.0f}) | apply->fp16 planes {t_apply * 1e3:6.1f} us, apply->fp32 {t_gn * 1e3:6.1f} us",
              flush=True)


if __name__ == "__main__":
    if "--only" in sys.argv:      # --only B R C: one timing case (for ncu)
        i = sys.argv.index("--only")
        b, r, c = (int(v) for v in sys.argv[i + 1:i + 4])
        one(b, r, c, c, accuracy=False, timing=True)
        sys.exit(0)
    for (b, r, cin, cout) in ((2, 8, 32, 32), (3, 16, 64, 64), (2, 32, 32, 32), (2, 16, 128, 128), (1, 32, 64, 64), (2, 8, 256, 128),
                              (2, 16, 64, 32), (2, 16, 32, 64)):
        one(b, r, cin, cout, accuracy=True, timing=False)
    if not quick:
        for (b, r, c) in ((32, 32, 64), (32, 32, 32), (32, 16, 128), (32, 16, 64), (32, 8, 128)):
            one(b, r, c, c, accuracy=False, timing=True)
