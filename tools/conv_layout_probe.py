"""Probe: cuDNN Conv3d time with NCDHW vs channels_last_3d activations on the denoiser's voxel shapes."""
import torch
import torch.nn as nn

torch.manual_seed(0)
shapes = [(390, 32, 32), (32, 32, 32), (64, 64, 32), (128, 128, 16), (64, 64, 16), (256, 256, 8)]
B = 16


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


for cin, cout, r in shapes:
    conv = nn.Conv3d(cin, cout, 3, padding=1).cuda().eval()
    x = torch.randn(B, cin, r, r, r, device="cuda")
    with torch.no_grad():
        t_nc = timeit(lambda: conv(x))
        xcl = x.contiguous(memory_format=torch.channels_last_3d)
        conv_cl = nn.Conv3d(cin, cout, 3, padding=1).cuda().eval().to(memory_format=torch.channels_last_3d)
        conv_cl.load_state_dict(conv.state_dict())
        conv_cl = conv_cl.to(memory_format=torch.channels_last_3d)
        t_cl = timeit(lambda: conv_cl(xcl))
        y, ycl = conv(x), conv_cl(xcl)
        err = (y - ycl).abs().max().item() / y.abs().max().item()
        print(f"Conv3d {cin:3d}->{cout:3d} R={r:2d}: NCDHW {t_nc:7.1f} us   channels_last_3d {t_cl:7.1f} us   "
              f"out CL: {ycl.is_contiguous(memory_format=torch.channels_last_3d)}   rel diff {err:.1e}", flush=True)
