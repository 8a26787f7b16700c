import torch, sys
sys.path.insert(0, '.')
from bdm_b200 import backend as B
for (b, t, scale) in [(2, 512, 1.0), (3, 4096, 1.0), (1, 1024, 3.0), (2, 128, 0.2), (16, 4096, 0.6)]:
    g = torch.Generator(device="cuda").manual_seed(t + b)
    q, k, v = (torch.randn(b, 64, t, device="cuda", generator=g) * scale for _ in range(3))
    got = B.attention(q, k, v)
    ref = torch.matmul(v.double(), torch.softmax(torch.matmul(q.double().transpose(1, 2), k.double()), -1).transpose(1, 2))
    f32 = torch.matmul(v, torch.softmax(torch.matmul(q.transpose(1, 2), k), -1).transpose(1, 2))
    peak = ref.abs().max().item()
    print(b, t, scale, "ours", (got.double() - ref).abs().max().item() / peak, "torch fp32", (f32.double() - ref).abs().max().item() / peak, flush=True)
    def timeit(fn, iters=5):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3
    print("   us: ours", timeit(lambda: B.attention(q, k, v)), "torch", timeit(lambda: torch.matmul(v, torch.softmax(torch.matmul(q.transpose(1, 2), k), -1).transpose(1, 2))), flush=True)
