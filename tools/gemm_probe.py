import torch, time
torch.backends.cuda.matmul.allow_tf32 = True
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (cout, cin, L) in ((32, 390, 4096), (64, 35, 32768), (128, 67, 8192), (256, 131, 2048), (128, 387 + 64 + 0, 4096)):
    b = 32
    x = torch.randn(b, cin, L, device="cuda")
    w = torch.randn(cout, cin, device="cuda")
    kp = (cin + 3) // 4 * 4
    wp = torch.zeros(cout, kp, device="cuda"); wp[:, :cin] = w
    wv = wp[:, :cin]
    head = cin & ~3
    wh, wt = w[:, :head].contiguous(), w[:, head:].contiguous()
    def split():
        y = torch.matmul(wh, x[:, :head]); y.baddbmm_(wt.expand(b, -1, -1), x[:, head:]); return y
    y0 = torch.matmul(w, x); y1 = torch.matmul(wv, x); y2 = split()
    print(cout, cin, L, "plain", round(t(lambda: torch.matmul(w, x)), 1), "ld-padded view", round(t(lambda: torch.matmul(wv, x)), 1),
          "head+tail", round(t(split), 1), "maxdiff", float((y1 - y0).abs().max()), float((y2 - y0).abs().max()))
