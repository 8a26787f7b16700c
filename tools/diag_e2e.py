import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
x, feats, cams = bench.make_inputs(16, 1234, "cuda:0")
sampler = bench.build_sampler(feats, cams, "cuda:0", mode="vanilla")
x_host = x.cpu().pin_memory(); out_host = torch.empty_like(x_host).pin_memory()
with torch.no_grad():
    for _ in range(3): sampler.pc2_step(x, 500)
    sampler.enable_cuda_graphs(x)
    for _ in range(3): sampler.pc2_step(x, 500)
    torch.cuda.synchronize()
    def loop(fn, n=10):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
    print("resident, no sync      ", loop(lambda: sampler.pc2_step(x, 500)))
    def a():
        sampler.pc2_step(x, 500); torch.cuda.current_stream().synchronize()
    print("resident, sync per step", loop(a))
    def b():
        xi = x_host.to("cuda:0", non_blocking=True); y = sampler.pc2_step(xi, 500); torch.cuda.current_stream().synchronize()
    print("h2d + step + sync      ", loop(b))
    def c():
        xi = x_host.to("cuda:0", non_blocking=True); y = sampler.pc2_step(xi, 500); out_host.copy_(y, non_blocking=True); torch.cuda.current_stream().synchronize()
    print("h2d + step + d2h + sync", loop(c))
    def d():
        y = sampler.pc2_step(x, 500); out_host.copy_(y, non_blocking=True); torch.cuda.current_stream().synchronize()
    print("step + d2h + sync      ", loop(d))
    # phases of one synced step
    g = sampler._graphs[("pc2", tuple(x.shape))]
    tt = torch.full((16,), 500, device="cuda:0", dtype=torch.long)
    def e():
        g(x, tt); torch.cuda.current_stream().synchronize()
    print("graph replay + sync    ", loop(e))
