#!/usr/bin/env python
"""bench.py -- BDM sampling on B200: the metric's own configuration, measured.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[4]): BDM-Merging, 32 shapes of 4096 points per GPU, the shipped schedule
(roll_step=16, milestones=[1000,968,936,872,128,64,32,0]: 995 PC^2 + 75 PVD + 5 fusion denoiser forwards per
shape, reference main_merging.py:369-523), 224x224x387 projection conditioning, random-init weights,
synthetic inputs.  ONE STEP = ONE COMPLETE SAMPLING JOB of a rank's 32 shapes: load the batch's conditioning,
1000 sampler iterations, all-gather of the finished clouds over the ranks (NCCL), Chamfer distance +
F-score of the rank's shapes against synthetic ground truth, all-reduce of the metric partials.
Metric: shapes/sec = shapes of all ranks / time, timed with CUDA events around exactly K jobs after W
warm-up jobs, barrier + synchronize on both sides, max over ranks.  Weak scaling (32 shapes per GPU).

One JSON line on stdout (rank 0):
  value        feature maps, cameras and ground truth already resident in HBM
  e2e          the same job through the public API (ProjectionConditioner.load, BDMSampler.sample_merging,
               distributed.gather_samples, evaluation.evaluate) with HOST buffers: per job the feature maps,
               cameras and ground truth are copied from pinned host memory and the finished clouds and the
               metrics are read back; the copies are inside the timed region
  roofline     the kernel group of libbdm_b200.so with the largest share of a PC^2 step (timed live with
               CUDA events); roofline_extra: the kernels north_star sets a bar on, timed stand-alone
  cpu_baseline / --impl reference: the reference's CPU route for the same workload on the host cores --
               eager PyTorch for the dense layers and the oracle port (oracle/, the one place this file
               may execute it) for the sparse ops the reference only has in CUDA -- on a bounded sample:
               one PC^2 sampler iteration of the full 32-shape batch per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS = 4096
C_IMG = 387          # 3 RGB + 384 ViT-S/16 channels (config/structured.py:79, use_mask=False)
IMG = 224
SHAPES_PER_GPU = 32  # BASELINE.json configs[4]: batch 256 over 8 GPUs
T_MID = 500
METRIC = "shapes_per_sec_1000step_sampling_4096pts"
TIME_BUDGET_S = 640.0   # timed regions of one run (the driver allows 870 s per GPU count)


def measured_peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def workload_config(mode, world, batch):
    """The `config` object of the JSON line: identical for both arms (the driver compares them)."""
    from bdm_b200.diffusion import DEFAULT_MILESTONES, DEFAULT_ROLL_STEP, forward_counts
    fw = forward_counts(mode=mode) if mode != "vanilla" else dict(pc2=1000, pvd=0, fuse=0)
    name = {"merging": "bdm_merging_1000step_b32_per_gpu (BASELINE.json configs[4])",
            "blending": "bdm_blending_1000step (BASELINE.json configs[3])",
            "vanilla": "pc2_vanilla_1000step (BASELINE.json configs[2])"}[mode]
    return {"workload": name, "shapes_per_gpu": batch, "points": N_POINTS, "image_feature_map": [C_IMG, IMG, IMG],
            "schedule": {"roll_step": DEFAULT_ROLL_STEP, "milestones": list(DEFAULT_MILESTONES)} if mode != "vanilla" else "1000 DDPM steps",
            "forwards_per_shape": fw, "step": "one complete sampling job of the rank's batch (conditioning load, 1000 sampler "
            "iterations, all-gather of the clouds, CD + F-score, all-reduce of the partials)",
            "parallelism": f"shapes sharded over {world} rank(s), no per-iteration collective",
            "l2": "a sampler iteration streams >3 GB (2.5 GB feature map + activations) through the 126 MB L2; no flush needed"}


# ---------------------------------------------------------------------------------------------------
# synthetic inputs (seeded)
# ---------------------------------------------------------------------------------------------------
def make_cameras(batch, g):
    import torch
    from bdm_b200.projection import look_at_cameras
    return look_at_cameras(torch.rand(batch, generator=g) * 360.0, 25.0 + 5.0 * torch.rand(batch, generator=g),
                           (0.65 + 0.30 * torch.rand(batch, generator=g)) * 1.75)


def make_inputs(batch, seed, device):
    """-> (x_t at t=500 (B,N,3), feature maps (B,387,224,224), cameras), all on `device`"""
    import numpy as np
    import torch
    from bdm_b200.diffusion import DDPMSchedule
    from tests.cases import cloud
    rng = np.random.default_rng(seed)
    g = torch.Generator().manual_seed(seed)
    shape = torch.from_numpy(cloud(rng, batch, N_POINTS, "shape")).permute(0, 2, 1).contiguous()  # (B,N,3)
    a = float(DDPMSchedule().alphas_cumprod[T_MID])
    x_t = (a ** 0.5) * shape + ((1 - a) ** 0.5) * torch.randn(shape.shape, generator=g)          # q(x_t | x_0)
    if device == "cpu":
        feats = torch.randn(batch, C_IMG, IMG, IMG, generator=g)
    else:
        feats = torch.randn(batch, C_IMG, IMG, IMG, generator=torch.Generator(device=device).manual_seed(seed), device=device)
    cams = make_cameras(batch, g)
    return x_t.to(device), feats, cams.to(device)


def make_ground_truth(batch, seed):
    import numpy as np
    import torch
    from tests.cases import cloud
    return torch.as_tensor(cloud(np.random.default_rng(2003 + seed), batch, N_POINTS, "shape")).permute(0, 2, 1).contiguous()


def build_sampler(feats, cams, device, mode="merging", seed=42):
    import torch
    from bdm_b200.denoiser import PVCNN2_PVD, PointCloudModel, PVCNNFuse
    from bdm_b200.diffusion import BDMSampler
    from bdm_b200.projection import ProjectionConditioner
    torch.manual_seed(seed)  # structured.py:20 run seed
    net = PointCloudModel(in_channels=3 + C_IMG, out_channels=3, embed_dim=64).to(device).eval()
    cond = ProjectionConditioner(feats, cams, radius=0.0075, scale_factor=1.0, channel_last=(device != "cpu"))
    gen = torch.Generator(device=device).manual_seed(seed)
    sampler = BDMSampler(net, cond, generator=gen)
    if mode != "vanilla":
        torch.manual_seed(seed + 1)
        sampler.pvd_net = PVCNN2_PVD(3, 64, True, 0.1, extra_feature_channels=0).to(device).eval()   # main_blending.py:133-139
    if mode == "merging":
        sampler.fuse_net = PVCNNFuse(sampler.pvd_net, sampler.pc2_net.model, extra_feature_channels=C_IMG).to(device).eval()
    return sampler


# ---------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=200):
        self.index, self.rows, self.proc, self.period_ms = index, [], None, int(period_ms)

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        return False

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: eager PyTorch + oracle port (the reference has no CPU implementation of the sparse ops)
# ---------------------------------------------------------------------------------------------------
def cpu_step_runner(batch, seed, mode="merging"):
    """Returns (step_fn, cores).  step_fn = one PC^2 sampler iteration (projection conditioning, denoiser,
    DDPM update) of `batch` shapes on the host.  The module tree is ours, but every sparse op is routed to the
    oracle (CPU restatement of the reference kernels) and every dense layer runs in eager PyTorch."""
    import torch
    import bdm_b200.functional.ops as ops
    import oracle
    from oracle.torch_backend import OracleBackend
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    ops._B = OracleBackend()
    ops.REFERENCE_CALL_PATTERN = True
    x, feats, cams = make_inputs(batch, seed, "cpu")
    sampler = build_sampler(feats, cams, "cpu", mode="vanilla")
    feats_np, R, T = feats.numpy(), cams.R.numpy(), cams.T.numpy()
    focal, pp = cams.focal.numpy(), cams.principal.numpy()

    class CpuCond:
        def get_input_with_conditioning(self, x_t):
            proj, _ = oracle.surface_projection(x_t.numpy(), R, T, focal, pp, feats_np, radius=0.0075)
            return torch.cat([x_t, torch.from_numpy(proj)], dim=2)
    sampler.cond = CpuCond()

    def step():
        with torch.no_grad():
            return sampler.pc2_step(x, T_MID)
    return step, cores


def cpu_sample_text(batch, mode):
    from bdm_b200.diffusion import forward_counts
    fw = forward_counts(mode=mode) if mode != "vanilla" else dict(pc2=1000, pvd=0, fuse=0)
    total = sum(fw.values())
    return (f"one PC^2 sampler iteration of the full {batch}-shape batch per step (1/{total} of a job: the job's "
            f"{fw['pc2']} PC^2 + {fw['pvd']} PVD + {fw['fuse']} fusion forwards are all priced as PC^2 iterations; the "
            f"gather and the evaluation, <0.1 % of a job, are left out); eager-PyTorch dense layers + oracle "
            f"port of the CUDA-only sparse ops, all host threads"), total


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.cpu_sample_shapes or args.batch
    step, cores = cpu_step_runner(batch, args.seed, args.mode)
    sample, iters_per_job = cpu_sample_text(batch, args.mode)
    for _ in range(max(1, min(args.warmup, 2))):     # the host path has no clocks / caches to settle beyond this
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = batch / (iters_per_job * dt)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "shapes/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.mode, args.gpus, args.batch),
        "cpu_baseline": {"value": value, "unit": "shapes/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "shapes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "value = shapes of the batch / (sampler iterations per job x seconds per measured iteration): the "
                "host cores do not scale with --gpus; a step here is the bounded sample described in cpu_baseline.sample",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# rooflines
# ---------------------------------------------------------------------------------------------------
def _numel(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n


def algorithmic_work(op, shp):
    """(bound, units per launch, unit) of one call of a libbdm_b200 op from its recorded argument shapes:
    SURVEY.md section 8(d) for the boundary ops, DESIGN.md section 4 for the fused routes."""
    kw = shp[-1] if shp and isinstance(shp[-1], dict) else {}
    if op == "attention":                                   # q [B,C,T]: 2 GEMMs of 2*T*T*C flops per shape
        b, c, t = shp[0]
        return "tensor", 4.0 * b * t * t * c, "flop"
    if op == "attention_qkv":                               # qkv [B,T,3C]
        b, t, c3 = shp[0]
        return "tensor", 4.0 * b * t * t * (c3 // 3), "flop"
    if op == "conv3_tc05":                                  # (HalfPlanes, prepared, c_out): 27 taps x Cin x Cout MACs per voxel
        pl, c_out = shp[0], int(shp[2])
        return "tensor", 2.0 * pl.b * pl.r ** 3 * 27 * pl.c * c_out, "flop"
    if op == "groupnorm_swish_half_planar":                 # f32 read + f16 write of the grid
        return "hbm", 6.0 * _numel(shp[0]), "byte"
    if op == "groupnorm_act":
        x = shp[0]
        n = _numel(x)
        mol = kw.get("max_over_last", False)
        return "hbm", 4.0 * n * (1.0 + (1.0 / x[-1] if mol else 1.0)), "byte"
    if op == "groupnorm_act_cl":
        return "hbm", 8.0 * _numel(shp[0]), "byte"
    if op == "avg_voxelize_compact":
        return "hbm", 8.0 * _numel(shp[0]), "byte"
    if op == "trilinear_devoxelize_cl":
        b, c, n = shp[0][0], shp[0][-1], shp[1][2]
        r3 = _numel(shp[0]) // (b * c)
        res = 4.0 * b * c * n if kw.get("residual") is not None else 0.0
        return "hbm", b * (12.0 * n + 4.0 * c * min(r3, 8 * n) + 4.0 * c * n) + res, "byte"
    if op == "three_nn_interpolate":
        b, c, m = shp[0]
        n = shp[1][2]
        return "hbm", b * (4.0 * c * m + 4.0 * c * n + 24.0 * n), "byte"
    if op in ("grouping_into", "grouping_forward"):
        b, c, n = shp[0]
        m, u = shp[1][1], shp[1][2]
        return "hbm", b * (4.0 * m * u + 4.0 * c * min(n, m * u) + 4.0 * c * m * u), "byte"
    if op == "conditioning_input":
        b, n = shp[0][0], shp[0][1]
        h, w, c = shp[5][1], shp[5][2], shp[5][3]
        return "hbm", b * (12.0 * n + 24.0 * h * w + 4.0 * n * (c + 3) + 4.0 * c * min(n, h * w)), "byte"
    return None


def pick_roofline(prof, steps, ms_step, occupied, peaks):
    """The (op, shapes) group of libbdm_b200 calls with the largest share of the profiled PC^2 iteration."""
    groups = {}
    for op, calls in prof.items():
        for ms, shp in calls:
            if op == "conv3_tc05":
                # one kernel per (operand, c_out): first (window-skipping) and second (dense) convolutions of a block rank
                # together; the fraction below is taken from the dense launches only
                key = (op, shp[0].describe(), int(shp[2]))
                kw = shp[-1] if isinstance(shp[-1], dict) else {}
                g = groups.setdefault(key, {"op": op, "shapes": shp, "ms": [], "ms_dense": []})
                if not kw.get("sparse"):
                    g["ms_dense"].append(ms)
                    g["shapes"] = shp
            else:
                key = (op, json.dumps(shp, default=lambda o: getattr(o, "describe", lambda: type(o).__name__)()))
                g = groups.setdefault(key, {"op": op, "shapes": shp, "ms": []})
            g["ms"].append(ms)
    def robust_total(g):
        # per-op events on eager launches also see the host: when the GPU runs ahead of Python the gap between an op's
        # two events includes enqueue latency.  Median x count ranks the groups by device time.
        ms = sorted(g["ms"])
        return ms[len(ms) // 2] * len(ms)
    ranked = sorted(groups.values(), key=lambda g: -robust_total(g))
    for g in ranked:
        op, shp = g["op"], g["shapes"]
        if op == "sparse_conv3_gather":
            b, n, k = shp[0]
            cout = k // 27
            plan = shp[1]
            work = ("hbm", 4.0 * cout * (b * plan.r ** 3 + 27.0 * occupied.get(plan.r, 0)), "byte")
        else:
            work = algorithmic_work(op, shp)
        if work is None:
            continue
        bound, units, unit = work
        timed = sorted(g.get("ms_dense") or g["ms"])
        ms = timed[len(timed) // 2]
        if bound == "hbm":
            ach, peak, u = units / (ms * 1e-3) / 1e9, peaks["hbm_gbs"], "GB/s"
        else:
            ach, peak, u = units / (ms * 1e-3) / 1e12, peaks["bf16_tflops_sustained"], "TFLOP/s"
        shown = [list(s) if isinstance(s, tuple) else (s if isinstance(s, (int, float, str, bool, dict, type(None)))
                                                       else getattr(s, "describe", lambda: type(s).__name__)())
                 for s in shp]
        extra = {}
        if op == "conv3_tc05":
            extra = {"launches_in_fraction": len(timed),
                     "note": "share_of_iteration counts every launch of this kernel instance (first convolutions skip all-zero tap "
                             "windows and are faster; achieved / frac come from the dense second convolutions only). "
                             "algorithmic flops = 2*27*Cin*Cout per real voxel (the padded rows the flat layout also computes, "
                             "6 % at R=32 / 13 % at R=16, are not counted); fp16 operands (11 significant bits, as TF32), "
                             "fp32 accumulation; SS-mode MMA: the kernel is bound by shared-memory operand bandwidth (DESIGN.md)"}
        if op in ("attention", "attention_qkv"):
            # every fp32 product is three fp16 tensor-core products (lo*hi + hi*lo + hi*hi): the tensor pipe executes 3x
            extra = {"executed_tflops": 3.0 * ach, "frac_executed": 3.0 * ach / peak,
                     "note": "algorithmic flops = the fp32 attention (4*B*T*T*C); executed = 3x (fp16 hi/lo split). The launch "
                             "group is amax + operand prep + the tcgen05 kernel."}
        return {"bound": bound, "kernel": f"{op} {json.dumps(shown)}", "achieved": ach, "peak": peak, "unit": u,
                "frac": ach / peak, "traffic": None, "peak_source": peaks["source"], **extra,
                "algorithmic_%ss_per_launch" % unit: units, "ms_per_launch": ms,
                "launches_timed": len(g["ms"]), "launches_per_iteration": len(g["ms"]) // max(steps, 1),
                "share_of_iteration": robust_total(g) / max(steps, 1) / ms_step,
                "timed_in": "eager single-stream pass of one PC^2 sampler iteration of the job's batch (per-op CUDA events, median per "
                            "launch); share_of_iteration = launches x median / the graph-replayed iteration (iteration_ms.pc2)"}
    return None


def standalone_rooflines(device, batch, peaks):
    """roofline_extra: the kernels north_star sets a bar on (voxelize, devoxelize, ball query), the projection
    (a10) and the evaluation NN (a11), each timed stand-alone with CUDA events, L2 flushed between calls."""
    import numpy as np
    import torch
    from bdm_b200 import backend as B
    from bdm_b200.projection import look_at_cameras
    from tests import cases
    rng = np.random.default_rng(1234)
    b, n = batch, N_POINTS
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=device)
    out = []

    def timeit(fn, reps=8):
        for _ in range(2):
            fn()
        ms = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return statistics.median(ms)

    def hbm(name, nbytes, ms, note=None):
        ach = nbytes / (ms * 1e-3) / 1e9
        d = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
             "frac": ach / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": nbytes, "ms_per_launch": ms}
        if note:
            d["note"] = note
        out.append(d)

    co = torch.from_numpy(cases.cloud(rng, b, n, "shape")).to(device)
    for c, r in ((64, 32), (390, 32)):
        vox, nc = cases.vox_coords(co.cpu().numpy(), r)
        vox_t, nc_t = torch.from_numpy(vox).to(device), torch.from_numpy(nc).to(device)
        feat = torch.randn(b, c, n, device=device)
        ms = timeit(lambda: B.avg_voxelize_forward(feat, vox_t, r))
        hbm(f"avg_voxelize C={c} N={n} R={r} (plan + fill)", b * (4 * c * n + 12 * n + 4 * c * r ** 3 + 4 * n + 4 * r ** 3), ms)
        if c == 64:
            grid = torch.randn(b, c, r ** 3, device=device)
            ms = timeit(lambda: B.trilinear_devoxelize_forward(r, False, nc_t, grid))
            hbm(f"trilinear_devoxelize C={c} N={n} R={r} (binning + gather)", b * (12 * n + 4 * c * min(r ** 3, 8 * n) + 4 * c * n), ms)
            grid_cl = grid.view(b, c, r, r, r).permute(0, 2, 3, 4, 1).contiguous()
            ms = timeit(lambda: B.trilinear_devoxelize_cl(grid_cl, nc_t, r))
            hbm(f"trilinear_devoxelize_cl C={c} N={n} R={r} (channels-last grid, the route inside the step)",
                b * (12 * n + 4 * c * min(r ** 3, 8 * n) + 4 * c * n), ms)
    idx = B.furthest_point_sampling(co, 1024)
    cen = B.gather_features_forward(co, idx)
    ms = timeit(lambda: B.ball_query(cen, co, 0.1, 32))
    tests_ps = b * 1024.0 * n / (ms * 1e-3)
    out.append({"kernel": f"ball_query M=1024 N={n} r=0.1 U=32", "bound": "fp32 issue (B*M*N distance tests; 4 MB of algorithmic bytes)",
                "achieved": tests_ps / 1e12, "unit": "T pair-tests/s", "peak": None, "frac": None, "ms_per_launch": ms,
                "algorithmic_bytes_per_launch": b * (12 * (n + 1024) + 4 * 1024 * 32),
                "hbm_frac_on_algorithmic_bytes": b * (12 * (n + 1024) + 4 * 1024 * 32) / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]})
    ms = timeit(lambda: B.furthest_point_sampling(co, 1024))
    out.append({"kernel": f"furthest_point_sampling N={n} M=1024", "bound": "latency (M-1 dependent rounds, one CTA per shape)",
                "achieved": 1023.0 / (ms * 1e3), "unit": "rounds/us", "peak": None, "frac": None, "ms_per_launch": ms})
    # a10: projection conditioning straight into the denoiser's channel-first input
    g = torch.Generator().manual_seed(5)
    cams = make_cameras(b, g).to(device)
    feat_hwc = torch.randn(b, IMG, IMG, C_IMG, device=device)
    pts = co.permute(0, 2, 1).contiguous() * 0.3
    pix = B.conditioning_input(pts, cams.R, cams.T, cams.focal, cams.principal, feat_hwc, 0.0075)[1]
    winners = int((pix >= 0).sum().item())
    ms = timeit(lambda: B.conditioning_input(pts, cams.R, cams.T, cams.focal, cams.principal, feat_hwc, 0.0075))
    hbm(f"surface_projection -> channel-first input B={b} N={n} C={C_IMG} {IMG}x{IMG} (5 kernels)",
        b * (12 * n + 24 * IMG * IMG + 4 * n * (C_IMG + 3)) + 4 * C_IMG * winners, ms,
        note=f"{winners} of {b * n} points win a pixel")
    # a11: evaluation nearest neighbour, fp64
    gt = torch.as_tensor(np.ascontiguousarray(cases.cloud(rng, 8, n, "shape").transpose(0, 2, 1).astype(np.float64))).to(device)
    pred = (gt[:, torch.randperm(n, device=device)] + 0.05 * torch.randn(gt.shape, device=device, dtype=torch.float64)).contiguous()
    ms = timeit(lambda: B.nn_f64(pred, gt, expanded=False, return_index=False))
    out.append({"kernel": f"nn_f64 8 pairs x {n} x {n} (direct form)", "bound": "fp64 issue", "achieved": 8.0 * n * n / (ms * 1e-3) / 1e12,
                "unit": "T pair-tests/s (8 fp64 flops each)", "peak": None, "frac": None, "ms_per_launch": ms})
    return out


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from bdm_b200 import backend
    from bdm_b200 import distributed as D
    from bdm_b200 import evaluation as E
    from bdm_b200.diffusion import forward_counts
    from bdm_b200.projection import Cameras

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: bdm_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    B, mode = args.batch, args.mode
    seed = D.rank_seed(args.seed, rank)                     # per-rank seed (training_utils.py:373-379)
    t_start = time.perf_counter()

    # ---- the job's inputs: resident copies and pinned host copies ----
    x_mid, feats_dev, cams_dev = make_inputs(B, seed, device)
    gt_dev = make_ground_truth(B, seed).to(device)
    feats_host = feats_dev.cpu().pin_memory()
    cams_host = Cameras(*(t.cpu().pin_memory() for t in (cams_dev.R, cams_dev.T, cams_dev.focal, cams_dev.principal)))
    gt_host = gt_dev.cpu().pin_memory()
    clouds_host = torch.empty((B, N_POINTS, 3), dtype=torch.float32).pin_memory()
    gt_stage = torch.empty_like(gt_dev)

    sampler = build_sampler(feats_dev, cams_dev, device, mode)
    if not args.no_graph:
        sampler.enable_cuda_graphs(x_mid)
    total = B * world
    sample_fn = {"merging": sampler.sample_merging, "blending": sampler.sample_blending,
                 "vanilla": sampler.sample_vanilla}[mode]
    last = {}

    def finish(x, gt):
        clouds = D.gather_samples(x.contiguous(), total)          # NCCL all-gather of the finished clouds
        cd, f1 = E.evaluate(x, gt)                                # CD x1e3 and F-score@0.01 of the rank's shapes
        last["metrics"] = D.reduce_metrics(cd, f1)                # all-reduce of (sum CD, sum F, count); host read
        last["clouds"] = clouds
        return clouds

    def job_resident():
        sampler.cond.load(feats_dev, cams_dev)
        return finish(sample_fn(B, N_POINTS, device), gt_dev)

    def job_e2e():
        # host buffers in, host buffers out: the batch's feature maps / cameras / ground truth come from pinned
        # host memory, the finished clouds and the metrics go back to the host
        sampler.cond.load(feats_host, cams_host)
        gt_stage.copy_(gt_host, non_blocking=True)
        x = sample_fn(B, N_POINTS, device)
        clouds = finish(x, gt_stage)
        clouds_host.copy_(x, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return clouds

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    # ---- warm-up jobs, then the timed regions ----
    warm = max(args.warmup, 3)
    t0 = time.perf_counter()
    job_resident()
    barrier()
    job_s = time.perf_counter() - t0
    for _ in range(warm - 1):
        job_resident()
    job_e2e()
    barrier()
    # K jobs for `value`; the e2e region gets the same K unless that would overrun the time budget
    k_e2e = max(1, min(args.steps, int((TIME_BUDGET_S - (warm + 1 + args.steps) * job_s) / max(job_s, 1e-3))))
    if world > 1:
        kk = torch.tensor([k_e2e], device=device)
        dist.all_reduce(kk, op=dist.ReduceOp.MIN)
        k_e2e = int(kk.item())
    launches0 = backend.LAUNCHES
    with ClockSampler(local, args.clock_ms) as clocks:
        ms_job = timed(job_resident, args.steps)
        launches = backend.LAUNCHES - launches0
        ms_e2e = timed(job_e2e, k_e2e)
    fw_done = dict(sampler.forwards)
    jobs_done = warm + 1 + args.steps + k_e2e
    mean_cd, mean_f1, count = last["metrics"]
    finite = bool(torch.isfinite(last["clouds"]).all())
    gathered = list(last["clouds"].shape)

    value = total / (ms_job * 1e-3)
    e2e_value = total / (ms_e2e * 1e-3)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- extras (rank 0; the other ranks wait at the barrier above) ----
    peaks = measured_peaks()
    expected_fw = forward_counts(mode=mode) if mode != "vanilla" else dict(pc2=1000, pvd=0, fuse=0)
    line = {
        "metric": METRIC, "value": value, "unit": "shapes/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_job, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "precision": "fp32 activations and accumulation throughout; convolution / 1x1-convolution operands carry 11 significant "
                     "bits as in the reference under torch's default TF32 conv policy (TF32 in cuDNN / cuBLAS; fp16 after a "
                     "power-of-two scaling in the tcgen05 3x3x3 convolutions); attention is fp32-equivalent (fp16 hi/lo split, "
                     "3 products); sparse ops are fp32, integer outputs bit-exact",
        "data": "synthetic", "config": workload_config(mode, world, B),
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": "shapes/s", "ms_per_step": ms_e2e, "steps": k_e2e,
                "h2d_bytes_per_step": (feats_host.numel() + gt_host.numel() + 9 * B + 3 * B + 4 * B) * 4 * world,
                "d2h_bytes_per_step": (clouds_host.numel() * 4 + 24) * world,
                "api": "ProjectionConditioner.load(pinned host) -> BDMSampler.sample_%s -> distributed.gather_samples -> "
                       "evaluation.evaluate -> distributed.reduce_metrics -> clouds to pinned host" % mode},
        "gpu_launches": launches,
        "job": {"seconds": ms_job * 1e-3, "shapes": total, "gathered": gathered, "finite": finite,
                "forwards_per_shape_per_job": {k: v // jobs_done for k, v in fw_done.items()},
                "expected_forwards": expected_fw, "mean_cd_x1e3_vs_synthetic_gt": mean_cd,
                "mean_fscore_vs_synthetic_gt": mean_f1, "evaluated": count,
                "launch": "every sampler iteration is one CUDA-graph replay (conditioning + denoiser + noise + update + "
                          "timestep decrement)" if not args.no_graph else "eager",
                "libbdm_b200_kernels_per_replay": sampler.graph_launches_per_step,
                "note": "random-init weights: the metric values only exercise the evaluation path"},
        "implementation": {
            "conv3_tc05": "3x3x3 convolutions on 16^3 / 32^3 grids: libbdm_b200's tcgen05 implicit GEMM (fp16 chunk planes, TMEM "
                          "accumulators, bias + GroupNorm statistics in the epilogue); a block's first convolution reads planes "
                          "scattered from the compact voxel averages and skips all-zero tap windows (c_in <= 128)",
            "sparse_first_conv": "first convolution of the 8^3 blocks and of the 390-channel input block: compact averages -> GEMM "
                                 "(cuBLAS, TF32 like the Conv3d) -> sparse_conv3_gather; BDM_SPARSE_CONV=0 restores the dense route",
            "dense_layers": "second Conv3d of the 8^3 blocks (cuDNN) and 1x1 convs (cuBLAS): torch, PyTorch default TF32 conv policy; "
                            "conv bias + GroupNorm + Swish (+ SE squeeze, + max over neighbours, + devoxelize), attention (tcgen05): "
                            "libbdm_b200",
        },
    }

    # ---- one sampler iteration of each kind: graph replay times, per-op events, roofline ----
    if not args.no_breakdown:
        with torch.no_grad():
            def replay_ms(kind, x, t, reps):
                g = sampler._graphs.get((kind, tuple(x.shape)))
                if g is None:
                    return None
                extra = {"prior": x} if kind == "fuse" else {}
                g.run(x, t, 2, **extra)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g.x.copy_(x)
                g.t.fill_(t)
                e0.record()
                for _ in range(reps):
                    g.graph.replay()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / reps
            it = {"pc2": replay_ms("pc2", x_mid, 700, 40), "pvd": replay_ms("pvd", x_mid.permute(0, 2, 1).contiguous(), 700, 20),
                  "fuse": replay_ms("fuse", x_mid, 700, 5)}
            line["iteration_ms"] = it
            if all(v is not None for v in it.values()):
                line["job"]["seconds_from_iterations"] = sum(expected_fw[k] * it[k] for k in it) * 1e-3

            import bdm_b200.denoiser as denoiser_mod
            from bdm_b200.modules.point_voxel import coordinate_plan
            occupied = {r: int((coordinate_plan(x_mid.transpose(1, 2).contiguous(), r)[2].cnt > 0).sum().item())
                        for r in (32,)}
            plan_ahead_default = denoiser_mod.PLAN_AHEAD
            denoiser_mod.PLAN_AHEAD = False              # events on one stream: clean per-op durations
            tt = torch.full((B,), T_MID, device=device, dtype=torch.long)
            psteps = 3
            try:
                sampler._pc2_eps(x_mid, tt)
                backend.profile_start()
                for _ in range(psteps):
                    sampler._pc2_eps(x_mid, tt)
                prof = backend.profile_stop()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(psteps):
                    sampler._pc2_eps(x_mid, tt)
                e1.record()
                torch.cuda.synchronize()
                ms_eager = e0.elapsed_time(e1) / psteps
            finally:
                denoiser_mod.PLAN_AHEAD = plan_ahead_default
            by_op = {k: sum(ms for ms, _ in v) / psteps for k, v in prof.items()}
            line["sparse_path"] = {"ms_per_iteration_by_op": by_op, "ms_total": sum(by_op.values()),
                                   "ms_eager_iteration": ms_eager,
                                   "note": "per-op CUDA events on eager single-stream launches of one PC^2 iteration (host gaps "
                                           "inside an op included); the graph replay of the same iteration is iteration_ms.pc2"}
            roof = pick_roofline(prof, psteps, it.get("pc2") or ms_eager, occupied, peaks)   # share of the graph-replayed iteration
            if roof is not None:
                try:
                    traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
                    key = roof["kernel"].split(" ")[0]
                    roof["traffic"] = traffic.get(key, {}).get("dram_bytes_per_launch")
                    roof["traffic_source"] = traffic.get(key, {}).get("source")
                except Exception:
                    pass
            line["roofline"] = roof
            try:
                line["roofline_extra"] = standalone_rooflines(device, B, peaks)
            except Exception as e:
                line["roofline_extra"] = {"unavailable": repr(e)[:300]}

    # ---- reference CUDA kernels (recompiled for sm_100a) under the reference's call pattern ----
    if world == 1 and not args.no_ref_cuda:
        try:
            from oracle import build_ref
            ref = build_ref.load_ref()
            if ref is not None:
                import bdm_b200.functional.ops as ops
                saved = (ops._B, ops.REFERENCE_CALL_PATTERN)
                ops._B, ops.REFERENCE_CALL_PATTERN = ref, True
                tt = torch.full((B,), T_MID, device=device, dtype=torch.long)
                try:
                    with torch.no_grad():
                        for _ in range(2):
                            sampler._pc2_eps(x_mid, tt)
                        torch.cuda.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        for _ in range(5):
                            sampler.ddpm.step(sampler._pc2_eps(x_mid, tt), T_MID, x_mid, sampler.gen)
                        e1.record()
                        torch.cuda.synchronize()
                        ms_ref = e0.elapsed_time(e1) / 5
                finally:
                    ops._B, ops.REFERENCE_CALL_PATTERN = saved
                line["reference_cuda"] = {"what": "one PC^2 sampler iteration of the same batch with the reference's own kernels "
                                                  "(oracle/_ref, unmodified sources recompiled for sm_100a) and its native-call "
                                                  "pattern, eager (they launch on the legacy default stream)",
                                          "ms_per_iteration": ms_ref,
                                          "ours_ms_per_iteration": line.get("iteration_ms", {}).get("pc2"),
                                          "ours_eager_ms_per_iteration": line.get("sparse_path", {}).get("ms_eager_iteration")}
        except Exception as e:  # the reference extension is optional evidence, never required
            line["reference_cuda"] = {"unavailable": repr(e)[:200]}

    # ---- BASELINE configs[0]: CD + F-score@0.01 on 8 synthetic 4096-point pairs (evaluation kNN) ----
    if world == 1 and not args.no_breakdown:
        try:
            import numpy as np
            from tests.cases import cloud
            rng = np.random.default_rng(2003)
            gt_np = cloud(rng, 8, N_POINTS, "shape").transpose(0, 2, 1).astype(np.float64)
            pred_np = gt_np[:, rng.permutation(N_POINTS)] + 0.05 * rng.standard_normal(gt_np.shape)
            gt_t, pred_t = torch.as_tensor(gt_np).to(device), torch.as_tensor(pred_np).to(device)
            for _ in range(3):
                cd, f1 = E.evaluate(pred_t, gt_t)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                cd, f1 = E.evaluate(pred_t, gt_t)
            e1.record()
            torch.cuda.synchronize()
            gpu_ms = e0.elapsed_time(e1) / 10
            import oracle
            t0 = time.perf_counter()
            pc, gc = pred_np - pred_np.mean(1, keepdims=True), gt_np - gt_np.mean(1, keepdims=True)
            cd_cpu, f1_cpu = oracle.chamfer_distance(pc, gc) * 1000.0, oracle.fscore(gc, pc)
            cpu_ms = (time.perf_counter() - t0) * 1e3
            line["eval_knn"] = {"what": "BASELINE configs[0]: Chamfer x1e3 + F-score@0.01, 8 pairs x 4096 points, fp64",
                                "gpu_ms": gpu_ms, "pairs_per_s": 8 / (gpu_ms * 1e-3), "cpu_port_ms": cpu_ms,
                                "mean_cd_x1e3": float(cd.mean()), "mean_fscore": float(f1.mean()),
                                "max_abs_diff_vs_cpu_port": [float(np.abs(cd.cpu().numpy() - cd_cpu).max()),
                                                             float(np.abs(f1.cpu().numpy() - f1_cpu).max())]}
        except Exception as e:
            line["eval_knn"] = {"unavailable": repr(e)[:200]}

    # ---- CPU baseline: bounded sample on the host cores ----
    if world == 1 and not args.no_cpu_baseline:
        import bdm_b200.functional.ops as ops
        saved = (ops._B, ops.REFERENCE_CALL_PATTERN)
        try:
            cb = args.cpu_sample_shapes or B
            step, cores = cpu_step_runner(cb, args.seed, mode)
            sample, iters_per_job = cpu_sample_text(cb, mode)
            step()
            reps, t0 = 0, time.perf_counter()
            while reps < 2 or (time.perf_counter() - t0 < 15.0 and reps < 20):
                step()
                reps += 1
            dt = (time.perf_counter() - t0) / reps
            line["cpu_baseline"] = {"value": cb / (iters_per_job * dt), "unit": "shapes/s", "cores": cores, "kind": "port",
                                    "ms_per_iteration": dt * 1e3, "sample": f"{reps} x " + sample}
        finally:
            ops._B, ops.REFERENCE_CALL_PATTERN = saved

    line["wall_s"] = time.perf_counter() - t_start
    print(json.dumps(line, default=lambda o: type(o).__name__), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3, help="timed sampling jobs (a step is one complete job)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="merging", choices=["merging", "blending", "vanilla"],
                    help="sampling procedure (default: BDM-Merging, the configuration the metric is quoted on)")
    ap.add_argument("--batch", type=int, default=SHAPES_PER_GPU, help="shapes per GPU (configs[4]: 32)")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--cpu-sample-shapes", type=int, default=0, help="shapes per CPU iteration (0 = the full per-GPU batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true", help="skip the per-iteration breakdown / rooflines")
    ap.add_argument("--clock-ms", type=int, default=200, help="nvidia-smi sampling period during the timed regions")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of one CUDA graph per iteration")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
