"""GPU: the fused reverse-diffusion update (csrc/sampler.cu through the C-ABI) against the oracle and against the
eager torch op sequence, and the one-graph-per-step chains against eager sampling -- all bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n", [3 * 4096 * 2, 1031])       # a multiple of four and a ragged tail
def test_update_kernel_matches_oracle_and_eager(cuda_backend, mode, n):
    import torch
    import oracle as O
    from bdm_b200.diffusion import DDPMSchedule, PVDSchedule
    sched = DDPMSchedule() if mode == 0 else PVDSchedule()
    table = sched.table("cuda")
    g = torch.Generator().manual_seed(31 + mode)
    x, eps, noise = (torch.randn(n, generator=g) for _ in range(3))
    xd, ed, zd = x.cuda(), eps.cuda(), noise.cuda()
    rows = O.ddpm_coefficients() if mode == 0 else None
    for t in (999, 640, 1, 0):
        t_dev = torch.tensor([t], dtype=torch.int32, device="cuda")
        got = cuda_backend.sampler_update(xd, ed, zd, table, t_dev, mode)
        if mode == 0:
            want = O.ddpm_step(x.numpy(), eps.numpy(), noise.numpy(), t, rows)
        else:
            want = O.pvd_p_sample(x.numpy(), eps.numpy(), noise.numpy(), t)
        assert np.array_equal(got.cpu().numpy(), want), f"mode {mode} t {t}: differs from the oracle"
        eager = sched.step(ed, t, xd, noise=zd)             # the separate torch kernels, on the GPU
        assert torch.equal(got, eager), f"mode {mode} t {t}: differs from the eager torch sequence"
    # in place, and the device-side timestep is clamped
    t_dev = torch.tensor([5000], dtype=torch.int32, device="cuda")
    y = xd.clone()
    cuda_backend.sampler_update(y, ed, zd, table, t_dev, mode, out=y)
    assert torch.equal(y, sched.step(ed, 999, xd, noise=zd))


def _sampler(seed, graphs, b=2, n=512, c_img=5, hw=32):
    import torch
    from bdm_b200.denoiser import PointCloudModel, PVCNN2_PVD, PVCNNFuse
    from bdm_b200.diffusion import BDMSampler
    from bdm_b200.projection import ProjectionConditioner, look_at_cameras
    torch.manual_seed(seed)
    feats = torch.randn(b, c_img, hw, hw, device="cuda")
    cams = look_at_cameras([30.0, 200.0][:b], [27.0, 29.0][:b], [1.4, 1.5][:b]).to("cuda")
    cond = ProjectionConditioner(feats, cams, radius=0.05)
    pc2 = PointCloudModel(in_channels=3 + c_img).cuda().eval()
    pvd = PVCNN2_PVD(3, 64, True, 0.1, extra_feature_channels=0).cuda().eval()
    fuse = PVCNNFuse(pvd, pc2.model, extra_feature_channels=c_img).cuda().eval()
    with torch.no_grad():                       # zero-initialised projections would hide the prior branch
        for proj in fuse.projs:
            proj[-1].weight.normal_(0, 0.02)
    s = BDMSampler(pc2, cond, pvd_net=pvd, fuse_net=fuse, generator=torch.Generator(device="cuda").manual_seed(seed))
    if graphs:
        s.enable_cuda_graphs(torch.zeros(b, n, 3, device="cuda"))
    return s, b, n


MILESTONES, ROLL = (12, 9, 6, 0), 2


def test_graphed_merging_and_blending_are_bit_identical_to_eager():
    """Every step of every chain (PC^2, PVD, fusion) is one graph replay that also draws the noise, applies
    the update and decrements the device-side timestep; the clouds equal the eager run's bit for bit."""
    import torch
    from bdm_b200 import backend
    from bdm_b200.diffusion import forward_counts
    for mode in ("merging", "blending"):
        outs = []
        for graphs in (False, True):
            s, b, n = _sampler(7, graphs)
            n0 = backend.LAUNCHES
            fn = s.sample_merging if mode == "merging" else s.sample_blending
            outs.append(fn(b, n, "cuda", milestones=MILESTONES, roll_step=ROLL))
            torch.cuda.synchronize()
            assert s.forwards == forward_counts(MILESTONES, ROLL, mode)
            if graphs:
                per = s.graph_launches_per_step
                assert per["pc2"] > 50 and per["pvd"] > 50 and per["fuse"] > 100
                want = sum(per[k] * v for k, v in s.forwards.items())
                assert backend.LAUNCHES - n0 == want          # nothing of ours ran outside the replays
        assert torch.isfinite(outs[0]).all()
        assert torch.equal(outs[0], outs[1]), mode


def test_graphs_are_dropped_when_the_conditioner_or_a_network_is_replaced():
    import torch
    from bdm_b200.projection import ProjectionConditioner
    s, b, n = _sampler(9, True)
    assert len(s._graphs) == 3
    s.cond = s.cond                       # same object: nothing to invalidate
    assert len(s._graphs) == 3
    s.cond = ProjectionConditioner(torch.randn(b, 5, 32, 32, device="cuda"), s.cond.cameras, radius=0.05)
    assert not s._graphs                  # a replay would have read the old conditioner's memory
    s.enable_cuda_graphs(torch.zeros(b, n, 3, device="cuda"))
    s.pvd_net = None
    assert not s._graphs


def test_conditioner_load_serves_a_new_batch_through_the_same_graphs():
    import torch
    from bdm_b200.projection import ProjectionConditioner, look_at_cameras
    s, b, n = _sampler(11, True)
    x = torch.randn(b, n, 3, device="cuda") * 0.3
    feats2 = torch.randn(b, 5, 32, 32, device="cuda")
    cams2 = look_at_cameras([77.0, 310.0], [25.0, 30.0], [1.3, 1.6]).to("cuda")
    s.cond.load(feats2.cpu().pin_memory(), cams2)          # pinned host tensors: copy + transpose in place
    s.gen.manual_seed(3)
    got = s.pc2_step(x, 400)
    fresh, _, _ = _sampler(11, False)
    fresh.cond = ProjectionConditioner(feats2, cams2, radius=0.05)
    fresh.gen.manual_seed(3)
    assert torch.equal(got, fresh.pc2_step(x, 400))


def test_voxelization_module_matches_oracle(cuda_backend):
    """Voxelization.forward (modules/voxelization.py:16-25): integer voxel coordinates bit-exact, float
    coordinates 1e-5, voxel averages 1e-4 -- against the oracle's restatement of the module plus its
    restatement of avg_voxelize, on both input regimes and all three resolutions of the network."""
    import torch
    import oracle as O
    from bdm_b200.modules import Voxelization
    from bdm_b200.modules.point_voxel import coordinate_plan
    from . import cases, runners
    rng = np.random.default_rng(1234)
    for regime in ("noise", "shape"):
        for r, n in ((32, 4096), (16, 1024), (8, 256)):
            co = cases.cloud(rng, 4, n, regime)
            feat = rng.standard_normal((4, 7, n)).astype(np.float32)
            vox_want, nc_want = O.voxelization_coords(co, r)
            grid_want, ind_want, cnt_want = O.avg_voxelize_forward(feat, vox_want, r)
            co_t = torch.from_numpy(co).cuda()
            with torch.no_grad():
                grid, nc = Voxelization(r)(torch.from_numpy(feat).cuda(), co_t)
                _, vox, plan = coordinate_plan(co_t, r)
            label = f"voxelization[{regime},r={r}]"
            assert np.array_equal(vox.cpu().numpy(), vox_want), f"{label}: integer voxel coordinates differ"
            assert np.array_equal(plan.ind.cpu().numpy(), ind_want) and np.array_equal(plan.cnt.cpu().numpy(), cnt_want)
            runners.assert_close(f"{label}.norm_coords", nc.cpu().numpy(), nc_want, 1e-5)
            runners.assert_close(f"{label}.grid", grid.reshape(4, 7, -1).cpu().numpy(), grid_want, 1e-4)


@pytest.mark.parametrize("normalize", [True, False])
def test_fused_voxel_coordinates(cuda_backend, normalize):
    """csrc/voxel_coords.cu (one kernel for the eight torch launches of Voxelization.forward's coordinate half):
    bit-exact -- floats and integers -- against the oracle's statement of the same arithmetic, and the same integer
    voxel coordinates as the torch op sequence on the GPU (whose mean differs in summation order only)."""
    import torch
    import oracle as O
    from bdm_b200.modules.point_voxel import normalized_voxel_coords
    from . import cases, runners
    rng = np.random.default_rng(77)
    for regime in ("noise", "shape"):
        for r, n in ((32, 4096), (16, 1024), (8, 256), (8, 64), (16, 1000), (4, 7)):
            co = cases.cloud(rng, 3, n, regime) * (1.0 if normalize else 0.3)
            eps = 0.0 if n != 1000 else 1e-3
            vox_want, nc_want = O.voxelization_coords_rounded_mean(co, r, normalize, eps)
            co_t = torch.from_numpy(co).cuda()
            nc, vox = cuda_backend.voxelize_coords(co_t, r, normalize, eps)
            label = f"voxel_coords[{regime},r={r},n={n},normalize={normalize}]"
            assert np.array_equal(vox.cpu().numpy(), vox_want), f"{label}: integer coordinates differ from the oracle"
            assert np.array_equal(nc.cpu().numpy(), nc_want), f"{label}: float coordinates differ from the oracle"
            nc_torch = normalized_voxel_coords(co_t, r, normalize, eps)
            vox_torch = torch.round(nc_torch).to(torch.int32)
            assert torch.equal(vox, vox_torch), f"{label}: integer coordinates differ from the torch op sequence"
            runners.assert_close(f"{label} vs torch", nc.cpu().numpy(), nc_torch.cpu().numpy(), 2e-6)
