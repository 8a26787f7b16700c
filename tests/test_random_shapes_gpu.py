"""GPU: randomised (seeded) odd sizes for every op against the oracle -- ragged tails, sizes that are not
multiples of 4 / 32, single points / centres / channels, more centres than points, batch 1."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rng(seed):
    return np.random.default_rng(seed)


def _cloud(rng, b, n, scale=1.0):
    return (rng.standard_normal((b, 3, n)) * scale).astype(np.float32)


@pytest.mark.parametrize("seed", range(12))
def test_voxelize_devoxelize_random(seed, cuda_backend):
    import torch

    import oracle as O
    rng = _rng(seed)
    b, c, n = int(rng.integers(1, 5)), int(rng.integers(1, 70)), int(rng.integers(1, 3000))
    r = int(rng.choice([1, 2, 3, 5, 8, 11, 16, 20, 27, 32, 33, 40]))
    nc = (rng.random((b, 3, n)) * (r - 1)).astype(np.float32)
    vox = np.round(nc).astype(np.int32)
    feat = rng.standard_normal((b, c, n)).astype(np.float32)
    out, ind, cnt = cuda_backend.avg_voxelize_forward(torch.as_tensor(feat).cuda(), torch.as_tensor(vox).cuda(), r)
    oo, oi, oc = O.avg_voxelize_forward(feat, vox, r)
    assert np.array_equal(ind.cpu().numpy(), oi) and np.array_equal(cnt.cpu().numpy(), oc)
    assert np.abs(out.cpu().numpy() - oo).max() <= 1e-4 * max(np.abs(oo).max(), 1e-30)
    for training in (False, True):
        got = cuda_backend.trilinear_devoxelize_forward(r, training, torch.as_tensor(nc).cuda(), torch.as_tensor(oo).cuda())
        want = O.trilinear_devoxelize_forward(r, training, nc, oo)
        assert np.array_equal(got[0].cpu().numpy(), want[0])
        if training:
            assert np.array_equal(got[1].cpu().numpy(), want[1]) and np.array_equal(got[2].cpu().numpy(), want[2])


@pytest.mark.parametrize("seed", range(12))
def test_fps_ball_group_random(seed, cuda_backend):
    import torch

    import oracle as O
    rng = _rng(100 + seed)
    b, n = int(rng.integers(1, 4)), int(rng.integers(1, 2500))
    m, u = int(rng.integers(1, 400)), int(rng.choice([1, 2, 7, 16, 32, 33, 64]))
    c = int(rng.integers(1, 40))
    co = _cloud(rng, b, n)
    radius = float(rng.choice([0.05, 0.2, 0.7, 3.0]))
    idx = cuda_backend.furthest_point_sampling(torch.as_tensor(co).cuda(), m)
    oidx = O.furthest_point_sampling(co, m)
    assert np.array_equal(idx.cpu().numpy(), oidx)
    cen = O.gather_features_forward(co, oidx)
    assert np.array_equal(cuda_backend.gather_features_forward(torch.as_tensor(co).cuda(), idx).cpu().numpy(), cen)
    nb = cuda_backend.ball_query(torch.as_tensor(cen).cuda(), torch.as_tensor(co).cuda(), radius, u)
    onb = O.ball_query(cen, co, radius, u)
    assert np.array_equal(nb.cpu().numpy(), onb)
    feat = rng.standard_normal((b, c, n)).astype(np.float32)
    g = cuda_backend.grouping_forward(torch.as_tensor(feat).cuda(), nb)
    assert np.array_equal(g.cpu().numpy(), O.grouping_forward(feat, onb))


@pytest.mark.parametrize("seed", range(12))
def test_three_nn_random(seed, cuda_backend):
    import torch

    import oracle as O
    rng = _rng(200 + seed)
    b, n, m, c = int(rng.integers(1, 4)), int(rng.integers(1, 3000)), int(rng.integers(1, 1200)), int(rng.integers(1, 50))
    pts, cen = _cloud(rng, b, n), _cloud(rng, b, m)
    if m > 4:
        cen[:, :, m // 2] = cen[:, :, 0]   # duplicate centre: equal distances
    feat = rng.standard_normal((b, c, m)).astype(np.float32)
    out, idx, w = cuda_backend.three_nearest_neighbors_interpolate_forward(
        torch.as_tensor(pts).cuda(), torch.as_tensor(cen).cuda(), torch.as_tensor(feat).cuda())
    oo, oi, ow = O.three_nearest_neighbors_interpolate_forward(pts, cen, feat)
    assert np.array_equal(idx.cpu().numpy(), oi)
    assert np.array_equal(w.cpu().numpy(), ow) and np.array_equal(out.cpu().numpy(), oo)


def test_empty_inputs(cuda_backend):
    """zero-sized dimensions: no launch, right shapes, no crash"""
    import torch
    z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device="cuda")  # noqa: E731
    assert cuda_backend.furthest_point_sampling(z(2, 3, 5), 0).shape == (2, 0)
    assert cuda_backend.gather_features_forward(z(1, 0, 5), z(1, 3, dt=torch.int32)).shape == (1, 0, 3)
    out, ind, cnt = cuda_backend.avg_voxelize_forward(z(1, 2, 0), z(1, 3, 0, dt=torch.int32), 2)
    torch.cuda.synchronize()
    assert out.shape == (1, 2, 8) and not out.any() and not cnt.any()
    assert cuda_backend.grouping_forward(z(0, 4, 8), z(0, 2, 2, dt=torch.int32)).shape == (0, 4, 2, 2)
    assert cuda_backend.ball_query(z(1, 3, 0), z(1, 3, 4), 0.1, 3).shape == (1, 0, 3)
    nb = cuda_backend.ball_query(z(1, 3, 2) + 50.0, z(1, 3, 4), 0.1, 3)
    assert (nb == 0).all()


@pytest.mark.parametrize("seed", range(10))
def test_sparse_conv_route_random(seed, cuda_backend):
    """compact averages -> gather (both layouts, with / without bias and statistics) and the channels-last
    devoxelize on odd sizes: against the oracle"""
    import torch

    import oracle as O
    rng = _rng(100 + seed)
    b, c, n = int(rng.integers(1, 4)), int(rng.integers(1, 40)), int(rng.integers(1, 2500))
    r = int(rng.choice([1, 2, 4, 8, 16, 32]))
    cout = int(rng.choice([1, 3, 8, 32, 33, 64, 70]))
    nc = (rng.random((b, 3, n)) * (r - 1)).astype(np.float32)
    vox = np.round(nc).astype(np.int32)
    feat = rng.standard_normal((b, c, n)).astype(np.float32)
    plan = cuda_backend.voxel_plan(torch.as_tensor(vox).cuda(), r)
    comp = cuda_backend.avg_voxelize_compact(torch.as_tensor(feat).cuda(), plan).cpu().numpy()
    want_comp, occupied = O.avg_voxelize_compact(feat, vox, r)
    assert np.abs(comp - want_comp).max() <= 1e-4 * max(np.abs(want_comp).max(), 1e-30)
    taps = rng.standard_normal((b, n, 27 * cout)).astype(np.float32)
    bias = rng.standard_normal((cout,)).astype(np.float32)
    tt = torch.as_tensor(taps).cuda()
    want = O.sparse_conv3_gather(taps, occupied, r, bias)
    got = cuda_backend.sparse_conv3_gather(tt, plan, torch.as_tensor(bias).cuda()).cpu().numpy()
    assert np.array_equal(got, want)
    got_cl, stats = cuda_backend.sparse_conv3_gather(tt, plan, torch.as_tensor(bias).cuda(), channels_last=True, stats=True)
    assert np.array_equal(got_cl.permute(0, 4, 1, 2, 3).cpu().numpy(), want)
    nobias = O.sparse_conv3_gather(taps, occupied, r, None).astype(np.float64).reshape(b, cout, -1)
    assert np.abs(stats.sum(1)[..., 0].cpu().numpy() - nobias.sum(-1)).max() <= 1e-4 * max(np.abs(nobias).sum(-1).max(), 1.0)
    grid = rng.standard_normal((b, cout, r, r, r)).astype(np.float32)
    want_dev = O.trilinear_devoxelize_forward(r, False, nc, grid.reshape(b, cout, -1))[0]
    grid_cl = torch.as_tensor(grid).cuda().permute(0, 2, 3, 4, 1).contiguous()
    got_dev = cuda_backend.trilinear_devoxelize_cl(grid_cl, torch.as_tensor(nc).cuda(), r).cpu().numpy()
    assert np.array_equal(got_dev, want_dev)


@pytest.mark.parametrize("seed", range(10))
def test_groupnorm_variants_random(seed, cuda_backend):
    """channel-first / channels-last, one-pass / two-pass selection on random shapes: against torch fp32"""
    import torch
    import torch.nn.functional as TF
    rng = _rng(200 + seed)
    b = int(rng.integers(1, 5))
    c = int(rng.choice([8, 16, 24, 32, 64, 128, 256]))
    spatial = tuple(int(v) for v in rng.integers(1, 13, size=3))
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn((b, c) + spatial, device="cuda", generator=g) * 1.7 + 0.4
    w, bias, cb = (torch.randn(c, device="cuda", generator=g) for _ in range(3))
    want = TF.group_norm(x + cb.view(1, c, 1, 1, 1), 8, w, bias, 1e-5)
    want = want * torch.sigmoid(want)
    peak = want.abs().max().item()
    got = cuda_backend.groupnorm_act(x, 8, w, bias, 1e-5, True, conv_bias=cb)
    assert (got - want).abs().max().item() <= 1e-5 * peak
    if cuda_backend.groupnorm_cl_supported(c, 8):
        x_cl = x.permute(0, 2, 3, 4, 1).contiguous()
        got_cl, sums = cuda_backend.groupnorm_act_cl(x_cl, 8, w, bias, 1e-5, True, conv_bias=cb, channel_sums=True)
        assert (got_cl.permute(0, 4, 1, 2, 3) - want).abs().max().item() <= 1e-5 * peak
        want_sums = want.double().flatten(2).sum(-1)
        assert (sums.double() - want_sums).abs().max().item() <= 1e-5 * max(want_sums.abs().max().item(), 1.0)
