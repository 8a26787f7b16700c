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
