"""The sparse first convolution of a PVConv block (csrc/sparse_conv.cu, modules/point_voxel.py).

CPU: the oracle's restatement (oracle.avg_voxelize_compact + a GEMM + oracle.sparse_conv3_gather) equals
torch's Conv3d on the reference's dense voxel grid -- that pins the checker to what the reference computes
(modules/pvconv.py:75-76,91-97).
GPU: the two kernels against that oracle (compact: bit-exact; gather: bit-exact, same fp32 add order), and
the whole PVConv block sparse vs dense.
"""
import numpy as np
import pytest

import oracle


def _cloud(b, c, n, r, seed, clustered=True):
    rng = np.random.default_rng(seed)
    if clustered:   # a surface-like cloud: most voxels stay empty
        u = rng.normal(size=(b, 3, n)).astype(np.float32)
        u /= np.linalg.norm(u, axis=1, keepdims=True)
        pts = 0.5 + 0.35 * u + 0.01 * rng.normal(size=(b, 3, n)).astype(np.float32)
    else:
        pts = rng.random((b, 3, n), dtype=np.float32)
    coords = np.clip(np.round(pts * r), 0, r - 1).astype(np.int32)
    feats = rng.normal(size=(b, c, n)).astype(np.float32)
    return feats, coords


def _taps_matrix(w):
    """Conv3d weight [Cout,Cin,3,3,3] -> [Cin, 27*Cout], column k*Cout+co"""
    return np.ascontiguousarray(np.transpose(w, (1, 2, 3, 4, 0)).reshape(w.shape[1], -1))


@pytest.mark.parametrize("b,cin,cout,n,r", [(2, 5, 4, 300, 8), (1, 3, 8, 64, 4), (2, 6, 3, 500, 16)])
def test_oracle_sparse_conv_equals_dense_conv3d(b, cin, cout, n, r):
    import torch
    feats, coords = _cloud(b, cin, n, r, seed=n + r)
    rng = np.random.default_rng(7)
    w = rng.normal(size=(cout, cin, 3, 3, 3)).astype(np.float32)
    bias = rng.normal(size=(cout,)).astype(np.float32)
    dense, _, _ = oracle.avg_voxelize_forward(feats, coords, r)
    want = torch.nn.functional.conv3d(torch.from_numpy(dense.reshape(b, cin, r, r, r)).double(),
                                      torch.from_numpy(w).double(), torch.from_numpy(bias).double(), padding=1).numpy()
    compact, occupied = oracle.avg_voxelize_compact(feats, coords, r)
    taps = np.einsum("bcn,ck->bnk", compact.astype(np.float64), _taps_matrix(w).astype(np.float64)).astype(np.float32)
    got = oracle.sparse_conv3_gather(taps, occupied, r, bias)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()


GPU_CASES = [(2, 7, 32, 1000, 32, True), (3, 64, 64, 4096, 32, True), (2, 5, 40, 777, 16, False),
             (1, 3, 8, 100, 8, False), (2, 16, 32, 4096, 32, False), (1, 4, 33, 50, 4, True)]


@pytest.mark.gpu
@pytest.mark.parametrize("b,c,cout,n,r,clustered", GPU_CASES)
def test_compact_and_gather_bit_exact(b, c, cout, n, r, clustered, cuda_backend):
    import torch
    feats, coords = _cloud(b, c, n, r, seed=b * 1000 + n, clustered=clustered)
    plan = cuda_backend.voxel_plan(torch.from_numpy(coords).cuda(), r)
    got_compact = cuda_backend.avg_voxelize_compact(torch.from_numpy(feats).cuda(), plan).cpu().numpy()
    # the compact columns are the dense CUDA grid's non-empty columns, bit for bit
    dense_cuda = cuda_backend.avg_voxelize_fill(torch.from_numpy(feats).cuda(), plan).cpu().numpy()
    cnt = plan.cnt.cpu().numpy()
    for i in range(b):
        occ = np.flatnonzero(cnt[i] > 0)
        assert np.array_equal(got_compact[i][:, :len(occ)], dense_cuda[i][:, occ])
        assert not got_compact[i][:, len(occ):].any()
    # ... and match the oracle's within the voxel-average tolerance (fp32 summation order inside a voxel)
    want_compact, occupied = oracle.avg_voxelize_compact(feats, coords, r)
    assert np.abs(got_compact - want_compact).max() <= 1e-4 * max(np.abs(want_compact).max(), 1.0)

    rng = np.random.default_rng(3)
    taps = rng.normal(size=(b, n, 27 * cout)).astype(np.float32)
    bias = rng.normal(size=(cout,)).astype(np.float32)
    for use_bias in (False, True):
        got = cuda_backend.sparse_conv3_gather(torch.from_numpy(taps).cuda(), plan,
                                               torch.from_numpy(bias).cuda() if use_bias else None).cpu().numpy()
        want = oracle.sparse_conv3_gather(taps, occupied, r, bias if use_bias else None)
        assert np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("cin,cout,n,r,attention", [(19, 32, 4096, 32, False), (64, 64, 2048, 32, False), (12, 16, 512, 16, True)])
def test_pvconv_sparse_equals_dense(cin, cout, n, r, attention, cuda_backend):
    """whole block, TF32 off on both sides: 1e-5 of the output's peak"""
    import torch

    import bdm_b200.modules.point_voxel as PV
    torch.manual_seed(cin + n)
    blk = PV.PVConv(cin, cout, 3, r, attention=attention, with_se=True).cuda().eval()
    feats = torch.randn(3, cin, n, device="cuda")
    u = torch.randn(3, 3, n, device="cuda")
    coords = u / u.norm(dim=1, keepdim=True) * (0.5 + 0.02 * torch.randn(3, 1, n, device="cuda"))
    temb = torch.randn(3, 8, n, device="cuda")
    saved = (PV.SPARSE_FIRST_CONV, PV.SPARSE_MAX_FILL, torch.backends.cudnn.allow_tf32)
    try:
        torch.backends.cudnn.allow_tf32 = False
        PV.SPARSE_MAX_FILL = 1.0
        with torch.no_grad():
            PV.SPARSE_FIRST_CONV = True
            assert blk._sparse_eligible(feats)
            y_sparse = blk((feats, coords, temb))[0]
            PV.SPARSE_FIRST_CONV = False
            y_dense = blk((feats, coords, temb))[0]
        err = (y_sparse - y_dense).abs().max().item() / y_dense.abs().max().item()
        assert err <= 1e-5, err
        # default precision (TF32 on both sides): same sums, different rounding points
        torch.backends.cudnn.allow_tf32 = True
        with torch.no_grad():
            PV.SPARSE_FIRST_CONV = True
            y_sparse = blk((feats, coords, temb))[0]
            PV.SPARSE_FIRST_CONV = False
            y_dense32 = blk((feats, coords, temb))[0]
        e_sparse = (y_sparse - y_dense).abs().max().item()
        e_dense = (y_dense32 - y_dense).abs().max().item()
        assert e_sparse <= max(3 * e_dense, 1e-5 * y_dense.abs().max().item()), (e_sparse, e_dense)
    finally:
        PV.SPARSE_FIRST_CONV, PV.SPARSE_MAX_FILL, torch.backends.cudnn.allow_tf32 = saved


@pytest.mark.gpu
@pytest.mark.parametrize("b,cout,n,r", [(2, 32, 1000, 32), (3, 64, 4096, 32), (2, 40, 500, 16), (1, 8, 60, 8)])
def test_gather_channels_last_is_the_same_tensor(b, cout, n, r, cuda_backend):
    import torch
    feats, coords = _cloud(b, 4, n, r, seed=n)
    plan = cuda_backend.voxel_plan(torch.from_numpy(coords).cuda(), r)
    g = torch.Generator(device="cuda").manual_seed(1)
    taps = torch.randn(b, n, 27 * cout, device="cuda", generator=g)
    bias = torch.randn(cout, device="cuda", generator=g)
    a = cuda_backend.sparse_conv3_gather(taps, plan, bias)
    c = cuda_backend.sparse_conv3_gather(taps, plan, bias, channels_last=True)
    assert c.shape == (b, r, r, r, cout) and c.is_contiguous()
    assert torch.equal(c.permute(0, 4, 1, 2, 3), a)


@pytest.mark.gpu
@pytest.mark.parametrize("b,c,n,r", [(2, 64, 4096, 32), (3, 32, 1000, 32), (1, 5, 77, 8), (2, 130, 300, 16)])
def test_devoxelize_channels_last_bit_exact(b, c, n, r, cuda_backend):
    """same weights, corner order and fma chain as the channel-first op: bit-exact against the oracle"""
    import torch
    rng = np.random.default_rng(c + n)
    grid = rng.normal(size=(b, c, r, r, r)).astype(np.float32)
    coords = (rng.random((b, 3, n), dtype=np.float32) * (r - 1)).astype(np.float32)
    coords[:, :, : min(n, 16)] = np.round(coords[:, :, : min(n, 16)])          # points on voxel centres / faces
    want = oracle.trilinear_devoxelize_forward(r, False, coords, grid.reshape(b, c, -1))[0]
    grid_cl = torch.from_numpy(grid).cuda().permute(0, 2, 3, 4, 1).contiguous()
    got = cuda_backend.trilinear_devoxelize_cl(grid_cl, torch.from_numpy(coords).cuda(), r).cpu().numpy()
    assert np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("b,cout,n,r", [(2, 64, 4096, 32), (3, 32, 900, 32), (2, 40, 300, 16), (1, 16, 5, 2)])
def test_gather_statistics_feed_the_norm(b, cout, n, r, cuda_backend):
    """the per-channel (sum, sumsq) blocks the gather emits equal the tensor's, and GroupNorm from them equals
    GroupNorm with its own statistics pass"""
    import torch
    feats, coords = _cloud(b, 4, n, r, seed=n + 1)
    plan = cuda_backend.voxel_plan(torch.from_numpy(coords).cuda(), r)
    g = torch.Generator(device="cuda").manual_seed(5)
    taps = torch.randn(b, n, 27 * cout, device="cuda", generator=g)
    out, stats = cuda_backend.sparse_conv3_gather(taps, plan, channels_last=True, stats=True)
    assert torch.equal(out, cuda_backend.sparse_conv3_gather(taps, plan, channels_last=True))
    flat = out.double().reshape(b, -1, cout)
    tot = stats.sum(dim=1)
    assert (tot[..., 0] - flat.sum(1)).abs().max().item() <= 1e-4 * max(flat.abs().sum(1).max().item(), 1.0)
    assert (tot[..., 1] - (flat * flat).sum(1)).abs().max().item() <= 1e-5 * (flat * flat).sum(1).max().item()
    if cuda_backend.groupnorm_cl_supported(cout, 8):
        w = torch.randn(cout, device="cuda", generator=g)
        bias = torch.randn(cout, device="cuda", generator=g)
        cb = torch.randn(cout, device="cuda", generator=g)
        a = cuda_backend.groupnorm_act_cl(out, 8, w, bias, 1e-5, True, conv_bias=cb)
        c = cuda_backend.groupnorm_act_cl(out, 8, w, bias, 1e-5, True, conv_bias=cb, partials=stats)
        assert (a - c).abs().max().item() <= 2e-6 * a.abs().max().item()
