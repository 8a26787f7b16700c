"""GPU: the tcgen05 3x3x3 convolution route (csrc/conv3_tc05.cu) -- GroupNorm+Swish -> fp16 chunk planes ->
implicit-GEMM convolution with bias and GroupNorm statistics -- against float64 torch, and against cuDNN's TF32
convolution, which is what the reference runs here (modules/pvconv.py:75-88 under torch's default conv policy).

Tolerance: the route rounds operands to 11 significant bits exactly like TF32 does, so its error against float64
must not exceed cuDNN-TF32's error on the same input (x 1.5 for sampling noise), and both sit near 3e-4 of the
output's peak; statistics are compared at 1e-5."""
import pytest

pytestmark = pytest.mark.gpu

# (batch, resolution, c_in, c_out): every template instance of the kernel, odd batches, ragged last units
CASES = [(2, 8, 32, 32), (3, 16, 64, 64), (1, 32, 32, 32), (2, 16, 128, 128), (1, 32, 64, 64), (2, 8, 256, 128),
         (2, 16, 64, 32), (5, 16, 32, 64), (3, 8, 32, 128), (1, 4, 64, 64)]


def _inputs(b, r, cin, cout, seed, gamma_scale=1.0, w_scale=1.0):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(b, r, r, r, cin, device="cuda", generator=g) * 1.7 + 0.3
    cb = torch.randn(cin, device="cuda", generator=g) * 0.2
    gamma = (torch.rand(cin, device="cuda", generator=g) + 0.5) * gamma_scale
    beta = torch.randn(cin, device="cuda", generator=g) * 0.1 * gamma_scale
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda", generator=g) / (27 * cin) ** 0.5 * w_scale
    bias = torch.randn(cout, device="cuda", generator=g) * 0.1 * w_scale * gamma_scale
    return x, cb, gamma, beta, w, bias


def _partials(x):
    import torch
    b, c = x.shape[0], x.shape[-1]
    xs = x.reshape(b, -1, c).double()
    return torch.stack([xs.sum(1), (xs * xs).sum(1)], dim=-1).reshape(b, 1, c, 2).contiguous()


def _route(B, x, cb, gamma, beta, w, bias, groups=8, eps=1e-5):
    b, r, cin = x.shape[0], x.shape[1], x.shape[-1]
    prepared = B.conv3_tc05_prepare(w, gamma, beta, (cin // groups) * r ** 3)
    planes = B.HalfPlanes(b, cin, r, x.device)
    B.groupnorm_swish_half_planar(x, groups, gamma, beta, eps, True, cb, _partials(x), prepared, planes)
    out, stats = B.conv3_tc05(planes, prepared, w.shape[0], bias=bias, stats=True)
    return out, stats, planes, prepared


def _reference64(x, cb, gamma, beta, w, bias, groups=8, eps=1e-5):
    import torch
    import torch.nn.functional as TF
    xd = (x.double() + cb.double()).permute(0, 4, 1, 2, 3)
    act = TF.group_norm(xd, groups, gamma.double(), beta.double(), eps)
    act = act * torch.sigmoid(act)
    return act, TF.conv3d(act, w.double(), bias.double(), padding=1).permute(0, 2, 3, 4, 1)


@pytest.mark.parametrize("b,r,cin,cout", CASES)
def test_conv3_tc05_vs_float64_and_cudnn_tf32(b, r, cin, cout, cuda_backend):
    import torch
    import torch.nn.functional as TF
    B = cuda_backend
    assert B.conv3_tc05_supported(cin, cout, r)
    x, cb, gamma, beta, w, bias = _inputs(b, r, cin, cout, 1000 * b + r + cin)
    out, stats, _, _ = _route(B, x, cb, gamma, beta, w, bias)
    act, ref = _reference64(x, cb, gamma, beta, w, bias)
    saved = torch.backends.cudnn.allow_tf32
    try:
        torch.backends.cudnn.allow_tf32 = True
        tf32 = TF.conv3d(act.float().contiguous(memory_format=torch.channels_last_3d),
                         w.contiguous(memory_format=torch.channels_last_3d), bias, padding=1).permute(0, 2, 3, 4, 1)
    finally:
        torch.backends.cudnn.allow_tf32 = saved
    peak = ref.abs().max().item()
    e_ours = (out.double() - ref).abs().max().item() / peak
    e_tf32 = (tf32.double() - ref).abs().max().item() / peak
    assert e_ours <= 1.5 * e_tf32 + 2e-6, (e_ours, e_tf32)
    assert e_ours <= 1e-3, e_ours
    # statistics of the result: per-group sums in the group's first channel slot, zeros elsewhere
    od = out.double().reshape(b, -1, cout)
    cg = cout // 8
    want1 = od.sum(1).reshape(b, 8, cg).sum(-1)
    want2 = (od * od).sum(1).reshape(b, 8, cg).sum(-1)
    got = stats.reshape(b, 8, cg, 2)
    assert ((got[:, :, 0, 0] - want1).abs() <= 1e-5 * (want2.sqrt() + 1)).all()
    assert ((got[:, :, 0, 1] - want2).abs() <= 1e-5 * (want2 + 1)).all()
    if cg > 1:
        assert got[:, :, 1:].abs().max().item() == 0.0


@pytest.mark.parametrize("b,r,c", [(3, 32, 64), (2, 16, 128), (2, 32, 32)])
def test_group_statistics_go_to_the_consumers_unfolded(b, r, c, cuda_backend):
    """conv3_tc05(stats="groups") leaves its per-unit group partials; every consumer folds them in its prologue and must
    agree with the folded form (stats=True) to rounding of the double sums"""
    import torch
    B = cuda_backend
    x, cb, gamma, beta, w, bias = _inputs(b, r, c, c, 31 + r + c)
    prepared = B.conv3_tc05_prepare(w, gamma, beta, (c // 8) * r ** 3)
    planes = B.HalfPlanes(b, c, r, x.device)
    B.groupnorm_swish_half_planar(x, 8, gamma, beta, 1e-5, True, cb, _partials(x), prepared, planes)
    out, folded = B.conv3_tc05(planes, prepared, c, bias=bias, stats=True)
    out2, groups = B.conv3_tc05(planes, prepared, c, bias=bias, stats="groups")
    assert torch.equal(out, out2)
    cg = c // 8
    assert torch.allclose(groups.data.sum(1), folded[:, 0, ::cg], rtol=1e-12, atol=0)
    of = out.reshape(b, -1, c)
    y1, s1 = B.groupnorm_act_cl(of, 8, gamma, beta, 1e-5, True, channel_sums="tiles", partials=folded)
    y2, s2 = B.groupnorm_act_cl(of, 8, gamma, beta, 1e-5, True, channel_sums="tiles", partials=groups)
    assert (y1 - y2).abs().max().item() <= 1e-6 * y1.abs().max().item()
    assert torch.allclose(s1, s2, rtol=1e-5, atol=1e-4)
    t1, c1 = B.groupnorm_cl_sums(of, 8, gamma, beta, 1e-5, True, None, folded)
    t2, c2 = B.groupnorm_cl_sums(of, 8, gamma, beta, 1e-5, True, None, groups)
    assert torch.allclose(c1, c2, rtol=1e-6, atol=1e-7)
    p1, p2 = B.HalfPlanes(b, c, r, x.device), B.HalfPlanes(b, c, r, x.device)
    B.groupnorm_swish_half_planar(out, 8, gamma, beta, 1e-5, True, None, folded, prepared, p1)
    B.groupnorm_swish_half_planar(out, 8, gamma, beta, 1e-5, True, None, groups, prepared, p2)
    assert (p1.data.float() - p2.data.float()).abs().max().item() <= 2e-3 * p1.data.float().abs().max().item()


@pytest.mark.parametrize("b,r,c", [(2, 16, 64), (3, 8, 32), (1, 32, 128)])
def test_half_planes_hold_the_activation_and_zero_pads(b, r, c, cuda_backend):
    """decode the fp16 chunk planes back to [B,R,R,R,C]: GroupNorm+Swish to fp16 rounding; every other row is zero"""
    import torch
    import torch.nn.functional as TF
    B = cuda_backend
    x, cb, gamma, beta, w, bias = _inputs(b, r, c, c, 77 + r)
    _, _, planes, prepared = _route(B, x, cb, gamma, beta, w, bias)
    act_scale = prepared[:12].view(torch.float32)[1].item()
    assert act_scale == 1.0
    q = r + 1
    guard = (q * q + q + 1 + 7) // 8 * 8
    srows = (guard + q ** 3 + 7) // 8 * 8
    data = planes.data                                            # [C/8, rows, 8]
    full = data.permute(1, 0, 2).reshape(planes.rows, c)          # [rows, C]
    want = TF.group_norm((x + cb).permute(0, 4, 1, 2, 3), 8, gamma, beta, 1e-5)
    want = (want * torch.sigmoid(want)).permute(0, 2, 3, 4, 1)
    mask = torch.zeros(planes.rows, dtype=torch.bool, device="cuda")
    for i in range(b):
        vol = full[guard + i * srows: guard + i * srows + q ** 3].reshape(q, q, q, c)
        got = vol[:r, :r, :r].float()
        assert (got - want[i]).abs().max().item() <= 1e-3 * want.abs().max().item()     # fp16: 2^-11 relative
        assert vol[r].abs().max().item() == 0 and vol[:, r].abs().max().item() == 0 and vol[:, :, r].abs().max().item() == 0
        m = torch.zeros(q, q, q, dtype=torch.bool, device="cuda")
        m[:r, :r, :r] = True
        mask[guard + i * srows: guard + i * srows + q ** 3] = m.reshape(-1)
    assert full[~mask].abs().max().item() == 0


def test_scaling_keeps_extreme_operands_in_range(cuda_backend):
    """huge GroupNorm affine (activations beyond fp16's range without the scale) and tiny weights (below fp16's
    normal range without the scale): same relative accuracy"""
    import torch
    B = cuda_backend
    for gs, ws in ((3.0e3, 1.0), (1.0, 1.0e-7), (2.0e3, 1.0e-6)):
        x, cb, gamma, beta, w, bias = _inputs(2, 16, 64, 64, 5, gamma_scale=gs, w_scale=ws)
        out, _, _, prepared = _route(B, x, cb, gamma, beta, w, bias)
        _, ref = _reference64(x, cb, gamma, beta, w, bias)
        assert torch.isfinite(out).all()
        err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
        assert err <= 1e-3, (gs, ws, err)
        hdr = prepared[:12].view(torch.float32)
        if gs > 100:
            assert hdr[1].item() < 1.0      # activations were scaled down


def test_voxel_stack_takes_the_tcgen05_route_and_matches_cudnn(cuda_backend):
    """PVConv blocks at the network's (channels, resolution) pairs: default route (tcgen05 second convolution)
    against the same block with the route off (cuDNN TF32) and against fp32 convolutions"""
    import torch

    import bdm_b200.modules.layers as L
    import bdm_b200.modules.point_voxel as PV
    B = cuda_backend
    for cin, cout, n, r, attention in ((32, 32, 4096, 32, False), (64, 64, 2048, 32, False), (64, 128, 1024, 16, False),
                                       (32, 64, 1024, 16, True)):
        torch.manual_seed(cin + n)
        blk = PV.PVConv(cin, cout, 3, r, attention=attention, with_se=True).cuda().eval()
        feats = torch.randn(3, cin, n, device="cuda")
        u = torch.randn(3, 3, n, device="cuda")
        coords = u / u.norm(dim=1, keepdim=True) * (0.5 + 0.02 * torch.randn(3, 1, n, device="cuda"))
        temb = torch.randn(3, 8, n, device="cuda")
        saved = (L.CONV3_TC05, torch.backends.cudnn.allow_tf32)
        try:
            with torch.no_grad():
                torch.backends.cudnn.allow_tf32 = True
                L.CONV3_TC05 = True
                n0 = B.LAUNCHES
                B.profile_start()
                y_tc = blk((feats, coords, temb))[0]
                prof = B.profile_stop()
                assert "conv3_tc05" in prof and "groupnorm_swish_half_planar" in prof, sorted(prof)
                L.CONV3_TC05 = False
                y_cudnn = blk((feats, coords, temb))[0]
                torch.backends.cudnn.allow_tf32 = False
                y_fp32 = blk((feats, coords, temb))[0]
        finally:
            L.CONV3_TC05, torch.backends.cudnn.allow_tf32 = saved
        peak = y_fp32.abs().max().item()
        e_tc = (y_tc - y_fp32).abs().max().item() / peak
        e_cudnn = (y_cudnn - y_fp32).abs().max().item() / peak
        assert e_tc <= max(2.0 * e_cudnn, 2e-4), (cin, cout, r, e_tc, e_cudnn)


@pytest.mark.parametrize("b,c,n,r", [(3, 64, 2048, 32), (2, 32, 1000, 16), (2, 128, 1024, 16)])
def test_fill_planes_from_a_voxelized_cloud(b, c, n, r, cuda_backend):
    """occupied-voxel averages + plan -> fp16 planes == the dense voxel grid (fp16 rounding of the scaled values),
    zeros at empty voxels and pads; the scale in the prepared header is the power of two the kernel derived"""
    import numpy as np
    import torch
    from tests import cases
    B = cuda_backend
    rng = np.random.default_rng(b * 7 + r)
    co = cases.cloud(rng, b, n, "shape")
    vox, _ = cases.vox_coords(co, r)
    plan = B.voxel_plan(torch.from_numpy(vox).cuda(), r)
    feats = torch.randn(b, c, n, device="cuda") * 37.0
    dense = B.avg_voxelize_fill(feats, plan).reshape(b, c, r, r, r)
    w = torch.randn(c, c, 3, 3, 3, device="cuda")
    prepared = B.conv3_tc05_prepare(w, None, None, 1)
    planes = B.HalfPlanes(b, c, r, "cuda")
    planes.data.fill_(7.0)                                     # stale contents of real rows must be overwritten
    q = r + 1
    guard = (q * q + q + 1 + 7) // 8 * 8
    srows = (guard + q ** 3 + 7) // 8 * 8
    B.conv3_tc05_fill_planes(B.avg_voxelize_compact(feats, plan), plan, prepared, planes)
    hdr = prepared[:16].view(torch.float32)
    act_scale = hdr[1].item()
    amax = dense.abs().max().item()
    assert 2 ** 13 <= amax * act_scale < 2 ** 14
    assert hdr[0].item() == hdr[3].item() / act_scale
    full = planes.data.permute(1, 0, 2).reshape(planes.rows, c).float()
    for i in range(b):
        vol = full[guard + i * srows: guard + i * srows + q ** 3].reshape(q, q, q, c)[:r, :r, :r]
        want = dense[i].permute(1, 2, 3, 0) * act_scale
        assert (vol - want).abs().max().item() <= 2 ** -10 * amax * act_scale
        assert (vol[want == 0] == 0).all()
    # the same with max|average| measured by the compact voxelize kernel itself: identical planes and scales
    prepared2 = B.conv3_tc05_prepare(w, None, None, 1)
    planes2 = B.HalfPlanes(b, c, r, "cuda")
    planes2.data.fill_(7.0)
    B.conv3_tc05_fill_planes(B.avg_voxelize_compact(feats, plan, amax_into=prepared2), plan, prepared2, planes2, amax_ready=True)
    assert torch.equal(planes2.data, planes.data)
    assert torch.equal(prepared2[:20], prepared[:20])


@pytest.mark.parametrize("b,cin,cout,n,r,regime", [(3, 64, 64, 4096, 32, "shape"), (2, 32, 32, 4096, 32, "noise"), (2, 128, 128, 1024, 16, "shape"),
                                                    (2, 64, 32, 300, 32, "shape"), (1, 64, 64, 8, 16, "shape")])
def test_skipping_empty_windows_is_exact(b, cin, cout, n, r, regime, cuda_backend):
    """conv3_tc05(sparse=True) -- loads and MMAs of all-zero tap windows skipped via the occupancy bits the plane
    writer leaves -- is bit-identical to the dense pass over the same planes (the skipped products are zeros), down
    to a cloud of 8 points where almost every unit is skipped entirely"""
    import numpy as np
    import torch
    from tests import cases
    B = cuda_backend
    rng = np.random.default_rng(b + cin + n)
    co = cases.cloud(rng, b, n, regime)
    vox, _ = cases.vox_coords(co, r)
    plan = B.voxel_plan(torch.from_numpy(vox).cuda(), r)
    feats = torch.randn(b, cin, n, device="cuda")
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda") / (27 * cin) ** 0.5
    bias = torch.randn(cout, device="cuda")
    prepared = B.conv3_tc05_prepare(w, None, None, 1)
    planes = B.HalfPlanes(b, cin, r, "cuda")
    B.conv3_tc05_fill_planes(B.avg_voxelize_compact(feats, plan, amax_into=prepared), plan, prepared, planes, amax_ready=True)
    # the occupancy bits are exactly the non-zero rows' positions
    q = r + 1
    guard = (q * q + q + 1 + 7) // 8 * 8
    bits = planes.occ.cpu().numpy().view(np.uint32)
    for i in range(b):
        p = np.unique((vox[i, 0].astype(np.int64) * q + vox[i, 1]) * q + vox[i, 2]) + guard
        want = np.zeros(bits.shape[1] * 32, dtype=bool)
        want[p] = True
        got = np.unpackbits(bits[i].view(np.uint8), bitorder="little").astype(bool)
        assert (got == want).all()
    dense, st_d = B.conv3_tc05(planes, prepared, cout, bias=bias, stats=True, sparse=False)
    sparse, st_s = B.conv3_tc05(planes, prepared, cout, bias=bias, stats=True, sparse=True)
    assert torch.equal(dense, sparse)
    assert torch.equal(st_d, st_s)
    # and both are the convolution of the voxelized grid
    grid = B.avg_voxelize_fill(feats, plan).reshape(b, cin, r, r, r)
    ref = torch.nn.functional.conv3d(grid.double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 4, 1)
    assert (sparse.double() - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()


def test_first_convolution_dense_route_matches_sparse_route(cuda_backend):
    """block level: first convolution through the tcgen05 kernel (DENSE_FIRST_TC05) against the tap-product route"""
    import torch

    import bdm_b200.modules.point_voxel as PV
    B = cuda_backend
    for cin, cout, n, r in ((64, 64, 4096, 32), (128, 128, 1024, 16), (32, 32, 4096, 32), (128, 64, 1024, 16)):
        torch.manual_seed(cin + r)
        blk = PV.PVConv(cin, cout, 3, r, with_se=True).cuda().eval()
        feats = torch.randn(2, cin, n, device="cuda") * 2.0
        u = torch.randn(2, 3, n, device="cuda")
        coords = u / u.norm(dim=1, keepdim=True) * (0.5 + 0.02 * torch.randn(2, 1, n, device="cuda"))
        temb = torch.randn(2, 8, n, device="cuda")
        saved = (PV.DENSE_FIRST_TC05, torch.backends.cudnn.allow_tf32)
        try:
            with torch.no_grad():
                torch.backends.cudnn.allow_tf32 = True
                PV.DENSE_FIRST_TC05 = True
                assert blk._dense_first_eligible(feats)
                B.profile_start()
                y_dense = blk((feats, coords, temb))[0]
                prof = B.profile_stop()
                assert "conv3_tc05_fill_planes" in prof and len(prof["conv3_tc05"]) == 2 and "sparse_conv3_gather" not in prof
                PV.DENSE_FIRST_TC05 = False
                y_sparse = blk((feats, coords, temb))[0]
                torch.backends.cudnn.allow_tf32 = False
                y_fp32 = blk((feats, coords, temb))[0]
        finally:
            PV.DENSE_FIRST_TC05, torch.backends.cudnn.allow_tf32 = saved
        peak = y_fp32.abs().max().item()
        e_dense = (y_dense - y_fp32).abs().max().item() / peak
        e_sparse = (y_sparse - y_fp32).abs().max().item() / peak
        assert e_dense <= max(2.0 * e_sparse, 2e-4), (cin, cout, r, e_dense, e_sparse)


@pytest.mark.parametrize("b,c,r,n", [(3, 64, 32, 4096), (2, 32, 16, 1000), (2, 128, 16, 1024), (1, 256, 8, 64)])
def test_norm_on_the_fly_devoxelize_is_bit_identical(b, c, r, n, cuda_backend):
    """groupnorm_cl_sums + trilinear_devoxelize_cl(norm_coef=) == groupnorm_act_cl -> trilinear_devoxelize_cl, bit for
    bit: SE squeeze sums, and the devoxelized, gated, residual-added features"""
    import torch
    B = cuda_backend
    g = torch.Generator(device="cuda").manual_seed(b + c + r)
    x = torch.randn(b, r, r, r, c, device="cuda", generator=g) * 2.0 + 0.5
    gamma = torch.rand(c, device="cuda", generator=g) + 0.5
    beta = torch.randn(c, device="cuda", generator=g) * 0.2
    coords = torch.rand(b, 3, n, device="cuda", generator=g) * (r - 1)
    gate = torch.rand(b, c, device="cuda", generator=g)
    res = torch.randn(b, c, n, device="cuda", generator=g)
    xs = x.reshape(b, -1, c).double()
    cg = c // 8
    part = torch.zeros(b, 1, c, 2, dtype=torch.float64, device="cuda")      # per-group sums in the group's first slot
    part[:, 0, ::cg, 0] = xs.sum(1).reshape(b, 8, cg).sum(-1)
    part[:, 0, ::cg, 1] = (xs * xs).sum(1).reshape(b, 8, cg).sum(-1)
    xf = x.reshape(b, -1, c)
    y, sums_ref = B.groupnorm_act_cl(xf, 8, gamma, beta, 1e-5, True, channel_sums="tiles", partials=part)
    want = B.trilinear_devoxelize_cl(y.reshape(b, r, r, r, c), coords, r, gate=gate, residual=res)
    sums, coef = B.groupnorm_cl_sums(xf, 8, gamma, beta, 1e-5, True, None, part)
    got = B.trilinear_devoxelize_cl(x, coords, r, gate=gate, residual=res, norm_coef=coef)
    if sums_ref.shape == sums.shape:       # (small groups take the one-pass norm kernel, which reports whole sums)
        assert torch.equal(sums, sums_ref)
    else:
        assert torch.allclose(sums.sum(1), sums_ref.sum(1), rtol=1e-5, atol=1e-3)
        y2 = torch.nn.functional.group_norm(x.permute(0, 4, 1, 2, 3), 8, gamma, beta, 1e-5)
        want = B.trilinear_devoxelize_cl((y2 * torch.sigmoid(y2)).permute(0, 2, 3, 4, 1).contiguous(), coords, r, gate=gate, residual=res)
        assert (got - want).abs().max().item() <= 1e-5 * want.abs().max().item()
        return
    assert torch.equal(got, want)


def test_block_with_fused_tail_is_bit_identical(cuda_backend):
    import torch

    import bdm_b200.modules.layers as L
    import bdm_b200.modules.point_voxel as PV
    B = cuda_backend
    for cin, cout, n, r in ((64, 64, 4096, 32), (128, 128, 1024, 16), (32, 32, 2048, 32)):
        torch.manual_seed(cin + r + 1)
        blk = PV.PVConv(cin, cout, 3, r, with_se=True).cuda().eval()
        feats = torch.randn(2, cin, n, device="cuda")
        u = torch.randn(2, 3, n, device="cuda")
        coords = u / u.norm(dim=1, keepdim=True) * (0.5 + 0.02 * torch.randn(2, 1, n, device="cuda"))
        temb = torch.randn(2, 8, n, device="cuda")
        saved = (L.FUSED_TAIL_NORM, L.FUSED_TAIL_NORM_MIN_R, torch.backends.cudnn.allow_tf32)
        try:
            with torch.no_grad():
                torch.backends.cudnn.allow_tf32 = True        # the convolutions' statistics come from the tcgen05 route
                L.FUSED_TAIL_NORM, L.FUSED_TAIL_NORM_MIN_R = True, 16
                B.profile_start()
                y_fused = blk((feats, coords, temb))[0]
                prof = B.profile_stop()
                assert "groupnorm_cl_sums" in prof, sorted(prof)
                L.FUSED_TAIL_NORM = False
                y_plain = blk((feats, coords, temb))[0]
        finally:
            L.FUSED_TAIL_NORM, L.FUSED_TAIL_NORM_MIN_R, torch.backends.cudnn.allow_tf32 = saved
        assert torch.equal(y_fused, y_plain), (cin, cout, r, (y_fused - y_plain).abs().max().item())


def test_prepared_weights_follow_in_place_updates(cuda_backend):
    """the fp16 weight stages are cached per parameter version: an in-place update of a convolution's weight (or of the
    GroupNorm affine that bounds its activations) must be picked up by the next forward"""
    import torch

    import bdm_b200.modules.point_voxel as PV
    torch.manual_seed(3)
    blk = PV.PVConv(64, 64, 3, 32, with_se=True).cuda().eval()
    feats = torch.randn(2, 64, 4096, device="cuda")
    u = torch.randn(2, 3, 4096, device="cuda")
    coords = u / u.norm(dim=1, keepdim=True) * 0.5
    temb = torch.randn(2, 8, 4096, device="cuda")
    saved = torch.backends.cudnn.allow_tf32
    try:
        torch.backends.cudnn.allow_tf32 = True
        with torch.no_grad():
            y0 = blk((feats, coords, temb))[0].clone()
            assert torch.equal(blk((feats, coords, temb))[0], y0)                   # deterministic, cache hit
            for conv in (blk.voxel_layers[0], blk.voxel_layers[4]):
                conv.weight.mul_(1.5)                                                # in place: same storage, new version
            y1 = blk((feats, coords, temb))[0]
            torch.backends.cudnn.allow_tf32 = False                                  # cuDNN fp32 with the updated weights
            y_ref = blk((feats, coords, temb))[0]
    finally:
        torch.backends.cudnn.allow_tf32 = saved
    assert (y1 - y0).abs().max().item() > 1e-3 * y0.abs().max().item()
    assert (y1 - y_ref).abs().max().item() <= 2e-3 * y_ref.abs().max().item()


def test_route_is_off_when_tf32_convolutions_are_off(cuda_backend):
    import torch

    import bdm_b200.modules.layers as L
    conv = torch.nn.Conv3d(64, 64, 3, padding=1).cuda()
    gn = torch.nn.GroupNorm(8, 64).cuda()
    saved = torch.backends.cudnn.allow_tf32
    try:
        torch.backends.cudnn.allow_tf32 = True
        assert L.conv3_tc05_applicable(conv, gn, 64, 32)
        assert not L.conv3_tc05_applicable(conv, gn, 64, 8)          # cuDNN is faster on 8^3 grids
        torch.backends.cudnn.allow_tf32 = False
        assert not L.conv3_tc05_applicable(conv, gn, 64, 32)
    finally:
        torch.backends.cudnn.allow_tf32 = saved
