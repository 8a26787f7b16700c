"""GPU: the fused GroupNorm(+Swish) kernel against a plain PyTorch fp32 reference of the same op
(torch.nn.functional.group_norm followed by x*sigmoid(x)).  Tolerance: 1e-5 relative to the output's peak."""
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [(16, 32, (32, 32, 32)), (16, 64, (1024, 32)), (4, 256, (8, 8, 8)), (2, 512, (16,)), (3, 16, (5, 7)),
          (2, 16, (3,)), (1, 128, (4096,)), (16, 64, (4096,)), (2, 24, (33,))]


@pytest.mark.parametrize("swish", [True, False])
@pytest.mark.parametrize("b,c,spatial", SHAPES)
def test_groupnorm_act_vs_torch(b, c, spatial, swish, cuda_backend):
    import torch
    import torch.nn.functional as TF
    g = torch.Generator(device="cuda").manual_seed(b * 131 + c)
    x = torch.randn((b, c) + spatial, device="cuda", generator=g) * 3.0 + 0.7
    w = torch.randn(c, device="cuda", generator=g)
    bias = torch.randn(c, device="cuda", generator=g)
    got = cuda_backend.groupnorm_act(x, 8, w, bias, 1e-5, swish)
    want = TF.group_norm(x, 8, w, bias, 1e-5)
    if swish:
        want = want * torch.sigmoid(want)
    err = (got - want).abs().max().item() / max(want.abs().max().item(), 1e-30)
    assert err <= 1e-5, err
    # against float64 too (the fused kernel reduces in double): must be at least as close as torch is
    ref64 = TF.group_norm(x.double(), 8, w.double(), bias.double(), 1e-5)
    if swish:
        ref64 = ref64 * torch.sigmoid(ref64)
    e_ours = (got.double() - ref64).abs().max().item()
    e_torch = (want.double() - ref64).abs().max().item()
    assert e_ours <= max(4 * e_torch, 1e-6 * ref64.abs().max().item())


def test_groupnorm_large_offset_is_stable(cuda_backend):
    """mean >> std: E[x^2]-mean^2 in double must not lose the variance"""
    import torch
    import torch.nn.functional as TF
    x = torch.randn(2, 16, 4096, device="cuda") * 0.01 + 100.0
    got = cuda_backend.groupnorm_act(x, 8, None, None, 1e-5, False)
    want = TF.group_norm(x.double(), 8, None, None, 1e-5).float()
    assert (got - want).abs().max().item() <= 2e-3  # torch's own fp32 result is no closer


def test_denoiser_fused_vs_unfused(cuda_backend):
    import torch

    import bdm_b200.modules.layers as L
    from bdm_b200.denoiser import PVCNN2_PC2
    torch.manual_seed(5)
    net = PVCNN2_PC2(num_classes=3, embed_dim=64, extra_feature_channels=6).cuda().eval()
    x = torch.randn(4, 9, 2048, device="cuda")
    t = torch.tensor([500.0, 3.0, 999.0, 0.0], device="cuda")
    with torch.no_grad():
        saved = (L.FUSED_NORM_ACT, torch.backends.cudnn.allow_tf32)
        try:
            # TF32 convolutions round their inputs to 10 mantissa bits, which turns a 1e-7 difference in
            # front of a conv into ~1e-3 behind it; the fused kernels are compared with fp32 convs
            torch.backends.cudnn.allow_tf32 = False
            L.FUSED_NORM_ACT = True
            y_fused = net(x, t)
            L.FUSED_NORM_ACT = False
            y_plain = net(x, t)
        finally:
            L.FUSED_NORM_ACT, torch.backends.cudnn.allow_tf32 = saved
    err = (y_fused - y_plain).abs().max().item() / y_plain.abs().max().item()
    assert err <= 1e-5, err   # 60+ norm layers deep


def test_attention_fused_vs_plain(cuda_backend):
    import torch

    import bdm_b200.modules.layers as L
    torch.manual_seed(9)
    att = L.Attention(64, 8, D=3).cuda().eval()
    x = torch.randn(2, 64, 16, 16, 16, device="cuda")   # 64 query tiles: below the fused-attention threshold
    with torch.no_grad():
        saved = L.FUSED_NORM_ACT
        try:
            L.FUSED_NORM_ACT = True
            y_fused = att(x)
            L.FUSED_NORM_ACT = False
            y_plain = att(x)
        finally:
            L.FUSED_NORM_ACT = saved
    err = (y_fused - y_plain).abs().max().item() / y_plain.abs().max().item()
    assert err <= 1e-5, err


@pytest.mark.parametrize("b,c,spatial", [(16, 32, (32, 32, 32)), (4, 64, (256, 32)), (2, 512, (16, 32)), (3, 16, (9, 4)),
                                         (16, 256, (8, 8, 8)), (2, 64, (16, 16, 16)), (3, 128, (128,)), (2, 256, (64,))])
def test_groupnorm_conv_bias_max_and_sums(b, c, spatial, cuda_backend):
    """conv bias folded into the statistics; max over the last dim; per-channel sums of the output"""
    import torch
    import torch.nn.functional as TF
    g = torch.Generator(device="cuda").manual_seed(c)
    x = torch.randn((b, c) + spatial, device="cuda", generator=g) * 2.0
    cb = torch.randn(c, device="cuda", generator=g)
    w = torch.randn(c, device="cuda", generator=g)
    bias = torch.randn(c, device="cuda", generator=g)
    shape = (1, c) + (1,) * len(spatial)
    want = TF.group_norm(x + cb.view(shape), 8, w, bias, 1e-5)
    want = want * torch.sigmoid(want)
    peak = want.abs().max().item()
    got = cuda_backend.groupnorm_act(x, 8, w, bias, 1e-5, True, conv_bias=cb)
    assert (got - want).abs().max().item() / peak <= 1e-5
    got_y, sums = cuda_backend.groupnorm_act(x, 8, w, bias, 1e-5, True, conv_bias=cb, channel_sums=True)
    assert (got_y - got).abs().max().item() <= 2e-6 * peak   # may come from different kernels (one-pass / two-kernel)
    want_sums = want.double().flatten(2).sum(-1)
    assert (sums.double() - want_sums).abs().max().item() <= 1e-5 * max(want_sums.abs().max().item(), 1.0)
    if cuda_backend.groupnorm_max_supported(spatial[-1]):
        got_max = cuda_backend.groupnorm_act(x, 8, w, bias, 1e-5, True, conv_bias=cb, max_over_last=True)
        assert got_max.shape == want.shape[:-1]
        # (the full-size result may come from the one-pass kernel, the max from the two-kernel path)
        assert (got_max - got.max(dim=-1).values).abs().max().item() <= 2e-6 * peak


def test_fused_sequential_matches_modules(cuda_backend):
    """PVConv voxel stack and SA shared MLP: fused execution vs module-by-module execution"""
    import torch

    import bdm_b200.modules.layers as L
    from bdm_b200.modules import PVConv, SharedMLP
    torch.manual_seed(2)
    pv = PVConv(16, 32, 3, 16, attention=False, with_se=True, with_se_relu=True).cuda().eval()
    mlp = SharedMLP(19, (32, 64), dim=2).cuda().eval()
    grid = torch.randn(4, 16, 16, 16, 16, device="cuda")
    grouped = torch.randn(4, 19, 128, 32, device="cuda")
    with torch.no_grad():
        saved = (L.FUSED_NORM_ACT, torch.backends.cudnn.allow_tf32)
        try:
            torch.backends.cudnn.allow_tf32 = False   # TF32 convs amplify 1e-7 input differences to ~1e-4
            L.FUSED_NORM_ACT = True
            a1, a2 = pv.voxel_layers(grid), mlp.forward_max(grouped)
            L.FUSED_NORM_ACT = False
            b1, b2 = pv.voxel_layers(grid), mlp(grouped).max(dim=-1).values
        finally:
            L.FUSED_NORM_ACT, torch.backends.cudnn.allow_tf32 = saved
    for a, b in ((a1, b1), (a2, b2)):
        assert a.shape == b.shape
        assert (a - b).abs().max().item() / b.abs().max().item() <= 2e-5


@pytest.mark.parametrize("cin,cout,spatial", [(390, 32, (4096,)), (67, 64, (256, 32)), (35, 32, (100,))])
def test_misaligned_pointwise_conv_split(cin, cout, spatial, cuda_backend):
    """1x1 conv with Cin % 4 != 0 as aligned head + tail GEMMs == the conv (fp32 on both sides)"""
    import torch
    import torch.nn as nn

    import bdm_b200.modules.layers as L
    torch.manual_seed(cin)
    conv = (nn.Conv1d if len(spatial) == 1 else nn.Conv2d)(cin, cout, 1).cuda().eval()
    x = torch.randn((3, cin) + spatial, device="cuda")
    saved = torch.backends.cudnn.allow_tf32
    try:
        torch.backends.cudnn.allow_tf32 = False
        with torch.no_grad():
            got = L.conv_no_bias(conv, x)
            want = conv._conv_forward(x, conv.weight, None)
    finally:
        torch.backends.cudnn.allow_tf32 = saved
    assert got.shape == want.shape
    assert (got - want).abs().max().item() <= 1e-5 * want.abs().max().item()
    with torch.no_grad():   # weight edits invalidate the cached split
        conv.weight.mul_(2.0)
        torch.backends.cudnn.allow_tf32 = False
        try:
            got2 = L.conv_no_bias(conv, x)
        finally:
            torch.backends.cudnn.allow_tf32 = saved
    assert (got2 - 2.0 * want).abs().max().item() <= 2e-5 * want.abs().max().item()


def test_deferred_se_gate(cuda_backend):
    """SE gate applied after devoxelization == gate applied to the grid (devoxelize is linear)"""
    import torch

    import bdm_b200.modules.point_voxel as PV
    torch.manual_seed(11)
    blk = PV.PVConv(24, 32, 3, 16, with_se=True).cuda().eval()
    feats = torch.randn(2, 24, 700, device="cuda")
    coords = torch.randn(2, 3, 700, device="cuda")
    saved = (PV.DEFER_SE_GATE, torch.backends.cudnn.allow_tf32)
    try:
        torch.backends.cudnn.allow_tf32 = False
        with torch.no_grad():
            PV.DEFER_SE_GATE = True
            y1 = blk((feats, coords, None))[0]
            PV.DEFER_SE_GATE = False
            y0 = blk((feats, coords, None))[0]
    finally:
        PV.DEFER_SE_GATE, torch.backends.cudnn.allow_tf32 = saved
    assert (y1 - y0).abs().max().item() <= 1e-5 * y0.abs().max().item()


@pytest.mark.parametrize("b,t,scale", [(2, 512, 1.0), (3, 4096, 1.0), (1, 1024, 3.0), (2, 128, 0.2)])
def test_fused_attention_vs_float64(b, t, scale, cuda_backend):
    """softmax(q^T k) applied to v, un-scaled logits: at least as close to float64 as torch's fp32 route"""
    import torch
    g = torch.Generator(device="cuda").manual_seed(t + b)
    q, k, v = (torch.randn(b, 64, t, device="cuda", generator=g) * scale for _ in range(3))
    got = cuda_backend.attention(q, k, v)
    ref = torch.matmul(v.double(), torch.softmax(torch.matmul(q.double().transpose(1, 2), k.double()), -1).transpose(1, 2))
    f32 = torch.matmul(v, torch.softmax(torch.matmul(q.transpose(1, 2), k), -1).transpose(1, 2))
    peak = ref.abs().max().item()
    e_ours = (got.double() - ref).abs().max().item() / peak
    e_torch = (f32.double() - ref).abs().max().item() / peak
    assert e_ours <= max(3 * e_torch, 1e-5), (e_ours, e_torch)


def test_attention_block_fused_kernel(cuda_backend):
    import torch

    import bdm_b200.modules.layers as L
    torch.manual_seed(3)
    blk = L.Attention(64, 8).cuda().eval()
    x = torch.randn(4, 64, 16, 16, 16, device="cuda")
    saved = (L.FUSED_ATTENTION, torch.backends.cudnn.allow_tf32)
    try:
        torch.backends.cudnn.allow_tf32 = False
        with torch.no_grad():
            L.FUSED_ATTENTION = True
            y1 = blk(x)
            L.FUSED_ATTENTION = False
            y0 = blk(x)
    finally:
        L.FUSED_ATTENTION, torch.backends.cudnn.allow_tf32 = saved
    assert (y1 - y0).abs().max().item() <= 1e-5 * y0.abs().max().item()


@pytest.mark.parametrize("b,c,t", [(32, 512, 16), (3, 64, 64), (2, 128, 4)])
def test_attention_block_short_point_set(b, c, t, cuda_backend):
    """Attention(D=1) on the bottleneck's short point set: the seven-launch route (one q|k|v GEMM, residual folded into
    the out-projection GEMM, out-conv bias folded into the norm) against the module-by-module route, fp32 both."""
    import torch

    import bdm_b200.modules.layers as L
    torch.manual_seed(c + t)
    blk = L.Attention(c, 8, D=1).cuda().eval()
    x = torch.randn(b, c, t, device="cuda")
    saved = (L.FUSED_ATTENTION, torch.backends.cudnn.allow_tf32)
    try:
        torch.backends.cudnn.allow_tf32 = False
        with torch.no_grad():
            assert blk.small_applicable(x)
            y1 = blk(x)
            L.FUSED_ATTENTION = False
            y0 = blk(x)
            assert not blk.small_applicable(x)
            ref = blk.double()(x.double())
    finally:
        L.FUSED_ATTENTION, torch.backends.cudnn.allow_tf32 = saved
        blk.float()
    peak = ref.abs().max().item()
    e1 = (y1.double() - ref).abs().max().item() / peak
    e0 = (y0.double() - ref).abs().max().item() / peak
    assert y1.shape == y0.shape and y1.is_contiguous()
    assert e1 <= max(3 * e0, 1e-5), (e1, e0)


def test_fp_module_without_concatenation(cuda_backend):
    """PointNetFPModule: first conv over (interpolated, skip) separately == conv over their concatenation"""
    import torch

    import bdm_b200.modules.layers as L
    from bdm_b200.modules import PointNetFPModule
    torch.manual_seed(4)
    fp = PointNetFPModule(in_channels=64 + 390, out_channels=(128, 64)).cuda().eval()
    pts = torch.randn(2, 3, 1024, device="cuda")
    cen = pts[:, :, :256].contiguous()
    cen_feats = torch.randn(2, 64, 256, device="cuda")
    full = torch.randn(2, 393, 1024, device="cuda")
    skip = full[:, 3:, :]                      # a channel-sliced view, as in the denoiser
    temb = torch.randn(2, 16, 256, device="cuda")
    saved = (L.FUSED_NORM_ACT, torch.backends.cudnn.allow_tf32)
    try:
        torch.backends.cudnn.allow_tf32 = False
        with torch.no_grad():
            y1 = fp((pts, cen, cen_feats, skip, temb))[0]
            # the denoiser marks the skip view as a slice of the network input: its 390 channels are then read as
            # full[:, 1:] (392, aligned) against two zero weight columns -- one GEMM, same sums
            wide = full[:, 3:, :]
            wide._bdm_slice_of = (full, 3)
            y2 = fp((pts, cen, cen_feats, wide, temb))[0]
            assert fp.mlp.layers[0]._concat_weight[0][-1] == (0, 2)
            L.FUSED_NORM_ACT = False             # module-by-module route, with torch.cat
            y0 = fp((pts, cen, cen_feats, skip, temb))[0]
    finally:
        L.FUSED_NORM_ACT, torch.backends.cudnn.allow_tf32 = saved
    assert (y1 - y0).abs().max().item() <= 1e-5 * y0.abs().max().item()
    assert (y2 - y0).abs().max().item() <= 1e-5 * y0.abs().max().item()


@pytest.mark.parametrize("swish", [True, False])
@pytest.mark.parametrize("b,c,spatial", [(16, 32, (32, 32, 32)), (4, 64, (16, 16, 16)), (2, 256, (8, 8, 8)), (3, 16, (5, 7, 3)),
                                         (2, 128, (4, 4, 4))])
def test_groupnorm_channels_last_vs_torch(b, c, spatial, swish, cuda_backend):
    import torch
    import torch.nn.functional as TF
    g = torch.Generator(device="cuda").manual_seed(b + c)
    x = torch.randn((b, c) + spatial, device="cuda", generator=g) * 2.0 + 0.3
    cb = torch.randn(c, device="cuda", generator=g)
    w = torch.randn(c, device="cuda", generator=g)
    bias = torch.randn(c, device="cuda", generator=g)
    want = TF.group_norm(x + cb.view(1, c, 1, 1, 1), 8, w, bias, 1e-5)
    if swish:
        want = want * torch.sigmoid(want)
    peak = want.abs().max().item()
    x_cl = x.permute(0, 2, 3, 4, 1).contiguous()
    assert cuda_backend.groupnorm_cl_supported(c, 8)
    got, sums = cuda_backend.groupnorm_act_cl(x_cl, 8, w, bias, 1e-5, swish, conv_bias=cb, channel_sums=True)
    assert (got.permute(0, 4, 1, 2, 3) - want).abs().max().item() / peak <= 1e-5
    want_sums = want.double().flatten(2).sum(-1)
    assert (sums.double() - want_sums).abs().max().item() <= 1e-5 * max(want_sums.abs().max().item(), 1.0)
    got2 = cuda_backend.groupnorm_act_cl(x_cl, 8, None, None, 1e-5, swish)
    want2 = TF.group_norm(x, 8, None, None, 1e-5)
    if swish:
        want2 = want2 * torch.sigmoid(want2)
    assert (got2.permute(0, 4, 1, 2, 3) - want2).abs().max().item() / want2.abs().max().item() <= 1e-5


@pytest.mark.parametrize("c,use_relu,tiles", [(64, False, 1), (256, True, 5), (32, False, 37)])
def test_se_gate_kernel(c, use_relu, tiles, cuda_backend):
    import torch

    import bdm_b200.modules.layers as L
    torch.manual_seed(c)
    se = L.SE3d(c, use_relu=use_relu).cuda().eval()
    sums = torch.randn(3, tiles, c, device="cuda") * 50.0
    count = 512.0
    with torch.no_grad():
        want = se.fc(sums.sum(dim=1) / count)
        got = cuda_backend.se_gate(sums if tiles > 1 else sums[:, 0].contiguous(), count, se.fc[0].weight, se.fc[2].weight, use_relu)
    assert got.shape == want.shape
    assert (got - want).abs().max().item() <= 1e-5


def test_devoxelize_cl_epilogue(cuda_backend):
    import torch
    g = torch.Generator(device="cuda").manual_seed(2)
    b, c, n, r = 2, 32, 500, 16
    grid = torch.randn(b, r, r, r, c, device="cuda", generator=g)
    coords = torch.rand(b, 3, n, device="cuda", generator=g) * (r - 1)
    gate = torch.rand(b, c, device="cuda", generator=g)
    resid = torch.randn(b, c, n, device="cuda", generator=g)
    plain = cuda_backend.trilinear_devoxelize_cl(grid, coords, r)
    got = cuda_backend.trilinear_devoxelize_cl(grid, coords, r, gate=gate, residual=resid)
    want = plain * gate[:, :, None] + resid
    assert (got - want).abs().max().item() <= 1e-6 * want.abs().max().item()


def test_whole_denoiser_all_fused_routes_vs_plain_torch(cuda_backend):
    """Every inference-only route at once (sparse first conv, channels-last branch, fused norms, attention,
    SE gate, split / concatenation-free 1x1 convs) against the module-by-module torch execution of the same
    network on the same 12 kernels, fp32 convolutions on both sides."""
    import torch

    import bdm_b200.modules.layers as L
    import bdm_b200.modules.point_voxel as PV
    from bdm_b200.denoiser import PVCNN2_PC2
    torch.manual_seed(7)
    net = PVCNN2_PC2(num_classes=3, embed_dim=64, extra_feature_channels=387).cuda().eval()
    x = torch.randn(8, 390, 4096, device="cuda")
    x[:, :3] = (torch.nn.functional.normalize(torch.randn(8, 3, 4096, device="cuda"), dim=1) * 0.45
                + 0.02 * torch.randn(8, 3, 4096, device="cuda"))
    t = torch.randint(0, 1000, (8,), device="cuda").float()
    saved = (L.FUSED_NORM_ACT, L.FUSED_ATTENTION, PV.SPARSE_FIRST_CONV, PV.CHANNELS_LAST_VOXELS, PV.DEFER_SE_GATE,
             torch.backends.cudnn.allow_tf32)
    try:
        torch.backends.cudnn.allow_tf32 = False
        with torch.no_grad():
            y_fused = net(x, t)
            L.FUSED_NORM_ACT = L.FUSED_ATTENTION = False
            PV.SPARSE_FIRST_CONV = PV.CHANNELS_LAST_VOXELS = PV.DEFER_SE_GATE = False
            y_plain = net(x, t)
    finally:
        (L.FUSED_NORM_ACT, L.FUSED_ATTENTION, PV.SPARSE_FIRST_CONV, PV.CHANNELS_LAST_VOXELS, PV.DEFER_SE_GATE,
         torch.backends.cudnn.allow_tf32) = saved
    assert torch.isfinite(y_fused).all()
    err = (y_fused - y_plain).abs().max().item() / y_plain.abs().max().item()
    assert err <= 5e-5, err   # ~70 layers deep, each route within 1e-5 of the ops it replaces


@pytest.mark.parametrize("b,c,spatial,u", [(2, 64, (1024, 32), 32), (3, 32, (1024, 32), 0), (2, 128, (256, 32), 32),
                                           (2, 128, (4096,), 0), (2, 512, (16, 32), 32), (1, 24, (40, 4), 4)])
def test_groupnorm_cluster_kernel_channel_first(b, c, spatial, u, cuda_backend):
    """Groups of 64 KB .. 256 KB (and every max-over-neighbours call up to that size) take the thread-block
    cluster kernel -- slices in shared memory by TMA, moments exchanged through DSMEM, one read of the tensor --;
    the 0.5 and 1 MB groups here take the two-kernel path.  Against torch, with and without the conv bias."""
    import torch
    import torch.nn.functional as TF
    g = torch.Generator(device="cuda").manual_seed(c + u)
    x = torch.randn((b, c) + spatial, device="cuda", generator=g) * 1.5 - 0.4
    cb = torch.randn(c, device="cuda", generator=g)
    w = torch.randn(c, device="cuda", generator=g)
    bias = torch.randn(c, device="cuda", generator=g)
    shape = (1, c) + (1,) * len(spatial)
    for conv_bias in (cb, None):
        want = TF.group_norm(x + (conv_bias.view(shape) if conv_bias is not None else 0.0), 8, w, bias, 1e-5)
        want = want * torch.sigmoid(want)
        if u:
            want = want.max(dim=-1).values
        got = cuda_backend.groupnorm_act(x, 8, w, bias, 1e-5, True, conv_bias=conv_bias, max_over_last=bool(u))
        assert got.shape == want.shape
        assert (got - want).abs().max().item() <= 1e-5 * want.abs().max().item()
        again = cuda_backend.groupnorm_act(x, 8, w, bias, 1e-5, True, conv_bias=conv_bias, max_over_last=bool(u))
        assert torch.equal(got, again)            # fixed reduction order: deterministic


@pytest.mark.parametrize("b,c,r,swish", [(2, 64, 32, True), (3, 32, 32, True), (2, 128, 16, False), (1, 16, 32, True)])
def test_groupnorm_channels_last_large_samples(b, c, r, swish, cuda_backend):
    """Samples of 1 MB and more in channels-last memory (the R = 32 / R = 16 voxel grids of the step): statistics +
    software-pipelined apply, conv bias folded in, per-tile channel sums for the SE gate."""
    import torch
    import torch.nn.functional as TF
    g = torch.Generator(device="cuda").manual_seed(c + r)
    x = torch.randn(b, c, r, r, r, device="cuda", generator=g) * 2.0 + 0.3
    cb = torch.randn(c, device="cuda", generator=g)
    w = torch.randn(c, device="cuda", generator=g)
    bias = torch.randn(c, device="cuda", generator=g)
    want = TF.group_norm(x + cb.view(1, c, 1, 1, 1), 8, w, bias, 1e-5)
    if swish:
        want = want * torch.sigmoid(want)
    peak = want.abs().max().item()
    x_cl = x.permute(0, 2, 3, 4, 1).contiguous()
    got, sums = cuda_backend.groupnorm_act_cl(x_cl, 8, w, bias, 1e-5, swish, conv_bias=cb, channel_sums=True)
    assert (got.permute(0, 4, 1, 2, 3) - want).abs().max().item() / peak <= 1e-5
    want_sums = want.double().flatten(2).sum(-1)
    assert (sums.double() - want_sums).abs().max().item() <= 1e-5 * max(want_sums.abs().max().item(), 1.0)
    got2, tiles = cuda_backend.groupnorm_act_cl(x_cl, 8, w, bias, 1e-5, swish, conv_bias=cb, channel_sums="tiles")
    assert torch.equal(got2, got) and tiles.dim() == 3 and torch.allclose(tiles.sum(dim=1), sums)
    plain = cuda_backend.groupnorm_act_cl(x_cl, 8, None, None, 1e-5, swish)
    want2 = TF.group_norm(x, 8, None, None, 1e-5)
    if swish:
        want2 = want2 * torch.sigmoid(want2)
    assert (plain.permute(0, 4, 1, 2, 3) - want2).abs().max().item() / want2.abs().max().item() <= 1e-5


@pytest.mark.parametrize("b,t,scale", [(2, 512, 1.0), (3, 4096, 1.0), (1, 1024, 3.0), (2, 128, 0.2)])
def test_attention_from_fused_projection_vs_float64(b, t, scale, cuda_backend):
    """bdm_attention_qkv: q | k | v of a token side by side (one projection GEMM), biases added as the kernel reads,
    token-major output -- as close to float64 as torch's fp32 route, like the channel-first entry point"""
    import torch
    g = torch.Generator(device="cuda").manual_seed(t + b)
    qkv = torch.randn(b, t, 192, device="cuda", generator=g) * scale
    bias = torch.randn(192, device="cuda", generator=g) * 0.3
    got = cuda_backend.attention_qkv(qkv, bias)                                   # [B,T,64]
    q, k, v = ((qkv + bias)[..., i * 64:(i + 1) * 64] for i in range(3))           # [B,T,64] each
    ref = torch.matmul(torch.softmax(torch.matmul(q.double(), k.double().transpose(1, 2)), -1), v.double())
    f32 = torch.matmul(torch.softmax(torch.matmul(q, k.transpose(1, 2)), -1), v)
    peak = ref.abs().max().item()
    e_ours = (got.double() - ref).abs().max().item() / peak
    e_torch = (f32.double() - ref).abs().max().item() / peak
    assert e_ours <= max(3 * e_torch, 1e-5), (e_ours, e_torch)
    assert torch.equal(cuda_backend.attention_qkv(qkv, None), cuda_backend.attention_qkv(qkv, torch.zeros_like(bias)))


def test_attention_block_on_channels_last_grid(cuda_backend):
    """The attention block of the R=16 PVConv stage on a channels-last grid: fused route (one projection GEMM, the
    tcgen05 attention kernel, residual GEMM, norm kernel with the out-conv bias folded in and the SE sums) against
    the module-by-module torch route; and inside a voxel stack, where the SE gate takes the block's sums."""
    import torch

    import bdm_b200.modules.layers as L
    torch.manual_seed(5)
    blk = L.Attention(64, 8).cuda().eval()
    x = torch.randn(4, 64, 16, 16, 16, device="cuda").contiguous(memory_format=torch.channels_last_3d)
    saved = (L.FUSED_ATTENTION, torch.backends.cudnn.allow_tf32)
    try:
        torch.backends.cudnn.allow_tf32 = False            # fp32 projections on both sides
        with torch.no_grad():
            L.FUSED_ATTENTION = True
            assert blk.fused_applicable(x)
            y1 = blk(x)
            y1s, sums = blk.forward_fused(x, channel_sums=True)
            L.FUSED_ATTENTION = False
            y0 = blk(x)
    finally:
        L.FUSED_ATTENTION, torch.backends.cudnn.allow_tf32 = saved
    assert y1.shape == y0.shape and L.is_channels_last_3d(y1)
    assert (y1 - y0).abs().max().item() <= 1e-5 * y0.abs().max().item()
    assert torch.equal(y1s, y1)
    want_sums = y0.double().flatten(2).sum(-1)
    assert (sums.sum(dim=1).double() - want_sums).abs().max().item() <= 1e-5 * want_sums.abs().max().item()


@pytest.mark.parametrize("cin,widths,m,u,n", [(32, (32, 64), 256, 32, 1024), (64, (64, 128), 64, 16, 256), (128, (128,), 16, 8, 64)])
def test_grouped_channels_padded_to_a_multiple_of_four(cin, widths, m, u, n, cuda_backend):
    """SA module with the grouped tensor's 3 + C channels zero-padded to a multiple of 4 (aligned GEMM against a
    zero-padded weight) == the unpadded route; fp32 GEMMs on both sides: 1e-5 of the output's peak"""
    import torch

    import bdm_b200.modules.layers as L
    from bdm_b200.modules.pointnet2 import PointNetSAModule
    torch.manual_seed(cin + m)
    sa = PointNetSAModule(m, 0.4, u, cin, list(widths)).cuda().eval()
    feats = torch.randn(3, cin, n, device="cuda")
    coords = torch.rand(3, 3, n, device="cuda")
    temb = torch.randn(3, 8, n, device="cuda")
    saved = (L.PAD_GROUPED_CHANNELS, torch.backends.cudnn.allow_tf32)
    try:
        torch.backends.cudnn.allow_tf32 = False
        with torch.no_grad():
            L.PAD_GROUPED_CHANNELS = True
            assert L.pads_grouped_channels(feats, sa.mlps[0].layers[0])
            y_pad = sa((feats, coords, temb))[0]
            L.PAD_GROUPED_CHANNELS = False
            y_ref = sa((feats, coords, temb))[0]
    finally:
        L.PAD_GROUPED_CHANNELS, torch.backends.cudnn.allow_tf32 = saved
    err = (y_pad - y_ref).abs().max().item() / y_ref.abs().max().item()
    assert err <= 1e-5, err
