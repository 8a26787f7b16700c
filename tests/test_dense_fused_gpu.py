"""GPU: the fused GroupNorm(+Swish) kernel against a plain PyTorch fp32 reference of the same op
(torch.nn.functional.group_norm followed by x*sigmoid(x)).  Tolerance: 1e-5 relative to the output's peak."""
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [(16, 32, (32, 32, 32)), (16, 64, (1024, 32)), (4, 256, (8, 8, 8)), (2, 512, (16,)), (3, 16, (5, 7)),
          (2, 8, (1,)), (1, 128, (4096,)), (16, 64, (4096,)), (2, 24, (33,))]


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
        saved = L.FUSED_NORM_ACT
        try:
            L.FUSED_NORM_ACT = True
            y_fused = net(x, t)
            L.FUSED_NORM_ACT = False
            y_plain = net(x, t)
        finally:
            L.FUSED_NORM_ACT = saved
    err = (y_fused - y_plain).abs().max().item() / y_plain.abs().max().item()
    assert err <= 1e-4, err   # 60+ norm layers deep; each within 1e-5
