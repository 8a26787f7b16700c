"""CPU tests of the host-side mirror: module tree / state_dict parity with the reference, drop-in
injection, and forward equality of our modules vs the reference's own Python modules when both run
on the same (oracle-backed) `_backend`.  The reference-dependent tests need /root/reference and are
skipped where it does not exist (the GPU box)."""
import os
import sys

import numpy as np
import pytest
import torch

REF_EXP = "/root/reference/experiments"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF_EXP), reason="reference checkout not present")


@pytest.fixture()
def oracle_backend(monkeypatch):
    from oracle.torch_backend import OracleBackend
    import bdm_b200.functional.ops as ops
    ob = OracleBackend()
    monkeypatch.setattr(ops, "_B", ob)
    return ob


def test_public_surface():
    import bdm_b200.functional as F
    import bdm_b200.modules as M
    for name in ['ball_query', 'trilinear_devoxelize', 'grouping', 'nearest_neighbor_interpolate', 'kl_loss',
                 'huber_loss', 'gather', 'furthest_point_sample', 'logits_mask', 'avg_voxelize']:
        assert callable(getattr(F, name)), name          # functional/__init__.py:1-7
    for name in ['BallQuery', 'FrustumPointNetLoss', 'KLLoss', 'PointNetAModule', 'PointNetSAModule',
                 'PointNetFPModule', 'PVConv', 'Attention', 'Swish', 'PVConvReLU', 'SE3d', 'SharedMLP',
                 'Voxelization']:
        assert hasattr(M, name), name                    # modules/__init__.py:1-8
    from bdm_b200 import backend
    for name in ['gather_features_forward', 'gather_features_backward', 'furthest_point_sampling', 'ball_query',
                 'grouping_forward', 'grouping_backward', 'three_nearest_neighbors_interpolate_forward',
                 'three_nearest_neighbors_interpolate_backward', 'trilinear_devoxelize_forward',
                 'trilinear_devoxelize_backward', 'avg_voxelize_forward', 'avg_voxelize_backward']:
        assert callable(getattr(backend, name)), name    # bindings.cpp:11-36


def test_parameter_counts():
    from bdm_b200.denoiser import PVCNN2_PC2, PVCNN2_PVD
    pc2 = PVCNN2_PC2(num_classes=3, embed_dim=64, extra_feature_channels=387)
    pvd = PVCNN2_PVD(3, 64, True, 0.1, extra_feature_channels=0)
    assert sum(p.numel() for p in pc2.parameters()) == 28046275   # SURVEY.md section 3.1 (probe P3)
    assert sum(p.numel() for p in pvd.parameters()) == 27649987


def test_forward_counts_match_schedule():
    from bdm_b200.diffusion import forward_counts
    assert forward_counts(mode="blending") == dict(pc2=1000, pvd=80, fuse=0)
    assert forward_counts(mode="merging") == dict(pc2=995, pvd=75, fuse=5)


def test_backend_rejects_cpu_tensors():
    """No CPU fallback: the product backend refuses non-CUDA inputs like the reference (utils.hpp:7)."""
    from bdm_b200 import backend
    with pytest.raises(RuntimeError):
        backend.grouping_forward(torch.zeros(1, 2, 8), torch.zeros(1, 2, 2, dtype=torch.int32))
    with pytest.raises(RuntimeError):
        backend.furthest_point_sampling(torch.zeros(1, 3, 8), 4)


def _small_inputs(b=2, n=96, extra=5, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(b, 3 + extra, n, generator=g), torch.tensor([500.0, 17.0])[:b]


def test_denoiser_runs_on_oracle_backend(oracle_backend):
    """Our module tree end to end on CPU (oracle-backed): shapes, op-call census of SURVEY.md section 3.1."""
    from bdm_b200.denoiser import PVCNN2_PC2
    torch.manual_seed(1)
    net = PVCNN2_PC2(num_classes=3, embed_dim=64, extra_feature_channels=5).eval()
    x, t = _small_inputs()
    with torch.no_grad():
        y = net(x, t)
    assert y.shape == (2, 3, 96) and torch.isfinite(y).all()
    names = [c[0] for c in oracle_backend.calls]
    assert names.count("avg_voxelize_forward") == 14
    assert names.count("trilinear_devoxelize_forward") == 14
    assert names.count("furthest_point_sampling") == 4
    assert names.count("ball_query") == 4
    # the reference does 12 groupings and 8 three-NN calls per forward; the mirror drops the 4 groupings
    # of the point-constant time embedding, searches once per FP stage and interpolates once per FP stage (the
    # interpolated time embedding is the tail of the interpolated [features, temb])
    assert names.count("grouping_forward") == 8
    assert names.count("three_nn_search") == 4 and names.count("three_nn_interpolate") == 4


@needs_ref
def test_state_dict_and_forward_match_reference_modules(oracle_backend):
    """Reference PVCNN2_PC2 (its own Python, via bdm_b200.dropin with an oracle-backed _backend) vs
    our PVCNN2_PC2 with the same weights: identical state_dict layout, bit-identical forward."""
    from bdm_b200 import dropin
    from bdm_b200.denoiser import PVCNN2_PC2
    saved = {k: sys.modules.get(k) for k in ("model", "pvd") + dropin.BACKEND_MODULE_NAMES}
    try:
        dropin.install(REF_EXP, backend=oracle_backend, stub_packages=True)
        from model.pvcnn.pvcnn import PVCNN2_PC2 as RefNet
        torch.manual_seed(3)
        ref = RefNet(num_classes=3, embed_dim=64, extra_feature_channels=5).eval()
        ours = PVCNN2_PC2(num_classes=3, embed_dim=64, extra_feature_channels=5).eval()
        rs, os_ = ref.state_dict(), ours.state_dict()
        assert list(rs.keys()) == list(os_.keys())
        assert all(rs[k].shape == os_[k].shape for k in rs)
        ours.load_state_dict(rs)
        x, t = _small_inputs(seed=4)
        with torch.no_grad():
            yr = ref(x, t)
            yo = ours(x, t)
        assert torch.equal(yr, yo)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in [k for k in sys.modules if k.startswith("model.") or k.startswith("pvd.")]:
            sys.modules.pop(k, None)


@needs_ref
def test_pvd_state_dict_matches_reference():
    from bdm_b200 import dropin
    from bdm_b200.denoiser import PVCNN2_PVD
    from oracle.torch_backend import OracleBackend
    saved = {k: sys.modules.get(k) for k in ("model", "pvd") + dropin.BACKEND_MODULE_NAMES}
    try:
        dropin.install(REF_EXP, backend=OracleBackend(), stub_packages=True)
        from pvd.model.pvcnn_generation import PVCNN2Base_PVD

        class RefPVD(PVCNN2Base_PVD):
            from bdm_b200.denoiser import FP_BLOCKS as fp_blocks, SA_BLOCKS as sa_blocks
        ref = RefPVD(num_classes=3, embed_dim=64, use_att=True, dropout=0.1, extra_feature_channels=0)
        ours = PVCNN2_PVD(3, 64, True, 0.1, extra_feature_channels=0)
        assert [(k, tuple(v.shape)) for k, v in ref.state_dict().items()] == \
               [(k, tuple(v.shape)) for k, v in ours.state_dict().items()]
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in [k for k in sys.modules if k.startswith("model.") or k.startswith("pvd.")]:
            sys.modules.pop(k, None)


def test_ddpm_schedule_properties():
    from bdm_b200.diffusion import DDPMSchedule, PVDSchedule
    s = DDPMSchedule()
    assert s.timesteps[0] == 999 and s.timesteps[-1] == 0
    sqrt_a, sqrt_b, c0, ct, sig = s.coefficients(0)
    assert c0 == pytest.approx(1.0) and ct == pytest.approx(0.0) and sig == 0.0  # last step returns x0
    x = torch.randn(2, 8, 3)
    assert torch.allclose(s.step(torch.zeros_like(x), 0, x), x / sqrt_a)
    p = PVDSchedule()
    assert torch.allclose(p.step(torch.zeros(2, 3, 8), 0, x.transpose(1, 2)),
                          float(p.coef1[0]) * float(p.sqrt_recip_ac[0]) * x.transpose(1, 2)
                          + float(p.coef2[0]) * x.transpose(1, 2))


def test_modules_run_under_inference_mode(oracle_backend):
    """Activation tensors made under torch.inference_mode() carry no version counter; the plan / geometry
    memos must not ask for one (the reference's modules run under inference_mode)."""
    from bdm_b200.denoiser import PVCNN2_PC2
    from bdm_b200.modules import PVConv
    torch.manual_seed(2)
    block = PVConv(8, 16, kernel_size=3, resolution=8, with_se=True).eval()
    net = PVCNN2_PC2(num_classes=3, embed_dim=64, extra_feature_channels=5).eval()
    x, t = _small_inputs(seed=6)
    with torch.no_grad():
        want_block = block((x[:, :8].contiguous(), x[:, :3].contiguous(), None))[0]
        want = net(x, t)
    with torch.inference_mode():
        feats, coords = x[:, :8].contiguous(), x[:, :3].contiguous()
        got_block = block((feats, coords, None))[0]
        got_block2 = block((feats, coords, None))[0]          # second call hits the one-entry plan memo
        got = net(x.clone(), t)
    assert torch.equal(got_block, want_block) and torch.equal(got_block2, want_block)
    assert torch.equal(got, want)


def test_sampler_drops_graphs_on_reassignment():
    from bdm_b200.diffusion import BDMSampler
    s = BDMSampler(object(), object())
    s._graphs["x"] = 1
    s.cond = s.cond
    assert s._graphs
    s.cond = object()
    assert not s._graphs
    s._graphs["x"] = 1
    s.pc2_net = object()
    assert not s._graphs


def test_concat_conv_widens_a_misaligned_channel_slice():
    """layers.conv_no_bias_concat: a part marked as the channel slice base[:, c0:] of a wider tensor whose width is not a
    multiple of 4 is read as base[:, c0 - e:] against e zero weight columns (one aligned GEMM); same sums as the
    convolution over the concatenation.  Unmarked parts, slices without room below them and non-contiguous bases keep
    the head + tail split."""
    import torch
    import torch.nn as nn

    from bdm_b200.modules import layers as L
    torch.manual_seed(0)
    conv = nn.Conv1d(64 + 387, 128, 1)
    full = torch.randn(2, 390, 32)
    up = torch.randn(2, 64, 32)
    with torch.no_grad():
        want = conv._conv_forward(torch.cat([up, full[:, 3:, :]], 1), conv.weight, None)
        plain = L.conv_no_bias_concat(conv, [up, full[:, 3:, :]])
        assert conv._concat_weight[0][-1] == (0, 0)
        marked = full[:, 3:, :]
        marked._bdm_slice_of = (full, 3)
        wide = L.conv_no_bias_concat(conv, [up, marked])
        assert conv._concat_weight[0][-2:] == ((64, 388), (0, 1))
        assert [(off, width) for off, width, _ in conv._concat_weight[1]] == [(0, 64), (64, 388)]
        # no room below the slice / a base that is not plain-contiguous: not widened
        conv2 = nn.Conv1d(387, 16, 1)
        low = full[:, :387, :]
        low._bdm_slice_of = (full, 0)
        L.conv_no_bias_concat(conv2, [low])
        assert conv2._concat_weight[0][-1] == (0,)
        strided = full.transpose(1, 2).contiguous().transpose(1, 2)       # [2, 390, 32] with channel stride 1
        part = strided[:, 3:, :]
        part._bdm_slice_of = (strided, 3)
        L.conv_no_bias_concat(conv2, [part.contiguous()])                # (the unmarked copy: plain route)
        assert conv2._concat_weight[0][-1] == (0,)
    scale = want.abs().max().item()
    assert (plain - want).abs().max().item() <= 1e-5 * scale
    assert (wide - want).abs().max().item() <= 1e-5 * scale


def test_fp_stages_reuse_the_interpolated_embedding(oracle_backend, monkeypatch):
    """denoiser._decode marks cat([features, temb]); PointNetFPModule then takes the interpolated time embedding from the
    tail of the interpolated concatenation instead of interpolating temb a second time: same bits, half the calls."""
    import bdm_b200.functional as F
    from bdm_b200 import denoiser
    from bdm_b200.denoiser import PVCNN2_PC2
    torch.manual_seed(3)
    net = PVCNN2_PC2(num_classes=3, embed_dim=64, extra_feature_channels=5).eval()
    x, t = _small_inputs(seed=7)
    calls = []
    original = F.three_nn_interpolate

    def counting(feats, idx, w):
        calls.append(int(feats.shape[1]))
        return original(feats, idx, w)

    monkeypatch.setattr(F, "three_nn_interpolate", counting)
    with torch.no_grad():
        got = net(x, t)
    marked_calls = list(calls)
    calls.clear()

    def unmarked_decode(fp_layers, features, coords, temb, coords_per_stage, skips_per_stage):
        for fp_idx, stage in enumerate(fp_layers):
            features, coords, temb = stage((coords_per_stage[-1 - fp_idx], coords, torch.cat([features, temb], dim=1),
                                            skips_per_stage[-1 - fp_idx], temb))
        return features

    monkeypatch.setattr(denoiser, "_decode", unmarked_decode)
    with torch.no_grad():
        want = net(x, t)
    assert torch.equal(got, want)
    assert len(marked_calls) == len(net.fp_layers) and len(calls) == 2 * len(net.fp_layers)
