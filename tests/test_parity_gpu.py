"""GPU parity tests: the CUDA path (bdm_b200.backend -> libbdm_b200.so C-ABI) against
  (1) the CPU oracle (oracle/bdm_oracle.c) on the seeded cases of tests/cases.py,
  (2) the reference's own CUDA kernels (oracle/_ref, when the prebuilt .so travelled with the repo),
  (3) the committed golden vectors tests/golden/ref_*.npz (reference outputs recorded on a B200).
Integer outputs must be bit-exact; floats within 1e-5 relative (1e-4 for atomic-order sums)."""
import glob
import os

import numpy as np
import pytest

from . import cases, runners

pytestmark = pytest.mark.gpu
GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", list(cases.GOLDEN_CASES))
def test_cuda_vs_oracle(name, cuda_backend):
    inp = cases.build_case(name)
    runners.compare(name, runners.run_backend(name, cuda_backend, inp), runners.run_oracle(name, inp))


@pytest.mark.parametrize("name", list(cases.GOLDEN_CASES))
def test_cuda_vs_reference_kernels(name, cuda_backend, ref_backend):
    inp = cases.build_case(name)
    runners.compare(name, runners.run_backend(name, cuda_backend, inp), runners.run_backend(name, ref_backend, inp))


@pytest.mark.parametrize("name", list(cases.GOLDEN_CASES))
def test_cuda_vs_golden(name, cuda_backend):
    path = os.path.join(GOLDEN_DIR, f"ref_{name}.npz")
    if not os.path.exists(path):
        pytest.skip("golden vector not generated yet")
    z = np.load(path)
    inp = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    want = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    runners.compare(name, runners.run_backend(name, cuda_backend, inp), want)


# ---- full-size chain (BASELINE config 2 shapes: N=4096 -> 1024 -> 256 -> 64 -> 16) -------------------
SA = [(1024, 0.1), (256, 0.2), (64, 0.4), (16, 0.8)]


def _chain(be, co_t, torch):
    """FPS -> gather -> ball query -> grouping at the four SA levels; 3-NN back up."""
    outs = {}
    levels = [co_t]
    for li, (m, rad) in enumerate(SA):
        pts = levels[-1]
        idx = be.furthest_point_sampling(pts, m)
        cen = be.gather_features_forward(pts, idx)
        nb = be.ball_query(cen, pts, rad, 32)
        grp = be.grouping_forward(pts, nb)
        outs[f"fps{li}"] = idx
        outs[f"bq{li}"] = nb
        outs[f"grp{li}"] = grp
        levels.append(cen)
    for li in range(len(SA) - 1, -1, -1):
        pts, cen = levels[li], levels[li + 1]
        feat = torch.sin(cen.repeat(1, 3, 1) * 3.0).contiguous()
        o, i3, w3 = be.three_nearest_neighbors_interpolate_forward(pts, cen, feat)
        outs[f"nn_idx{li}"] = i3
        outs[f"nn_w{li}"] = w3
        outs[f"nn_out{li}"] = o
    return outs


@pytest.mark.parametrize("regime", ["noise", "shape"])
def test_full_size_chain_vs_reference_kernels(regime, cuda_backend, ref_backend):
    import torch
    rng = np.random.default_rng(1234)
    co = torch.as_tensor(cases.cloud(rng, 16, 4096, regime)).cuda()
    a = _chain(cuda_backend, co, torch)
    b = _chain(ref_backend, co, torch)
    for k in a:
        if a[k].dtype == torch.int32:
            assert torch.equal(a[k], b[k]), f"{regime}.{k}: integer output differs from the reference kernels"
        else:
            err = (a[k] - b[k]).abs().max().item() / max(b[k].abs().max().item(), 1e-30)
            assert err <= 1e-5, f"{regime}.{k}: rel err {err:.2e}"


@pytest.mark.parametrize("regime", ["noise", "shape"])
def test_full_size_chain_vs_oracle(regime, cuda_backend):
    import torch

    import oracle as O
    rng = np.random.default_rng(4321)
    co_np = cases.cloud(rng, 2, 4096, regime)
    a = _chain(cuda_backend, torch.as_tensor(co_np).cuda(), torch)
    pts = co_np
    for li, (m, rad) in enumerate(SA):
        idx = O.furthest_point_sampling(pts, m)
        assert np.array_equal(a[f"fps{li}"].cpu().numpy(), idx), f"fps level {li}"
        cen = O.gather_features_forward(pts, idx)
        nb = O.ball_query(cen, pts, rad, 32)
        assert np.array_equal(a[f"bq{li}"].cpu().numpy(), nb), f"ball query level {li}"
        assert np.array_equal(a[f"grp{li}"].cpu().numpy(), O.grouping_forward(pts, nb)), f"grouping level {li}"
        pts = cen


VOX_SHAPES = [(390, 4096, 32), (32, 4096, 32), (128, 1024, 16), (192, 256, 8), (256, 64, 8), (64, 4096, 32)]


@pytest.mark.parametrize("c,n,r", VOX_SHAPES)
def test_voxelize_devoxelize_full_size(c, n, r, cuda_backend, ref_backend):
    """BASELINE config-2 voxel shapes at B=16 against the reference kernels + size-independent
    properties (mass conservation, partition of unity)."""
    import torch
    rng = np.random.default_rng(c * 7 + n + r)
    b = 16
    co = cases.cloud(rng, b, n, "shape")
    vox, nc = cases.vox_coords(co, r)
    feat = torch.as_tensor(rng.standard_normal((b, c, n)).astype(np.float32)).cuda()
    vox_t, nc_t = torch.as_tensor(vox).cuda(), torch.as_tensor(nc).cuda()
    out, ind, cnt = cuda_backend.avg_voxelize_forward(feat, vox_t, r)
    ro, ri, rc = ref_backend.avg_voxelize_forward(feat, vox_t, r)
    assert torch.equal(ind, ri) and torch.equal(cnt, rc)
    err = (out - ro).abs().max().item() / ro.abs().max().item()
    assert err <= 1e-4, f"voxelize rel err {err:.2e}"
    # properties: counts partition the cloud; sum_v out*cnt == sum_i feat (mass conservation)
    assert int(cnt.sum().item()) == b * n
    mass = (out.double() * cnt.unsqueeze(1).double()).sum(-1)
    assert torch.allclose(mass, feat.double().sum(-1), rtol=1e-4, atol=1e-3)
    # devoxelize the reference grid with both implementations
    dv = cuda_backend.trilinear_devoxelize_forward(r, False, nc_t, ro)[0]
    rv = ref_backend.trilinear_devoxelize_forward(r, False, nc_t, ro)[0]
    err = (dv - rv).abs().max().item() / max(rv.abs().max().item(), 1e-30)
    assert err <= 1e-5, f"devoxelize rel err {err:.2e}"
    ones = torch.ones((b, 2, r ** 3), device="cuda")
    pu = cuda_backend.trilinear_devoxelize_forward(r, False, nc_t, ones)[0]
    assert (pu - 1.0).abs().max().item() <= 1e-5  # trilinear weights sum to one


def test_voxelize_deterministic(cuda_backend):
    import torch
    rng = np.random.default_rng(5)
    co = cases.cloud(rng, 4, 4096, "noise")
    vox, _ = cases.vox_coords(co, 32)
    feat = torch.as_tensor(rng.standard_normal((4, 16, 4096)).astype(np.float32)).cuda()
    vt = torch.as_tensor(vox).cuda()
    a = cuda_backend.avg_voxelize_forward(feat, vt, 32)[0]
    for _ in range(3):
        assert torch.equal(a, cuda_backend.avg_voxelize_forward(feat, vt, 32)[0])


@pytest.mark.parametrize("n,m", [(33, 20), (700, 90), (1024, 256), (1025, 64), (2500, 300), (4096, 1024), (6000, 128), (8192, 64)])
def test_fps_every_launch_configuration(n, m, cuda_backend):
    """each register-path instantiation (4/16/8 points per thread) against the oracle, with duplicates"""
    import torch

    import oracle as O
    rng = np.random.default_rng(n * 31 + m)
    co = cases.cloud(rng, 2, n, "shape")
    co[:, :, rng.integers(0, n, n // 5)] = co[:, :, rng.integers(0, n, n // 5)]
    idx = cuda_backend.furthest_point_sampling(torch.as_tensor(co).cuda(), m)
    assert np.array_equal(idx.cpu().numpy(), O.furthest_point_sampling(co, m))


@pytest.mark.parametrize("c,n,r", [(5, 777, 8), (7, 1000, 16), (64, 4096, 32), (3, 300, 5), (9, 5000, 32), (2, 100, 1)])
@pytest.mark.parametrize("b", [1, 16, 60])
def test_devoxelize_fast_path_vs_oracle(b, c, n, r, cuda_backend):
    """slice-ring kernel (CT = 1, 2, 4 depending on b*c) and odd sizes, bit-exact against the oracle"""
    import torch

    import oracle as O
    if b * c * r ** 3 > 40e6:
        b = 4
    rng = np.random.default_rng(b + c + n + r)
    nc = (rng.random((b, 3, n)) * (r - 1)).astype(np.float32)
    nc[:, :, : min(n, 16)] = np.round(nc[:, :, : min(n, 16)])
    nc[0, :, -1] = r - 1
    grid = rng.standard_normal((b, c, r ** 3)).astype(np.float32)
    got = cuda_backend.trilinear_devoxelize_forward(r, False, torch.as_tensor(nc).cuda(), torch.as_tensor(grid).cuda())[0]
    want = O.trilinear_devoxelize_forward(r, False, nc, grid)[0]
    assert np.array_equal(got.cpu().numpy(), want)


def test_voxel_plan_reuse(cuda_backend):
    """one plan, several feature tensors == independent avg_voxelize calls"""
    import torch
    rng = np.random.default_rng(3)
    co = cases.cloud(rng, 4, 2048, "shape")
    vox, _ = cases.vox_coords(co, 16)
    vt = torch.as_tensor(vox).cuda()
    plan = cuda_backend.voxel_plan(vt, 16)
    for c in (1, 3, 8, 33):
        f = torch.randn(4, c, 2048, device="cuda")
        assert torch.equal(cuda_backend.avg_voxelize_fill(f, plan), cuda_backend.avg_voxelize_forward(f, vt, 16)[0])


def test_generic_paths(cuda_backend):
    """sizes outside the shared-memory fast paths: R^3 > 32768 (voxelize), N > 8192 (FPS)"""
    import torch

    import oracle as O
    rng = np.random.default_rng(11)
    co = cases.cloud(rng, 1, 3000, "shape")
    vox, nc = cases.vox_coords(co, 40)
    feat = rng.standard_normal((1, 3, 3000)).astype(np.float32)
    out, ind, cnt = cuda_backend.avg_voxelize_forward(torch.as_tensor(feat).cuda(), torch.as_tensor(vox).cuda(), 40)
    oo, oi, oc = O.avg_voxelize_forward(feat, vox, 40)
    assert np.array_equal(ind.cpu().numpy(), oi) and np.array_equal(cnt.cpu().numpy(), oc)
    assert np.abs(out.cpu().numpy() - oo).max() <= 1e-4 * np.abs(oo).max()
    co2 = cases.cloud(rng, 1, 9000, "noise")
    idx = cuda_backend.furthest_point_sampling(torch.as_tensor(co2).cuda(), 64)
    assert np.array_equal(idx.cpu().numpy(), O.furthest_point_sampling(co2, 64))


def test_input_checks_raise(cuda_backend):
    import torch
    f = torch.zeros((1, 2, 8), device="cuda")
    with pytest.raises(RuntimeError):
        cuda_backend.grouping_forward(f.cpu(), torch.zeros((1, 2, 2), dtype=torch.int32, device="cuda"))
    with pytest.raises(RuntimeError):
        cuda_backend.grouping_forward(f, torch.zeros((1, 2, 2), dtype=torch.int64, device="cuda"))
    with pytest.raises(RuntimeError):
        cuda_backend.ball_query(f.transpose(1, 2), f, 0.1, 4)


@pytest.mark.parametrize("b,c,n,m,u", [(2, 5, 300, 40, 8), (3, 64, 4096, 1024, 32), (1, 7, 50, 9, 3), (2, 32, 1024, 256, 32)])
def test_grouping_into_concatenated_tensor(b, c, n, m, u, cuda_backend):
    """BallQuery's grouping -> subtract centres -> cat, written in place: bit-exact against the oracle's
    grouping followed by the same fp32 subtraction"""
    import torch

    import oracle
    rng = np.random.default_rng(b * 7 + u)
    pts = rng.normal(size=(b, 3, n)).astype(np.float32)
    feats = rng.normal(size=(b, c, n)).astype(np.float32)
    cen = rng.normal(size=(b, 3, m)).astype(np.float32)
    idx = rng.integers(0, n, size=(b, m, u)).astype(np.int32)
    want = np.concatenate([oracle.grouping_forward(pts, idx) - cen[..., None], oracle.grouping_forward(feats, idx)], axis=1)
    out = torch.full((b, 3 + c, m, u), float("nan"), device="cuda")
    ti = torch.from_numpy(idx).cuda()
    cuda_backend.grouping_into(torch.from_numpy(pts).cuda(), ti, out, 0, centers=torch.from_numpy(cen).cuda())
    cuda_backend.grouping_into(torch.from_numpy(feats).cuda(), ti, out, 3)
    assert np.array_equal(out.cpu().numpy(), want)
