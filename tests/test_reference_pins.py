"""CPU: pins taken from the reference's own Python sources (run here, where /root/reference exists; skipped on
the GPU box).  The functions are executed from the reference files themselves -- cut out by name with `ast`,
because the files' module-level imports (pytorch3d, open3d, hydra, ipdb ...) are not installable here.

  * evaluation/evaluation_f1.py:90-110  compute_pc_to_pc_dist / cal_fscore   -> oracle.nn_expanded / oracle.fscore
  * pvd/__init__.py:18-224              GaussianDiffusion (coefficients, p_sample) -> oracle.pvd_* and PVDSchedule
  * model/pvcnn/pvcnn_fuse.py:14-237    PVCNN_fuse state_dict layout and forward   -> bdm_b200.denoiser.PVCNNFuse
"""
import ast
import os
import sys
import types

import numpy as np
import pytest
import torch

REF_EXP = "/root/reference/experiments"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_EXP), reason="reference checkout not present")


def _exec_definitions(path, names, namespace):
    """exec the top-level functions / classes `names` of a reference file, and nothing else of it"""
    src = open(path).read()
    tree = ast.parse(src)
    found = []
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), namespace)
            found.append(node.name)
    assert sorted(found) == sorted(names), f"{path}: expected {names}, found {found}"
    return namespace


# ---------------------------------------------------------------------------------------------------
# F-score
# ---------------------------------------------------------------------------------------------------
def _eval_pairs(b=3, n=512, seed=2003):
    from tests.cases import cloud
    rng = np.random.default_rng(seed)
    gt = cloud(rng, b, n, "shape").transpose(0, 2, 1).astype(np.float64)
    pred = gt[:, rng.permutation(n)] + 0.05 * rng.standard_normal(gt.shape)
    gt -= gt.mean(1, keepdims=True)
    pred -= pred.mean(1, keepdims=True)
    return gt, pred


def test_fscore_matches_reference_source():
    import oracle as O
    ns = _exec_definitions(os.path.join(REF_EXP, "evaluation", "evaluation_f1.py"),
                           ["compute_pc_to_pc_dist", "cal_fscore"], {"torch": torch})
    gt, pred = _eval_pairs()
    d_or = O.nn_expanded(gt, pred)
    f_or = O.fscore(gt, pred)
    for i in range(gt.shape[0]):
        g, p = torch.from_numpy(gt[i]), torch.from_numpy(pred[i])       # float64, like evaluation_f1.py:129-141
        d_ref = np.asarray(ns["compute_pc_to_pc_dist"](g, p))
        # the reference's distances come out of a BLAS matmul whose summation order is the library's;
        # same formula, same clamp: agreement to the last few bits, and never across the 0.01 threshold
        assert np.allclose(d_or[i], d_ref, rtol=0, atol=4e-15)
        assert np.array_equal(d_or[i] < 0.01, d_ref < 0.01)
        assert f_or[i] == ns["cal_fscore"](g, p)                           # the metric itself: exactly equal
    # distances straddle the threshold (the comparison above is not vacuous)
    assert 0.05 < (d_or < 0.01).mean() < 0.95


# ---------------------------------------------------------------------------------------------------
# PVD GaussianDiffusion
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ref_diffusion():
    ns = _exec_definitions(os.path.join(REF_EXP, "pvd", "__init__.py"), ["GaussianDiffusion"],
                           {"torch": torch, "np": np, "tqdm": lambda it, **kw: it})
    betas = np.linspace(0.0001, 0.02, 1000)            # pvd/__init__.py:476-478 get_betas('linear', ...)
    return ns["GaussianDiffusion"](betas, "mse", "eps", "fixedsmall")


def test_pvd_coefficients_match_reference(ref_diffusion):
    import oracle as O
    from bdm_b200.diffusion import PVDSchedule
    c, p, gd = O.pvd_coefficients(), PVDSchedule(), ref_diffusion
    for ours, theirs in ((c["sqrt_recip_ac"], gd.sqrt_recip_alphas_cumprod), (c["sqrt_recipm1_ac"], gd.sqrt_recipm1_alphas_cumprod),
                         (c["coef1"], gd.posterior_mean_coef1), (c["coef2"], gd.posterior_mean_coef2),
                         (c["sigma"], torch.exp(0.5 * gd.posterior_log_variance_clipped))):
        assert np.array_equal(ours, theirs.numpy())
    assert torch.equal(p.sqrt_recip_ac, gd.sqrt_recip_alphas_cumprod) and torch.equal(p.coef1, gd.posterior_mean_coef1)
    assert torch.equal(p.coef2, gd.posterior_mean_coef2) and torch.equal(p.post_log_var, gd.posterior_log_variance_clipped)
    for t in (0, 1, 499, 999):
        assert np.array_equal(p.row(t)[:5], [c["sqrt_recip_ac"][t], c["sqrt_recipm1_ac"][t], c["coef1"][t], c["coef2"][t], c["sigma"][t]])


@pytest.mark.parametrize("t", [999, 872, 500, 17, 1, 0])
def test_pvd_step_matches_reference_p_sample(ref_diffusion, t):
    """GaussianDiffusion.p_sample (pvd/__init__.py:196-224) under a shared noise tensor and a shared noise
    prediction == PVDSchedule.step == oracle.pvd_p_sample, bit for bit."""
    import oracle as O
    from bdm_b200.diffusion import PVDSchedule
    g = torch.Generator().manual_seed(100 + t)
    x = torch.randn(2, 3, 257, generator=g)
    eps = torch.randn(2, 3, 257, generator=g)
    noise = torch.randn(2, 3, 257, generator=g)
    t_vec = torch.full((2,), t, dtype=torch.int64)
    want = ref_diffusion.p_sample(denoise_fn=lambda data, tt: eps, data=x, t=t_vec,
                                  noise_fn=lambda size, dtype, device: noise, clip_denoised=False,
                                  return_pred_xstart=False)
    got = PVDSchedule().step(eps, t, x, noise=noise)
    assert torch.equal(got, want)
    assert np.array_equal(O.pvd_p_sample(x.numpy(), eps.numpy(), noise.numpy(), t), want.numpy())


def test_ddpm_step_product_matches_oracle():
    """diffusers is not vendored (parity unpinned): the product's eager step and the oracle restate the same
    published update independently and must agree bit for bit."""
    import oracle as O
    from bdm_b200.diffusion import DDPMSchedule
    s, rows = DDPMSchedule(), O.ddpm_coefficients()
    g = torch.Generator().manual_seed(5)
    x, eps, noise = (torch.randn(2, 130, 3, generator=g) for _ in range(3))
    for t in (999, 500, 1, 0):
        assert np.array_equal(s.row(t)[:5], rows[t])
        assert np.array_equal(s.step(eps, t, x, noise=noise).numpy(), O.ddpm_step(x.numpy(), eps.numpy(), noise.numpy(), t, rows))
    # t = 0 returns the predicted clean sample, no noise
    assert np.array_equal(O.ddpm_step(x.numpy(), np.zeros_like(x.numpy()), noise.numpy(), 0, rows), (x.numpy() - 0) * rows[0][1] * rows[0][2] + rows[0][3] * x.numpy())


# ---------------------------------------------------------------------------------------------------
# PVCNN_fuse
# ---------------------------------------------------------------------------------------------------
def _with_reference_modules(fn):
    from bdm_b200 import dropin
    from oracle.torch_backend import OracleBackend
    import bdm_b200.functional.ops as ops
    saved = {k: sys.modules.get(k) for k in ("model", "pvd") + dropin.BACKEND_MODULE_NAMES}
    saved_b = ops._B
    ob = OracleBackend()
    try:
        ops._B = ob
        dropin.install(REF_EXP, backend=ob, stub_packages=True)
        return fn()
    finally:
        ops._B = saved_b
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in [k for k in sys.modules if k.startswith("model.") or k.startswith("pvd.")]:
            sys.modules.pop(k, None)


def test_pvcnn_fuse_layout_and_forward_match_reference():
    """Reference PVCNN_fuse (pvcnn_fuse.py:14-237, its own Python over the oracle backend) vs PVCNNFuse with
    the same weights: identical state_dict layout; identical forward at N = 16 points, the only cloud size
    at which the reference's forward stays in bounds (it hands the PVD encoder the PC^2 encoder's leftover
    time embedding of length 16 and gathers it with point indices, pvcnn_fuse.py:176-186)."""
    def run():
        from model.pvcnn.pvcnn import PVCNN2_PC2 as RefPC2
        from model.pvcnn.pvcnn_fuse import PVCNN_fuse as RefFuse
        from pvd.model.pvcnn_generation import PVCNN2Base_PVD
        from bdm_b200.denoiser import FP_BLOCKS, SA_BLOCKS, PVCNN2_PC2, PVCNN2_PVD, PVCNNFuse

        class RefPVD(PVCNN2Base_PVD):
            sa_blocks, fp_blocks = SA_BLOCKS, FP_BLOCKS
        extra = 5
        torch.manual_seed(11)
        ref_pc2 = RefPC2(num_classes=3, embed_dim=64, extra_feature_channels=extra).eval()
        ref_pvd = RefPVD(num_classes=3, embed_dim=64, use_att=True, dropout=0.1, extra_feature_channels=0).eval()
        wrap_pvd = types.SimpleNamespace(model=types.SimpleNamespace(module=ref_pvd))
        wrap_pc2 = types.SimpleNamespace(point_cloud_model=types.SimpleNamespace(model=ref_pc2))
        ref = RefFuse(wrap_pvd, wrap_pc2, num_classes=3, embed_dim=64, extra_feature_channels=extra).eval()
        with torch.no_grad():                      # make the zero-initialised projections matter
            for proj in ref.projs:
                proj[-1].weight.normal_(0, 0.05)
                proj[-1].bias.normal_(0, 0.05)

        pc2 = PVCNN2_PC2(num_classes=3, embed_dim=64, extra_feature_channels=extra).eval()
        pvd = PVCNN2_PVD(3, 64, True, 0.1, extra_feature_channels=0).eval()
        ours = PVCNNFuse(pvd, pc2, extra_feature_channels=extra).eval()
        rs, os_ = ref.state_dict(), ours.state_dict()
        assert [(k, tuple(v.shape)) for k, v in rs.items()] == [(k, tuple(v.shape)) for k, v in os_.items()]
        ours.load_state_dict(rs)

        n = 16
        g = torch.Generator().manual_seed(12)
        recon = torch.randn(2, 3 + extra, n, generator=g)
        prior = torch.randn(2, 3, n, generator=g)
        t = torch.tensor([500.0, 17.0])
        with torch.no_grad():
            want = ref(recon, prior, t)
            got = ours(recon, prior, t)
        assert want.shape == (2, 3, n)
        assert torch.equal(got, want)
        # the PVD branch is live in this comparison
        with torch.no_grad():
            assert not torch.equal(ours(recon, prior + 0.5, t), got)
    _with_reference_modules(run)
