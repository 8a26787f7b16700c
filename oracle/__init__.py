"""oracle -- CPU restatement of the reference's hot-path ops.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this package; nothing under bdm_b200/ does (tests/test_boundary.py greps for it).  The C file
bdm_oracle.c holds the algorithms (each citing the reference file:line it follows); this module is
the numpy/ctypes binding plus the few pieces of reference *Python* glue that sit on the path
(`Voxelization.forward`, Chamfer / F-score reductions).

All functions take and return numpy arrays with the reference's layouts ([B,C,N] channel-first,
[B,3,N] coordinate planes, int32 indices).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "bdm_oracle.c")
_SO = os.path.join(_HERE, "_build", "libbdm_oracle.so")
_lib = None


def build(force=False):
    """gcc -O2 -fopenmp -ffp-contract=off (every fma in the C file is explicit)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-fPIC",
                               "-shared", "-o", _SO, _SRC, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


_I = ctypes.c_int
_F = ctypes.c_float


# ------------------------------------------------------------------------------------------------
# the 7 forward ops (+ 5 backward)
# ------------------------------------------------------------------------------------------------
def avg_voxelize_forward(features, coords, r):
    """vox.cpp:17-43 -> (out[B,C,R^3], ind[B,N], cnt[B,R^3])"""
    features, coords = _f32(features), _i32(coords)
    b, c, n = features.shape
    r3 = r * r * r
    out = np.empty((b, c, r3), np.float32)
    ind = np.empty((b, n), np.int32)
    cnt = np.empty((b, r3), np.int32)
    lib().orc_avg_voxelize_forward(_I(b), _I(c), _I(n), _I(r), _p(coords), _p(features), _p(ind), _p(cnt), _p(out))
    return out, ind, cnt


def avg_voxelize_backward(grad_y, ind, cnt):
    """vox.cpp:54-76"""
    grad_y, ind, cnt = _f32(grad_y), _i32(ind), _i32(cnt)
    b, c, s = grad_y.shape
    n = ind.shape[1]
    gx = np.empty((b, c, n), np.float32)
    lib().orc_avg_voxelize_backward(_I(b), _I(c), _I(n), _I(s), _p(ind), _p(cnt), _p(grad_y), _p(gx))
    return gx


def trilinear_devoxelize_forward(r, is_training, coords, features):
    """trilinear_devox.cpp:18-55 -> (outs[B,C,N], inds, wgts); inds/wgts are [1] zeros when not training"""
    coords, features = _f32(coords), _f32(features)
    b, c = features.shape[:2]
    n = coords.shape[2]
    outs = np.empty((b, c, n), np.float32)
    if is_training:
        inds = np.empty((b, 8, n), np.int32)
        wgts = np.empty((b, 8, n), np.float32)
    else:
        inds = np.zeros((1,), np.int32)
        wgts = np.zeros((1,), np.float32)
    lib().orc_trilinear_devoxelize_forward(_I(b), _I(c), _I(n), _I(r), _I(1 if is_training else 0), _p(coords),
                                           _p(features), _p(inds), _p(wgts), _p(outs))
    return outs, inds, wgts


def trilinear_devoxelize_backward(grad_y, inds, wgts, r):
    """trilinear_devox.cpp:68-94"""
    grad_y, inds, wgts = _f32(grad_y), _i32(inds), _f32(wgts)
    b, c, n = grad_y.shape
    r3 = r * r * r
    gx = np.empty((b, c, r3), np.float32)
    lib().orc_trilinear_devoxelize_backward(_I(b), _I(c), _I(n), _I(r3), _p(inds), _p(wgts), _p(grad_y), _p(gx))
    return gx


def gather_features_forward(features, indices):
    """sampling.cpp:6-23"""
    features, indices = _f32(features), _i32(indices)
    b, c, n = features.shape
    m = indices.shape[1]
    out = np.empty((b, c, m), np.float32)
    lib().orc_gather_features_forward(_I(b), _I(c), _I(n), _I(m), _p(features), _p(indices), _p(out))
    return out


def gather_features_backward(grad_y, indices, n):
    """sampling.cpp:25-41"""
    grad_y, indices = _f32(grad_y), _i32(indices)
    b, c, m = grad_y.shape
    gx = np.empty((b, c, n), np.float32)
    lib().orc_gather_features_backward(_I(b), _I(c), _I(n), _I(m), _p(grad_y), _p(indices), _p(gx))
    return gx


def furthest_point_sampling(coords, m):
    """sampling.cpp:43-58 -> int32[B,M]"""
    coords = _f32(coords)
    b, _, n = coords.shape
    idx = np.zeros((b, max(m, 0)), np.int32)
    lib().orc_furthest_point_sampling(_I(b), _I(n), _I(m), _p(coords), _p(idx))
    return idx


def ball_query(centers, points, radius, u):
    """ball_query.cpp:6-30 -> int32[B,M,U]; r2 = radius*radius in fp32 like the host wrapper (:24)"""
    centers, points = _f32(centers), _f32(points)
    b, _, m = centers.shape
    n = points.shape[2]
    r2 = np.float32(radius) * np.float32(radius)
    out = np.empty((b, m, u), np.int32)
    lib().orc_ball_query(_I(b), _I(n), _I(m), _F(float(r2)), _I(u), _p(centers), _p(points), _p(out))
    return out


def grouping_forward(features, indices):
    """grouping.cpp:6-24"""
    features, indices = _f32(features), _i32(indices)
    b, c, n = features.shape
    _, m, u = indices.shape
    out = np.empty((b, c, m, u), np.float32)
    lib().orc_grouping_forward(_I(b), _I(c), _I(n), _I(m), _I(u), _p(features), _p(indices), _p(out))
    return out


def grouping_backward(grad_y, indices, n):
    """grouping.cpp:26-43"""
    grad_y, indices = _f32(grad_y), _i32(indices)
    b, c, m, u = grad_y.shape
    gx = np.empty((b, c, n), np.float32)
    lib().orc_grouping_backward(_I(b), _I(c), _I(n), _I(m), _I(u), _p(grad_y), _p(indices), _p(gx))
    return gx


def three_nn(points, centers):
    """neighbor_interpolate.cu:20-75 -> (idx int32[B,3,N], w f32[B,3,N])"""
    points, centers = _f32(points), _f32(centers)
    b, _, n = points.shape
    m = centers.shape[2]
    w = np.empty((b, 3, n), np.float32)
    idx = np.empty((b, 3, n), np.int32)
    lib().orc_three_nn(_I(b), _I(n), _I(m), _p(points), _p(centers), _p(w), _p(idx))
    return idx, w


def three_interpolate(features, idx, w):
    """neighbor_interpolate.cu:90-116"""
    features, idx, w = _f32(features), _i32(idx), _f32(w)
    b, c, m = features.shape
    n = idx.shape[2]
    out = np.empty((b, c, n), np.float32)
    lib().orc_three_interpolate(_I(b), _I(c), _I(m), _I(n), _p(features), _p(idx), _p(w), _p(out))
    return out


def three_nearest_neighbors_interpolate_forward(points, centers, features):
    """neighbor_interpolate.cpp:6-40 -> (out[B,C,N], idx[B,3,N], w[B,3,N])"""
    idx, w = three_nn(points, centers)
    return three_interpolate(features, idx, w), idx, w


def three_nearest_neighbors_interpolate_backward(grad_y, idx, w, m):
    """neighbor_interpolate.cpp:42-66"""
    grad_y, idx, w = _f32(grad_y), _i32(idx), _f32(w)
    b, c, n = grad_y.shape
    gx = np.empty((b, c, m), np.float32)
    lib().orc_three_interpolate_backward(_I(b), _I(c), _I(n), _I(m), _p(grad_y), _p(idx), _p(w), _p(gx))
    return gx


# ------------------------------------------------------------------------------------------------
# Reference Python glue on the path
# ------------------------------------------------------------------------------------------------
def voxelization_coords(coords, r, normalize=True, eps=0.0):
    """modules/voxelization.py:16-25 restated with torch CPU ops in the same order (the reference
    runs the very same torch ops, on the GPU).  -> (vox int32[B,3,N], norm_coords f32[B,3,N])"""
    import torch
    c = torch.as_tensor(np.asarray(coords, dtype=np.float32))
    nc = c - c.mean(2, keepdim=True)
    if normalize:
        nc = nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0 + eps) + 0.5
    else:
        nc = (nc + 1) / 2.0
    nc = torch.clamp(nc * r, 0, r - 1)
    vox = torch.round(nc).to(torch.int32)
    return vox.numpy(), nc.numpy()


# ------------------------------------------------------------------------------------------------
# Projection conditioning (PARITY UNPINNED: pytorch3d is not vendored; see bdm_oracle.c)
# ------------------------------------------------------------------------------------------------
def project_points(points, R, T, focal, pp):
    """points [B,N,3] -> ndc xy + view z [B,N,3]"""
    points, R, T, focal, pp = _f32(points), _f32(R), _f32(T), _f32(focal), _f32(pp)
    b, n, _ = points.shape
    ndc = np.empty((b, n, 3), np.float32)
    lib().orc_project_points(_I(b), _I(n), _p(points), _p(R), _p(T), _p(focal), _p(pp), _p(ndc))
    return ndc


def rasterize_points(ndc, H, W, radius):
    """-> int32[B,H,W] winning point index or -1 (K=1, nearest z, earlier index wins ties)"""
    ndc = _f32(ndc)
    b, n, _ = ndc.shape
    zi = np.empty((b, H, W), np.int32)
    lib().orc_rasterize_points(_I(b), _I(n), _I(H), _I(W), _F(radius), _p(ndc), _p(zi))
    return zi


def surface_projection(points, R, T, focal, pp, local_features, radius=0.0075, scale_factor=1.0):
    """projection_model.py:127-157 -> (out f32[B,N,C], zbuf_idx int32[B,H,W])"""
    local_features = _f32(local_features)
    b, C, H, W = local_features.shape
    T = _f32(T) * np.float32(scale_factor)  # projection_model.py:136-137
    ndc = project_points(points, R, T, focal, pp)
    zi = rasterize_points(ndc, H, W, radius)
    n = ndc.shape[1]
    out = np.empty((b, n, C), np.float32)
    lib().orc_splat_features(_I(b), _I(n), _I(C), _I(H), _I(W), _p(zi), _p(local_features), _p(out))
    return out, zi


# ------------------------------------------------------------------------------------------------
# Evaluation (fp64)
# ------------------------------------------------------------------------------------------------
def nn_direct(src, tgt):
    """src [B,N,3], tgt [B,M,3] (fp64) -> (min squared distance [B,N], argmin int32[B,N])"""
    src, tgt = _f64(src), _f64(tgt)
    b, n, _ = src.shape
    m = tgt.shape[1]
    d = np.empty((b, n), np.float64)
    i = np.empty((b, n), np.int32)
    lib().orc_nn_direct_f64(_I(b), _I(n), _I(m), _p(src), _p(tgt), _p(d), _p(i))
    return d, i


def nn_expanded(src, tgt):
    """evaluation_f1.py:90-98 compute_pc_to_pc_dist, batched: -> min clamped squared distance [B,N]"""
    src, tgt = _f64(src), _f64(tgt)
    b, n, _ = src.shape
    m = tgt.shape[1]
    d = np.empty((b, n), np.float64)
    lib().orc_nn_expanded_f64(_I(b), _I(n), _I(m), _p(src), _p(tgt), _p(d))
    return d


def chamfer_distance(pred, gt):
    """evaluation_cd.py:111-125: inputs are mean-centred by the caller; pytorch3d chamfer_distance
    defaults (squared L2, point_reduction='mean', batch_reduction='mean') -> per-pair CD [B]"""
    d1, _ = nn_direct(pred, gt)
    d2, _ = nn_direct(gt, pred)
    return d1.mean(axis=1) + d2.mean(axis=1)


def fscore(gt, pred, thr=0.01):
    """evaluation_f1.py:101-110 cal_fscore, batched -> per-pair F [B]"""
    d1 = nn_expanded(gt, pred)
    d2 = nn_expanded(pred, gt)
    precision = (d1 < thr).sum(axis=1) / float(d1.shape[1])
    recall = (d2 < thr).sum(axis=1) / float(d2.shape[1])
    return 2 * recall * precision / (recall + precision + 1e-12)


# ------------------------------------------------------------------------------------------------
# sparse first convolution of a PVConv block (checker for csrc/sparse_conv.cu)
# ------------------------------------------------------------------------------------------------
def avg_voxelize_compact(features, coords, r):
    """The non-empty columns of avg_voxelize's grid, in ascending voxel id, zero-padded to N columns:
    -> (compact f32[B,C,N], occupied voxel ids: list of int arrays).  The dense grid is the reference's
    (vox.cpp:17-43); this only selects from it."""
    dense, _, cnt = avg_voxelize_forward(features, coords, r)
    b, c, n = np.asarray(features).shape
    out = np.zeros((b, c, n), np.float32)
    occupied = []
    for i in range(b):
        occ = np.flatnonzero(cnt[i] > 0)
        occupied.append(occ)
        out[i, :, :len(occ)] = dense[i][:, occ]
    return out, occupied


def sparse_conv3_gather(taps, occupied, r, bias=None):
    """What `nn.Conv3d(cin, cout, 3, stride=1, padding=1)` (modules/pvconv.py:75-76, cross-correlation, zero
    padding) yields on a grid whose only non-zero voxels are `occupied`, given the per-voxel tap products
    taps[b, j, k, co] = sum_ci W[co, ci, kd, kh, kw] * x[b, ci, occupied[b][j]], k = (kd*3+kh)*3+kw:
    out[b, co, x, y, z] = bias[co] + sum_k taps[b, slot(x+kd-1, y+kh-1, z+kw-1), k, co], k ascending, fp32."""
    taps = _f32(taps)
    b, n = taps.shape[0], taps.shape[1]
    cout = taps.shape[2] // 27
    taps = taps.reshape(b, n, 27, cout)
    out = np.zeros((b, cout, r, r, r), np.float32)
    for i in range(b):
        occ = np.asarray(occupied[i], dtype=np.int64)
        ox, oy, oz = occ // (r * r), (occ // r) % r, occ % r
        for k in range(27):
            kd, kh, kw = k // 9, (k // 3) % 3, k % 3
            x, y, z = ox - (kd - 1), oy - (kh - 1), oz - (kw - 1)      # the output voxel this tap lands on
            ok = (x >= 0) & (x < r) & (y >= 0) & (y < r) & (z >= 0) & (z < r)
            # each output voxel receives at most one contribution per k, so += is an ordered fp32 add
            out[i][:, x[ok], y[ok], z[ok]] += taps[i, np.flatnonzero(ok), k, :].T
    if bias is not None:
        out += _f32(bias).reshape(1, cout, 1, 1, 1)
    return out


# ------------------------------------------------------------------------------------------------
# reverse-diffusion updates around the path (checker for csrc/sampler.cu and bdm_b200.diffusion)
# ------------------------------------------------------------------------------------------------
def pvd_coefficients(b_start=1e-4, b_end=0.02, time_num=1000):
    """GaussianDiffusion.__init__, pvd/__init__.py:18-68, with the betas of :477 (linear, float64 -> float32
    exactly where the reference casts).  -> dict of float32 arrays indexed by timestep."""
    import torch
    betas = np.linspace(b_start, b_end, time_num).astype(np.float64)
    alphas = 1.0 - betas
    ac = torch.from_numpy(np.cumprod(alphas, axis=0)).float()
    ac_prev = torch.from_numpy(np.append(1.0, ac[:-1])).float()
    b32, a32 = torch.from_numpy(betas).float(), torch.from_numpy(alphas).float()
    post_var = b32 * (1.0 - ac_prev) / (1.0 - ac)
    post_log_var = torch.log(torch.max(post_var, 1e-20 * torch.ones_like(post_var)))
    return dict(sqrt_recip_ac=torch.sqrt(1.0 / ac).numpy(), sqrt_recipm1_ac=torch.sqrt(1.0 / ac - 1).numpy(),
                coef1=(b32 * torch.sqrt(ac_prev) / (1.0 - ac)).numpy(),
                coef2=((1.0 - ac_prev) * torch.sqrt(a32) / (1.0 - ac)).numpy(),
                sigma=torch.exp(0.5 * post_log_var).numpy())


def pvd_p_sample(x_t, eps, noise, t, coef=None):
    """GaussianDiffusion.p_sample, pvd/__init__.py:196-224 with model_mean_type='eps' (:184-193),
    model_var_type='fixedsmall', clip_denoised=False: every product / sum a separate fp32 rounding."""
    c = coef if coef is not None else pvd_coefficients()
    x_t, eps, noise = _f32(x_t), _f32(eps), _f32(noise)
    f = np.float32
    x0 = f(c["sqrt_recip_ac"][t]) * x_t - f(c["sqrt_recipm1_ac"][t]) * eps
    mean = f(c["coef1"][t]) * x0 + f(c["coef2"][t]) * x_t
    if t == 0:
        return mean
    return mean + f(c["sigma"][t]) * noise


def ddpm_coefficients(beta_start=1e-5, beta_end=8e-3, steps=1000):
    """diffusers 0.21.0 DDPMScheduler(beta_schedule='linear', variance_type='fixed_small', clip_sample=False,
    prediction_type='epsilon') as configured by model/model.py:51-62 and config/structured.py:105-107.
    diffusers is not vendored in the reference: PARITY UNPINNED (restated from the published algorithm)."""
    import torch
    betas = torch.linspace(beta_start, beta_end, steps, dtype=torch.float32)
    ac = torch.cumprod(1.0 - betas, dim=0)
    rows = np.zeros((steps, 5), np.float32)
    for t in range(steps):
        a_t = ac[t]
        a_prev = ac[t - 1] if t > 0 else torch.tensor(1.0)
        cur_alpha = a_t / a_prev
        cur_beta = 1 - cur_alpha
        coef_x0 = (a_prev ** 0.5 * cur_beta) / (1 - a_t)
        coef_xt = cur_alpha ** 0.5 * (1 - a_prev) / (1 - a_t)
        var = torch.clamp((1 - a_prev) / (1 - a_t) * cur_beta, min=1e-20)
        sigma = float(var ** 0.5) if t > 0 else 0.0
        rows[t] = [float((1 - a_t) ** 0.5), 1.0 / float(a_t ** 0.5), float(coef_x0), float(coef_xt), sigma]
    return rows


def ddpm_step(x_t, eps, noise, t, rows=None):
    """prev_sample of DDPMScheduler.step: x0 = (x_t - sqrt(1-abar) eps) / sqrt(abar) -- evaluated, as torch's
    CUDA kernel for `tensor / host scalar` does, as a product with the reciprocal rounded to float --,
    mean = coef_x0 x0 + coef_xt x_t, plus sigma * noise for t > 0."""
    r = (rows if rows is not None else ddpm_coefficients())[t]
    x_t, eps, noise = _f32(x_t), _f32(eps), _f32(noise)
    x0 = (x_t - r[0] * eps) * r[1]
    prev = r[2] * x0 + r[3] * x_t
    if t > 0:
        prev = prev + r[4] * noise
    return prev


def voxelization_coords_rounded_mean(coords, r, normalize=True, eps=0.0):
    """modules/voxelization.py:16-25 with every elementwise step in fp32, exactly as torch evaluates it, and the mean
    taken as the CORRECTLY ROUNDED fp32 mean (double accumulation) instead of in a particular fp32 summation order
    -- the arithmetic of csrc/voxel_coords.cu.  -> (vox int32[B,3,N], norm_coords f32[B,3,N])"""
    x = np.asarray(coords, dtype=np.float32)
    f = np.float32
    mean = (x.astype(np.float64).sum(axis=2, keepdims=True) / x.shape[2]).astype(np.float32)
    c = x - mean
    if normalize:
        sq = (c[:, 0] * c[:, 0] + c[:, 1] * c[:, 1]) + c[:, 2] * c[:, 2]          # three rounded squares, summed left to right
        extent = (np.sqrt(sq).max(axis=1) * f(2.0) + f(eps)).astype(np.float32).reshape(-1, 1, 1)
        u = c / extent + f(0.5)
    else:
        u = (c + f(1.0)) / f(2.0)
    nc = np.clip(u * f(r), f(0.0), f(r - 1)).astype(np.float32)
    return np.rint(nc).astype(np.int32), nc
