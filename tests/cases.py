"""Seeded synthetic inputs shared by the golden-vector generator (tests/golden/make_golden.py), the
oracle tests and the GPU parity tests.  numpy only, so the same bytes are produced everywhere."""
import numpy as np


def cloud(rng, b, n, regime="noise"):
    """[b,3,n] float32.  'noise' ~ N(0,1) (what the sampler sees at t~999); 'shape' = points on an
    ellipsoid surface + 1% jitter, zero-mean / unit-std (dataset normalisation)."""
    if regime == "noise":
        x = rng.standard_normal((b, n, 3))
    else:
        d = rng.standard_normal((b, n, 3))
        d /= np.linalg.norm(d, axis=-1, keepdims=True)
        x = d * np.array([1.0, 0.6, 0.8]) + 0.01 * rng.standard_normal((b, n, 3))
        x = x - x.mean(axis=1, keepdims=True)
        x = x / x.reshape(b, -1).std(axis=1).reshape(b, 1, 1)
    return np.ascontiguousarray(x.transpose(0, 2, 1)).astype(np.float32)


def vox_coords(coords, r):
    """modules/voxelization.py:16-25 in numpy float32 (used only to build *inputs*; the exact torch
    restatement lives in oracle.voxelization_coords)."""
    c = coords - coords.mean(axis=2, keepdims=True, dtype=np.float32)
    nrm = np.sqrt((c * c).sum(axis=1, keepdims=True)).max(axis=2, keepdims=True)
    nc = (c / (nrm * np.float32(2.0)) + np.float32(0.5)).astype(np.float32)
    nc = np.clip(nc * np.float32(r), 0, r - 1).astype(np.float32)
    return np.round(nc).astype(np.int32), nc


def fps_chain_centers(coords, idx):
    b = coords.shape[0]
    return np.stack([coords[i][:, idx[i]] for i in range(b)]).astype(np.float32)


# name -> builder(rng) -> dict of inputs.  Sizes are small: these feed committed golden files.
def case_voxelize(rng, b=2, c=5, n=1000, r=8, regime="shape"):
    co = cloud(rng, b, n, regime)
    vox, nc = vox_coords(co, r)
    feat = rng.standard_normal((b, c, n)).astype(np.float32)
    return dict(features=feat, coords=vox, norm_coords=nc, r=np.int32(r))


def case_devoxelize(rng, b=2, c=6, n=777, r=8):
    co = cloud(rng, b, n, "noise")
    _, nc = vox_coords(co, r)
    # force some exact-integer coordinates and the r-1 boundary (trilinear_devox.cu:64-75 trick)
    nc[:, :, :8] = np.round(nc[:, :, :8])
    nc[0, 0, 8] = r - 1
    nc[0, 1, 9] = r - 1
    nc[0, 2, 10] = r - 1
    nc[1, :, 11] = r - 1
    nc[1, :, 12] = 0
    grid = rng.standard_normal((b, c, r * r * r)).astype(np.float32)
    return dict(coords=nc.astype(np.float32), features=grid, r=np.int32(r))


def case_fps(rng, b=3, n=1500, m=300, regime="noise", dup=0):
    co = cloud(rng, b, n, regime)
    if dup:  # exact duplicates -> exact distance ties -> exercises the (k mod 512, k) tie rule
        src = rng.integers(0, n, size=dup)
        dst = rng.integers(0, n, size=dup)
        co[:, :, dst] = co[:, :, src]
    return dict(coords=co, m=np.int32(m))


def case_fps_grid(rng, b=2, side=9, m=200):
    """points on an integer lattice: massive exact ties in every round"""
    g = np.stack(np.meshgrid(*[np.arange(side)] * 3, indexing="ij"), 0).reshape(3, -1).astype(np.float32)
    co = np.stack([g[:, rng.permutation(g.shape[1])] for _ in range(b)])
    return dict(coords=np.ascontiguousarray(co), m=np.int32(m))


def case_ball_query(rng, b=2, n=1024, m=256, radius=0.2, u=32, regime="shape"):
    co = cloud(rng, b, n, regime)
    idx = np.stack([rng.permutation(n)[:m] for _ in range(b)]).astype(np.int32)
    cen = fps_chain_centers(co, idx)
    return dict(centers=cen, points=co, radius=np.float32(radius), u=np.int32(u))


def case_ball_query_nohit(rng, b=2, n=300, m=37, u=8):
    co = cloud(rng, b, n, "noise")
    cen = (cloud(rng, b, m, "noise") + np.float32(100.0)).astype(np.float32)
    cen[:, :, :5] = co[:, :, 10:15]  # a few centres that do hit
    return dict(centers=cen, points=co, radius=np.float32(0.05), u=np.int32(u))


def case_grouping(rng, b=2, c=7, n=500, m=60, u=16):
    feat = rng.standard_normal((b, c, n)).astype(np.float32)
    idx = rng.integers(0, n, size=(b, m, u)).astype(np.int32)
    return dict(features=feat, indices=idx)


def case_gather(rng, b=2, c=3, n=500, m=123):
    feat = rng.standard_normal((b, c, n)).astype(np.float32)
    idx = rng.integers(0, n, size=(b, m)).astype(np.int32)
    return dict(features=feat, indices=idx)


def case_three_nn(rng, b=2, c=9, n=1000, m=100, regime="shape", dup_centers=0):
    co = cloud(rng, b, n, regime)
    idx = np.stack([rng.permutation(n)[:m] for _ in range(b)]).astype(np.int32)
    cen = fps_chain_centers(co, idx)
    if dup_centers:
        cen[:, :, m - dup_centers:] = cen[:, :, :dup_centers]  # equal distances -> earlier index first
    feat = rng.standard_normal((b, c, m)).astype(np.float32)
    return dict(points=co, centers=cen, features=feat)


GOLDEN_CASES = {
    "voxelize_r8": (case_voxelize, dict(r=8)),
    "voxelize_r16_noise": (case_voxelize, dict(r=16, n=1024, c=4, regime="noise")),
    "voxelize_r5_odd": (case_voxelize, dict(r=5, n=333, c=3)),
    "devoxelize_r8": (case_devoxelize, dict()),
    "devoxelize_r16": (case_devoxelize, dict(r=16, n=1024, c=4)),
    "fps_noise": (case_fps, dict()),
    "fps_dups": (case_fps, dict(b=2, n=1100, m=400, dup=500)),
    "fps_small": (case_fps, dict(b=2, n=64, m=16)),
    "fps_m_gt_n": (case_fps, dict(b=1, n=40, m=60)),
    "fps_lattice": (case_fps_grid, dict()),
    "ball_query": (case_ball_query, dict()),
    "ball_query_dense": (case_ball_query, dict(n=512, m=64, radius=0.9, u=16)),
    "ball_query_nohit": (case_ball_query_nohit, dict()),
    "grouping": (case_grouping, dict()),
    "gather": (case_gather, dict()),
    "three_nn": (case_three_nn, dict()),
    "three_nn_dups": (case_three_nn, dict(n=300, m=40, c=4, dup_centers=10)),
    "three_nn_m2": (case_three_nn, dict(n=50, m=2, c=3)),
}


def build_case(name):
    fn, kw = GOLDEN_CASES[name]
    seed = int.from_bytes(name.encode(), "little") % (2 ** 31)
    return fn(np.random.default_rng(seed), **kw)
