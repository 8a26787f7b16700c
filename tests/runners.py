"""Run one seeded case through (a) the CPU oracle, (b) a torch-tensor `_backend` (ours on CUDA, or the
reference's own extension) and return numpy outputs with identical keys, so tests can diff them."""
import numpy as np

from . import cases


def _family(name):
    for fam in ("voxelize", "devoxelize", "fps", "ball_query", "grouping", "gather", "three_nn"):
        if name.startswith(fam):
            return fam
    raise KeyError(name)


def _grad_like(shape, name):
    seed = (int.from_bytes(name.encode(), "little") + 7) % (2 ** 31)
    return np.random.default_rng(seed).standard_normal(shape).astype(np.float32)


def run_oracle(name, inp=None):
    import oracle as O
    inp = inp if inp is not None else cases.build_case(name)
    fam = _family(name)
    if fam == "voxelize":
        out, ind, cnt = O.avg_voxelize_forward(inp["features"], inp["coords"], int(inp["r"]))
        gy = _grad_like(out.shape, name)
        return dict(out=out, ind=ind, cnt=cnt, grad_x=O.avg_voxelize_backward(gy, ind, cnt))
    if fam == "devoxelize":
        r = int(inp["r"])
        outs, inds, wgts = O.trilinear_devoxelize_forward(r, True, inp["coords"], inp["features"])
        outs_eval, _, _ = O.trilinear_devoxelize_forward(r, False, inp["coords"], inp["features"])
        gy = _grad_like(outs.shape, name)
        return dict(outs=outs, inds=inds, wgts=wgts, outs_eval=outs_eval,
                    grad_x=O.trilinear_devoxelize_backward(gy, inds, wgts, r))
    if fam == "fps":
        idx = O.furthest_point_sampling(inp["coords"], int(inp["m"]))
        return dict(indices=idx, centers=O.gather_features_forward(inp["coords"], idx))
    if fam == "ball_query":
        return dict(neighbors=O.ball_query(inp["centers"], inp["points"], float(inp["radius"]), int(inp["u"])))
    if fam == "grouping":
        out = O.grouping_forward(inp["features"], inp["indices"])
        gy = _grad_like(out.shape, name)
        return dict(out=out, grad_x=O.grouping_backward(gy, inp["indices"], inp["features"].shape[2]))
    if fam == "gather":
        out = O.gather_features_forward(inp["features"], inp["indices"])
        gy = _grad_like(out.shape, name)
        return dict(out=out, grad_x=O.gather_features_backward(gy, inp["indices"], inp["features"].shape[2]))
    if fam == "three_nn":
        out, idx, w = O.three_nearest_neighbors_interpolate_forward(inp["points"], inp["centers"], inp["features"])
        gy = _grad_like(out.shape, name)
        return dict(out=out, idx=idx, w=w,
                    grad_x=O.three_nearest_neighbors_interpolate_backward(gy, idx, w, inp["centers"].shape[2]))
    raise KeyError(name)


def run_backend(name, be, inp=None, device="cuda"):
    """`be` exposes the reference's 12 pybind names over torch CUDA tensors (bindings.cpp:10-37)."""
    import torch
    inp = inp if inp is not None else cases.build_case(name)
    fam = _family(name)
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(device)  # noqa: E731
    N = lambda t: t.detach().cpu().numpy()  # noqa: E731
    if fam == "voxelize":
        out, ind, cnt = be.avg_voxelize_forward(T(inp["features"]), T(inp["coords"]), int(inp["r"]))
        gy = T(_grad_like(tuple(out.shape), name))
        gx = be.avg_voxelize_backward(gy, ind, cnt)
        return dict(out=N(out), ind=N(ind), cnt=N(cnt), grad_x=N(gx))
    if fam == "devoxelize":
        r = int(inp["r"])
        outs, inds, wgts = be.trilinear_devoxelize_forward(r, True, T(inp["coords"]), T(inp["features"]))
        outs_eval, di, dw = be.trilinear_devoxelize_forward(r, False, T(inp["coords"]), T(inp["features"]))
        assert tuple(di.shape) == (1,) and tuple(dw.shape) == (1,)
        gy = T(_grad_like(tuple(outs.shape), name))
        gx = be.trilinear_devoxelize_backward(gy, inds, wgts, r)
        return dict(outs=N(outs), inds=N(inds), wgts=N(wgts), outs_eval=N(outs_eval), grad_x=N(gx))
    if fam == "fps":
        co = T(inp["coords"])
        idx = be.furthest_point_sampling(co, int(inp["m"]))
        return dict(indices=N(idx), centers=N(be.gather_features_forward(co, idx)))
    if fam == "ball_query":
        nb = be.ball_query(T(inp["centers"]), T(inp["points"]), float(inp["radius"]), int(inp["u"]))
        return dict(neighbors=N(nb))
    if fam == "grouping":
        idx = T(inp["indices"])
        out = be.grouping_forward(T(inp["features"]), idx)
        gy = T(_grad_like(tuple(out.shape), name))
        return dict(out=N(out), grad_x=N(be.grouping_backward(gy, idx, inp["features"].shape[2])))
    if fam == "gather":
        idx = T(inp["indices"])
        out = be.gather_features_forward(T(inp["features"]), idx)
        gy = T(_grad_like(tuple(out.shape), name))
        return dict(out=N(out), grad_x=N(be.gather_features_backward(gy, idx, inp["features"].shape[2])))
    if fam == "three_nn":
        out, idx, w = be.three_nearest_neighbors_interpolate_forward(T(inp["points"]), T(inp["centers"]),
                                                                     T(inp["features"]))
        gy = T(_grad_like(tuple(out.shape), name))
        gx = be.three_nearest_neighbors_interpolate_backward(gy, idx, w, inp["centers"].shape[2])
        return dict(out=N(out), idx=N(idx), w=N(w), grad_x=N(gx))
    raise KeyError(name)


# Comparison policy (BASELINE.json north_star): integer outputs bit-exact; floats within 1e-5 RELATIVE error,
# element by element; atomic-order-dependent sums (voxel averages, all backward scatter-adds) within 1e-4.
# A sum of terms of mixed sign can land arbitrarily close to zero while its rounding error stays at the size
# of the terms, so every element also gets an absolute floor of FLOOR_FRAC * rtol * (largest magnitude in the
# tensor): |got - want| <= rtol * |want| + rtol * FLOOR_FRAC * max|want|.
EXACT_KEYS = {"ind", "cnt", "inds", "indices", "neighbors", "idx", "centers"}
ATOMIC_KEYS = {"grad_x"}
FLOOR_FRAC = 1e-2


def assert_close(label, got, want, rtol, floor_frac=FLOOR_FRAC):
    g, w = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert g.shape == w.shape, f"{label}: shape {g.shape} vs {w.shape}"
    if w.size == 0:
        return
    assert np.isfinite(g).all() or not np.isfinite(w).all(), f"{label}: non-finite values"
    peak = float(np.abs(w).max())
    bound = rtol * np.abs(w) + rtol * floor_frac * max(peak, 1e-30)
    err = np.abs(g - w)
    worst = int(np.argmax(err - bound))
    assert (err <= bound).all(), (f"{label}: element {worst}: got {g.flat[worst]!r}, want {w.flat[worst]!r} "
                                  f"(|diff| {err.flat[worst]:.3e} > {bound.flat[worst]:.3e}; rtol {rtol:.0e})")


def compare(name, got, want, atol_scale=1.0):
    fam = _family(name)
    for k, w in want.items():
        g = got[k]
        assert g.shape == w.shape, f"{name}.{k}: shape {g.shape} vs {w.shape}"
        if k in EXACT_KEYS or w.dtype.kind in "iu":
            bad = int((g != w).sum())
            assert bad == 0, f"{name}.{k}: {bad} of {w.size} entries differ (must be bit-exact)"
        else:
            rtol = 1e-4 if (k in ATOMIC_KEYS or (fam == "voxelize" and k == "out")) else 1e-5
            assert_close(f"{name}.{k}", g, w, rtol * atol_scale)
