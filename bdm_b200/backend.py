"""`_backend`: the reference's 12-function native module, served by libbdm_b200.so.

Mirrors the pybind module `_pvcnn_backend`
(/root/reference/experiments/model/pvcnn/modules/functional/src/bindings.cpp:10-37; loaded by
functional/backend.py:12-31 as `_backend`): same function names, same positional arguments, same
return structure, same input checks (CUDA / contiguous / dtype -> RuntimeError, src/utils.hpp:7-18).

Differences that are invisible to callers: outputs are allocated with torch.empty (every kernel
writes its whole output; the reference zero-fills first), all launches go to the *current* stream of
the tensor's device (the reference mixes the legacy default stream and the current stream, and never
sets the device), and launch failures raise instead of calling exit(-1).

Extras beyond the 12 names (used by bdm_b200.modules to avoid duplicated work, never required):
three_nn_search, three_nn_interpolate, surface_projection, nn_f64.
"""
import numpy as np
import torch

from . import _lib

_L = _lib.lib
_check = _lib.check
_I32 = torch.int32
_F32 = torch.float32


def _req(cond, msg):
    if not cond:
        raise RuntimeError(msg)


# ---------------------------------------------------------------------------------------------
# launch accounting + optional per-call CUDA-event timing (used by bench.py; off by default)
# ---------------------------------------------------------------------------------------------
LAUNCHES = 0          # kernels of libbdm_b200.so enqueued by this process (memsets not counted)
_PROFILE = None       # None, or {op name: [(start_event, end_event, shapes), ...]}


def profile_start():
    global _PROFILE
    _PROFILE = {}


def profile_stop():
    """-> {op name: [(milliseconds, shapes), ...]}; synchronises the device."""
    global _PROFILE
    rec, _PROFILE = _PROFILE, None
    torch.cuda.synchronize()
    return {k: [(e0.elapsed_time(e1), shp) for e0, e1, shp in v] for k, v in (rec or {}).items()}


def _count_launches(n):
    global LAUNCHES
    LAUNCHES += int(n)


def _op(launches):
    """Decorator: count kernel launches; when profiling, bracket the call with CUDA events on the
    current stream of the first tensor argument's device."""
    def deco(fn):
        name = fn.__name__

        def wrapper(*args, **kwargs):
            global LAUNCHES
            LAUNCHES += launches
            if _PROFILE is None:
                return fn(*args, **kwargs)
            shapes = tuple(tuple(a.shape) if isinstance(a, torch.Tensor) else a for a in args)
            if kwargs:   # keyword arguments as one trailing dict (tensor -> shape, None / scalars as they are)
                shapes += ({k: (tuple(v.shape) if isinstance(v, torch.Tensor) else v) for k, v in kwargs.items()},)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*args, **kwargs)
            e1.record()
            _PROFILE.setdefault(name, []).append((e0, e1, shapes))
            return out
        wrapper.__name__ = name
        wrapper.__doc__ = fn.__doc__
        wrapper.__wrapped__ = fn
        return wrapper
    return deco


def _chk_float(x, name):
    _req(x.is_cuda, f"{name} must be a CUDA tensor")
    _req(x.is_contiguous(), f"{name} must be a contiguous tensor")
    _req(x.dtype == _F32, f"{name} must be a float tensor")


def _chk_int(x, name):
    _req(x.is_cuda, f"{name} must be a CUDA tensor")
    _req(x.is_contiguous(), f"{name} must be a contiguous tensor")
    _req(x.dtype == _I32, f"{name} must be an int tensor")


def _chk_channel_vector(t, name, c):
    """optional per-channel f32 parameter (GroupNorm affine, conv bias): CUDA, contiguous, float32, C values"""
    if t is None:
        return
    _chk_float(t, name)
    _req(t.numel() == c, f"{name} must hold one value per channel ({c})")


class _Launch:
    """Device guard + current stream of the device that owns `ref`."""
    __slots__ = ("dev", "prev", "stream")

    def __init__(self, ref):
        self.dev = ref.device.index
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.dev:
            self.prev = cur
            torch.cuda.set_device(self.dev)
        self.stream = torch.cuda.current_stream().cuda_stream
        return self.stream

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


def _workspace(nbytes, device):
    # torch's caching allocator makes this stream-safe and cheap; contents are irrelevant
    return torch.empty((max(int(nbytes), 16),), dtype=torch.uint8, device=device)


# ---------------------------------------------------------------------------------------------
# voxelization  (vox.cpp:17-43, :54-76)
# ---------------------------------------------------------------------------------------------
class VoxelPlan:
    """Everything avg_voxelize derives from the coordinates alone: per-point voxel index, per-voxel
    count, and the sorted lookup tables living in `workspace` (opaque)."""
    __slots__ = ("b", "n", "r", "ind", "cnt", "workspace")

    def __init__(self, b, n, r, ind, cnt, workspace):
        self.b, self.n, self.r, self.ind, self.cnt, self.workspace = b, n, r, ind, cnt, workspace


@_op(1)
def voxelize_coords(coords, resolution, normalize=True, eps=0.0):
    """coords f32[B,3,N] -> (float voxel coordinates f32[B,3,N] in [0, R-1], int32[B,3,N] = their round-half-even):
    the torch op sequence of Voxelization.forward (modules/voxelization.py:17-24) as one kernel"""
    _chk_float(coords, "coords")
    _req(coords.dim() == 3 and coords.shape[1] == 3, "coords must be [B,3,N]")
    b, n = coords.shape[0], coords.shape[2]
    norm = torch.empty_like(coords)
    vox = torch.empty((b, 3, n), dtype=_I32, device=coords.device)
    with _Launch(coords) as st:
        _check(_L.bdm_voxelize_coords(b, n, int(resolution), 1 if normalize else 0, float(eps), coords.data_ptr(),
                                      norm.data_ptr(), vox.data_ptr(), st))
    return norm, vox


@_op(1)
def voxel_plan(coords, resolution):
    """coords int32[B,3,N] (voxel coordinates in [0,R)) -> VoxelPlan"""
    _chk_int(coords, "coords")
    b, n = coords.shape[0], coords.shape[2]
    r = int(resolution)
    dev = coords.device
    ind = torch.empty((b, n), dtype=_I32, device=dev)
    cnt = torch.empty((b, r * r * r), dtype=_I32, device=dev)
    ws = _workspace(_L.bdm_avg_voxelize_workspace_bytes(b, n, r), dev)
    with _Launch(coords) as st:
        _check(_L.bdm_voxel_plan(b, n, r, coords.data_ptr(), ind.data_ptr(), cnt.data_ptr(), ws.data_ptr(),
                                 ws.numel(), st))
    return VoxelPlan(b, n, r, ind, cnt, ws)


@_op(1)
def avg_voxelize_fill(features, plan):
    """features f32[B,C,N] + VoxelPlan -> dense grid f32[B,C,R^3]"""
    _chk_float(features, "features")
    b, c, n = features.shape
    _req(b == plan.b and n == plan.n, "features do not match the voxel plan")
    out = torch.empty((b, c, plan.r ** 3), dtype=_F32, device=features.device)
    with _Launch(features) as st:
        _check(_L.bdm_avg_voxelize_fill(b, c, n, plan.r, plan.ind.data_ptr(), plan.cnt.data_ptr(),
                                        features.data_ptr(), out.data_ptr(), plan.workspace.data_ptr(),
                                        plan.workspace.numel(), st))
    return out


@_op(1)
def avg_voxelize_compact(features, plan, amax_into=None):
    """features f32[B,C,N] + VoxelPlan -> f32[B,C,N]: column j = average of the j-th occupied voxel
    (ascending voxel id), zero past the shape's occupied count.
    amax_into: a prepared conv3_tc05 weight buffer whose header also receives max|out| (the dynamic fp16 scale of
    conv3_tc05_fill_planes(amax_ready=True))."""
    _chk_float(features, "features")
    b, c, n = features.shape
    _req(b == plan.b and n == plan.n, "features do not match the voxel plan")
    out = torch.empty((b, c, n), dtype=_F32, device=features.device)
    with _Launch(features) as st:
        if amax_into is not None:
            _req(amax_into.is_cuda and amax_into.dtype == torch.uint8 and amax_into.numel() >= 256, "amax_into must be a prepared weight buffer")
            _check(_L.bdm_avg_voxelize_compact_amax(b, c, n, plan.r, features.data_ptr(), out.data_ptr(),
                                                    plan.workspace.data_ptr(), plan.workspace.numel(),
                                                    amax_into.data_ptr() + 16, st))
        else:
            _check(_L.bdm_avg_voxelize_compact(b, c, n, plan.r, features.data_ptr(), out.data_ptr(),
                                               plan.workspace.data_ptr(), plan.workspace.numel(), st))
    return out


@_op(1)
def sparse_conv3_gather(taps, plan, bias=None, channels_last=False, stats=False):
    """taps f32[B,N,27*Cout] (per-occupied-voxel tap products) + VoxelPlan -> the dense output
    f32[B,Cout,R,R,R] of the zero-padded 3x3x3 convolution (f32[B,R,R,R,Cout] when channels_last).
    stats: also return f64[B,blocks,Cout,2], the per-channel (sum, sum of squares) of the bias-less output
    per block of rows, which groupnorm_act_cl(..., partials=) takes instead of reading the tensor again."""
    _chk_float(taps, "taps")
    b, n, k = taps.shape
    _req(b == plan.b and n == plan.n and k % 27 == 0, "taps do not match the voxel plan")
    cout, r = k // 27, plan.r
    if bias is not None:
        _chk_float(bias, "bias")
        _req(bias.numel() == cout, "bias must hold one value per output channel")
    out = torch.empty((b, r, r, r, cout) if channels_last else (b, cout, r, r, r), dtype=_F32, device=taps.device)
    part = None
    if stats:
        part = torch.empty((b, _L.bdm_sparse_conv3_stats_blocks(r), cout, 2), dtype=torch.float64, device=taps.device)
    with _Launch(taps) as st:
        _check(_L.bdm_sparse_conv3_gather(b, cout, n, r, taps.data_ptr(), bias.data_ptr() if bias is not None else None,
                                          out.data_ptr(), 1 if channels_last else 0,
                                          part.data_ptr() if part is not None else None, plan.workspace.data_ptr(),
                                          plan.workspace.numel(), st))
    return (out, part) if stats else out


@_op(1)
def trilinear_devoxelize_cl(grid, coords, resolution, gate=None, residual=None, norm_coef=None, swish=True):
    """grid f32[B,R,R,R,C] (channels last) + float voxel coordinates f32[B,3,N] -> f32[B,C,N]; inference only.
    gate f32[B,C] / residual f32[B,C,N]: out = devox * gate + residual (the tail of a PVConv block).
    norm_coef f32[B,C,2] (groupnorm_cl_sums): `grid` is un-normalised and every corner value goes through
    act(x*A + B) first -- devoxelization of GroupNorm(+Swish)(grid) without materialising it."""
    _chk_float(grid, "grid")
    _chk_float(coords, "coords")
    b, c, n, r = grid.shape[0], grid.shape[-1], coords.shape[2], int(resolution)
    _req(grid.numel() == b * r * r * r * c, "grid does not match the resolution")
    if gate is not None:
        _chk_float(gate, "gate")
        _req(tuple(gate.shape) == (b, c), "gate must be [B,C]")
    if residual is not None:
        _chk_float(residual, "residual")
        _req(tuple(residual.shape) == (b, c, n), "residual must be [B,C,N]")
    if norm_coef is not None:
        _chk_float(norm_coef, "norm_coef")
        _req(tuple(norm_coef.shape) == (b, c, 2), "norm_coef must be [B,C,2]")
    out = torch.empty((b, c, n), dtype=_F32, device=grid.device)
    with _Launch(grid) as st:
        _check(_L.bdm_trilinear_devoxelize_cl_norm(b, c, n, r, coords.data_ptr(), grid.data_ptr(),
                                                   norm_coef.data_ptr() if norm_coef is not None else None,
                                                   1 if swish else 0,
                                                   gate.data_ptr() if gate is not None else None,
                                                   residual.data_ptr() if residual is not None else None,
                                                   out.data_ptr(), st))
    return out


@_op(1)
def se_gate(sums, count, w1, w2, use_relu):
    """sums f32[B,C] or f32[B,tiles,C] (per-channel sums of the activations, e.g. from groupnorm_act*) ->
    gate f32[B,C] = sigmoid(w2 @ act(w1 @ (sums / count)))"""
    _chk_float(sums, "sums")
    _chk_float(w1, "w1")
    _chk_float(w2, "w2")
    if sums.dim() == 2:
        b, c = sums.shape
        tiles, sb, st_, sc = 1, c, 0, 1
    else:
        b, tiles, c = sums.shape
        sb, st_, sc = tiles * c, c, 1
    hidden = w1.shape[0]
    _req(tuple(w1.shape) == (hidden, c) and tuple(w2.shape) == (c, hidden), "SE weights do not match the channels")
    gate = torch.empty((b, c), dtype=_F32, device=sums.device)
    with _Launch(sums) as st:
        _check(_L.bdm_se_gate(b, c, hidden, tiles, float(count), sb, st_, sc, sums.data_ptr(), w1.data_ptr(),
                              w2.data_ptr(), 1 if use_relu else 0, gate.data_ptr(), st))
    return gate


def attention_supported(channels, tokens, batch=None):
    """shapes the fused kernel takes; with `batch` given, also whether there is enough work to fill the
    GPU (one CTA per 128 queries; below ~1 CTA per SM torch's batched GEMMs are faster)"""
    ok = channels == 64 and tokens >= 128 and tokens % 128 == 0
    if ok and batch is not None:
        ok = batch * (tokens // 128) >= 128
    return ok


@_op(3)
def attention(q, k, v):
    """q, k, v f32[B,64,T] -> f32[B,64,T]: out[b,c,i] = sum_j softmax_j(q[b,:,i].k[b,:,j]) v[b,c,j]"""
    _chk_float(q, "q")
    _chk_float(k, "k")
    _chk_float(v, "v")
    b, c, t = q.shape
    _req(k.shape == q.shape and v.shape == q.shape, "q, k, v must have one shape")
    _req(attention_supported(c, t), "attention kernel needs 64 channels and a multiple of 128 tokens")
    out = torch.empty_like(q)
    ws = _workspace(_L.bdm_attention_workspace_bytes(b, c, t), q.device)
    with _Launch(q) as st:
        _check(_L.bdm_attention(b, c, t, q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), ws.data_ptr(),
                                ws.numel(), st))
    return out


@_op(3)
def attention_qkv(qkv, bias=None):
    """qkv f32[B,T,3*64] (q | k | v per token: one fused projection of channels-last activations) + bias f32[192]
    -> f32[B,T,64] token-major: softmax(q.k) applied to v, the biases added on the way in"""
    _chk_float(qkv, "qkv")
    b, t, ld = qkv.shape
    c = ld // 3
    _req(ld == 3 * c and attention_supported(c, t), "attention kernel needs 3 x 64 columns and a multiple of 128 tokens")
    if bias is not None:
        _chk_float(bias, "bias")
        _req(bias.numel() == ld, "bias must hold 3*64 values")
    out = torch.empty((b, t, c), dtype=_F32, device=qkv.device)
    ws = _workspace(_L.bdm_attention_workspace_bytes(b, c, t), qkv.device)
    with _Launch(qkv) as st:
        _check(_L.bdm_attention_qkv(b, c, t, qkv.data_ptr(), ld, bias.data_ptr() if bias is not None else None,
                                    out.data_ptr(), ws.data_ptr(), ws.numel(), st))
    return out


def sparse_conv3_supported(n, resolution):
    r = int(resolution)
    return 1 <= r <= 32 and (r & (r - 1)) == 0 and 1 <= n <= 16384


def avg_voxelize_forward(features, coords, resolution):
    _chk_float(features, "features")
    _chk_int(coords, "coords")
    plan = voxel_plan(coords, resolution)
    return [avg_voxelize_fill(features, plan), plan.ind, plan.cnt]


@_op(1)
def avg_voxelize_backward(grad_y, indices, cnt):
    _chk_float(grad_y, "grad_y")
    _chk_int(indices, "indices")
    _chk_int(cnt, "cnt")
    b, c, s = grad_y.shape
    n = indices.shape[1]
    grad_x = torch.empty((b, c, n), dtype=_F32, device=grad_y.device)
    with _Launch(grad_y) as st:
        _check(_L.bdm_avg_voxelize_grad(b, c, n, s, indices.data_ptr(), cnt.data_ptr(), grad_y.data_ptr(),
                                        grad_x.data_ptr(), st))
    return grad_x


# ---------------------------------------------------------------------------------------------
# devoxelization  (trilinear_devox.cpp:18-55, :68-94)
# ---------------------------------------------------------------------------------------------
@_op(1)
def devoxelize_plan(coords, r):
    """Coordinate-only half of inference devoxelization (x-slice binning) -> opaque workspace tensor"""
    _chk_float(coords, "coords")
    b, n = coords.shape[0], coords.shape[2]
    r = int(r)
    ws = _workspace(_L.bdm_trilinear_devoxelize_workspace_bytes(b, n, r), coords.device)
    with _Launch(coords) as st:
        _check(_L.bdm_trilinear_devoxelize_plan(b, n, r, coords.data_ptr(), ws.data_ptr(), ws.numel(), st))
    return ws


@_op(1)
def trilinear_devoxelize_forward(r, is_training, coords, features, plan=None):
    _chk_float(features, "features")
    _chk_float(coords, "coords")
    b, c = features.shape[0], features.shape[1]
    n = coords.shape[2]
    r = int(r)
    dev = features.device
    outs = torch.empty((b, c, n), dtype=_F32, device=dev)
    if is_training:
        inds = torch.empty((b, 8, n), dtype=_I32, device=dev)
        wgts = torch.empty((b, 8, n), dtype=_F32, device=dev)
        ip, wp = inds.data_ptr(), wgts.data_ptr()
    else:  # the reference returns 1-element dummies (trilinear_devox.cpp:45-53)
        inds = torch.zeros((1,), dtype=_I32, device=dev)
        wgts = torch.zeros((1,), dtype=_F32, device=dev)
        ip, wp = None, None
    ws = plan if plan is not None else _workspace(_L.bdm_trilinear_devoxelize_workspace_bytes(b, n, r), dev)
    with _Launch(features) as st:
        _check(_L.bdm_trilinear_devoxelize(b, c, n, r, 1 if is_training else 0, coords.data_ptr(),
                                           features.data_ptr(), ip, wp, outs.data_ptr(), ws.data_ptr(),
                                           ws.numel(), 1 if plan is not None else 0, st))
    return [outs, inds, wgts]


@_op(1)
def trilinear_devoxelize_backward(grad_y, indices, weights, r):
    _chk_float(grad_y, "grad_y")
    _chk_float(weights, "weights")
    _chk_int(indices, "indices")
    b, c, n = grad_y.shape
    r3 = int(r) ** 3
    grad_x = torch.empty((b, c, r3), dtype=_F32, device=grad_y.device)
    with _Launch(grad_y) as st:
        _check(_L.bdm_trilinear_devoxelize_grad(b, c, n, r3, indices.data_ptr(), weights.data_ptr(),
                                                grad_y.data_ptr(), grad_x.data_ptr(), st))
    return grad_x


# ---------------------------------------------------------------------------------------------
# sampling  (sampling.cpp:6-58)
# ---------------------------------------------------------------------------------------------
@_op(1)
def gather_features_forward(features, indices):
    _chk_float(features, "features")
    _chk_int(indices, "indices")
    b, c, n = features.shape
    m = indices.shape[1]
    out = torch.empty((b, c, m), dtype=_F32, device=features.device)
    with _Launch(features) as st:
        _check(_L.bdm_gather_features(b, c, n, m, features.data_ptr(), indices.data_ptr(), out.data_ptr(), st))
    return out


@_op(1)
def gather_features_backward(grad_y, indices, n):
    _chk_float(grad_y, "grad_y")
    _chk_int(indices, "indices")
    b, c = grad_y.shape[0], grad_y.shape[1]
    m = indices.shape[1]
    n = int(n)
    grad_x = torch.empty((b, c, n), dtype=_F32, device=grad_y.device)
    with _Launch(grad_y) as st:
        _check(_L.bdm_gather_features_grad(b, c, n, m, grad_y.data_ptr(), indices.data_ptr(),
                                           grad_x.data_ptr(), st))
    return grad_x


@_op(1)
def furthest_point_sampling(coords, num_samples):
    _chk_float(coords, "coords")
    b, n = coords.shape[0], coords.shape[2]
    m = int(num_samples)
    dev = coords.device
    if m <= 0 or b == 0:
        return torch.zeros((b, max(m, 0)), dtype=_I32, device=dev)
    indices = torch.empty((b, m), dtype=_I32, device=dev)
    ws = _workspace(_L.bdm_furthest_point_sampling_workspace_bytes(b, n), dev)
    with _Launch(coords) as st:
        _check(_L.bdm_furthest_point_sampling(b, n, m, coords.data_ptr(), indices.data_ptr(), ws.data_ptr(),
                                              ws.numel(), st))
    return indices


# ---------------------------------------------------------------------------------------------
# ball query  (ball_query.cpp:6-30)
# ---------------------------------------------------------------------------------------------
@_op(1)
def ball_query(centers_coords, points_coords, radius, num_neighbors):
    _chk_float(centers_coords, "centers_coords")
    _chk_float(points_coords, "points_coords")
    b, m = centers_coords.shape[0], centers_coords.shape[2]
    n = points_coords.shape[2]
    u = int(num_neighbors)
    r = np.float32(radius)
    r2 = float(r * r)  # fp32 product, like `radius * radius` at ball_query.cpp:24
    out = torch.empty((b, m, u), dtype=_I32, device=centers_coords.device)
    with _Launch(centers_coords) as st:
        _check(_L.bdm_ball_query(b, n, m, r2, u, centers_coords.data_ptr(), points_coords.data_ptr(),
                                 out.data_ptr(), st))
    return out


# ---------------------------------------------------------------------------------------------
# grouping  (grouping.cpp:6-43)
# ---------------------------------------------------------------------------------------------
@_op(1)
def grouping_forward(features, indices):
    _chk_float(features, "features")
    _chk_int(indices, "indices")
    b, c, n = features.shape
    m, u = indices.shape[1], indices.shape[2]
    out = torch.empty((b, c, m, u), dtype=_F32, device=features.device)
    with _Launch(features) as st:
        _check(_L.bdm_grouping(b, c, n, m, u, features.data_ptr(), indices.data_ptr(), out.data_ptr(), st))
    return out


@_op(1)
def grouping_into(features, indices, out, channel_offset, centers=None):
    """out[:, channel_offset : channel_offset + C] = features gathered by indices (minus centers[b,c,m] when
    given); `out` f32[B,OC,M,U] contiguous.  Lets BallQuery fill its concatenated tensor in place."""
    _chk_float(features, "features")
    _chk_int(indices, "indices")
    _chk_float(out, "out")
    b, c, n = features.shape
    m, u = indices.shape[1], indices.shape[2]
    _req(out.dim() == 4 and out.shape[0] == b and out.shape[2] == m and out.shape[3] == u, "out must be [B,OC,M,U]")
    _req(0 <= channel_offset and channel_offset + c <= out.shape[1], "channel slice outside out")
    if centers is not None:
        _chk_float(centers, "centers")
        _req(tuple(centers.shape) == (b, c, m), "centers must be [B,C,M]")
    with _Launch(features) as st:
        _check(_L.bdm_grouping_into(b, c, n, m, u, features.data_ptr(), indices.data_ptr(),
                                    centers.data_ptr() if centers is not None else None, out.data_ptr(),
                                    out.shape[1], channel_offset, st))
    return out


@_op(1)
def grouping_backward(grad_y, indices, n):
    _chk_float(grad_y, "grad_y")
    _chk_int(indices, "indices")
    b, c = grad_y.shape[0], grad_y.shape[1]
    m, u = indices.shape[1], indices.shape[2]
    n = int(n)
    grad_x = torch.empty((b, c, n), dtype=_F32, device=grad_y.device)
    with _Launch(grad_y) as st:
        _check(_L.bdm_grouping_grad(b, c, n, m, u, grad_y.data_ptr(), indices.data_ptr(), grad_x.data_ptr(), st))
    return grad_x


# ---------------------------------------------------------------------------------------------
# three nearest neighbours  (neighbor_interpolate.cpp:6-66)
# ---------------------------------------------------------------------------------------------
@_op(1)
def three_nn_search(points_coords, centers_coords):
    """-> (indices int32[B,3,N], weights f32[B,3,N]); the search half of the reference op."""
    _chk_float(points_coords, "points_coords")
    _chk_float(centers_coords, "centers_coords")
    b, n = points_coords.shape[0], points_coords.shape[2]
    m = centers_coords.shape[2]
    dev = points_coords.device
    indices = torch.empty((b, 3, n), dtype=_I32, device=dev)
    weights = torch.empty((b, 3, n), dtype=_F32, device=dev)
    with _Launch(points_coords) as st:
        _check(_L.bdm_three_nn_search(b, n, m, points_coords.data_ptr(), centers_coords.data_ptr(),
                                      weights.data_ptr(), indices.data_ptr(), st))
    return indices, weights


@_op(1)
def three_nn_interpolate(centers_features, indices, weights):
    _chk_float(centers_features, "centers_features")
    _chk_int(indices, "indices")
    _chk_float(weights, "weights")
    b, c, m = centers_features.shape
    n = indices.shape[2]
    out = torch.empty((b, c, n), dtype=_F32, device=centers_features.device)
    with _Launch(centers_features) as st:
        _check(_L.bdm_three_nn_interpolate(b, c, m, n, centers_features.data_ptr(), indices.data_ptr(),
                                           weights.data_ptr(), out.data_ptr(), st))
    return out


def three_nearest_neighbors_interpolate_forward(points_coords, centers_coords, centers_features):
    _chk_float(centers_features, "centers_features")
    indices, weights = three_nn_search(points_coords, centers_coords)
    return [three_nn_interpolate(centers_features, indices, weights), indices, weights]


@_op(1)
def three_nearest_neighbors_interpolate_backward(grad_y, indices, weights, m):
    _chk_float(grad_y, "grad_y")
    _chk_int(indices, "indices")
    _chk_float(weights, "weights")
    b, c, n = grad_y.shape
    m = int(m)
    grad_x = torch.empty((b, c, m), dtype=_F32, device=grad_y.device)
    with _Launch(grad_y) as st:
        _check(_L.bdm_three_nearest_neighbors_interpolate_grad(b, c, n, m, grad_y.data_ptr(), indices.data_ptr(),
                                                               weights.data_ptr(), grad_x.data_ptr(), st))
    return grad_x


# ---------------------------------------------------------------------------------------------
# dense side: fused GroupNorm (+ Swish)
# ---------------------------------------------------------------------------------------------
@_op(0)
def groupnorm_act(x, num_groups, weight, bias, eps, swish=True, conv_bias=None, max_over_last=False,
                  channel_sums=False):
    """x f32[B,C,*] -> act(group_norm(x + conv_bias[c])), act = swish or identity.
    max_over_last: return the max over the last dim instead (shape x.shape[:-1]).
    channel_sums:  also return f32[B,C] sums of the output over the trailing dims (SE squeeze)."""
    _chk_float(x, "x")
    b, c = x.shape[0], x.shape[1]
    for t_, nm in ((weight, "weight"), (bias, "bias"), (conv_bias, "conv_bias")):
        _chk_channel_vector(t_, nm, c)
    s = x.numel() // max(b * c, 1)
    dev = x.device
    u = int(x.shape[-1]) if max_over_last else 0
    y = torch.empty(x.shape[:-1] if max_over_last else x.shape, dtype=_F32, device=dev)
    sums = None
    if channel_sums:
        sums = torch.empty((b * c, _L.bdm_groupnorm_tiles(b, c, s)), dtype=_F32, device=dev)
    ws = _workspace(_L.bdm_groupnorm_workspace_bytes(b, c, s), dev)
    with _Launch(x) as st:
        _check(_L.bdm_groupnorm_act(b, c, s, int(num_groups), float(eps), 1 if swish else 0, u, x.data_ptr(),
                                    conv_bias.data_ptr() if conv_bias is not None else None,
                                    weight.data_ptr() if weight is not None else None,
                                    bias.data_ptr() if bias is not None else None, y.data_ptr(),
                                    sums.data_ptr() if sums is not None else None, ws.data_ptr(), ws.numel(), st))
    _count_launches(_L.bdm_groupnorm_last_launches())
    if channel_sums:
        return y, sums.sum(dim=1).view(b, c)
    return y


def groupnorm_cl_supported(channels, num_groups):
    return bool(_L.bdm_groupnorm_cl_supported(int(channels), int(num_groups)))


@_op(0)
def groupnorm_act_cl(x, num_groups, weight, bias, eps, swish=True, conv_bias=None, channel_sums=False, partials=None):
    """Channels-last flavour of groupnorm_act: x f32[B,*,C] contiguous (channel innermost) -> same layout;
    channel_sums: also f32[B,C] sums of the output over the voxels.
    partials: f64[B,chunks,C,2] per-channel (sum, sumsq) of x over disjoint voxel blocks made by x's producer
    (sparse_conv3_gather(stats=True)): the statistics pass is skipped."""
    _chk_float(x, "x")
    b, c = x.shape[0], x.shape[-1]
    for t_, nm in ((weight, "weight"), (bias, "bias"), (conv_bias, "conv_bias")):
        _chk_channel_vector(t_, nm, c)
    s = x.numel() // max(b * c, 1)
    dev = x.device
    y = torch.empty_like(x)
    sums = None
    if channel_sums:
        sums = torch.empty((b, _L.bdm_groupnorm_cl_tiles(b, c, s, int(num_groups)), c), dtype=_F32, device=dev)
    chunks = 0
    if partials is not None:
        ws, chunks = _partials_args(partials, b, c, int(num_groups), conv_bias)
        ws_bytes = ws.numel() * 8
    else:
        ws = _workspace(_L.bdm_groupnorm_cl_workspace_bytes(b, c, s), dev)
        ws_bytes = ws.numel()
    with _Launch(x) as st:
        _check(_L.bdm_groupnorm_act_cl(b, c, s, int(num_groups), float(eps), 1 if swish else 0, x.data_ptr(),
                                       conv_bias.data_ptr() if conv_bias is not None else None,
                                       weight.data_ptr() if weight is not None else None,
                                       bias.data_ptr() if bias is not None else None, y.data_ptr(),
                                       sums.data_ptr() if sums is not None else None, ws.data_ptr(), ws_bytes, chunks, st))
    _count_launches(_L.bdm_groupnorm_last_launches())
    if channel_sums == "tiles":     # f32[B,tiles,C], for se_gate (which folds the tiles itself)
        return y, sums
    if channel_sums:
        return y, sums.sum(dim=1)
    return y


@_op(1)
def groupnorm_cl_sums(x, num_groups, weight, bias, eps, swish, conv_bias, partials):
    """x f32[B,*,C] channels-last + producer statistics f64[B,chunks,C,2] -> (tile sums f32[B,tiles,C] of
    act(group_norm(x + conv_bias)) for se_gate, coefficients f32[B,C,2] = (A, B) with y = act(x*A + B)): the
    normalised tensor itself is not written (trilinear_devoxelize_cl(norm_coef=) applies it on the fly)."""
    _chk_float(x, "x")
    b, c = x.shape[0], x.shape[-1]
    for t_, nm in ((weight, "weight"), (bias, "bias"), (conv_bias, "conv_bias")):
        _chk_channel_vector(t_, nm, c)
    s = x.numel() // max(b * c, 1)
    pt, chunks = _partials_args(partials, b, c, int(num_groups), conv_bias)
    sums = torch.empty((b, _L.bdm_groupnorm_cl_sums_tiles(b, c, s), c), dtype=_F32, device=x.device)
    coef = torch.empty((b, c, 2), dtype=_F32, device=x.device)
    with _Launch(x) as st:
        _check(_L.bdm_groupnorm_cl_sums(b, c, s, int(num_groups), float(eps), 1 if swish else 0, x.data_ptr(),
                                        conv_bias.data_ptr() if conv_bias is not None else None,
                                        weight.data_ptr() if weight is not None else None,
                                        bias.data_ptr() if bias is not None else None, pt.data_ptr(),
                                        chunks, sums.data_ptr(), coef.data_ptr(), st))
    return sums, coef


def groupnorm_max_supported(u):
    return 4 <= u <= 128 and (u & (u - 1)) == 0


# ---------------------------------------------------------------------------------------------
# secondary boundaries
# ---------------------------------------------------------------------------------------------
@_op(4)
def surface_projection(points, R, T, focal, principal, feat, radius, feat_is_hwc=False):
    """points f32[B,N,3]; R [B,3,3]; T [B,3]; focal, principal [B,2]; feat [B,C,H,W] (or [B,H,W,C]
    when feat_is_hwc) -> (out f32[B,N,C], pix int32[B,N]: lowest won pixel index or -1)"""
    for x, nm in ((points, "points"), (R, "R"), (T, "T"), (focal, "focal"), (principal, "principal"),
                  (feat, "feat")):
        _chk_float(x, nm)
    b, n = points.shape[0], points.shape[1]
    if feat_is_hwc:
        H, W, C = feat.shape[1], feat.shape[2], feat.shape[3]
    else:
        C, H, W = feat.shape[1], feat.shape[2], feat.shape[3]
    dev = points.device
    zbuf = torch.empty((b, H, W), dtype=torch.int64, device=dev)
    pix = torch.empty((b, n), dtype=_I32, device=dev)
    out = torch.empty((b, n, C), dtype=_F32, device=dev)
    fn = _L.bdm_surface_projection_hwc if feat_is_hwc else _L.bdm_surface_projection
    with _Launch(points) as st:
        _check(fn(b, n, C, H, W, float(radius), points.data_ptr(), R.data_ptr(), T.data_ptr(), focal.data_ptr(),
                  principal.data_ptr(), feat.data_ptr(), zbuf.data_ptr(), pix.data_ptr(), out.data_ptr(), st))
    return out, pix


@_op(6)
def conditioning_input(points, R, T, focal, principal, feat_hwc, radius):
    """Fused get_input_with_conditioning: points f32[B,N,3] + feat_hwc f32[B,H,W,C] -> channel-first denoiser
    input f32[B,3+C,N] (channels 0-2 = the coordinates, 3.. = projected features), pix int32[B,N]."""
    for x, nm in ((points, "points"), (R, "R"), (T, "T"), (focal, "focal"), (principal, "principal"),
                  (feat_hwc, "feat_hwc")):
        _chk_float(x, nm)
    b, n = points.shape[0], points.shape[1]
    H, W, C = feat_hwc.shape[1], feat_hwc.shape[2], feat_hwc.shape[3]
    dev = points.device
    zbuf = torch.empty((b, H, W), dtype=torch.int64, device=dev)
    pix = torch.empty((b, n), dtype=_I32, device=dev)
    out = torch.empty((b, 3 + C, n), dtype=_F32, device=dev)
    out[:, :3, :].copy_(points.transpose(1, 2))
    with _Launch(points) as st:
        _check(_L.bdm_surface_projection_cf(b, n, C, H, W, float(radius), points.data_ptr(), R.data_ptr(),
                                            T.data_ptr(), focal.data_ptr(), principal.data_ptr(), feat_hwc.data_ptr(),
                                            zbuf.data_ptr(), pix.data_ptr(), out.data_ptr(), 3 + C, 3, st))
    return out, pix



# ---------------------------------------------------------------------------------------------
# reverse-diffusion update (model.py:182-194 via diffusers DDPMScheduler.step; pvd/__init__.py:136-224)
# ---------------------------------------------------------------------------------------------
@_op(1)
def sampler_update(x, eps, noise, table, t_dev, mode, out=None):
    """out = posterior sample from (x_t, predicted noise, fresh noise) with the coefficients of row *t_dev of
    `table` f32[T,8]; mode 0 = DDPM (PC^2 side), 1 = PVD.  x, eps, noise: same-shaped contiguous f32 tensors;
    t_dev int32[1] on the device (nothing is read back).  out may be x (in place)."""
    for t_, nm in ((x, "x"), (eps, "eps"), (noise, "noise"), (table, "table")):
        _chk_float(t_, nm)
    _chk_int(t_dev, "t_dev")
    _req(eps.shape == x.shape and noise.shape == x.shape, "x, eps and noise must have one shape")
    _req(table.dim() == 2 and table.shape[1] == 8, "table must be f32[T,8]")
    if out is None:
        out = torch.empty_like(x)
    else:
        _chk_float(out, "out")
        _req(out.shape == x.shape, "out must be shaped like x")
    with _Launch(x) as st:
        _check(_L.bdm_sampler_update(x.numel(), int(mode), x.data_ptr(), eps.data_ptr(), noise.data_ptr(),
                                     table.data_ptr(), table.shape[0], t_dev.data_ptr(), out.data_ptr(), st))
    return out


def nn_f64(src, tgt, expanded=False, return_index=True):
    """src f64[B,N,3], tgt f64[B,M,3] -> (min squared distance f64[B,N], argmin int32[B,N] | None)"""
    for x, nm in ((src, "src"), (tgt, "tgt")):
        _req(x.is_cuda, f"{nm} must be a CUDA tensor")
        _req(x.is_contiguous(), f"{nm} must be a contiguous tensor")
        _req(x.dtype == torch.float64, f"{nm} must be a double tensor")
    b, n = src.shape[0], src.shape[1]
    m = tgt.shape[1]
    dist = torch.empty((b, n), dtype=torch.float64, device=src.device)
    idx = torch.empty((b, n), dtype=_I32, device=src.device) if return_index else None
    with _Launch(src) as st:
        _check(_L.bdm_nn_f64(b, n, m, 1 if expanded else 0, src.data_ptr(), tgt.data_ptr(), dist.data_ptr(),
                             idx.data_ptr() if idx is not None else None, st))
    return dist, idx


@_op(1)
def nn_f64_reduce(src, tgt, expanded=False, thr=0.01):
    """src f64[B,N,3], tgt f64[B,M,3] -> (sum over i of min_j |s_i - t_j|^2  f64[B],  #{i: that minimum < thr} int64[B]):
    the nearest-neighbour search with the Chamfer / F-score reductions fused in (per-block partials, summed here)"""
    for x, nm in ((src, "src"), (tgt, "tgt")):
        _req(x.is_cuda, f"{nm} must be a CUDA tensor")
        _req(x.is_contiguous(), f"{nm} must be a contiguous tensor")
        _req(x.dtype == torch.float64, f"{nm} must be a double tensor")
    b, n = src.shape[0], src.shape[1]
    m = tgt.shape[1]
    blocks = max(int(_L.bdm_nn_f64_reduce_blocks(b, n)), 1)
    psum = torch.zeros((b, blocks), dtype=torch.float64, device=src.device)
    pcnt = torch.zeros((b, blocks), dtype=torch.int32, device=src.device)
    with _Launch(src) as st:
        _check(_L.bdm_nn_f64_reduce(b, n, m, 1 if expanded else 0, float(thr), src.data_ptr(), tgt.data_ptr(),
                                    psum.data_ptr(), pcnt.data_ptr(), st))
    return psum.sum(dim=1), pcnt.sum(dim=1, dtype=torch.int64)


# ---------------------------------------------------------------------------------------------
# tcgen05 3x3x3 convolution of the voxel branch (csrc/conv3_tc05.cu; reference: modules/pvconv.py:75-88)
# ---------------------------------------------------------------------------------------------
def conv3_tc05_supported(c_in, c_out, resolution):
    return bool(_L.bdm_conv3_tc05_supported(int(c_in), int(c_out), int(resolution)))


class GroupStats:
    """Group-level producer statistics f64[B,blocks,groups,2] -- (sum, sum of squares) per normalisation group over
    disjoint blocks of voxels, of the tensor as it is (bias included) -- as conv3_tc05(stats="groups") leaves them per
    unit; groupnorm_act_cl / groupnorm_cl_sums / groupnorm_swish_half_planar fold them in their prologue."""
    __slots__ = ("data",)

    def __init__(self, data):
        self.data = data


def _partials_args(partials, b, c, groups, conv_bias):
    """-> (tensor, chunks argument of the C entry): per-channel partials f64[B,chunks,C,2] -> +chunks; GroupStats -> -blocks"""
    if isinstance(partials, GroupStats):
        t = partials.data
        _req(t.dtype == torch.float64 and t.is_contiguous() and t.dim() == 4 and t.shape[0] == b and t.shape[2] == groups
             and t.shape[3] == 2, "group statistics must be f64[B,blocks,groups,2]")
        _req(conv_bias is None, "group statistics already include the bias")
        return t, -int(t.shape[1])
    _req(partials.dtype == torch.float64 and partials.is_contiguous() and partials.dim() == 4 and partials.shape[0] == b
         and partials.shape[2] == c and partials.shape[3] == 2, "partials must be f64[B,chunks,C,2]")
    return partials, int(partials.shape[1])


class HalfPlanes:
    """fp16 chunk planes [C/8, rows, 8] of the flat padded grid for (batch, resolution): the convolution's operand.
    Zero-filled at creation; producers write real voxels only, so pad rows stay zero for the buffer's lifetime."""
    __slots__ = ("b", "c", "r", "rows", "data", "occ")

    def __init__(self, b, c, r, device):
        self.b, self.c, self.r = int(b), int(c), int(r)
        self.rows = int(_L.bdm_conv3_tc05_plane_rows(self.b, self.r))
        self.data = torch.zeros((self.c // 8, self.rows, 8), dtype=torch.float16, device=device)
        # one bit per non-zero row (conv3_tc05_fill_planes): lets conv3_tc05(sparse=True) skip all-zero tap windows
        self.occ = torch.zeros((self.b, int(_L.bdm_conv3_tc05_occ_words(self.r))), dtype=torch.int32, device=device)

    def describe(self):
        return f"HalfPlanes(b={self.b}, c={self.c}, r={self.r})"


@_op(2)
def conv3_tc05_prepare(weight, gamma, beta, group_elems):
    """Conv3d weight f32[Cout,Cin,3,3,3] (+ the affine of the GroupNorm that produces the conv's input and the
    element count of one of its groups) -> the prepared weight buffer (scales + fp16 stage images)."""
    _chk_float(weight, "weight")
    c_out, c_in = weight.shape[:2]
    _req(tuple(weight.shape[2:]) == (3, 3, 3), "weight must be [Cout,Cin,3,3,3]")
    for t_, nm in ((gamma, "gamma"), (beta, "beta")):
        _chk_channel_vector(t_, nm, c_in)
    nbytes = int(_L.bdm_conv3_tc05_weight_bytes(c_in, c_out))
    prepared = torch.empty((nbytes,), dtype=torch.uint8, device=weight.device)
    with _Launch(weight) as st:
        _check(_L.bdm_conv3_tc05_prepare(c_in, c_out, weight.data_ptr(), gamma.data_ptr() if gamma is not None else None,
                                         beta.data_ptr() if beta is not None else None, int(group_elems),
                                         prepared.data_ptr(), nbytes, st))
    return prepared


@_op(1)
def groupnorm_swish_half_planar(x, num_groups, weight, bias, eps, swish, conv_bias, partials, prepared, planes):
    """x f32[B,R,R,R,C] channels-last + producer statistics f64[B,chunks,C,2] -> act(group_norm(x + conv_bias)) written
    into `planes` (HalfPlanes) in the scaling `prepared` prescribes."""
    _chk_float(x, "x")
    b, r, c = x.shape[0], x.shape[1], x.shape[-1]
    _req(x.dim() == 5 and x.shape[2] == r and x.shape[3] == r, "x must be [B,R,R,R,C]")
    _req(planes.b == b and planes.c == c and planes.r == r and planes.data.device == x.device, "planes do not match x")
    for t_, nm in ((weight, "weight"), (bias, "bias"), (conv_bias, "conv_bias")):
        _chk_channel_vector(t_, nm, c)
    pt, chunks = _partials_args(partials, b, c, int(num_groups), conv_bias)
    with _Launch(x) as st:
        _check(_L.bdm_groupnorm_swish_half_planar(b, c, r, int(num_groups), float(eps), 1 if swish else 0, x.data_ptr(),
                                                  conv_bias.data_ptr() if conv_bias is not None else None,
                                                  weight.data_ptr() if weight is not None else None,
                                                  bias.data_ptr() if bias is not None else None, pt.data_ptr(),
                                                  chunks, prepared.data_ptr(), planes.data.data_ptr(),
                                                  planes.rows, st))
    return planes


@_op(1)
def conv3_tc05_fill_planes(compact, plan, prepared, planes, amax_ready=False):
    """per-occupied-voxel averages f32[B,C,N] (avg_voxelize_compact) + VoxelPlan -> `planes` (zeros at empty voxels),
    scaled by a power of two derived on the device from max|average| and recorded in `prepared`.
    amax_ready: avg_voxelize_compact(..., amax_into=prepared) already measured max|average|."""
    if not amax_ready:
        _count_launches(1)
    _chk_float(compact, "compact")
    b, c, n = compact.shape
    _req(b == plan.b and n == plan.n, "compact does not match the voxel plan")
    _req(planes.b == b and planes.c == c and planes.r == plan.r and planes.data.device == compact.device, "planes do not match")
    with _Launch(compact) as st:
        _check(_L.bdm_conv3_tc05_fill_planes(b, c, n, plan.r, compact.data_ptr(), plan.workspace.data_ptr(), plan.workspace.numel(),
                                             prepared.data_ptr(), planes.data.data_ptr(), planes.rows,
                                             1 if amax_ready else 0, planes.occ.data_ptr(), st))
    return planes


@_op(1)
def conv3_tc05(planes, prepared, c_out, bias=None, stats=False, sparse=False):
    """HalfPlanes + prepared weights -> f32[B,R,R,R,Cout] channels-last (= conv + bias) and, with stats, the result's
    GroupNorm(8) statistics f64[B,1,Cout,2] for groupnorm_act_cl(partials=).
    sparse: the planes were just written by conv3_tc05_fill_planes (whose occupancy bits are current): tap windows that
    hold only zero rows are skipped.  Exact: the skipped products are zeros."""
    b, r, c_in = planes.b, planes.r, planes.c
    dev = planes.data.device
    _chk_channel_vector(bias, "bias", c_out)
    out = torch.empty((b, r, r, r, c_out), dtype=_F32, device=dev)
    part = ws = None
    ws_bytes = 0
    if stats == "groups":       # the per-unit group partials as they are: no folding kernel (see GroupStats)
        part = torch.empty((b, int(_L.bdm_conv3_tc05_units(c_in, int(c_out), r)), 8, 2), dtype=torch.float64, device=dev)
    elif stats:
        part = torch.empty((b, 1, c_out, 2), dtype=torch.float64, device=dev)
        ws_bytes = int(_L.bdm_conv3_tc05_workspace_bytes(b, r))
        ws = _workspace(ws_bytes, dev)
        _count_launches(1)
    with _Launch(planes.data) as st:
        _check(_L.bdm_conv3_tc05(b, c_in, int(c_out), r, planes.data.data_ptr(), planes.rows, prepared.data_ptr(),
                                 bias.data_ptr() if bias is not None else None, out.data_ptr(),
                                 part.data_ptr() if part is not None else None, ws.data_ptr() if ws is not None else None,
                                 ws_bytes, planes.occ.data_ptr() if sparse else None, st))
    if stats == "groups":
        return out, GroupStats(part)
    return (out, part) if stats else out
